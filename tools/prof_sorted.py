"""Small driver for ncu: C3-shaped ordered evaluation (2^26 queries) + prefilter, few launches."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200_loader
m = b200_loader.load()
n = 512
nq = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 26)
A = torch.randn((n, n, n), device="cuda", dtype=torch.float32)
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
g = torch.Generator(device="cuda").manual_seed(2)
coords = [(torch.rand(nq, device="cuda", dtype=torch.float32, generator=g) * (n - 1) + 1).clamp(1, n) for _ in range(3)]
for _ in range(3):
    out = m.evaluate_sorted(itp, *coords)
torch.cuda.synchronize()
print("ok", float(out.sum()))

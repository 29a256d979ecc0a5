"""Small driver for ncu / timing: C3-shaped ordered evaluation (default 2^27 queries), checked against the
direct kernel on a prefix."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200_loader
m = b200_loader.load()
n = 512
nq = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 27)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
A = torch.randn((n, n, n), device="cuda", dtype=torch.float32)
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
g = torch.Generator(device="cuda").manual_seed(2)
coords = [(torch.rand(nq, device="cuda", dtype=torch.float32, generator=g) * (n - 1) + 1).clamp(1, n) for _ in range(3)]
out = m.evaluate_sorted(itp, *coords)
torch.cuda.synchronize()
k = min(nq, 1 << 22)
ref = m.evaluate_direct(itp, *[c[:k] for c in coords])
assert torch.equal(out[:k], ref), "ordered path differs from the direct kernel"
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = m.evaluate_sorted(itp, *coords)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"ok nq={nq} ms={min(ts):.3f} gevals={nq / min(ts) / 1e6:.2f} checksum={float(out.double().sum()):.6f}")

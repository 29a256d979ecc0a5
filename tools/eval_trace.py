"""Per-CTA phase cycles of ord_eval_kernel (trace build: B200ITP_LIB=tools/libb200itp_trace.so)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_loader
m = b200_loader.load()
dev = torch.device("cuda", 0)
n = 512
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
g = torch.Generator(device=dev).manual_seed(1)
A = torch.rand((n, n, n), device=dev, dtype=torch.float32, generator=g)
itp = m.interpolate(A, it, device=dev)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
coords = [torch.rand(nq, device=dev, dtype=torch.float32, generator=g) * (n - 1) + 1 for _ in range(3)]
out = itp(*coords)
torch.cuda.synchronize()

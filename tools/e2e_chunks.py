"""e2e (host buffers in and out) throughput of evaluate_from_host as a function of the chunk size."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import b200_loader
m = b200_loader.load()
n, nq = 512, int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
itp = m.interpolate(torch.randn((n, n, n), device="cuda"), it)
g = torch.Generator().manual_seed(1)
xs = [(torch.rand(nq, generator=g) * (n - 1) + 1).pin_memory() for _ in range(3)]
out = torch.empty(nq, dtype=torch.float32).pin_memory()
for cs in (1, 3):
    for chunk in (1 << 26, 1 << 24):
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            m.evaluate_from_host(itp, *xs, out=out, chunk=chunk, copy_streams=cs)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"copy_streams {cs} chunk 2^{chunk.bit_length() - 1}: {dt * 1e3:.1f} ms, {nq / dt / 1e9:.3f} Gevals/s ({12 * nq / dt / 1e9:.1f} GB/s H2D)", flush=True)
# raw link rate: one big pinned -> device copy
d = torch.empty(nq, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(xs[0], non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"plain H2D copy of {4 * nq / 1e9:.1f} GB: {4 * nq / dt / 1e9:.1f} GB/s")

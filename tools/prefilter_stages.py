"""Per-stage device times of interpolate(A, it) for the BASELINE prefilter shapes (C1, C2, C4, C3 f32/f64)."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
g = torch.Generator(device="cuda").manual_seed(1)
def colmajor(shape, dt):
    t = torch.rand(tuple(reversed(shape)), device="cuda", dtype=dt, generator=g)
    return t.permute(*reversed(range(len(shape)))) if len(shape) > 1 else t
cases = [("C1", (10**6,), torch.float64, m.BSpline(m.Cubic(m.Line(m.OnGrid())))),
         ("C2", (4096, 4096), torch.float32, m.BSpline(m.Cubic(m.Flat(m.OnCell())))),
         ("C4", (8192, 8192), torch.float32, m.BSpline(m.Quadratic(m.Reflect(m.OnCell())))),
         ("C4f64", (8192, 8192), torch.float64, m.BSpline(m.Quadratic(m.Reflect(m.OnCell())))),
         ("C3", (512, 512, 512), torch.float32, m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))),
         ("C3f64", (512, 512, 512), torch.float64, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))]
flush = torch.zeros(64 << 20, device="cuda")
for name, shape, dt, it in cases:
    A = colmajor(shape, dt)
    for _ in range(3):
        itp = m.interpolate(A, it)
    ts = []
    for _ in range(5):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); itp = m.interpolate(A, it); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    flush.add_(1.0)
    L.b200itp_profile_enable(1); itp = m.interpolate(A, it); torch.cuda.synchronize(); L.b200itp_profile_enable(0)
    st = {k: round(min(v), 4) for k, v in m._lib.profile_stages().items()}
    import numpy as np
    es = 4 if dt == torch.float32 else 8
    nfilt = len(shape)
    alg = 2 * es * int(np.prod(itp.size)) * nfilt
    print(f"{name:6s} whole call {min(ts):.3f} ms ({alg / min(ts) / 1e6:.0f} GB/s algorithmic), stages {st}, sum {sum(st.values()):.3f}", flush=True)
    del A, itp

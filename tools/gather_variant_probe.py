"""Direct-kernel throughput of the BASELINE shapes for one build of the library (B200ITP_LIB selects the build)."""
import sys, os
sys.path.insert(0, ".")
import numpy as np
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
g = torch.Generator(device="cuda").manual_seed(1)
def colmajor(shape, dt):
    t = torch.rand(tuple(reversed(shape)), device="cuda", dtype=dt, generator=g)
    return t.permute(*reversed(range(len(shape)))) if len(shape) > 1 else t
flush = torch.zeros(64 << 20, device="cuda")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
def U(n, lo, hi, dt): return (torch.rand(n, device="cuda", dtype=dt, generator=g) * (hi - lo) + lo).clamp_(lo, hi)
itp1 = m.interpolate(colmajor((10**6,), torch.float64), m.BSpline(m.Cubic(m.Line(m.OnGrid())))); q1 = [U(10**7, 1.0, 1e6, torch.float64)]
itp5 = m.interpolate(colmajor((128,) * 4, torch.float64), m.BSpline(m.Linear())); q5 = [U(1 << 27, 1.0, 128.0, torch.float64) for _ in range(4)]
i4 = m.interpolate(colmajor((8192, 8192), torch.float32), m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))))
r = m.jrange(np.float32(0), np.float32(1), length=8192, dtype=np.float32)
itp4 = m.extrapolate(m.scale(i4, r, r), m.Line()); q4 = [U(10**8, -0.1, 1.1, torch.float32) for _ in range(2)]
itp2 = m.interpolate(colmajor((4096, 4096), torch.float32), m.BSpline(m.Cubic(m.Flat(m.OnCell())))); q2 = [U(10**8, 0.5, 4096.5, torch.float32) for _ in range(2)]
itp3 = m.interpolate(colmajor((512,) * 3, torch.float32), m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))); q3 = [U(1 << 26, 1.0, 512.0, torch.float32) for _ in range(3)]
t1 = timeit(lambda: itp1(*q1)); t5 = timeit(lambda: itp5(*q5)); t4 = timeit(lambda: itp4(*q4)); t4g = timeit(lambda: m.gradient(itp4, *q4))
t2 = timeit(lambda: itp2(*q2)); t2g = timeit(lambda: m.gradient(itp2, *q2)); t3 = timeit(lambda: m.evaluate_direct(itp3, *q3))
print(f"{os.environ.get('B200ITP_LIB', 'default')}: C1 {1e7 / t1 / 1e6:6.2f} | C2 {1e8 / t2 / 1e6:6.2f} grad {1e8 / t2g / 1e6:6.2f} | C3 direct {(1 << 26) / t3 / 1e6:5.2f} | "
      f"C4 {1e8 / t4 / 1e6:6.2f} grad {1e8 / t4g / 1e6:6.2f} | C5 {(1 << 27) / t5 / 1e6:6.2f} Gevals/s", flush=True)

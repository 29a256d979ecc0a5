"""Prefilter error table: fast kernels and the strict build against the CPU oracle (norm-wise and the element-wise
SURVEY §8d metric with its 1e-3 floor), per element type / shape / spec."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
import oracle as orc
rng = np.random.default_rng(0)
def errs(got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64)); r = np.abs(ref.astype(np.float64))
    return d.max() / r.max(), (d / np.maximum(r, 1e-3 * r.max())).max()
for dt in (np.float32, np.float64):
    for shape in ((200000,), (1500, 1300), (256, 128, 128), (160, 136, 130)):
        for it in (m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), m.BSpline(m.Cubic(m.Flat(m.OnCell()))), m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))), m.BSpline(m.Cubic(m.Line(m.OnGrid())))):
            A = rng.standard_normal(shape).astype(dt)
            ref = orc.prefilter(A, it, nthreads=8)
            fast = m.interpolate(A, it).coefs_array()
            L.b200itp_prefilter_set_strict(1)
            strict = m.interpolate(A, it).coefs_array()
            L.b200itp_prefilter_set_strict(0)
            nf, ef = errs(fast, ref); ns, es = errs(strict, ref)
            print(f"{np.dtype(dt).name} {str(shape):18s} {str(it):40s} fast norm {nf:.2e} elem {ef:.2e} | strict norm {ns:.2e} elem {es:.2e} equal {np.array_equal(strict, ref)}", flush=True)

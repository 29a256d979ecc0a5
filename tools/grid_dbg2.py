"""Debug driver of the separable grid evaluation against the CPU oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
import oracle as orc
rng = np.random.default_rng(99)
coef_dtype = coord_dtype = np.float32
shape = (33,)
A = rng.standard_normal(shape).astype(coef_dtype)
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
rngs = [m.jrange(np.float32(-1), np.float32(2), length=33, dtype=np.float32)]
for obj in (itp, m.scale(itp, *rngs), m.extrapolate(itp, m.Flat()), m.extrapolate(m.scale(itp, *rngs), m.Periodic())):
    (lb, ub), = obj.bounds()
    v = np.array([lb, lb + 0.3 * (ub - lb), ub], np.float32)
    got = m.evaluate_grid(obj, v).cpu().numpy()
    pt = m.evaluate_direct(obj, v).cpu().numpy()
    host = orc.with_coefs(obj, obj.coefficients().cpu().numpy())
    ov, _, bi, br, _ = orc.evaluate(host, v, debug=True)
    gv, _, gbi, gbr = m.evaluate_debug(obj, v)
    print(type(obj).__name__, repr(obj)[:120])
    print('  x', v, 'bounds', lb, ub)
    print('  grid', got, got.dtype, 'direct', pt, 'oracle', ov, 'oracle f32-repr', ov == ov.astype(np.float32))
    print('  base idx gpu', gbi.cpu().numpy().ravel(), 'orc', bi.ravel(), 'branch', gbr.cpu().numpy(), br)

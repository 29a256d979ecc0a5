"""Generates tests/golden/bspline_golden.npz from the CPU oracle (oracle/, pinned to the reference's known answers
and property tests by tests/test_oracle_pinning.py; the reference itself is Julia and cannot run in this image, and
it ships no golden vectors for this path — SURVEY §8c).  The fixture freezes the oracle's outputs so that a later
change of the oracle or of the CUDA path shows up as a diff against committed numbers.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import b200_loader  # noqa: E402

m = b200_loader.load()
import oracle as orc  # noqa: E402


def cases():
    G, Cc = m.OnGrid, m.OnCell
    yield "c1_cubic_line_ongrid_f64", (41,), np.float64, m.BSpline(m.Cubic(m.Line(G()))), None, None
    yield "c2_cubic_flat_oncell_f32", (19, 17), np.float32, m.BSpline(m.Cubic(m.Flat(Cc()))), None, None
    yield "c3_cubic_periodic_ongrid_f32", (12, 11, 10), np.float32, m.BSpline(m.Cubic(m.Periodic(G()))), None, None
    yield ("c4_quadratic_reflect_oncell_scaled_line_f32", (15, 14), np.float32,
           m.BSpline(m.Quadratic(m.Reflect(Cc()))), "scale", m.Line())
    yield "c5_linear_4d_f64", (5, 6, 4, 5), np.float64, m.BSpline(m.Linear()), None, None
    yield "quadratic_free_ongrid_f64", (23,), np.float64, m.BSpline(m.Quadratic(m.Free(G()))), None, m.Flat()
    yield "cubic_free_oncell_periodic_etp_f64", (16, 13), np.float64, m.BSpline(m.Cubic(m.Free(Cc()))), None, m.Periodic()
    yield "quadratic_periodic_oncell_reflect_etp_f32", (14, 12), np.float32, m.BSpline(m.Quadratic(m.Periodic(Cc()))), None, m.Reflect()


def build(name, shape, dtype, it, scale, etp):
    rng = np.random.default_rng(abs(hash(name)) % 2**32 if False else sum(map(ord, name)))
    A = rng.standard_normal(shape).astype(dtype)
    itp = orc.interpolate(A, it)
    obj = itp
    lo = [float(b[0]) for b in itp.bounds()]
    hi = [float(b[1]) for b in itp.bounds()]
    if scale:
        ranges = [m.jrange(np.float32(0), np.float32(1), length=n) if dtype == np.float32 else m.jrange(0.0, 1.0, length=n)
                  for n in shape]
        obj = m.scale(obj, *ranges)
        lo = [float(b[0]) for b in obj.bounds()]
        hi = [float(b[1]) for b in obj.bounds()]
    if etp is not None:
        obj = m.extrapolate(obj, etp)
    nq = 64
    qs = []
    for d in range(len(shape)):
        span = hi[d] - lo[d]
        margin = 0.35 * span if etp is not None else 0.0
        q = rng.uniform(lo[d] - margin, hi[d] + margin, nq)
        q[:4] = [lo[d], hi[d], lo[d] + 0.5 * span, lo[d] + 1e-3 * span]
        qs.append(q.astype(dtype))
    val, _, idx, br, _ = orc.evaluate(obj, *qs, debug=True)
    _, grad, _, _, _ = orc.evaluate(obj, *qs, want_value=False, want_grad=True)
    return A, itp.coefs, qs, val, grad, idx, br


def main():
    out = {}
    for name, shape, dtype, it, scale, etp in cases():
        A, coefs, qs, val, grad, idx, br = build(name, shape, dtype, it, scale, etp)
        out[name + "/A"] = A
        out[name + "/coefs"] = np.asarray(coefs)
        for d, q in enumerate(qs):
            out[f"{name}/q{d}"] = q
        out[name + "/value"] = val
        out[name + "/gradient"] = grad
        out[name + "/index"] = idx
        out[name + "/branch"] = br
    path = os.path.join(HERE, "bspline_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

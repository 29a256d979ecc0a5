"""Committed golden vectors (tests/golden/bspline_golden.npz, written by tests/golden/make_golden.py from the CPU
oracle): the oracle must still reproduce them (CPU), and the CUDA path must match them (GPU) — stencil indices and
branch records bit-exact, values within 1e-5 (Float32) / 1e-12 (Float64) of max(|ref|, 1e-3 ||ref||_inf)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "bspline_golden.npz")


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _close(got, ref, tol):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
    return np.all(np.abs(got - ref) <= tol * scale)


def _objects(m, mod, name, shape, dtype, it, scale, etp, interp):
    g = np.load(GOLD)
    A = g[name + "/A"]
    itp = interp(A, it)
    obj = itp
    if scale:
        ranges = [m.jrange(np.float32(0), np.float32(1), length=n) if dtype == np.float32 else m.jrange(0.0, 1.0, length=n)
                  for n in shape]
        obj = m.scale(obj, *ranges)
    if etp is not None:
        obj = m.extrapolate(obj, etp)
    qs = [g[f"{name}/q{d}"] for d in range(len(shape))]
    return g, itp, obj, qs


def test_oracle_reproduces_golden(pkg, orc):
    mod = _gen()
    for name, shape, dtype, it, scale, etp in mod.cases():
        g, itp, obj, qs = _objects(pkg, mod, name, shape, dtype, it, scale, etp, orc.interpolate)
        assert np.array_equal(np.asarray(itp.coefs), g[name + "/coefs"]), name
        val, _, idx, br, _ = orc.evaluate(obj, *qs, debug=True)
        _, grad, _, _, _ = orc.evaluate(obj, *qs, want_value=False, want_grad=True)
        assert np.array_equal(val, g[name + "/value"], equal_nan=True), name
        assert np.array_equal(grad, g[name + "/gradient"], equal_nan=True), name
        assert np.array_equal(idx, g[name + "/index"]) and np.array_equal(br, g[name + "/branch"]), name


@pytest.mark.gpu
def test_cuda_path_matches_golden(pkg):
    m = pkg
    mod = _gen()
    for name, shape, dtype, it, scale, etp in mod.cases():
        tol = 1e-5 if dtype == np.float32 else 1e-12
        g, itp, obj, qs = _objects(m, mod, name, shape, dtype, it, scale, etp, m.interpolate)
        ref_c = g[name + "/coefs"]
        got_c = itp.coefs.cpu().numpy()
        assert np.abs(got_c - ref_c).max() <= tol * np.abs(ref_c).max(), name
        # evaluate on the golden coefficients so that the comparison isolates the evaluation path
        import torch
        itp.coefs.copy_(torch.from_numpy(ref_c).to(itp.coefs.device))
        val, _, idx, br = m.evaluate_debug(obj, *qs)
        grad = m.gradient(obj, *qs)
        assert np.array_equal(idx.cpu().numpy(), g[name + "/index"]), name
        assert np.array_equal(br.cpu().numpy().astype(np.uint32), g[name + "/branch"].astype(np.uint32)), name
        assert _close(val.cpu().numpy(), g[name + "/value"], tol), name
        assert _close(grad.cpu().numpy(), g[name + "/gradient"], tol), name

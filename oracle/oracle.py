"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front end of the CPU oracle (oracle/oracle.c), a literal restatement of the reference's
B-spline hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  The package's pure-host classes (flags, ranges,
wrapper objects) are used as CONTAINERS of the specification only: the descriptor the C oracle
consumes (padding, first coefficient index, bounds, types) is derived independently in
oracle/describe.py from the reference's rules, not by the product's `descriptor()`; all arithmetic
is done by liboracle.so on the CPU.  Parity is unpinned at the last ulp: the
reference holds no golden vectors for this path and Julia is absent from the image.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import b200_loader  # noqa: E402

pkg = b200_loader.load()
from interpolations_jl_b200 import _lib as _pl  # noqa: E402  (struct layouts only)
from interpolations_jl_b200.core import BSplineInterpolation, _its_tuple  # noqa: E402
from interpolations_jl_b200.flags import F32, F64, needs_padding  # noqa: E402
from describe import describe  # noqa: E402  (the oracle's own descriptor derivation)

_libs = {}


def build(force=False):
    """Compile liboracle.so / liboracle_fast.so with the committed Makefile."""
    if force or not (os.path.exists(os.path.join(_HERE, "liboracle.so"))
                     and os.path.exists(os.path.join(_HERE, "liboracle_fast.so"))):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "all"])


def lib(fast=False) -> C.CDLL:
    key = "fast" if fast else "parity"
    if key in _libs:
        return _libs[key]
    build()
    L = C.CDLL(os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so"))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    pi32, pi64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.orc_num_threads.restype = C.c_int
    L.orc_prefiltering_system.restype = C.c_int
    L.orc_prefiltering_system.argtypes = [i32, i64, i32, i32, i32, vp, vp, vp, pi32, pi32, pi32, vp]
    L.orc_prefilter.restype = C.c_int
    L.orc_prefilter.argtypes = [i32, i32, pi64, vp, pi32, pi32, pi32, vp, i32]
    L.orc_prefilter_inplace.restype = C.c_int
    L.orc_prefilter_inplace.argtypes = [i32, i32, pi64, vp, pi32, pi32, pi32, i32]
    L.orc_eval.restype = C.c_int
    L.orc_eval.argtypes = [C.POINTER(_pl.Desc), vp, C.POINTER(vp), i64, vp, vp, pi64, vp, vp, i32]
    L.orc_weights.restype = C.c_int
    L.orc_weights.argtypes = [i32, i32, C.c_double, i64, C.POINTER(C.c_double), C.POINTER(C.c_double), pi32]
    L.orc_eval_tricubic_periodic_f32.restype = C.c_int
    L.orc_eval_tricubic_periodic_f32.argtypes = [vp, i64, i64, i64, vp, vp, vp, i64, vp, i32]
    _libs[key] = L
    return L


def num_threads() -> int:
    return lib().orc_num_threads()


def _codes(its):
    N = len(its)
    deg = (C.c_int32 * N)(*[it.degree.code for it in its])
    bc = (C.c_int32 * N)(*[it.degree.bc.code for it in its])
    gt = (C.c_int32 * N)(*[it.degree.bc.gt.code for it in its])
    return deg, bc, gt


def prefiltering_system(dtype, n, degree_code, bc_code, grid_code):
    """(dl, d, du, urow, vcol, Cp) as the reference's `prefiltering_system` + `lut!` + Woodbury give."""
    dt = np.dtype(dtype)
    code = F32 if dt == np.float32 else F64
    dl = np.zeros(max(n - 1, 1), dt)
    d = np.zeros(n, dt)
    du = np.zeros(max(n - 1, 1), dt)
    k = C.c_int32(0)
    urow = (C.c_int32 * 8)()
    vcol = (C.c_int32 * 8)()
    Cp = np.zeros(64, dt)
    rc = lib().orc_prefiltering_system(code, n, degree_code, bc_code, grid_code, dl.ctypes.data, d.ctypes.data,
                                       du.ctypes.data, C.byref(k), urow, vcol, Cp.ctypes.data)
    if rc:
        raise ValueError(f"orc_prefiltering_system rc={rc}")
    kk = k.value
    return dl[:n - 1], d, du[:n - 1], list(urow[:kk]), list(vcol[:kk]), Cp[:kk * kk].reshape(kk, kk, order="F")


def prefilter(A: np.ndarray, it, nthreads=1, fast=False) -> np.ndarray:
    """`prefilter(TW, TC, A, it)` (prefiltering.jl:32-38): returns the padded coefficient array with
    Julia's shape (Fortran-ordered numpy array)."""
    A = np.asarray(A)
    if A.dtype not in (np.float32, np.float64):
        A = A.astype(np.float64)
    N = A.ndim
    its = _its_tuple(it, N)
    pad = [1 if needs_padding(i) else 0 for i in its]
    psize = tuple(A.shape[d] + 2 * pad[d] for d in range(N))
    src = np.ascontiguousarray(A.reshape(-1, order="F"))
    out = np.empty(int(np.prod(psize)), A.dtype)
    deg, bc, gt = _codes(its)
    code = F32 if A.dtype == np.float32 else F64
    rc = lib(fast).orc_prefilter(code, N, (C.c_int64 * N)(*A.shape), src.ctypes.data, deg, bc, gt, out.ctypes.data,
                                 nthreads)
    if rc:
        raise ValueError(f"orc_prefilter rc={rc}")
    return out.reshape(psize, order="F")


def prefilter_inplace(coefs_flat: np.ndarray, size, it, nthreads=1, fast=False):
    """`prefilter!(TW, ret, it)` on an already padded column-major flat array."""
    N = len(size)
    its = _its_tuple(it, N)
    deg, bc, gt = _codes(its)
    code = F32 if coefs_flat.dtype == np.float32 else F64
    rc = lib(fast).orc_prefilter_inplace(code, N, (C.c_int64 * N)(*size), coefs_flat.ctypes.data, deg, bc, gt,
                                         nthreads)
    if rc:
        raise ValueError(f"orc_prefilter_inplace rc={rc}")
    return coefs_flat


def interpolate(A, it, nthreads=1) -> BSplineInterpolation:
    """`interpolate(A, it)` on the CPU; the result's `coefs` is a flat column-major numpy array."""
    A = np.asarray(A)
    if A.dtype not in (np.float32, np.float64):
        A = A.astype(np.float64)
    its = _its_tuple(it, A.ndim)
    c = prefilter(A, its, nthreads)
    return BSplineInterpolation(np.ascontiguousarray(c.reshape(-1, order="F")), c.shape, A.shape, its, None)


def interpolate_inplace(A, it, nthreads=1) -> BSplineInterpolation:
    """`interpolate!(A, it)` on the CPU (b-splines.jl:236-241): the system of size n is solved on the samples
    themselves and the axes a degree would have padded are trimmed to 2:n-1."""
    A = np.asarray(A)
    if A.dtype not in (np.float32, np.float64):
        A = A.astype(np.float64)
    its = _its_tuple(it, A.ndim)
    flat = np.ascontiguousarray(A.reshape(-1, order="F")).copy()
    prefilter_inplace(flat, A.shape, its, nthreads)
    inset = [1 if needs_padding(i) else 0 for i in its]
    return BSplineInterpolation(flat, A.shape, tuple(n - 2 * k for n, k in zip(A.shape, inset)), its, None,
                                parent_first=tuple(1 + k for k in inset))


def with_coefs(itp, coefs_flat: np.ndarray):
    """The same wrapper chain around other coefficients (e.g. the GPU's, copied to the host)."""
    from interpolations_jl_b200.core import Extrapolation, FilledExtrapolation, ScaledInterpolation
    b, s, e = itp._chain()
    nb = BSplineInterpolation(np.ascontiguousarray(coefs_flat), b.size, b.parent_len, b.its, None,
                              parent_first=b.parent_first)
    out = nb
    if s is not None:
        out = ScaledInterpolation(nb, s.ranges)
    if isinstance(e, Extrapolation):
        out = Extrapolation(out, e.flags)
    elif isinstance(e, FilledExtrapolation):
        out = FilledExtrapolation(out, e.fillvalue)
        out.fill_tag = e.fill_tag
    return out


def evaluate(itp, *xs, want_value=True, want_grad=False, debug=False, nthreads=1, fast=False, raise_oob=True):
    """itp.(xs...) / gradient.(Ref(itp), xs...) on the CPU.  Returns (value, grad, base_index, branch,
    first_oob); value/grad are float64 arrays holding numbers already rounded to their Julia type."""
    N = itp.ndims
    arrs = [np.asarray(x) for x in xs]
    dt = np.float64 if any(a.dtype == np.float64 or a.dtype.kind in "iu" for a in arrs) else np.float32
    arrs = np.broadcast_arrays(*[a.astype(dt) for a in arrs])
    shape = arrs[0].shape
    cs = [np.ascontiguousarray(a.reshape(-1)) for a in arrs]
    nq = cs[0].size
    D = describe(itp, dt)
    coefs = itp.coefficients()
    if not isinstance(coefs, np.ndarray):
        coefs = coefs.cpu().numpy()
    val = np.empty(nq, np.float64) if want_value else None
    G = sum(1 for d in range(N) if D.degree[d] != 7)  # NoInterp axes (code 7) have no gradient component
    grad = np.empty(nq * G, np.float64) if want_grad else None
    bi = np.empty(nq * N, np.int32) if debug else None
    br = np.empty(nq, np.uint32) if debug else None
    ptrs = (C.c_void_p * N)(*[c.ctypes.data for c in cs])
    foob = C.c_int64(-1)

    def p(a):
        return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)

    rc = lib(fast).orc_eval(C.byref(D), p(coefs), ptrs, nq, p(val), p(grad), C.byref(foob), p(bi), p(br), nthreads)
    if rc:
        raise ValueError(f"orc_eval rc={rc}")
    if raise_oob and foob.value >= 0:
        raise pkg.BoundsError(f"oracle: BoundsError at query {foob.value}")
    return (val.reshape(shape) if want_value else None, grad.reshape(shape + (G,)) if want_grad else None,
            bi.reshape(shape + (N,)) if debug else None, br.reshape(shape) if debug else None, foob.value)


def evaluate_tricubic_periodic_f32(coefs_flat, shape, xs, ys, zs, nthreads=1, fast=True):
    """Type-specialised `itp.(xs, ys, zs)` for a Float32 Cubic(Periodic(OnGrid())) volume (what Julia compiles for this
    interpolant; bench.py's CPU legs).  In-bounds coordinates only; same bits as `evaluate`."""
    n0, n1, n2 = (int(v) for v in shape)
    xs, ys, zs = (np.ascontiguousarray(a, np.float32) for a in (xs, ys, zs))
    out = np.empty(xs.size, np.float32)
    rc = lib(fast).orc_eval_tricubic_periodic_f32(coefs_flat.ctypes.data, n0, n1, n2, xs.ctypes.data, ys.ctypes.data,
                                                  zs.ctypes.data, xs.size, out.ctypes.data, nthreads)
    if rc:
        raise ValueError(f"orc_eval_tricubic_periodic_f32 rc={rc}")
    return out


def hessian(itp, *xs, nthreads=1, fast=False, raise_oob=True):
    """hessian.(Ref(itp), xs...) on the CPU: array of shape xs.shape + (N, N) (float64 holding numbers already
    rounded to their Julia type)."""
    N = itp.ndims
    arrs = [np.asarray(x) for x in xs]
    dt = np.float64 if any(a.dtype == np.float64 or a.dtype.kind in "iu" for a in arrs) else np.float32
    arrs = np.broadcast_arrays(*[a.astype(dt) for a in arrs])
    shape = arrs[0].shape
    cs = [np.ascontiguousarray(a.reshape(-1)) for a in arrs]
    nq = cs[0].size
    D = describe(itp, dt)
    if D.etp_kind == 1:
        D.etp_kind = 0  # hessian(etp, x) = hessian(parent(etp), x) in bounds, error outside (extrapolation.jl:84-90)
    coefs = itp.coefficients()
    if not isinstance(coefs, np.ndarray):
        coefs = coefs.cpu().numpy()
    out = np.empty(nq * N * N, np.float64)
    ptrs = (C.c_void_p * N)(*[c.ctypes.data for c in cs])
    foob = C.c_int64(-1)
    fn = lib(fast).orc_hessian
    fn.restype = C.c_int
    rc = fn(C.byref(D), C.c_void_p(coefs.ctypes.data), ptrs, C.c_int64(nq), C.c_void_p(out.ctypes.data), C.byref(foob),
            C.c_int(nthreads))
    if rc:
        raise ValueError(f"orc_hessian rc={rc}")
    if raise_oob and foob.value >= 0:
        raise pkg.BoundsError(f"oracle: BoundsError at query {foob.value}")
    return out.reshape(shape + (N, N))


def weights(degree_code, x, n, dtype=np.float64):
    """(value_weights, gradient_weights, first tap index) at coordinate x on an axis 1:n."""
    wv = (C.c_double * 4)()
    wg = (C.c_double * 4)()
    base = C.c_int32(0)
    tag = 1 if np.dtype(dtype) == np.float32 else 2
    L = lib().orc_weights(degree_code, tag, float(x), n, wv, wg, C.byref(base))
    return list(wv[:L]), list(wg[:L]), base.value

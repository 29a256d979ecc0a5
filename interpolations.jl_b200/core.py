"""Host-side mirror of the reference's operator interface for the B-spline hot path.

Same names, argument meaning and error behaviour as Interpolations.jl v0.16.3:
`interpolate`, `scale`, `extrapolate`, broadcast call `itp(xs...)`, `gradient`,
`bounds/lbounds/ubounds`.  Everything numeric runs in libb200itp.so (hand-written
sm_100a CUDA) through the C ABI of include/b200itp.h; this file only flattens the
wrapper chain into the POD descriptor, exactly what the Julia glue does
(INTEGRATION.md), and owns the device buffers (torch tensors).

Reference: src/b-splines/b-splines.jl:74-78,141-158,167-193; src/scaling/scaling.jl:5-8,
33-48,60-83,116-141; src/extrapolation/extrapolation.jl:1-4,45-58,71-82; filled.jl:1-40;
src/gpu_support.jl:52-80 (the broadcast seam).

Arrays follow Julia's convention: `A[i1, i2, ...]` with the FIRST index contiguous in
device memory.  A numpy/torch array of shape (n1, ..., nN) is interpreted with that
index order and stored column-major.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from .flags import (BC_FLAT, BC_LINE, BC_PERIODIC, BC_REFLECT, BC_THROW, CUBIC, ETP_FILL, ETP_FLAGS, ETP_NONE,
                    CONSTANT_NEAREST, NOINTERP, F32, F64, LINEAR, ONCELL, ONGRID, QUADRATIC, TAG_EXACT, TAG_F32, TAG_F64, BoundaryCondition,
                    BSpline, Flat, Free, Line, OnCell, OnGrid, Periodic, Reflect, Throw, iscomplete, needs_padding)
from .ranges import AbstractRange, UnitRange

MAX_DIMS = _lib.MAX_DIMS


class BoundsError(IndexError):
    """Raised where the reference throws BoundsError (indexing.jl:6, extrapolation.jl:115-129)."""


def _torch():
    import torch
    return torch


def _np_dtype(code):
    return np.float32 if code == F32 else np.float64


def _code_of(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.dtype(np.float32):
        return F32
    if dt == np.dtype(np.float64):
        return F64
    raise TypeError(f"only Float32/Float64 reach the GPU kernels, got {dt}")


def _tag_of_code(code):
    return TAG_F32 if code == F32 else TAG_F64


def _its_tuple(it, ndims) -> Tuple[BSpline, ...]:
    if isinstance(it, BSpline):
        return (it,) * ndims
    its = tuple(it)
    if len(its) != ndims or not all(isinstance(i, BSpline) for i in its):
        raise TypeError("interpolation type must be a BSpline or one BSpline per dimension")
    return its


def stage_tags(D):
    """Julia type (TAG_F32/TAG_F64) of each axis' coordinate after every wrapper, following the
    promotion rules of `maxp/minp` (extrapolation.jl:153-156), `coordlookup` (scaling.jl:102-115) and
    `clamp` (Interpolations.jl:450-452).  Returns (t_after_extrapolation, t_weights) per axis."""
    t0 = TAG_F32 if D.coord_dtype == F32 else TAG_F64
    t1, tw = [], []
    for d in range(D.ndims):
        t = t0
        if D.etp_kind == ETP_FLAGS and not (D.etp_lo[d] == BC_THROW and D.etp_hi[d] == BC_THROW):
            t = max(t, D.outer_tag[d] if D.has_scale else D.inner_tag[d])
        t1.append(t)
        if D.has_scale:
            t = max(t, D.range_tag[d], D.inner_tag[d])
        tw.append(t)
    return t1, tw


def out_dtype_code(D) -> int:
    """promote(weights type of every axis, eltype(coefs) [, fill value]) -> F32/F64; mirrors
    b200itp_out_dtype."""
    _, tw = stage_tags(D)
    # Constant axes carry Int weights (constant.jl:106): they do not widen the result
    tw = [t for d, t in enumerate(tw) if not (CONSTANT_NEAREST <= D.degree[d] <= NOINTERP)]
    t = max(tw + [TAG_F32 if D.coef_dtype == F32 else TAG_F64])
    if D.etp_kind == ETP_FILL:
        t = max(t, D.fill_tag)
    return F32 if t == TAG_F32 else F64


# ----------------------------------------------------------------------------- interpolants
class AbstractInterpolation:
    ndims: int

    # -- pure host logic -------------------------------------------------------
    def lbounds(self):
        raise NotImplementedError

    def ubounds(self):
        raise NotImplementedError

    def bounds(self):
        """`bounds(itp)`: ((l1,u1), ..., (lN,uN))   (Interpolations.jl:155)."""
        return tuple(zip(self.lbounds(), self.ubounds()))

    def _chain(self):
        """(bspline, scaled-or-None, extrapolation-or-None); only this wrapper order is on the hot path."""
        raise NotImplementedError

    def descriptor(self, coord_dtype) -> _lib.Desc:
        bsp, sc, et = self._chain()
        D = _lib.Desc()
        N = bsp.ndims
        D.ndims = N
        D.coef_dtype = bsp.dtype_code
        D.coord_dtype = _code_of(coord_dtype)
        D.etp_pair_mask = 0
        for d in range(N):
            it = bsp.its[d]
            D.size[d] = bsp.size[d]
            D.parent_len[d] = bsp.parent_len[d]
            D.coef_first[d] = 0 if (needs_padding(it) and bsp.parent_first[d] == 1) else 1
            D.parent_first[d] = bsp.parent_first[d]
            D.degree[d] = it.degree.code
            D.bc[d] = it.degree.bc.code
            D.gridtype[d] = it.degree.bc.gt.code
            lo, hi, tag = bsp._bound(d)
            D.inner_lb[d], D.inner_ub[d], D.inner_tag[d] = lo, hi, tag
        D.has_scale = 1 if sc is not None else 0
        if sc is not None:
            for d in range(N):
                r = sc.ranges[d]
                D.range_tag[d] = r.tag
                D.range_first[d] = float(r.first)
                D.range_step[d] = float(r.step)
                lo, hi, tag = sc._bound(d)
                D.outer_lb[d], D.outer_ub[d], D.outer_tag[d] = lo, hi, tag
        D.etp_kind = ETP_NONE
        if isinstance(et, Extrapolation):
            D.etp_kind = ETP_FLAGS
            for d in range(N):
                f = et.flags[d]
                if isinstance(f, tuple):
                    D.etp_pair_mask |= 1 << d
                    D.etp_lo[d], D.etp_hi[d] = f[0].code, f[1].code
                else:
                    D.etp_lo[d] = D.etp_hi[d] = f.code
        elif isinstance(et, FilledExtrapolation):
            D.etp_kind = ETP_FILL
            D.fill_value = float(et.fillvalue)
            D.fill_tag = et.fill_tag
        return D

    # -- evaluation through the C ABI -----------------------------------------
    def __call__(self, *xs):
        """Broadcast evaluation `itp.(xs...)` (gpu_support.jl:52-57).  Large scattered broadcasts over arrays that do
        not fit the L2 take the ordered pipeline (`b200itp_eval_sorted`), everything else the direct kernel; the
        results are bit-identical."""
        return _broadcast_eval(self, xs, want_value=True, want_grad=False, sorted_queries=True)[0]

    def coefficients(self):
        return self._chain()[0].coefs


class BSplineInterpolation(AbstractInterpolation):
    """Fields as in b-splines.jl:74-78: `coefs` (device, column-major, padded), `parentaxes`, `it`."""

    def __init__(self, coefs, size, parent_len, its, device_index, parent_first=None):
        self.coefs = coefs  # 1-D torch CUDA tensor (column-major linearisation)
        self.size = tuple(int(s) for s in size)
        self.parent_len = tuple(int(s) for s in parent_len)
        self.its = tuple(its)
        self.ndims = len(self.size)
        # first.(parentaxes): 1, or 2 on the axes `interpolate!` insets (b-splines.jl:236-241)
        self.parent_first = tuple(int(f) for f in parent_first) if parent_first is not None else (1,) * self.ndims
        self.device_index = device_index
        self.dtype_code = F32 if str(coefs.dtype).endswith("float32") else F64  # torch or numpy dtype

    @property
    def parentaxes(self):
        return tuple(UnitRange(f, f + n - 1) for f, n in zip(self.parent_first, self.parent_len))

    def _bound(self, d):
        """lbound/ubound (b-splines.jl:141-158): OnGrid -> (first, last) Ints; OnCell -> first-0.5, last+0.5 Float64."""
        n = self.parent_len[d]
        f = self.parent_first[d]
        gt = self.its[d].degree.bc.gt
        if isinstance(gt, OnCell):
            return f - 0.5, f + n - 1 + 0.5, TAG_F64
        return float(f), float(f + n - 1), TAG_EXACT

    def lbounds(self):
        return tuple(self._bound(d)[0] if self._bound(d)[2] != TAG_EXACT else int(self._bound(d)[0])
                     for d in range(self.ndims))

    def ubounds(self):
        return tuple(self._bound(d)[1] if self._bound(d)[2] != TAG_EXACT else int(self._bound(d)[1])
                     for d in range(self.ndims))

    def _chain(self):
        return self, None, None

    def coefs_array(self) -> np.ndarray:
        """Host copy of the coefficient array with Julia's shape (size) in Fortran order."""
        return self.coefs.cpu().numpy().reshape(self.size, order="F")

    def __repr__(self):
        return f"BSplineInterpolation(size={self.parent_len}, coefs={self.size}, it={self.its})"


class ScaledInterpolation(AbstractInterpolation):
    """`scale(itp, ranges...)` (scaling.jl:5-8,33-48)."""

    def __init__(self, itp: BSplineInterpolation, ranges: Sequence[AbstractRange]):
        if not isinstance(itp, BSplineInterpolation):
            raise ValueError("scale(): only scale(interpolate(...)) is on the B200 hot path "
                             "(scale(extrapolate(...)) is not supported; use extrapolate(scale(...)))")
        if any(f != 1 for f in itp.parent_first):
            raise ValueError("scale(): an interpolate_inplace object with inset axes cannot be scaled "
                             "(coordlookup maps onto 1:n, scaling.jl:102-115)")
        ranges = tuple(ranges)
        if len(ranges) != itp.ndims:
            raise ValueError("scale(): need one range per dimension")
        for d, r in enumerate(ranges):  # check_ranges scaling.jl:39-48
            if not isinstance(r, AbstractRange):
                raise TypeError("scale(): ranges must be range objects (see ranges.jrange)")
            if len(r) != itp.parent_len[d]:
                raise ValueError(f"The range {r} is incommensurate with the corresponding axis 1:{itp.parent_len[d]}")
            if not r.step > 0:
                raise ValueError(f"The range {r} is in decreasing order; only ranges with positive steps are supported by scale().")
            if itp.its[d].degree.code == NOINTERP and not (float(r.first) == 1.0 and float(r.step) == 1.0):  # scaling.jl:46
                raise ValueError(f"The range {r} did not equal the corresponding axis of the interpolation object "
                                 f"1:{itp.parent_len[d]}")
        self.itp = itp
        self.ranges = ranges
        self.ndims = itp.ndims
        self.device_index = itp.device_index

    def _bound(self, d):
        """scaling.jl:60-75,116: OnGrid -> first(r), last(r) in the range's eltype;
        OnCell -> first(r) - boundstep(r), last(r) + boundstep(r), always Float64."""
        r = self.ranges[d]
        gt = self.itp.its[d].degree.bc.gt
        if isinstance(gt, OnCell):
            half = 0.5 * float(r.step)  # boundstep: r.step/2, 0.5*step(r) (Float64 arithmetic)
            return float(r.first) - half, float(r.last) + half, TAG_F64
        return float(r.first), float(r.last), r.tag

    def lbounds(self):
        return tuple(self._bound(d)[0] for d in range(self.ndims))

    def ubounds(self):
        return tuple(self._bound(d)[1] for d in range(self.ndims))

    def _chain(self):
        return self.itp, self, None

    def __repr__(self):
        return f"ScaledInterpolation({self.itp!r}, ranges={self.ranges})"


def _normalise_etp_flags(et, ndims):
    """ExtrapDimSpec (extrapolation.jl:15-19): one BC, a per-dim tuple, or per-dim (lo, hi) pairs."""
    ok = (Throw, Flat, Line, Periodic, Reflect)

    def one(f):
        if isinstance(f, type) and issubclass(f, BoundaryCondition):
            raise ValueError(f"cannot create a filled extrapolation with a type {f.__name__}, use a value of this "
                             f"type instead (e.g., {f.__name__}())")
        if isinstance(f, BoundaryCondition):
            if not isinstance(f, ok):
                raise ValueError(f"extrapolation scheme {f!r} is not supported")
            return f
        if isinstance(f, tuple) and len(f) == 2 and all(isinstance(g, ok) for g in f):
            return (f[0], f[1])
        raise TypeError(f"bad extrapolation flag {f!r}")

    if isinstance(et, BoundaryCondition):
        return (one(et),) * ndims
    if isinstance(et, tuple):
        if len(et) != ndims:
            raise ValueError("extrapolate(): need one scheme per dimension")
        return tuple(one(f) for f in et)
    return None


class Extrapolation(AbstractInterpolation):
    """`extrapolate(itp, scheme)` (extrapolation.jl:1-4,45-46)."""

    def __init__(self, itp, flags):
        self.itp = itp
        self.flags = flags
        self.ndims = itp.ndims
        self.device_index = itp.device_index

    def lbounds(self):
        return self.itp.lbounds()

    def ubounds(self):
        return self.itp.ubounds()

    def _chain(self):
        b, s, _ = self.itp._chain()
        return b, s, self

    def __repr__(self):
        return f"Extrapolation({self.itp!r}, {self.flags})"


class FilledExtrapolation(AbstractInterpolation):
    """`extrapolate(itp, fillvalue)` (filled.jl:1-4,29)."""

    def __init__(self, itp, fillvalue):
        self.itp = itp
        self.fillvalue = fillvalue
        if isinstance(fillvalue, (bool, int, np.integer)):
            self.fill_tag = TAG_EXACT
        elif isinstance(fillvalue, np.float32):
            self.fill_tag = TAG_F32
        else:
            self.fill_tag = TAG_F64
        self.ndims = itp.ndims
        self.device_index = itp.device_index

    def lbounds(self):
        return self.itp.lbounds()

    def ubounds(self):
        return self.itp.ubounds()

    def _chain(self):
        b, s, _ = self.itp._chain()
        return b, s, self

    def __repr__(self):
        return f"FilledExtrapolation({self.itp!r}, {self.fillvalue!r})"


# ----------------------------------------------------------------------------- constructors
def _device_index(device) -> int:
    torch = _torch()
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index or 0


def _to_colmajor_device(A, device_index, alias_ok=False):
    """(flat column-major CUDA tensor, shape).  Integer arrays are converted like `tcoef` does
    (b-splines.jl:229-233): Float32 stays Float32, integers become Float64.  A CUDA tensor that already is
    column-major (its reversed-axes view is contiguous) is not transposed; it is cloned unless `alias_ok`
    (`interpolate!`), because the prefilter works in place."""
    torch = _torch()
    dev = torch.device("cuda", device_index)
    if isinstance(A, np.ndarray):
        if A.dtype not in (np.float32, np.float64):
            A = A.astype(np.float64)
        shape = A.shape
        flat = np.ascontiguousarray(A.reshape(-1, order="F"))
        return torch.from_numpy(flat).to(dev), shape
    if isinstance(A, torch.Tensor):
        if A.dtype not in (torch.float32, torch.float64):
            A = A.to(torch.float64)
        shape = tuple(A.shape)
        t = A.to(dev)
        if t.dim() > 1:
            t = t.permute(*reversed(range(t.dim()))).contiguous()  # no copy when A already is column-major
        flat = t.reshape(-1)
        if flat.data_ptr() == A.data_ptr() and not alias_ok:  # interpolate(): never alias the caller's memory
            flat = flat.clone()
        return flat, shape
    return _to_colmajor_device(np.asarray(A), device_index)


def _stream_ptr(device_index):
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


def prefilter_(coefs, size, its, device_index, src=None):
    """`prefilter!(TW, ret, it)` on a padded device array (prefiltering.jl:49-59,108-122).  With `src` (same size and
    layout) the first axis pass reads the samples from `src` and writes `coefs`: `src` stays untouched and no copy
    is made up front."""
    L = _lib.lib()
    N = len(size)
    sz = (C.c_int64 * N)(*size)
    deg = (C.c_int32 * N)(*[it.degree.code for it in its])
    bc = (C.c_int32 * N)(*[it.degree.bc.code for it in its])
    gt = (C.c_int32 * N)(*[it.degree.bc.gt.code for it in its])
    code = F32 if str(coefs.dtype).endswith("float32") else F64
    if src is not None and src.data_ptr() != coefs.data_ptr():
        rc = L.b200itp_prefilter_from(device_index, C.c_void_p(src.data_ptr()), C.c_void_p(coefs.data_ptr()), code, N, sz,
                                      deg, bc, gt, _stream_ptr(device_index))
        _lib.check(rc, "b200itp_prefilter_from")
        return coefs
    rc = L.b200itp_prefilter(device_index, C.c_void_p(coefs.data_ptr()), code, N, sz, deg, bc, gt,
                             _stream_ptr(device_index))
    _lib.check(rc, "b200itp_prefilter")
    return coefs


def _has_prefiltering_system(it: BSpline) -> bool:
    """Does the reference define `prefiltering_system` for this spec (cubic.jl:101-264, quadratic.jl:71-189)?"""
    deg, bc, gt = it.degree.code, it.degree.bc.code, it.degree.bc.gt.code
    if deg == CUBIC:
        return bc in (BC_FLAT, BC_LINE, 3, BC_PERIODIC)  # Flat, Line, Free, Periodic
    if deg == QUADRATIC:
        if bc in (6, 7):  # InPlace / InPlaceQ: OnCell only
            return gt == ONCELL
        return bc in (BC_FLAT, BC_LINE, 3, BC_PERIODIC, BC_REFLECT)
    return True


# A Cubic / Quadratic boundary condition the reference has no `prefiltering_system` method for (`Cubic(Reflect)`):
# `prefiltering_system(::Any...)` returns `(nothing, nothing)` (prefiltering.jl:124) and
# `A_ldiv_B_md!(dest, ::Nothing, src, dim, ::Nothing) = dest` (filter1d.jl:6) makes the axis pass a no-op in BOTH
# `prefilter!` methods (single spec, prefiltering.jl:49-59, and per-axis tuple, :108-122): the reference returns the
# padded, UNFILTERED array without an error, and so does the library (`_has_prefiltering_system` only reports it).


def interpolate(A, it, device=None) -> BSplineInterpolation:
    """`interpolate(A, it)` (b-splines.jl:167-170,191-193): pad (`copy_with_padding`), prefilter on
    the GPU, wrap.  `A`: numpy array or torch tensor, shape (n1..nN) in Julia index order."""
    torch = _torch()
    L = _lib.lib()
    dev = _device_index(device)
    # the caller's array is only READ: the padded copy, or the first axis pass, writes a new array
    src, shape = _to_colmajor_device(A, dev, alias_ok=True)
    aliased = isinstance(A, torch.Tensor) and src.data_ptr() == A.data_ptr()
    N = len(shape)
    if not 1 <= N <= MAX_DIMS:
        raise ValueError(f"1..{MAX_DIMS} dimensions are supported, got {N}")
    its = _its_tuple(it, N)
    for i in its:
        if not iscomplete(i):
            raise ValueError(f"OnGrid/OnCell is not supplied for some of the interpolation modes in {it}")
    for d in range(N):  # is_singleton_ok b-splines.jl:114-121
        if shape[d] == 1 and its[d].degree.code != NOINTERP:
            raise ValueError(f"size {shape} is inconsistent with {it}, use NoInterp along singleton dimensions")
    pad = [1 if needs_padding(i) else 0 for i in its]
    psize = tuple(shape[d] + 2 * pad[d] for d in range(N))
    code = F32 if src.dtype == torch.float32 else F64
    with torch.cuda.device(dev):
        if any(pad):
            # copy_with_padding + prefilter! in one library call: the padding is folded into the first axis pass
            coefs = torch.empty(int(np.prod(psize)), dtype=src.dtype, device=src.device)
            rc = L.b200itp_prefilter_padded_from(
                dev, C.c_void_p(src.data_ptr()), C.c_void_p(coefs.data_ptr()), code, N, (C.c_int64 * N)(*shape),
                (C.c_int32 * N)(*pad), (C.c_int32 * N)(*[i.degree.code for i in its]),
                (C.c_int32 * N)(*[i.degree.bc.code for i in its]), (C.c_int32 * N)(*[i.degree.bc.gt.code for i in its]),
                _stream_ptr(dev))
            _lib.check(rc, "b200itp_prefilter_padded_from")
        elif aliased:
            coefs = torch.empty_like(src)
            prefilter_(coefs, psize, its, dev, src=src)  # first pass reads A, writes coefs
        else:
            coefs = src  # already a private device copy (numpy / CPU / row-major input)
            prefilter_(coefs, psize, its, dev)
    return BSplineInterpolation(coefs, psize, shape, its, dev)


def interpolate_inplace(A, it, device=None) -> BSplineInterpolation:
    """`interpolate!(A, it)` (b-splines.jl:236-241): the system of size n is solved on the samples themselves — no
    padded copy — and the domain is trimmed where the degree would have padded: those axes are `2:n-1` (`padinset`).
    A 1-D CUDA tensor, or an N-D CUDA tensor whose memory is already column-major (`A.permute(N-1..0)` contiguous, e.g.
    `torch.empty(n3, n2, n1).permute(2, 1, 0)`), is prefiltered IN PLACE like the reference does; every other input
    (numpy, CPU, row-major) is converted to a column-major device copy first and the caller's array is left alone."""
    torch = _torch()
    dev = _device_index(device)
    src, shape = _to_colmajor_device(A, dev, alias_ok=True)
    N = len(shape)
    if not 1 <= N <= MAX_DIMS:
        raise ValueError(f"1..{MAX_DIMS} dimensions are supported, got {N}")
    its = _its_tuple(it, N)
    for i in its:
        if not iscomplete(i):
            raise ValueError(f"OnGrid/OnCell is not supplied for some of the interpolation modes in {it}")
    inset = [1 if needs_padding(i) else 0 for i in its]
    for d in range(N):
        if shape[d] - 2 * inset[d] < 1 or (shape[d] == 1 and its[d].degree.code != NOINTERP):
            raise ValueError(f"size {shape} is inconsistent with {it}")
    with torch.cuda.device(dev):
        prefilter_(src, shape, its, dev)
    return BSplineInterpolation(src, shape, tuple(shape[d] - 2 * inset[d] for d in range(N)), its, dev,
                                parent_first=tuple(1 + inset[d] for d in range(N)))


def scale(itp, *ranges) -> ScaledInterpolation:
    """`scale(itp, ranges...)` (scaling.jl:33-37)."""
    if len(ranges) == 1 and isinstance(ranges[0], (tuple, list)):
        ranges = tuple(ranges[0])
    return ScaledInterpolation(itp, ranges)


def extrapolate(itp, et) -> Union[Extrapolation, FilledExtrapolation]:
    """`extrapolate(itp, scheme)` / `extrapolate(itp, fillvalue)` (extrapolation.jl:45-46, filled.jl:29)."""
    if isinstance(itp, (Extrapolation, FilledExtrapolation)):
        raise ValueError("extrapolate(): nested extrapolation is not on the B200 hot path")
    flags = _normalise_etp_flags(et, itp.ndims)
    if flags is not None:
        return Extrapolation(itp, flags)
    if isinstance(et, (bool, int, float, np.integer, np.floating)):
        return FilledExtrapolation(itp, et)
    raise TypeError(f"extrapolate(): bad scheme {et!r}")


# ----------------------------------------------------------------------------- rrule
def rrule(itp, *xs):
    """`ChainRulesCore.rrule(itp, x...)` (chainrules/chainrules.jl:9-16): returns `(y, pullback)`; `pullback(dy)` gives the
    cotangents of the evaluation points, `dy .* gradient(itp, x...)` per axis (none for the data inside `itp`)."""
    y, g = value_and_gradient(itp, *xs)

    def interpolate_pullback(dy):
        torch = _torch()
        dy = torch.as_tensor(dy, dtype=g.dtype, device=g.device)
        return tuple(dy * g[..., k] for k in range(g.shape[-1]))

    return y, interpolate_pullback


def differentiable(itp):
    """The rrule wired into torch.autograd: `differentiable(itp)(x, y, ...)` evaluates `itp` on CUDA tensors and
    back-propagates to the coordinates through `Interpolations.gradient` (both passes run in libb200itp)."""
    torch = _torch()

    class _Eval(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *xs):
            y, g = value_and_gradient(itp, *[x.detach() for x in xs])
            ctx.save_for_backward(g)
            ctx.shapes = [x.shape for x in xs]
            ctx.dtypes = [x.dtype for x in xs]
            return y

        @staticmethod
        def backward(ctx, dy):
            (g,) = ctx.saved_tensors
            outs = []
            for k, (shape, dt) in enumerate(zip(ctx.shapes, ctx.dtypes)):
                t = (dy * g[..., k]).to(dt)
                outs.append(t.sum_to_size(shape) if tuple(t.shape) != tuple(shape) else t)  # broadcast arguments
            return tuple(outs)

    return _Eval.apply


# ----------------------------------------------------------------------------- convenience constructors
def _knot_ranges(knots):
    ranges = tuple(knots) if isinstance(knots, (tuple, list)) else (knots,)
    if not all(isinstance(r, AbstractRange) for r in ranges):
        raise TypeError("only range knots (the scaled B-spline path) are on the B200 hot path; "
                        "vector knots are Gridded interpolation")
    return ranges


def constant_interpolation(knots, vs, extrapolation_bc=None, device=None):
    """`constant_interpolation(ranges, vs; extrapolation_bc = Throw())` (convenience-constructors.jl:3-4,16-18):
    extrapolate(scale(interpolate(vs, BSpline(Constant())), ranges...), extrapolation_bc)."""
    from .flags import Constant
    etp = Throw() if extrapolation_bc is None else extrapolation_bc
    return extrapolate(scale(interpolate(vs, BSpline(Constant()), device), *_knot_ranges(knots)), etp)


def linear_interpolation(knots, vs, extrapolation_bc=None, device=None):
    """`linear_interpolation(ranges, vs; extrapolation_bc = Throw())` (convenience-constructors.jl:7-8,22-24)."""
    from .flags import Linear
    etp = Throw() if extrapolation_bc is None else extrapolation_bc
    return extrapolate(scale(interpolate(vs, BSpline(Linear()), device), *_knot_ranges(knots)), etp)


def cubic_spline_interpolation(knots, vs, bc=None, extrapolation_bc=None, device=None):
    """`cubic_spline_interpolation(ranges, vs; bc = Line(OnGrid()), extrapolation_bc = Throw())`
    (convenience-constructors.jl:11-13,28-30)."""
    from .flags import Cubic
    etp = Throw() if extrapolation_bc is None else extrapolation_bc
    return extrapolate(scale(interpolate(vs, BSpline(Cubic(Line(OnGrid()) if bc is None else bc)), device),
                             *_knot_ranges(knots)), etp)


# ----------------------------------------------------------------------------- evaluation
def _prepare_coords(itp, xs):
    """Broadcast the N coordinate arguments (numpy broadcasting == Julia's `itp.(xs, ys')`) and put
    them on the device as contiguous 1-D arrays of one float type."""
    torch = _torch()
    N = itp.ndims
    if len(xs) != N:
        raise ValueError(f"expected {N} coordinate arguments, got {len(xs)}")
    dev = torch.device("cuda", itp.device_index)
    x0 = xs[0] if xs else None
    if isinstance(x0, torch.Tensor) and x0.dim() == 1 and x0.dtype in (torch.float32, torch.float64) and all(
            isinstance(x, torch.Tensor) and x.device == dev and x.dtype == x0.dtype and x.shape == x0.shape
            and x.is_contiguous() for x in xs):
        return list(xs), tuple(x0.shape), x0.dtype  # the hot path: SoA device arrays of one type, nothing to do
    ts = []
    for x in xs:
        if isinstance(x, torch.Tensor):
            t = x
        else:
            a = np.asarray(x)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            t = torch.from_numpy(np.ascontiguousarray(a))
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        ts.append(t.to(dev))
    dt = torch.float64 if any(t.dtype == torch.float64 for t in ts) else torch.float32
    ts = [t.to(dt) for t in ts]
    shape = torch.broadcast_shapes(*[t.shape for t in ts])
    ts = [t.expand(shape).contiguous().reshape(-1) for t in ts]
    return ts, tuple(shape), dt


def out_dtype(itp, coord_dtype):
    """Element type of `itp.(xs...)`: promote(coordinate type after all wrappers, eltype(coefs))."""
    D = itp.descriptor(coord_dtype)
    code = _lib.lib().b200itp_out_dtype(C.byref(D))
    if code < 0:
        raise ValueError("invalid descriptor")
    return _np_dtype(code)


_scratch_cache = {}


def _scratch_tensor(nbytes, device):
    """Ordered-pipeline scratch (~32 B/query), kept per (device, stream) and grown geometrically instead of a fresh
    `torch.empty` per broadcast (10^9 queries: 32 GB)."""
    torch = _torch()
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _scratch_cache.get(key)
    if t is None or t.numel() < nbytes:
        _scratch_cache.pop(key, None)
        t = None
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _scratch_cache[key] = t
    return t


def release_scratch():
    """Drop the cached ordered-pipeline scratch buffers."""
    _scratch_cache.clear()


def _can_throw(D) -> bool:
    """Can a query raise BoundsError?  Not when every side of every axis extrapolates (or fills)."""
    if D.etp_kind == ETP_FILL:
        return False
    if D.etp_kind == ETP_FLAGS:
        return any(D.etp_lo[d] == BC_THROW or D.etp_hi[d] == BC_THROW for d in range(D.ndims)) or any(
            D.degree[d] == NOINTERP for d in range(D.ndims))
    return True


def _broadcast_eval(itp, xs, want_value, want_grad, debug=False, sorted_queries=None):
    torch = _torch()
    L = _lib.lib()
    coords, shape, dt = _prepare_coords(itp, xs)
    nq = coords[0].numel() if coords else 0
    N = itp.ndims
    npdt = np.float32 if dt == torch.float32 else np.float64
    D = itp.descriptor(npdt)
    ocode = L.b200itp_out_dtype(C.byref(D))
    if ocode < 0:
        raise ValueError("invalid descriptor")
    odt = torch.float32 if ocode == F32 else torch.float64
    dev = itp.device_index
    coefs = itp.coefficients()
    with torch.cuda.device(dev):
        tdev = coefs.device
        val = torch.empty(nq, dtype=odt, device=tdev) if want_value else None
        G = sum(1 for d in range(N) if D.degree[d] != NOINTERP)  # NoInterp axes have no component (indexing.jl:108)
        grad = torch.empty(nq * G, dtype=odt, device=tdev) if want_grad else None
        dbg_i = torch.empty(nq * N, dtype=torch.int32, device=tdev) if debug else None
        dbg_b = torch.empty(nq, dtype=torch.int32, device=tdev) if debug else None
        ptrs = (C.c_void_p * N)(*[c.data_ptr() for c in coords])
        foob = C.c_int64(-1)
        # the out-of-bounds report costs a stream synchronisation: skipped when no query can throw
        foob_ref = C.byref(foob) if _can_throw(D) else None

        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(None)

        ev = getattr(itp._chain()[0], "ready_event", None)
        if nq > 0 and ev is not None:
            # coefficients still on their way on another stream (parallel.broadcast_interpolant(..., overlap=True)): the
            # library waits for this event right before its first kernel that reads them
            L.b200itp_eval_set_coefs_event(C.c_void_p(ev.cuda_event))
        try:
            # scratch_bytes is 0 when the ordered pipeline would not serve this call (small / L2-resident / wrapped)
            nbytes = 0
            if nq > 0 and sorted_queries and (want_value != want_grad) and not debug:
                nbytes = L.b200itp_eval_sorted_scratch_bytes(C.byref(D), nq)
            if nq > 0 and nbytes:
                scratch = _scratch_tensor(int(nbytes), tdev)
                if want_value:
                    rc = L.b200itp_eval_sorted(dev, C.byref(D), p(coefs), ptrs, nq, p(val), foob_ref, p(scratch),
                                               nbytes, _stream_ptr(dev))
                    _lib.check(rc, "b200itp_eval_sorted")
                else:
                    rc = L.b200itp_gradient_sorted(dev, C.byref(D), p(coefs), ptrs, nq, p(grad), foob_ref,
                                                   p(scratch), nbytes, _stream_ptr(dev))
                    _lib.check(rc, "b200itp_gradient_sorted")
            elif nq > 0:
                rc = L.b200itp_eval(dev, C.byref(D), p(coefs), ptrs, nq, p(val), p(grad), foob_ref, p(dbg_i),
                                    p(dbg_b), _stream_ptr(dev))
                _lib.check(rc, "b200itp_eval")
        finally:
            if ev is not None:  # a call that failed before it consumed the event must not leave it for the next one
                L.b200itp_eval_set_coefs_event(None)
    if foob.value >= 0:
        q = foob.value
        xq = tuple(float(c[q].item()) for c in coords)
        raise BoundsError(f"attempt to access {itp!r} at index {xq}")
    out = [None, None, None, None]
    if want_value:
        out[0] = val.reshape(shape)
    if want_grad:
        out[1] = grad.reshape(shape + (G,))
    if debug:
        out[2] = dbg_i.reshape(shape + (N,))
        out[3] = dbg_b.reshape(shape)
    return out


def gradient(itp, *xs):
    """`Interpolations.gradient.(Ref(itp), xs...)`: result has a trailing axis of length N holding the
    SVector components contiguously (indexing.jl:25-31, scaling.jl:123-128, extrapolation.jl:71-82)."""
    return _broadcast_eval(itp, xs, want_value=False, want_grad=True, sorted_queries=True)[1]


def value_and_gradient(itp, *xs):
    r = _broadcast_eval(itp, xs, want_value=True, want_grad=True)
    return r[0], r[1]


def evaluate_debug(itp, *xs, want_grad=False):
    """value (+gradient) plus the bit-exact contract: base tap index per axis and branch record."""
    return _broadcast_eval(itp, xs, want_value=True, want_grad=want_grad, debug=True)


def evaluate_direct(itp, *xs):
    """`itp.(xs...)` through the direct kernel only (one thread per query, no reordering): the reference point the
    ordered pipeline is compared with bit for bit."""
    return _broadcast_eval(itp, xs, want_value=True, want_grad=False, sorted_queries=False)[0]


def gradient_direct(itp, *xs):
    """`Interpolations.gradient.(Ref(itp), xs...)` through the direct kernel only."""
    return _broadcast_eval(itp, xs, want_value=False, want_grad=True, sorted_queries=False)[1]


def evaluate_sorted(itp, *xs):
    """`itp.(xs...)` with on-device L2-friendly query reordering (results in the caller's order)."""
    return _broadcast_eval(itp, xs, want_value=True, want_grad=False, sorted_queries=True)[0]


def evaluate_from_host(itp, *host_xs, out=None, chunk=1 << 26, ordered=True, copy_streams=1):
    """`itp.(xs...)` for coordinate arrays that live in (pinned) HOST memory: the queries are cut into chunks and the
    host->device copy of chunk i+1, the evaluation of chunk i and the device->host copy of chunk i-1 run on three
    CUDA streams, so the call is bound by the PCIe transfer of the coordinates instead of the sum of the three.
    `host_xs`: N 1-D CPU torch tensors of one float dtype (pinned for asynchronous copies); `out`: optional pinned CPU
    tensor for the results.  Returns the result tensor on the host (values identical to `itp(*xs)`)."""
    torch = _torch()
    N = itp.ndims
    if len(host_xs) != N:
        raise ValueError(f"expected {N} coordinate arguments, got {len(host_xs)}")
    nq = int(host_xs[0].numel())
    dt = host_xs[0].dtype
    if any(x.dtype != dt or x.numel() != nq or x.device.type != "cpu" for x in host_xs):
        raise ValueError("evaluate_from_host: N CPU tensors of equal length and dtype are required")
    dev_index = itp.device_index
    dev = torch.device("cuda", dev_index)
    odt = torch.float32 if out_dtype(itp, np.float32 if dt == torch.float32 else np.float64) == np.float32 else torch.float64
    if out is None:
        out = torch.empty(nq, dtype=odt, pin_memory=True)
    if out.numel() != nq or out.dtype != odt:
        raise ValueError("evaluate_from_host: `out` has the wrong length or dtype")
    if nq == 0:
        return out
    chunk = max(1, int(chunk))
    nchunks = (nq + chunk - 1) // chunk
    with torch.cuda.device(dev_index):
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        s_in.wait_stream(main)
        # optionally one upload stream per coordinate array (copy_streams > 1): concurrent copies on several copy engines
        s_ins = [s_in] + [torch.cuda.Stream() for _ in range(max(0, min(int(copy_streams), N) - 1))]
        for s_extra in s_ins[1:]:
            s_extra.wait_stream(main)
        nbuf = min(3, nchunks)
        bufs = [[torch.empty(min(chunk, nq), dtype=dt, device=dev) for _ in range(N)] for _ in range(nbuf)]
        free_ev = [None] * nbuf   # buffer may be refilled once its evaluation has finished
        results = [None] * nbuf
        done_ev = [None] * nbuf   # result buffer may be reused once its D2H copy has finished
        ready = [None] * nbuf

        def upload(c):
            lo, hi = c * chunk, min(nq, (c + 1) * chunk)
            b = c % nbuf
            for k, s_k in enumerate(s_ins[1:], start=1):  # extra upload streams: their arrays first
                with torch.cuda.stream(s_k):
                    if free_ev[b] is not None:
                        s_k.wait_event(free_ev[b])
                    for d in range(k, N, len(s_ins)):
                        bufs[b][d][: hi - lo].copy_(host_xs[d][lo:hi], non_blocking=True)
            with torch.cuda.stream(s_in):
                if free_ev[b] is not None:
                    s_in.wait_event(free_ev[b])
                for d in range(0, N, len(s_ins)):
                    bufs[b][d][: hi - lo].copy_(host_xs[d][lo:hi], non_blocking=True)
                for s_k in s_ins[1:]:
                    s_in.wait_stream(s_k)
                ready[b] = torch.cuda.Event()
                ready[b].record(s_in)

        upload(0)
        for c in range(nchunks):
            lo, hi = c * chunk, min(nq, (c + 1) * chunk)
            b = c % nbuf
            if c + 1 < nchunks:
                upload(c + 1)  # queued before the (host-blocking) evaluation of chunk c: the copy engine never idles
            main.wait_event(ready[b])
            if done_ev[b] is not None:
                main.wait_event(done_ev[b])
            xs = [bufs[b][d][: hi - lo] for d in range(N)]
            results[b] = evaluate_sorted(itp, *xs) if ordered else itp(*xs)
            free_ev[b] = torch.cuda.Event()
            free_ev[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(free_ev[b])
                out[lo:hi].copy_(results[b], non_blocking=True)
                results[b].record_stream(s_out)
                done_ev[b] = torch.cuda.Event()
                done_ev[b].record(s_out)
        main.wait_stream(s_out)
        for s_k in s_ins:
            main.wait_stream(s_k)
        # `out` lives in host memory: the caller may read it as soon as this returns, so the last device->host copies
        # must have LANDED (stream ordering alone does not block the CPU)
        s_out.synchronize()
    return out


def evaluate_grid(itp, *vectors):
    """`itp(xvec, yvec, ...)`: tensor-product evaluation (indexing.jl:16-23, scaling.jl:98-100, extrapolation.jl:59-70).
    Returns a tensor of shape (len(xvec), len(yvec), ...) whose element [i, j, ...] is `itp(xvec[i], yvec[j], ...)`
    (a permuted view of the Julia-ordered result: the first axis is the fastest in memory)."""
    torch = _torch()
    L = _lib.lib()
    N = itp.ndims
    if len(vectors) != N:
        raise ValueError(f"expected {N} coordinate vectors, got {len(vectors)}")
    dev = itp.device_index
    tdev = torch.device("cuda", dev)
    ts = []
    for v in vectors:
        t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.atleast_1d(np.asarray(v)))
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        ts.append(t.reshape(-1).to(tdev))
    dt = torch.float64 if any(t.dtype == torch.float64 for t in ts) else torch.float32
    ts = [t.to(dt).contiguous() for t in ts]
    npdt = np.float32 if dt == torch.float32 else np.float64
    D = itp.descriptor(npdt)
    ocode = L.b200itp_out_dtype(C.byref(D))
    if ocode < 0:
        raise ValueError("invalid descriptor")
    odt = torch.float32 if ocode == F32 else torch.float64
    lens = [int(t.numel()) for t in ts]
    nq = int(np.prod(lens))
    coefs = itp.coefficients()
    foob = C.c_int64(-1)
    with torch.cuda.device(dev):
        out = torch.empty(nq, dtype=odt, device=tdev)
        if nq > 0:
            ptrs = (C.c_void_p * N)(*[t.data_ptr() for t in ts])
            clens = (C.c_int64 * N)(*lens)
            nbytes = int(L.b200itp_eval_grid_scratch_bytes(C.byref(D), clens))
            scratch = _scratch_tensor(max(nbytes, 1), tdev)
            rc = L.b200itp_eval_grid(dev, C.byref(D), C.c_void_p(coefs.data_ptr()), ptrs, clens,
                                     C.c_void_p(out.data_ptr()), C.byref(foob), C.c_void_p(scratch.data_ptr()), nbytes,
                                     _stream_ptr(dev))
            _lib.check(rc, "b200itp_eval_grid")
    if foob.value >= 0:
        idx = np.unravel_index(foob.value, lens, order="F")
        xq = tuple(float(ts[d][idx[d]].item()) for d in range(N))
        raise BoundsError(f"attempt to access {itp!r} at index {xq}")
    if isinstance(itp, Extrapolation):
        # extrapolation.jl:61-64: the result array has eltype promote_type(T, eltype.(x)...) and stores converted values
        b = itp._chain()[0]
        tret = torch.float64 if (dt == torch.float64 or b.dtype_code == F64) else torch.float32
        out = out.to(tret)
    return out.reshape(tuple(reversed(lens))).permute(*reversed(range(N)))


def eachvalue(sitp):
    """`collect(eachvalue(sitp))` (scaling.jl:174-256): the scaled interpolant at every integer point of its ranges,
    `ceil(Int, first(r)):floor(Int, last(r))` per axis — one separable grid evaluation (each axis' weights computed once
    per coordinate, each coefficient row reused across the axis-1 run, which is what the reference's ScaledIterator
    caches by hand)."""
    if not isinstance(sitp, ScaledInterpolation):
        raise TypeError("eachvalue: a ScaledInterpolation is required")
    vecs = []
    for r in sitp.ranges:
        lo, hi = int(np.ceil(float(r.first))), int(np.floor(float(r.last)))
        dt = np.float32 if r.tag == TAG_F32 else np.float64
        vecs.append(np.arange(lo, hi + 1).astype(dt))
    return evaluate_grid(sitp, *vecs)


def hessian(itp, *xs):
    """`Interpolations.hessian.(Ref(itp), xs...)`: trailing axes (N, N) hold the symmetric matrix of every query
    (indexing.jl:37-41, scaling.jl:150-160; extrapolated interpolants: in bounds only, extrapolation.jl:84-90)."""
    torch = _torch()
    L = _lib.lib()
    coords, shape, dt = _prepare_coords(itp, xs)
    nq = coords[0].numel() if coords else 0
    N = itp.ndims
    npdt = np.float32 if dt == torch.float32 else np.float64
    D = itp.descriptor(npdt)
    # output type: the in-bounds path never meets the extrapolation bounds, so they do not widen Float32 coordinates
    Dn = itp.descriptor(npdt)
    Dn.etp_kind = 0
    ocode = L.b200itp_out_dtype(C.byref(Dn))
    if ocode < 0:
        raise ValueError("invalid descriptor")
    odt = torch.float32 if ocode == F32 else torch.float64
    dev = itp.device_index
    coefs = itp.coefficients()
    foob = C.c_int64(-1)
    with torch.cuda.device(dev):
        out = torch.empty(nq * N * N, dtype=odt, device=coefs.device)
        if nq > 0:
            ptrs = (C.c_void_p * N)(*[c.data_ptr() for c in coords])
            rc = L.b200itp_hessian(dev, C.byref(D), C.c_void_p(coefs.data_ptr()), ptrs, nq, C.c_void_p(out.data_ptr()),
                                   C.byref(foob), _stream_ptr(dev))
            _lib.check(rc, "b200itp_hessian")
    if foob.value >= 0:
        q = foob.value
        xq = tuple(float(c[q].item()) for c in coords)
        raise BoundsError(f"attempt to access the hessian of {itp!r} at index {xq}")
    return out.reshape(shape + (N, N))


def gradient_sorted(itp, *xs):
    """`Interpolations.gradient.(Ref(itp), xs...)` through the ordered pipeline (results in the caller's order, identical
    to `gradient`)."""
    return _broadcast_eval(itp, xs, want_value=False, want_grad=True, sorted_queries=True)[1]


def bounds(itp):
    return itp.bounds()


def lbounds(itp):
    return itp.lbounds()


def ubounds(itp):
    return itp.ubounds()

"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

The oracle's OWN description of an interpolant: every number the C oracle consumes (padded sizes, index of the first
coefficient, bounds of the bare and the scaled object, their Julia types, extrapolation flags) is derived here from the
primitive specification — (degree, boundary condition, grid type) per axis, the parent axes, the range parameters, the
flags — following the reference line by line.  The product builds its `b200itp_desc` in
`interpolations.jl_b200/core.py:descriptor()`; the two derivations share no code, and
tests/test_oracle_pinning.py::test_descriptor_derivations_agree compares them field by field, so a wrong bound or
padding rule on one side shows up as a disagreement instead of being bit-exactly wrong on both.

Only the field LAYOUT (the ctypes struct of include/b200itp.h) and the integer codes are shared.
"""
from __future__ import annotations

import os
import sys
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import b200_loader  # noqa: E402

b200_loader.load()
from interpolations_jl_b200 import _lib as _pl  # noqa: E402  (struct layout only)

# codes of include/b200itp.h
_F32, _F64 = 0, 1
_EXACT, _T32, _T64 = 0, 1, 2          # Julia type of a bound / range parameter: Int or Rational, Float32, Float64
_LINEAR, _QUADRATIC, _CUBIC, _NOINTERP = 1, 2, 3, 7
_PERIODIC, _INPLACE, _INPLACEQ = 4, 6, 7
_ONCELL = 1
_ETP_NONE, _ETP_FLAGS, _ETP_FILL = 0, 1, 2


def pads(degree: int, bc: int) -> bool:
    """padded_axis: Cubic pads by one cell per side unless Periodic (cubic.jl:86-87); Quadratic unless Periodic /
    InPlace / InPlaceQ (quadratic.jl:64-65); Linear, Constant, NoInterp never (linear.jl:60, constant.jl:108)."""
    if degree == _CUBIC:
        return bc != _PERIODIC
    if degree == _QUADRATIC:
        return bc not in (_PERIODIC, _INPLACE, _INPLACEQ)
    return False


def bspline_bounds(first: int, last: int, degree: int, grid: int):
    """lbound / ubound of a BSplineInterpolation axis (b-splines.jl:141-158).  Constant, Linear, Quadratic and Cubic are
    all `DegreeBC` (constant.jl:32, linear.jl:1, quadratic.jl:1, cubic.jl:1): on an OnCell grid the axis extends half a
    cell, `first(ax) - 0.5` / `last(ax) + 0.5` (Float64, :155-156); on OnGrid, and for NoInterp
    (`lbound(ax, deg::Degree) = first(ax)`, :148), the bounds are the Ints `first(ax)`, `last(ax)`."""
    if degree != _NOINTERP and grid == _ONCELL:
        return first - 0.5, last + 0.5, _T64
    return float(first), float(last), _EXACT


def scaled_bounds(r_first, r_step, r_last, r_tag: int, grid: int, unit_range: bool, degree: int):
    """lbounds / ubounds of a ScaledInterpolation axis (scaling.jl:60-75,116): OnGrid -> first(r), last(r) in the
    range's element type; OnCell -> first(r) - boundstep(r), last(r) + boundstep(r) with boundstep = 0.5*step(r)
    (scaling.jl:116; `r.step/2` for StepRange, `1//2` for UnitRange: the same Float64 numbers)."""
    if grid == _ONCELL and degree != _NOINTERP:
        half = 0.5 * float(r_step)
        return float(r_first) - half, float(r_last) + half, _T64
    return float(r_first), float(r_last), r_tag


def describe(itp, coord_dtype) -> "_pl.Desc":
    """b200itp_desc of `itp` for coordinates of `coord_dtype`, derived from the specification alone."""
    b, s, e = itp._chain()
    N = b.ndims
    D = _pl.Desc()
    D.ndims = N
    D.coef_dtype = _F32 if str(b.coefs.dtype).endswith("float32") else _F64
    D.coord_dtype = _F32 if np.dtype(coord_dtype) == np.float32 else _F64
    D.etp_pair_mask = 0
    for d in range(N):
        deg = b.its[d].degree.code
        bc = b.its[d].degree.bc.code
        gt = b.its[d].degree.bc.gt.code
        first = int(b.parent_first[d])           # first(parentaxes[d]): 1, or 2 after interpolate! (prefiltering.jl:6)
        n = int(b.parent_len[d])
        last = first + n - 1
        padded = pads(deg, bc)
        # interpolate (b-splines.jl:167-170): coefs axes = padded_axis(1:n) = 0:n+1 when padded, else 1:n.
        # interpolate! (b-splines.jl:236-241): coefs axes = axes(A) = 1:m, parent axes = padinset = 2:m-1 when padded.
        if first == 1:
            size = n + 2 if padded else n
            coef_first = 0 if padded else 1
        else:
            assert first == 2 and padded, "parent axes start at 2 only for interpolate! on a padding degree"
            size = n + 2
            coef_first = 1
        D.size[d] = size
        D.parent_len[d] = n
        D.parent_first[d] = first
        D.coef_first[d] = coef_first
        D.degree[d], D.bc[d], D.gridtype[d] = deg, bc, gt
        lo, hi, tag = bspline_bounds(first, last, deg, gt)
        D.inner_lb[d], D.inner_ub[d], D.inner_tag[d] = lo, hi, tag
    D.has_scale = 0
    if s is not None:
        D.has_scale = 1
        for d in range(N):
            r = s.ranges[d]
            # first(r), step(r), last(r): primitives of the range object, each already rounded once to the range's
            # element type by Julia's range arithmetic (ranges.py); last(r) is NOT first + (n-1)*step(r) in floating point
            tag = r.tag
            D.range_first[d], D.range_step[d] = float(r.first), float(r.step)
            last_v = float(r.last)
            D.range_tag[d] = tag
            lo, hi, btag = scaled_bounds(D.range_first[d], D.range_step[d], last_v, tag, b.its[d].degree.bc.gt.code,
                                         type(r).__name__ == "UnitRange", b.its[d].degree.code)
            D.outer_lb[d], D.outer_ub[d], D.outer_tag[d] = lo, hi, btag
    D.etp_kind = _ETP_NONE
    if e is not None and hasattr(e, "flags"):      # Extrapolation (extrapolation.jl:1-4)
        D.etp_kind = _ETP_FLAGS
        for d in range(N):
            f = e.flags[d]
            if isinstance(f, tuple):               # per-side pair (extrapolation.jl:118-127)
                D.etp_pair_mask |= 1 << d
                D.etp_lo[d], D.etp_hi[d] = f[0].code, f[1].code
            else:
                D.etp_lo[d] = D.etp_hi[d] = f.code
    elif e is not None:                            # FilledExtrapolation (filled.jl:1-4)
        D.etp_kind = _ETP_FILL
        fv = e.fillvalue
        D.fill_value = float(fv)
        D.fill_tag = _EXACT if isinstance(fv, (bool, int, np.integer)) else (_T32 if isinstance(fv, np.float32) else _T64)
    return D


FIELDS_SCALAR = ["ndims", "coef_dtype", "coord_dtype", "etp_pair_mask", "has_scale", "etp_kind"]
FIELDS_AXIS = ["size", "parent_len", "coef_first", "degree", "bc", "gridtype", "inner_lb", "inner_ub", "inner_tag",
               "parent_first"]
FIELDS_SCALE = ["range_tag", "range_first", "range_step", "outer_lb", "outer_ub", "outer_tag"]
FIELDS_ETP = ["etp_lo", "etp_hi"]


def differences(A, B):
    """Field-by-field comparison of two descriptors: list of (field, axis, a, b)."""
    out = []
    for f in FIELDS_SCALAR:
        if getattr(A, f) != getattr(B, f):
            out.append((f, None, getattr(A, f), getattr(B, f)))
    N = A.ndims
    fields = list(FIELDS_AXIS)
    if A.has_scale:
        fields += FIELDS_SCALE
    if A.etp_kind == _ETP_FLAGS:
        fields += FIELDS_ETP
    for f in fields:
        for d in range(N):
            a, b = getattr(A, f)[d], getattr(B, f)[d]
            if a != b:
                out.append((f, d, a, b))
    if A.etp_kind == _ETP_FILL:
        for f in ("fill_value", "fill_tag"):
            if getattr(A, f) != getattr(B, f) and not (getattr(A, f) != getattr(A, f) and getattr(B, f) != getattr(B, f)):
                out.append((f, None, getattr(A, f), getattr(B, f)))
    return out

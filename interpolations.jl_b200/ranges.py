"""Host-side stand-ins for the Julia range types `scale` accepts.

Reference: src/scaling/scaling.jl:33-48 (`scale`, `check_ranges`), :60-75,116 (bounds,
`boundstep`), :102-115 (`coordlookup`), :116,137-141 (`rescale_gradient`).

Only three numbers of a range reach the kernels: `first(r)`, `step(r)` and `last(r)`,
each with the Julia element type it had (Int ranges do not widen Float32 coordinates,
Float64 ranges do).  In a Julia deployment the glue passes `first/step/last` computed by
Julia itself (TwicePrecision and all); here they are computed with exact rationals and
rounded once, which is what Julia's range arithmetic is designed to deliver.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from .flags import TAG_EXACT, TAG_F32, TAG_F64


def _frac(x) -> Fraction:
    """The rational a Julia user meant by a float literal (cf. Base.rat used by `range`)."""
    if isinstance(x, (int, np.integer)):
        return Fraction(int(x))
    if isinstance(x, Fraction):
        return x
    if isinstance(x, np.float32):
        return Fraction(np.format_float_positional(x, unique=True, trim="-"))
    return Fraction(repr(float(x)))


class AbstractRange:
    tag: int
    first: float
    step: float
    last: float
    length: int

    def __len__(self):
        return self.length

    def _np_dtype(self):
        return {TAG_EXACT: np.int64, TAG_F32: np.float32, TAG_F64: np.float64}[self.tag]

    def collect(self) -> np.ndarray:
        """`collect(r)`: every element rounded once from the exact rational value."""
        f, s = Fraction(self._first_exact), Fraction(self._step_exact)
        dt = self._np_dtype()
        return np.array([dt(float(f + i * s)) if self.tag != TAG_EXACT else int(f + i * s)
                         for i in range(self.length)], dtype=dt)

    def __repr__(self):
        return f"{type(self).__name__}(first={self.first!r}, step={self.step!r}, last={self.last!r}, length={self.length})"


class UnitRange(AbstractRange):
    """`a:b` with integer endpoints."""

    def __init__(self, start: int, stop: int):
        self.tag = TAG_EXACT
        self.first, self.step, self.last = int(start), 1, int(stop)
        self.length = max(0, int(stop) - int(start) + 1)
        self._first_exact, self._step_exact = Fraction(int(start)), Fraction(1)
        self.is_unit = True


class StepRange(AbstractRange):
    """`a:s:b` with integer start/step."""

    def __init__(self, start: int, step: int, stop: int):
        if step == 0:
            raise ValueError("step cannot be zero")
        self.tag = TAG_EXACT
        n = max(0, (int(stop) - int(start)) // int(step) + 1)
        self.first, self.step, self.length = int(start), int(step), n
        self.last = int(start) + (n - 1) * int(step)
        self._first_exact, self._step_exact = Fraction(int(start)), Fraction(int(step))
        self.is_unit = False


class FloatRange(AbstractRange):
    """StepRangeLen / LinRange with Float32 or Float64 elements."""

    def __init__(self, first: Fraction, step: Fraction, length: int, dtype):
        dt = np.dtype(dtype)
        if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("FloatRange dtype must be float32 or float64")
        self.tag = TAG_F32 if dt == np.dtype(np.float32) else TAG_F64
        self._first_exact, self._step_exact = first, step
        self.length = int(length)
        cast = dt.type
        self.first = float(cast(float(first)))
        self.step = float(cast(float(step)))
        self.last = float(cast(float(first + (length - 1) * step)))
        self.is_unit = False


def jrange(start, stop=None, *, length=None, step=None, dtype=None) -> AbstractRange:
    """`range(start, stop; length)` / `range(start; step, length)` / `start:step:stop`.

    Integer arguments without `length` give UnitRange/StepRange; anything else a FloatRange whose
    dtype defaults to float64 (float32 if every float argument is np.float32)."""
    args = [a for a in (start, stop, step) if a is not None]
    all_int = all(isinstance(a, (int, np.integer)) for a in args)
    if all_int and length is None and dtype is None:
        if step is None:
            return UnitRange(start, stop)
        return StepRange(start, step, stop)
    if dtype is None:
        fl = [a for a in args if not isinstance(a, (int, np.integer))]
        dtype = np.float32 if fl and all(isinstance(a, np.float32) for a in fl) else np.float64
    a = _frac(start)
    if length is not None and stop is not None:
        if length < 2:
            raise ValueError("range needs length >= 2")
        s = (_frac(stop) - a) / (length - 1)
        return FloatRange(a, s, length, dtype)
    if step is not None and length is not None:
        return FloatRange(a, _frac(step), length, dtype)
    if step is not None and stop is not None:
        s = _frac(step)
        n = int((_frac(stop) - a) // s) + 1
        return FloatRange(a, s, max(n, 0), dtype)
    raise ValueError("jrange: need (start, stop, length), (start, step, length) or (start, step, stop)")

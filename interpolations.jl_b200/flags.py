"""Interpolation-type flags, mirroring the reference's type lattice.

Reference: src/Interpolations.jl:49-123 (Flag, GridType, BoundaryCondition),
src/b-splines/b-splines.jl:10-24 (BSpline), linear.jl:1-10, quadratic.jl:1-8, cubic.jl:1-9.
Behaviour is selected entirely by these values, as in the reference (no config files).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

# integer codes shared with include/b200itp.h
F32, F64 = 0, 1
TAG_EXACT, TAG_F32, TAG_F64 = 0, 1, 2
LINEAR, QUADRATIC, CUBIC = 1, 2, 3
CONSTANT_NEAREST, CONSTANT_PREVIOUS, CONSTANT_NEXT = 4, 5, 6
NOINTERP = 7
BC_THROW, BC_FLAT, BC_LINE, BC_FREE, BC_PERIODIC, BC_REFLECT = 0, 1, 2, 3, 4, 5
BC_INPLACE, BC_INPLACEQ = 6, 7
ONGRID, ONCELL = 0, 1
ETP_NONE, ETP_FLAGS, ETP_FILL = 0, 1, 2


class GridType:
    code: int = -1

    def __repr__(self):
        return f"{type(self).__name__}()"

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)


class OnGrid(GridType):
    """Boundary condition applies at the first & last nodes (Interpolations.jl:53-54)."""
    code = ONGRID


class OnCell(GridType):
    """Boundary condition applies half a grid spacing beyond the edge nodes (Interpolations.jl:55-56)."""
    code = ONCELL


def _as_gt(gt):
    if gt is None or isinstance(gt, GridType):
        return gt
    if isinstance(gt, type) and issubclass(gt, GridType):
        return gt()
    raise TypeError(f"expected OnGrid()/OnCell(), got {gt!r}")


@dataclass(frozen=True, eq=False)
class BoundaryCondition:
    """`BC(gt)`; `BC()` has gt = None and is what extrapolation uses (Interpolations.jl:89-118)."""
    gt: Optional[GridType] = None
    code = -1

    def __init__(self, gt=None):
        object.__setattr__(self, "gt", _as_gt(gt))

    def __call__(self, gt):
        return type(self)(gt)

    def __repr__(self):
        return f"{type(self).__name__}({'' if self.gt is None else repr(self.gt)})"

    def __eq__(self, other):
        return type(self) is type(other) and self.gt == other.gt

    def __hash__(self):
        return hash((type(self).__name__, self.gt))


class Throw(BoundaryCondition):
    code = BC_THROW


class Flat(BoundaryCondition):
    code = BC_FLAT


class Line(BoundaryCondition):
    code = BC_LINE


class Free(BoundaryCondition):
    code = BC_FREE


class Periodic(BoundaryCondition):
    code = BC_PERIODIC


class Reflect(BoundaryCondition):
    code = BC_REFLECT


class InPlace(BoundaryCondition):
    """`InPlace(gt)`: prefiltering in place, no padding (Interpolations.jl:104-105; Quadratic OnCell only,
    quadratic.jl:91-95)."""
    code = BC_INPLACE


class InPlaceQ(BoundaryCondition):
    """`InPlaceQ(gt)`: like InPlace, exact for an underlying quadratic (Interpolations.jl:106-109, quadratic.jl:98-110)."""
    code = BC_INPLACEQ


Natural = Line  # Interpolations.jl:110


class Degree:
    code = -1
    bc: BoundaryCondition

    def __repr__(self):
        return f"{type(self).__name__}({self.bc!r})"

    def __eq__(self, other):
        return type(self) is type(other) and self.bc == other.bc

    def __hash__(self):
        return hash((type(self).__name__, self.bc))


class Linear(Degree):
    """`Linear()` == Linear(Throw(OnGrid())); `Linear(Periodic())` == Linear(Periodic(OnCell())) (linear.jl:1-10)."""
    code = LINEAR

    def __init__(self, bc: Optional[BoundaryCondition] = None):
        if bc is None:
            bc = Throw(OnGrid())
        elif isinstance(bc, Periodic) and bc.gt is None:
            bc = Periodic(OnCell())
        ok = (isinstance(bc, Throw) and isinstance(bc.gt, OnGrid)) or (
            isinstance(bc, Periodic) and isinstance(bc.gt, OnCell))
        if not ok:
            raise TypeError(f"Linear accepts only Throw(OnGrid()) or Periodic(OnCell()), got {bc!r}")
        self.bc = bc


class ConstantInterpType:
    """`Nearest` / `Previous` / `Next` (constant.jl:1-31)."""
    code = -1

    def __repr__(self):
        return type(self).__name__

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)


class Nearest(ConstantInterpType):
    code = CONSTANT_NEAREST


class Previous(ConstantInterpType):
    code = CONSTANT_PREVIOUS


class Next(ConstantInterpType):
    code = CONSTANT_NEXT


class Constant(Degree):
    """`Constant()` == Constant{Nearest}(Throw(OnGrid())); `Constant(Previous)`, `Constant(Next)`;
    `Constant(kind, Periodic())` == Constant{kind}(Periodic(OnCell())) (constant.jl:33-49)."""

    def __init__(self, kind=None, bc: Optional[BoundaryCondition] = None):
        if isinstance(kind, BoundaryCondition) and bc is None:  # Constant(bc) defaults to Nearest
            kind, bc = None, kind
        if kind is None:
            kind = Nearest()
        elif isinstance(kind, type) and issubclass(kind, ConstantInterpType):
            kind = kind()
        if not isinstance(kind, ConstantInterpType):
            raise TypeError("Constant(kind): kind must be Nearest, Previous or Next")
        if bc is None:
            bc = Throw(OnGrid())
        elif isinstance(bc, Periodic) and bc.gt is None:
            bc = Periodic(OnCell())
        ok = (isinstance(bc, Throw) and isinstance(bc.gt, OnGrid)) or (
            isinstance(bc, Periodic) and isinstance(bc.gt, OnCell))
        if not ok:
            raise TypeError(f"Constant accepts only Throw(OnGrid()) or Periodic(OnCell()), got {bc!r}")
        self.kind = kind
        self.bc = bc
        self.code = kind.code

    def __repr__(self):
        return f"Constant({self.kind!r}, {self.bc!r})"

    def __eq__(self, other):
        return type(self) is type(other) and self.kind == other.kind and self.bc == other.bc

    def __hash__(self):
        return hash(("Constant", self.kind, self.bc))


class Quadratic(Degree):
    """`Quadratic(bc)`; default Line(OnGrid()) (quadratic.jl:4-8)."""
    code = QUADRATIC

    def __init__(self, bc: Optional[BoundaryCondition] = None):
        self.bc = Line(OnGrid()) if bc is None else bc
        if not isinstance(self.bc, BoundaryCondition):
            raise TypeError("Quadratic(bc): bc must be a BoundaryCondition")


class Cubic(Degree):
    """`Cubic(bc)`; default Line(OnGrid()) (cubic.jl:5-9)."""
    code = CUBIC

    def __init__(self, bc: Optional[BoundaryCondition] = None):
        self.bc = Line(OnGrid()) if bc is None else bc
        if not isinstance(self.bc, BoundaryCondition):
            raise TypeError("Cubic(bc): bc must be a BoundaryCondition")


class BSpline:
    """`BSpline(degree)` (b-splines.jl:10-24)."""

    def __init__(self, degree: Degree):
        if not isinstance(degree, Degree):
            raise TypeError("BSpline(degree): degree must be Constant/Linear/Quadratic/Cubic")
        self.degree = degree

    def __repr__(self):
        return f"BSpline({self.degree!r})"

    def __eq__(self, other):
        return isinstance(other, BSpline) and self.degree == other.degree

    def __hash__(self):
        return hash(("BSpline", self.degree))


class _NoInterpDegree(Degree):
    code = NOINTERP

    def __init__(self):
        self.bc = Throw(OnGrid())  # bounds first(ax)..last(ax) (nointerp.jl:14-15); no boundary condition of its own


class NoInterp(BSpline):
    """`NoInterp()`: the axis takes integer indices, no interpolation (Interpolations.jl:51-52, nointerp.jl).  Modelled
    as a BSpline-like flag so that per-axis tuples `(BSpline(Linear()), NoInterp())` go through the same plumbing."""

    def __init__(self):
        self.degree = _NoInterpDegree()

    def __repr__(self):
        return "NoInterp()"

    def __eq__(self, other):
        return isinstance(other, NoInterp)

    def __hash__(self):
        return hash("NoInterp")


def iscomplete(it: BSpline) -> bool:
    """OnGrid/OnCell must be supplied (b-splines.jl:96)."""
    return it.degree.bc.gt is not None


def needs_padding(it: BSpline) -> bool:
    """padded_axis: Cubic/Quadratic pad by one on each side unless Periodic
    (cubic.jl:86-87, quadratic.jl:64-65); Linear never (linear.jl:60)."""
    d = it.degree
    return d.code in (QUADRATIC, CUBIC) and not isinstance(d.bc, (Periodic, InPlace, InPlaceQ))

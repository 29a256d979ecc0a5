"""interpolations.jl_b200 — B200-native drop-in for the B-spline hot path of Interpolations.jl.

Public names mirror the reference (src/Interpolations.jl:3-31).  The directory name contains a
dot, so it is imported through `b200_loader.load()` (repo root) under the module name
`interpolations_jl_b200`.
"""
from .flags import (InPlace, InPlaceQ)
from .flags import (BSpline, Constant, Cubic, Flat, Free, Line, Linear, Natural, Nearest, Next, NoInterp, OnCell, OnGrid, Periodic,
                    Previous, Quadratic, Reflect, Throw)
from .ranges import FloatRange, StepRange, UnitRange, jrange
from .core import (constant_interpolation, cubic_spline_interpolation, differentiable, linear_interpolation, rrule)
from .core import (BoundsError, BSplineInterpolation, Extrapolation, FilledExtrapolation, ScaledInterpolation,
                   bounds, eachvalue, evaluate_debug, evaluate_direct, evaluate_from_host, evaluate_grid, evaluate_sorted, gradient_direct, extrapolate, gradient, gradient_sorted, hessian, interpolate, interpolate_inplace, lbounds, out_dtype,
                   prefilter_, scale, ubounds, value_and_gradient)
from . import _lib, parallel

__all__ = [
    "constant_interpolation", "cubic_spline_interpolation", "differentiable", "linear_interpolation", "rrule",
    "InPlace", "InPlaceQ", "BSpline", "Constant", "Nearest", "Previous", "Next", "NoInterp", "Cubic", "Flat", "Free", "Line", "Linear", "Natural", "OnCell", "OnGrid", "Periodic", "Quadratic",
    "Reflect", "Throw", "FloatRange", "StepRange", "UnitRange", "jrange", "BoundsError", "BSplineInterpolation",
    "Extrapolation", "FilledExtrapolation", "ScaledInterpolation", "evaluate_direct", "gradient_direct", "eachvalue", "bounds", "evaluate_debug", "evaluate_from_host", "evaluate_grid",
    "evaluate_sorted",
    "extrapolate", "gradient", "gradient_sorted", "hessian", "interpolate", "interpolate_inplace", "lbounds", "out_dtype", "prefilter_", "scale", "ubounds",
    "value_and_gradient",
]

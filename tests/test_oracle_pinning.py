"""Pins the CPU oracle (oracle/) to the reference.

The reference ships no golden vectors for the B-spline path and Julia is not in the image, so the
oracle is pinned by
  (a) the reference's known-answer examples (docs/src/devdocs.md:96-190, src/Interpolations.jl:326-329),
  (b) every property the reference's own tests assert for this path (test/b-splines/cubic.jl,
      quadratic.jl, linear.jl, test/gradient.jl, test/extrapolation/*.jl, test/scaling/*.jl),
  (c) independent SciPy splines that solve the same boundary-value problems.
CPU only (`-m "not gpu"`).
"""
import itertools

import numpy as np
import pytest
from scipy.interpolate import CubicSpline


def bc_list(m):
    return [m.Line, m.Flat, m.Free, m.Periodic]


# ---------------------------------------------------------------------------- (a) known answers
def test_devdocs_linear_known_answer(orc):
    """docs/src/devdocs.md:96-190: value 8.7, the three WeightedAdjIndex weights, gradient (1.0, 3.000000000000001, 9.0)."""
    m = orc.pkg
    A = np.arange(1, 28, dtype=np.float64).reshape(3, 3, 3, order="F")
    itp = orc.interpolate(A, m.BSpline(m.Linear()))
    v, g, bi, _, _ = orc.evaluate(itp, [1.2], [1.4], [1.7], want_grad=True, debug=True)
    assert v[0] == 8.7
    assert tuple(g[0]) == (1.0, 3.000000000000001, 9.0)
    assert tuple(bi[0]) == (1, 1, 1)
    assert orc.weights(1, 1.2, 3)[0] == [0.8, 0.19999999999999996]
    assert orc.weights(1, 1.4, 3)[0] == [0.6000000000000001, 0.3999999999999999]
    assert orc.weights(1, 1.7, 3)[0] == [0.30000000000000004, 0.7]
    assert orc.weights(1, 1.2, 3)[1] == [-1.0, 1.0]


def test_value_weights_doctest(orc):
    """src/Interpolations.jl:326-329: value_weights(Linear(), 0.2) == (0.8, 0.2)."""
    wv, _, _ = orc.weights(1, 1.2 - 1.0 + 1.0, 5)  # x = 1.2 -> dx = 0.19999999999999996
    # the doctest is for dx = 0.2 exactly: evaluate at x = 0.2 + 0 (base 0 is outside 1:n, weights only)
    wv, _, _ = orc.weights(1, 0.2, 5)
    assert wv == [0.8, 0.2]


def test_weights_sum_and_constants(orc):
    """Cubic/quadratic weights: partition of unity, gradient weights sum to 0; constants rounded once in
    the working type (cubic.jl:63-77, quadratic.jl:45-53)."""
    for dt in (np.float32, np.float64):
        eps = np.finfo(dt).eps
        for deg in (2, 3):
            for x in np.linspace(2.0, 3.0, 23):
                wv, wg, _ = orc.weights(deg, x, 10, dt)
                assert abs(sum(wv) - 1) < 8 * eps
                assert abs(sum(wg)) < 16 * eps
    wv, _, base = orc.weights(3, 4.0, 10)  # dx = 0: (1/6, 2/3, 1/6, 0)
    assert base == 3
    assert wv == [1.0 / 6.0, 2.0 / 3.0, 2.0 / 3.0 - 1.0 + 0.5, 0.0]
    wv32, _, _ = orc.weights(3, 4.0, 10, np.float32)
    assert wv32[0] == float(np.float32(1) / np.float32(6))


def test_lu_matches_julia_lu_nopivot(orc):
    """lut! = lu!(Tridiagonal, NoPivot()): dl[i] /= d[i]; d[i+1] -= dl[i]*du[i] (b-splines.jl:257)."""
    dl, d, du, urow, vcol, Cp = orc.prefiltering_system(np.float64, 12, 3, 2, 0)  # Cubic(Line(OnGrid()))
    a, b = 1 / 6, 2 / 3
    D = np.full(12, b)
    L = np.full(11, a)
    U = np.full(11, a)
    D[0] = D[-1] = 1
    U[0] = L[-1] = -2
    for i in range(11):
        L[i] = L[i] / D[i]
        D[i + 1] = D[i + 1] - L[i] * U[i]
    assert np.array_equal(dl, L) and np.array_equal(d, D) and np.array_equal(du, U)
    assert urow == [1, 12] and vcol == [3, 10]
    # pivots converge to (2+sqrt(3))/6
    assert abs(d[10] - (2 + np.sqrt(3)) / 6) < 1e-10


@pytest.mark.parametrize("deg,bc,gt", [(3, 1, 0), (3, 1, 1), (3, 2, 0), (3, 2, 1), (3, 3, 0), (3, 4, 0),
                                       (2, 1, 0), (2, 1, 1), (2, 2, 0), (2, 3, 1), (2, 4, 0), (2, 5, 0), (2, 5, 1)])
def test_system_solution_satisfies_dense_system(orc, deg, bc, gt):
    """The prefilter must solve M c = v where M is the dense matrix spelled out in cubic.jl:114-264 /
    quadratic.jl:84-189 (tridiagonal + Woodbury entries)."""
    n = 17
    a, b = (1 / 6, 2 / 3) if deg == 3 else (1 / 8, 3 / 4)
    M = np.zeros((n, n))
    for i in range(n):
        M[i, i] = b
        if i:
            M[i, i - 1] = a
        if i < n - 1:
            M[i, i + 1] = a
    rows = {  # (d1, du1, [(col offset from the edge, value)])
        (3, 1, 0): (-1, 0, [(3, 1)]), (3, 1, 1): (-9, 11, [(3, -3), (4, 1)]), (3, 2, 1): (3, -7, [(3, 5), (4, -1)]),
        (3, 2, 0): (1, -2, [(3, 1)]), (3, 3, 0): (-1, 4, [(3, -6), (4, 4), (5, -1)]),
        (2, 1, 1): (-1, 1, []), (2, 5, 1): (-1, 1, []), (2, 1, 0): (-1, 0, [(3, 1)]), (2, 5, 0): (-1, 0, [(3, 1)]),
        (2, 2, 0): (1, -2, [(3, 1)]), (2, 3, 1): (1, -3, [(3, 3), (4, -1)]),
    }
    if bc == 4:
        M[0, n - 1] = a
        M[n - 1, 0] = a
    else:
        d1, u1, extra = rows[(deg, bc, gt)]
        M[0, 0] = M[n - 1, n - 1] = d1
        M[0, 1] = M[n - 1, n - 2] = u1
        for c, v in extra:
            M[0, c - 1] = v
            M[n - 1, n - c] = v
    rng = np.random.default_rng(1)
    v = rng.standard_normal(n)
    x = v.copy()
    m = orc.pkg
    BC = {1: m.Flat, 2: m.Line, 3: m.Free, 4: m.Periodic, 5: m.Reflect}[bc]
    GT = m.OnGrid if gt == 0 else m.OnCell
    Deg = m.Cubic if deg == 3 else m.Quadratic
    orc.prefilter_inplace(x, (n,), m.BSpline(Deg(BC(GT()))))
    assert np.allclose(M @ x, v, rtol=0, atol=1e-13)


# ---------------------------------------------------------------------------- (b) reference test properties
def f1(x, n):
    return np.sin((x - 3) * 2 * np.pi / (n - 1) - 1)


@pytest.mark.parametrize("deg", ["Cubic", "Quadratic"])
@pytest.mark.parametrize("gt", ["OnGrid", "OnCell"])
def test_1d_battery(orc, deg, gt):
    """test/b-splines/cubic.jl:15-37 and quadratic.jl:7-39: bounds, node reproduction, OOB throws, interior
    within 10 % of the analytic function."""
    m = orc.pkg
    n = 10
    A = f1(np.arange(1, n + 1), n)
    GT = getattr(m, gt)
    bcs = bc_list(m) + ([m.Reflect] if deg == "Quadratic" else [])
    for BC in bcs:
        itp = orc.interpolate(A, m.BSpline(getattr(m, deg)(BC(GT()))))
        lb, ub = itp.bounds()[0]
        isoncell = gt == "OnCell" and True
        assert (lb, ub) == ((0.5, n + 0.5) if isoncell else (1, n))
        v = orc.evaluate(itp, np.arange(1.0, n + 1))[0]
        assert np.allclose(v, A, rtol=1.5e-8, atol=1e-14)
        for bad in (lb - 0.1, ub + 0.1, lb - 1, ub + 1):
            with pytest.raises(m.BoundsError):
                orc.evaluate(itp, [bad])
        assert np.all(np.isfinite(orc.evaluate(itp, [lb + 0.1, ub - 0.1, lb, ub])[0]))
        xs = np.arange(3.1, n - 3, 0.1)
        assert np.allclose(orc.evaluate(itp, xs)[0], f1(xs, n), atol=np.abs(0.1 * f1(xs, n)).max())


def test_2d_battery_and_periodic_oncell(orc):
    """test/b-splines/cubic.jl:52-66,76-79 (2-D 30x10)."""
    m = orc.pkg

    def f2(x, y):
        return np.sin(x / 10) * np.cos(y / 6)

    xmax, ymax = 30, 10
    X, Y = np.meshgrid(np.arange(1, xmax + 1), np.arange(1, ymax + 1), indexing="ij")
    A = f2(X, Y)
    for BC, GT in itertools.product(bc_list(m), (m.OnGrid, m.OnCell)):
        itp = orc.interpolate(A, m.BSpline(m.Cubic(BC(GT()))))
        v = orc.evaluate(itp, X.astype(float), Y.astype(float))[0]
        assert np.allclose(v, A, rtol=1.5e-8, atol=1e-14)
        xs, ys = np.meshgrid(np.arange(3.1, xmax - 3, 0.53), np.arange(3.1, ymax - 3, 0.29), indexing="ij")
        assert np.allclose(orc.evaluate(itp, xs, ys)[0], f2(xs, ys), atol=np.abs(0.1 * f2(xs, ys)).max())
    itp = orc.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnCell()))))
    (l1, u1), (l2, u2) = itp.bounds()
    ys = np.arange(1.0, ymax + 1)
    assert np.allclose(orc.evaluate(itp, np.full_like(ys, l1), ys)[0], orc.evaluate(itp, np.full_like(ys, u1), ys)[0])
    xs = np.arange(1.0, xmax + 1)
    assert np.allclose(orc.evaluate(itp, xs, np.full_like(xs, l2))[0], orc.evaluate(itp, xs, np.full_like(xs, u2))[0])


def test_flat_boundary_gradient_zero(orc):
    """test/b-splines/cubic.jl:81-87: Flat => gradient ~ 0 (atol eps) at both bounds."""
    m = orc.pkg
    n = 10
    A = f1(np.arange(1, n + 1), n)
    for GT in (m.OnGrid, m.OnCell):
        for Deg in (m.Cubic, m.Quadratic):
            itp = orc.interpolate(A, m.BSpline(Deg(m.Flat(GT()))))
            lb, ub = itp.bounds()[0]
            g = orc.evaluate(itp, [lb, ub], want_value=False, want_grad=True)[1]
            assert np.all(np.abs(g) <= 4 * np.finfo(float).eps), (Deg, GT, g)


def test_free_bc_exact_on_polynomials(orc):
    """test/b-splines/cubic.jl:95-109 (Cubic Free reproduces cubics) and test/gradient.jl:95-116
    (Quadratic Free reproduces quadratics), 1-D and 2-D."""
    m = orc.pkg
    xs = np.arange(1.0, 16)
    p3 = lambda x: 1.5 + 0.3 * x - 0.07 * x**2 + 0.011 * x**3
    for GT in (m.OnGrid, m.OnCell):
        itp = orc.interpolate(p3(xs), m.BSpline(m.Cubic(m.Free(GT()))))
        lb, ub = itp.bounds()[0]
        q = np.linspace(lb, ub, 201)
        assert np.allclose(orc.evaluate(itp, q)[0], p3(q), rtol=1e-11, atol=1e-11)
        p2 = lambda x: 0.3 - 1.1 * x + 0.21 * x**2
        itp = orc.interpolate(p2(xs), m.BSpline(m.Quadratic(m.Free(GT()))))
        v, g, *_ = orc.evaluate(itp, q, want_grad=True)
        assert np.allclose(v, p2(q), rtol=1e-11, atol=1e-11)
        assert np.allclose(g[:, 0], -1.1 + 0.42 * q, rtol=1e-10, atol=1e-10)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    itp = orc.interpolate(p3(X) * p3(Y), m.BSpline(m.Cubic(m.Free(m.OnGrid()))))
    qx, qy = np.meshgrid(np.linspace(1, 15, 37), np.linspace(1, 15, 41), indexing="ij")
    assert np.allclose(orc.evaluate(itp, qx, qy)[0], p3(qx) * p3(qy), rtol=1e-10, atol=1e-9)


def test_gradient_matches_finite_differences(orc):
    """test/gradient.jl:24-41 (check_gradient vs ForwardDiff): here vs central differences of the oracle's
    own value, for Linear, Quadratic x 6 BC x 2, Cubic x 4 BC x 2 in 2-D."""
    m = orc.pkg
    rng = np.random.default_rng(3)
    A = rng.standard_normal((9, 8))
    specs = [m.BSpline(m.Linear())]
    for GT in (m.OnGrid, m.OnCell):
        specs += [m.BSpline(m.Quadratic(BC(GT()))) for BC in (m.Flat, m.Line, m.Free, m.Periodic, m.Reflect)]
        specs += [m.BSpline(m.Cubic(BC(GT()))) for BC in (m.Flat, m.Line, m.Free, m.Periodic)]
    qx, qy = np.meshgrid(np.arange(1.4, 8.5, 1 / 3), np.arange(1.4, 7.5, 1 / 3), indexing="ij")
    h = 1e-6
    for it in specs:
        itp = orc.interpolate(A, it)
        g = orc.evaluate(itp, qx, qy, want_value=False, want_grad=True)[1]
        fx = (orc.evaluate(itp, qx + h, qy)[0] - orc.evaluate(itp, qx - h, qy)[0]) / (2 * h)
        fy = (orc.evaluate(itp, qx, qy + h)[0] - orc.evaluate(itp, qx, qy - h)[0]) / (2 * h)
        if it.degree.code == 1:  # piecewise linear: skip points within h of a kink
            ok = (np.abs(qx - np.round(qx)) > 1e-3) & (np.abs(qy - np.round(qy)) > 1e-3)
        else:
            ok = np.ones_like(qx, bool)
        assert np.allclose(g[..., 0][ok], fx[ok], rtol=1e-5, atol=1e-6), it
        assert np.allclose(g[..., 1][ok], fy[ok], rtol=1e-5, atol=1e-6), it


def test_linear_periodic(orc):
    """test/b-splines/linear.jl:65-98: Linear(Periodic()) has bounds [0.5, n+0.5] and wraps."""
    m = orc.pkg
    n = 10
    A = f1(np.arange(1, n + 1), n)
    itp = orc.interpolate(A, m.BSpline(m.Linear(m.Periodic())))
    assert itp.bounds()[0] == (0.5, n + 0.5)
    v = orc.evaluate(itp, [0.5, n + 0.5, 0.75, n + 0.25])[0]
    assert v[0] == pytest.approx(0.5 * A[-1] + 0.5 * A[0]) and v[1] == pytest.approx(v[0])
    assert v[2] == pytest.approx(0.25 * A[-1] + 0.75 * A[0])
    assert v[3] == pytest.approx(0.75 * A[-1] + 0.25 * A[0])


def test_extrapolation_reference_cases(orc):
    """test/extrapolation/runtests.jl:13-62."""
    m = orc.pkg
    f = lambda x: np.sin((x - 3) * 2 * np.pi / 9 - 1)
    xmax = 10
    A = f(np.arange(1.0, xmax + 1))
    itpg = orc.interpolate(A, m.BSpline(m.Linear()))
    etpg = m.extrapolate(itpg, m.Flat())
    v = orc.evaluate(etpg, [-3, -4.5, 0.9, 1.0, 10.1, 11, 148.298452])[0]
    assert np.all(v[:4] == A[0]) and np.all(v[4:] == A[-1])
    etpf = m.extrapolate(itpg, float("nan"))
    v = orc.evaluate(etpf, [-2.5, 0.999, 1, 10, 10.001])[0]
    assert np.isnan(v[[0, 1, 4]]).all() and v[2] == pytest.approx(f(1)) and v[3] == pytest.approx(f(10))
    etpl = m.extrapolate(itpg, m.Line())
    k_lo, k_hi = A[1] - A[0], A[-1] - A[-2]
    v = orc.evaluate(etpl, [-3.2, xmax + 5.7])[0]
    assert v[0] == pytest.approx(A[0] + k_lo * (-3.2 - 1)) and v[1] == pytest.approx(A[-1] + k_hi * 5.7)
    # 2-D, per-dimension and per-side tuples
    g = lambda x, y: (x**2 + 3 * x - 8) * (-2 * y**2 + y + 1)
    X, Y = np.meshgrid(np.arange(1.0, 9), np.arange(1.0, 9), indexing="ij")
    itp2g = orc.interpolate(g(X, Y), (m.BSpline(m.Quadratic(m.Free(m.OnGrid()))), m.BSpline(m.Linear())))
    ev = lambda itp, x, y: orc.evaluate(itp, [x], [y])[0][0]
    gr = lambda itp, x, y: orc.evaluate(itp, [x], [y], want_value=False, want_grad=True)[1][0]
    etp2g = m.extrapolate(itp2g, (m.Line(), m.Flat()))
    assert ev(etp2g, -0.5, 4) == pytest.approx(ev(itp2g, 1, 4) - 1.5 * gr(itp2g, 1, 4)[0])
    assert ev(etp2g, 5, 100) == pytest.approx(ev(itp2g, 5, 8))
    etp2ud = m.extrapolate(itp2g, ((m.Line(), m.Flat()), m.Flat()))
    assert ev(etp2ud, -0.5, 4) == pytest.approx(ev(itp2g, 1, 4) - 1.5 * gr(itp2g, 1, 4)[0])
    assert ev(etp2ud, 5, -4) == ev(etp2ud, 5, 1)
    assert ev(etp2ud, 100, 4) == ev(etp2ud, 8, 4)
    assert ev(etp2ud, -0.5, 100) == ev(itp2g, 1, 8) - 1.5 * gr(itp2g, 1, 8)[0]
    etp2ll = m.extrapolate(itp2g, m.Line())
    assert ev(etp2ll, -0.5, 100) == pytest.approx(
        (ev(itp2g, 1, 8) - 1.5 * gr(itp2g, 1, 8)[0]) + (100 - 8) * gr(itp2g, 1, 8)[1])
    # Throw in one dimension (runtests.jl:87-92 without the NoInterp axis)
    itpe = m.extrapolate(itp2g, (m.Line(), m.Throw()))
    assert np.isfinite(ev(itpe, 10.1, 1))
    with pytest.raises(m.BoundsError):
        ev(itpe, 5.0, 0.0)


def test_periodic_extrapolation(orc):
    """test/extrapolation/periodic.jl:2-22 (degrees 1-3, OnCell) and Reflect symmetry."""
    m = orc.pkg
    xmax = 7
    f = lambda x: np.sin((x - 3) * 2 * np.pi / xmax - 1)
    A = f(np.arange(1.0, xmax + 1))
    xs = np.arange(-xmax, 2 * xmax, 0.4)
    for it, atol in ((m.BSpline(m.Linear(m.Periodic())), 0.1), (m.BSpline(m.Quadratic(m.Periodic(m.OnCell()))), 0.01),
                     (m.BSpline(m.Cubic(m.Periodic(m.OnCell()))), 0.003)):
        etp = m.extrapolate(orc.interpolate(A, it), m.Periodic())
        assert np.allclose(orc.evaluate(etp, xs)[0], f(xs), rtol=0, atol=atol)
    itp = orc.interpolate(A, m.BSpline(m.Cubic(m.Line(m.OnGrid()))))
    etp = m.extrapolate(itp, m.Reflect())
    assert np.allclose(orc.evaluate(etp, 7 + np.arange(0.1, 5, 0.3))[0], orc.evaluate(itp, 7 - np.arange(0.1, 5, 0.3))[0])
    assert np.allclose(orc.evaluate(etp, 1 - np.arange(0.1, 5, 0.3))[0], orc.evaluate(itp, 1 + np.arange(0.1, 5, 0.3))[0])


def test_extrapolated_gradients(orc):
    """test/gradient.jl:208-244 flavour: Flat gives zero OOB gradient, Line keeps the edge gradient."""
    m = orc.pkg
    A = np.sin(np.arange(1.0, 11) / 2)
    itp = orc.interpolate(A, m.BSpline(m.Cubic(m.Line(m.OnGrid()))))
    gedge = orc.evaluate(itp, [1.0, 10.0], want_value=False, want_grad=True)[1][:, 0]
    g = orc.evaluate(m.extrapolate(itp, m.Flat()), [-2.0, 0.5, 10.5, 5.0], want_value=False, want_grad=True)[1][:, 0]
    assert np.all(g[:3] == 0) and g[3] != 0
    g = orc.evaluate(m.extrapolate(itp, m.Line()), [-2.0, 12.0], want_value=False, want_grad=True)[1][:, 0]
    assert np.array_equal(g, gedge)
    g = orc.evaluate(m.extrapolate(itp, 0.0), [-2.0, 12.0, 5.0], want_value=False, want_grad=True)[1][:, 0]
    assert g[0] == 0 and g[1] == 0 and g[2] != 0


def test_scaling_reference_cases(orc):
    """test/scaling/scaling.jl:8-61 and withextrap.jl:3-41 (scale then extrapolate, +-Inf)."""
    m = orc.pkg
    # scaling.jl:8-20: itp = interpolate(1:1.0:10, Linear); sitp = scale(itp, -3:.5:1.5)
    itp = orc.interpolate(np.arange(1.0, 11), m.BSpline(m.Linear()))
    sitp = m.scale(itp, m.jrange(-3.0, 1.5, step=0.5))
    assert orc.evaluate(sitp, [-3.0, -2.5, 0.0, 1.5, -2.75])[0].tolist() == [1.0, 2.0, 7.0, 10.0, 1.5]
    for bad in (-3.1, 1.6):
        with pytest.raises(m.BoundsError):
            orc.evaluate(sitp, [bad])
    # 2-D quadratic reproduces samples on the scaled grid (scaling.jl:45-50)
    tf = lambda x, y: np.sin(x) * y**2
    rx, ry = m.jrange(-5.0, 5.0, step=0.5), m.jrange(-4.0, 4.0, step=0.2)
    X, Y = np.meshgrid(rx.collect(), ry.collect(), indexing="ij")
    sitp2 = m.scale(orc.interpolate(tf(X, Y), m.BSpline(m.Quadratic(m.Flat(m.OnGrid())))), rx, ry)
    assert np.allclose(orc.evaluate(sitp2, X, Y)[0], tf(X, Y), rtol=1e-8, atol=1e-12)
    # scaled gradient (scaling.jl:52-61)
    r = m.jrange(-np.pi, np.pi, length=10000)
    sitp = m.scale(orc.interpolate(np.sin(r.collect()), m.BSpline(m.Cubic(m.Line(m.OnGrid())))), r)
    xq = np.arange(-3.0, 3.0, 0.37)
    g = orc.evaluate(sitp, xq, want_value=False, want_grad=True)[1][:, 0]
    assert np.allclose(g, np.cos(xq), atol=1e-7)
    # withextrap.jl: scale then extrapolate Flat, OnGrid and OnCell, including +-Inf
    xs = m.jrange(-5.0, 5.0, length=10)
    ys = np.sin(xs.collect())
    for GT in (m.OnGrid, m.OnCell):
        itp = orc.interpolate(ys, m.BSpline(m.Quadratic(m.Flat(GT()))))
        esitp = m.extrapolate(m.scale(itp, xs), m.Flat())
        lb, ub = itp.bounds()[0]
        lo, hi = orc.evaluate(itp, [float(lb), float(ub)])[0]
        half = (xs.collect()[1] - xs.collect()[0]) / 2
        far = [5 + half, 5 + 1.1 * half, 15.8, np.inf] if GT is m.OnCell else [5.0, 5.1, 15.8, np.inf]
        assert np.all(orc.evaluate(esitp, np.array(far))[0] == hi)
        assert np.all(orc.evaluate(esitp, -np.array(far))[0] == lo)


def test_float32_stays_float32(orc):
    """test/typing.jl:10-27 and test/scaling/scaling.jl:96-97: Float32 data + Float32 coordinates (OnGrid,
    Float32 ranges) give Float32; an OnCell bound (Float64, b-splines.jl:152-153) widens under scale."""
    m = orc.pkg
    from interpolations_jl_b200.core import out_dtype_code
    A = np.arange(1, 11, dtype=np.float32) ** 2
    itp = orc.interpolate(A, m.BSpline(m.Quadratic(m.Flat(m.OnGrid()))))
    assert itp.coefs.dtype == np.float32
    assert out_dtype_code(itp.descriptor(np.float32)) == 0
    assert out_dtype_code(itp.descriptor(np.float64)) == 1
    r32 = m.jrange(np.float32(-5), np.float32(4), step=np.float32(1))
    assert out_dtype_code(m.scale(itp, r32).descriptor(np.float32)) == 0
    assert out_dtype_code(m.scale(itp, m.jrange(-5.0, 4.0, step=1.0)).descriptor(np.float32)) == 1
    itpc = orc.interpolate(A, m.BSpline(m.Quadratic(m.Flat(m.OnCell()))))
    assert out_dtype_code(itpc.descriptor(np.float32)) == 0            # bare: no clamp, stays Float32
    assert out_dtype_code(m.scale(itpc, r32).descriptor(np.float32)) == 1  # clamp(x, 0.5, n+0.5) is Float64
    v = orc.evaluate(itp, np.array([3.5], np.float32))[0][0]
    assert v == float(np.float32(v))


# ---------------------------------------------------------------------------- (c) independent splines
@pytest.mark.parametrize("bc,scipy_bc", [("Line", "natural"), ("Flat", "clamped"), ("Free", "not-a-knot"),
                                         ("Periodic", "periodic")])
def test_cubic_ongrid_equals_scipy_cubicspline(orc, bc, scipy_bc):
    """A cubic B-spline interpolant with OnGrid Line/Flat/Free/Periodic conditions is the natural / clamped /
    not-a-knot / periodic cubic spline, which SciPy computes independently (different algorithm)."""
    m = orc.pkg
    rng = np.random.default_rng(11)
    n = 24
    y = rng.standard_normal(n)
    itp = orc.interpolate(y, m.BSpline(m.Cubic(getattr(m, bc)(m.OnGrid()))))
    if bc == "Periodic":  # period n: node n+1 is node 1 again
        cs = CubicSpline(np.arange(1, n + 2), np.r_[y, y[0]], bc_type="periodic")
    else:
        cs = CubicSpline(np.arange(1, n + 1), y, bc_type=scipy_bc)
    q = np.linspace(1, n, 997)
    v, g, *_ = orc.evaluate(itp, q, want_grad=True)
    assert np.allclose(v, cs(q), rtol=0, atol=2e-13 * np.abs(y).max() * 10)
    assert np.allclose(g[:, 0], cs(q, 1), rtol=0, atol=1e-11)


def test_float32_prefilter_close_to_float64(orc):
    m = orc.pkg
    rng = np.random.default_rng(5)
    A = rng.standard_normal((20, 18, 16))
    for it in (m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), m.BSpline(m.Cubic(m.Flat(m.OnCell()))),
               m.BSpline(m.Quadratic(m.Reflect(m.OnCell())))):
        c64 = orc.prefilter(A, it)
        c32 = orc.prefilter(A.astype(np.float32), it)
        assert c32.dtype == np.float32 and c64.shape == c32.shape
        assert np.abs(c32 - c64).max() <= 2e-6 * np.abs(c64).max()


def test_multithreaded_oracle_is_bit_identical(orc):
    m = orc.pkg
    rng = np.random.default_rng(6)
    A = rng.standard_normal((33, 21, 9))
    it = m.BSpline(m.Cubic(m.Line(m.OnGrid())))
    assert np.array_equal(orc.prefilter(A, it, nthreads=1), orc.prefilter(A, it, nthreads=4))
    assert np.array_equal(orc.prefilter(A, it, nthreads=1), orc.prefilter(A, it, nthreads=3, fast=True))
    itp = orc.interpolate(A, it)
    q = [rng.uniform(1, s, 5000) for s in A.shape]
    a = orc.evaluate(itp, *q, want_grad=True, nthreads=1)
    b = orc.evaluate(itp, *q, want_grad=True, nthreads=4, fast=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_hessian_against_gradient_differences_and_exact_polynomials(orc):
    """`hessian` (indexing.jl:37-41,114-144, cubic.jl:79, quadratic.jl:55, linear.jl:58, scaling.jl:150-160):
    the reference checks it against ForwardDiff (test/hessian.jl via InterpolationTestUtils.jl:134-143); here against
    central differences of the (already pinned) gradient, against exact polynomials under Free boundary conditions
    (test/gradient.jl:95-116 style), and the chain rule of the scaled wrapper."""
    m = orc.pkg
    rng = np.random.default_rng(5)
    A = rng.standard_normal((9, 8, 7))
    eps = 1e-5
    for it in (m.BSpline(m.Cubic(m.Line(m.OnGrid()))), m.BSpline(m.Cubic(m.Periodic(m.OnCell()))),
               m.BSpline(m.Quadratic(m.Free(m.OnCell())))):
        itp = orc.interpolate(A, it)
        q = [rng.uniform(2.2, 5.8, 40) for _ in range(3)]
        H = orc.hessian(itp, *q)
        assert np.array_equal(H, np.swapaxes(H, 1, 2))
        if it.degree.code == 3:  # cubic: C2, central differences of the gradient are accurate everywhere
            for d in range(3):
                qp, qm = [c.copy() for c in q], [c.copy() for c in q]
                qp[d] += eps
                qm[d] -= eps
                gp = orc.evaluate(itp, *qp, want_value=False, want_grad=True)[1]
                gm = orc.evaluate(itp, *qm, want_value=False, want_grad=True)[1]
                assert np.abs((gp - gm) / (2 * eps) - H[:, d, :]).max() < 1e-7
    # exact on quadratics: f(x, y) = 2x^2 - 3xy + 0.5y^2 + x - y  ->  H = [[4, -3], [-3, 1]]
    x = np.arange(1, 11.0)[:, None]
    y = np.arange(1, 13.0)[None, :]
    F = 2 * x**2 - 3 * x * y + 0.5 * y**2 + x - y
    for it in (m.BSpline(m.Quadratic(m.Free(m.OnGrid()))), m.BSpline(m.Cubic(m.Free(m.OnGrid())))):
        itp = orc.interpolate(F, it)
        qs = [rng.uniform(1, 10, 30), rng.uniform(1, 12, 30)]
        H = orc.hessian(itp, *qs)
        assert np.abs(H - np.array([[4.0, -3.0], [-3.0, 1.0]])).max() < 1e-9
        # scaled: h ./ (steps .* steps')
        sitp = m.scale(itp, m.jrange(0.0, 4.5, length=10), m.jrange(-1.0, 4.5, length=12))
        Hs = orc.hessian(sitp, (qs[0] - 1) * 0.5, (qs[1] - 1) * 0.5 - 1.0)
        assert np.abs(Hs - np.array([[4.0, -3.0], [-3.0, 1.0]]) / 0.25).max() < 1e-8
    # Linear: hessian_weights == (0, 0) (Interpolations.jl:361-365 doctest)
    lin = orc.interpolate(A, m.BSpline(m.Linear()))
    Hl = orc.hessian(lin, *[rng.uniform(1, 6, 10) for _ in range(3)])
    assert not np.diagonal(Hl, axis1=1, axis2=2).any()


def test_constant_reference_testset(orc):
    """test/b-splines/constant.jl:20-182: Constant{Nearest|Previous|Next}, Throw(OnGrid()) and Periodic(OnCell()):
    bounds, node reproduction, constancy between the nodes, the periodic wrap of the first/last half cells, zero
    gradients."""
    import interpolations_jl_b200 as m
    rng = np.random.default_rng(11)
    N1 = 10
    A1 = rng.random(N1) * 100
    A2 = rng.random((N1, N1)) * 100
    A3 = rng.random((N1, N1, N1)) * 100
    assert m.Constant() == m.Constant(m.Nearest) == m.Constant(m.Nearest(), m.Throw(m.OnGrid()))
    assert m.Constant(m.Periodic()) == m.Constant(m.Nearest, m.Periodic(m.OnCell()))

    def ev(itp, *x):
        return orc.evaluate(itp, *[np.array([float(c)]) for c in x])[0][0]

    for kind, lo_off, hi_off in ((m.Nearest, 0, 0), (m.Previous, -1, 0), (m.Next, 0, 1)):
        it = m.BSpline(m.Constant(kind))
        itp1, itp2, itp3 = orc.interpolate(A1, it), orc.interpolate(A2, it), orc.interpolate(A3, it)
        assert itp1.lbounds() == (1,) and itp1.ubounds() == (N1,)
        assert np.array_equal(itp1.coefs, A1)  # parent(itp) === itp.coefs: no padding, no prefilter
        for i in range(1, N1 + 1):  # check_inbounds_values
            assert ev(itp1, i) == A1[i - 1]
        for i in range(2, N1):
            assert ev(itp1, i + .3) == A1[i - 1 + (1 if kind is m.Next else 0)]
            assert ev(itp1, i - .3) == A1[i - 1 + (-1 if kind is m.Previous else 0)]
            for j in range(2, N1):
                want = {m.Nearest: A2[i - 1, j - 1], m.Previous: A2[i - 1, j - 2], m.Next: A2[i, j - 1]}[kind]
                assert ev(itp2, i + .4, j - .3) == want
        want3 = {m.Nearest: A3[4, 5, 6], m.Previous: A3[4, 4, 6], m.Next: A3[5, 5, 7]}[kind]
        assert ev(itp3, 5.4, 5.7, 7.1) == ev(itp3, 5.4, 5.7, 7.2) == want3
        with pytest.raises(m.BoundsError):
            ev(itp1, 0.9)
        with pytest.raises(m.BoundsError):
            ev(itp1, N1 + 0.1)
        g = orc.evaluate(itp2, np.array([3.3]), np.array([4.6]), want_grad=True)[1]
        assert np.array_equal(g, np.zeros((1, 2)))

    per = orc.interpolate(A1, m.BSpline(m.Constant(m.Periodic())))
    prev = orc.interpolate(A1, m.BSpline(m.Constant(m.Previous, m.Periodic())))
    nxt = orc.interpolate(A1, m.BSpline(m.Constant(m.Next, m.Periodic())))
    for itp in (per, prev, nxt):
        assert itp.lbounds() == (0.5,) and itp.ubounds() == (N1 + 0.5,)
        for i in range(1, N1 + 1):
            assert ev(itp, i) == A1[i - 1]
    for i in range(2, N1):
        assert A1[i - 1] == ev(per, i + .3) == ev(per, i - .3)
        assert A1[i - 1] == ev(prev, i + .3) == ev(prev, i + .6)
        assert A1[i - 2] == ev(prev, i - .3) == ev(prev, i - .6)
        assert A1[i] == ev(nxt, i + .3) == ev(nxt, i + .6)
        assert A1[i - 1] == ev(nxt, i - .3) == ev(nxt, i - .6)
    assert A1[0] == ev(per, 1 - .3) == ev(per, 1 + .3)
    assert A1[-1] == ev(per, N1 - .3) == ev(per, N1 + .3)
    assert A1[0] == ev(prev, 1 + .3)
    assert A1[-1] == ev(prev, N1 + .3) == ev(prev, 1 - .3)
    assert A1[0] == ev(nxt, 1 - .3) == ev(nxt, N1 + .3)
    assert A1[-1] == ev(nxt, N1 - .3)


def test_constant_axes_do_not_widen_the_result(orc):
    """Constant weights are Ints (constant.jl:106): Float32 samples stay Float32 under Float64 coordinates, and a mixed
    (Constant, Linear) interpolant takes its type from the Linear axis."""
    import interpolations_jl_b200 as m
    A = np.arange(1, 31, dtype=np.float32).reshape(5, 6)
    allc = orc.interpolate(A, m.BSpline(m.Constant()))
    assert m.out_dtype(allc, np.float64) == np.float32
    mixed = orc.interpolate(A, (m.BSpline(m.Constant(m.Previous)), m.BSpline(m.Linear())))
    assert m.out_dtype(mixed, np.float64) == np.float64 and m.out_dtype(mixed, np.float32) == np.float32
    v = orc.evaluate(mixed, np.array([2.7]), np.array([3.25]))[0][0]
    assert v == 0.75 * A[1, 2] + 0.25 * A[1, 3]


def test_nointerp_reference_testset(orc):
    """test/nointerp.jl:1-10, test/gradient.jl:134-160, test/scaling/nointerp.jl:3-21: NoInterp axes index with
    integers (a non-integer is an error), reproduce the samples, carry no gradient component, and keep the other
    axes' interpolation untouched."""
    import interpolations_jl_b200 as m
    a = np.arange(1, 13, dtype=np.float64).reshape(4, 3).T  # reshape(1:12, 3, 4)
    ai = orc.interpolate(a, m.NoInterp())
    ii, jj = np.meshgrid(np.arange(1, 4), np.arange(1, 5), indexing="ij")
    assert np.array_equal(orc.evaluate(ai, ii.astype(np.float64), jj.astype(np.float64))[0], a)
    for bad in ((2.2, 2.0), (2.0, 2.2), (0.0, 1.0), (4.0, 1.0)):
        with pytest.raises(m.BoundsError):
            orc.evaluate(ai, np.array([bad[0]]), np.array([bad[1]]))
    rng = np.random.default_rng(5)
    nx = 10
    A2 = rng.random((nx, nx)) * 100
    for BC in (m.Flat, m.Line, m.Free, m.Periodic, m.Reflect):
        for GT in (m.OnGrid, m.OnCell):
            q = m.BSpline(m.Quadratic(BC(GT())))
            itp_c = orc.interpolate(A2, (m.NoInterp(), q))
            itp_d = orc.interpolate(A2, (q, m.NoInterp()))
            col = [orc.interpolate(A2[i, :], q) for i in range(nx)]   # the 1-D interpolants a NoInterp axis selects
            row = [orc.interpolate(A2[:, j], q) for j in range(nx)]
            for _ in range(6):
                x, y = rng.random(2) * (nx - 2) + 1.5
                ix, iy = int(round(x)), int(round(y))
                v, g = orc.evaluate(itp_c, np.array([float(ix)]), np.array([y]), want_grad=True)[:2]
                v1, g1 = orc.evaluate(col[ix - 1], np.array([y]), want_grad=True)[:2]
                assert g.shape == (1, 1) and v[0] == v1[0] and g[0, 0] == g1[0, 0]
                v, g = orc.evaluate(itp_d, np.array([x]), np.array([float(iy)]), want_grad=True)[:2]
                v1, g1 = orc.evaluate(row[iy - 1], np.array([x]), want_grad=True)[:2]
                assert g.shape == (1, 1) and v[0] == v1[0] and g[0, 0] == g1[0, 0]
    # scaled: the range of a NoInterp axis is the axis itself
    xs = m.jrange(-np.pi, np.pi, length=11)
    xv = np.linspace(-np.pi, np.pi, 11)
    A = np.stack([np.sin(xv), np.cos(xv), np.sin(xv) * np.cos(xv)], axis=1)[:-1]
    xs = m.jrange(float(xv[0]), float(xv[-2]), length=10)
    itp = orc.interpolate(A, (m.BSpline(m.Quadratic(m.Periodic(m.OnGrid()))), m.NoInterp()))
    sitp = m.scale(itp, xs, m.UnitRange(1, 3))
    for k, f in enumerate((np.sin, np.cos, lambda t: np.sin(t) * np.cos(t))):
        v = orc.evaluate(sitp, xv[:-1], np.full(10, k + 1.0))[0]
        assert np.abs(v - f(xv[:-1])).max() < 0.05
    g = orc.evaluate(sitp, np.array([np.pi / 3]), np.array([2.0]), want_grad=True)[1]
    assert g.shape == (1, 1)
    with pytest.raises(ValueError):
        m.scale(itp, xs, m.jrange(0.0, 1.0, length=3))


def test_interpolate_inplace_reference_properties(orc):
    """`interpolate!` (b-splines.jl:236-241; exercised by every `(interpolate!, copy)` loop of test/b-splines/*.jl with
    InterpolationTestUtils.check_axes / check_inbounds_values / check_oob / can_eval_near_boundaries): the axes a
    degree would pad are trimmed to 2:n-1, the samples are reproduced on them, the bounds move with them."""
    import interpolations_jl_b200 as m
    rng = np.random.default_rng(21)
    A1 = rng.random(12) * 10
    A2 = rng.random((11, 9)) * 10
    for deg in (m.Cubic, m.Quadratic):
        for BC in (m.Flat, m.Line, m.Free, m.Periodic) + ((m.Reflect,) if deg is m.Quadratic else ()):
            for GT in (m.OnGrid, m.OnCell):
                it = m.BSpline(deg(BC(GT())))
                itp = orc.interpolate_inplace(A1, it)
                padded = BC is not m.Periodic
                first, last = (2, 11) if padded else (1, 12)
                assert (itp.parentaxes[0].first, itp.parentaxes[0].last) == (first, last)  # check_axes
                half = 0.5 if GT is m.OnCell else 0
                assert itp.lbounds() == (first - half,) and itp.ubounds() == (last + half,)
                nodes = np.arange(first, last + 1, dtype=np.float64)
                assert np.allclose(orc.evaluate(itp, nodes)[0], A1[first - 1:last], rtol=0, atol=1e-12)
                for x in (first - half, last + half):  # can_eval_near_boundaries
                    assert np.isfinite(orc.evaluate(itp, np.array([float(x)]))[0][0])
                for x in (first - half - 0.1, last + half + 0.1):  # check_oob
                    with pytest.raises(m.BoundsError):
                        orc.evaluate(itp, np.array([x]))
                itp2 = orc.interpolate_inplace(A2, (it, m.BSpline(m.Linear())))
                xs, ys = np.meshgrid(nodes[nodes <= 11 - (1 if padded else 0)], np.arange(1.0, 10.0), indexing="ij")
                got = orc.evaluate(itp2, xs, ys)[0]
                assert np.allclose(got, A2[xs.astype(int) - 1, ys.astype(int) - 1], rtol=0, atol=1e-12)
    it = m.BSpline(m.Cubic(m.Line(m.OnGrid())))
    with pytest.raises(ValueError):
        m.scale(orc.interpolate_inplace(A1, it), m.jrange(0.0, 1.0, length=10))
    lin = orc.interpolate_inplace(A1, m.BSpline(m.Linear()))  # nothing to trim
    assert lin.lbounds() == (1,) and lin.ubounds() == (12,) and np.array_equal(lin.coefs, A1)


def test_inplace_boundary_conditions_reference_cases(orc):
    """test/b-splines/quadratic.jl:47-62 and test/gradient.jl:108-131: Quadratic(InPlace(OnCell())) keeps the axes of
    A and reproduces the samples; InPlaceQ is exact for an underlying quadratic (values and gradients)."""
    import interpolations_jl_b200 as m
    A = np.array([np.sin((x - 3) * 2 * np.pi / 9 - 1) for x in range(1, 11)])
    itp1 = orc.interpolate_inplace(A, m.BSpline(m.Quadratic(m.InPlace(m.OnCell()))))
    assert (itp1.parentaxes[0].first, itp1.parentaxes[0].last) == (1, 10) and itp1.size == (10,)
    assert itp1.lbounds() == (0.5,) and itp1.ubounds() == (10.5,)
    assert np.allclose(orc.evaluate(itp1, np.arange(1.0, 11.0))[0], A, rtol=0, atol=1e-13)
    assert np.array_equal(orc.interpolate(A, m.BSpline(m.Quadratic(m.InPlace(m.OnCell())))).coefs, itp1.coefs)  # never padded
    A2 = np.array([[np.sin(x / 10) * np.cos(y / 6) for y in range(1, 11)] for x in range(1, 31)])
    itp2 = orc.interpolate_inplace(A2, m.BSpline(m.Quadratic(m.InPlace(m.OnCell()))))
    xs, ys = np.meshgrid(np.arange(1.0, 31.0), np.arange(1.0, 11.0), indexing="ij")
    assert np.allclose(orc.evaluate(itp2, xs, ys)[0], A2, rtol=0, atol=1e-13)
    assert np.isfinite(orc.evaluate(itp2, np.array([0.5, 30.5]), np.array([0.5, 10.5]))[0]).all()  # clamped outer taps
    p = np.array([(x - 1.75) ** 2 for x in range(1, 8)])
    A = np.outer(p, p)
    iq = orc.interpolate_inplace(A, m.BSpline(m.Quadratic(m.InPlace(m.OnCell()))))
    v, g = orc.evaluate(iq, np.array([4.0, 4.0]), np.array([4.0, 3.0]), want_grad=True)[:2]
    assert np.allclose(v, [(4 - 1.75) ** 4, (4 - 1.75) ** 2 * (3 - 1.75) ** 2])
    assert abs(g[1, 0] - 2 * (4 - 1.75) * (3 - 1.75) ** 2) < 0.03 and abs(g[1, 1] - 2 * (4 - 1.75) ** 2 * (3 - 1.75)) < 0.2
    iq = orc.interpolate_inplace(A, m.BSpline(m.Quadratic(m.InPlaceQ(m.OnCell()))))
    v, g = orc.evaluate(iq, np.array([4.0, 4.0]), np.array([4.0, 3.0]), want_grad=True)[:2]
    assert np.allclose(v, [(4 - 1.75) ** 4, (4 - 1.75) ** 2 * (3 - 1.75) ** 2])
    assert np.allclose(g[1], [2 * (4 - 1.75) * (3 - 1.75) ** 2, 2 * (4 - 1.75) ** 2 * (3 - 1.75)])
    # InPlaceQ: exact between the nodes too, wherever no tap is clamped (1.5 <= x < n - 0.5)
    xq = np.array([1.5, 1.8, 2.3, 3.7, 5.9, 6.4])
    v = orc.evaluate(iq, xq, np.full(6, 2.0))[0]
    assert np.allclose(v, (xq - 1.75) ** 2 * (2 - 1.75) ** 2, rtol=0, atol=1e-12)


# ---------------------------------------------------------------------------- (d) the oracle's descriptor is its own
def test_descriptor_derivations_agree(orc):
    """The oracle derives padding, first coefficient index, bounds and types itself (oracle/describe.py, following
    b-splines.jl:141-158, cubic.jl:86-87, quadratic.jl:64-65, prefiltering.jl:6, scaling.jl:60-75,116); the product
    derives them in core.descriptor().  Field-by-field agreement over every spec / wrapper combination in scope."""
    import describe as dsc
    m = orc.pkg
    specs = []
    for gt in (m.OnGrid, m.OnCell):
        for bc in (m.Line, m.Flat, m.Free, m.Periodic):
            specs.append(m.BSpline(m.Cubic(bc(gt()))))
        for bc in (m.Line, m.Flat, m.Free, m.Periodic, m.Reflect):
            specs.append(m.BSpline(m.Quadratic(bc(gt()))))
    specs += [m.BSpline(m.Quadratic(m.InPlace(m.OnCell()))), m.BSpline(m.Quadratic(m.InPlaceQ(m.OnCell()))),
              m.BSpline(m.Linear()), m.BSpline(m.Linear(m.Periodic(m.OnCell()))), m.BSpline(m.Constant()),
              m.BSpline(m.Constant(m.Previous(), m.Periodic(m.OnCell())))]
    rng = np.random.default_rng(3)
    n_checked = 0
    for it in specs:
        for ndims, dt in ((1, np.float64), (2, np.float32), (3, np.float64)):
            shape = tuple(int(v) for v in rng.integers(6, 11, ndims))
            A = rng.standard_normal(shape).astype(dt)
            itps = [orc.interpolate(A, it)]
            if it.degree.code in (2, 3):
                itps.append(orc.interpolate_inplace(A, it))
            for itp in itps:
                chains = [itp]
                if all(f == 1 for f in itp.parent_first):
                    for rdt in (np.float32, np.float64, None):
                        if rdt is None:
                            ranges = [m.jrange(2, 2 + n - 1) for n in shape]                   # UnitRange
                        else:
                            ranges = [m.jrange(-1.0, 2.5, length=n, dtype=rdt) for n in shape]  # StepRangeLen
                        chains.append(m.scale(itp, *ranges))
                wrapped = list(chains)
                for c in chains:
                    wrapped.append(m.extrapolate(c, m.Line()))
                    wrapped.append(m.extrapolate(c, ((m.Flat(), m.Periodic()),)) if ndims == 1 else
                                   m.extrapolate(c, tuple([m.Reflect()] + [m.Flat()] * (ndims - 1))))
                    wrapped.append(m.extrapolate(c, np.float32(1.5)))
                    wrapped.append(m.extrapolate(c, 0))
                for w in wrapped:
                    for cdt in (np.float32, np.float64):
                        diff = dsc.differences(dsc.describe(w, cdt), w.descriptor(cdt))
                        assert not diff, (it, shape, w, cdt, diff)
                        n_checked += 1
    assert n_checked > 1000


def test_oracle_does_not_use_the_product_descriptor():
    """oracle.py must build its descriptors with oracle/describe.py, never with core.descriptor()."""
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "oracle.py")).read()
    assert ".descriptor(" not in src


def test_specialised_tricubic_evaluator_equals_the_interpreter(orc):
    """bench.py's CPU legs time a type-specialised Float32 tricubic evaluator (what Julia compiles for C3); it must
    return the interpreter's bits, in the parity build and in the -O3 timing build."""
    m = orc.pkg
    rng = np.random.default_rng(5)
    shape = (19, 23, 17)
    A = rng.standard_normal(shape).astype(np.float32)
    itp = orc.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
    q = [rng.uniform(1, n, 20000).astype(np.float32) for n in shape]
    for k, n in enumerate(shape):  # integers, both bounds
        q[k][:50] = rng.integers(1, n + 1, 50).astype(np.float32)
        q[k][50], q[k][51] = 1.0, float(n)
    ref = orc.evaluate(itp, *q)[0].astype(np.float32)
    for fast in (False, True):
        got = orc.evaluate_tricubic_periodic_f32(itp.coefs, shape, *q, nthreads=2, fast=fast)
        assert np.array_equal(got, ref), fast

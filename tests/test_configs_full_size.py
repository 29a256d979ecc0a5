"""BASELINE.json configs at their full sizes on the GPU.  The CPU oracle cannot finish these in seconds, so each
config is checked through size-independent properties of the domain plus an oracle comparison on a sample:

  * node reproduction  itp(i...) == A[i...]  (test/InterpolationTestUtils.jl:52-57) on a strided subset of nodes;
  * M c = v: applying the degree's stencil to the coefficients gives the samples back (prefilter correctness);
  * oracle on the first 2*10^5 queries, on the GPU's coefficients: values, gradients, indices, branches all `==`;
  * the ordered pipeline returns exactly what the direct kernel returns (C2, C3).
Query counts are reduced where the config's count only repeats the same work (stated per test)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sample_check(pkg, orc, obj, itp, coords, k=200_000, grad=True):
    m = pkg
    sub = [c[:k].cpu().numpy() for c in coords]
    host = orc.with_coefs(obj, itp.coefs.cpu().numpy())
    v, _, bi, br = m.evaluate_debug(obj, *sub)
    ov, og, obi, obr, _ = orc.evaluate(host, *sub, want_grad=grad, debug=True, nthreads=orc.num_threads())
    odt = m.out_dtype(obj, sub[0].dtype)
    assert np.array_equal(bi.cpu().numpy(), obi)
    assert np.array_equal(br.cpu().numpy().view(np.uint32), obr)
    assert np.array_equal(v.cpu().numpy(), ov.astype(odt), equal_nan=True)
    if grad:
        assert np.array_equal(m.gradient(obj, *sub).cpu().numpy(), og.astype(odt), equal_nan=True)


def test_c1_cubic_line_ongrid_1e6_f64(pkg, orc):
    """C1: 1-D Cubic(Line(OnGrid())) on 10^6 Float64 samples, 10^7 random queries (full size; the CPU oracle also
    runs this one completely: prefilter compared at 1e-12, all 10^7 values compared on equal coefficients)."""
    import torch
    m = pkg
    rng = np.random.default_rng(11)
    n, nq = 1_000_000, 10_000_000
    A = np.sin(np.arange(1, n + 1) * 1e-3) + 0.1 * rng.standard_normal(n)
    it = m.BSpline(m.Cubic(m.Line(m.OnGrid())))
    itp = m.interpolate(A, it)
    ref = orc.prefilter(A, it, nthreads=orc.num_threads())
    got = itp.coefs_array()
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    x = torch.from_numpy(rng.uniform(1, n, nq)).cuda()
    v = itp(x)
    host = orc.with_coefs(itp, itp.coefs.cpu().numpy())
    ov, *_ = orc.evaluate(host, x.cpu().numpy(), nthreads=orc.num_threads(), fast=False)
    assert np.array_equal(v.cpu().numpy(), ov)
    nodes = torch.arange(1, n + 1, 997, dtype=torch.float64).cuda()
    assert np.allclose(itp(nodes).cpu().numpy(), A[::997], rtol=0, atol=1e-10)


def test_c2_bicubic_flat_oncell_4096_f32(pkg, orc):
    """C2: 4096^2 Float32 image, Cubic(Flat(OnCell())), 10^8 scattered queries + gradient (full size)."""
    import torch
    m = pkg
    n, nq = 4096, 100_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand((n, n), device="cuda", dtype=torch.float32, generator=g)
    it = m.BSpline(m.Cubic(m.Flat(m.OnCell())))
    itp = m.interpolate(A, it)
    assert itp.size == (n + 2, n + 2)
    # M c = v in the interior: (1/6, 2/3, 1/6) along both axes (cubic.jl:101-112)
    c = itp.coefs.reshape(n + 2, n + 2)
    back = c
    for dim in range(2):
        back = (torch.roll(back, 1, dim) + torch.roll(back, -1, dim)) / 6 + back * (2.0 / 3.0)
    err = (back[1:-1, 1:-1] - A.t()).abs().max().item()
    assert err < 5e-6, err
    coords = [torch.rand(nq, device="cuda", dtype=torch.float32, generator=g) * n + 0.5 for _ in range(2)]
    coords = [c_.clamp_(0.5, n + 0.5) for c_ in coords]
    v = m.evaluate_direct(itp, *coords)
    gr = m.gradient(itp, *coords)
    assert torch.isfinite(v).all() and torch.isfinite(gr).all() and gr.shape == (nq, 2)
    assert torch.equal(v, m.evaluate_sorted(itp, *coords))
    assert torch.equal(v, itp(*coords))
    _sample_check(m, orc, itp, itp, coords)
    # node reproduction on a strided subset
    ii = torch.arange(1, n + 1, 61, device="cuda", dtype=torch.float32)
    X, Y = torch.meshgrid(ii, ii, indexing="ij")
    vals = itp(X.reshape(-1), Y.reshape(-1)).reshape(len(ii), len(ii))
    assert (vals - A[::61, ::61]).abs().max().item() < 5e-6


def test_c3_tricubic_periodic_512_f32(pkg, orc):
    """C3: 512^3 Float32 Cubic(Periodic(OnGrid())): prefilter (property test above) + queries.  2*10^8 of the 10^9
    queries (the rest repeats the same kernels on more batches; bench.py runs the full count)."""
    import torch
    m = pkg
    n, nq = 512, 200_000_000
    g = torch.Generator(device="cuda").manual_seed(4)
    A = torch.randn((n, n, n), device="cuda", dtype=torch.float32, generator=g)
    itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
    coords = [(torch.rand(nq, device="cuda", dtype=torch.float32, generator=g) * (n - 1) + 1).clamp_(1, n) for _ in range(3)]
    L = m._lib.lib()
    n0 = L.b200itp_eval_sorted_ordered_calls()
    v = m.evaluate_sorted(itp, *coords)
    assert L.b200itp_eval_sorted_ordered_calls() == n0 + 1, "C3 must be served by the ordered pipeline"
    k = 20_000_000
    assert torch.equal(v[:k], m.evaluate_direct(itp, *[c[:k] for c in coords]))          # ordered pipeline == direct kernel
    assert torch.equal(v[-k:], m.evaluate_direct(itp, *[c[-k:] for c in coords]))
    n2 = L.b200itp_eval_sorted_ordered_calls()
    assert torch.equal(v, itp(*coords)) and L.b200itp_eval_sorted_ordered_calls() == n2 + 1   # the drop-in call takes it too
    _sample_check(m, orc, itp, itp, coords, grad=True)
    # gradients of 5*10^7 queries through the ordered pipeline == the direct kernel's
    kg = 50_000_000
    n1 = L.b200itp_eval_sorted_ordered_calls()
    gs = m.gradient_sorted(itp, *[c[:kg] for c in coords])
    assert L.b200itp_eval_sorted_ordered_calls() == n1 + 1
    kd = 4_000_000
    assert torch.equal(gs[:kd], m.gradient_direct(itp, *[c[:kd] for c in coords]))
    assert torch.equal(gs[-kd:], m.gradient_direct(itp, *[c[kg - kd:kg] for c in coords]))
    del gs
    # checksum of checksums: a permutation of the queries permutes the results
    perm = torch.randperm(1 << 22, device="cuda", generator=g)
    a = m.evaluate_sorted(itp, *[c[: 1 << 22] for c in coords])
    b = m.evaluate_sorted(itp, *[c[: 1 << 22][perm] for c in coords])
    assert torch.equal(a[perm], b)
    ii = torch.arange(1, n + 1, 37, device="cuda", dtype=torch.float32)
    X, Y, Z = torch.meshgrid(ii, ii, ii, indexing="ij")
    vals = itp(X.reshape(-1), Y.reshape(-1), Z.reshape(-1)).reshape(len(ii), len(ii), len(ii))
    assert (vals - A[::37, ::37, ::37]).abs().max().item() < 2e-5


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c4_quadratic_reflect_scaled_line_8192(pkg, orc, dtype):
    """C4: 8192^2 Quadratic(Reflect(OnCell())) with scale(ranges) and extrapolate(Line()) fused; 10^8 queries over
    [-0.1, 1.1]^2 (f32) / 2*10^7 (f64 secondary)."""
    import torch
    m = pkg
    n = 8192
    nq = 100_000_000 if dtype == np.float32 else 20_000_000
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand((n, n), device="cuda", dtype=tdt, generator=g)
    itp = m.interpolate(A, m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))))
    one = dtype(1)
    rngs = [m.jrange(dtype(0), one, length=n) for _ in range(2)]
    obj = m.extrapolate(m.scale(itp, *rngs), m.Line())
    coords = [torch.rand(nq, device="cuda", dtype=tdt, generator=g) * 1.2 - 0.1 for _ in range(2)]
    v = obj(*coords)
    assert torch.isfinite(v).all()
    _sample_check(m, orc, obj, itp, coords, grad=True)
    # node reproduction through the scaled wrapper: sitp(range[i], range[j]) == A[i, j]
    idx = torch.arange(0, n, 211, device="cuda")
    xs = torch.from_numpy(rngs[0].collect()[::211].copy()).cuda()
    X, Y = torch.meshgrid(xs, xs, indexing="ij")
    vals = obj(X.reshape(-1), Y.reshape(-1)).reshape(len(idx), len(idx))
    # Float32 ranges: coordlookup (x - first) / step + 1 carries ~8192 * eps(Float32) ~ 1e-3 cells of rounding, and the
    # random image changes by O(1) per cell (the oracle comparison above is bit-exact; this bounds the property only)
    tol = 2e-3 if dtype == np.float32 else 1e-10
    assert (vals - A[::211, ::211]).abs().max().item() < tol


def test_c5_linear_4d_128_f64(pkg, orc):
    """C5: 4-D Linear() 128^4 Float64 grid (2 GiB, no prefilter); 10^8 of the 10^9 queries (2^4-tap gather)."""
    import torch
    m = pkg
    n, nq = 128, 100_000_000
    g = torch.Generator(device="cuda").manual_seed(6)
    A = torch.rand((n,) * 4, device="cuda", dtype=torch.float64, generator=g)
    itp = m.interpolate(A, m.BSpline(m.Linear()))
    assert itp.size == (n,) * 4
    coords = [(torch.rand(nq, device="cuda", dtype=torch.float64, generator=g) * (n - 1) + 1).clamp_(1, n) for _ in range(4)]
    v = itp(*coords)
    assert torch.isfinite(v).all()
    _sample_check(m, orc, itp, itp, coords, grad=True)
    ii = torch.arange(1, n + 1, 9, device="cuda", dtype=torch.float64)
    W, X, Y, Z = torch.meshgrid(ii, ii, ii, ii, indexing="ij")
    vals = itp(W.reshape(-1), X.reshape(-1), Y.reshape(-1), Z.reshape(-1)).reshape((len(ii),) * 4)
    assert torch.equal(vals, A[::9, ::9, ::9, ::9])  # linear weights at nodes are exactly (1, 0)


@pytest.mark.parametrize("shape", [(530, 530, 530), (12100, 12100)])
def test_ordered_path_beyond_8192_bricks(pkg, shape):
    """Arrays with more than 8192 bricks (512^3 cells / 11585^2 cells) take a third split level (super-groups): the
    ordered pipeline must still serve the call and return exactly the direct kernel's values and gradients."""
    import torch
    m = pkg
    L = m._lib.lib()
    g = torch.Generator(device="cuda").manual_seed(8)
    A = torch.rand(shape, device="cuda", dtype=torch.float32, generator=g)
    it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))) if len(shape) == 3 else m.BSpline(m.Quadratic(m.Reflect(m.OnCell())))
    itp = m.interpolate(A, it)
    del A
    nq = 30_000_000
    coords = []
    for (lb, ub) in itp.bounds():
        lb, ub = float(lb), float(ub)
        coords.append((torch.rand(nq, device="cuda", dtype=torch.float32, generator=g) * (ub - lb) + lb).clamp_(lb, ub))
    n0 = L.b200itp_eval_sorted_ordered_calls()
    v = m.evaluate_sorted(itp, *coords)
    assert L.b200itp_eval_sorted_ordered_calls() == n0 + 1, "the three-level ordered pipeline must serve this call"
    k = 6_000_000
    assert torch.equal(v[:k], m.evaluate_direct(itp, *[c[:k] for c in coords]))
    assert torch.equal(v[-k:], m.evaluate_direct(itp, *[c[-k:] for c in coords]))
    gs = m.gradient_sorted(itp, *[c[:k] for c in coords])
    assert L.b200itp_eval_sorted_ordered_calls() == n0 + 2
    assert torch.equal(gs, m.gradient_direct(itp, *[c[:k] for c in coords]))

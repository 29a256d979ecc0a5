"""C2 (4096^2 bicubic) and C4-like timings: direct vs ordered path."""
import sys, time
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for n, it, lo in ((4096, m.BSpline(m.Cubic(m.Flat(m.OnCell()))), 0.5), (8192, m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))), 0.5),
                  (16384, m.BSpline(m.Cubic(m.Flat(m.OnCell()))), 0.5)):
    nq = 10**8
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand((n, n), device="cuda", generator=g)
    itp = m.interpolate(A, it)
    c = [(torch.rand(nq, device="cuda", generator=g) * n + lo).clamp_(lo, n + lo) for _ in range(2)]
    td = t(lambda: itp(*c)); ts = t(lambda: m.evaluate_sorted(itp, *c)); tg = t(lambda: m.gradient(itp, *c))
    assert torch.equal(itp(*c), m.evaluate_sorted(itp, *c))
    print(f"n={n} {it}: direct {nq/td/1e6:.1f} Gevals/s, ordered {nq/ts/1e6:.1f} Gevals/s, gradient(direct) {nq/tg/1e6:.1f} Gevals/s")

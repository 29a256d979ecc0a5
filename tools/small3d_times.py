"""3-D volumes that fit the L2: direct kernel vs ordered pipeline (threshold lifted)."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
def t(fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
nq = 1 << 27
for n in (64, 128, 256, 320):
    A = torch.randn((n, n, n), device="cuda")
    itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
    g = torch.Generator(device="cuda").manual_seed(2)
    c = [(torch.rand(nq, device="cuda", generator=g) * (n - 1) + 1).clamp(1, n) for _ in range(3)]
    td = t(lambda: itp(*c))
    L.b200itp_eval_sorted_set_min_queries(1)
    n0 = L.b200itp_eval_sorted_ordered_calls()
    ts = t(lambda: m.evaluate_sorted(itp, *c))
    used = L.b200itp_eval_sorted_ordered_calls() - n0
    L.b200itp_eval_sorted_set_min_queries(1 << 20)
    print(f"n={n}: direct {nq/td/1e6:.2f} Gevals/s, ordered {nq/ts/1e6:.2f} Gevals/s (ordered calls {used})")

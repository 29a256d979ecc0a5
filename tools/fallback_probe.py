import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200_loader
m = b200_loader.load(); L = m._lib.lib()
n = 512; nq = 1 << 28
A = torch.randn((n, n, n), device="cuda", dtype=torch.float32)
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
g = torch.Generator(device="cuda").manual_seed(2)
coords = [(torch.rand(nq, device="cuda", dtype=torch.float32, generator=g) * (n - 1) + 1).clamp(1, n) for _ in range(3)]
def timed(label, fn):
    c0 = L.b200itp_eval_sorted_ordered_calls()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); out = fn(); b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{label}: {a.elapsed_time(b):.2f} ms (wall {(t1-t0)*1e3:.2f}), ordered calls +{L.b200itp_eval_sorted_ordered_calls()-c0}", flush=True)
    return out
out = timed("sorted #1", lambda: m.evaluate_sorted(itp, *coords))
out = timed("sorted #2", lambda: m.evaluate_sorted(itp, *coords))
k = 1 << 22
ref = timed("direct 2^22", lambda: m.evaluate_direct(itp, *[c[:k] for c in coords]))
out = timed("sorted #3", lambda: m.evaluate_sorted(itp, *coords))
out = timed("sorted #4", lambda: m.evaluate_sorted(itp, *coords))
out = timed("call   #5", lambda: itp(*coords))

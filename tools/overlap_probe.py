"""Does the ordered pipeline gain from running sub-batches of the queries on concurrent streams (issue-bound evaluation
of one batch under the memory-bound passes of another)?  K host threads, one CUDA stream each, contiguous query shares."""
import sys, time, threading
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
n = 512
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10**9
A = torch.randn((n, n, n), device="cuda")
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
g = torch.Generator(device="cuda").manual_seed(2)
c = [(torch.rand(nq, device="cuda", generator=g) * (n - 1) + 1).clamp(1, n) for _ in range(3)]
ref = m.evaluate_sorted(itp, *c)
torch.cuda.synchronize()


def run(K):
    outs = [None] * K
    streams = [torch.cuda.Stream() for _ in range(K)]
    bounds = [nq * i // K for i in range(K + 1)]

    def work(i):
        with torch.cuda.stream(streams[i]):
            outs[i] = m.evaluate_sorted(itp, *[x[bounds[i]:bounds[i + 1]] for x in c])

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = all(torch.equal(outs[i], ref[bounds[i]:bounds[i + 1]]) for i in range(K))
    return dt * 1e3, ok


for K in (1, 2, 3, 4, 8):
    ts = [run(K) for _ in range(3)]
    print(f"K={K}: {min(t for t, _ in ts):.2f} ms  {nq / min(t for t, _ in ts) / 1e6:.2f} Gevals/s  identical={all(o for _, o in ts)}", flush=True)

"""Single-process multi-GPU prefilter through the C ABI (b200itp_prefilter_sharded): device times for 512^3 and 1024^3
Float32 against the single-device prefilter + b200itp_bcast_coefs."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
G = torch.cuda.device_count()
comm = m.parallel.Comm(list(range(G)))
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
for n in (512, 1024):
    A = torch.randn((n, n, n), device="cuda:0")
    slabs = []
    for r in range(G):
        lo, hi = m.parallel.shard_bounds(n, G, r)
        slabs.append(A[lo:hi].to(f"cuda:{r}").contiguous())
    best = None
    for _ in range(5):
        itps, ms = comm.interpolate_sharded(slabs, it)
        best = ms if best is None else (min(best[0], ms[0]), min(best[1], ms[1]))
    # single device: prefilter + broadcast
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    bufs = [torch.empty(n ** 3, device=f"cuda:{r}") for r in range(G)]
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        a.record()
        ref = m.interpolate(A.permute(2, 1, 0), it, device="cuda:0")
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    import time
    bufs[0] = ref.coefs
    tb = []
    for _ in range(4):
        t0 = time.perf_counter(); comm.bcast_coefs(bufs, root=0); tb.append((time.perf_counter() - t0) * 1e3)
    ok = all(torch.equal(itp.coefs.to("cuda:0"), ref.coefs) for itp in itps)
    alg = 2 * 4 * n ** 3 * 3
    print(f"n={n} G={G}: sharded passes+a2a {best[0]:.3f} ms ({alg / best[0] / 1e6:.0f} GB/s), with replication {best[1]:.3f} ms | "
          f"one device {min(ts):.3f} ms + bcast {min(tb):.3f} ms (host-timed) | identical={ok}", flush=True)
    del A, slabs, itps, bufs, ref
    torch.cuda.empty_cache()
comm.close()

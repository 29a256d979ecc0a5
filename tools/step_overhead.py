"""Where does a C3 step spend time besides its kernels?  One GPU, the query share of a rank at N = 8 (1.25e8) and the
whole 1e9: step time by CUDA events, the library's stage sum, host time per step and the Python profile."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_loader
m = b200_loader.load()
L = m._lib.lib()
dev = torch.device("cuda", 0)
n = 512
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
g = torch.Generator(device=dev).manual_seed(1)
A = torch.rand((n, n, n), device=dev, dtype=torch.float32, generator=g)
side = torch.cuda.Stream(device=dev)

def make_step(coords, overlap=True):
    def step():
        main = torch.cuda.current_stream()
        if not overlap:
            itp = m.interpolate(A, it, device=dev)
            return itp(*coords)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            coefs = m.interpolate(A, it, device=dev).coefs
            ready = torch.cuda.Event(); ready.record(side)
        coefs.record_stream(main)
        itp = m.BSplineInterpolation(coefs, (n,) * 3, (n,) * 3, (it,) * 3, 0)
        itp.ready_event = ready
        return itp(*coords)
    return step

for nq in (125_000_000, 1_000_000_000):
    coords = [torch.rand(nq, device=dev, dtype=torch.float32, generator=g) * (n - 1) + 1 for _ in range(3)]
    for overlap in (True, False):
        step = make_step(coords, overlap)
        for _ in range(3): out = step()
        torch.cuda.synchronize()
        K = 10
        L.b200itp_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(K): out = step()
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        L.b200itp_profile_enable(0)
        st = m._lib.profile_stages()
        ssum = sum(float(np.sum(v)) for k, v in st.items() if k.startswith("ord_")) / K
        psum = sum(float(np.sum(v)) for k, v in st.items() if not k.startswith("ord_")) / K
        print(f"nq {nq:.2e} overlap={overlap}: step {e0.elapsed_time(e1)/K:.3f} ms (wall {(t1-t0)/K*1e3:.3f}), ord stage sum {ssum:.3f}, prefilter stage sum {psum:.3f}")
        print("    ", {k: round(float(np.mean(v)), 3) for k, v in st.items()})
        del out
    if nq == 125_000_000:
        # host time of one step with the GPU idle at the start (what a step costs the CPU)
        step = make_step(coords, True)
        torch.cuda.synchronize()
        pr = cProfile.Profile(); pr.enable()
        for _ in range(20): out = step()
        pr.disable(); torch.cuda.synchronize()
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
    del coords
    m.core.release_scratch(); torch.cuda.empty_cache()

"""A/B of the tensor-map tile copy of ord_eval_kernel in ONE process (b200itp_eval_sorted_set_tma): 1.25e8 and 1e9
queries on the C3 volume."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_loader
m = b200_loader.load(); L = m._lib.lib()
dev = torch.device("cuda", 0); n = 512
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
g = torch.Generator(device=dev).manual_seed(1)
A = torch.rand((n, n, n), device=dev, dtype=torch.float32, generator=g)
itp = m.interpolate(A, it, device=dev)
for nq in (125_000_000, 1_000_000_000):
    coords = [torch.rand(nq, device=dev, dtype=torch.float32, generator=g) * (n - 1) + 1 for _ in range(3)]
    for on in (1, 0, 1, 0):
        L.b200itp_eval_sorted_set_tma(on)
        for _ in range(2): out = itp(*coords)
        torch.cuda.synchronize(); L.b200itp_profile_enable(1)
        for _ in range(5): out = itp(*coords)
        torch.cuda.synchronize(); L.b200itp_profile_enable(0)
        st = m._lib.profile_stages()
        print(f"nq {nq:.2e} tma={on}: ord_eval {np.mean(st['ord_eval']):.3f} ms (min {np.min(st['ord_eval']):.3f})  stage sum {sum(float(np.sum(v)) for v in st.values())/5:.3f}")
    del coords, out
    m.core.release_scratch(); torch.cuda.empty_cache()

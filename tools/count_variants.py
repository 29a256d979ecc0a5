"""C3 ordered evaluation, 1e9 queries: ord_count / ord_split1 stage times and the stage sum for one build of the library
(B200ITP_LIB selects the build: sweeps of ORD_CG_* and ORD_FINE_SHARE)."""
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
nq = 1_000_000_000
coords = [torch.rand(nq, device=dev, dtype=torch.float32, generator=g) * (n - 1) + 1 for _ in range(3)]
for _ in range(2): out = itp(*coords)
torch.cuda.synchronize(); L.b200itp_profile_enable(1)
for _ in range(4): out = itp(*coords)
torch.cuda.synchronize(); L.b200itp_profile_enable(0)
st = m._lib.profile_stages()
print(os.environ.get("B200ITP_LIB", "default"), {k: round(float(np.mean(v)), 3) for k, v in st.items() if k in ("ord_count", "ord_split1")}, "sum", round(sum(float(np.sum(v)) for k, v in st.items()) / 4, 3))

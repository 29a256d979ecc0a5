"""Driver for ncu: C3 prefilter (512^3 Float32, Cubic(Periodic(OnGrid()))), a few repetitions."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
n = 512
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
torch.manual_seed(0)
A = torch.randn((n, n, n), device="cuda")
work = A.reshape(-1).clone()
for _ in range(3):
    m.prefilter_(work, (n, n, n), (it,) * 3, 0)
torch.cuda.synchronize()
L.b200itp_profile_enable(1)
for _ in range(10):
    m.prefilter_(work, (n, n, n), (it,) * 3, 0)
torch.cuda.synchronize()
L.b200itp_profile_enable(0)
st = m._lib.profile_stages()
print({k: round(sum(v) / len(v), 4) for k, v in st.items()}, "total", round(sum(sum(v) / len(v) for v in st.values()), 4))
chk = A.reshape(-1).clone()
m.prefilter_(chk, (n, n, n), (it,) * 3, 0)
torch.cuda.synchronize()
print("checksum", float(chk.double().abs().sum()), float(chk[12345678]), float(chk[-1]))

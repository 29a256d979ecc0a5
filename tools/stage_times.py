"""Stage times (b200itp_profile_*) of one C3 ordered evaluation call."""
import sys
sys.path.insert(0,".")
import b200_loader, torch
m=b200_loader.load(); L=m._lib.lib()
n=512; nq=int(sys.argv[1]) if len(sys.argv)>1 else 10**9
A=torch.randn((n,n,n),device="cuda"); itp=m.interpolate(A,m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
g=torch.Generator(device="cuda").manual_seed(2)
c=[(torch.rand(nq,device="cuda",generator=g)*(n-1)+1).clamp(1,n) for _ in range(3)]
out=m.evaluate_sorted(itp,*c); torch.cuda.synchronize()
k=1<<22
assert torch.equal(out[:k], m.evaluate_direct(itp, *[x[:k] for x in c]))
L.b200itp_profile_enable(1); m.evaluate_sorted(itp,*c); m.evaluate_sorted(itp,*c); torch.cuda.synchronize(); L.b200itp_profile_enable(0)
st=m._lib.profile_stages()
print({k:round(min(v),3) for k,v in st.items()}, 'total', round(sum(min(v) for v in st.values()),2))

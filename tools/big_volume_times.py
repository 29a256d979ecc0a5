"""Ordered evaluation on a 1024^3 Float32 volume (4 GiB, 32768 bricks: three split levels), 1e9 queries."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10**9
A = torch.randn((n, n, n), device="cuda"); itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))); del A
g = torch.Generator(device="cuda").manual_seed(2)
c = [(torch.rand(nq, device="cuda", generator=g) * (n - 1) + 1).clamp(1, n) for _ in range(3)]
out = m.evaluate_sorted(itp, *c); torch.cuda.synchronize()
k = 1 << 22
assert torch.equal(out[:k], itp(*[x[:k] for x in c]))
L.b200itp_profile_enable(1); m.evaluate_sorted(itp, *c); torch.cuda.synchronize(); L.b200itp_profile_enable(0)
st = m._lib.profile_stages()
tot = sum(min(v) for v in st.values())
print(f"n={n}", {k_: round(min(v), 3) for k_, v in st.items()}, "total", round(tot, 2), "Gevals/s", round(nq / tot / 1e6, 2))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
kd = 1 << 26
a.record(); itp(*[x[:kd] for x in c]); b.record(); torch.cuda.synchronize()
print("direct kernel:", round(kd / a.elapsed_time(b) / 1e6, 2), "Gevals/s")

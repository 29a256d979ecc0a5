"""Resampling with tensor-product evaluation: 512^3 Float32 volume -> m^3 output grid (itp(xs, ys, zs))."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load(); L = m._lib.lib()
n = 512; mo = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
A = torch.randn((n, n, n), device="cuda")
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
vs = [torch.linspace(1, n, mo, device="cuda", dtype=torch.float32) for _ in range(3)]
out = m.evaluate_grid(itp, *vs); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = m.evaluate_grid(itp, *vs); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
nq = mo ** 3
print(f"grid {mo}^3 outputs: {ms:.2f} ms, {nq / ms / 1e6:.1f} Gevals/s, out write {4 * nq / ms / 1e6:.0f} GB/s")
k = 100
ref = itp(vs[0][:k, None, None].expand(k, k, k), vs[1][None, :k, None].expand(k, k, k), vs[2][None, None, :k].expand(k, k, k))
assert torch.equal(out[:k, :k, :k], ref)
L.b200itp_profile_enable(1); out = m.evaluate_grid(itp, *vs); torch.cuda.synchronize(); L.b200itp_profile_enable(0)
print({k_: round(min(v), 3) for k_, v in m._lib.profile_stages().items()})
prev = L.b200itp_eval_grid_set_separable(0)
del out
a.record(); out2 = m.evaluate_grid(itp, *vs); b.record(); torch.cuda.synchronize()
L.b200itp_eval_grid_set_separable(prev)
print(f"expansion path: {a.elapsed_time(b):.2f} ms, {nq / a.elapsed_time(b) / 1e6:.1f} Gevals/s")

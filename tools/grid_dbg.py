"""Debug driver of the separable grid evaluation (small case against the pointwise kernel)."""
import sys
sys.path.insert(0, ".")
import b200_loader, torch
m = b200_loader.load()
n = 512
A = torch.randn((n, n, n), device="cuda")
itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))))
def t(fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
for mo in (256, 512, 640):
    vs = [torch.linspace(1, n, mo, device="cuda", dtype=torch.float32) for _ in range(3)]
    tg = t(lambda: m.evaluate_grid(itp, *vs))
    # Julia order: first axis fastest -> flat coordinate arrays in that order
    X = vs[0][None, None, :].expand(mo, mo, mo).reshape(-1); Y = vs[1][None, :, None].expand(mo, mo, mo).reshape(-1); Z = vs[2][:, None, None].expand(mo, mo, mo).reshape(-1)
    tb = t(lambda: itp(X, Y, Z))
    # random permutation of the same points
    print(f"m={mo}: grid {mo**3/tg/1e6:.1f} Gevals/s, broadcast(same order) {mo**3/tb/1e6:.1f} Gevals/s")

"""Debug: warp-per-line prefilter kernel vs the oracle for every system, several line lengths."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import b200_loader
m = b200_loader.load()
sys.path.insert(0, "oracle")
import oracle as orc
L = m._lib.lib()
def specs():
    out = []
    for GT in (m.OnGrid, m.OnCell):
        out += [m.BSpline(m.Cubic(BC(GT()))) for BC in (m.Flat, m.Line, m.Free, m.Periodic)]
        out += [m.BSpline(m.Quadratic(BC(GT()))) for BC in (m.Flat, m.Line, m.Free, m.Periodic, m.Reflect)]
    return out
for n0 in (256, 384, 1024):
    rng = np.random.default_rng(n0)
    for it in specs():
        pad = 2 if m.core.needs_padding(it) else 0
        A = rng.standard_normal((n0 - pad, 2400)).astype(np.float32)
        ref = orc.prefilter(A, it)
        L.b200itp_profile_enable(1)
        got = m.interpolate(A, it).coefs_array()
        torch.cuda.synchronize()
        L.b200itp_profile_enable(0)
        used = "prefilter_axis0_warpline" in m._lib.profile_stages()
        err = np.abs(got.astype(np.float64) - ref) / np.abs(ref).max()
        i = np.unravel_index(np.argmax(err), err.shape)
        rows = np.where(err.max(axis=1) > 1e-5)[0]
        print(n0, it, "warpline" if used else "other", "max", float(err.max()), "at", i, "bad rows", rows[:12], len(rows))

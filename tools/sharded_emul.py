"""Emulates the N-rank slab-sharded prefilter (parallel.prefilter_sharded) on ONE GPU, slab by slab, and compares it
with the single-GPU prefilter: bit-identical?  max relative error?"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import b200_loader
m = b200_loader.load()
n = 512
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
skip = m.BSpline(m.Linear())
torch.manual_seed(0)
A = torch.randn((n, n, n), device="cuda")          # (z, y, x), x contiguous
full = A.reshape(-1).clone()
m.prefilter_(full, (n, n, n), (it,) * 3, 0)
full = full.view(n, n, n)
zb = [m.parallel.shard_bounds(n, W, r) for r in range(W)]
L = m._lib.lib()
if len(sys.argv) > 2 and sys.argv[2] == "noseg":
    L.b200itp_prefilter_set_segmentation(0)  # what parallel.prefilter_sharded does around its solves
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
slabs = []
for lo, hi in zb:
    s = A[lo:hi].contiguous().clone()
    m.prefilter_(s.view(-1), (n, n, hi - lo), (it, skip, skip), 0)
    m.prefilter_(s.view(-1), (n, n, hi - lo), (skip, it, skip), 0)
    slabs.append(s)
xy = torch.cat(slabs, 0)
out = torch.empty_like(xy)
for lo, hi in zb:  # y-shards
    ys = xy[:, lo:hi, :].contiguous()
    m.prefilter_(ys.view(-1), (n, hi - lo, n), (skip, skip, it), 0)
    out[:, lo:hi, :] = ys
t1.record()
torch.cuda.synchronize()
print(f"all slabs, 3 passes + torch packing: {t0.elapsed_time(t1):.3f} ms")
err = (out - full).abs().max().item() / full.abs().max().item()
print(f"W={W}: bit-identical {torch.equal(out, full)}, max |diff| / max |c| = {err:.3e}, "
      f"differing elements {(out != full).sum().item()} of {out.numel()}")

"""Probe: does running the x and y passes of the C3 prefilter slab by slab (z-slabs small enough for the L2), and the z
pass afterwards, save a DRAM round trip?  Everything through the C ABI (b200itp_prefilter_from / b200itp_prefilter)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200_loader
m = b200_loader.load()
core = m.core
dev = 0
n = 512
it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
lin = m.BSpline(m.Linear())
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand(n * n * n, device="cuda", dtype=torch.float32, generator=g)   # column-major (x, y, z) flat
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def whole(out):
    core.prefilter_(out, (n, n, n), (it, it, it), dev, src=A)

def slabbed(out, nzs, inplace_first=False):
    plane = n * n
    for z0 in range(0, n, nzs):
        a = A[z0 * plane:(z0 + nzs) * plane]
        o = out[z0 * plane:(z0 + nzs) * plane]
        core.prefilter_(o, (n, n, nzs), (it, it, lin), dev, src=a)
    core.prefilter_(out, (n, n, n), (lin, lin, it), dev)

def timeit(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

def graphed(fn):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        fn()
    return gr.replay

ref = torch.empty_like(A); whole(ref); torch.cuda.synchronize()
print("whole call      median %.3f ms  min %.3f" % timeit(lambda: whole(ref)))
try:
    rp = graphed(lambda: whole(ref))
    print("whole call as a CUDA graph: median %.3f ms  min %.3f" % timeit(rp))
except Exception as e:
    print("graph capture failed:", str(e)[:200])
for nzs in (8, 16, 32, 64, 128):
    out = torch.empty_like(A)
    slabbed(out, nzs); torch.cuda.synchronize()
    same = torch.equal(out, ref)
    print("slabs of %3d planes (%3d MiB): median %.3f ms  min %.3f  identical=%s" % ((nzs, nzs * n * n * 4 >> 20) + timeit(lambda: slabbed(out, nzs)) + (same,)))
    try:
        rp = graphed(lambda: slabbed(out, nzs))
        print("    as a CUDA graph: median %.3f ms  min %.3f" % timeit(rp))
    except Exception as e:
        print("    graph capture failed:", str(e)[:200])

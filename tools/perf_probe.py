"""Quick device-side timing of the hot-path kernels (CUDA events, warm-up, inputs larger than L2).
Used during development; bench.py is the contract-conforming benchmark."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200_loader

m = b200_loader.load()
L = m._lib.lib()
PEAK = 6554.9
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def prefilter_case(name, shape, it, dtype):
    A = torch.randn(shape, device="cuda", dtype=dtype)
    itp = m.interpolate(A, it)
    size = itp.size
    N = len(size)
    base = itp.coefs.clone()
    work = base.clone()
    its = itp.its
    es = work.element_size()
    out = {}
    # per axis
    sz = (C.c_int64 * N)(*size)
    for ax in range(N):
        deg = [its[d].degree.code if d == ax else 1 for d in range(N)]
        degc = (C.c_int32 * N)(*deg)
        bc = (C.c_int32 * N)(*[i.degree.bc.code for i in its])
        gt = (C.c_int32 * N)(*[i.degree.bc.gt.code for i in its])
        code = 0 if dtype == torch.float32 else 1

        def run():
            L.b200itp_prefilter(0, C.c_void_p(work.data_ptr()), code, N, sz, degc, bc, gt, None)
        med, mn = timeit(run)
        nbytes = 2 * es * work.numel()
        out[f"axis{ax}"] = dict(ms=med, ms_min=mn, gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / PEAK)
        work.copy_(base)

    def run_all():
        m.prefilter_(work, size, its, 0)
    med, mn = timeit(run_all)
    nfilt = sum(1 for i in its if i.degree.code > 1)
    nbytes = 2 * es * work.numel() * nfilt
    out["all"] = dict(ms=med, ms_min=mn, gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / PEAK)
    print(name, json.dumps(out), flush=True)
    return out


def eval_case(name, shape, it, dtype, nq, wrap=None, grad=False, sorted_q=False):
    A = torch.randn(shape, device="cuda", dtype=dtype)
    itp = m.interpolate(A, it)
    obj = wrap(itp) if wrap else itp
    bnds = obj.bounds()
    g = torch.Generator(device="cuda").manual_seed(2)
    coords = [(torch.rand(nq, device="cuda", dtype=dtype, generator=g) * (float(u) - float(l)) + float(l)).clamp(float(l), float(u))
              for (l, u) in bnds]
    res = {}

    def run_v():
        return obj(*coords)
    med, mn = timeit(run_v, warm=2, reps=5)
    res["value"] = dict(ms=med, gevals=nq / med / 1e6)
    if grad:
        def run_g():
            return m.gradient(obj, *coords)
        med, mn = timeit(run_g, warm=2, reps=5)
        res["gradient"] = dict(ms=med, gevals=nq / med / 1e6)
    if sorted_q:
        def run_s():
            return m.evaluate_sorted(obj, *coords)
        med, mn = timeit(run_s, warm=2, reps=5)
        res["sorted"] = dict(ms=med, gevals=nq / med / 1e6)
    print(name, json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    f32, f64 = torch.float32, torch.float64
    R = {}
    if which in ("all", "prefilter"):
        R["c3_prefilter"] = prefilter_case("C3 512^3 f32 Cubic(Periodic(OnGrid))", (512, 512, 512),
                                           m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), f32)
        R["c2_prefilter"] = prefilter_case("C2 4096^2 f32 Cubic(Flat(OnCell))", (4096, 4096),
                                           m.BSpline(m.Cubic(m.Flat(m.OnCell()))), f32)
        R["c4_prefilter"] = prefilter_case("C4 8192^2 f32 Quadratic(Reflect(OnCell))", (8192, 8192),
                                           m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))), f32)
        R["c1_prefilter"] = prefilter_case("C1 1e6 f64 Cubic(Line(OnGrid))", (1000000,),
                                           m.BSpline(m.Cubic(m.Line(m.OnGrid()))), f64)
        R["c3_f64"] = prefilter_case("512^3 f64 Cubic(Line(OnGrid))", (384, 384, 384),
                                     m.BSpline(m.Cubic(m.Line(m.OnGrid()))), f64)
    if which in ("all", "eval"):
        R["c3_eval"] = eval_case("C3 eval 2^27 q", (512, 512, 512), m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), f32,
                                 1 << 27, grad=True, sorted_q=True)
        R["c3_eval_1e9"] = eval_case("C3 eval 1e9 q", (512, 512, 512), m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), f32,
                                     10**9, sorted_q=True)
        R["c3_eval_small"] = eval_case("64^3 eval 2^27 q (L2 resident)", (64, 64, 64),
                                       m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), f32, 1 << 27)
        R["c2_eval"] = eval_case("C2 eval 1e8 q", (4096, 4096), m.BSpline(m.Cubic(m.Flat(m.OnCell()))), f32, 10**8,
                                 grad=True, sorted_q=True)
        rng = [m.jrange(np.float32(0), np.float32(1), length=8192, dtype=np.float32)] * 2
        R["c4_eval"] = eval_case("C4 eval 1e8 q", (8192, 8192), m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))), f32,
                                 10**8, wrap=lambda i: m.extrapolate(m.scale(i, *rng), m.Line()))
        R["c5_eval"] = eval_case("C5 eval 2^27 q", (128, 128, 128, 128), m.BSpline(m.Linear()), f64, 1 << 27)
        R["c1_eval"] = eval_case("C1 eval 1e7 q", (1000000,), m.BSpline(m.Cubic(m.Line(m.OnGrid()))), f64, 10**7)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(R, open(os.path.join(ROOT, "gpurun_out", f"probe_{which}.json"), "w"), indent=1)

#!/usr/bin/env python
"""bench.py — headline benchmark of the B-spline hot path on B200 (BASELINE.json metric).

Headline workload (BASELINE.json configs[2], the configuration the metric is quoted on):
  C3: 3-D tricubic `Cubic(Periodic(OnGrid()))`, 512^3 Float32 volume: prefilter + 10^9 scattered queries.
One step = `interpolate(A, it)` (copy + 3 prefilter passes) on rank 0, NCCL broadcast of the coefficient array (N > 1),
then `itp.(xs, ys, zs)` over the rank's shard of the 10^9 queries (STRONG scaling: the query total is fixed, each rank
evaluates 10^9 / N).  `value` = queries evaluated by all ranks / step time (Gevals/s), inputs resident in HBM.  The weak
variant (10^9 queries on every rank) is reported beside it under "weak_scaling".
`e2e` = the same step through the public API with HOST buffers (pinned): H2D of the step's samples and coordinates and
D2H of its results inside the timed region.
`configs` holds the other four BASELINE shapes (C1, C2 value+gradient, C4 scale+extrapolate fused, C5 4-D linear), each
with prefilter GB/s, evaluation Gevals/s and its HBM-roofline fraction, query-sharded over the N ranks as well.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--queries Q] [--impl reference] [--configs 0|1]
N > 1 is launched by torchrun (one rank per GPU, NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 512
DEFAULT_QUERIES = 10**9
# DRAM bytes per query of one ordered evaluation call and of the three C3 prefilter passes: ncu dram__bytes_read+write
# per launch, summed (profiles/).  Refreshed from the launch list named in profiles/README.md when the pipeline changes.
NCU = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
METRIC = "Gevals/s tricubic + prefilter GB/s (512^3 f32) at 1/2/4/8 B200 vs HBM peak"
C3_WORKLOAD = "C3: 3-D Cubic(Periodic(OnGrid())) 512^3 Float32, prefilter + 1e9 scattered queries"


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def host_threads():
    """Threads the CPU legs use: the cores this process may run on (NOT the inherited OMP_NUM_THREADS, which torchrun
    sets to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank's CPU threads (and therefore its pinned host buffers, first-touch) to the NUMA node the GPU's PCIe
    slot hangs off: host<->device copies of 8 ranks otherwise cross the socket interconnect.  Best effort; returns a
    short description for the JSON line."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                                      text=True, timeout=10).strip().lower()
        # sysfs uses a 4-digit PCI domain, nvidia-smi prints 8
        parts = bus.split(":")
        sysbus = ":".join([parts[0][-4:]] + parts[1:])
        node = int(open(f"/sys/bus/pci/devices/{sysbus}/numa_node").read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa node {node} ({len(allowed)} cpus)"
        return f"numa node {node}: no allowed cpus, unchanged"
    except Exception as e:  # no sysfs / no permission: leave the affinity alone
        return f"numa: unchanged ({type(e).__name__})"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                               "-i", str(self.gpu)], text=True, timeout=5)
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def synthetic_volume(torch, dev, seed=20240601):
    """Smooth sin-product field plus uniform noise, generated on the device.  Returned as a COLUMN-MAJOR
    tensor (shape (n1, n2, n3) in Julia index order, first index contiguous): what a Julia array is in memory, so
    `interpolate` does not have to transpose it."""
    g = torch.Generator(device=dev).manual_seed(seed)
    ax = torch.arange(1, N_GRID + 1, device=dev, dtype=torch.float32)
    s = torch.sin(ax * (2 * np.pi * 3 / N_GRID))
    A = s[:, None, None] * torch.cos(ax * (2 * np.pi * 5 / N_GRID))[None, :, None] * s[None, None, :]
    A = A + 0.1 * (torch.rand((N_GRID,) * 3, device=dev, dtype=torch.float32, generator=g) - 0.5)
    return A.contiguous().permute(2, 1, 0)  # a view: element [i1, i2, i3] at offset i1 + n*(i2 + n*i3)


def uniform(torch, dev, n, lo, hi, dtype, gen):
    x = torch.rand(n, device=dev, dtype=dtype, generator=gen) * (hi - lo) + lo
    return x.clamp_(lo, hi)


def synthetic_queries(torch, dev, nq, rank, seed=20240602):
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    return [uniform(torch, dev, nq, 1.0, float(N_GRID), torch.float32, g) for _ in range(3)]


# ------------------------------------------------------------------------------------------ CPU legs (oracle/)
def cpu_c3_sample(steps, warmup, threads, nq_s):
    """The reference's CPU path for C3 on this box's host cores: prefilter of the FULL 512^3 Float32 volume (the
    oracle's LU + explicit Woodbury sweeps, OpenMP over lines) and a bounded sample of the 1e9 queries through the
    type-specialised tricubic evaluator (what Julia compiles for this interpolant), OpenMP over queries.
    Returns median seconds (prefilter, evaluation of nq_s queries, 1-thread Mevals/s)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc  # test infrastructure: here as the timed CPU baseline only
    m = orc.pkg
    rng = np.random.default_rng(20240601)
    A = rng.standard_normal((N_GRID,) * 3).astype(np.float32)
    it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
    q = [rng.uniform(1, N_GRID, nq_s).astype(np.float32) for _ in range(3)]
    tp, te = [], []
    coefs = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        c = orc.prefilter(A, it, nthreads=threads, fast=True)
        t1 = time.perf_counter()
        if coefs is None:
            coefs = np.ascontiguousarray(c.reshape(-1, order="F"))
        orc.evaluate_tricubic_periodic_f32(coefs, (N_GRID,) * 3, *q, nthreads=threads, fast=True)
        t2 = time.perf_counter()
        if i >= warmup:
            tp.append(t1 - t0)
            te.append(t2 - t1)
    k1 = min(nq_s, 400_000)
    t3 = time.perf_counter()
    orc.evaluate_tricubic_periodic_f32(coefs, (N_GRID,) * 3, *[c[:k1] for c in q], nthreads=1, fast=True)
    t4 = time.perf_counter()
    return float(np.median(tp)), float(np.median(te)), k1 / (t4 - t3) / 1e6


def cpu_baseline_dict(steps, warmup, nq_total, nq_s=20_000_000):
    threads = host_threads()
    tp, te, one = cpu_c3_sample(steps, warmup, threads, nq_s)
    step_s = tp + te * (nq_total / nq_s)
    pre_gbs = 2 * 4 * N_GRID**3 * 3 / tp / 1e9
    return {"value": nq_total / step_s / 1e9, "unit": "Gevals/s", "cores": threads, "kind": "port",
            "sample": f"prefilter of the full 512^3 Float32 volume ({tp*1e3:.0f} ms = {pre_gbs:.2f} GB/s) + {nq_s} of the "
                      f"{nq_total} tricubic queries ({nq_s/te/1e6:.1f} Mevals/s on {threads} OpenMP threads, "
                      f"{one:.2f} Mevals/s on 1), evaluation time scaled to {nq_total} queries; Julia absent: C "
                      "restatement of the reference (oracle/), type-specialised Float32 tricubic loop",
            "prefilter_gbs": pre_gbs, "eval_gevals": nq_s / te / 1e9, "eval_mevals_1thread": one,
            "step_s": step_s}


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone times the reference's CPU path (C port, all host threads) and prints the line."""
    if rank != 0:
        return
    cb = cpu_baseline_dict(args.steps, args.warmup, args.queries)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Gevals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["step_s"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C3_WORKLOAD, "queries_per_step": args.queries,
                   "sampled": "prefilter at full size; evaluation on a sample of the queries, time scaled linearly"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "Gevals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU helpers
class Ctx:
    pass


def flush_l2(ctx):
    """Between timed iterations of workloads that fit the L2: overwrite a buffer larger than the L2 (256 MiB)."""
    ctx.flush.add_(1.0)


def timed(ctx, fn, steps, warmup, flush=False):
    """Per-iteration CUDA-event times (ms) of fn() on the current stream, max over ranks of the mean."""
    torch = ctx.torch
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        if flush:
            flush_l2(ctx)
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([float(np.mean(ts))], device=ctx.dev, dtype=torch.float64)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return t.item()


def shard(nq, world, rank):
    base, rem = divmod(nq, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_config(ctx, name, workload, make_samples, it_of, wrap, coord_bounds, cdtype, nq_total, want_grad, alg):
    """One BASELINE shape: prefilter (replicated: every rank, SURVEY §8e 'replicas only') and query-sharded evaluation.
    alg = dict(pre_bytes, eval_bytes(nq), grad_bytes(nq))."""
    torch, m, L = ctx.torch, ctx.m, ctx.L
    dev = ctx.dev
    A = make_samples()
    it = it_of(m)
    peak = ctx.peak
    out = {"workload": workload}
    holder = {}

    def build():
        holder["itp"] = m.interpolate(A, it, device=dev)

    pre_ms = timed(ctx, build, ctx.csteps, 3, flush=True)
    itp = wrap(m, holder["itp"])
    if alg["pre_bytes"]:
        out["prefilter"] = {"ms": pre_ms, "algorithmic_bytes": alg["pre_bytes"],
                            "gbs": alg["pre_bytes"] / (pre_ms * 1e-3) / 1e9,
                            "frac": alg["pre_bytes"] / (pre_ms * 1e-3) / 1e9 / peak,
                            "note": "interpolate(A, it): copy_with_padding + one pass per filtered axis, whole call timed"}
    lo, hi = shard(nq_total, ctx.world, ctx.rank)
    nq = hi - lo
    g = torch.Generator(device=dev).manual_seed(977 + ctx.rank)
    coords = [uniform(torch, dev, nq, b[0], b[1], cdtype, g) for b in coord_bounds]
    before = L.b200itp_eval_sorted_ordered_calls()
    res = {}

    def ev():
        res["v"] = itp(*coords)

    ev_ms = timed(ctx, ev, ctx.csteps, 3, flush=True)
    path = "ordered" if L.b200itp_eval_sorted_ordered_calls() > before else "direct"
    eb = alg["eval_bytes"](nq_total)
    out["eval"] = {"ms": ev_ms, "queries": nq_total, "gevals": nq_total / (ev_ms * 1e-3) / 1e9, "path": path,
                   "algorithmic_bytes": eb, "gbs": eb / (ev_ms * 1e-3) / 1e9, "frac": eb / (ev_ms * 1e-3) / 1e9 / peak,
                   "checksum": float(res["v"].double().sum().item())}
    if want_grad:
        before = L.b200itp_eval_sorted_ordered_calls()

        def gr():
            res["g"] = m.gradient(itp, *coords)

        g_ms = timed(ctx, gr, ctx.csteps, 3, flush=True)
        gb = alg["grad_bytes"](nq_total)
        out["gradient"] = {"ms": g_ms, "queries": nq_total, "gevals": nq_total / (g_ms * 1e-3) / 1e9,
                           "path": "ordered" if L.b200itp_eval_sorted_ordered_calls() > before else "direct",
                           "algorithmic_bytes": gb, "gbs": gb / (g_ms * 1e-3) / 1e9,
                           "frac": gb / (g_ms * 1e-3) / 1e9 / peak}
    del coords, res, holder, itp, A
    m.core.release_scratch()
    torch.cuda.empty_cache()
    return out


def grid_config(ctx):
    """Resampling: `itp(xvec, yvec, zvec)` of the C3 volume onto a 1000^3 grid: the separable kernels, one
    pass per axis.  N > 1: the output z-range is sharded.  Algorithmic bytes: coefficients once + the output once."""
    torch, m, L, dev = ctx.torch, ctx.m, ctx.L, ctx.dev
    no = 1000
    g = torch.Generator(device=dev).manual_seed(5)
    A = torch.rand((N_GRID,) * 3, device=dev, dtype=torch.float32, generator=g).permute(2, 1, 0)
    itp = m.interpolate(A, m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))), device=dev)
    lo, hi = shard(no, ctx.world, ctx.rank)
    vecs = [torch.linspace(1.0, float(N_GRID), no, device=dev, dtype=torch.float32) for _ in range(3)]
    vecs[2] = vecs[2][lo:hi].contiguous()
    c0 = L.b200itp_eval_grid_separable_calls()
    res = {}

    def ev():
        res["v"] = None
        res["v"] = m.evaluate_grid(itp, *vecs)

    ms = timed(ctx, ev, ctx.csteps, 3, flush=False)
    nout = no ** 3
    b = 4 * nout + 4 * N_GRID ** 3
    out = {"workload": "grid: itp(xvec, yvec, zvec), 512^3 Float32 Cubic(Periodic(OnGrid())) resampled onto 1000^3",
           "eval": {"ms": ms, "queries": nout, "gevals": nout / (ms * 1e-3) / 1e9,
                    "path": "separable" if L.b200itp_eval_grid_separable_calls() > c0 else "expansion",
                    "algorithmic_bytes": b, "gbs": b / (ms * 1e-3) / 1e9, "frac": b / (ms * 1e-3) / 1e9 / ctx.peak,
                    "l2": "output 4 GB and intermediates 3 GB per call: larger than the L2",
                    "checksum": float(res["v"].double().sum().item())}}
    del res, itp, A, vecs
    torch.cuda.empty_cache()
    return out


def other_configs(ctx):
    """C1, C2, C4, C5 of BASELINE.json (SURVEY §8 sizes and synthetic inputs; §8d algorithmic bytes)."""
    torch, dev = ctx.torch, ctx.dev
    g = torch.Generator(device=dev).manual_seed(20240601)
    out = {}

    def rnd(shape, dt):  # column-major view (Julia memory order), no transposition inside interpolate
        t = torch.rand(tuple(reversed(shape)), device=dev, dtype=dt, generator=g)
        return t.permute(*reversed(range(len(shape)))) if len(shape) > 1 else t

    # C1: 1-D Cubic(Line(OnGrid())), 1e6 Float64 samples, 1e7 queries
    n1 = 10**6
    out["C1"] = run_config(
        ctx, "C1", "C1: 1-D Cubic(Line(OnGrid())) 1e6 Float64 samples, 1e7 random queries",
        lambda: rnd((n1,), torch.float64), lambda m: m.BSpline(m.Cubic(m.Line(m.OnGrid()))), lambda m, itp: itp,
        [(1.0, float(n1))], torch.float64, 10**7, False,
        {"pre_bytes": 2 * 8 * (n1 + 2), "eval_bytes": lambda q: q * 16 + 8 * (n1 + 2)})
    # C2: 2-D bicubic Cubic(Flat(OnCell())) 4096^2 Float32, 1e8 queries + gradient
    n2 = 4096
    out["C2"] = run_config(
        ctx, "C2", "C2: 2-D Cubic(Flat(OnCell())) 4096^2 Float32, 1e8 scattered queries + gradient",
        lambda: rnd((n2, n2), torch.float32), lambda m: m.BSpline(m.Cubic(m.Flat(m.OnCell()))), lambda m, itp: itp,
        [(0.5, n2 + 0.5)] * 2, torch.float32, 10**8, True,
        {"pre_bytes": 2 * 4 * (n2 + 2) ** 2 * 2, "eval_bytes": lambda q: q * (8 + 4) + 4 * (n2 + 2) ** 2,
         "grad_bytes": lambda q: q * (8 + 8) + 4 * (n2 + 2) ** 2})
    # C4: 2-D Quadratic(Reflect(OnCell())) 8192^2 Float32, scale(ranges) + extrapolate(Line()) fused; ~30 % of the
    # queries fall outside [0,1]^2 and take the Line branch
    n4 = 8192

    def wrap4(m, itp):
        r = m.jrange(np.float32(0), np.float32(1), length=n4, dtype=np.float32)
        return m.extrapolate(m.scale(itp, r, r), m.Line())

    out["C4"] = run_config(
        ctx, "C4", "C4: 2-D Quadratic(Reflect(OnCell())) 8192^2 Float32, scale(range(0f0,1f0,8192)) + extrapolate(Line()), "
                   "1e8 queries uniform over [-0.1,1.1]^2",
        lambda: rnd((n4, n4), torch.float32), lambda m: m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))), wrap4,
        [(-0.1, 1.1)] * 2, torch.float32, 10**8, True,
        {"pre_bytes": 2 * 4 * (n4 + 2) ** 2 * 2, "eval_bytes": lambda q: q * (8 + 4) + 4 * (n4 + 2) ** 2,
         "grad_bytes": lambda q: q * (8 + 8) + 4 * (n4 + 2) ** 2})
    # C5: 4-D Linear() 128^4 Float64, 1e9 queries (no prefilter)
    n5 = 128
    out["C5"] = run_config(
        ctx, "C5", "C5: 4-D Linear() 128^4 Float64 grid, 1e9 queries (2^4 stencil gather, no prefilter)",
        lambda: rnd((n5,) * 4, torch.float64), lambda m: m.BSpline(m.Linear()), lambda m, itp: itp,
        [(1.0, float(n5))] * 4, torch.float64, ctx.c5_queries, False,
        {"pre_bytes": 0, "eval_bytes": lambda q: q * (32 + 8) + 8 * n5**4})
    out["grid"] = grid_config(ctx)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--queries", type=int, default=DEFAULT_QUERIES, help="queries per step in TOTAL (sharded over the GPUs)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", type=int, default=1, help="also run C1, C2, C4, C5")
    ap.add_argument("--c5-queries", type=int, default=10**9)
    ap.add_argument("--lead-deficit", type=int, default=14_000_000,
                    help="N > 1: queries rank 0 (prefilter + broadcast source) gets fewer than each other rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import b200_loader
    m = b200_loader.load()
    L = m._lib.lib()  # raises if the CUDA library is missing: there is no fallback
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "numa: not bound (single rank)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            # the coefficient broadcast runs beside the evaluation's persistent count / split kernels (one CTA per SM,
            # nearly all registers): on a normal-priority stream NCCL's CTAs only get an SM when such a kernel ends
            # and then queue behind the next one; high priority puts them first at those boundaries
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.m, ctx.L, ctx.dev, ctx.rank, ctx.world, ctx.barrier = torch, dist, m, L, dev, rank, world, barrier
    ctx.peak, peak_kind = measured_peak()
    ctx.flush = torch.zeros(64 << 20, device=dev, dtype=torch.float32)
    ctx.csteps = max(2, min(args.steps, 5))
    ctx.c5_queries = args.c5_queries
    peak = ctx.peak

    it = m.BSpline(m.Cubic(m.Periodic(m.OnGrid())))
    A = synthetic_volume(torch, dev)
    ncoef = N_GRID**3

    side = torch.cuda.Stream(device=dev, priority=-1)  # prefilter + broadcast: ahead of the queued split kernels

    def make_step(coords):
        def step():
            # rank 0 prefilters; the coefficient array is replicated over NVLink; every rank evaluates its shard.
            # The coefficients are produced on a side stream: the ordered pipeline's count / split kernels only read the
            # query coordinates, so they run meanwhile, and the gather kernel waits for `ready_event`
            # (b200itp_eval_set_coefs_event).  Same work per step, the serial prologue is off the critical path.
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if rank == 0:
                    coefs = m.interpolate(A, it, device=dev).coefs
                else:
                    coefs = torch.empty(ncoef, dtype=torch.float32, device=dev)
                if world > 1:
                    dist.broadcast(coefs, src=0)
                ready = torch.cuda.Event()
                ready.record(side)
            coefs.record_stream(main)
            itp = m.BSplineInterpolation(coefs, (N_GRID,) * 3, (N_GRID,) * 3, (it,) * 3, local_rank)
            itp.ready_event = ready
            return itp(*coords)  # the drop-in call: takes the ordered pipeline for this size
        return step

    def run_steps(step, steps, warmup):
        for _ in range(warmup):
            out = step()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        L.b200itp_profile_enable(1)
        barrier()
        ev[0].record()
        for _ in range(steps):
            out = step()
        ev[1].record()
        barrier()
        L.b200itp_profile_enable(0)
        t_ms = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return t_ms.item() / steps, out

    # ---- headline: 1e9 queries in total, sharded (strong scaling); every kernel launch of the timed steps is bracketed
    #      by CUDA events on the launching stream (b200itp_profile_*): the roofline's per-kernel durations
    # rank 0 also prefilters the volume and feeds the broadcast while its own count / split kernels run (+0.6 ms of
    # device time per step, whatever N): it takes --lead-deficit fewer queries than each other rank
    from interpolations_jl_b200 import parallel as _par
    lead = min(args.lead_deficit, args.queries // (4 * world)) if world > 1 else 0
    lo, hi = _par.shard_bounds_lead(args.queries, world, rank, lead)
    nq = hi - lo
    nq_others = (lambda b: b[1] - b[0])(_par.shard_bounds_lead(args.queries, world, world - 1, lead))
    coords = synthetic_queries(torch, dev, nq, rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = L.b200itp_launch_count()
    oc0 = L.b200itp_eval_sorted_ordered_calls()
    ms_per_step, out = run_steps(make_step(coords), args.steps, args.warmup)
    stages = m._lib.profile_stages()
    launches = L.b200itp_launch_count() - l0
    if L.b200itp_eval_sorted_ordered_calls() == oc0:
        raise RuntimeError("bench: the ordered evaluation pipeline did not run (fell back to the direct kernel)")
    launches_per_step = launches / (args.steps + args.warmup)
    value = args.queries / (ms_per_step * 1e-3) / 1e9
    checksum = float(out.double().sum().item())
    del out

    # ---- roofline (HBM): algorithmic bytes per launch / live CUDA-event duration
    per_query = {"ord_count": 12, "ord_split1": 12 + 16, "ord_split2": 16 + 16, "ord_eval": 16 + 8,
                 "ord_unsortA": 8 + 8, "ord_unsortB": 8 + 8, "ord_place": 8 + 4, "eval_direct_value": 12 + 4}
    stage_rows = []
    eval_ms = 0.0
    pre_ms = 0.0
    for name, ts in stages.items():
        avg = float(np.mean(ts))
        per_step = float(np.sum(ts)) / args.steps
        if name in per_query:
            b = per_query[name] * nq + (4 * ncoef if name in ("ord_eval", "eval_direct_value") else 0)
            eval_ms += per_step
        elif name.startswith("prefilter_axis"):
            b = 2 * 4 * ncoef
            pre_ms += per_step
        elif name == "copy_with_padding":
            b = 2 * 4 * ncoef
        else:
            b = 0
            if name.startswith("ord_"):
                eval_ms += per_step
        row = {"kernel": name, "launches": len(ts), "ms": avg, "stream_bytes": b,
               "gbs": b / (avg * 1e-3) / 1e9 if avg > 0 else None,
               "frac": b / (avg * 1e-3) / 1e9 / peak if avg > 0 else None}
        if name.startswith("prefilter_axis") or name == "copy_with_padding":
            # side stream: these launches share the device with ord_count / ord_split1 of the same step, so their
            # event-to-event time is not the kernel's own; the `prefilter` roofline below is the isolated measurement
            row["overlapped_with"] = "ord_count / ord_split1 (side stream)"
            row["gbs"] = row["frac"] = None
        stage_rows.append(row)
    eval_bytes = nq * (12 + 4) + 4 * ncoef          # coordinates in, results out, coefficients once
    pre_bytes = 2 * 4 * ncoef * 3                   # one read + one write per axis pass
    # Inside the step the prefilter runs on the side stream, concurrently with the count / split kernels of the
    # evaluation (they share the HBM and the SMs), so its in-step stage times are not its own: the prefilter roofline is
    # measured in isolation here (interpolate(A, it) on an idle device, CUDA-event stage times of the three passes).
    pre_ms_in_step = pre_ms
    L.b200itp_profile_enable(1)
    for _ in range(5):
        flush_l2(ctx)
        tmp_itp = m.interpolate(A, it, device=dev)
    torch.cuda.synchronize()
    L.b200itp_profile_enable(0)
    st2 = m._lib.profile_stages()
    pre_ms = float(sum(np.mean(v) for k, v in st2.items() if k.startswith("prefilter_axis")))
    del tmp_itp
    roof = {"bound": "hbm", "kernel": "ordered evaluation: all launches of one b200itp_eval_sorted call on this rank's "
                                      f"{nq} queries",
            "achieved": eval_bytes / (eval_ms * 1e-3) / 1e9, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": eval_bytes / (eval_ms * 1e-3) / 1e9 / peak,
            "traffic": NCU["ordered_eval_bytes_per_query"] * nq + 4 * ncoef,
            "traffic_source": NCU["source"], "algorithmic_bytes": eval_bytes, "ms": eval_ms, "stages": stage_rows}
    roof_pre = {"bound": "hbm", "kernel": "prefilter (3 axis passes, 512^3 f32)",
                "achieved": pre_bytes / (pre_ms * 1e-3) / 1e9, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": pre_bytes / (pre_ms * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": pre_bytes / (pre_ms * 1e-3) / 1e9 / 8000,
                "traffic": NCU["prefilter_c3_bytes"], "traffic_source": NCU["source"],
                "algorithmic_bytes": pre_bytes, "ms": pre_ms, "ms_in_step_overlapped": pre_ms_in_step,
                "note": "timed alone (idle device, L2 flushed); inside the step the passes overlap the evaluation's "
                        "count / split kernels on a side stream"}

    # ---- N > 1: the weak variant (every rank its own 1e9 queries) and the slab-sharded prefilter
    weak = None
    sharded = None
    if world > 1:
        del coords
        m.core.release_scratch()
        torch.cuda.empty_cache()
        wcoords = synthetic_queries(torch, dev, args.queries, rank)
        w_ms, wout = run_steps(make_step(wcoords), max(2, args.steps // 2), 2)
        weak = {"value": args.queries * world / (w_ms * 1e-3) / 1e9, "unit": "Gevals/s", "ms_per_step": w_ms,
                "queries_per_gpu_per_step": args.queries, "scaling": "weak"}
        del wcoords, wout
        m.core.release_scratch()
        torch.cuda.empty_cache()
        coords = synthetic_queries(torch, dev, nq, rank)
        from interpolations_jl_b200 import parallel as par
        itp_check = m.interpolate(A, it, device=dev) if rank == 0 else None
        zlo, zhi = par.shard_bounds(N_GRID, world, rank)
        slab = A.permute(2, 1, 0)[zlo:zhi].contiguous()   # torch layout (nz_local, ny, nx) of the Julia array (nx, ny, nz)
        ts_solve, ts_all = [], []
        for i in range(4):
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            ysh, meta = par.prefilter_sharded(slab, it, counts=[b - a for a, b in (par.shard_bounds(N_GRID, world, r) for r in range(world))])
            e1.record()
            full = par.allgather_coefficients(ysh, meta)
            e2.record()
            torch.cuda.synchronize()
            if i > 0:
                ts_solve.append(e0.elapsed_time(e1))
                ts_all.append(e0.elapsed_time(e2))
        t2 = torch.tensor([float(np.mean(ts_solve)), float(np.mean(ts_all))], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(full, itp_check.coefs)) if rank == 0 else True
        sharded = {"prefilter_ms": t2[0].item(), "prefilter_plus_allgather_ms": t2[1].item(),
                   "prefilter_gbs": pre_bytes / (t2[0].item() * 1e-3) / 1e9, "matches_single_gpu": ok,
                   "single_gpu_prefilter_ms": pre_ms,
                   "note": "z-slabs, local x/y passes, one all-to-all, local z pass; then all-gather for evaluation"}
        del full, ysh, slab
        # the same sequence through the C ABI's own entry points (b200itp_comm_init / b200itp_prefilter_sharded): ONE
        # process — rank 0 — drives all N devices, the other ranks wait with idle GPUs
        barrier()
        if rank == 0:
            try:
                comm = par.Comm(list(range(world)))
                base = A.permute(2, 1, 0)
                slabs = []
                for r in range(world):
                    a, b = par.shard_bounds(N_GRID, world, r)
                    slabs.append(base[a:b].to(f"cuda:{r}").contiguous())
                best = None
                for _ in range(4):
                    itps, ms2 = comm.interpolate_sharded(slabs, it)
                    best = ms2 if best is None else (min(best[0], ms2[0]), min(best[1], ms2[1]))
                same = all(bool(torch.equal(i.coefs.to(dev), itp_check.coefs)) for i in itps)
                sharded["c_abi"] = {"prefilter_ms": best[0], "prefilter_plus_allgather_ms": best[1],
                                    "prefilter_gbs": pre_bytes / (best[0] * 1e-3) / 1e9, "matches_single_gpu": same,
                                    "note": "b200itp_prefilter_sharded: one process drives the N devices (ncclSend/Recv "
                                            "all-to-all, ncclAllGather), strided pack kernels, no torch in the path"}
                del itps, slabs
                comm.close()
            except Exception as e:  # reported, never fatal for the headline
                sharded["c_abi"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.set_device(local_rank)
        # the other ranks wait on the HOST (store key), not inside a NCCL barrier kernel that would spin on the very
        # devices rank 0 is timing
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("b200itp_c_abi_done", "1")
        else:
            store.wait(["b200itp_c_abi_done"])
        barrier()
        del itp_check

    # ---- end to end through the public API with host buffers (pinned): H2D samples + coords, D2H results, per step
    e2e = None
    if not args.no_e2e:
        host_coords = [torch.empty(nq, dtype=torch.float32, pin_memory=True) for _ in range(3)]
        for h, c in zip(host_coords, coords):
            h.copy_(c)
        host_A = torch.empty((N_GRID,) * 3, dtype=torch.float32, pin_memory=True)  # memory order of the Julia array
        host_A.copy_(A.permute(2, 1, 0))
        host_out = torch.empty(nq, dtype=torch.float32, pin_memory=True)
        torch.cuda.synchronize()

        def step_e2e():
            if rank == 0:
                Ad = host_A.to(dev, non_blocking=True).permute(2, 1, 0)
                coefs = m.interpolate(Ad, it, device=dev).coefs
            else:
                coefs = torch.empty(ncoef, dtype=torch.float32, device=dev)
            if world > 1:
                dist.broadcast(coefs, src=0)
            itp2 = m.BSplineInterpolation(coefs, (N_GRID,) * 3, (N_GRID,) * 3, (it,) * 3, local_rank)
            m.evaluate_from_host(itp2, *host_coords, out=host_out)

        for _ in range(2):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_steps = max(2, min(args.steps, 3))
        e0.record()
        for _ in range(e_steps):
            step_e2e()
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": args.queries / (te.item() / e_steps * 1e-3) / 1e9, "unit": "Gevals/s",
               "h2d_bytes_per_step": 12 * args.queries + 4 * ncoef, "d2h_bytes_per_step": 4 * args.queries,
               "queries_per_step": args.queries, "queries_per_gpu_per_step": nq, "ms_per_step": te.item() / e_steps,
               "host": numa}
        del host_coords, host_A, host_out
    if sampler:
        sampler.stop_flag.set()
        sampler.join(timeout=3)
    del coords
    m.core.release_scratch()
    torch.cuda.empty_cache()

    configs = other_configs(ctx) if args.configs else None
    if configs is not None:
        configs["C3"] = {"workload": C3_WORKLOAD,
                         "prefilter": {"ms": pre_ms, "gbs": roof_pre["achieved"], "frac": roof_pre["frac"],
                                       "algorithmic_bytes": pre_bytes},
                         "eval": {"ms": eval_ms, "queries": args.queries, "gevals": value, "path": "ordered",
                                  "algorithmic_bytes": eval_bytes, "frac": roof["frac"]}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gevals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C3_WORKLOAD, "queries_per_step": args.queries, "queries_per_gpu_per_step": nq,
                       "queries_per_gpu_per_step_other_ranks": nq_others,
                       "sharding": "contiguous chunks in rank order; rank 0 (prefilter + broadcast source) takes "
                                   f"{lead} fewer queries than each other rank" if world > 1 else "one rank",
                       "call": "itp(xs, ys, zs) (the drop-in broadcast; served by the ordered pipeline)",
                       "l2": "inputs larger than L2 (coefficients 512 MiB, coordinates 12 B/query); the small configs "
                             "flush the L2 between timed iterations",
                       "parallelism": f"query-sharded x{world} (fixed total), coefficients replicated by NCCL broadcast"},
            "prefilter_gbs": roof_pre["achieved"], "prefilter_ms": pre_ms,
            "eval_gevals_kernel": nq / (eval_ms * 1e-3) / 1e9 * world,  # rank 0's rate x N
            "roofline": roof, "roofline_prefilter": roof_pre, "weak_scaling": weak, "prefilter_sharded": sharded,
            "configs": configs, "e2e": e2e,
            "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
            "clocks": sampler.summary() if sampler else None, "checksum": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_dict(1, 1, args.queries)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

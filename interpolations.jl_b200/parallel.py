"""Multi-GPU use of the hot path on one 8xB200 box: one process per GPU, `torch.distributed` (NCCL over
NVLink / NVSwitch) for the plumbing, the CUDA library for every flop (BASELINE.json north_star).

* Evaluation shards: queries are independent, so each rank evaluates a contiguous chunk of the query arrays
  (`shard_bounds`, `shard_queries`) against its own replica of the coefficient array; the replica comes from
  one broadcast (`broadcast_interpolant`) or from the all-gather that ends `interpolate_sharded`.  There is no
  collective on the evaluation data path.
* Large 3-D prefilters shard: every rank owns a z-slab of the (padded) array.  The passes along x and y are
  local; one all-to-all re-shards the array along y so that the pass along z is local too
  (`prefilter_sharded`); `allgather_coefficients` then replicates the result for evaluation.
* Small arrays (C1, C2, C4 of BASELINE.json): replicas only — run `interpolate` on every rank.

Array convention: a Julia array of size (n1, n2, n3), column-major, is held as a C-contiguous torch tensor of
shape (n3, n2, n1); a z-slab is a chunk of the leading torch dimension and is contiguous in memory.

The axis solves go through `solve(t, julia_size, its)`, by default the CUDA library (`core.prefilter_`, which
raises if libb200itp.so is missing).  The CPU tests (gloo, world_size 2) inject a host solver of their own to
check the data movement; the product never does.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import core
from .flags import BSpline, Linear, needs_padding


def _torch():
    import torch
    return torch


def _dist():
    import torch.distributed as dist
    return dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, order-preserving split of range(n) into `world` chunks whose sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bounds_lead(n: int, world: int, rank: int, lead_deficit: int = 0) -> Tuple[int, int]:
    """`shard_bounds` with a lighter first chunk: rank 0 — the rank that also prefilters the volume and feeds the
    coefficient broadcast while its own count / split kernels run — gets `lead_deficit` fewer queries than each of
    the other ranks, so that all ranks finish together.  Chunks stay contiguous and in rank order."""
    n, world, d = int(n), int(world), int(lead_deficit)
    if world <= 1 or d <= 0:
        return shard_bounds(n, world, rank)
    n0 = max(0, (n - (world - 1) * d) // world)
    if rank == 0:
        return 0, n0
    lo, hi = shard_bounds(n - n0, world - 1, rank - 1)
    return n0 + lo, n0 + hi


def shard_queries(xs: Sequence, world: int, rank: int):
    """The rank's chunk of every coordinate array (1-D, equal lengths): `itp.(xs...)` over the full arrays is the
    concatenation of the per-rank results in rank order."""
    n = len(xs[0])
    lo, hi = shard_bounds(n, world, rank)
    return [x[lo:hi] for x in xs]


def _world(group):
    dist = _dist()
    return dist.get_world_size(group), dist.get_rank(group)


_side_streams = {}


def _side_stream(device):
    """One high-priority stream per device for coefficient traffic that overlaps an evaluation: the ordered pipeline's
    count / split kernels are persistent (one CTA per SM, nearly all registers), so work queued at normal priority only
    gets an SM when such a kernel ends and then waits behind the next one.  (For the NCCL kernels themselves create the
    process group with `ProcessGroupNCCL.Options(is_high_priority_stream=True)`, as bench.py does.)"""
    torch = _torch()
    key = device.index
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device, priority=-1)
    return _side_streams[key]


def broadcast_interpolant(itp: Optional[core.BSplineInterpolation], src: int, device=None, group=None, overlap=False):
    """Replicate `itp` (held by rank `src` OF THE GROUP) on every rank: metadata by object broadcast, the coefficient
    array by one `broadcast` (NCCL: 512 MiB cross NVLink once).

    `overlap=True` (CUDA tensors): the broadcast is issued on a high-priority side stream and the returned object
    carries `ready_event`; the next evaluation hands it to the library (`b200itp_eval_set_coefs_event`), whose count /
    split kernels — they read only the coordinates — run meanwhile, and whose gather kernel waits for it."""
    torch, dist = _torch(), _dist()
    world, rank = _world(group)
    gsrc = dist.get_global_rank(group, src) if group is not None else src  # the collectives take GLOBAL ranks
    meta = [None]
    if rank == src:
        meta = [(tuple(itp.size), tuple(itp.parent_len), itp.its, str(itp.coefs.dtype), tuple(itp.parent_first))]
    dist.broadcast_object_list(meta, src=gsrc, group=group)
    size, parent_len, its, dtype, parent_first = meta[0]
    tdtype = torch.float32 if dtype.endswith("float32") else torch.float64
    if rank == src:
        coefs = itp.coefs
    else:
        dev = torch.device("cuda", core._device_index(device)) if device is not None else torch.device("cpu")
        coefs = torch.empty(int(np.prod(size)), dtype=tdtype, device=dev)
    ready = None
    if overlap and coefs.device.type == "cuda":
        main = torch.cuda.current_stream(coefs.device)
        side = _side_stream(coefs.device)
        side.wait_stream(main)  # the coefficients were produced on the caller's stream
        with torch.cuda.stream(side):
            dist.broadcast(coefs, src=gsrc, group=group)
            ready = torch.cuda.Event()
            ready.record(side)
        coefs.record_stream(side)
    else:
        dist.broadcast(coefs, src=gsrc, group=group)
    if rank == src:
        out = itp
    else:
        dev_index = coefs.device.index if coefs.device.type == "cuda" else None
        out = core.BSplineInterpolation(coefs, size, parent_len, its, dev_index, parent_first=parent_first)
    if ready is not None:
        out.ready_event = ready
    return out


def _cuda_solver(device_index: int) -> Callable:
    def solve(t, julia_size, its, src=None):
        # segmentation off: a line is then solved by the same arithmetic whatever slab it sits in, so the sharded
        # result has the bits of the single-GPU one at every world size (include/b200itp.h)
        L = core._lib.lib()
        prev = L.b200itp_prefilter_set_segmentation(0)
        try:
            # raises without the CUDA library; with `src` the first pass reads the samples and writes `t`
            core.prefilter_(t.view(-1), tuple(julia_size), its, device_index, src=None if src is None else src.reshape(-1))
        finally:
            L.b200itp_prefilter_set_segmentation(prev)
    solve.reads_src = True
    return solve


def prefilter_sharded(A_slab, it, group=None, solve: Optional[Callable] = None, counts=None):
    """Slab-sharded `interpolate(A, it)` for a 3-D array.

    `A_slab`: this rank's z-slab of the samples, torch tensor of shape (nz_local, ny, nx) (z-planes
    `shard_bounds(nz, world, rank)` of the full array), on the rank's GPU.
    Returns `(coefs_yshard, meta)`: the rank's y-shard of the padded coefficient array, shape (pz, py_local, px),
    and the metadata `allgather_coefficients` needs.  Pass order and systems are the reference's
    (prefiltering.jl:49-59): x, then y, then z.

    Data movement: the x and y passes run in ONE library call on the slab (the first reads `A_slab`, so the samples are
    never copied first); the all-to-all send buffer is one strided copy of the slab into peer-major order
    (`view(pz_l, world, py_l, px).permute(1, 0, 2, 3)`); the blocks arrive in rank order == z order, so the receive
    buffer IS the y-shard.  No zero-filled staging arrays unless the degree pads.
    """
    torch, dist = _torch(), _dist()
    world, rank = _world(group)
    if A_slab.dim() != 3:
        raise ValueError("prefilter_sharded: 3-D arrays only (smaller problems: replicas, SURVEY §8e)")
    its = core._its_tuple(it, 3)
    pad = [1 if needs_padding(i) else 0 for i in its]
    nzl, ny, nx = (int(s) for s in A_slab.shape)
    if counts is None:
        # slab sizes of the other ranks: one small tensor all-gather (an object all-gather pickles on the CPU and costs
        # milliseconds); callers that split with `shard_bounds` can pass `counts` and skip it
        mine = torch.tensor([nzl], dtype=torch.int64, device=A_slab.device)
        allc = torch.empty(world, dtype=torch.int64, device=A_slab.device)
        dist.all_gather_into_tensor(allc, mine, group=group)
        counts = [int(v) for v in allc.tolist()]
    counts = [int(c) for c in counts]
    nz = int(sum(counts))
    px, py, pz = nx + 2 * pad[0], ny + 2 * pad[1], nz + 2 * pad[2]
    if solve is None:
        if A_slab.device.type != "cuda":
            raise RuntimeError("prefilter_sharded: the product path needs CUDA tensors (there is no CPU fallback)")
        solve = _cuda_solver(A_slab.device.index)
    skip = BSpline(Linear())

    # ---- local slab: padded copy (copy_with_padding, prefiltering.jl:17-28; the zero z-planes belong to the end
    #      ranks) only when some axis pads; otherwise the first pass reads the samples directly
    lead = pad[2] if rank == 0 else 0
    trail = pad[2] if rank == world - 1 else 0
    pzl = nzl + lead + trail
    pz_counts = [c + (pad[2] if r == 0 else 0) + (pad[2] if r == world - 1 else 0) for r, c in enumerate(counts)]
    src = None
    if any(pad):
        slab = torch.zeros((pzl, py, px), dtype=A_slab.dtype, device=A_slab.device)
        slab[lead:lead + nzl, pad[1]:pad[1] + ny, pad[0]:pad[0] + nx] = A_slab
    elif getattr(solve, "reads_src", False) and A_slab.is_contiguous():
        slab = torch.empty_like(A_slab)
        src = A_slab
    else:
        slab = A_slab.clone()

    # ---- passes along x and y: local, one call
    if pzl > 0:
        if src is not None:
            solve(slab, (px, py, pzl), (its[0], its[1], skip), src=src)
        else:
            solve(slab, (px, py, pzl), (its[0], its[1], skip))

    # ---- one all-to-all: z-slabs -> y-slabs
    ybounds = [shard_bounds(py, world, r) for r in range(world)]
    ylo, yhi = ybounds[rank]
    pyl = yhi - ylo
    if py % world == 0 and pzl > 0:
        send = slab.view(pzl, world, pyl, px).permute(1, 0, 2, 3).contiguous().view(-1)  # peer-major: one strided copy
    else:
        send = torch.cat([slab[:, lo:hi, :].reshape(-1) for lo, hi in ybounds]) if pzl > 0 else slab.reshape(-1)
    in_splits = [pzl * (hi - lo) * px for lo, hi in ybounds]
    out_splits = [pz_counts[r] * pyl * px for r in range(world)]
    recv = torch.empty(int(sum(out_splits)), dtype=slab.dtype, device=slab.device)
    dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
    yshard = recv.view(pz, pyl, px)  # blocks arrive in rank order == z order

    # ---- pass along z: local
    if pyl > 0:
        solve(yshard, (px, pyl, pz), (skip, skip, its[2]))
    meta = {"size": (px, py, pz), "parent_len": (nx, ny, nz), "its": its, "ybounds": ybounds}
    return yshard, meta


def allgather_coefficients(yshard, meta, group=None):
    """Replicate the y-sharded coefficient array on every rank.  Equal shards: ONE all-gather into a rank-major buffer
    and one strided copy into the (z, y, x) layout; uneven shards: one broadcast per shard (fine on NCCL and gloo).
    Returns the flat column-major coefficient tensor."""
    torch, dist = _torch(), _dist()
    world, rank = _world(group)
    px, py, pz = meta["size"]
    sizes = [hi - lo for lo, hi in meta["ybounds"]]
    if len(set(sizes)) == 1 and sizes[0] > 0:
        pyl = sizes[0]
        gath = torch.empty((world, pz, pyl, px), dtype=yshard.dtype, device=yshard.device)
        dist.all_gather_into_tensor(gath.view(-1), yshard.contiguous().view(-1), group=group)
        return gath.permute(1, 0, 2, 3).reshape(-1)  # (pz, world, pyl, px) -> (pz, py, px): the one copy
    full = torch.empty((pz, py, px), dtype=yshard.dtype, device=yshard.device)
    for r, (lo, hi) in enumerate(meta["ybounds"]):
        if hi == lo:
            continue
        block = yshard.contiguous() if r == rank else torch.empty((pz, hi - lo, px), dtype=yshard.dtype,
                                                                   device=yshard.device)
        dist.broadcast(block, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        full[:, lo:hi, :] = block
    return full.view(-1)


def interpolate_sharded(A_slab, it, group=None, solve: Optional[Callable] = None) -> core.BSplineInterpolation:
    """`interpolate(A, it)` with the prefilter sharded over the ranks of `group`; every rank returns the same
    (replicated) interpolant, ready for query-sharded evaluation."""
    yshard, meta = prefilter_sharded(A_slab, it, group=group, solve=solve)
    coefs = allgather_coefficients(yshard, meta, group=group)
    dev_index = coefs.device.index if coefs.device.type == "cuda" else None
    return core.BSplineInterpolation(coefs, meta["size"], meta["parent_len"], meta["its"], dev_index)


# ----------------------------------------------------------------------------------------------------------------------
# The C ABI's own multi-GPU entry points (include/b200itp.h: b200itp_comm_init, b200itp_bcast_coefs,
# b200itp_prefilter_sharded): ONE host process drives several devices — what a Julia session does — NCCL inside the
# library.  The torch.distributed functions above are the one-process-per-GPU form of the same sequence.
class Comm:
    """`b200itp_comm`: NCCL communicator over `devices` (CUDA device indices) owned by this process."""

    def __init__(self, devices: Sequence[int]):
        import ctypes as C
        self._C = C
        self.devices = [int(d) for d in devices]
        self._L = core._lib.lib()
        self._h = C.c_void_p(None)
        devs = (C.c_int32 * len(self.devices))(*self.devices)
        core._lib.check(self._L.b200itp_comm_init(len(self.devices), devs, C.byref(self._h)), "b200itp_comm_init")

    def __len__(self):
        return len(self.devices)

    def close(self):
        if self._h:
            self._L.b200itp_comm_destroy(self._h)
            self._h = self._C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bcast_coefs(self, buffers, root=0):
        """Replicate `buffers[root]` (flat torch CUDA tensors, one per device, equal sizes) on every device."""
        C = self._C
        nbytes = buffers[root].numel() * buffers[root].element_size()
        ptrs = (C.c_void_p * len(buffers))(*[b.data_ptr() for b in buffers])
        _torch().cuda.synchronize()
        core._lib.check(self._L.b200itp_bcast_coefs(self._h, root, ptrs, nbytes), "b200itp_bcast_coefs")

    def interpolate_sharded(self, slabs, it):
        """`interpolate(A, it)` with device i holding the z-slab `slabs[i]` (torch CUDA tensor of shape (nz_i, ny, nx):
        the memory of the Julia array's slab).  Returns one replicated `BSplineInterpolation` per device and the device
        times (ms) up to the end of the z pass and including the replication."""
        C = self._C
        torch = _torch()
        its = core._its_tuple(it, 3)
        nz = int(sum(int(s.shape[0]) for s in slabs))
        ny, nx = int(slabs[0].shape[1]), int(slabs[0].shape[2])
        code = core.F32 if slabs[0].dtype == torch.float32 else core.F64
        outs = [torch.empty(nx * ny * nz, dtype=slabs[0].dtype, device=s.device) for s in slabs]
        size = (C.c_int64 * 3)(nx, ny, nz)
        deg = (C.c_int32 * 3)(*[i.degree.code for i in its])
        bc = (C.c_int32 * 3)(*[i.degree.bc.code for i in its])
        gt = (C.c_int32 * 3)(*[i.degree.bc.gt.code for i in its])
        sp = (C.c_void_p * len(slabs))(*[s.contiguous().data_ptr() for s in slabs])
        op = (C.c_void_p * len(slabs))(*[o.data_ptr() for o in outs])
        ms = (C.c_double * 2)(0.0, 0.0)
        for s in slabs:
            torch.cuda.synchronize(s.device)
        core._lib.check(self._L.b200itp_prefilter_sharded(self._h, sp, op, code, size, deg, bc, gt, ms),
                        "b200itp_prefilter_sharded")
        itps = [core.BSplineInterpolation(o, (nx, ny, nz), (nx, ny, nz), its, o.device.index) for o in outs]
        return itps, (ms[0], ms[1])

"""Host logic of the multi-GPU path on CPU: `gloo`, world_size 2 and 3.

The collectives (all-to-all transpose between the y and z passes, coefficient replication, query sharding) are the
product's; the per-axis solves are injected from the CPU oracle, because no GPU exists here (on a GPU box the same
functions run with the CUDA library as the solver, see tests/test_gpu_parity.py::test_sharded_single_rank)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, spec_name, out_dir):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import b200_loader
    m = b200_loader.load()
    import oracle as orc  # test infrastructure: stands in for the CUDA axis solver
    from interpolations_jl_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        it = _spec(m, spec_name)
        rng = np.random.default_rng(1234)
        A = rng.standard_normal(shape)  # Julia shape (nx, ny, nz), same on every rank
        nz = shape[2]
        lo, hi = par.shard_bounds(nz, world, rank)
        # torch layout (nz, ny, nx), C-contiguous == Julia column-major
        slab = torch.from_numpy(np.ascontiguousarray(A.transpose(2, 1, 0)[lo:hi]))

        def solve(t, julia_size, its):
            flat = t.view(-1).numpy()
            orc.prefilter_inplace(flat, tuple(julia_size), its)

        itp = par.interpolate_sharded(slab, it, solve=solve)
        got = itp.coefs.numpy().reshape(itp.size, order="F")
        ref = orc.prefilter(A, it)
        assert got.shape == ref.shape
        np.save(os.path.join(out_dir, f"err_{rank}.npy"), np.array([np.abs(got - ref).max()]))
        # replication by broadcast: rank 0's interpolant reaches everybody
        itp0 = par.broadcast_interpolant(itp if rank == 0 else None, src=0)
        assert itp0.size == itp.size and itp0.parent_len == itp.parent_len
        assert np.array_equal(itp0.coefs.numpy(), itp.coefs.numpy())
        # query sharding: concatenation in rank order restores the caller's order
        xs = [np.arange(101, dtype=np.float64) + d for d in range(3)]
        mine = par.shard_queries(xs, world, rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, [c.tolist() for c in mine])
        for d in range(3):
            cat = np.concatenate([np.asarray(g[d]) for g in gathered])
            assert np.array_equal(cat, xs[d])
    finally:
        dist.destroy_process_group()


def _spec(m, name):
    return {
        "cubic_periodic": m.BSpline(m.Cubic(m.Periodic(m.OnGrid()))),
        "cubic_line": m.BSpline(m.Cubic(m.Line(m.OnGrid()))),
        "quadratic_reflect": m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))),
        "mixed": (m.BSpline(m.Cubic(m.Flat(m.OnCell()))), m.BSpline(m.Linear()), m.BSpline(m.Quadratic(m.Free(m.OnGrid())))),
    }[name]


@pytest.mark.parametrize("world,shape,spec", [(2, (12, 10, 9), "cubic_periodic"), (2, (9, 7, 8), "cubic_line"),
                                               (3, (8, 11, 10), "quadratic_reflect"), (2, (10, 6, 7), "mixed")])
def test_sharded_prefilter_matches_single_process(tmp_path, orc, world, shape, spec):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, spec, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err = float(np.load(tmp_path / f"err_{r}.npy")[0])
        assert err == 0.0, (r, err)  # same solver, same per-line arithmetic: bit-identical to the unsharded prefilter


def test_shard_bounds_cover_range(pkg):
    from interpolations_jl_b200 import parallel as par
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [par.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_shard_bounds_lead_gives_rank0_a_lighter_contiguous_chunk(pkg):
    """Rank 0 prefilters and feeds the broadcast: `shard_bounds_lead` hands it `lead_deficit` fewer queries than each
    other rank; the chunks stay contiguous, ordered, and cover the range."""
    from interpolations_jl_b200 import parallel as par
    for n in (0, 7, 1000, 10**9):
        for w in (1, 2, 3, 8):
            for d in (0, 10, 14_000_000):
                b = [par.shard_bounds_lead(n, w, r, d) for r in range(w)]
                assert b[0][0] == 0 and b[-1][1] == n
                assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
                sizes = [hi - lo for lo, hi in b]
                assert all(s >= 0 for s in sizes)
                if w > 1:
                    assert max(sizes[1:]) - min(sizes[1:]) <= 1
                    if d > 0 and n >= w * d:
                        assert d - 1 <= sizes[1] - sizes[0] <= d + 1, (n, w, d, sizes)
                if d == 0:
                    assert b == [par.shard_bounds(n, w, r) for r in range(w)]

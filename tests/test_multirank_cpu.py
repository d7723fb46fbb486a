"""world_size-2 gloo test of the N > 1 path on CPU: range sharding, all-gather of the XYZZ
partials, host combination.  Per-rank partials come from the CPU oracle here (no GPU in this
container); on the GPU box the same functions are driven by bench.py with NCCL."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, curve, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import cpu_oracle
    from tachyon_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = cpu_oracle.CurveOracle(curve)
    bases, scalars = o.generate_points(900, n), o.generate_scalars(901, n, "witness")
    lo, hi = sharding.shard_range(n, rank, world)
    part = o.msm(bases[lo:hi], scalars[lo:hi], threads=2)
    parts = sharding.gather_partials(part, world)
    total = sharding.combine_partials(curve, list(parts))
    if rank == 0:
        want = o.msm_affine(bases, scalars, threads=2)
        q.put(bool((o.xyzz_to_affine(total) == want).all()) and bool((parts[rank] == part).all()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("curve,n", [("bn254", 1001), ("bls12_381", 300)])
def test_two_rank_sharded_msm(curve, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, curve, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_cover():
    from tachyon_b200 import sharding
    for n in (0, 1, 7, 1 << 24, 1000003):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_deal_msms_covers_every_point_once():
    from tachyon_b200 import sharding
    sizes = [(1 << 10) + 1, 1 << 10, 1000, 1 << 10]     # A, B1, L, H of a small circuit
    for world in (1, 2, 3, 4, 8):
        for split in ("auto", "msm", "range"):
            work = sharding.deal_msms(sizes, world, split)
            assert len(work) == world
            for j, n in enumerate(sizes):
                pieces = sorted((lo, hi) for items in work for (m, lo, hi) in items if m == j)
                assert pieces[0][0] == 0 and pieces[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    # 4 ranks: one MSM each; 8 ranks: two ranks per MSM; 2 ranks: two MSMs each
    assert [len(w) for w in sharding.deal_msms(sizes, 4)] == [1, 1, 1, 1]
    assert [w[0][0] for w in sharding.deal_msms(sizes, 8)] == [0, 0, 1, 1, 2, 2, 3, 3]
    assert [[m for m, _, _ in w] for w in sharding.deal_msms(sizes, 2)] == [[0, 2], [1, 3]]


def _set_worker(rank, world, port, curve, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import cpu_oracle
    from tachyon_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = cpu_oracle.CurveOracle(curve)
    sizes = [301, 300, 250]
    sets = [(o.generate_points(910 + j, n), o.generate_scalars(920 + j, n, "witness" if j < 2 else "uniform"))
            for j, n in enumerate(sizes)]
    ok = True
    for split in ("msm", "range"):
        partials = np.stack([o.xyzz_zero() for _ in sizes])
        for j, lo, hi in sharding.deal_msms(sizes, world, split)[rank]:
            partials[j] = o.msm(sets[j][0][lo:hi], sets[j][1][lo:hi], threads=2)
        total = sharding.combine_set(curve, sharding.gather_set_partials(partials, world))
        if rank == 0:
            for j in range(len(sizes)):
                ok = ok and bool((o.xyzz_to_affine(total[j]) == o.msm_affine(*sets[j], threads=2)).all())
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_msm_set():
    """The Groth16 MSM set (several independent MSMs) over 2 ranks, both ways of dealing it."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_set_worker, args=(r, 2, port, "bn254", q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

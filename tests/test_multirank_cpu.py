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

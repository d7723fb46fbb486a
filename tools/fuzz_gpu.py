"""Randomised parity sweep on a GPU (development aid, uses the CPU oracle as the checker):
random sizes, window sizes, window balancing, reduction / sort modes, running-sum block lengths,
range counts, ladder placement and window groups, task lengths, point staging, registered bases
with and without the precomputed table, scalar distributions, host and device inputs, both G1
curves (G2 with --g2).
    python tools/fuzz_gpu.py [cases] [seed] [--g2]"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cpu_oracle
from tachyon_b200 import msm

cases = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1
curves = ["bn254", "bls12_381"] + (["bn254_g2", "bls12_381_g2"] if "--g2" in sys.argv else [])
rng = random.Random(seed)
oracles = {c: cpu_oracle.CurveOracle(c) for c in curves}
ctxs = {c: msm.MSMGpu(c) for c in curves}
t0 = time.time()
bad = 0
for case in range(cases):
    c = rng.choice(curves)
    o, ctx = oracles[c], ctxs[c]
    big = rng.random() < 0.15
    n = rng.randint(1 << 15, 1 << 17) if big else rng.choice([rng.randint(1, 70), rng.randint(70, 3000), rng.randint(3000, 20000)])
    if c.endswith("_g2"):
        n = min(n, 4000)
    dist = rng.choice(["uniform", "uniform", "witness", "non_uniform"])
    opts = {
        "window_bits": rng.choice([0, 0, 0] + list(range(4, 19))),
        "balance": rng.choice([1, 1, 0]),
        "reduce_mode": rng.choice([1, 1, 0, 2]),
        "reduce_roll": rng.choice([-1, -1, 0, 1, 2]),
        "reduce_inline": rng.choice([-1, -1, 0, 1]),
        "acc_lockstep": rng.choice([-1, 0, 1]),
        "sort_mode": rng.choice([-1, -1, 0, 1]),
        "ranges": rng.choice([0, 0, 1, 2, 3, 7]),
        "sample_scalars": rng.choice([1, 1, 0]),
        "device_ladder": rng.choice([0, 0, 1]),
        "low_windows": rng.choice([-1, 0, 1, 2, 3, 9]),
        "level_fill": rng.choice([0, 0, 48, 96, 3000]),      # running-sum block lengths 4 .. 64
        "stage_points": rng.choice([0, 0, 1]),
        "acc_variant": rng.choice([-1, -1, 0, 1, 2, 3]),
        "segment": rng.choice([0, 0, 0, 16, 32]),
    }
    for k, v in opts.items():
        ctx.set_option(k, v)
    s = rng.randrange(1 << 30)
    bases, scalars = o.generate_points(s, n), o.generate_scalars(s + 1, n, dist)
    want = o.msm_affine(bases, scalars)
    mode = rng.random()
    if mode < 0.2:
        # registered bases, with or without the table of window multiples, through the batch call
        ctx.set_option("precompute", rng.choice([0, 1]))
        ctx.register_bases(bases)
        m2 = rng.randint(1, n)
        out = msm.batch_normalize(c, ctx.commit_batch([scalars, scalars[:m2]], [n, m2]))
        ctx.set_option("precompute", 0)
        ok = bool((out[0] == np.asarray(want).reshape(-1)).all()) and \
            bool((out[1] == np.asarray(o.msm_affine(bases[:m2], scalars[:m2])).reshape(-1)).all())
        if not ok:
            bad += 1
            print("MISMATCH (registered)", c, n, m2, dist, opts, "seed", s, flush=True)
        continue
    if rng.random() < 0.5:
        got = ctx.affine_msm(bases, scalars)
        where = "host"
    else:
        db = torch.from_numpy(bases.view(np.int64)).cuda()
        ds = torch.from_numpy(scalars.view(np.int64)).cuda()
        got = ctx.affine_msm(db.data_ptr(), ds.data_ptr(), n)
        where = "device"
    ok = bool((o.jacobian_to_affine(got) == want).all())
    if not ok:
        bad += 1
        print("MISMATCH", c, n, dist, where, opts, "seed", s, flush=True)
print("%d cases, %d mismatches, %.1f s" % (cases, bad, time.time() - t0), flush=True)
sys.exit(1 if bad else 0)

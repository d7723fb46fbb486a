"""Randomised parity sweep on a GPU (development aid, uses the CPU oracle as the checker):
random sizes, window sizes, window balancing, reduction / sort modes, range counts and scalar
distributions, host and device inputs, both G1 curves (G2 with --g2).
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
        "reduce_mode": rng.choice([1, 1, 0]),
        "sort_mode": rng.choice([-1, -1, 0, 1]),
        "ranges": rng.choice([0, 0, 1, 2, 3, 7]),
        "sample_scalars": rng.choice([1, 1, 0]),
    }
    for k, v in opts.items():
        ctx.set_option(k, v)
    s = rng.randrange(1 << 30)
    bases, scalars = o.generate_points(s, n), o.generate_scalars(s + 1, n, dist)
    want = o.msm_affine(bases, scalars)
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

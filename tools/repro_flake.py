"""Counts how often a registered-table MSM batch goes wrong under option variations (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cpu_oracle
from tachyon_b200 import msm

c, n, m2, s = "bls12_381", 11708, 10609, 617450610
o = cpu_oracle.CurveOracle(c)
bases, scalars = o.generate_points(s, n), o.generate_scalars(s + 1, n, "witness")
uni = o.generate_scalars(s + 2, n, "uniform")
want = {"witness": np.asarray(o.msm_affine(bases, scalars)).reshape(-1), "uniform": np.asarray(o.msm_affine(bases, uni)).reshape(-1)}
base = {'window_bits': 18, 'balance': 1, 'reduce_mode': 1, 'sort_mode': -1, 'ranges': 1, 'sample_scalars': 1,
        'device_ladder': 1, 'low_windows': -1, 'level_fill': 0, 'stage_points': 0, 'segment': 0}
variants = [("base", {}), ("host ladder", {"device_ladder": 0}), ("wb 16", {"window_bits": 16}), ("wb 14", {"window_bits": 14}),
            ("level_fill 3000", {"level_fill": 3000}), ("reduce_mode 0", {"reduce_mode": 0}), ("segment 16", {"segment": 16})]
for name, delta in variants:
    for pre in (1, 0):
        for dist, sc in (("witness", scalars), ("uniform", uni)):
            bad_first = bad_single = 0
            reps = 12
            for rep in range(reps):
                ctx = msm.MSMGpu(c)
                for k, v in {**base, **delta}.items():
                    ctx.set_option(k, v)
                ctx.set_option("precompute", pre)
                ctx.register_bases(bases)
                out = msm.batch_normalize(c, ctx.commit_batch([sc, sc[:m2]], [n, m2]))
                bad_first += not bool((out[0] == want[dist]).all())
                out = msm.batch_normalize(c, ctx.commit_batch([sc], [n]))
                bad_single += not bool((out[0] == want[dist]).all())
                ctx.close()
            print(f"{name:16s} precompute {pre} {dist:8s}: first-of-batch wrong {bad_first}/{reps}, single wrong {bad_single}/{reps}", flush=True)

"""Replays fuzz mismatches with one option flipped at a time (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cpu_oracle
from tachyon_b200 import msm

def run(c, n, m2, dist, opts, s, registered, pre):
    o = cpu_oracle.CurveOracle(c)
    ctx = msm.MSMGpu(c)
    for k, v in opts.items():
        ctx.set_option(k, v)
    bases, scalars = o.generate_points(s, n), o.generate_scalars(s + 1, n, dist)
    want = np.asarray(o.msm_affine(bases, scalars)).reshape(-1)
    if registered:
        ctx.set_option("precompute", pre)
        ctx.register_bases(bases)
        out = msm.batch_normalize(c, ctx.commit_batch([scalars, scalars[:m2]], [n, m2]))
        w2 = np.asarray(o.msm_affine(bases[:m2], scalars[:m2])).reshape(-1)
        ok = (bool((out[0] == want).all()), bool((out[1] == w2).all()))
    else:
        ok = bool((np.asarray(o.jacobian_to_affine(ctx.affine_msm(bases, scalars))).reshape(-1) == want).all())
    ctx.close()
    return ok

cases = [
 ("bls12_381", 11708, 10609, "witness", {'window_bits': 18, 'balance': 1, 'reduce_mode': 0, 'sort_mode': -1, 'ranges': 1, 'sample_scalars': 1, 'device_ladder': 1, 'low_windows': -1, 'level_fill': 96, 'stage_points': 0, 'segment': 0}, 617450610, True),
 ("bls12_381", 13195, 1461, "witness", {'window_bits': 5, 'balance': 1, 'reduce_mode': 1, 'sort_mode': -1, 'ranges': 2, 'sample_scalars': 1, 'device_ladder': 1, 'low_windows': 9, 'level_fill': 48, 'stage_points': 0, 'segment': 0}, 657950901, True),
]
for c, n, m2, dist, opts, s, reg in cases:
    for pre in (0, 1):
        print(c, n, "precompute", pre, "base:", run(c, n, m2, dist, opts, s, reg, pre), flush=True)
        for k, alt in (("device_ladder", 0), ("low_windows", 0), ("level_fill", 0), ("reduce_mode", 1), ("ranges", 1), ("window_bits", 9)):
            if opts[k] == alt:
                continue
            o2 = dict(opts); o2[k] = alt
            print("   flip", k, "->", alt, run(c, n, m2, dist, o2, s, reg, pre), flush=True)
        print("   not registered:", run(c, n, m2, dist, opts, s, False, 0), flush=True)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cpu_oracle
from tachyon_b200 import msm
c, s = "bls12_381", 617450610
o = cpu_oracle.CurveOracle(c)
N = 11708
bases = o.generate_points(s, N)
uni = o.generate_scalars(s + 2, N, "uniform")
def bad_count(n, sc, wb, reps=8, extra=None):
    want = np.asarray(o.msm_affine(bases[:n], sc[:n])).reshape(-1)
    bad = 0
    for rep in range(reps):
        ctx = msm.MSMGpu(c)
        ctx.set_option("window_bits", wb)
        for k, v in (extra or {}).items():
            ctx.set_option(k, v)
        ctx.set_option("precompute", 1)
        ctx.register_bases(bases[:n].copy())
        bad += not bool((msm.batch_normalize(c, ctx.commit_batch([sc[:n].copy()], [n]))[0] == want).all())
        ctx.close()
    return bad
for n in (7000, 8000, 8192, 8193, 9000, 10000, 11000, 11708):
    print("n", n, "wb14 wrong", bad_count(n, uni, 14), "/8   wb9 wrong", bad_count(n, uni, 9), "/8", flush=True)
suffix = uni.copy(); suffix[:6000] = 0
prefix = uni.copy(); prefix[6000:] = 0
print("suffix-only scalars wrong", bad_count(N, suffix, 14), "/8; prefix-only", bad_count(N, prefix, 14), "/8", flush=True)
print("sort_mode 0:", bad_count(N, uni, 14, extra={"sort_mode": 0}), " ranges 2:", bad_count(N, uni, 14, extra={"ranges": 2}),
      " ranges 4:", bad_count(N, uni, 14, extra={"ranges": 4}), flush=True)
# bn254 same size
c = "bn254"; o = cpu_oracle.CurveOracle(c); bases = o.generate_points(s, N); uni = o.generate_scalars(s + 2, N, "uniform")
print("bn254 n=11708 wb14 wrong", bad_count(N, uni, 14), "/8", flush=True)

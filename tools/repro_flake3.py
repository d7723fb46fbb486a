"""Table wrong, or pipeline nondeterministic? (debug aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cpu_oracle
from tachyon_b200 import msm

c, s = "bls12_381", 617450610
o = cpu_oracle.CurveOracle(c)
for n in (11708, 6000, 5000, 4200):
    bases = o.generate_points(s, n)
    uni = o.generate_scalars(s + 2, n, "uniform")
    want = np.asarray(o.msm_affine(bases, uni)).reshape(-1)
    for rep in range(6):
        ctx = msm.MSMGpu(c)
        ctx.set_option("window_bits", 14)
        ctx.set_option("precompute", 1)
        ctx.register_bases(bases)
        outs = [msm.batch_normalize(c, ctx.commit_batch([uni], [n]))[0] for _ in range(4)]
        ok = [bool((x == want).all()) for x in outs]
        same = all((x == outs[0]).all() for x in outs)
        # prefix MSMs over the same table
        pre = []
        for m in (4096, 4097, 5000):
            if m <= n:
                w2 = np.asarray(o.msm_affine(bases[:m], uni[:m])).reshape(-1)
                pre.append((m, bool((msm.batch_normalize(c, ctx.commit_batch([uni[:m]], [m]))[0] == w2).all())))
        print(n, "rep", rep, "ok", ok, "all four identical", same, "prefixes", pre, flush=True)
        ctx.close()

"""Which ingredient makes the precomputed-table MSM flaky (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cpu_oracle
from tachyon_b200 import msm

c, s = "bls12_381", 617450610
o = cpu_oracle.CurveOracle(c)
def trial(name, n, bases, sc, opts, reps=10):
    want = np.asarray(o.msm_affine(bases, sc)).reshape(-1)
    bad = 0
    outs = set()
    for rep in range(reps):
        ctx = msm.MSMGpu(c)
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_option("precompute", 1)
        ctx.register_bases(bases)
        out = msm.batch_normalize(c, ctx.commit_batch([sc], [n]))
        bad += not bool((out[0] == want).all())
        outs.add(out[0].tobytes())
        ctx.close()
    print(f"{name:40s}: wrong {bad}/{reps}, distinct results {len(outs)}", flush=True)

n = 11708
chain = o.generate_points(s, n)
spread = o.generate_points(s, n * 64)[::64].copy()
uni = o.generate_scalars(s + 2, n, "uniform")
base = {'window_bits': 14, 'device_ladder': 0}
trial("chain bases", n, chain, uni, base)
trial("spread bases (no duplicates)", n, spread, uni, base)
trial("chain, aggregate 0", n, chain, uni, {**base, "aggregate": 0})
trial("chain, segment 1024", n, chain, uni, {**base, "segment": 1024})
trial("chain, segment 16", n, chain, uni, {**base, "segment": 16})
trial("chain, reduce_mode 0", n, chain, uni, {**base, "reduce_mode": 0})
for nn in (4096, 4000, 2000, 1000):
    trial(f"chain n={nn}", nn, chain[:nn].copy(), uni[:nn].copy(), base)
trial("chain n=4096 wb 9", 4096, chain[:4096].copy(), uni[:4096].copy(), {**base, "window_bits": 9})
trial("chain n=4096 wb 12", 4096, chain[:4096].copy(), uni[:4096].copy(), {**base, "window_bits": 12})

"""Does the order in which the grow-only workspace was allocated change the memory-bound stages?
(sort and running-sum level of a 2^24-point MSM after smaller MSMs on the same context)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tachyon_b200 import _lib, msm

curve = "bn254"
fq = _lib.element_limbs(curve)
nmax = 1 << 24
bases = torch.empty((nmax, 2 * fq), dtype=torch.int64, device="cuda")
scalars = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
msm.generate_bases_device(curve, 1, nmax, bases.data_ptr())
msm.generate_scalars_device(curve, 2, nmax, scalars.data_ptr(), "uniform")
torch.cuda.synchronize()


def run(ctx, logs, label):
    for lg in logs:
        for _ in range(3):
            ctx.msm_xyzz(bases.data_ptr(), scalars.data_ptr(), 1 << lg)
        t = ctx.last_timing()
        print("%-34s 2^%d total %.3f sort %.3f acc %.3f reduce %.3f" %
              (label, lg, t["total_ms"], t["sort_ms"], t["accumulate_ms"], t["reduce_ms"]), flush=True)


with msm.MSMGpu(curve, degree=20) as ctx:
    run(ctx, [24], "degree 20, then 24")
with msm.MSMGpu(curve, degree=20) as ctx:
    run(ctx, [16, 20, 24], "degree 20, then 16, 20, 24")
    ctx.set_option("release_workspace", 1)
    run(ctx, [24], "... workspace released, 24")
with msm.MSMGpu(curve, degree=24) as ctx:
    run(ctx, [24, 16, 24], "degree 24")
with msm.MSMGpu(curve, degree=16) as ctx:
    run(ctx, [16, 18, 20, 22, 24], "degree 16, growing")

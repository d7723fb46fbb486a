"""Quick on-GPU look: IMAD peak and stage timings for a few sizes (development aid)."""
import sys
import time

import numpy as np
import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tachyon_b200 import msm

curve = sys.argv[1] if len(sys.argv) > 1 else "bn254"
logs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 18, 20, 22]
opts = dict(kv.split("=") for kv in sys.argv[3:])
from tachyon_b200 import _lib
fq = _lib.element_limbs(curve)
print("imad peak variant0 %.3e  variant1 %.3e products/s" % (msm.imad_peak(0, 0), msm.imad_peak(0, 1)))
nmax = 1 << max(logs)
bases = torch.empty((nmax, 2 * fq), dtype=torch.int64, device="cuda")
scalars = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
t0 = time.time()
msm.generate_bases_device(curve, 1, nmax, bases.data_ptr())
msm.generate_scalars_device(curve, 2, nmax, scalars.data_ptr(), opts.get("dist", "uniform"))
torch.cuda.synchronize()
print("generated 2^%d inputs in %.2fs" % (max(logs), time.time() - t0))
ctx = msm.MSMGpu(curve)
for k, v in opts.items():
    if k != "dist":
        ctx.set_option(k, int(v))
for lg in logs:
    n = 1 << lg
    for it in range(4):
        t0 = time.time()
        ctx.msm_xyzz(bases.data_ptr(), scalars.data_ptr(), n)
        wall = (time.time() - t0) * 1e3
        t = ctx.last_timing()
    print("2^%d wall %.3f ms | total %.3f sort %.3f acc %.3f reduce %.3f host %.3f enq %.3f wait %.3f | c=%d W=%d R=%d tasks=%d entries=%d launches=%d low=%d acck=%.3f comb=%.3f"
          % (lg, wall, t["total_ms"], t["sort_ms"], t["accumulate_ms"], t["reduce_ms"], t["host_ms"], t["enqueue_ms"], t["wait_ms"],
             t["window_bits"], t["windows"], t["pair_rounds"], t["tasks"], t["entries"], t["kernel_launches"], t["low_windows"], t["acc_kernel_ms"], t["combine_ms"]))

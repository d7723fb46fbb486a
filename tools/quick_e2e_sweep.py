"""Development aid: end-to-end time of one BN254 2^24 MSM from page-locked host buffers for several
(window size, most ranges) pairs of the automatic host-input pipeline.
    python tools/quick_e2e_sweep.py [curve] [log_n]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tachyon_b200 import _lib, msm

curve = sys.argv[1] if len(sys.argv) > 1 else "bn254"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
fq = _lib.element_limbs(curve)
n = 1 << lg
bases = torch.empty((n, 2 * fq), dtype=torch.int64, device="cuda")
scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
msm.generate_bases_device(curve, 1, n, bases.data_ptr())
msm.generate_scalars_device(curve, 2, n, scalars.data_ptr(), "uniform")
hb = torch.empty((n, 2 * fq), dtype=torch.int64).pin_memory()
hs = torch.empty((n, 4), dtype=torch.int64).pin_memory()
hb.copy_(bases)
hs.copy_(scalars)
torch.cuda.synchronize()
ctx = msm.MSMGpu(curve, degree=lg)
ref = ctx.msm_xyzz(bases.data_ptr(), scalars.data_ptr(), n)
for wb in (0, 19, 20):
    for hr in (4, 5, 6, 8, 10, 12):
        ctx.set_option("window_bits", wb)
        ctx.set_option("host_ranges", hr)
        best = 1e9
        for it in range(5):
            t0 = time.perf_counter()
            out = ctx.msm_xyzz(hb.data_ptr(), hs.data_ptr(), n)
            best = min(best, (time.perf_counter() - t0) * 1e3)
        t = ctx.last_timing()
        print("window_bits=%d host_ranges=%d -> ranges %d c=%d wall %.3f ms | total %.3f h2d %.3f sort %.3f acc %.3f reduce %.3f same_bits=%s"
              % (wb, hr, t["ranges"], t["window_bits"], best, t["total_ms"], t["h2d_ms"], t["sort_ms"],
                 t["accumulate_ms"], t["reduce_ms"], bool((out == ref).all())), flush=True)

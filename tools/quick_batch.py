"""Development aid: batched commitments over registered bases vs one call per MSM.
python tools/quick_batch.py bn254 20 16"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tachyon_b200 import msm

curve = sys.argv[1] if len(sys.argv) > 1 else "bn254"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
count = int(sys.argv[3]) if len(sys.argv) > 3 else 16
from tachyon_b200 import _lib
fq = _lib.element_limbs(curve)
n = 1 << lg
bases = torch.empty((n, 2 * fq), dtype=torch.int64, device="cuda")
msm.generate_bases_device(curve, 1, n, bases.data_ptr())
hs = []
for i in range(count):
    s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    msm.generate_scalars_device(curve, 100 + i, n, s.data_ptr(), "uniform")
    h = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    h.copy_(s)
    hs.append(h)
hb = torch.empty((n, 2 * fq), dtype=torch.int64).pin_memory()
hb.copy_(bases)
torch.cuda.synchronize()
ctx = msm.MSMGpu(curve)
for rep in range(3):
    t0 = time.perf_counter()
    one = [ctx.msm_xyzz(hb.data_ptr(), h.data_ptr(), n) for h in hs]
    t_naive = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    two = [ctx.msm_xyzz(bases.data_ptr(), h.data_ptr(), n) for h in hs]
    t_resident = (time.perf_counter() - t0) * 1e3
    ctx.register_bases(bases.data_ptr(), n)
    t0 = time.perf_counter()
    out = ctx.commit_batch([h.data_ptr() for h in hs], [n] * count)
    t_batch = (time.perf_counter() - t0) * 1e3
    t = ctx.last_timing()
    print("2^%d x %d: host bases+scalars per call %.2f ms | resident bases, one call each %.2f ms | commit_batch %.2f ms "
          "(sort %.2f acc %.2f reduce %.2f host %.2f h2d %.2f)" %
          (lg, count, t_naive, t_resident, t_batch, t["sort_ms"], t["accumulate_ms"], t["reduce_ms"], t["host_ms"], t["h2d_ms"]),
          flush=True)
a = msm.batch_normalize(curve, out)
b = msm.batch_normalize(curve, np.stack(one))
print("batch == per-call:", bool((a == b).all()))

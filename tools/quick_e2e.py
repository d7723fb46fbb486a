"""Development aid: end-to-end time of one MSM from pinned host buffers for several
point-range counts.   python tools/quick_e2e.py bn254 24 1,2,4,8,16 uniform [option=value ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tachyon_b200 import msm

curve = sys.argv[1] if len(sys.argv) > 1 else "bn254"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ranges = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 4, 8, 16]
from tachyon_b200 import _lib
fq = _lib.element_limbs(curve)
dist = sys.argv[4] if len(sys.argv) > 4 else "uniform"
n = 1 << lg
bases = torch.empty((n, 2 * fq), dtype=torch.int64, device="cuda")
scalars = torch.empty((n, 4), dtype=torch.int64, device="cuda")
msm.generate_bases_device(curve, 1, n, bases.data_ptr())
msm.generate_scalars_device(curve, 2, n, scalars.data_ptr(), dist)
hb = torch.empty((n, 2 * fq), dtype=torch.int64).pin_memory()
hs = torch.empty((n, 4), dtype=torch.int64).pin_memory()
hb.copy_(bases)
hs.copy_(scalars)
torch.cuda.synchronize()
pb = hb.numpy().copy()   # pageable copies
ps = hs.numpy().copy()
ctx = msm.MSMGpu(curve)
for kv in sys.argv[5:]:          # engine options, e.g. host_ranges=6 window_bits=19
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ref = ctx.msm_xyzz(bases.data_ptr(), scalars.data_ptr(), n)
for r in ranges:
    ctx.set_option("ranges", r)
    for kind, b, s in (("pinned", hb.data_ptr(), hs.data_ptr()), ("pageable", pb.ctypes.data, ps.ctypes.data),
                       ("dev", bases.data_ptr(), scalars.data_ptr())):
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            out = ctx.msm_xyzz(b, s, n)
            best = min(best, (time.perf_counter() - t0) * 1e3)
        t = ctx.last_timing()
        same = bool((out == ref).all())
        print("2^%d %s ranges=%d(%d) wall %.3f ms | enq %.3f total %.3f h2d %.3f sort %.3f acc %.3f reduce %.3f host %.3f | c=%d tasks=%d same_bits=%s"
              % (lg, kind, r, t["ranges"], best, t["enqueue_ms"], t["total_ms"], t["h2d_ms"], t["sort_ms"], t["accumulate_ms"],
                 t["reduce_ms"], t["host_ms"], t["window_bits"], t["tasks"], same), flush=True)

"""Development aid: wall clock of the FIRST calls of a context (un-warmed, pageable host inputs),
the protocol of `bench.py -k ... `.   python tools/repro_first_call.py uniform 24 16,18,20,20,22"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tachyon_b200 import msm

curve, fq = "bn254", 4
dist = sys.argv[1] if len(sys.argv) > 1 else "uniform"
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 20, 20, 16, 20]
nmax = 1 << max(ks)
b = torch.empty((nmax, 2 * fq), dtype=torch.int64, device="cuda")
sc = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
msm.generate_bases_device(curve, 5, nmax, b.data_ptr())
msm.generate_scalars_device(curve, 6, nmax, sc.data_ptr(), dist)
torch.cuda.synchronize()
hb, hs = b.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64)
del b, sc
t0 = time.perf_counter()
ctx = msm.MSMGpu(curve, degree=degree)
print("create(degree=%d) %.1f ms" % (degree, (time.perf_counter() - t0) * 1e3), flush=True)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
for k in ks:
    n = 1 << k
    t0 = time.perf_counter()
    ctx.affine_msm(hb[:n], hs[:n], n)
    dt = (time.perf_counter() - t0) * 1e3
    t = ctx.last_timing()
    print(k, "wall %.2f ms enq %.2f wait %.2f total %.2f h2d %.2f sort %.2f acc %.2f reduce %.2f ranges %d c=%d"
          % (dt, t["enqueue_ms"], t["wait_ms"], t["total_ms"], t["h2d_ms"], t["sort_ms"], t["accumulate_ms"],
             t["reduce_ms"], t["ranges"], t["window_bits"]), flush=True)

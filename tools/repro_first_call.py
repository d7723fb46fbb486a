import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tachyon_b200 import msm
curve="bn254"; fq=4
nmax=1<<20
b = torch.empty((nmax, 2*fq), dtype=torch.int64, device="cuda")
sc = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
msm.generate_bases_device(curve, 5, nmax, b.data_ptr())
msm.generate_scalars_device(curve, 6, nmax, sc.data_ptr(), sys.argv[1])
torch.cuda.synchronize()
hb, hs = b.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64)
ctx = msm.MSMGpu(curve)
for k in (16, 20, 20, 16, 20):
    n=1<<k
    t0=time.perf_counter()
    ctx.affine_msm(hb[:n], hs[:n], n)
    dt=(time.perf_counter()-t0)*1e3
    t=ctx.last_timing()
    print(k, "wall %.2f ms enq %.2f wait %.2f total %.2f h2d %.2f sort %.2f acc %.2f reduce %.2f ranges %d"%(dt, t["enqueue_ms"], t["wait_ms"], t["total_ms"], t["h2d_ms"], t["sort_ms"], t["accumulate_ms"], t["reduce_ms"], t["ranges"]), flush=True)

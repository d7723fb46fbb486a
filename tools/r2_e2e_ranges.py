"""Per-rank end-to-end time of a 2^24 / world shard from page-locked host memory while all ranks
copy at once, for several point-range counts and both kinds of page-locked memory (run under
torchrun on an 8-GPU box; development aid behind the e2e numbers of DESIGN.md section 6)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tachyon_b200 import msm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = (1 << 24) // world
b = torch.empty((n, 8), dtype=torch.int64, device="cuda")
s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
msm.generate_bases_device("bn254", 5, n, b.data_ptr(), first=rank * n)
msm.generate_scalars_device("bn254", 6, n, s.data_ptr(), "uniform", first=rank * n)
ctx = msm.MSMGpu("bn254", degree=24, device=local)
for wc in (0, 1):
    hb, hs = msm.HostBuffer(n, 8, write_combined=bool(wc)), msm.HostBuffer(n, 4, write_combined=bool(wc))
    torch.from_numpy(hb.array.view(np.int64)).copy_(b)
    torch.from_numpy(hs.array.view(np.int64)).copy_(s)
    torch.cuda.synchronize()
    for ranges in (0, 1, 2, 4, 8, 16):
        ctx.set_option("ranges", ranges)
        ctx.msm_xyzz(hb.ptr, hs.ptr, n)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.msm_xyzz(hb.ptr, hs.ptr, n)
        ms = (time.perf_counter() - t0) * 1e3 / 5
        t = ctx.last_timing()
        v = torch.tensor([ms, t["h2d_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world {world} write_combined {wc} ranges {ranges}({t['ranges']}): e2e {v[0].item():.2f} ms, h2d {v[1].item():.2f} ms "
                  f"({n * 96 / v[1].item() / 1e6:.1f} GB/s per GPU), c={t['window_bits']}", flush=True)
    hb.free(); hs.free()
dist.destroy_process_group()

"""Point-range sharding of one MSM over ranks (one process per GPU) — the split of
tachyon/math/elliptic_curves/msm/algorithms/pippenger/pippenger_adapter.h:82-113 taken across
GPUs instead of threads.  The only exchange is an all-gather of one XYZZ partial per rank."""
import numpy as np

from . import msm


def shard_range(n, rank, world):
    """Contiguous range of points owned by `rank`."""
    return n * rank // world, n * (rank + 1) // world


def gather_partials(partial, world, device=None, out=None):
    """all_gather of this rank's XYZZ partial (4, fq_limbs) uint64 -> (world, 4, fq_limbs).
    NCCL when `device` is a CUDA device, gloo on CPU tensors otherwise."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return partial[None]
    flat = torch.from_numpy(np.ascontiguousarray(partial).view(np.int64).reshape(-1))
    if device is not None:
        flat = flat.to(device, non_blocking=True)
    if out is None:
        out = torch.empty((world * flat.numel(),), dtype=torch.int64, device=flat.device)
    dist.all_gather_into_tensor(out.view(-1), flat)
    return out.cpu().numpy().view(np.uint64).reshape((world,) + partial.shape)


def combine_partials(curve, parts):
    """Host-side sum of the per-rank partials (pippenger_adapter.h:110-113)."""
    total = parts[0]
    for g in range(1, len(parts)):
        total = msm.xyzz_add(curve, total, parts[g])
    return total

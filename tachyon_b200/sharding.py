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


def deal_msms(sizes, world, split="auto"):
    """Work of every rank for a SET of independent MSMs (the A, B1, L, H queries of one
    Groth16 proof, zk/r1cs/groth16/prove.h:100-131): list over ranks of (msm, lo, hi) items.

    split "msm":   whole MSMs are dealt to ranks; with more ranks than MSMs (and a multiple),
                   world // len(sizes) ranks share one MSM by point range.
    split "range": every MSM is cut into `world` point ranges, each rank takes one of each.
    "auto" = "msm" when the ranks divide evenly among the MSMs or the MSMs among the ranks,
    else "range"."""
    k = len(sizes)
    if split == "auto":
        split = "msm" if (world % k == 0 or k % world == 0) else "range"
    work = [[] for _ in range(world)]
    if split == "range" or (world > k and world % k):
        for j, n in enumerate(sizes):
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                if hi > lo:
                    work[r].append((j, lo, hi))
        return work
    if world <= k:
        for j, n in enumerate(sizes):
            if n:
                work[j % world].append((j, 0, n))
        return work
    g = world // k                      # ranks per MSM
    for r in range(world):
        j, part = r // g, r % g
        lo, hi = shard_range(sizes[j], part, g)
        if hi > lo:
            work[r].append((j, lo, hi))
    return work


def gather_set_partials(partials, world, device=None):
    """all_gather of a (k, 4, fq_limbs) array of per-MSM partial sums (identity where the rank
    had no share) -> (world, k, 4, fq_limbs)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return partials[None]
    flat = torch.from_numpy(np.ascontiguousarray(partials).view(np.int64).reshape(-1))
    if device is not None:
        flat = flat.to(device, non_blocking=True)
    out = torch.empty((world * flat.numel(),), dtype=torch.int64, device=flat.device)
    dist.all_gather_into_tensor(out, flat)
    return out.cpu().numpy().view(np.uint64).reshape((world,) + partials.shape)


def combine_set(curve, gathered):
    """(world, k, 4, fq_limbs) -> (k, 4, fq_limbs): per-MSM host sums of the rank partials."""
    world, k = gathered.shape[:2]
    return np.stack([combine_partials(curve, [gathered[r, j] for r in range(world)]) for j in range(k)])

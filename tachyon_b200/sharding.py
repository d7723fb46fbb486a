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


class PartialGather:
    """Reusable buffers for the per-step all-gather of XYZZ partials over NCCL: the partial goes
    through a pinned host word to the device (asynchronous copy), one all_gather_into_tensor on
    the caller's stream, one asynchronous copy back to pinned memory and ONE stream
    synchronisation — instead of a pageable copy each way (two implicit synchronisations and an
    allocation per step).  `shape` = shape of one rank's contribution, e.g. (4, fq_limbs) or
    (k, 4, fq_limbs)."""

    def __init__(self, shape, world, device="cuda"):
        import torch
        self.shape, self.world = tuple(shape), world
        words = int(np.prod(self.shape))
        self.pin_in = torch.empty(words, dtype=torch.int64).pin_memory()
        self.pin_out = torch.empty(world * words, dtype=torch.int64).pin_memory()
        self.dev_in = torch.empty(words, dtype=torch.int64, device=device)
        self.dev_out = torch.empty(world * words, dtype=torch.int64, device=device)
        self._in_np = self.pin_in.numpy().view(np.uint64)
        self._out_np = self.pin_out.numpy().view(np.uint64).reshape((world,) + self.shape)

    def __call__(self, partial, stream=None):
        """(shape) uint64 -> (world,) + shape uint64 (a view of the pinned result buffer, valid
        until the next call)."""
        import torch
        import torch.distributed as dist
        self._in_np[:] = np.ascontiguousarray(partial).reshape(-1)
        stream = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            self.dev_in.copy_(self.pin_in, non_blocking=True)
            dist.all_gather_into_tensor(self.dev_out, self.dev_in)
            self.pin_out.copy_(self.dev_out, non_blocking=True)
        stream.synchronize()
        return self._out_np


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

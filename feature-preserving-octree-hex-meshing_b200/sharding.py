"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

Every sub-path of SURVEY.md §8(e) except the octree closure shards with NO data-path collective:
queries / hexes / samples are split into contiguous ranges, the triangle mesh and its trees are replicated.
The only collectives are the final gather of per-range results and the 5-scalar statistics reduction; both are
fixed-order so an N-rank result is bit-identical to the 1-rank result for integer outputs and for min/max, and
equal up to the documented summation order for fp sums.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`: the first n % world ranks get one extra item."""
    q, r = divmod(int(n), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def z_slabs(nz: int, world: int):
    """z-slab decomposition of a grid with nz layers (VoxelGrid::raw_layer is z-major, voxelization.h:68)."""
    return [shard_range(nz, r, world) for r in range(world)]


def gather_ranges(part: torch.Tensor, n: int, rank: int, world: int) -> torch.Tensor:
    """All-gather variable-length per-rank slices (dim 0) back into the full array, in rank order."""
    if world == 1:
        return part
    sizes = [shard_range(n, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    pad[: part.shape[0]] = part
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def reduce_stats(st: dict) -> dict:
    """min / max / sum / sumsq / count over ranks (scaled Jacobian and Hausdorff statistics)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return st
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mn = torch.tensor([st["min"]], dtype=torch.float64, device=dev); mx = torch.tensor([st["max"]], dtype=torch.float64, device=dev)
    sm = torch.tensor([st["sum"], st["sumsq"], float(st["count"])], dtype=torch.float64, device=dev)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return dict(min=float(mn), max=float(mx), sum=float(sm[0]), sumsq=float(sm[1]), count=int(round(float(sm[2]))))

"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

Every sub-path of SURVEY.md §8(e) except the octree closure shards with NO data-path collective:
queries / hexes / samples are split into contiguous ranges, the triangle mesh and its trees are replicated.
The only collectives are the final gather of per-range results and the 5-scalar statistics reduction; both are
fixed-order so an N-rank result is bit-identical to the 1-rank result for integer outputs and for min/max, and
equal up to the documented summation order for fp sums.

The octree closure is the one path with a real exchange step: `build_octree_sharded` cuts the grid into z slabs,
each rank refines and balances its slab on its GPU, and per level the 2:1 candidates that cross a slab face (the
halo) are all-gathered as Morton codes.  The communicator is a two-method object (`allreduce_max`, `allgather_var`):
`TorchComm` is torch.distributed (NCCL on GPUs, gloo on CPU tensors), `ThreadComm` runs W ranks as threads of one
process for single-GPU tests of the same code path.
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


# ---------------------------------------------------------------------------------------------------------------------
# z-slab sharded octree build
class TorchComm:
    """torch.distributed plumbing for variable-length int64 device buffers."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def allreduce_max(self, v: int, device) -> int:
        t = torch.tensor([int(v)], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def allgather_var(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenation, in rank order, of every rank's 1-D tensor (lengths may differ, may be 0)."""
        n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
        ns = torch.empty(self.world, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(ns, n, group=self.group)
        ns = ns.tolist()
        mx = max(ns)
        if mx == 0:
            return t.new_empty(0)
        pad = t.new_empty(mx)
        pad[: t.numel()] = t
        out = t.new_empty(self.world * mx)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        if all(k == mx for k in ns):
            return out
        return torch.cat([out[r * mx: r * mx + k] for r, k in enumerate(ns)])


class ThreadComm:
    """W ranks as W threads of one process (all on one GPU): the same protocol without a process group."""

    class _Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.slots = [None] * world
            self.barrier = threading.Barrier(world)

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def make(cls, world):
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def _exchange(self, v):
        self.sh.slots[self.rank] = v
        self.sh.barrier.wait()
        got = list(self.sh.slots)
        self.sh.barrier.wait()
        return got

    def allreduce_max(self, v: int, device) -> int:
        return max(self._exchange(int(v)))

    def allgather_var(self, t: torch.Tensor) -> torch.Tensor:
        torch.cuda.current_stream(t.device).synchronize() if t.is_cuda else None
        parts = self._exchange(t)
        out = torch.cat([p.to(t.device) for p in parts]) if any(p.numel() for p in parts) else t.new_empty(0)
        torch.cuda.current_stream(t.device).synchronize() if t.is_cuda else None
        self.sh.barrier.wait()       # nobody reuses its buffer before everyone has copied
        return out


def build_octree_sharded(fp, ctx, mesh, params, comm, device=None, stats: dict | None = None):
    """fpohm_octree_build over `comm.world` z slabs (include/fpohm.h protocol).  Every rank returns the complete,
    canonically numbered octree, bit-identical to `fp.Octree.build` on one GPU.  `fp` is the fpohm_b200 module.
    `stats`, if given, receives halo sizes (codes sent per level) and the slab cut."""
    import time
    device = device if device is not None else torch.device("cuda", ctx.device)
    ph = {}

    def tick(name, t0):
        ctx.sync()
        ph[name] = ph.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t0 = time.perf_counter()
    sh = fp.OctreeShard(ctx, mesh, params, comm.rank, comm.world)
    try:
        lmax = sh.refine()
        t0 = tick("phase1_predicate_own_slab", t0)
        G = comm.allreduce_max(lmax, device)
        if stats is not None:
            stats.update(sh.info()); stats["halo_codes"] = {}; stats["global_max_level"] = G
        for l in range(G, -1, -1):
            n = sh.level_outgoing(G, l)
            buf = torch.empty(n, dtype=torch.int64, device=device)
            if n:
                sh.outgoing_copy(buf.data_ptr())
            t0 = tick("phase2_closure_own_slab", t0)
            got = comm.allgather_var(buf)
            _sync(got)
            t0 = tick("phase2_halo_exchange", t0)
            sh.level_close(l, got.data_ptr() if got.numel() else 0, got.numel())
            if stats is not None:
                stats["halo_codes"][l] = n
        t0 = tick("phase2_closure_own_slab", t0)
        ptrs, counts, keep = [], [], []
        for l in range(G + 1):
            n = sh.level_result(l)
            buf = torch.empty(n, dtype=torch.int64, device=device)
            if n:
                sh.level_result(l, buf.data_ptr())
            got = comm.allgather_var(buf)
            _sync(got)
            keep.append(got)
            ptrs.append(got.data_ptr() if got.numel() else 0)
            counts.append(got.numel())
        t0 = tick("gather_closed_sets", t0)
        out = sh.finish(ptrs, counts)
        t0 = tick("phase3_numbering_replicated", t0)
        if stats is not None:
            stats["phase_ms"] = {k: round(v, 2) for k, v in ph.items()}
        return out
    finally:
        sh.close()


def _sync(t: torch.Tensor):
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()

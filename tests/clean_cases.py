"""Inputs of the clean_hex_mesh parity tests (shared by the golden generator, the CPU and the GPU tests)."""
import numpy as np


def block(nx, ny, nz):
    """Regular hex block, corner order of hex_ref_shape; vertex (i,j,k) at id (i*(ny+1)+j)*(nz+1)+k."""
    idx = np.arange((nx + 1) * (ny + 1) * (nz + 1), dtype=np.int64).reshape(nx + 1, ny + 1, nz + 1)
    g = np.stack(np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    c = lambda dx, dy, dz: idx[dx:nx + dx, dy:ny + dy, dz:nz + dz].reshape(-1)
    H = np.stack([c(0, 0, 0), c(1, 0, 0), c(1, 1, 0), c(0, 1, 0), c(0, 0, 1), c(1, 0, 1), c(1, 1, 1), c(0, 1, 1)], -1)
    return g, np.ascontiguousarray(H.astype(np.uint32))


def carved_block(dims, p, seed):
    """Block with jittered vertices, random inside flags (density p: plenty of non-manifold vertices / edges, peninsulas and
    loose pieces) and a copy of the hex list with ~30 % of the hexes mirrored (negative volume) for reorder_hex_mesh."""
    V, H = block(*dims)
    rng = np.random.default_rng(seed)
    V = V + rng.uniform(-0.2, 0.2, V.shape)
    flag = (rng.random(len(H)) < p).astype(np.uint8)
    Hm = H.copy(); m = rng.random(len(H)) < 0.3
    Hm[m] = Hm[m][:, [3, 2, 1, 0, 7, 6, 5, 4]]
    return V, H, flag, Hm


def lattice_around(tV, n):
    """n cells along the longest axis of the bounding box of tV, padded by one cell (what voxel_meshing produces up to the
    float grid_length, ghm.cpp:215-291)."""
    lo, hi = tV.min(0), tV.max(0)
    h = (hi - lo).max() / n
    dims = np.maximum(np.ceil((hi - lo) / h).astype(int) + 2, 3)
    V, H = block(*dims)
    return V * h + (lo - h), H

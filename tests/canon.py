"""Canonical forms for order-dependent outputs: the reference numbers nodes and cells in
sequential split order, the GPU path numbers them canonically, so parity is defined on id-free forms."""
import numpy as np

DELTA = np.array([[(i & 1) ^ ((i >> 1) & 1), (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], np.int64)  # common.h:147


def _rows_sorted(a):
    a = np.ascontiguousarray(a)
    if len(a) == 0:
        return a
    idx = np.lexsort(a.T[::-1])
    return a[idx]


def canon_octree(ex):
    """ex: dict from {RefOctree,Octree}.export().  Returns id-free sorted arrays."""
    npos = ex["node_pos"].astype(np.int64)
    corner = ex["corner"].astype(np.int64)
    fc = ex["first_child"]
    c0 = npos[corner[:, 0]]
    ext = npos[corner[:, 1]][:, 0] - c0[:, 0]
    cell_key = np.concatenate([c0, ext[:, None]], 1)          # (x, y, z, extent) identifies a cell
    # every corner must sit at c0 + delta*extent
    for k in range(8):
        assert np.array_equal(npos[corner[:, k]], c0 + DELTA[k] * ext[:, None]), f"corner {k} misplaced"
    leaf = fc < 0
    out = {}
    out["cells"] = _rows_sorted(cell_key)
    out["leaves"] = _rows_sorted(cell_key[leaf])
    out["nodes"] = _rows_sorted(npos)
    # children: block of 8 at firstChild, child k at c0 + delta(k)*ext/2 (octree.cpp:549-556)
    ii = np.nonzero(~leaf)[0]
    for k in range(8):
        ch = fc[ii] + k
        assert np.array_equal(cell_key[ch][:, :3], c0[ii] + DELTA[k] * (ext[ii] // 2)[:, None]), f"child {k} misplaced"
        assert np.array_equal(cell_key[ch][:, 3], ext[ii] // 2)
    # cell neighbours as keys
    neigh = ex["neigh"]
    nk = np.full((len(neigh), 6, 4), -1, np.int64)
    m = neigh >= 0
    nk[m] = cell_key[neigh[m]]
    out["cell_neigh"] = _rows_sorted(np.concatenate([cell_key, nk.reshape(len(neigh), 24)], 1))
    # node neighbours as positions
    nn = ex["node_neigh"]
    pk = np.full((len(nn), 6, 3), -1, np.int64)
    m = nn >= 0
    pk[m] = npos[nn[m]]
    out["node_neigh"] = _rows_sorted(np.concatenate([npos, pk.reshape(len(nn), 18)], 1))
    return out


def assert_octree_equal(a, b):
    ca, cb = canon_octree(a), canon_octree(b)
    for k in ca:
        assert ca[k].shape == cb[k].shape, f"{k}: {ca[k].shape} vs {cb[k].shape}"
        assert np.array_equal(ca[k], cb[k]), f"{k} differs"


def canon_hexes(Vpos, hexa):
    """hex mesh -> rows of 24 doubles (8 corner positions in corner order), sorted."""
    P = Vpos[hexa.astype(np.int64)].reshape(len(hexa), 24)
    return _rows_sorted(P)

"""Deterministic procedural input meshes for tests and bench.

All generators return (V float64 [nV,3] C-contiguous, F int32 [nF,3]) of a closed,
consistently outward-oriented triangle mesh.  `normalise_unit_box` mirrors what the
reference does to every input before the hot path sees it (meshing.cpp:103-152, the
AABB flavour): translate the bbox centre to the origin and scale the longest extent to 1.
"""
from __future__ import annotations

import hashlib

import numpy as np


def normalise_unit_box(V: np.ndarray) -> np.ndarray:
    lo, hi = V.min(0), V.max(0)
    c = (lo + hi) / 2
    s = (hi - lo).max()
    return np.ascontiguousarray((V - c) / s)


def mesh_sha256(V: np.ndarray, F: np.ndarray) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(V, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(F, dtype=np.int32).tobytes())
    return h.hexdigest()


def _grid_quads_to_tris(nu: int, nv: int, wrap_u=True, wrap_v=True, flip=False) -> np.ndarray:
    iu = np.arange(nu if wrap_u else nu - 1)
    iv = np.arange(nv if wrap_v else nv - 1)
    U, W = np.meshgrid(iu, iv, indexing="ij")
    a = (U % nu) * nv + (W % nv)
    b = ((U + 1) % nu) * nv + (W % nv)
    c = ((U + 1) % nu) * nv + ((W + 1) % nv)
    d = (U % nu) * nv + ((W + 1) % nv)
    t0 = np.stack([a, b, c], -1).reshape(-1, 3)
    t1 = np.stack([a, c, d], -1).reshape(-1, 3)
    F = np.concatenate([t0, t1], 0)
    if flip:
        F = F[:, ::-1]
    return np.ascontiguousarray(F.astype(np.int32))


def torus(nu: int = 100, nv: int = 100, R: float = 1.0, r: float = 0.35):
    """C1: closed genus-1 torus, 2*nu*nv triangles (100x100 -> 20 000)."""
    u = np.arange(nu) * (2 * np.pi / nu)
    v = np.arange(nv) * (2 * np.pi / nv)
    U, W = np.meshgrid(u, v, indexing="ij")
    x = (R + r * np.cos(W)) * np.cos(U)
    y = (R + r * np.cos(W)) * np.sin(U)
    z = r * np.sin(W)
    V = np.stack([x, y, z], -1).reshape(-1, 3)
    F = _grid_quads_to_tris(nu, nv)
    return normalise_unit_box(V), F


def gear(teeth: int = 24, n_radial: int = 12, n_axial: int = 40, n_arc: int = 8,
         r_root: float = 0.8, r_tip: float = 1.0, r_bore: float = 0.3, height: float = 0.4,
         target_tris: int | None = None):
    """C2: extruded spur gear with a bore: trapezoidal teeth, sharp rims and tooth edges.

    Built as a (profile polygon) x (axial) tensor mesh: outer wall, inner bore wall, top and
    bottom annular caps.  `target_tris` scales the tessellation to roughly that many triangles.
    Returns (V, F, crease_edges[int32 nE,2]) — crease edges are the sharp feature curves.
    """
    if target_tris is not None:
        # tris ~= 2*np_*(2*n_axial + 2*n_radial), np_ = teeth*4*n_arc
        base = 2 * (teeth * 4 * n_arc) * (2 * n_axial + 2 * n_radial)
        s = max(1.0, (target_tris / base) ** 0.5)
        n_arc = max(1, int(round(n_arc * s)))
        n_axial = max(1, int(round(n_axial * s)))
        n_radial = max(1, int(round(n_radial * s)))
    # outer profile: per tooth 4 segments (root arc, rising flank, tip arc, falling flank)
    prof = []
    corner_idx = []
    dth = 2 * np.pi / teeth
    for t in range(teeth):
        th0 = t * dth
        knots = [(th0 + 0.00 * dth, r_root), (th0 + 0.30 * dth, r_root), (th0 + 0.40 * dth, r_tip),
                 (th0 + 0.80 * dth, r_tip), (th0 + 0.90 * dth, r_root), (th0 + 1.0 * dth, r_root)]
        # merge last root arc into first (5 knots define 4 segs + closing root seg); use 5 segs with n_arc each
        for k in range(5):
            (ta, ra), (tb, rb) = knots[k], knots[k + 1]
            if k in (1, 2, 3, 4):
                corner_idx.append(len(prof))
            for i in range(n_arc):
                s_ = i / n_arc
                th = ta + (tb - ta) * s_
                xa, ya = ra * np.cos(ta), ra * np.sin(ta)
                xb, yb = rb * np.cos(tb), rb * np.sin(tb)
                if ra == rb:
                    prof.append((ra * np.cos(th), ra * np.sin(th)))
                else:
                    prof.append((xa + (xb - xa) * s_, ya + (yb - ya) * s_))
    prof = np.asarray(prof, dtype=np.float64)
    npf = len(prof)
    ang = np.arctan2(prof[:, 1], prof[:, 0])
    bore = np.stack([r_bore * np.cos(ang), r_bore * np.sin(ang)], -1)

    verts = []
    index = {}

    def vid(kind, i, j):
        key = (kind, i % npf, j)
        if key not in index:
            index[key] = len(verts)
            verts.append(None)
        return index[key]

    # canonical vertex positions: rings r in [0..n_radial] (0 = bore, n_radial = outer), levels z in [0..n_axial]
    def pos(i, ring, lev):
        p = bore[i] + (prof[i] - bore[i]) * (ring / n_radial)
        return (p[0], p[1], -height / 2 + height * lev / n_axial)

    def v(i, ring, lev):
        key = (i % npf, ring, lev)
        if key not in index:
            index[key] = len(verts)
            verts.append(pos(i % npf, ring, lev))
        return index[key]

    index.clear()
    tris = []

    def quad(a, b, c, d):
        tris.append((a, b, c))
        tris.append((a, c, d))

    for i in range(npf):
        for l in range(n_axial):
            # outer wall, outward normal
            quad(v(i, n_radial, l), v(i + 1, n_radial, l), v(i + 1, n_radial, l + 1), v(i, n_radial, l + 1))
            # bore wall, normal towards the axis
            quad(v(i + 1, 0, l), v(i, 0, l), v(i, 0, l + 1), v(i + 1, 0, l + 1))
        for r in range(n_radial):
            # top cap (+z)
            quad(v(i, r, n_axial), v(i, r + 1, n_axial), v(i + 1, r + 1, n_axial), v(i + 1, r, n_axial))
            # bottom cap (-z)
            quad(v(i, r + 1, 0), v(i, r, 0), v(i + 1, r, 0), v(i + 1, r + 1, 0))
    V = np.asarray(verts, dtype=np.float64)
    F = np.asarray(tris, dtype=np.int32)
    crease = []
    for i in range(npf):
        for lev in (0, n_axial):
            crease.append((v(i, n_radial, lev), v(i + 1, n_radial, lev)))
            crease.append((v(i, 0, lev), v(i + 1, 0, lev)))
    for i in corner_idx:
        for l in range(n_axial):
            crease.append((v(i, n_radial, l), v(i, n_radial, l + 1)))
    return normalise_unit_box(V), np.ascontiguousarray(F), np.asarray(crease, dtype=np.int32)


def linked_tori(n: int = 4, nu: int = 64, nv: int = 32):
    """C3 base: n^3 lattice of disjoint closed tori (high genus = n^3), alternating orientation."""
    Vs, Fs = [], []
    off = 0
    k = 0
    for a in range(n):
        for b in range(n):
            for c in range(n):
                V, F = torus(nu, nv, 1.0, 0.3)
                axis = k % 3
                V = np.roll(V, axis, axis=1)
                if axis == 1 or axis == 2:
                    pass
                V = V * 0.8 + np.array([a, b, c], dtype=np.float64)
                # np.roll of coordinates is an even permutation (cyclic) -> orientation preserved
                Vs.append(V)
                Fs.append(F + off)
                off += len(V)
                k += 1
    return normalise_unit_box(np.concatenate(Vs)), np.ascontiguousarray(np.concatenate(Fs).astype(np.int32))


def midpoint_subdivide(V: np.ndarray, F: np.ndarray, levels: int = 1):
    """1->4 midpoint subdivision (keeps the surface, multiplies triangles by 4 per level)."""
    for _ in range(levels):
        E = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], 0).astype(np.int64)
        Es = np.sort(E, 1)
        key = Es[:, 0] * (len(V) + 1) + Es[:, 1]
        uk, inv = np.unique(key, return_inverse=True)
        a = (uk // (len(V) + 1)).astype(np.int64)
        b = (uk % (len(V) + 1)).astype(np.int64)
        mid = (V[a] + V[b]) * 0.5
        m = inv.reshape(3, -1).T + len(V)  # [nF,3]: mids of edges 01,12,20
        V = np.concatenate([V, mid], 0)
        f0, f1, f2 = F[:, 0], F[:, 1], F[:, 2]
        m01, m12, m20 = m[:, 0], m[:, 1], m[:, 2]
        F = np.concatenate([
            np.stack([f0, m01, m20], -1), np.stack([m01, f1, m12], -1),
            np.stack([m20, m12, f2], -1), np.stack([m01, m12, m20], -1)], 0).astype(np.int32)
    return np.ascontiguousarray(V), np.ascontiguousarray(F)


def warped_hex_block(n: int = 16, amp: float = 0.3, seed: int = 42):
    """C4: structured n^3 hex block (corner order = hex_ref_shape, global_types.h:216-226) with a smooth
    seeded displacement of amplitude amp*h.  Returns (V [nV,3] float64, H [nH,8] uint32)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    g = np.arange(n + 1, dtype=np.float64) / n
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    ph = rng.uniform(0, 2 * np.pi, size=(3, 3))
    h = 1.0 / n
    D = [amp * h * np.sin(2 * np.pi * (1 + k) * X + ph[k, 0]) * np.cos(2 * np.pi * (2 + k) * Y + ph[k, 1]) *
         np.sin(2 * np.pi * (1 + k) * Z + ph[k, 2]) for k in range(3)]
    V = np.stack([X + D[0], Y + D[1], Z + D[2]], -1).reshape(-1, 3)
    idx = np.arange((n + 1) ** 3, dtype=np.int64).reshape(n + 1, n + 1, n + 1)
    c = lambda dx, dy, dz: idx[dx:n + dx, dy:n + dy, dz:n + dz].reshape(-1)
    H = np.stack([c(0, 0, 0), c(1, 0, 0), c(1, 1, 0), c(0, 1, 0), c(0, 0, 1), c(1, 0, 1), c(1, 1, 1), c(0, 1, 1)], -1)
    return np.ascontiguousarray(V), np.ascontiguousarray(H.astype(np.uint32))


def hex_lattice_around(tV, n: int):
    """Regular hex lattice (corner order of hex_ref_shape, vertex (i,j,k) at id (i*(ny+1)+j)*(nz+1)+k) with n cells along the
    longest axis of the bounding box of tV, padded by one cell — the shape voxel_meshing hands to clean_hex_mesh."""
    tV = np.asarray(tV, np.float64)
    lo, hi = tV.min(0), tV.max(0)
    h = (hi - lo).max() / n
    nx, ny, nz = (int(d) for d in np.maximum(np.ceil((hi - lo) / h).astype(int) + 2, 3))
    idx = np.arange((nx + 1) * (ny + 1) * (nz + 1), dtype=np.int64).reshape(nx + 1, ny + 1, nz + 1)
    g = np.stack(np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    c = lambda dx, dy, dz: idx[dx:nx + dx, dy:ny + dy, dz:nz + dz].reshape(-1)
    H = np.stack([c(0, 0, 0), c(1, 0, 0), c(1, 1, 0), c(0, 1, 0), c(0, 0, 1), c(1, 0, 1), c(1, 1, 1), c(0, 1, 1)], -1)
    return np.ascontiguousarray(g * h + (lo - h)), np.ascontiguousarray(H.astype(np.uint32))


def c3_mesh():
    """BASELINE config C3: 4^3 lattice of tori, one midpoint subdivision -> 2 027 520 triangles (high genus, closed)."""
    return midpoint_subdivide(*linked_tori(4, 90, 44), 1)


def c4_queries(V, F, n_jitter: int = 1_500_000, n_block: int = 216, h: float = 1.0 / 1024, seed: int = 7):
    """BASELINE config C4 query sets against the surface (V, F):
      project  — `n_jitter` points jittered +-2h around the surface (one barycentric sample of facet k*nF/n_jitter each, in
                 facet order: what projection_smooth / dirty_graph_projection see, ghm.cpp:3760-3781) followed by the
                 boundary vertices of the n_block^3 hex block laid over the bounding box (far-field queries);
      classify — the n_block^3 hex centres of that block in the lattice's own order (id = i*ny*nz + j*nz + k, z fastest:
                 voxel_meshing ghm.cpp:242 -> clean_hex_mesh's points_inside_mesh, ghm.cpp:1937-1951).
    Returns (project [n,3], classify [n_block^3,3]) float64, C-contiguous."""
    rng = np.random.Generator(np.random.PCG64(seed))
    V = np.asarray(V, np.float64); F = np.asarray(F)
    fid = (np.arange(n_jitter, dtype=np.int64) * len(F)) // n_jitter
    b = rng.random((n_jitter, 2))
    flip = b.sum(1) > 1.0
    b[flip] = 1.0 - b[flip]
    tri = V[F[fid].astype(np.int64)]
    pts = tri[:, 0] + b[:, :1] * (tri[:, 1] - tri[:, 0]) + b[:, 1:] * (tri[:, 2] - tri[:, 0])
    pts += (rng.random(pts.shape) - 0.5) * (4.0 * h)
    lo, hi = V.min(0), V.max(0)
    g = np.arange(n_block + 1, dtype=np.float64) / n_block
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    on_b = np.zeros(X.shape, bool)
    on_b[[0, -1], :, :] = True; on_b[:, [0, -1], :] = True; on_b[:, :, [0, -1]] = True
    bverts = np.stack([X[on_b], Y[on_b], Z[on_b]], -1) * (hi - lo) + lo
    gc = (np.arange(n_block, dtype=np.float64) + 0.5) / n_block
    Xc, Yc, Zc = np.meshgrid(gc, gc, gc, indexing="ij")
    centres = np.stack([Xc, Yc, Zc], -1).reshape(-1, 3) * (hi - lo) + lo
    return np.ascontiguousarray(np.concatenate([pts, bverts])), np.ascontiguousarray(centres)

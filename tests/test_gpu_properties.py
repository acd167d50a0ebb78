"""GPU tests at BASELINE.json's full sizes through size-independent properties, plus the edge cases of the path.
(The bit-level parity against the reference at sizes the CPU finishes in seconds lives in test_gpu_parity.py.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gear(fp):
    V, F, crease = fp.procedural.gear()          # config[1]: ~200 k triangles with sharp feature curves
    return V, F, crease


def test_full_size_octree_depth8_properties(fp, ctx, gear):
    V, F, _ = gear
    m = fp.TriMesh(ctx, V, F)
    p = fp.octree_grid_setup(V, 1 << 20)
    p.c.stop_extent = 1 << 12                    # --e 12: depth 8
    o = fp.Octree.build(ctx, m, p)
    s = o.sizes()
    assert s["leaves"] > 1_000_000 and s["cells"] == s["roots"] + 8 * (s["cells"] - s["leaves"])      # every internal cell has 8 children
    assert o.flags() == (True, True)             # is2to1Graded, isPaired evaluated on the device
    ex = o.export()
    npos, corner, fc, neigh, nn = ex["node_pos"].astype(np.int64), ex["corner"], ex["first_child"], ex["neigh"], ex["node_neigh"]
    c0 = npos[corner[:, 0]]; ext = npos[corner[:, 1]][:, 0] - c0[:, 0]
    leaf = fc < 0
    assert ext[leaf].min() == 1 << 12 and (ext & (ext - 1) == 0).all()
    # leaves tile the grid exactly: volumes add up
    assert int((ext[leaf].astype(object) ** 3).sum()) == int(np.prod(p.grid_size.astype(object)))
    # nodes are unique and sorted (canonical Morton order => strictly increasing keys is checked on device; here: unique positions)
    assert len(np.unique(npos, axis=0)) == len(npos)
    # neighbour symmetry for same-size neighbours (assertIsValid, octree.cpp:874-895)
    for ax in range(3):
        nx = neigh[:, 2 * ax + 1]
        ok = nx >= 0
        same = ok & (ext[np.where(ok, nx, 0)] == ext)
        assert np.array_equal(neigh[nx[same], 2 * ax], np.nonzero(same)[0])
        # node links are symmetric too (octree.cpp:857-871)
        nxt = nn[:, 2 * ax + 1]
        okn = nxt >= 0
        assert np.array_equal(nn[nxt[okn], 2 * ax], np.nonzero(okn)[0])
    # idempotence: subdividing again with the same stop extent changes nothing
    o.subdivide(m, 1 << 12)
    assert o.sizes() == s
    # hexes of an axis-aligned octree are cubes: scaled Jacobian exactly 1 at every corner
    Vh, H, h2c = o.hexes()
    VJ, HJ, mad, fl = fp.scaled_jacobian(ctx, Vh, H)
    assert fl == 0 and np.all(np.abs(VJ - 1.0) < 1e-12) and abs(mad[0] - 1) < 1e-12 and mad[2] < 1e-20
    assert np.array_equal(h2c, np.nonzero(leaf)[0])
    o.close(); m.close()


def test_full_size_closest_point_properties(fp, ctx, gear):
    V, F, _ = gear
    m = fp.TriMesh(ctx, V, F)
    rng = np.random.default_rng(4)
    P = rng.uniform(-0.55, 0.55, (2_000_000, 3)); P[:, 2] *= 0.3
    S, I, Cp, N = m.signed_distance_pseudonormal(P)
    assert (I >= 0).all() and (I < len(F)).all()
    # |P - C| is the reported distance; C lies in the plane of facet I and inside its bounding box
    np.testing.assert_allclose(np.linalg.norm(P - Cp, axis=1), np.abs(S), rtol=1e-12, atol=1e-15)
    T = V[F[I]]
    n = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0]); n /= np.linalg.norm(n, axis=1)[:, None]
    assert np.abs(np.einsum("ij,ij->i", Cp - T[:, 0], n)).max() < 1e-9
    assert ((Cp >= T.min(1) - 1e-12) & (Cp <= T.max(1) + 1e-12)).all()
    # no vertex of the mesh is closer than the reported distance (vertices are on the surface)
    sub = rng.integers(0, len(P), 2000)
    dv = np.linalg.norm(P[sub, None, :] - V[None, ::37, :], axis=2).min(1)
    assert (dv >= np.abs(S[sub]) - 1e-12).all()
    # sign agrees with the independent z-ray parity classification away from the surface
    mn, ext = V.min(0), V.max(0) - V.min(0)
    g = fp.VoxelGrid(mn, ext, 1 / 200, 1)
    vox = fp.compute_sign_voxels(ctx, m, g)
    z, y, x = np.meshgrid(np.arange(g.dims[2]), np.arange(g.dims[1]), np.arange(g.dims[0]), indexing="ij")
    ctr = np.stack([(x + 0.5) * g.spacing + g.origin[0], (y + 0.5) * g.spacing + g.origin[1], (z + 0.5) * g.spacing + g.origin[2]], -1).reshape(-1, 3)
    Sv = m.signed_distance_pseudonormal(ctr, want=("S",))[0]
    far = np.abs(Sv) > 1e-9
    assert np.array_equal((Sv < 0)[far], vox.reshape(-1)[far] == 1)
    assert 0 < vox.sum() < vox.size
    m.close()


def test_full_size_connectivity_and_jacobian_properties(fp, ctx):
    V, H = fp.procedural.warped_hex_block(96)                     # 884 736 hexes
    c = fp.HexConnectivity(ctx, H, len(V))
    nH, n = len(H), 96
    assert len(c.F_vs) == 3 * n * n * (n + 1) and len(c.E_vs) == 3 * n * (n + 1) ** 2
    assert int(c.F_boundary.sum()) == 6 * n * n
    off, val = c.F_nhs
    assert off[-1] == 6 * nH and np.all(np.diff(off) == 2 - c.F_boundary)
    assert np.array_equal(np.sort(c.H_fs.reshape(-1)), np.repeat(np.arange(len(c.F_vs)), np.diff(off)))
    # faces are numbered in lexicographic order of their sorted vertex tuples (gf.cpp:132-147)
    sf = np.sort(c.F_vs.astype(np.int64), 1)
    key = ((sf[:, 0] * (len(V) + 1) + sf[:, 1]) * (len(V) + 1) + sf[:, 2])
    assert np.all(np.diff(key.astype(object)) >= 0)
    VJ, HJ, mad, fl = fp.scaled_jacobian(ctx, V, H)
    assert np.array_equal(HJ, np.minimum(1.0, VJ.reshape(-1, 8).min(1))) and mad[0] == HJ.min()
    np.testing.assert_allclose([mad[1], mad[2]], [HJ.mean(), HJ.var()], rtol=1e-10)
    assert fl == int((HJ < 0).sum())


def test_hausdorff_properties(fp, ctx, gear):
    V, F, _ = gear
    A = fp.TriMesh(ctx, V, F)
    h = fp.hausdorff(ctx, A, A)
    assert h["max"] == 0.0 and h["mean"] == 0.0                   # identical meshes
    shift = np.array([0.0, 0.0, 0.013])
    B = fp.TriMesh(ctx, V + shift, F)
    h = fp.hausdorff(ctx, A, B)
    assert 0 < h["max"] <= 0.013 + 1e-12 and h["mean"] <= h["max"] and h["n_ab"] == len(V)
    A.close(); B.close()


# ---- edge cases --------------------------------------------------------------------------------------------------------
def test_edge_cases_queries(fp, ctx, ref):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0], [2, 2, 2], [2, 2, 2.0]])
    # single facet (tree root is a leaf), empty query batch, a degenerate (zero-area) facet next to a regular one
    for F in ([[0, 1, 2]], [[0, 1, 2], [0, 2, 3], [4, 5, 4]]):
        F = np.array(F, np.int32)
        m = fp.TriMesh(ctx, V, F)
        S, I, Cp, N = m.signed_distance_pseudonormal(np.zeros((0, 3)))
        assert len(S) == 0
        P = np.array([[0.2, 0.2, 0.5], [0.2, 0.2, -0.5], [5, 5, 5], [0, 0, 0], [0.5, 0.5, 0.0], [-1, -1, 0.0]])
        S, I, Cp, N = m.signed_distance_pseudonormal(P)
        rS, rI, rC, rN = ref.RefTree(V, F).signed_distance(P)
        assert np.array_equal(S, rS, equal_nan=True) and np.array_equal(I, rI) and np.array_equal(Cp, rC) and np.array_equal(N, rN, equal_nan=True)
        m.close()


def test_edge_cases_abi_validation(fp, ctx):
    lib = fp.lib()
    V = np.zeros((3, 3)); F = np.array([[0, 1, 7]], np.int32)       # facet index out of range
    h = C.c_void_p()
    assert lib.fpohm_mesh_upload(ctx.h, V.ctypes.data_as(C.c_void_p), C.c_int64(3), F.ctypes.data_as(C.c_void_p), C.c_int64(1), C.byref(h)) == -1
    assert b"out of range" in lib.fpohm_last_error()
    assert lib.fpohm_mesh_upload(ctx.h, V.ctypes.data_as(C.c_void_p), C.c_int64(0), F.ctypes.data_as(C.c_void_p), C.c_int64(0), C.byref(h)) == -1
    with pytest.raises(fp.FpohmError):                                # grid size must be a power of two (octree.cpp:26-28)
        fp.Octree.from_marks(ctx, [12, 8, 8], np.zeros((0, 4), np.int32))
    with pytest.raises(fp.FpohmError):                                # hex corner id beyond nV
        fp.scaled_jacobian(ctx, np.zeros((4, 3)), np.array([[0, 1, 2, 3, 4, 5, 6, 9]], np.uint32))
    Vt, Ft = fp.procedural.torus(12, 8)
    m = fp.TriMesh(ctx, Vt, Ft)
    g = fp.VoxelGrid(Vt.min(0), Vt.max(0) - Vt.min(0), 1 / 40, 0)
    import torch
    buf = torch.empty(g.num_voxels(), dtype=torch.uint8, device="cuda")
    with pytest.raises(fp.FpohmError):                                # slab not aligned to 32 layers
        fp.voxel_sign_slab_dev(ctx, m, g, 3, int(g.dims[2]), buf.data_ptr(), 0)
    m.close()


def test_edge_cases_octree(fp, ctx, ref):
    # no split at all (stop extent >= root extent): only root cells, in Layout3D order, fully linked
    V, F = fp.procedural.torus(20, 12)
    m = fp.TriMesh(ctx, V, F)
    p = fp.octree_grid_setup(V, 1 << 20)
    p.c.stop_extent = int(p.grid_size.max())
    o = fp.Octree.build(ctx, m, p)
    gs, org, mt, vs = ref.octree_grid_setup(V, F, 1 << 20)
    r = ref.RefOctree.build(V, F, gs, org, mt, vs, int(gs.max()))
    from canon import assert_octree_equal
    assert_octree_equal(r.export(), o.export())
    assert o.sizes()["cells"] == o.sizes()["roots"] == o.sizes()["leaves"]
    # root ids follow Layout3D::toIndex (x fastest) exactly like the reference (createRootCells, octree.cpp:92-101)
    assert np.array_equal(o.export()["node_pos"][o.export()["corner"][:, 0]], r.export()["node_pos"][r.export()["corner"][:, 0]])
    o.close()
    # extent-1 limit: "Cannot subdivide cell of length 1" (octree.cpp:670-671) — marks at extent 1 are ignored
    marks = np.array([[0, 0, 0, 8], [0, 0, 0, 4], [0, 0, 0, 2], [0, 0, 0, 1]], np.int32)
    o = fp.Octree.from_marks(ctx, [8, 8, 8], marks)
    r = ref.RefOctree.from_marks([8, 8, 8], marks[:3])
    assert_octree_equal(r.export(), o.export())
    o.close(); m.close()


def _stacked_sheets(n_sheets):
    Vs, Fs = [], []
    for k in range(n_sheets):
        z = 0.6 * (k + 0.37) / n_sheets
        b = len(Vs)
        Vs += [[-1, -1, z], [1, -1, z], [1, 1, z], [-1, 1, z]]
        # alternate orientation: entry / exit sheets, so the parity rule has something to count
        Fs += [[b, b + 1, b + 2], [b, b + 2, b + 3]] if k % 2 == 1 else [[b, b + 2, b + 1], [b, b + 3, b + 2]]
    return np.array(Vs, float), np.array(Fs, np.int32)


@pytest.mark.parametrize("n_sheets", [40, 150, 600])
def test_ray_hit_lists_grow_with_the_input(fp, ctx, ref, n_sheets):
    """More than 32 ray/facet hits in one column: the reference collects them in a std::vector (voxelization.h:248-256), round 1
    refused with FPOHM_ERANGE.  The pass is now repeated with a larger list (128 / 512 / 2048) and must equal the reference for
    all three grid flavours — stacked alternating sheets, 40 .. 600 hits per column."""
    V, F = _stacked_sheets(n_sheets)
    m = fp.TriMesh(ctx, V, F)
    org, ext, sp = np.array([-0.5, -0.5, -0.1]), np.array([1.0, 1.0, 0.8]), 0.05
    g = fp.VoxelGrid(org, ext, sp, 0)
    vox = fp.compute_sign_voxels(ctx, m, g)
    rv, _ = ref.voxel_sign(V, F, org, ext, sp, 0)
    assert np.array_equal(vox.reshape(-1), rv.reshape(-1))
    off, val = fp.compute_sign_dexels(ctx, m, g)
    roff, rval, _ = ref.dexel_sign(V, F, org, ext, sp, 0)
    assert np.array_equal(off, roff) and np.array_equal(val, rval)
    assert len(val) > 0 and vox.any()
    m.close()


def test_ray_hit_lists_have_a_documented_limit(fp, ctx):
    V, F = _stacked_sheets(2100)
    m = fp.TriMesh(ctx, V, F)
    g = fp.VoxelGrid([-0.5, -0.5, -0.1], [1, 1, 0.8], 0.25, 0)
    with pytest.raises(fp.FpohmError) as e:
        fp.compute_sign_voxels(ctx, m, g)
    assert e.value.code == -5      # FPOHM_ERANGE: more than 2048 hits in one column
    m.close()


def test_full_size_clean_hex_mesh_properties(fp, ctx, gear):
    """clean_hex_mesh at a size the reference needs minutes for (1.5 M-hex lattice around the 200 k-facet gear), through
    properties that do not need it: every stage is idempotent on its own output, the result is one face-connected piece whose
    boundary is a closed, consistently oriented 2-manifold (every surface edge has exactly two faces running opposite ways),
    the maps are inverse to each other and the medial flags are exactly the boundary of the kept set."""
    V, F, _ = gear
    m = fp.TriMesh(ctx, V, F)
    Vl, Hl = fp.procedural.hex_lattice_around(V, 192)
    nV = len(Vl)
    assert len(Hl) > 1_500_000
    conn = fp.HexConnectivity(ctx, Hl, nV, keep=True)
    out = fp.clean_hex_mesh(ctx, m, Vl, Hl, conn)
    flag = out["H_flag"]
    assert out["stats"][4] == int(flag.sum()) > 500_000
    assert np.array_equal(flag <= (out["signed_dis"] < 0), np.ones(len(flag), bool)) or out["stats"][1] >= 1      # only tagging may add hexes
    # idempotence of every stage on the final flags
    t, _ = fp.tag_uneven_elements(ctx, conn, flag)
    n, rounds = fp.clean_non_manifold(ctx, out["hex"], nV, flag)
    d, pieces = fp.drop_small_pieces(ctx, out["hex"], nV, flag)
    assert np.array_equal(n, flag) and rounds == 0 and np.array_equal(d, flag) and pieces == 1
    assert np.array_equal(t, flag) or int((t != flag).sum()) < 10          # tagging ran BEFORE the other stages in the pipeline
    # maps
    s = fp.reindex_submesh(ctx, out["hex"], nV, flag)
    vr, vm, hr = s["V_map_reverse"], s["V_map"], s["H_map_reverse"]
    assert np.array_equal(vm[vr], np.arange(len(vr))) and (np.diff(vr) > 0).all() and np.array_equal(hr, np.nonzero(flag)[0])
    assert np.array_equal(vr[s["hex"].astype(np.int64)], out["hex"][hr].astype(np.int64))
    # medial flags == faces between kept and dropped hexes
    off, val = conn.F_nhs
    two = np.diff(off) == 2
    a = flag[val[off[:-1]]]; b = np.where(two, flag[val[np.minimum(off[:-1] + 1, len(val) - 1)]], 0)
    assert np.array_equal(out["F_medial"], (a != b).astype(np.uint8))
    conn.close()
    # the boundary surface of the kept hexes
    Vs = Vl[vr]
    sc = fp.HexConnectivity(ctx, s["hex"], len(Vs), keep=True)
    q = fp.extract_surface(ctx, sc, Vs, False)
    sc.close(); m.close()
    assert int(out["F_medial"].sum()) == len(q["F_vs"])
    eoff, _ = q["E_nfs"]
    assert (np.diff(eoff) == 2).all() and not q["E_boundary"].any()       # closed 2-manifold
    Fv = q["F_vs"].astype(np.int64)
    d0 = Fv.reshape(-1); d1 = np.roll(Fv, -1, 1).reshape(-1)
    fwd = np.bincount(q["F_es"].reshape(-1)[d0 < d1], minlength=len(q["E_vs"]))
    assert (fwd == 1).all()                                                   # each edge once in each direction: consistent orientation
    # outward orientation: signed volume of the quad surface is positive
    P = q["V"]; c = P[Fv].mean(1)
    vol = sum(np.einsum("ij,ij->i", np.cross(P[Fv[:, k]], P[Fv[:, (k + 1) % 4]]), c).sum() for k in range(4)) / 6
    kept_vol = float(flag.sum()) * float(np.prod(Vl[Hl[0, 6]] - Vl[Hl[0, 0]]))
    assert abs(vol - kept_vol) < 1e-9 * kept_vol          # orient_surface_mesh leaves res <= 0, i.e. a positive enclosed volume


def test_mesh_cache_is_keyed_on_content(fp, ctx, ref):
    """points_inside_mesh rebuilds its tree on every call in the reference; fpohm_mesh_upload_cached returns the SAME
    surface (trees built) for the same (V, F) content, a different one when a single coordinate changes, and the answers do not depend
    on which path served them."""
    import time
    V, F = fp.procedural.torus(60, 40)
    a = fp.TriMesh(ctx, V, F, cached=True)
    b = fp.TriMesh(ctx, V.copy(), F.copy(), cached=True)
    assert a.h.value == b.h.value
    V2 = V.copy(); V2[7, 1] += 1e-9
    c = fp.TriMesh(ctx, V2, F, cached=True)
    assert c.h.value != a.h.value
    P = np.random.default_rng(1).uniform(-0.6, 0.6, (5000, 3))
    S0 = fp.TriMesh(ctx, V, F).signed_distance_pseudonormal(P, want=("S",))[0]
    t = time.perf_counter(); S1 = fp.points_inside_mesh(ctx, P, V, F); t1 = time.perf_counter() - t
    t = time.perf_counter(); S2 = fp.points_inside_mesh(ctx, P, V, F); t2 = time.perf_counter() - t
    assert np.array_equal(S0, S1) and np.array_equal(S1, S2)
    assert np.array_equal(S1 < 0, ref.points_inside_mesh(V, F, P) < 0)
    for m in (a, b, c):
        m.close()
    assert fp.lib().fpohm_ctx_mesh_cache_clear(ctx.h) == 0

"""GPU parity tests: CUDA path (through the C-ABI) vs the reference's own code (oracle/_ref)."""
from pathlib import Path

import numpy as np
import pytest

from canon import assert_octree_equal, canon_hexes

pytestmark = pytest.mark.gpu


def _meshes(fp):
    pm = fp.procedural
    return {"torus": pm.torus(60, 40), "gear": pm.gear(teeth=12, n_radial=4, n_axial=6, n_arc=2)[:2],
            "tori": pm.linked_tori(2, 16, 8)}


def test_scaled_jacobian_bit_exact(fp, ctx, ref):
    pm = fp.procedural
    for n, amp in [(8, 0.3), (12, 1.5)]:  # amp 1.5 tangles the block: flipped hexes, negative Jacobians
        V, H = pm.warped_hex_block(n, amp)
        VJ, HJ, mad, fl = fp.scaled_jacobian(ctx, V, H)
        rVJ, rHJ, rmad, rfl = ref.scaled_jacobian(V, H)
        assert np.array_equal(VJ, rVJ) and np.array_equal(HJ, rHJ)   # same IEEE operations
        assert fl == rfl
        assert mad[0] == rmad[0]
        np.testing.assert_allclose(mad[1:], rmad[1:], rtol=1e-12)     # tolerance of the north star is 1e-5


def test_scaled_jacobian_degenerate(fp, ctx, ref):
    V, H = fp.procedural.warped_hex_block(3, 0.0)
    V = V.copy(); V[int(H[0, 1])] = V[int(H[0, 0])]  # zero-length edge -> "Potential Bug" branch (gf.cpp:2436)
    VJ, HJ, mad, fl = fp.scaled_jacobian(ctx, V, H)
    rVJ, rHJ, rmad, rfl = ref.scaled_jacobian(V, H)
    assert np.array_equal(VJ, rVJ) and np.array_equal(HJ, rHJ) and fl == rfl


def test_signed_distance_parity(fp, ctx, ref):
    rng = np.random.default_rng(7)
    for name, (V, F) in _meshes(fp).items():
        m = fp.TriMesh(ctx, V, F)
        rt = ref.RefTree(V, F)
        P = np.concatenate([rng.uniform(-0.6, 0.6, (20000, 3)),
                            V[rng.integers(0, len(V), 2000)],                                   # exactly on vertices
                            V[F[rng.integers(0, len(F), 2000)]].mean(1),                        # on faces
                            V[F[rng.integers(0, len(F), 2000)]][:, :2].mean(1)])                # on edges
        S, I, C, N = m.signed_distance_pseudonormal(P)
        rS, rI, rC, rN = rt.signed_distance(P)
        # the query tree is igl's own (same std:: calls on the host) and the search reproduces igl's tie-break: every bit
        assert np.array_equal(I, rI), (name, int((I != rI).sum()))
        assert np.array_equal(S, rS) and np.array_equal(C, rC) and np.array_equal(N, rN), name
        m.close()


def test_points_inside_mesh_occupancy(fp, ctx, ref):
    V, F = fp.procedural.torus(60, 40)
    g = (np.arange(24) + 0.5) / 24 - 0.5
    P = np.stack(np.meshgrid(g, g, g * 0.4, indexing="ij"), -1).reshape(-1, 3)
    S = fp.points_inside_mesh(ctx, P, V, F)
    rS = ref.points_inside_mesh(V, F, P)
    assert np.array_equal(S < 0, rS < 0)       # cell occupancy bit-exact (ghm.cpp:1952)
    np.testing.assert_allclose(S, rS, rtol=1e-12)


def test_point_mesh_squared_distance(fp, ctx, ref):
    V, F = fp.procedural.torus(40, 30)
    P = np.random.default_rng(3).uniform(-0.6, 0.6, (5000, 3))
    m = fp.TriMesh(ctx, V, F)
    D, I, C = m.point_mesh_squared_distance(P)
    rD, rI, rC = ref.point_mesh_sqdist(V, F, P)
    assert np.array_equal(I, rI) and np.array_equal(D, rD) and np.array_equal(C, rC)


@pytest.mark.parametrize("graded,paired", [(True, True), (True, False), (False, True), (False, False)])
def test_octree_from_marks_vs_reference(fp, ctx, ref, graded, paired):
    rng = np.random.default_rng(11)
    for gs in ([16, 16, 16], [32, 16, 8], [8, 8, 32]):
        gs = np.array(gs, np.int32)
        marks = []
        for e in (2, 4, 8, 16, 32):
            if e > gs.min():
                continue
            n = gs // e
            k = max(1, int(np.prod(n) * (0.15 if e > 2 else 0.03)))
            xyz = np.stack([rng.integers(0, n[d], k) for d in range(3)], -1) * e
            marks.append(np.concatenate([xyz, np.full((k, 1), e)], 1))
        marks = np.concatenate(marks).astype(np.int32)
        r = ref.RefOctree.from_marks(gs, marks, graded, paired)
        o = fp.Octree.from_marks(ctx, gs, marks, graded, paired)
        assert_octree_equal(r.export(), o.export())
        assert o.flags() == r.flags()
        o.close()


def test_octree_build_vs_reference(fp, ctx, ref):
    for name, (V, F) in _meshes(fp).items():
        p = fp.octree_grid_setup(V, 1 << 20)
        gs, org, mt, vs = ref.octree_grid_setup(V, F, 1 << 20)
        assert np.array_equal(p.grid_size, gs) and np.array_equal(p.origin, org) and np.array_equal(p.mesh_transform, mt) and p.voxel_size == vs
        m = fp.TriMesh(ctx, V, F)
        for E in (16, 15, 14):
            p.c.stop_extent = 1 << E
            r = ref.RefOctree.build(V, F, gs, org, mt, vs, 1 << E)
            o = fp.Octree.build(ctx, m, p)
            assert_octree_equal(r.export(), o.export())
            rV, rH, _ = r.hexes()
            oV, oH, _ = o.hexes()
            assert np.array_equal(canon_hexes(rV, rH), canon_hexes(oV, oH)), name   # vertex positions bit-exact
            assert o.flags() == (True, True)
            o.close()
        m.close()


def test_octree_subdivide_and_refine_vs_reference(fp, ctx, ref):
    V, F = fp.procedural.torus(60, 40)
    p = fp.octree_grid_setup(V, 1 << 20)
    gs, org, mt, vs = ref.octree_grid_setup(V, F, 1 << 20)
    m = fp.TriMesh(ctx, V, F)
    p.c.stop_extent = 1 << 16
    r = ref.RefOctree.build(V, F, gs, org, mt, vs, 1 << 16)
    o = fp.Octree.build(ctx, m, p)
    # incremental refinement of a few listed leaves (ghm.cpp:518-521): pick leaves by their (x,y,z,extent) key
    rex, oex = r.export(), o.export()

    def leaf_keys(ex):
        c0 = ex["node_pos"][ex["corner"][:, 0]]
        ext = ex["node_pos"][ex["corner"][:, 1]][:, 0] - c0[:, 0]
        return np.concatenate([c0, ext[:, None]], 1), ex["first_child"] < 0
    rk, rleaf = leaf_keys(rex)
    ok, oleaf = leaf_keys(oex)
    rng = np.random.default_rng(5)
    pick = rng.choice(np.nonzero(rleaf)[0], 25, replace=False)
    lut = {tuple(k): i for i, k in enumerate(ok)}
    opick = np.array([lut[tuple(rk[i])] for i in pick], np.int32)
    r.refine(pick.astype(np.int32), 1 << 14)
    o.refine(m, opick, 1 << 14)
    assert_octree_equal(r.export(), o.export())
    o.close(); m.close()


def test_hex_connectivity_vs_reference(fp, ctx, ref):
    pm = fp.procedural
    V, H = pm.warped_hex_block(6)
    cases = [(H, len(V))]
    # an octree hex mesh has T-junction faces (a big face vs four small ones are DIFFERENT faces): many boundary faces
    Vt, Ft = pm.torus(40, 24)
    m = fp.TriMesh(ctx, Vt, Ft)
    p = fp.octree_grid_setup(Vt); p.c.stop_extent = 1 << 16
    o = fp.Octree.build(ctx, m, p)
    Vh, Hh, _ = o.hexes()
    cases.append((Hh, len(Vh)))
    for Hx, nV in cases:
        c = fp.HexConnectivity(ctx, Hx, nV)
        r = ref.hex_connectivity(Hx, nV)
        for k in ("F_vs", "F_es", "F_boundary", "E_vs", "E_boundary", "V_boundary", "H_fs"):
            assert np.array_equal(getattr(c, k), r[k]), k
        for k in fp.HexConnectivity.NAMES:
            off, val = getattr(c, k)
            assert np.array_equal(off, r[k][0]) and np.array_equal(val, r[k][1]), k
    o.close(); m.close()


def test_voxel_sign_vs_reference(fp, ctx, ref):
    for name, (V, F) in _meshes(fp).items():
        m = fp.TriMesh(ctx, V, F)
        mn, ext = V.min(0), V.max(0) - V.min(0)
        for spacing, padding in ((1 / 24, 1), (1 / 37.5, 2)):
            g = fp.VoxelGrid(mn, ext, spacing, padding)
            vox = fp.compute_sign_voxels(ctx, m, g)
            rvox, rdims = ref.voxel_sign(V, F, mn, ext, spacing, padding)
            assert np.array_equal(g.dims, rdims)
            assert np.array_equal(vox, rvox), (name, int((vox != rvox).sum()))
            assert vox.sum() > 0
            off, val = fp.compute_sign_dexels(ctx, m, g)
            roff, rval, rd2 = ref.dexel_sign(V, F, mn, ext, spacing, padding)
            assert np.array_equal(off, roff) and np.array_equal(val, rval), name
        m.close()


def test_octree_cell_sign_vs_reference(fp, ctx, ref):
    V, F = fp.procedural.torus(60, 40)
    mn, ext = V.min(0), V.max(0) - V.min(0)
    spacing = 1 / 64
    # compute_octree (voxelization.cpp:353-391): mesh_transform = 0, split down to extent 1
    gs = np.array([1 << int(np.ceil(np.log2(np.ceil(e / spacing)))) for e in ext], np.int32)
    m = fp.TriMesh(ctx, V, F)
    p = fp.OctreeParams(gs, mn, [0, 0, 0], spacing, 1)
    o = fp.Octree.build(ctx, m, p)
    r = ref.RefOctree.build(V, F, gs, mn, [0, 0, 0], spacing, 1)
    assert_octree_equal(r.export(), o.export())
    ins = o.cell_sign(m, mn, spacing)
    rins = r.cell_sign(mn, spacing)

    def keyed(ex, val):
        c0 = ex["node_pos"][ex["corner"][:, 0]]
        e = ex["node_pos"][ex["corner"][:, 1]][:, 0] - c0[:, 0]
        k = np.concatenate([c0, e[:, None], val[:, None].astype(np.int64)], 1)
        return k[np.lexsort(k.T[::-1])]
    assert np.array_equal(keyed(o.export(), ins), keyed(r.export(), rins))
    assert 0 < ins.sum() < len(ins)
    # the one public end-to-end entry, compute_octree: same leaves / inside flags
    rV, rH, rin = ref.compute_octree(V, F, mn, ext, spacing, 0, True, True)
    assert len(rH) == o.sizes()["leaves"] and int(rin.sum()) == int(ins[o.export()["first_child"] < 0].sum())
    o.close(); m.close()


def test_voxel_occupancy_matches_octree_predicate(fp, ctx, ref):
    V, F = fp.procedural.torus(40, 24)
    mn, ext = V.min(0), V.max(0) - V.min(0)
    spacing = 1 / 32
    gs = np.array([1 << int(np.ceil(np.log2(np.ceil(e / spacing)))) for e in ext], np.int32)
    m = fp.TriMesh(ctx, V, F)
    g = fp.VoxelGrid(mn, gs * spacing - 1e-9, spacing, 0)
    assert np.array_equal(g.dims, gs)
    occ = fp.voxel_occupancy(ctx, m, g)
    # reference: ungraded, unpaired bbox-predicate octree down to extent 1: a cell is split iff every ancestor's and its
    # own predicate is true, so every occupied voxel lies in a split extent-2 cell
    r = ref.RefOctree.build(V, F, gs, mn, [0, 0, 0], spacing, 1, graded=False, paired=False)
    ex = r.export()
    c0 = ex["node_pos"][ex["corner"][:, 0]]; e = ex["node_pos"][ex["corner"][:, 1]][:, 0] - c0[:, 0]
    split2 = c0[(e == 2) & (ex["first_child"] >= 0)]
    cover = np.zeros_like(occ)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                cover[split2[:, 2] + dz, split2[:, 1] + dy, split2[:, 0] + dx] = 1
    assert occ.sum() > 0 and np.all(cover[occ > 0] == 1)
    m.close()


def test_polyline_projection_vs_reference(fp, ctx, ref):
    V, F, crease = fp.procedural.gear(teeth=8, n_radial=3, n_axial=5, n_arc=2)
    rng = np.random.default_rng(2)
    loops = [np.arange(0, 40, dtype=np.int32), np.arange(100, 130, dtype=np.int32), np.array([5, 9, 17, 33, 2], np.int32)]
    circle = np.array([1, 1, 0], np.uint8)
    off = np.concatenate([[0], np.cumsum([len(l) for l in loops])]).astype(np.int64)
    cvs = np.concatenate(loops)
    P = rng.uniform(-0.5, 0.5, (3000, 3)); cid = rng.integers(0, 3, 3000).astype(np.int32)
    oL, aL = fp.polyline_project(ctx, V, off, cvs, circle, P, cid)
    roL, raL = ref.polyline_project(V, off, cvs, circle, P, cid)
    assert np.array_equal(oL, roL) and np.array_equal(aL, raL)


def test_hausdorff_vs_reference(fp, ctx, ref):
    pm = fp.procedural
    VA, FA = pm.torus(60, 40)
    VB, FB = pm.torus(33, 21)
    VB = VB * 1.01 + 0.003
    A, B = fp.TriMesh(ctx, VA, FA), fp.TriMesh(ctx, VB, FB)
    h = fp.hausdorff(ctx, A, B)
    r = ref.hausdorff(VA, FA, VB, FB)
    # north star: Hausdorff distance within 1e-5 relative (VCG's PointDistanceEP is not our exact closest point)
    np.testing.assert_allclose([h["diag"], h["max"], h["mean"]], [r["diag"], r["max"], r["mean"]], rtol=1e-5)
    assert h["diag"] == r["diag"]
    ok, ratio = ref.hausdorff_ratio(VA, FA, VB, FB, 0.005)
    np.testing.assert_allclose(h["ratio"], ratio, rtol=1e-5)
    assert h["n_ab"] == len(VA) and h["n_ba"] == len(VB)
    h2 = fp.hausdorff(ctx, A, B, extra_face_samples=200000)
    assert h2["n_ab"] > h["n_ab"] and h2["max_ab"] >= h["max_ab"] - 1e-15
    A.close(); B.close()


def test_hausdorff_outliers_vs_reference(fp, ctx, ref):
    """a12: against the COMPILED hausdorff_dis(mesh0, mesh1, outlierVs, thr) (gf.cpp:3590-3628, oracle/ref/ref_driver.cpp), not a
    restatement: same vertex set for a threshold that fires at once, one that has to decay (x0.9 rounds) and a gear pair."""
    pm = fp.procedural
    VA, FA = pm.torus(40, 24)
    VB, FB = pm.torus(30, 20)
    VB = VB.copy(); VB[:50] *= 1.05
    gV, gF, _ = pm.gear(teeth=12, n_radial=4, n_axial=6, n_arc=2)
    gV2 = gV.copy(); gV2[::37] += 0.004
    for (Va, Fa, Vb, Fb, thr) in [(VA, FA, VB, FB, 0.02), (VA, FA, VB, FB, 0.5), (gV, gF, gV2, gF, 0.003), (VB, FB, VA, FA, 0.01)]:
        A, B = fp.TriMesh(ctx, Va, Fa), fp.TriMesh(ctx, Vb, Fb)
        out = fp.hausdorff_outliers(ctx, A, B, thr)
        r = ref.hausdorff_dis_outliers(Va, Fa, Vb, Fb, thr)
        assert len(r) > 0 and len(np.unique(r)) == len(r)
        assert np.array_equal(np.sort(out), np.sort(r)), (thr, len(out), len(r))
        A.close(); B.close()


def test_voxel_lattice(fp, ctx, ref):
    """a13: against the COMPILED grid_hex_meshing_bijective::voxel_meshing (ghm.cpp:215-296) run on a GEO::Mesh — vertex positions
    bit-exact (float grid_length x int, then + min_corner in double), hexes identical, for three bounding boxes / resolutions."""
    pm = fp.procedural
    for (V, F), nv in [(pm.torus(40, 24), 40), (pm.gear(teeth=12, n_radial=4, n_axial=6, n_arc=2)[:2], 30), (pm.linked_tori(2, 16, 8), 57)]:
        V = V * np.array([1.0, 0.83, 0.61]) + np.array([0.013, -0.2, 0.37])      # not a unit box: the three grid lengths differ
        rV, rH = ref.voxel_meshing(V, F, nv)
        Vp, H, dim = fp.voxel_lattice(ctx, V.min(0), V.max(0), nv)
        assert int(np.prod(dim)) == len(rV) and len(H) == len(rH)
        assert np.array_equal(Vp, rV)
        assert np.array_equal(H, rH)


def test_octree_subdivide_existing_tree_vs_reference(fp, ctx, ref):
    """The later passes of the outer loop (ghm.cpp:495-500,523-524): OctreeGrid::subdivide called AGAIN on the tree of the previous
    pass with a smaller stop extent — also after an incremental refine of listed cells in between (ghm.cpp:518-521)."""
    for name, (V, F) in _meshes(fp).items():
        p = fp.octree_grid_setup(V, 1 << 20)
        gs, org, mt, vs = ref.octree_grid_setup(V, F, 1 << 20)
        m = fp.TriMesh(ctx, V, F)
        p.c.stop_extent = 1 << 16
        r = ref.RefOctree.build(V, F, gs, org, mt, vs, 1 << 16)
        o = fp.Octree.build(ctx, m, p)
        for E in (15, 14):
            r.subdivide(1 << E)
            o.subdivide(m, 1 << E)
            assert_octree_equal(r.export(), o.export())
            rV, rH, _ = r.hexes(); oV, oH, _ = o.hexes()
            assert np.array_equal(canon_hexes(rV, rH), canon_hexes(oV, oH)), (name, E)
        # equal to building at the final extent from scratch (the closure is a least fix-point)
        p.c.stop_extent = 1 << 14
        o2 = fp.Octree.build(ctx, m, p)
        assert_octree_equal(o.export(), o2.export())
        o2.close()
        # refine a few leaves, then one more pass
        rex, oex = r.export(), o.export()

        def leaf_keys(ex):
            c0 = ex["node_pos"][ex["corner"][:, 0]]
            ext = ex["node_pos"][ex["corner"][:, 1]][:, 0] - c0[:, 0]
            return np.concatenate([c0, ext[:, None]], 1), ex["first_child"] < 0
        rk, rleaf = leaf_keys(rex)
        okk, _ = leaf_keys(oex)
        big = np.nonzero(rleaf & (rk[:, 3] >= (1 << 15)))[0]
        pick = np.random.default_rng(3).choice(big, min(12, len(big)), replace=False)
        lut = {tuple(k): i for i, k in enumerate(okk)}
        opick = np.array([lut[tuple(rk[i])] for i in pick], np.int32)
        r.refine(pick.astype(np.int32), 1 << 13)
        o.refine(m, opick, 1 << 13)
        assert_octree_equal(r.export(), o.export())
        r.subdivide(1 << 13)
        o.subdivide(m, 1 << 13)
        assert_octree_equal(r.export(), o.export())
        o.close(); m.close()


def test_cpp_shim_parity():
    """The C++ drop-in shim (host/fpohm_shim.hpp, reference signatures and types) against the reference's own functions,
    both linked into one executable (tests/cpp/shim_parity.cpp, built by `make -C oracle/ref shim_parity`)."""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "shim_parity"
    if not exe.exists():
        pytest.skip("oracle/_ref/shim_parity not built (needs the reference headers)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "SHIM PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    assert r.stdout.count("PASS") >= 11


def test_shard_invariance_single_gpu(fp, ctx, ref):
    """an N-shard result must equal the 1-shard result bit for bit.  Shards are run one after the other on
    one GPU here; bench.py --gpus N runs them on N GPUs."""
    import torch
    from fpohm_b200 import sharding
    pm = fp.procedural
    dev = torch.device("cuda", 0)
    V, F = pm.torus(60, 40)
    m = fp.TriMesh(ctx, V, F)
    # --- queries by range
    P = np.random.default_rng(9).uniform(-0.6, 0.6, (10007, 3))
    S, I, C, N = m.signed_distance_pseudonormal(P)
    for W in (2, 3):
        parts = [m.signed_distance_pseudonormal(P[slice(*sharding.shard_range(len(P), r, W))]) for r in range(W)]
        for k, whole in enumerate((S, I, C, N)):
            assert np.array_equal(np.concatenate([p[k] for p in parts]), whole)
    # --- voxel grid by z slabs (aligned to 32 layers)
    mn, ext = V.min(0), V.max(0) - V.min(0)
    g = fp.VoxelGrid(mn, ext, 1 / 160, 1)
    whole = torch.empty(g.num_voxels(), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fp.voxel_sign_dev(ctx, m, g, whole.data_ptr(), st)
    nz, layer = int(g.dims[2]), int(g.dims[0]) * int(g.dims[1])
    for W in (2, 3):
        chunks = (nz + 31) // 32
        pieces = []
        for r in range(W):
            c0, c1 = sharding.shard_range(chunks, r, W)
            z0, z1 = c0 * 32, min(c1 * 32, nz)
            if z0 >= z1:
                continue
            buf = torch.empty((z1 - z0) * layer, dtype=torch.uint8, device=dev)
            fp.voxel_sign_slab_dev(ctx, m, g, z0, z1, buf.data_ptr(), st)
            pieces.append(buf)
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(pieces), whole)
    assert int(whole.sum()) > 0
    # --- hexes by range: per-hex values identical, statistics through the fixed-order reduction
    Vh, H = pm.warped_hex_block(12, 1.2)
    VJ, HJ, mad, fl = fp.scaled_jacobian(ctx, Vh, H)
    parts = [fp.scaled_jacobian(ctx, Vh, H[slice(*sharding.shard_range(len(H), r, 2))]) for r in range(2)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), VJ) and np.array_equal(np.concatenate([p[1] for p in parts]), HJ)
    assert min(p[2][0] for p in parts) == mad[0] and sum(p[3] for p in parts) == fl
    np.testing.assert_allclose(sum(p[2][1] * len(p[1]) for p in parts) / len(HJ), mad[1], rtol=1e-14)
    m.close()


def test_voxel_occupancy_bit_exact_vs_predicate(fp, ctx, ref, port):
    """Dense occupancy == the reference's should_subdivide predicate evaluated per extent-1 cell (voxelization.cpp:367-380),
    checked bit for bit against the brute-force restatement (oracle/port: port_box_overlaps_any) on an awkward grid."""
    import ctypes as C
    V, F = fp.procedural.torus(40, 24)
    m = fp.TriMesh(ctx, V, F)
    mn = V.min(0) - 0.0137
    g = fp.VoxelGrid(mn, V.max(0) - mn + 0.021, 1 / 23.7, 0)
    occ = fp.voxel_occupancy(ctx, m, g)
    tb = port.facet_boxes(V, F)
    L = port.lib()
    exp = np.zeros_like(occ)
    for z in range(g.dims[2]):
        for y in range(g.dims[1]):
            for x in range(g.dims[0]):
                exp[z, y, x] = L.port_box_overlaps_any(tb.ctypes.data_as(C.c_void_p), C.c_int64(len(tb)), C.c_double(g.origin[0]), C.c_double(g.origin[1]),
                                                       C.c_double(g.origin[2]), C.c_double(g.spacing), C.c_int(x), C.c_int(y), C.c_int(z), C.c_int(1))
    assert np.array_equal(occ, exp) and 0 < occ.sum() < occ.size
    # facets larger than 64 cells take the load-balanced pair path: coarse mesh on a fine grid
    V2, F2 = fp.procedural.torus(8, 5)
    m2 = fp.TriMesh(ctx, V2, F2)
    g2 = fp.VoxelGrid(V2.min(0), V2.max(0) - V2.min(0), 1 / 40, 1)
    occ2 = fp.voxel_occupancy(ctx, m2, g2)
    tb2 = port.facet_boxes(V2, F2)
    idx = np.random.default_rng(0).integers(0, occ2.size, 4000)
    for i in idx:
        z, r = divmod(int(i), int(g2.dims[0]) * int(g2.dims[1])); y, x = divmod(r, int(g2.dims[0]))
        e = L.port_box_overlaps_any(tb2.ctypes.data_as(C.c_void_p), C.c_int64(len(tb2)), C.c_double(g2.origin[0]), C.c_double(g2.origin[1]),
                                    C.c_double(g2.origin[2]), C.c_double(g2.spacing), C.c_int(x), C.c_int(y), C.c_int(z), C.c_int(1))
        assert occ2[z, y, x] == e
    # x extent a multiple of 16: the 16-byte expansion path (one thread = four 4 x 4 x 2 bit tiles); odd y / z extents clip the last tiles
    g3 = fp.VoxelGrid(V.min(0) - 0.01, np.array([48, 37, 21]) / 30.0 - 1e-9, 1 / 30.0, 0)
    assert g3.dims[0] % 16 == 0 and g3.dims[1] % 4 and g3.dims[2] % 2
    occ3 = fp.voxel_occupancy(ctx, m, g3)
    for i in np.random.default_rng(1).integers(0, occ3.size, 6000):
        z, r = divmod(int(i), int(g3.dims[0]) * int(g3.dims[1])); y, x = divmod(r, int(g3.dims[0]))
        e = L.port_box_overlaps_any(tb.ctypes.data_as(C.c_void_p), C.c_int64(len(tb)), C.c_double(g3.origin[0]), C.c_double(g3.origin[1]),
                                    C.c_double(g3.origin[2]), C.c_double(g3.spacing), C.c_int(x), C.c_int(y), C.c_int(z), C.c_int(1))
        assert occ3[z, y, x] == e
    assert 0 < occ3.sum() < occ3.size
    m.close(); m2.close()


def _sharded_octree(fp, V, F, prm, W):
    """Run build_octree_sharded with W ranks as W threads on cuda:0 (sharding.ThreadComm); returns rank exports + stats."""
    import threading
    from fpohm_b200 import sharding
    comms = sharding.ThreadComm.make(W)
    out, err, stats = [None] * W, [None] * W, [dict() for _ in range(W)]

    def run(r):
        try:
            c = fp.Context(0)
            m = fp.TriMesh(c, V, F)
            o = sharding.build_octree_sharded(fp, c, m, prm, comms[r], stats=stats[r])
            out[r] = (o.export(), o.flags())
            o.close(); m.close(); c.close()
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            comms[r].sh.barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(W)]
    [t.start() for t in th]; [t.join(timeout=300) for t in th]
    for e in err:
        if e is not None:
            raise e
    return out, stats


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["torus_e11", "gear_e12", "linked_e12_ungraded_pairs", "torus_e11_unpaired"])
def test_octree_zslab_sharded_equals_single(fp, ctx, case):
    """z-slab build with halo exchange of 2:1 candidates == single-GPU build, bit for bit, for
    every world size (ids included: phase 3 numbers the gathered sets canonically)."""
    pm = fp.procedural
    graded, paired = True, True
    if case == "torus_e11":
        V, F = pm.torus(48, 24); e = 11
    elif case == "gear_e12":
        V, F, _ = pm.gear(); e = 12
    elif case == "linked_e12_ungraded_pairs":
        V, F = pm.linked_tori(); e = 12; graded = False
    else:
        V, F = pm.torus(40, 20); e = 11; paired = False
    prm = fp.octree_grid_setup(V, 1 << 20)
    prm = fp.OctreeParams(prm.grid_size, prm.origin, prm.mesh_transform, prm.voxel_size, 1 << e, graded, paired)
    m = fp.TriMesh(ctx, V, F)
    whole = fp.Octree.build(ctx, m, prm)
    ew = whole.export()
    for W in (1, 2, 3, 5):
        outs, stats = _sharded_octree(fp, V, F, prm, W)
        for r in range(W):
            ex, fl = outs[r]
            for k in ("node_pos", "node_neigh", "first_child", "corner", "neigh"):
                assert np.array_equal(ex[k], ew[k]), (case, W, r, k)
            assert fl == whole.flags()
        if W > 1:
            b = stats[0]["slab_bounds"]
            assert b[0] == 0 and all(b[i] <= b[i + 1] for i in range(W)) and all(s["slab_bounds"] == b for s in stats)
            if graded:
                assert sum(sum(s["halo_codes"].values()) for s in stats) > 0      # something did cross a slab face
            total = sum(s["owned_true_cells"] for s in stats)
            assert total > 0, stats[0]
            print(case, W, "slabs", b, "owned", [s["owned_true_cells"] for s in stats], "halo", [sum(s["halo_codes"].values()) for s in stats])


@pytest.mark.gpu
def test_signed_distance_exact_ties_and_packets(fp, ctx, ref):
    """The packet search must name the SAME facet as igl even where the visiting order decides (exact and 1-ulp ties),
    for coherent packets, incoherent packets, stragglers and heavy queries alike.  Every output bit-identical."""
    pm = fp.procedural
    rng = np.random.default_rng(21)
    cases = {}
    V, F, _ = pm.gear(teeth=12, n_radial=4, n_axial=8, n_arc=3)
    cases["gear"] = (V, F)
    cases["torus"] = pm.torus(48, 32)
    cases["linked"] = pm.linked_tori()
    for name, (V, F) in cases.items():
        m = fp.TriMesh(ctx, V, F)
        rt = ref.RefTree(V, F)
        mn, mx = V.min(0), V.max(0)
        c = (mn + mx) / 2
        g = np.linspace(0, 1, 41)
        lattice = mn + (mx - mn) * np.stack(np.meshgrid(g, g, g[::4], indexing="ij"), -1).reshape(-1, 3)   # dyadic-ish, symmetric
        axis = np.stack([np.full(257, c[0]), np.full(257, c[1]), np.linspace(mn[2] - 0.1, mx[2] + 0.1, 257)], 1)  # on the symmetry axis
        fi = rng.integers(0, len(F), 3000)
        verts = V[rng.integers(0, len(V), 3000)]
        edges = V[F[fi]][:, :2].mean(1)
        off = verts + 1e-3 * rng.standard_normal((3000, 3))                                             # just off a vertex: fans of near-ties
        far = rng.uniform(mn - 3, mx + 3, (2000, 3))
        near = np.repeat(V[F[fi[:1500]]].mean(1), 4, 0) + 2e-3 * rng.standard_normal((6000, 3))         # coherent runs of 4
        P = np.concatenate([lattice, axis, verts, edges, off, far, near, axis[:5]])                         # odd total
        for order in ("given", "shuffled"):
            Q = P if order == "given" else P[rng.permutation(len(P))]
            S, I, C, N = m.signed_distance_pseudonormal(Q)
            rS, rI, rC, rN = rt.signed_distance(Q)
            bad = np.flatnonzero(I != rI)
            assert len(bad) == 0, (name, order, len(bad), bad[:5], I[bad[:5]], rI[bad[:5]])
            assert np.array_equal(S, rS) and np.array_equal(C, rC) and np.array_equal(N, rN), (name, order)
        m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["marks_unpaired", "marks_paired", "torus", "gear"])
def test_conforming_mesh_vs_reference(fp, ctx, ref, case):
    """conforming_mesh (ghm.cpp:568-696).  The reference function itself (compiled from
    grid_hex_meshing.cpp) runs on the product's octree tables, so every id and every list order must be equal."""
    rng = np.random.default_rng(5)
    if case.startswith("marks"):
        gs = np.array([32, 16, 16], np.int32)
        marks = [[x, y, z, 16] for x in (0, 16) for y in (0,) for z in (0,)]
        for e, k in ((8, 6), (4, 10), (2, 14)):
            n = gs // e
            for _ in range(k):
                marks.append([int(rng.integers(0, n[0])) * e, int(rng.integers(0, n[1])) * e, int(rng.integers(0, n[2])) * e, e])
        o = fp.Octree.from_marks(ctx, gs, np.array(marks, np.int32), True, case == "marks_paired")
        grid = gs
    else:
        V, F = fp.procedural.torus(48, 24) if case == "torus" else fp.procedural.gear()[:2]
        prm = fp.octree_grid_setup(V, 1 << 20)
        prm.c.stop_extent = 1 << (14 if case == "torus" else 14)
        o = fp.Octree.build(ctx, fp.TriMesh(ctx, V, F), prm)
        grid = prm.grid_size
    ex = o.export()
    Vp, H, _ = o.hexes()
    got = fp.conforming_mesh(ctx, o, H)
    want = ref.conforming_mesh_tables(ex["node_pos"], ex["node_neigh"], Vp, H, grid)
    assert got["n_replaced"] > 0, "the case has no T-junction face"
    for k in ("nV", "nF", "nH", "nE"):
        assert got[k] == want[k], (case, k, got[k], want[k])
    for k in ("F_off", "F_vs", "H_foff", "H_fs", "H_voff", "H_vs", "E_vs", "F_es", "F_boundary", "E_boundary", "V_boundary", "F_nhoff", "F_nhs"):
        assert np.array_equal(got[k], want[k]), (case, k)
    sizes = np.bincount(np.diff(got["F_off"]))
    assert sizes[:4].sum() == 0 and sizes[5:].sum() > 0        # loops of 4..8 vertices, some of them with inserted mid vertices


@pytest.mark.gpu
def test_conforming_mesh_tables_vs_golden(fp, ctx):
    """The committed fixtures (reference octree numbering, reference conforming_mesh output) through the table entry."""
    g = dict(np.load(Path(__file__).resolve().parent / "golden" / "golden_conforming_v1.npz"))
    for name in ("a", "b", "c"):
        got, dual = fp.conforming_and_dual_tables(ctx, g[f"{name}_node_pos"], g[f"{name}_node_neigh"], g[f"{name}_Vpos"], g[f"{name}_hex"], g[f"{name}_grid"])
        for k, v in g.items():
            if k.startswith(f"{name}_out_"):
                kk = k[len(name) + 5:]
                assert np.array_equal(np.asarray(got[kk]).reshape(-1), np.asarray(v).reshape(-1)), (name, kk)
            if k.startswith(f"{name}_dual_"):
                kk = k[len(name) + 6:]
                assert np.array_equal(np.asarray(dual[kk]).reshape(-1), np.asarray(v).reshape(-1)), (name, "dual", kk)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["marks_unpaired", "marks_paired", "torus", "gear"])
def test_dual_conforming_mesh_vs_reference(fp, ctx, ref, case):
    """SURVEY.md §8f-1, second half: dual_conforming_mesh (ghm.cpp:697-872) — dual vertices (bit-exact centres), ring
    faces, cells, their connectivity, element types and the template-ordered vertex lists, against the reference method."""
    rng = np.random.default_rng(6)
    if case.startswith("marks"):
        gs = np.array([32, 16, 16], np.int32)
        marks = [[x, 0, 0, 16] for x in (0, 16)]
        for e, k in ((8, 6), (4, 10), (2, 14)):
            n = gs // e
            for _ in range(k):
                marks.append([int(rng.integers(0, n[0])) * e, int(rng.integers(0, n[1])) * e, int(rng.integers(0, n[2])) * e, e])
        o = fp.Octree.from_marks(ctx, gs, np.array(marks, np.int32), True, case == "marks_paired")
        grid = gs
    else:
        V, F = fp.procedural.torus(48, 24) if case == "torus" else fp.procedural.gear()[:2]
        prm = fp.octree_grid_setup(V, 1 << 20)
        prm.c.stop_extent = 1 << 14
        o = fp.Octree.build(ctx, fp.TriMesh(ctx, V, F), prm)
        grid = prm.grid_size
    ex = o.export()
    Vp, H, _ = o.hexes()
    hyb, dual = fp.conforming_and_dual(ctx, o)
    rh, rd = ref.conforming_and_dual_tables(ex["node_pos"], ex["node_neigh"], Vp, H, grid)
    for k in ("nV", "nF", "nH", "nE"):
        assert dual[k] == rd[k], (case, k, dual[k], rd[k])
    assert np.array_equal(dual["V"], rd["V"]), case                                   # centres: same sums, bit for bit
    for k in ("F_off", "F_vs", "H_foff", "H_fs", "h_type", "H_voff", "H_vs", "E_vs", "F_es", "F_boundary", "E_boundary", "V_boundary", "F_nhoff", "F_nhs"):
        assert np.array_equal(dual[k], rd[k]), (case, k)
    assert np.array_equal(dual["census"][1:], np.bincount(rd["h_type"], minlength=7)[1:])
    assert len(np.unique(rd["h_type"])) >= 3                                            # the case exercises several templates


@pytest.mark.gpu
def test_classify_hexes_vs_reference(fp, ctx, ref):
    """clean_hex_mesh head (ghm.cpp:1937-1951): bbox centre of every hex + points_inside_mesh + sign, vs the reference's
    points_inside_mesh on the same centres."""
    V, F = fp.procedural.torus(48, 24)
    m = fp.TriMesh(ctx, V, F)
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << 14
    o = fp.Octree.build(ctx, m, prm)
    Vh, H, _ = o.hexes()
    S, flag = fp.classify_hexes(ctx, m, Vh, H)
    corners = Vh[H.astype(np.int64)]
    P = (corners.max(1) + corners.min(1)) / 2
    rS = ref.points_inside_mesh(V, F, P)
    assert np.array_equal(S, rS)
    assert np.array_equal(flag, (rS < 0).astype(np.uint8)) and 0 < flag.sum() < len(flag)


@pytest.mark.gpu
def test_conforming_dual_edge_cases(fp, ctx, ref):
    """No T-junction at all (uniform tree), and a single unsplit root (no interior edge or vertex: empty dual)."""
    for gs, marks in (([8, 8, 8], [[0, 0, 0, 8], [0, 0, 0, 4]]), ([4, 4, 4], []), ([8, 4, 4], [])):
        gs = np.array(gs, np.int32)
        o = fp.Octree.from_marks(ctx, gs, np.array(marks, np.int32).reshape(-1, 4), True, True)
        ex = o.export(); Vp, H, _ = o.hexes()
        hyb, dual = fp.conforming_and_dual(ctx, o)
        rh, rd = ref.conforming_and_dual_tables(ex["node_pos"], ex["node_neigh"], Vp, H, gs)
        assert hyb["n_replaced"] == 0
        for got, want in ((hyb, rh), (dual, rd)):
            for k, v in want.items():
                assert np.array_equal(np.asarray(got[k]).reshape(-1), np.asarray(v).reshape(-1)), (gs.tolist(), k)


@pytest.mark.gpu
def test_signed_distance_incoherent_batch_is_sorted_internally(fp, ctx, ref):
    """A large batch in random order takes the Morton-ordered path of the host entry point (>= 65 536 queries failing the
    coherence probe); results must still come back in the caller's order and equal igl's bit for bit."""
    V, F = fp.procedural.torus(64, 40)
    m = fp.TriMesh(ctx, V, F)
    rng = np.random.default_rng(33)
    fi = rng.integers(0, len(F), 70001)
    P = V[F[fi]].mean(1) + 0.02 * rng.standard_normal((70001, 3))          # near the surface, random order, odd count
    S, I, C, N = m.signed_distance_pseudonormal(P)
    rS, rI, rC, rN = ref.RefTree(V, F).signed_distance(P)
    assert np.array_equal(I, rI) and np.array_equal(S, rS) and np.array_equal(C, rC) and np.array_equal(N, rN)
    D, I2, C2 = m.point_mesh_squared_distance(P)                            # unsigned entry, same path
    assert np.array_equal(I2, rI) and np.array_equal(C2, rC)


# ---- §8(f)-2: clean_hex_mesh stages, against the reference's own methods (oracle/_ref) and the golden fixtures ------------
def test_clean_stages_vs_reference(fp, ctx, ref):
    from clean_cases import carved_block
    for dims, p, seed in [((6, 5, 4), 0.6, 11), ((8, 8, 8), 0.5, 12), ((12, 7, 9), 0.75, 13), ((9, 9, 9), 0.35, 14), ((16, 16, 16), 0.55, 15),
                          ((5, 4, 3), 1.0, 16), ((7, 7, 7), 0.08, 17)]:
        V, H, flag, Hm = carved_block(dims, p, seed)
        nV = len(V)
        got, n_mirrored = fp.reorder_hexes(ctx, V, Hm)
        want = ref.RefClean(V, Hm).reorder()
        assert np.array_equal(got, want) and n_mirrored == int((want != Hm).any(1).sum()), dims
        rc = ref.RefClean(V, H)
        rc.set_flags(flag)
        conn = fp.HexConnectivity(ctx, H, nV, keep=True)
        t, sweeps = fp.tag_uneven_elements(ctx, conn, flag)
        assert np.array_equal(t, rc.tagging()), (dims, "tagging")
        if not t.any():
            continue
        s = rc.reindex()
        mine = fp.reindex_submesh(ctx, H, nV, t)
        for k in ("V_map", "V_map_reverse", "H_map_reverse", "hex"):
            assert np.array_equal(mine[k], s[k]), (dims, k)
        n, rounds = fp.clean_non_manifold(ctx, H, nV, t)
        assert np.array_equal(n, rc.non_manifold()), (dims, "non-manifold", rounds)
        if not n.any():
            continue
        d, pieces = fp.drop_small_pieces(ctx, H, nV, n)
        assert np.array_equal(d, rc.drop_small()), (dims, "drop", pieces)
        # medial flags of an arbitrary flag set against the port-free definition through the reference's tables
        Fm, Vm = fp.medial_surface_flags(ctx, conn, d)
        off, val = conn.F_nhs
        two = np.diff(off) == 2
        want_F = np.where(two, d[val[off[:-1]]] != d[val[np.minimum(off[:-1] + 1, len(val) - 1)]], d[val[off[:-1]]] != 0)
        assert np.array_equal(Fm, want_F.astype(np.uint8))
        want_V = np.zeros(nV, np.uint8); want_V[conn.F_vs[want_F].reshape(-1)] = 1
        assert np.array_equal(Vm, want_V)
        conn.close()


def test_clean_hex_mesh_vs_reference_and_golden(fp, ctx, ref):
    from clean_cases import lattice_around
    g = dict(np.load(Path(__file__).resolve().parent / "golden" / "golden_clean_v1.npz"))
    for name in ("torus", "tori"):
        V, H, tV, tF = g[f"{name}_V"], g[f"{name}_hex"], g[f"{name}_tV"], g[f"{name}_tF"]
        m = fp.TriMesh(ctx, tV, tF)
        conn = fp.HexConnectivity(ctx, H, len(V), keep=True)
        out = fp.clean_hex_mesh(ctx, m, V, H, conn)
        assert np.array_equal(out["H_flag"], g[f"{name}_flag"]), name
        assert np.array_equal(out["F_medial"], g[f"{name}_F_medial"]) and np.array_equal(out["V_medial"], g[f"{name}_V_medial"]), name
        out2 = fp.clean_hex_mesh(ctx, m, V, H)                      # connectivity built inside
        assert np.array_equal(out2["H_flag"], out["H_flag"]) and np.array_equal(out2["V_medial"], out["V_medial"])
        conn.close(); m.close()
    # live reference on a larger lattice with mirrored hexes (reorder inside the composition) and a gear
    gV, gF = fp.procedural.gear(teeth=12, n_radial=4, n_axial=6, n_arc=2)[:2]
    for (tV, tF), n in (((gV, gF), 28), (fp.procedural.linked_tori(2, 24, 12), 40)):
        V, H = lattice_around(tV, n)
        rng = np.random.default_rng(5)
        mir = rng.random(len(H)) < 0.25
        H = H.copy(); H[mir] = H[mir][:, [3, 2, 1, 0, 7, 6, 5, 4]]
        rc = ref.RefClean(V, H)
        want = rc.full(tV, tF)
        wFm, wVm = rc.medial()
        m = fp.TriMesh(ctx, tV, tF)
        conn = fp.HexConnectivity(ctx, H, len(V), keep=True)
        out = fp.clean_hex_mesh(ctx, m, V, H, conn)
        assert out["stats"][0] == int(mir.sum())
        assert np.array_equal(out["H_flag"], want), (n, int((out["H_flag"] != want).sum()))
        assert np.array_equal(out["F_medial"], wFm) and np.array_equal(out["V_medial"], wVm)
        sub = fp.reindex_submesh(ctx, out["hex"], len(V), out["H_flag"])
        assert len(sub["H_map_reverse"]) == out["stats"][4] and len(sub["V_map_reverse"]) == out["stats"][5]
        conn.close(); m.close()


def test_clean_edge_cases(fp, ctx, ref):
    from clean_cases import block
    V, H = block(4, 4, 4)
    nV = len(V)
    # nothing inside: every stage leaves the flags alone and reports empty maps
    z = np.zeros(len(H), np.uint8)
    conn = fp.HexConnectivity(ctx, H, nV, keep=True)
    t, _ = fp.tag_uneven_elements(ctx, conn, z)
    rc = ref.RefClean(V, H); rc.set_flags(z)
    assert np.array_equal(t, rc.tagging())
    s = fp.reindex_submesh(ctx, H, nV, z)
    assert len(s["hex"]) == 0 and (s["V_map"] == -1).all()
    # two hexes sharing only an edge / only a vertex; a checkerboard (every interior vertex and edge non-manifold)
    for pick in ([0, 5], [0, 21], None):
        f = np.zeros(len(H), np.uint8)
        if pick is None:
            i, j, k = np.unravel_index(np.arange(len(H)), (4, 4, 4)); f[(i + j + k) % 2 == 0] = 1
        else:
            f[pick] = 1
        rc = ref.RefClean(V, H); rc.set_flags(f); rc.reindex()
        want = rc.non_manifold()
        got, rounds = fp.clean_non_manifold(ctx, H, nV, f)
        assert np.array_equal(got, want), (pick, rounds)
        if want.any():
            rc.reindex()
            d, pieces = fp.drop_small_pieces(ctx, H, nV, got)
            assert np.array_equal(d, rc.drop_small()), pick
    # equal-sized pieces: the reference keeps the one found first
    f = np.zeros(len(H), np.uint8); f[[0, 1, 62, 63]] = 1
    rc = ref.RefClean(V, H); rc.set_flags(f); rc.reindex()
    d, pieces = fp.drop_small_pieces(ctx, H, nV, f)
    assert pieces == 2 and np.array_equal(d, rc.drop_small()) and d[0] == 1 and d[63] == 0
    conn.close()
    with pytest.raises(fp.FpohmError):
        fp.reindex_submesh(ctx, H + 1000, nV, f)


def _same_surface(got, want, what):
    for k, v in want.items():
        if isinstance(v, tuple):
            assert np.array_equal(got[k][0], v[0]) and np.array_equal(got[k][1], v[1]), (what, k)
        else:
            assert np.array_equal(np.asarray(got[k]), np.asarray(v)), (what, k)


def test_extract_surface_vs_reference(fp, ctx, ref):
    """extract_surface_conforming_mesh + orient_surface_mesh (gf.cpp:1021-1112): quad and triangle surfaces of cleaned sub-meshes,
    with mirrored hexes (inconsistent quads), a cavity (second component keeps its order) and a whole block."""
    from clean_cases import block, carved_block, lattice_around
    cases = []
    V, H = block(3, 3, 3); keep = np.ones(len(H), bool); keep[13] = False
    cases.append(("cavity", V, H[keep]))
    cases.append(("block", *block(5, 4, 3)))
    for dims, p, seed in [((8, 8, 8), 0.5, 2), ((14, 12, 10), 0.7, 3)]:
        V, H, flag, _ = carved_block(dims, p, seed)
        conn = fp.HexConnectivity(ctx, H, len(V), keep=True)
        f, _ = fp.tag_uneven_elements(ctx, conn, flag); conn.close()
        f, _ = fp.clean_non_manifold(ctx, H, len(V), f)
        f, _ = fp.drop_small_pieces(ctx, H, len(V), f)
        s = fp.reindex_submesh(ctx, H, len(V), f)
        sh = s["hex"].copy(); sh[::3] = sh[::3][:, [3, 2, 1, 0, 7, 6, 5, 4]]
        cases.append((f"carved{dims}", V[s["V_map_reverse"]], sh))
    tV, tF = fp.procedural.torus(60, 40)
    V, H = lattice_around(tV, 40)
    m = fp.TriMesh(ctx, tV, tF)
    out = fp.clean_hex_mesh(ctx, m, V, H); m.close()
    s = fp.reindex_submesh(ctx, out["hex"], len(V), out["H_flag"])
    cases.append(("torus", V[s["V_map_reverse"]], s["hex"]))
    for name, V, H in cases:
        conn = fp.HexConnectivity(ctx, H, len(V), keep=True)
        for tri in (False, True):
            got = fp.extract_surface(ctx, conn, V, tri)
            want = ref.extract_surface(V, H, tri)
            _same_surface(got, want, (name, tri))
        conn.close()
    # a non-manifold edge in the component of face 0: the reference's result is an accident of its queue order -> refused
    V, H = block(2, 2, 1)
    conn = fp.HexConnectivity(ctx, H[[0, 3]], len(V), keep=True)
    with pytest.raises(fp.FpohmError):
        fp.extract_surface(ctx, conn, V, False)
    conn.close()


# ---- §8(f)-3: SLIM per-element stages (slim_m.cpp), against the compiled reference functions; tolerance of the north star 1e-5 ----
def _slim_jacobian_set(n, seed):
    rng = np.random.default_rng(seed)
    J = np.eye(3)[None] + rng.normal(0, 0.35, (n, 3, 3))
    J[::7] = J[::7] @ np.diag([1, 1, -1.0])                      # inverted elements: reflection branch of polar_svd
    J[1] = np.eye(3); J[2] = np.diag([2.0, 1.0, 0.5]); J[3] = np.diag([1.0 + 5e-9, 1.0, 1.0 - 5e-9])     # |s - 1| < eps branch
    Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    J[4] = Q @ np.diag([3.0, 3.0, 0.2]) @ Q.T                      # repeated singular value
    J[5] = 1e-3 * J[5]; J[6] = 1e3 * J[6]                          # scale extremes
    return J.reshape(n, 9)


def test_slim_weights_rotations_and_energy_vs_reference(fp, ctx, ref):
    J = _slim_jacobian_set(20000, 11)
    areas = np.random.default_rng(12).uniform(0.5, 2.0, len(J))
    for en in fp.SLIM_ENERGIES:
        W, Ri = fp.slim_weights_rotations(ctx, J, en, 0.7)
        rW, rRi = ref.slim_weights_rotations(J, en, 0.7)
        fin = np.isfinite(rW).all(1) & np.isfinite(rRi).all(1)
        assert np.array_equal(fin, np.isfinite(W).all(1) & np.isfinite(Ri).all(1)), en      # overflow / 0-over-0 rows coincide
        scale_w = np.abs(rW[fin]).max(1, keepdims=True); scale_r = np.abs(rRi[fin]).max(1, keepdims=True)
        assert (np.abs(W[fin] - rW[fin]) <= 1e-9 * scale_w).all(), (en, float((np.abs(W[fin] - rW[fin]) / scale_w).max()))
        assert (np.abs(Ri[fin] - rRi[fin]) <= 1e-9 * scale_r).all(), (en, float((np.abs(Ri[fin] - rRi[fin]) / scale_r).max()))
        sub = fin & (np.abs(J).max(1) < 50)           # the energy sum of the whole set overflows for the exponential energies
        e = fp.slim_energy(ctx, J[sub], areas[sub], en, 0.7); re_ = ref.slim_energy(J[sub], areas[sub], en, 0.7)
        assert (np.isfinite(re_) and abs(e - re_) <= 1e-10 * abs(re_)) or (not np.isfinite(re_) and not np.isfinite(e)), (en, e, re_)
    # rotations really are rotations / W really is symmetric (size-independent properties)
    W, Ri = fp.slim_weights_rotations(ctx, J, "SYMMETRIC_DIRICHLET")
    R = Ri.reshape(-1, 3, 3)
    ok = np.isfinite(R).all((1, 2))
    assert np.abs(R[ok] @ np.swapaxes(R[ok], 1, 2) - np.eye(3)).max() < 1e-12 and (np.linalg.det(R[ok]) > 0).all()
    Wm = W.reshape(-1, 3, 3); okw = np.isfinite(Wm).all((1, 2))
    assert np.abs(Wm[okw] - np.swapaxes(Wm[okw], 1, 2)).max() <= 1e-12 * np.abs(Wm[okw]).max()
    import ctypes as C
    out = np.zeros(9)
    rc = fp.lib().fpohm_slim_weights_rotations(ctx.h, J.ctypes.data_as(C.c_void_p), C.c_int64(1), C.c_int32(9), C.c_double(1.0),
                                               out.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc != 0                                   # unknown SLIM_ENERGY is refused, not guessed


def test_slim_jacobians_vs_reference(fp, ctx, ref):
    """compute_jacobians on the gradient operator of a tet mesh: 4 non-zeros per row, Dx/Dy/Dz sharing the pattern (igl::grad)."""
    rng = np.random.default_rng(21)
    V, H = fp.procedural.warped_hex_block(12, 0.3)
    tets = H[:, [[0, 1, 3, 4], [1, 2, 0, 5], [2, 3, 1, 6], [3, 0, 2, 7], [4, 7, 5, 0], [5, 4, 6, 1], [6, 5, 7, 2], [7, 6, 4, 3]]].reshape(-1, 4).astype(np.int64)
    P = V[tets]                                                   # per tet: gradient of the 4 hat functions = rows of inv([p1-p0, p2-p0, p3-p0])
    E = np.stack([P[:, 1] - P[:, 0], P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]], 1)
    G = np.linalg.inv(E)                                          # G[:, :, k] = gradient of hat function k+1
    g = np.concatenate([-G.sum(2, keepdims=True), G], 2)          # 3 x 4 per tet
    off = np.arange(0, 4 * len(tets) + 1, 4); col = tets.reshape(-1).astype(np.int32)
    vx, vy, vz = g[:, 0].reshape(-1), g[:, 1].reshape(-1), g[:, 2].reshape(-1)
    uv = V + rng.normal(0, 0.01, V.shape)
    Ji = fp.slim_jacobians(ctx, off, col, vx, vy, vz, uv)
    rJ = ref.slim_jacobians(off, col, vx, vy, vz, uv, len(V))
    assert np.abs(Ji - rJ).max() <= 1e-12 * np.abs(rJ).max()
    # the deformation gradient of the identity map is the identity
    I = fp.slim_jacobians(ctx, off, col, vx, vy, vz, V)
    assert np.abs(I - np.eye(3).reshape(9)).max() < 1e-9


def test_slim_max_step_vs_reference(fp, ctx, ref):
    """compute_max_step_from_singularities (igl/flip_avoiding_line_search.cpp, tets): per-tet smallest positive roots and their minimum,
    for search directions from tiny to mesh-inverting, incl. a zero direction (no root at all) and a rigid translation."""
    V, H = fp.procedural.warped_hex_block(10, 0.3)
    T = H[:, [[0, 1, 3, 4], [1, 2, 0, 5], [2, 3, 1, 6], [3, 0, 2, 7], [4, 7, 5, 0], [5, 4, 6, 1], [6, 5, 7, 2], [7, 6, 4, 3]]].reshape(-1, 4).astype(np.int32)
    rng = np.random.default_rng(31)
    for sc in (1e-4, 0.02, 0.2, 2.0):
        d = rng.normal(0, sc, V.shape)
        m, r = fp.slim_max_step(ctx, V, T, d)
        rm, rr = ref.slim_max_step(V, T, d)
        fin = np.isfinite(rr)
        assert np.array_equal(fin, np.isfinite(r)), sc
        assert (np.abs(r[fin] - rr[fin]) <= 1e-8 * np.abs(rr[fin])).all(), (sc, float((np.abs(r[fin] - rr[fin]) / np.abs(rr[fin])).max()))
        assert abs(m - rm) <= 1e-9 * abs(rm), (sc, m, rm)
        if sc < 0.01:
            continue      # |det(D)| <= 1e-10 there: the reference (and we) drop the cubic term, the bound is only approximate
        # the bound is tight: just before it no tet is inverted, just after it one is
        vol = lambda X: np.einsum("ij,ij->i", X[T[:, 1]] - X[T[:, 0]], np.cross(X[T[:, 2]] - X[T[:, 0]], X[T[:, 3]] - X[T[:, 0]]))
        s0 = np.sign(vol(V))
        assert (np.sign(vol(V + 0.999 * m * d)) == s0).all() and (np.sign(vol(V + 1.001 * m * d)) != s0).any(), sc
    m, r = fp.slim_max_step(ctx, V, T, np.zeros_like(V))
    rm, rr = ref.slim_max_step(V, T, np.zeros_like(V))
    assert m == rm == np.inf and np.array_equal(r, rr)
    # rigid translation: no tet ever degenerates.  The difference form gives exactly zero coefficients -> +inf; the reference's
    # expansion in absolute coordinates keeps rounding noise in them and reports a spurious root of order 1e12 (documented deviation)
    m, r = fp.slim_max_step(ctx, V, T, np.ones_like(V) * 0.3)
    rm, rr = ref.slim_max_step(V, T, np.ones_like(V) * 0.3)
    assert m == np.inf and rm > 1e9
    with pytest.raises(fp.FpohmError):
        fp.slim_max_step(ctx, V, T + len(V), V)


def test_clean_hex_mesh_refuses_bad_ids(fp, ctx):
    from clean_cases import block
    V, H = block(3, 3, 3)
    tV, tF = fp.procedural.torus(20, 12)
    m = fp.TriMesh(ctx, tV, tF)
    bad = H.copy(); bad[5, 3] = len(V) + 7
    with pytest.raises(fp.FpohmError, match="out of range"):
        fp.clean_hex_mesh(ctx, m, V, bad)
    m.close()


def test_slim_rhs_terms_bit_exact(fp, ctx, ref):
    """per-element part of buildRhs (slim_m.cpp:1061-1083) against the reference function run with At = I: same IEEE operations."""
    J = _slim_jacobian_set(5000, 41)
    W, Ri = fp.slim_weights_rotations(ctx, J, "SYMMETRIC_DIRICHLET")
    ok = np.isfinite(W).all(1) & np.isfinite(Ri).all(1)
    got = fp.slim_rhs_terms(ctx, W[ok], Ri[ok]); want = ref.slim_rhs_terms(W[ok], Ri[ok])
    assert np.array_equal(got, want)


def test_device_tree_build_identical_to_host_and_reference(fp, ctx, ref):
    """a7: igl::AABB::init built ON THE DEVICE (tree_device.cu: radix sorts of the barycentre columns, level-synchronous median
    splits) — box, primitive and child arrays of every node equal the host builder's and the compiled igl tree's, on meshes whose
    barycentre coordinates are all distinct (no host involvement at all) and on meshes full of ties (symmetric gear, translated
    tori, an axis-aligned grid: the tied axes take their ranks from the host's std::sort)."""
    pm = fp.procedural
    rng = np.random.default_rng(17)
    Vt, Ft = pm.torus(60, 40)
    Vj = Vt + rng.normal(0, 1e-4, Vt.shape)                       # generic position: no two barycentre coordinates equal
    g = np.arange(13, dtype=np.float64) / 12
    X, Y = np.meshgrid(g, g, indexing="ij")
    Vg = np.stack([X.reshape(-1), Y.reshape(-1), 0.1 * np.sin(3 * X.reshape(-1))], -1)      # grid: massive ties on x and y
    q = np.arange(13 * 13).reshape(13, 13)
    Fg = np.concatenate([np.stack([q[:-1, :-1], q[1:, :-1], q[1:, 1:]], -1).reshape(-1, 3),
                         np.stack([q[:-1, :-1], q[1:, 1:], q[:-1, 1:]], -1).reshape(-1, 3)]).astype(np.int32)
    cases = {"jittered torus": (Vj, Ft), "torus": (Vt, Ft), "gear": pm.gear(teeth=12, n_radial=4, n_axial=6, n_arc=2)[:2],
             "tori": pm.linked_tori(2, 16, 8), "grid": (Vg, Fg), "gear full": pm.gear()[:2]}
    for name, (V, F) in cases.items():
        m = fp.TriMesh(ctx, V, F)
        m.build_aabb_tree()
        box, prim, lr = m.tree()
        hbox, hprim, hlr = fp.host_igl_tree(V, F)
        assert np.array_equal(prim, hprim) and np.array_equal(lr, hlr), name
        assert np.array_equal(box, hbox), name
        # the three normal sets + E / EMAP, built on the device (normals_device.cu; only acos of the corner angles runs on the host)
        dn = m.normals()
        hn = fp.host_igl_normals(V, F)
        for a, b, what in zip(dn, hn, ("FN", "VN", "EN", "E", "EMAP")):
            assert np.array_equal(a, b, equal_nan=True), (name, what)
        if len(F) < 50000:
            rt = ref.RefTree(V, F)
            rbox, rprim, rlr = rt.flatten()
            assert np.array_equal(prim, rprim) and np.array_equal(lr, rlr) and np.array_equal(box, rbox), name
            rFN, rVN, rEN, rE, rEMAP = rt.normals()
            assert np.array_equal(dn[0], rFN) and np.array_equal(dn[1], rVN, equal_nan=True) and np.array_equal(dn[2], rEN) and np.array_equal(dn[4], rEMAP), name
        m.close()

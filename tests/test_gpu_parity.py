"""GPU parity tests: CUDA path (through the C-ABI) vs the reference's own code (oracle/_ref)."""
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
        same = I == rI
        # where the facet agrees everything is the same arithmetic
        assert np.array_equal(S[same], rS[same]) and np.array_equal(C[same], rC[same]) and np.array_equal(N[same], rN[same]), name
        # disagreement only at distance ties (north star: "closest-primitive index equal except at ties")
        assert same.mean() > 0.999, (name, same.mean())
        np.testing.assert_allclose(np.abs(S[~same]), np.abs(rS[~same]), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(C[~same], rC[~same], rtol=1e-9, atol=1e-12)
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
    same = I == rI
    assert same.mean() > 0.999 and np.array_equal(D[same], rD[same]) and np.array_equal(C[same], rC[same])
    np.testing.assert_allclose(D, rD, rtol=1e-12)


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

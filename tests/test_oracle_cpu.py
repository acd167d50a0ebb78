"""CPU suite (-m "not gpu"): the oracle port against the golden vectors (and against oracle/_ref when present), the host
logic of the product library, the C-ABI surface, and the multi-process sharding logic over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from canon import canon_octree, canon_hexes

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def G():
    return dict(np.load(ROOT / "tests" / "golden" / "golden_v1.npz"))


def _canon_equal(canon, G, prefix):
    for k, v in canon.items():
        g = G[f"{prefix}_{k}"]
        assert v.shape == g.shape and np.array_equal(v, g), f"{prefix}_{k}"


# ---- oracle port vs golden (golden = reference output) ----------------------------------------------------------
def test_port_scaled_jacobian(port, G):
    VJ, HJ, mad, fl = port.scaled_jacobian(G["jac_V"], G["jac_H"])
    assert np.array_equal(VJ, G["jac_VJ"]) and np.array_equal(HJ, G["jac_HJ"]) and fl == int(G["jac_flipped"])
    assert np.array_equal(mad, G["jac_mad"])          # same sequential sums as gf.cpp:2341-2354
    assert fl > 0                                     # the fixture really has flipped hexes


def test_port_tree_normals_signed_distance(port, G):
    t = port.PortTree(G["sd_V"], G["sd_F"])
    box, prim, lr = t.flatten()
    assert np.array_equal(box, G["sd_box"]) and np.array_equal(prim, G["sd_prim"]) and np.array_equal(lr, G["sd_lr"])
    FN, VN, EN, EMAP = t.normals()
    assert np.array_equal(FN, G["sd_FN"]) and np.array_equal(VN, G["sd_VN"]) and np.array_equal(EN, G["sd_EN"]) and np.array_equal(EMAP, G["sd_EMAP"])
    S, I, Cc, N = t.signed_distance(G["sd_P"])
    assert np.array_equal(S, G["sd_S"]) and np.array_equal(I, G["sd_I"]) and np.array_equal(Cc, G["sd_C"]) and np.array_equal(N, G["sd_N"])
    D2 = t.signed_distance(G["sd_P"], with_sign=False)[0]
    assert np.array_equal(D2, G["sd_D2"])
    assert (S < 0).any() and (S > 0).any()


def test_port_signed_distance_on_tied_mesh(port, G):
    """Structured mesh with tied barycentre coordinates: the port's tree may differ from igl's (documented in
    fpohm_port.c), so the facet index may differ — but only at exact distance ties."""
    t = port.PortTree(G["tie_V"], G["tie_F"])
    S, I, Cc, N = t.signed_distance(G["tie_P"])
    assert np.array_equal(np.abs(S), np.abs(G["tie_S"]))
    same = I == G["tie_I"]
    assert np.array_equal(S[same], G["tie_S"][same]) and np.array_equal(Cc[same], G["tie_C"][same]) and np.array_equal(N[same], G["tie_N"][same])
    np.testing.assert_allclose(Cc, G["tie_C"], rtol=0, atol=1e-12)
    assert same.mean() > 0.9


def test_port_octree(port, G):
    gs, org, mt, vs = port.octree_grid_setup(G["sd_V"], 1 << 20)
    assert np.array_equal(gs, G["oct_gs"]) and np.array_equal(org, G["oct_origin"]) and np.array_equal(mt, G["oct_mt"]) and vs == float(G["oct_vs"])
    for E in (17, 16):
        o = port.octree_build(G["sd_V"], G["sd_F"], gs, org, mt, vs, 1 << E)
        ex = o.export()
        _canon_equal(canon_octree(ex), G, f"oct{E}")
        Vp, Hx, _ = port.octree_hexes(ex, org, mt, vs)
        assert np.array_equal(canon_hexes(Vp, Hx), G[f"oct{E}_hexpos"])


def test_port_octree_marks_all_modes(port, G):
    for ci in range(int(G["marks_n"])):
        gr, pa = G[f"marks{ci}_mode"]
        o = port.octree_from_marks(G[f"marks{ci}_gs"], G[f"marks{ci}_marks"], bool(gr), bool(pa))
        _canon_equal(canon_octree(o.export()), G, f"marks{ci}")


def test_port_ray_parity(port, G):
    V, F = G["sd_V"], G["sd_F"]
    mn, ext = V.min(0), V.max(0) - V.min(0)
    vox, dims = port.voxel_sign(V, F, mn, ext, float(G["vox_spacing"]), int(G["vox_pad"]))
    assert np.array_equal(dims, G["vox_dims"]) and np.array_equal(vox, G["vox_out"]) and vox.sum() > 0
    off, val, _ = port.dexel_sign(V, F, mn, ext, float(G["vox_spacing"]), int(G["vox_pad"]))
    assert np.array_equal(off, G["dex_off"]) and np.array_equal(val, G["dex_val"])
    sp = float(G["cs_spacing"])
    o = port.octree_build(V, F, G["cs_gs"], mn, [0, 0, 0], sp, 1)
    ex = o.export()
    ins = port.octree_cell_sign(V, F, ex, mn, sp)
    c0 = ex["node_pos"][ex["corner"][:, 0]]; e = ex["node_pos"][ex["corner"][:, 1]][:, 0] - c0[:, 0]
    key = np.concatenate([c0, e[:, None], ins[:, None].astype(np.int64)], 1)
    assert np.array_equal(key[np.lexsort(key.T[::-1])], G["cs_keyed"])


def test_port_connectivity(port, G):
    c = port.hex_connectivity(G["conn_H"], int(G["conn_nV"]))
    for k in ("F_vs", "F_es", "F_boundary", "E_vs", "E_boundary", "V_boundary", "H_fs"):
        assert np.array_equal(c[k], G[f"conn_{k}"]), k
    for k in ("F_nhs", "E_nfs", "E_nhs", "V_nvs", "V_nes", "V_nfs", "V_nhs"):
        assert np.array_equal(c[k][0], G[f"conn_{k}_off"]) and np.array_equal(c[k][1], G[f"conn_{k}_val"]), k


def test_port_polyline_and_hausdorff(port, G):
    oL, aL = port.polyline_project(G["sd_V"], G["pl_off"], G["pl_vs"], G["pl_circle"], G["pl_P"], G["pl_cid"])
    assert np.array_equal(oL, G["pl_origin"]) and np.array_equal(aL, G["pl_axis"])
    h = port.hausdorff(G["sd_V"], G["sd_F"], G["hd_VB"], G["hd_FB"])
    assert h["diag"] == G["hd_out"][0]
    np.testing.assert_allclose([h["max"], h["mean"]], G["hd_out"][1:], rtol=1e-5)   # tolerance of the north star


# ---- golden vs the live reference (only where oracle/_ref exists): keeps the fixtures honest ------------------------
def test_golden_matches_live_reference(ref, G):
    VJ, HJ, mad, fl = ref.scaled_jacobian(G["jac_V"], G["jac_H"])
    assert np.array_equal(VJ, G["jac_VJ"]) and np.array_equal(mad, G["jac_mad"])
    S, I, Cc, N = ref.RefTree(G["sd_V"], G["sd_F"]).signed_distance(G["sd_P"])
    assert np.array_equal(S, G["sd_S"]) and np.array_equal(I, G["sd_I"])
    ro = ref.RefOctree.build(G["sd_V"], G["sd_F"], G["oct_gs"], G["oct_origin"], G["oct_mt"], float(G["oct_vs"]), 1 << 16)
    _canon_equal(canon_octree(ro.export()), G, "oct16")


def test_reference_random_split_fuzz_is_valid(ref):
    """OctreeGrid::testSubdivideRandom (octree.cpp:900-961), the reference's only self-check, runs clean."""
    o = ref.RefOctree.random([16, 16, 8], True, True)
    assert o.flags() == (True, True) and o.sizes()["cells"] > 2


# ---- host logic of the product (no GPU needed) -------------------------------------------------------------------------
def test_host_igl_tree_identical_to_reference_even_with_ties(fp, G):
    for tag in ("sd", "tie"):     # "tie": structured gear, equal barycentre coordinates -> std::sort / nth_element tie order
        box, prim, lr = fp.host_igl_tree(G[f"{tag}_V"], G[f"{tag}_F"])
        assert np.array_equal(box, G[f"{tag}_box"]) and np.array_equal(prim, G[f"{tag}_prim"]) and np.array_equal(lr, G[f"{tag}_lr"]), tag
    FN, VN, EN, E, EMAP = fp.host_igl_normals(G["sd_V"], G["sd_F"])
    assert np.array_equal(FN, G["sd_FN"]) and np.array_equal(VN, G["sd_VN"]) and np.array_equal(EN, G["sd_EN"]) and np.array_equal(EMAP, G["sd_EMAP"])


def test_host_tree_of_a_large_tied_mesh_equals_compiled_igl(fp, ref):
    """Above 131 072 facets the barycentre columns are sorted by std::sort's own pieces on several threads (sort_like_std,
    igl_tree_host.cpp).  The full bench gear has 150 k - 200 k tied barycentre coordinates per axis, so any deviation from libstdc++'s
    order of equal keys shows up as a different tree: node for node against igl::AABB::init compiled from the reference."""
    V, F = fp.procedural.gear()[:2]
    assert len(F) >= 4 * 32768
    box, prim, lr = fp.host_igl_tree(V, F)
    rbox, rprim, rlr = ref.RefTree(V, F).flatten()
    assert np.array_equal(prim, rprim) and np.array_equal(lr, rlr) and np.array_equal(box, rbox)


@pytest.mark.parametrize("pattern", ["sorted", "reversed", "organ_pipe", "three_values", "sawtooth"])
def test_host_tree_sort_patterns_equal_compiled_igl(fp, ref, pattern):
    """The multi-threaded restatement of std::sort (sort_like_std) on inputs that stress introsort's partitioning: 140 000 disjoint
    triangles whose barycentre x follows a pattern (already sorted, reversed, organ pipe, three distinct values, sawtooth), y and z
    random with many ties.  The tree must equal igl::AABB::init compiled from the reference node for node."""
    n = 140000
    rng = np.random.default_rng(11)
    i = np.arange(n, dtype=np.float64)
    x = {"sorted": i, "reversed": n - i, "organ_pipe": np.minimum(i, n - i), "three_values": np.floor(i * 3 / n),
         "sawtooth": np.mod(i, 1000.0)}[pattern]
    c = np.stack([x * 1e-3, np.floor(rng.uniform(0, 50, n)) * 0.1, np.floor(rng.uniform(0, 7, n))], 1)
    off = np.array([[0.0, 0.0, 0.0], [0.03, 0.0, 0.0], [0.0, 0.03, 0.0]]) - np.array([0.01, 0.01, 0.0])
    V = (c[:, None, :] + off[None, :, :]).reshape(-1, 3)
    F = np.arange(3 * n, dtype=np.int32).reshape(-1, 3)
    box, prim, lr = fp.host_igl_tree(V, F)
    rbox, rprim, rlr = ref.RefTree(V, F).flatten()
    assert np.array_equal(prim, rprim) and np.array_equal(lr, rlr) and np.array_equal(box, rbox)


def test_host_grid_setups(fp, port, G):
    p = fp.octree_grid_setup(G["sd_V"], 1 << 20)
    assert np.array_equal(p.grid_size, G["oct_gs"]) and np.array_equal(p.origin, G["oct_origin"])
    assert np.array_equal(p.mesh_transform, G["oct_mt"]) and p.voxel_size == float(G["oct_vs"])
    V = G["sd_V"]
    g = fp.VoxelGrid(V.min(0), V.max(0) - V.min(0), float(G["vox_spacing"]), int(G["vox_pad"]))
    assert np.array_equal(g.dims, G["vox_dims"])
    dims, o = port.voxel_grid_setup(V.min(0), V.max(0) - V.min(0), float(G["vox_spacing"]), int(G["vox_pad"]))
    assert np.array_equal(g.origin, o)


def test_tiny_meshes_host_tree(fp):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    for F in ([[0, 1, 2]], [[0, 1, 2], [0, 1, 3]], [[0, 1, 2], [0, 1, 3], [0, 2, 3]]):   # 1, 2 (igl::sort2), 3 (igl::sort3) facets
        box, prim, lr = fp.host_igl_tree(V, np.array(F, np.int32))
        assert len(prim) == 2 * len(F) - 1 and sorted(prim[prim >= 0].tolist()) == list(range(len(F)))


# ---- C-ABI surface --------------------------------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol(fp):
    hdr = (ROOT / "include" / "fpohm.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fpohm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    lib = fp.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", str(fp.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fpohm_[a-z0-9_]+)", out))
    assert exported == set(declared), (exported ^ set(declared))


def test_abi_argument_validation_and_no_cpu_fallback(fp):
    lib = fp.lib()
    assert lib.fpohm_octree_grid_setup(None, C.c_int64(0), C.c_int32(1), None) == -1       # FPOHM_EINVAL
    assert b"bad argument" in lib.fpohm_last_error()
    if fp.device_count() == 0:
        h = C.c_void_p()
        assert lib.fpohm_ctx_create(C.c_int(0), C.byref(h)) == -2                           # FPOHM_ENODEV, loud
        assert b"no CPU fallback" in lib.fpohm_last_error()
        with pytest.raises(fp.FpohmError):
            fp.Context(0)


def test_product_never_imports_the_oracle():
    """The oracle is a checker: nothing under the product package (or the import shim) may reference it."""
    pkg = ROOT / "feature-preserving-octree-hex-meshing_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + [ROOT / "fpohm_b200.py"]:
        txt = p.read_text()
        assert "oracle" not in txt.replace("oracle/", "").lower() or p.name == "build.py", p
        assert "port_oracle" not in txt and "ref_oracle" not in txt and "libfpohm_ref" not in txt and "libfpohm_port" not in txt, p


# ---- multi-process sharding logic (gloo, world_size 2) ------------------------------------------------------------------
def test_query_sharding_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {str(ROOT)!r})
import fpohm_b200 as fp
from fpohm_b200 import sharding
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
n = 1001
lo, hi = sharding.shard_range(n, r, w)
x = torch.arange(n, dtype=torch.float64)
part = (x[lo:hi] * 2).contiguous()                     # stands in for a per-rank kernel result
full = sharding.gather_ranges(part, n, r, w)
assert torch.equal(full, x * 2), "gather"
st = sharding.reduce_stats(dict(min=float(part.min()), sum=float(part.sum()), sumsq=float((part*part).sum()), count=hi-lo, max=float(part.max())))
assert st["count"] == n and st["min"] == 0.0 and st["max"] == 2.0*(n-1) and abs(st["sum"] - float((x*2).sum())) < 1e-6
zs = sharding.z_slabs(64, w)
assert zs[0][0] == 0 and zs[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(zs, zs[1:]))
dist.destroy_process_group()
print("ok", r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_allgather_var_gloo_world2(tmp_path):
    """The halo exchange primitive of the z-slab octree build (sharding.TorchComm.allgather_var): ragged, empty and
    equal-length contributions over a real process group."""
    script = tmp_path / "ag.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {str(ROOT)!r})
from fpohm_b200 import sharding
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
c = sharding.TorchComm()
assert c.allreduce_max(3 + 4 * r, torch.device("cpu")) == 3 + 4 * (w - 1)
for lens in ([5, 2], [0, 7], [0, 0], [4, 4]):
    t = torch.arange(lens[r], dtype=torch.int64) + 1000 * r
    got = c.allgather_var(t)
    want = torch.cat([torch.arange(lens[k], dtype=torch.int64) + 1000 * k for k in range(w)])
    assert torch.equal(got, want), (lens, got)
dist.barrier()
print("OK", r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("OK") == 2


# ---- §8(f)-1 conforming_mesh: port vs golden, golden vs live reference --------------------------------------------------
def _conf_cases():
    g = dict(np.load(ROOT / "tests" / "golden" / "golden_conforming_v1.npz"))
    for name in ("a", "b", "c"):
        want = {k[len(name) + 5:]: v for k, v in g.items() if k.startswith(f"{name}_out_")}
        yield name, g[f"{name}_node_pos"], g[f"{name}_node_neigh"], g[f"{name}_hex"], g[f"{name}_grid"], want


def _dual_cases():
    g = dict(np.load(ROOT / "tests" / "golden" / "golden_conforming_v1.npz"))
    for name in ("a", "b", "c"):
        want = {k[len(name) + 6:]: v for k, v in g.items() if k.startswith(f"{name}_dual_")}
        yield name, g[f"{name}_node_pos"], g[f"{name}_node_neigh"], g[f"{name}_hex"], g[f"{name}_grid"], g[f"{name}_Vpos"], want


def _same_hybrid(got, want):
    for k, v in want.items():
        assert np.array_equal(np.asarray(got[k]).reshape(-1), np.asarray(v).reshape(-1)), k


def test_port_conforming_mesh_vs_golden(port):
    for name, npos, nn, H, gs, want in _conf_cases():
        _same_hybrid(port.conforming_mesh(npos, nn, H, gs), want)


def test_golden_conforming_matches_live_reference(ref):
    for name, npos, nn, H, gs, want in _conf_cases():
        Vp = npos.astype(np.float64)
        _same_hybrid(ref.conforming_mesh_tables(npos, nn, Vp, H, gs), want)


def test_port_dual_conforming_mesh_vs_golden(port):
    for name, npos, nn, H, gs, Vp, want in _dual_cases():
        _same_hybrid(port.dual_conforming_mesh(port.conforming_mesh(npos, nn, H, gs), Vp, H), want)


def test_golden_dual_matches_live_reference(ref):
    for name, npos, nn, H, gs, Vp, want in _dual_cases():
        _same_hybrid(ref.conforming_and_dual_tables(npos, nn, Vp, H, gs)[1], want)


# ---- §8(f)-2: clean_hex_mesh stages (golden_clean_v1.npz from the compiled reference, tests/golden/make_golden_clean.py) ----
def _clean_golden():
    return dict(np.load(ROOT / "tests" / "golden" / "golden_clean_v1.npz"))


def test_port_clean_stages_vs_golden(port):
    g = _clean_golden()
    for name in "abcd":
        V, H, flag = g[f"{name}_V"], g[f"{name}_hex"], g[f"{name}_flag"]
        assert np.array_equal(port.reorder_hex_mesh(V, g[f"{name}_hex_mirrored"]), g[f"{name}_reordered"]), name
        conn = port.hex_connectivity(H, len(V))
        t = port.tagging_uneven_element(conn, flag)
        assert np.array_equal(t, g[f"{name}_tagged"]), name
        vm, vr, hr, sh = port.re_indexing_connectivity(H, len(V), t)
        assert np.array_equal(vm, g[f"{name}_sub_V_map"]) and np.array_equal(vr, g[f"{name}_sub_V_map_reverse"])
        assert np.array_equal(hr, g[f"{name}_sub_H_map_reverse"]) and np.array_equal(sh, g[f"{name}_sub_hex"])
        n, _ = port.clean_non_manifold_ve(H, len(V), t)
        assert np.array_equal(n, g[f"{name}_manifold"]), name
        d, _ = port.drop_small_pieces(H, len(V), n)
        assert np.array_equal(d, g[f"{name}_dropped"]), name


def test_port_clean_hex_mesh_tail_vs_golden(port):
    """Whole clean_hex_mesh from the signed distances on: the port's stages chained the way ghm.cpp:1953-1981 chains them."""
    g = _clean_golden()
    for name in ("torus", "tori"):
        V, H, tV, tF = g[f"{name}_V"], g[f"{name}_hex"], g[f"{name}_tV"], g[f"{name}_tF"]
        P = (V[H].max(1) + V[H].min(1)) / 2
        t = port.PortTree(tV, tF)
        S = t.signed_distance(P)[0]
        flag = (S < 0).astype(np.uint8)
        conn = port.hex_connectivity(H, len(V))
        flag = port.tagging_uneven_element(conn, flag)
        flag, _ = port.clean_non_manifold_ve(H, len(V), flag)
        flag, _ = port.drop_small_pieces(H, len(V), flag)
        assert np.array_equal(flag, g[f"{name}_flag"]), name
        Fm, Vm = port.medial_surface_flags(conn, len(V), flag)
        assert np.array_equal(Fm, g[f"{name}_F_medial"]) and np.array_equal(Vm, g[f"{name}_V_medial"]), name


def test_golden_clean_matches_live_reference(ref):
    g = _clean_golden()
    for name in "ab":
        rc = ref.RefClean(g[f"{name}_V"], g[f"{name}_hex"])
        rc.set_flags(g[f"{name}_flag"])
        assert np.array_equal(rc.tagging(), g[f"{name}_tagged"])
        rc.reindex()
        assert np.array_equal(rc.non_manifold(), g[f"{name}_manifold"])
        assert np.array_equal(rc.drop_small(), g[f"{name}_dropped"])
    rc = ref.RefClean(g["torus_V"], g["torus_hex"])
    assert np.array_equal(rc.full(g["torus_tV"], g["torus_tF"]), g["torus_flag"])
    Fm, Vm = rc.medial()
    assert np.array_equal(Fm, g["torus_F_medial"]) and np.array_equal(Vm, g["torus_V_medial"])


def test_port_extract_surface_vs_golden(port):
    g = _clean_golden()
    for name in ("surf_cavity", "surf_carved"):
        for tri in (0, 1):
            got = port.extract_surface_conforming_mesh(g[f"{name}_V"], g[f"{name}_hex"], bool(tri))
            for k in ("V", "F_vs", "F_es", "E_vs", "E_boundary", "V_boundary", "V_map", "V_map_reverse", "F_map", "F_map_reverse"):
                assert np.array_equal(np.asarray(got[k]), g[f"{name}_{tri}_{k}"]), (name, tri, k)
            for k in ("E_nfs", "V_nvs", "V_nes", "V_nfs"):
                assert np.array_equal(got[k][0], g[f"{name}_{tri}_{k}_off"]) and np.array_equal(got[k][1], g[f"{name}_{tri}_{k}_val"]), (name, tri, k)


# ---- §8(f)-3: SLIM per-element stages (golden_slim_v1.npz from the compiled slim_m.cpp, tests/golden/make_golden_slim.py) ----
def test_port_slim_stages_vs_golden(port):
    g = dict(np.load(ROOT / "tests" / "golden" / "golden_slim_v1.npz"))
    J, areas, ef = g["J"], g["areas"], float(g["exp_factor"])
    for en in port.SLIM_ENERGIES:
        W, Ri = port.slim_weights_rotations(J, en, ef)
        fin = np.isfinite(g[f"{en}_W"]).all(1)
        assert np.array_equal(fin, np.isfinite(W).all(1)), en
        assert np.abs(W[fin] - g[f"{en}_W"][fin]).max() <= 1e-9 * np.abs(g[f"{en}_W"][fin]).max(), en      # LAPACK vs Eigen JacobiSVD: rounding only
        assert np.abs(Ri - g[f"{en}_Ri"]).max() <= 1e-9 * max(1.0, np.abs(g[f"{en}_Ri"]).max()), en
        e = port.slim_energy(J, areas, en, ef)
        want = float(g[f"{en}_energy"])
        assert (e == want) if not np.isfinite(want) else abs(e - want) <= 1e-10 * abs(want), en      # the exponential energy overflows to inf in both
    Ji = port.slim_jacobians(g["jac_off"], g["jac_col"], g["jac_vx"], g["jac_vy"], g["jac_vz"], g["jac_uv"])
    assert np.abs(Ji - g["jac_Ji"]).max() <= 1e-12 * np.abs(g["jac_Ji"]).max()


def test_port_slim_max_step_vs_golden(port):
    g = dict(np.load(ROOT / "tests" / "golden" / "golden_slim_v1.npz"))
    for k in range(3):
        m, r = port.slim_max_step(g["step_uv"], g["step_T"], g[f"step{k}_d"])
        want = g[f"step{k}_roots"]; fin = np.isfinite(want)
        assert np.array_equal(fin, np.isfinite(r))
        assert (np.abs(r[fin] - want[fin]) <= 1e-8 * np.abs(want[fin])).all() and abs(m - float(g[f"step{k}_max"])) <= 1e-9 * m


def test_port_slim_rhs_terms_vs_golden(port):
    g = dict(np.load(ROOT / "tests" / "golden" / "golden_slim_v1.npz"))
    assert np.array_equal(port.slim_rhs_terms(g["SYMMETRIC_DIRICHLET_W"], g["SYMMETRIC_DIRICHLET_Ri"]), g["rhs_terms"])      # same IEEE operations


def test_golden_slim_matches_live_reference(ref):
    g = dict(np.load(ROOT / "tests" / "golden" / "golden_slim_v1.npz"))
    W, Ri = ref.slim_weights_rotations(g["J"], "SYMMETRIC_DIRICHLET", float(g["exp_factor"]))
    assert np.array_equal(W, g["SYMMETRIC_DIRICHLET_W"]) and np.array_equal(Ri, g["SYMMETRIC_DIRICHLET_Ri"])
    assert ref.slim_energy(g["J"], g["areas"], "CONFORMAL", float(g["exp_factor"])) == float(g["CONFORMAL_energy"])


def test_io_wire_formats_byte_identical_to_reference(tmp_path):
    """§8(f)-4: .mesh / .vtk writers and the .fgraph reader against the reference's own h_io (io.cpp compiled into oracle/_ref):
    the files are compared BYTE for byte (doubles go through the same "%g", ids 1-based in .mesh), for every Mesh_type branch the
    writers have; .fgraph is read back through both and written by the product."""
    R = pytest.importorskip("oracle.ref_oracle")
    if not R.available():
        pytest.skip("oracle/_ref not built")
    import fpohm_b200 as fp
    pm = fp.procedural
    rng = np.random.default_rng(3)
    Vt, Ft = pm.torus(40, 24)
    Vt = Vt * np.array([1.0, 1e-3, 1e5]) + np.array([0.1, -2.5e-7, 123456.789])          # exponents on both sides of %g's switch to e-notation
    Vt[:4] = [[0.0, -0.0, 1.0], [1e-5, 123456.5, 1234567.0], [0.0001, 100000.0, 999999.5], [-1.5e300, 2.5e-300, 3.0]]
    Vh, H = pm.warped_hex_block(9, 0.3)
    vb = (rng.random(len(Vh)) < 0.3).astype(np.uint8)
    quads = H[:, [0, 1, 2, 3]]
    tets = H[:, [0, 1, 3, 4]]
    cases = [("Tri", Vt, Ft.astype(np.uint32)), ("HSur", Vt, Ft.astype(np.uint32)), ("Hex", Vh, H), ("Qua", Vh, quads)]
    for name, V, el in cases:
        a, b = tmp_path / f"{name}_ref.mesh", tmp_path / f"{name}_ours.mesh"
        R.io_write_mesh(a, V, fp.MESH_TYPES[name], el)
        fp.write_mesh(b, V, name, el)
        assert a.read_bytes() == b.read_bytes(), name
    for name, V, el in [("Tri", Vt, Ft.astype(np.uint32)), ("Qua", Vh, quads), ("Tet", Vh, tets), ("Hex", Vh, H)]:
        v_b = vb if len(V) == len(Vh) else (np.arange(len(V)) % 3 == 0).astype(np.uint8)
        a, b = tmp_path / f"{name}_ref.vtk", tmp_path / f"{name}_ours.vtk"
        R.io_write_vtk(a, V, fp.MESH_TYPES[name], el, v_b)
        fp.write_vtk(b, V, name, el, v_b)
        assert a.read_bytes() == b.read_bytes(), name
    # Hyb: polygons of 3..7 vertices in CSR
    sizes = rng.integers(3, 8, 500)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    poly = rng.integers(0, len(Vh), off[-1]).astype(np.uint32)
    a, b = tmp_path / "hyb_ref.vtk", tmp_path / "hyb_ours.vtk"
    R.io_write_vtk(a, Vh, 4, poly, vb, off)
    fp.write_vtk(b, Vh, "Hyb", poly, vb, off)
    assert a.read_bytes() == b.read_bytes()
    # .fgraph: the gear's crease edges, written by the product, read by both
    gV, gF, crease = pm.gear(teeth=12, n_radial=4, n_axial=6, n_arc=2)
    corners = np.unique(crease)[::7].astype(np.int32)
    fg = tmp_path / "gear.fgraph"
    fp.write_fgraph(fg, 30.5, 1, 0, corners, crease)
    mine, ref = fp.read_fgraph(fg), R.io_read_fgraph(fg)
    assert ref is not None
    for k in ("angle_threshold", "orphan_curve", "orphan_curve_single"):
        assert mine[k] == ref[k], k
    assert np.array_equal(mine["corners"], ref["corners"]) and np.array_equal(mine["pairs"], ref["pairs"]) and len(mine["pairs"]) == len(crease)
    assert np.array_equal(mine["pairs"], crease)
    with pytest.raises(fp.FpohmError):
        fp.read_fgraph(tmp_path / "missing.fgraph")
    assert R.io_read_fgraph(tmp_path / "missing.fgraph") is None

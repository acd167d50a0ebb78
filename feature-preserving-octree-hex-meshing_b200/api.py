"""ctypes binding of libfpohm.so + thin host classes that mirror the reference's interface names.

Nothing here computes: every function marshals numpy arrays (or raw device pointers for the `_dev`
variants) into the extern "C" entry points of include/fpohm.h.  If the shared library is missing the
import of this module raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

__all__ = [
    "FpohmError", "LIB_PATH", "lib", "device_count", "Context", "TriMesh", "OctreeParams", "Octree", "OctreeShard",
    "octree_grid_setup", "host_igl_tree", "host_igl_normals", "scaled_jacobian", "scaled_jacobian_dev", "points_inside_mesh", "HexConnectivity", "classify_hexes", "conforming_mesh", "conforming_mesh_tables", "conforming_and_dual", "conforming_and_dual_tables", "voxel_lattice", "VoxelGrid", "compute_sign_voxels", "voxel_sign_dev", "voxel_sign_slab_dev",
    "voxel_occupancy", "compute_sign_dexels", "polyline_project", "hausdorff", "hausdorff_outliers",
    "reorder_hexes", "tag_uneven_elements", "reindex_submesh", "clean_non_manifold", "drop_small_pieces", "medial_surface_flags", "clean_hex_mesh", "extract_surface",
    "MESH_TYPES", "write_mesh", "write_vtk", "read_fgraph", "write_fgraph",
    "SLIM_ENERGIES", "slim_jacobians", "slim_weights_rotations", "slim_energy", "slim_weights_rotations_dev", "slim_energy_dev", "slim_max_step", "slim_rhs_terms",
]

LIB_PATH = Path(__file__).resolve().parent / "libfpohm.so"
_lib = None


class FpohmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fpohm error {code}: {msg}")
        self.code = code


class _OctreeParams(C.Structure):
    _fields_ = [("grid_size", C.c_int32 * 3), ("origin", C.c_double * 3), ("mesh_transform", C.c_double * 3),
                ("voxel_size", C.c_double), ("stop_extent", C.c_int32), ("graded", C.c_int32), ("paired", C.c_int32),
                ("reserved", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                              "there is no CPU fallback")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.fpohm_last_error.restype = C.c_char_p
        _lib.fpohm_version.restype = C.c_char_p
    return _lib


def _chk(rc: int):
    if rc != 0:
        raise FpohmError(rc, lib().fpohm_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    return lib().fpohm_device_count()


class Context:
    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        _chk(lib().fpohm_ctx_create(C.c_int(device), C.byref(self.h)))
        self.device = device

    def sync(self):
        _chk(lib().fpohm_ctx_sync(self.h))

    def last_kernel_ms(self) -> float:
        ms = C.c_double()
        _chk(lib().fpohm_ctx_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def query_kernel_ms(self, last_n: int = 1) -> float:
        ms = C.c_double()
        _chk(lib().fpohm_ctx_query_kernel_ms(self.h, C.c_int32(last_n), C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_int64()
        _chk(lib().fpohm_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().fpohm_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TriMesh:
    """Device-resident triangle mesh (the reference's M_i / mf.tri) with lazily built trees."""

    def __init__(self, ctx: Context, V, F, cached: bool = False):
        """cached=True: keyed on the content of (V, F) in the context (fpohm_mesh_upload_cached) — a surface seen before comes back
        with its trees built; close() only drops the reference."""
        self.ctx = ctx
        self.V, self.F = _f64(V), _i32(F)
        self.h = C.c_void_p()
        up = lib().fpohm_mesh_upload_cached if cached else lib().fpohm_mesh_upload
        _chk(up(ctx.h, _p(self.V), C.c_int64(len(self.V)), _p(self.F), C.c_int64(len(self.F)), C.byref(self.h)))

    # build_aabb_tree, ghm.cpp:4231-4248
    def build_aabb_tree(self):
        _chk(lib().fpohm_mesh_build_query_tree(self.ctx.h, self.h))
        return self

    def normals(self):
        self.build_aabb_tree()
        nE = C.c_int64()
        _chk(lib().fpohm_mesh_num_edges(self.h, C.byref(nE)))
        FN = np.zeros((len(self.F), 3)); VN = np.zeros((len(self.V), 3)); EN = np.zeros((nE.value, 3))
        E = np.zeros((nE.value, 2), np.int32); EMAP = np.zeros(3 * len(self.F), np.int32)
        _chk(lib().fpohm_mesh_normals(self.h, _p(FN), _p(VN), _p(EN), _p(E), _p(EMAP)))
        return FN, VN, EN, E, EMAP

    def tree(self):
        self.build_aabb_tree()
        n = C.c_int64()
        _chk(lib().fpohm_mesh_tree_nodes(self.h, C.byref(n)))
        box = np.zeros((n.value, 6)); prim = np.zeros(n.value, np.int32); lr = np.zeros((n.value, 2), np.int32)
        _chk(lib().fpohm_mesh_tree_export(self.h, _p(box), _p(prim), _p(lr)))
        return box, prim, lr

    # igl::signed_distance_pseudonormal, igl/signed_distance.cpp:186-218
    def signed_distance_pseudonormal(self, P, want=("S", "I", "C", "N")):
        P = _f64(P).reshape(-1, 3); n = len(P)
        S = np.zeros(n) if "S" in want else None
        I = np.zeros(n, np.int32) if "I" in want else None
        Cc = np.zeros((n, 3)) if "C" in want else None
        N = np.zeros((n, 3)) if "N" in want else None
        _chk(lib().fpohm_signed_distance(self.ctx.h, self.h, _p(P), C.c_int64(n), _p(S), _p(I), _p(Cc), _p(N)))
        return S, I, Cc, N

    def signed_distance_dev(self, P_ptr: int, n: int, S_ptr=0, I_ptr=0, C_ptr=0, N_ptr=0, stream: int = 0):
        _chk(lib().fpohm_signed_distance_dev(self.ctx.h, self.h, C.c_void_p(P_ptr), C.c_int64(n), C.c_void_p(S_ptr),
                                             C.c_void_p(I_ptr), C.c_void_p(C_ptr), C.c_void_p(N_ptr), C.c_void_p(stream)))

    # igl::point_mesh_squared_distance
    def point_mesh_squared_distance(self, P):
        P = _f64(P).reshape(-1, 3); n = len(P)
        D = np.zeros(n); I = np.zeros(n, np.int32); Cc = np.zeros((n, 3))
        _chk(lib().fpohm_point_mesh_sqdist(self.ctx.h, self.h, _p(P), C.c_int64(n), _p(D), _p(I), _p(Cc)))
        return D, I, Cc

    def close(self):
        if self.h:
            lib().fpohm_mesh_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_igl_tree(V, F):
    """igl::AABB::init restatement, host only (no GPU): pre-order (box, primitive, left/right)."""
    V, F = _f64(V), _i32(F)
    n = C.c_int64(0)
    _chk(lib().fpohm_host_igl_tree(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), C.byref(n), None, None, None))
    box = np.zeros((n.value, 6)); prim = np.zeros(n.value, np.int32); lr = np.zeros((n.value, 2), np.int32)
    _chk(lib().fpohm_host_igl_tree(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), C.byref(n), _p(box), _p(prim), _p(lr)))
    return box, prim, lr


def host_igl_normals(V, F):
    V, F = _f64(V), _i32(F)
    n = C.c_int64(0)
    _chk(lib().fpohm_host_igl_normals(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), C.byref(n), None, None, None, None, None))
    FN = np.zeros((len(F), 3)); VN = np.zeros((len(V), 3)); EN = np.zeros((n.value, 3))
    E = np.zeros((n.value, 2), np.int32); EMAP = np.zeros(3 * len(F), np.int32)
    _chk(lib().fpohm_host_igl_normals(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), C.byref(n), _p(FN), _p(VN), _p(EN), _p(E), _p(EMAP)))
    return FN, VN, EN, E, EMAP


# points_inside_mesh, gf.cpp:4024-4048
def points_inside_mesh(ctx: Context, Ps, V, F):
    m = TriMesh(ctx, V, F, cached=True)      # the reference rebuilds its igl::AABB on every call (gf.cpp:4038)
    try:
        return m.signed_distance_pseudonormal(Ps, want=("S",))[0]
    finally:
        m.close()


class OctreeParams:
    def __init__(self, grid_size, origin, mesh_transform, voxel_size, stop_extent, graded=True, paired=True):
        self.c = _OctreeParams()
        for d in range(3):
            self.c.grid_size[d] = int(grid_size[d]); self.c.origin[d] = float(origin[d]); self.c.mesh_transform[d] = float(mesh_transform[d])
        self.c.voxel_size = float(voxel_size); self.c.stop_extent = int(stop_extent)
        self.c.graded = int(bool(graded)); self.c.paired = int(bool(paired))

    grid_size = property(lambda s: np.array(s.c.grid_size[:], np.int32))
    origin = property(lambda s: np.array(s.c.origin[:]))
    mesh_transform = property(lambda s: np.array(s.c.mesh_transform[:]))
    voxel_size = property(lambda s: s.c.voxel_size)


# ghm.cpp:463-493
def octree_grid_setup(V, num_voxels: int = 1 << 20) -> OctreeParams:
    V = _f64(V)
    p = OctreeParams([1, 1, 1], [0, 0, 0], [0, 0, 0], 1.0, 1)
    _chk(lib().fpohm_octree_grid_setup(_p(V), C.c_int64(len(V)), C.c_int32(num_voxels), C.byref(p.c)))
    return p


class Octree:
    """Mirror of OctreeGrid (octree.h:62-270) over a device-resident level-synchronous octree."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def build(cls, ctx: Context, mesh: TriMesh, params: OctreeParams):
        h = C.c_void_p()
        _chk(lib().fpohm_octree_build(ctx.h, mesh.h, C.byref(params.c), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_marks(cls, ctx: Context, grid_size, marks, graded=True, paired=True):
        gs = _i32(grid_size); m = _i32(marks).reshape(-1, 4)
        h = C.c_void_p()
        _chk(lib().fpohm_octree_build_from_marks(ctx.h, _p(gs), _p(m), C.c_int64(len(m)), C.c_int32(graded), C.c_int32(paired), C.byref(h)))
        return cls(ctx, h)

    def subdivide(self, mesh: TriMesh, stop_extent: int):
        _chk(lib().fpohm_octree_subdivide(self.h, mesh.h, C.c_int32(stop_extent)))

    def refine(self, mesh: TriMesh, cell_ids, stop_extent: int):
        c = _i32(cell_ids)
        _chk(lib().fpohm_octree_refine(self.h, mesh.h, _p(c), C.c_int64(len(c)), C.c_int32(stop_extent)))

    def sizes(self):
        nn, nc, nl = C.c_int64(), C.c_int64(), C.c_int64(); nr, md = C.c_int32(), C.c_int32()
        _chk(lib().fpohm_octree_sizes(self.h, C.byref(nn), C.byref(nc), C.byref(nl), C.byref(nr), C.byref(md)))
        return dict(nodes=nn.value, cells=nc.value, leaves=nl.value, roots=nr.value, max_depth=md.value)

    def export(self):
        s = self.sizes()
        node_pos = np.zeros((s["nodes"], 3), np.int32); node_neigh = np.zeros((s["nodes"], 6), np.int32)
        first_child = np.zeros(s["cells"], np.int32); corner = np.zeros((s["cells"], 8), np.int32)
        neigh = np.zeros((s["cells"], 6), np.int32)
        _chk(lib().fpohm_octree_export(self.h, _p(node_pos), _p(node_neigh), _p(first_child), _p(corner), _p(neigh)))
        return dict(node_pos=node_pos, node_neigh=node_neigh, first_child=first_child, corner=corner, neigh=neigh, **s)

    def hexes(self):
        s = self.sizes()
        Vp = np.zeros((s["nodes"], 3)); hexa = np.zeros((s["leaves"], 8), np.uint32); h2c = np.zeros(s["leaves"], np.int32)
        _chk(lib().fpohm_octree_hexes(self.h, _p(Vp), _p(hexa), _p(h2c)))
        return Vp, hexa, h2c

    def flags(self):
        f = C.c_int32()
        _chk(lib().fpohm_octree_check(self.h, C.byref(f)))
        return bool(f.value & 1), bool(f.value & 2)

    def cell_sign(self, mesh: TriMesh, origin, spacing: float):
        s = self.sizes(); o = _f64(origin)
        inside = np.zeros(s["cells"], np.float32)
        _chk(lib().fpohm_octree_cell_sign(self.h, mesh.h, _p(o), C.c_double(spacing), _p(inside)))
        return inside

    def close(self):
        if self.h:
            lib().fpohm_octree_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

class OctreeShard:
    """One rank of the z-slab sharded octree build (include/fpohm.h "z-slab sharded octree build").  The class only
    wraps the C-ABI steps; the collective between them is the caller's (sharding.build_octree_sharded)."""

    def __init__(self, ctx: Context, mesh: TriMesh, params: OctreeParams, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.h = C.c_void_p()
        _chk(lib().fpohm_octree_shard_create(ctx.h, mesh.h, C.byref(params.c), C.c_int32(rank), C.c_int32(world), C.byref(self.h)))

    def refine(self) -> int:
        lm = C.c_int32()
        _chk(lib().fpohm_octree_shard_refine(self.h, C.byref(lm)))
        return lm.value

    def info(self):
        rl = C.c_int32(); b = (C.c_int32 * (self.world + 1))(); t = C.c_int64()
        _chk(lib().fpohm_octree_shard_info(self.h, C.byref(rl), b, C.byref(t)))
        return dict(replicated_levels=rl.value, slab_bounds=list(b), owned_true_cells=t.value)

    def level_outgoing(self, global_max_level: int, level: int) -> int:
        n = C.c_int64()
        _chk(lib().fpohm_octree_shard_level_outgoing(self.h, C.c_int32(global_max_level), C.c_int32(level), C.byref(n)))
        return n.value

    def outgoing_copy(self, dst_ptr: int):
        _chk(lib().fpohm_octree_shard_outgoing_copy(self.h, C.c_void_p(dst_ptr)))

    def level_close(self, level: int, gathered_ptr: int, n: int) -> int:
        m = C.c_int64()
        _chk(lib().fpohm_octree_shard_level_close(self.h, C.c_int32(level), C.c_void_p(gathered_ptr), C.c_int64(n), C.byref(m)))
        return m.value

    def level_result(self, level: int, dst_ptr: int = 0) -> int:
        n = C.c_int64()
        _chk(lib().fpohm_octree_shard_level_result(self.h, C.c_int32(level), C.c_void_p(dst_ptr), C.byref(n)))
        return n.value

    def finish(self, ptrs, counts) -> "Octree":
        L = len(ptrs)
        pa = (C.c_void_p * max(L, 1))(*[C.c_void_p(int(x)) for x in ptrs]) if L else (C.c_void_p * 1)()
        ca = (C.c_int64 * max(L, 1))(*[int(c) for c in counts]) if L else (C.c_int64 * 1)()
        h = C.c_void_p()
        _chk(lib().fpohm_octree_shard_finish(self.h, pa, ca, C.byref(h)))
        return Octree(self.ctx, h)

    def close(self):
        if self.h:
            lib().fpohm_octree_shard_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass



# scaled_jacobian, gf.cpp:2309-2358 — returns (V_Js, H_Js, (min, ave, deviation), flipped) like Mesh_Quality
def scaled_jacobian(ctx: Context, V, hexa):
    V = _f64(V); hexa = np.ascontiguousarray(hexa, np.uint32); H = len(hexa)
    VJ = np.zeros(8 * H); HJ = np.zeros(H); mad = np.zeros(3); fl = C.c_int64()
    _chk(lib().fpohm_scaled_jacobian(ctx.h, _p(V), C.c_int64(len(V)), _p(hexa), C.c_int64(H), _p(VJ), _p(HJ), _p(mad), C.byref(fl)))
    return VJ, HJ, mad, fl.value


def scaled_jacobian_dev(ctx: Context, V_ptr, nV, hex_ptr, H, VJ_ptr, HJ_ptr, stats_ptr, flipped_ptr, stream=0):
    _chk(lib().fpohm_scaled_jacobian_dev(ctx.h, C.c_void_p(V_ptr), C.c_int64(nV), C.c_void_p(hex_ptr), C.c_int64(H),
                                         C.c_void_p(VJ_ptr), C.c_void_p(HJ_ptr), C.c_void_p(stats_ptr), C.c_void_p(flipped_ptr),
                                         C.c_void_p(stream)))


class HexConnectivity:
    """build_connectivity, Hex branch (gf.cpp:121-186,226-264)."""
    NAMES = ["F_nhs", "E_nfs", "E_nhs", "V_nvs", "V_nes", "V_nfs", "V_nhs"]

    def __init__(self, ctx: Context, hexa, nV: int, keep: bool = False):
        hexa = np.ascontiguousarray(hexa, np.uint32); H = len(hexa)
        h = C.c_void_p()
        self.h = None
        _chk(lib().fpohm_hex_connectivity(ctx.h, _p(hexa), C.c_int64(H), C.c_int64(nV), C.byref(h)))
        try:
            nF, nE = C.c_int64(), C.c_int64()
            _chk(lib().fpohm_conn_sizes(h, C.byref(nF), C.byref(nE)))
            nF, nE = nF.value, nE.value
            self.F_vs = np.zeros((nF, 4), np.uint32); self.F_es = np.zeros((nF, 4), np.uint32); self.F_boundary = np.zeros(nF, np.uint8)
            self.E_vs = np.zeros((nE, 2), np.uint32); self.E_boundary = np.zeros(nE, np.uint8); self.V_boundary = np.zeros(nV, np.uint8)
            self.H_fs = np.zeros((H, 6), np.uint32)
            _chk(lib().fpohm_conn_fixed(h, _p(self.F_vs), _p(self.F_es), _p(self.F_boundary), _p(self.E_vs), _p(self.E_boundary),
                                        _p(self.V_boundary), _p(self.H_fs)))
            sizes = [nF, nE, nE, nV, nV, nV, nV]
            for which, (nm, n) in enumerate(zip(self.NAMES, sizes)):
                tot = C.c_int64()
                _chk(lib().fpohm_conn_csr(h, C.c_int32(which), None, None, C.byref(tot)))
                off = np.zeros(n + 1, np.int64); val = np.zeros(tot.value, np.uint32)
                _chk(lib().fpohm_conn_csr(h, C.c_int32(which), _p(off), _p(val), C.byref(tot)))
                setattr(self, nm, (off, val))
        finally:
            if keep:
                self.h = h          # device tables stay alive for the cleaning stages (close() frees them)
            else:
                lib().fpohm_conn_free(h)

    def close(self):
        if self.h:
            lib().fpohm_conn_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def classify_hexes(ctx: Context, surface: "TriMesh", V, hexa):
    """clean_hex_mesh head (ghm.cpp:1937-1951): (signed_dis at the hex bbox centres, H_flag = inside)."""
    V = _f64(V); hexa = np.ascontiguousarray(hexa, np.uint32)
    S = np.zeros(len(hexa)); flag = np.zeros(len(hexa), np.uint8)
    _chk(lib().fpohm_classify_hexes(ctx.h, surface.h, _p(V), C.c_int64(len(V)), _p(hexa), C.c_int64(len(hexa)), _p(S), _p(flag)))
    return S, flag


# ---- clean_hex_mesh stages (ghm.cpp:1932-2124) ----------------------------------------------------------
def reorder_hexes(ctx: Context, V, hexa):
    """reorder_hex_mesh (gf.cpp:2199-2229): (reordered hex list, number of mirrored hexes)."""
    V = _f64(V); out = np.array(hexa, np.uint32, copy=True, order="C")
    n = C.c_int64()
    _chk(lib().fpohm_reorder_hexes(ctx.h, _p(V), C.c_int64(len(V)), _p(out), C.c_int64(len(out)), C.byref(n)))
    return out, n.value


def tag_uneven_elements(ctx: Context, conn: "HexConnectivity", H_flag):
    """tagging_uneven_element (ghm.cpp:1983-2005); conn = HexConnectivity(..., keep=True).  Returns (flags, sweeps)."""
    f = np.array(H_flag, np.uint8, copy=True); sweeps = C.c_int32()
    _chk(lib().fpohm_tag_uneven_elements(ctx.h, conn.h, _p(f), C.byref(sweeps)))
    return f, sweeps.value


def reindex_submesh(ctx: Context, hexa, nV: int, H_flag):
    """re_indexing_connectivity (gf.cpp:664-698): dict(V_map, V_map_reverse, H_map_reverse, hex)."""
    hexa = np.ascontiguousarray(hexa, np.uint32); f = np.ascontiguousarray(H_flag, np.uint8); H = len(hexa)
    V_map = np.zeros(nV, np.int32); V_rev = np.zeros(nV, np.int32); H_rev = np.zeros(H, np.int32); sub = np.zeros((H, 8), np.uint32)
    nv, nh = C.c_int64(), C.c_int64()
    _chk(lib().fpohm_reindex_submesh(ctx.h, _p(hexa), C.c_int64(H), C.c_int64(nV), _p(f), _p(V_map), _p(V_rev), C.byref(nv), _p(H_rev), C.byref(nh), _p(sub)))
    return dict(V_map=V_map, V_map_reverse=V_rev[:nv.value].copy(), H_map_reverse=H_rev[:nh.value].copy(), hex=sub[:nh.value].copy())


def clean_non_manifold(ctx: Context, hexa, nV: int, H_flag):
    """clean_non_manifold_ve (ghm.cpp:2006-2080): (flags, rounds)."""
    hexa = np.ascontiguousarray(hexa, np.uint32); f = np.array(H_flag, np.uint8, copy=True); r = C.c_int32()
    _chk(lib().fpohm_clean_non_manifold(ctx.h, _p(hexa), C.c_int64(len(hexa)), C.c_int64(nV), _p(f), C.byref(r)))
    return f, r.value


def drop_small_pieces(ctx: Context, hexa, nV: int, H_flag):
    """drop_small_pieces (ghm.cpp:2081-2124): (flags, number of pieces)."""
    hexa = np.ascontiguousarray(hexa, np.uint32); f = np.array(H_flag, np.uint8, copy=True); n = C.c_int64()
    _chk(lib().fpohm_drop_small_pieces(ctx.h, _p(hexa), C.c_int64(len(hexa)), C.c_int64(nV), _p(f), C.byref(n)))
    return f, n.value


def medial_surface_flags(ctx: Context, conn: "HexConnectivity", H_flag):
    """tail of clean_hex_mesh (ghm.cpp:1970-1981): (F_medial, V_medial)."""
    f = np.ascontiguousarray(H_flag, np.uint8)
    Fm = np.zeros(len(conn.F_vs), np.uint8); Vm = np.zeros(len(conn.V_boundary), np.uint8)
    _chk(lib().fpohm_medial_surface_flags(ctx.h, conn.h, _p(f), _p(Fm), _p(Vm)))
    return Fm, Vm


def clean_hex_mesh(ctx: Context, surface: "TriMesh", V, hexa, conn: "HexConnectivity | None" = None):
    """clean_hex_mesh (ghm.cpp:1932-1981, scaffold_type 1): dict(hex (reordered), signed_dis, H_flag, F_medial, V_medial, stats)."""
    V = _f64(V); hx = np.array(hexa, np.uint32, copy=True, order="C"); H = len(hx)
    S = np.zeros(H); flag = np.zeros(H, np.uint8); Vm = np.zeros(len(V), np.uint8)
    Fm = np.zeros(len(conn.F_vs), np.uint8) if conn is not None else None
    stats = (C.c_int64 * 6)()
    _chk(lib().fpohm_clean_hex_mesh(ctx.h, surface.h, _p(V), C.c_int64(len(V)), _p(hx), C.c_int64(H), conn.h if conn is not None else None,
                                    _p(S), _p(flag), _p(Fm) if Fm is not None else None, _p(Vm), stats))
    return dict(hex=hx, signed_dis=S, H_flag=flag, F_medial=Fm, V_medial=Vm, stats=list(stats))


def extract_surface(ctx: Context, conn: "HexConnectivity", V, as_triangles: bool = False):
    """extract_surface_conforming_mesh + orient_surface_mesh (gf.cpp:1021-1112) on the hex mesh behind conn (keep=True)."""
    V = _f64(V)
    h = C.c_void_p()
    _chk(lib().fpohm_extract_surface(ctx.h, conn.h, _p(V), C.c_int32(1 if as_triangles else 0), C.byref(h)))
    try:
        nV, nF, nE, lv = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        vn = C.c_int32()
        _chk(lib().fpohm_surface_sizes(h, C.byref(nV), C.byref(nF), C.byref(nE), C.byref(vn), C.byref(lv)))
        nV, nF, nE, vn = nV.value, nF.value, nE.value, vn.value
        out = dict(V=np.zeros((nV, 3)), F_vs=np.zeros((nF, vn), np.uint32), F_es=np.zeros((nF, vn), np.uint32), E_vs=np.zeros((nE, 2), np.uint32),
                   E_boundary=np.zeros(nE, np.uint8), V_boundary=np.zeros(nV, np.uint8), V_map=np.zeros(len(V), np.int32),
                   V_map_reverse=np.zeros(nV, np.int32), F_map=np.zeros(len(conn.F_vs), np.int32), F_map_reverse=np.zeros(nF, np.int32))
        _chk(lib().fpohm_surface_export(h, _p(out["V"]), _p(out["F_vs"]), _p(out["F_es"]), _p(out["E_vs"]), _p(out["E_boundary"]), _p(out["V_boundary"]),
                                        _p(out["V_map"]), _p(out["V_map_reverse"]), _p(out["F_map"]), _p(out["F_map_reverse"])))
        for which, (nm, n) in enumerate((("E_nfs", nE), ("V_nvs", nV), ("V_nes", nV), ("V_nfs", nV))):
            tot = C.c_int64()
            _chk(lib().fpohm_surface_csr(h, C.c_int32(which), None, None, C.byref(tot)))
            off = np.zeros(n + 1, np.int64); val = np.zeros(tot.value, np.uint32)
            _chk(lib().fpohm_surface_csr(h, C.c_int32(which), _p(off), _p(val), C.byref(tot)))
            out[nm] = (off, val)
        out["bfs_levels"] = lv.value
        return out
    finally:
        lib().fpohm_surface_free(h)


# ---- SLIM per-element stages, tet branch (slim_m.cpp:84-381, 861-913) ------------------------------------------
SLIM_ENERGIES = {"ARAP": 0, "LOG_ARAP": 1, "SYMMETRIC_DIRICHLET": 2, "CONFORMAL": 3, "EXP_CONFORMAL": 4, "EXP_SYMMETRIC_DIRICHLET": 5}


def slim_jacobians(ctx: Context, off, col, vx, vy, vz, uv):
    """compute_jacobians (slim_m.cpp:84-106): Dx, Dy, Dz as one CSR pattern with three value arrays; uv is nv x 3.  Returns Ji (n x 9)."""
    off = np.ascontiguousarray(off, np.int64); col = _i32(col); uv = _f64(uv); n = len(off) - 1
    Ji = np.zeros((n, 9))
    _chk(lib().fpohm_slim_jacobians(ctx.h, C.c_int64(n), C.c_int64(len(uv)), _p(off), _p(col), _p(_f64(vx)), _p(_f64(vy)), _p(_f64(vz)), _p(uv), _p(Ji)))
    return Ji


def slim_weights_rotations(ctx: Context, Ji, energy: str, exp_factor: float = 1.0):
    """update_weights_and_closest_rotations (slim_m.cpp:229-381): (W n x 9 = W_11..W_33, Ri n x 9 as s.Ri)."""
    Ji = _f64(Ji).reshape(-1, 9); n = len(Ji)
    W = np.zeros((n, 9)); Ri = np.zeros((n, 9))
    _chk(lib().fpohm_slim_weights_rotations(ctx.h, _p(Ji), C.c_int64(n), C.c_int32(SLIM_ENERGIES[energy]), C.c_double(exp_factor), _p(W), _p(Ri)))
    return W, Ri


def slim_energy(ctx: Context, Ji, areas, energy: str, exp_factor: float = 1.0) -> float:
    """compute_energy_with_jacobians (slim_m.cpp:861-913)."""
    Ji = _f64(Ji).reshape(-1, 9); a = _f64(areas); e = C.c_double()
    _chk(lib().fpohm_slim_energy(ctx.h, _p(Ji), C.c_int64(len(Ji)), _p(a), C.c_int32(SLIM_ENERGIES[energy]), C.c_double(exp_factor), C.byref(e)))
    return e.value


def slim_max_step(ctx: Context, uv, T, d):
    """compute_max_step_from_singularities (igl/flip_avoiding_line_search.cpp:273-299, tets): (max_step, per-tet roots)."""
    uv, d = _f64(uv), _f64(d); T = _i32(T); roots = np.zeros(len(T)); m = C.c_double()
    _chk(lib().fpohm_slim_max_step(ctx.h, _p(uv), C.c_int64(len(uv)), _p(T), C.c_int64(len(T)), _p(d), _p(roots), C.byref(m)))
    return m.value, roots


def slim_rhs_terms(ctx: Context, W, Ri):
    """per-element part of buildRhs (slim_m.cpp:1061-1083): f_rhs (9 n) from W (n x 9) and Ri (n x 9)."""
    W = _f64(W).reshape(-1, 9); Ri = _f64(Ri).reshape(-1, 9); f = np.zeros(9 * len(W))
    _chk(lib().fpohm_slim_rhs_terms(ctx.h, _p(W), _p(Ri), C.c_int64(len(W)), _p(f)))
    return f


def slim_weights_rotations_dev(ctx: Context, Ji_ptr: int, n: int, energy: str, exp_factor: float, W_ptr: int, Ri_ptr: int, stream: int = 0):
    _chk(lib().fpohm_slim_weights_rotations_dev(ctx.h, C.c_void_p(Ji_ptr), C.c_int64(n), C.c_int32(SLIM_ENERGIES[energy]), C.c_double(exp_factor),
                                                C.c_void_p(W_ptr), C.c_void_p(Ri_ptr), C.c_void_p(stream)))


def slim_energy_dev(ctx: Context, Ji_ptr: int, n: int, areas_ptr: int, energy: str, exp_factor: float, out_ptr: int, stream: int = 0):
    _chk(lib().fpohm_slim_energy_dev(ctx.h, C.c_void_p(Ji_ptr), C.c_int64(n), C.c_void_p(areas_ptr), C.c_int32(SLIM_ENERGIES[energy]), C.c_double(exp_factor),
                                     C.c_void_p(out_ptr), C.c_void_p(stream)))


def _hybrid_to_dict(hy):
    sizes = (C.c_int64 * 8)(); nrep = C.c_int64()
    _chk(lib().fpohm_hybrid_sizes(hy, sizes, C.byref(nrep)))
    nV, nF, nH, nE, fv, hf, hv, fn = [int(x) for x in sizes]
    out = dict(nV=nV, nF=nF, nH=nH, nE=nE, n_replaced=nrep.value,
               F_off=np.zeros(nF + 1, np.int64), F_vs=np.zeros(fv, np.uint32), F_es=np.zeros(fv, np.uint32), F_boundary=np.zeros(nF, np.uint8),
               E_vs=np.zeros((nE, 2), np.uint32), E_boundary=np.zeros(nE, np.uint8), V_boundary=np.zeros(nV, np.uint8),
               H_foff=np.zeros(nH + 1, np.int64), H_fs=np.zeros(hf, np.uint32), H_voff=np.zeros(nH + 1, np.int64), H_vs=np.zeros(hv, np.uint32),
               F_nhoff=np.zeros(nF + 1, np.int64), F_nhs=np.zeros(fn, np.uint32))
    _chk(lib().fpohm_hybrid_export(hy, _p(out["F_off"]), _p(out["F_vs"]), _p(out["F_es"]), _p(out["F_boundary"]), _p(out["E_vs"]),
                                   _p(out["E_boundary"]), _p(out["V_boundary"]), _p(out["H_foff"]), _p(out["H_fs"]), _p(out["H_voff"]),
                                   _p(out["H_vs"]), _p(out["F_nhoff"]), _p(out["F_nhs"])))
    return out


def conforming_mesh(ctx: Context, octree: "Octree", hexa=None, keep_timing: dict | None = None):
    """conforming_mesh (ghm.cpp:568-696) of the octree's hex mesh: polyhedral mesh as a dict of arrays (CSR for the
    variable-length relations).  `hexa` defaults to the octree's own hexes."""
    if hexa is None:
        hexa = octree.hexes()[1]
    hexa = np.ascontiguousarray(hexa, np.uint32)
    nV = octree.sizes()["nodes"]
    hc = C.c_void_p(); hy = C.c_void_p()
    _chk(lib().fpohm_hex_connectivity(ctx.h, _p(hexa), C.c_int64(len(hexa)), C.c_int64(nV), C.byref(hc)))
    try:
        if keep_timing is not None:
            keep_timing["connectivity_ms"] = ctx.last_kernel_ms()
        _chk(lib().fpohm_conforming_mesh(ctx.h, octree.h, hc, C.byref(hy)))
        if keep_timing is not None:
            keep_timing["conforming_ms"] = ctx.last_kernel_ms()
        return _hybrid_to_dict(hy)
    finally:
        if hy:
            lib().fpohm_hybrid_free(hy)
        lib().fpohm_conn_free(hc)


def conforming_and_dual(ctx: Context, octree: "Octree", keep_timing: dict | None = None):
    """conforming_mesh + dual_conforming_mesh (ghm.cpp:568-872) of the octree's hex mesh: (hybrid, dual) dicts; the dual
    carries "V" (cell centres), "h_type" (Element_Type per cell) and "census"."""
    Vp, hexa, _ = octree.hexes()
    hexa = np.ascontiguousarray(hexa, np.uint32); Vp = _f64(Vp)
    hc = C.c_void_p(); hy = C.c_void_p(); du = C.c_void_p()
    _chk(lib().fpohm_hex_connectivity(ctx.h, _p(hexa), C.c_int64(len(hexa)), C.c_int64(len(Vp)), C.byref(hc)))
    try:
        _chk(lib().fpohm_conforming_mesh(ctx.h, octree.h, hc, C.byref(hy)))
        if keep_timing is not None:
            keep_timing["conforming_ms"] = ctx.last_kernel_ms()
        _chk(lib().fpohm_dual_conforming_mesh(ctx.h, hy, _p(Vp), C.c_int64(len(Vp)), _p(hexa), C.c_int64(len(hexa)), C.byref(du)))
        if keep_timing is not None:
            keep_timing["dual_ms"] = ctx.last_kernel_ms()
        hyb = _hybrid_to_dict(hy); d = _hybrid_to_dict(du)
        d["V"] = np.zeros((d["nV"], 3)); d["h_type"] = np.zeros(d["nH"], np.int32); cen = (C.c_int64 * 7)()
        _chk(lib().fpohm_hybrid_dual_extra(du, _p(d["V"]), _p(d["h_type"]), cen))
        d["census"] = np.array(list(cen), np.int64)
        return hyb, d
    finally:
        if du:
            lib().fpohm_hybrid_free(du)
        if hy:
            lib().fpohm_hybrid_free(hy)
        lib().fpohm_conn_free(hc)


def conforming_and_dual_tables(ctx: Context, node_pos, node_neigh, Vpos, hexa, grid_size):
    """conforming_mesh + dual_conforming_mesh for an octree given as tables in any numbering (vertex i = node i)."""
    npos, nn, gs, Vp = _i32(node_pos), _i32(node_neigh), _i32(grid_size), _f64(Vpos)
    hexa = np.ascontiguousarray(hexa, np.uint32)
    hc = C.c_void_p(); hy = C.c_void_p(); du = C.c_void_p()
    _chk(lib().fpohm_hex_connectivity(ctx.h, _p(hexa), C.c_int64(len(hexa)), C.c_int64(len(npos)), C.byref(hc)))
    try:
        _chk(lib().fpohm_conforming_mesh_tables(ctx.h, _p(npos), _p(nn), C.c_int64(len(npos)), _p(gs), hc, C.byref(hy)))
        _chk(lib().fpohm_dual_conforming_mesh(ctx.h, hy, _p(Vp), C.c_int64(len(Vp)), _p(hexa), C.c_int64(len(hexa)), C.byref(du)))
        hyb = _hybrid_to_dict(hy); d = _hybrid_to_dict(du)
        d["V"] = np.zeros((d["nV"], 3)); d["h_type"] = np.zeros(d["nH"], np.int32); cen = (C.c_int64 * 7)()
        _chk(lib().fpohm_hybrid_dual_extra(du, _p(d["V"]), _p(d["h_type"]), cen))
        d["census"] = np.array(list(cen), np.int64)
        return hyb, d
    finally:
        if du:
            lib().fpohm_hybrid_free(du)
        if hy:
            lib().fpohm_hybrid_free(hy)
        lib().fpohm_conn_free(hc)


def conforming_mesh_tables(ctx: Context, node_pos, node_neigh, hexa, grid_size):
    """The same for an octree given as tables in any numbering (vertex i = node i)."""
    npos, nn, gs = _i32(node_pos), _i32(node_neigh), _i32(grid_size)
    hexa = np.ascontiguousarray(hexa, np.uint32)
    hc = C.c_void_p(); hy = C.c_void_p()
    _chk(lib().fpohm_hex_connectivity(ctx.h, _p(hexa), C.c_int64(len(hexa)), C.c_int64(len(npos)), C.byref(hc)))
    try:
        _chk(lib().fpohm_conforming_mesh_tables(ctx.h, _p(npos), _p(nn), C.c_int64(len(npos)), _p(gs), hc, C.byref(hy)))
        return _hybrid_to_dict(hy)
    finally:
        if hy:
            lib().fpohm_hybrid_free(hy)
        lib().fpohm_conn_free(hc)


class VoxelGrid:
    """VoxelGrid<num_t> (voxelization.h:41-91): origin/extent/spacing/padding -> dims + padded origin; x-fastest bytes."""

    def __init__(self, origin, extent, spacing: float, padding: int = 0):
        o, e = _f64(origin), _f64(extent)
        self.dims = np.zeros(3, np.int32); self.origin = np.zeros(3); self.spacing = float(spacing)
        _chk(lib().fpohm_voxel_grid_setup(_p(o), _p(e), C.c_double(spacing), C.c_int32(padding), _p(self.dims), _p(self.origin)))
        self.data = None

    def num_voxels(self):
        return int(self.dims[0]) * int(self.dims[1]) * int(self.dims[2])


# compute_sign(M, aabb, VoxelGrid&), voxelization.h:220-272
def compute_sign_voxels(ctx: Context, mesh: TriMesh, grid: VoxelGrid):
    out = np.zeros(grid.num_voxels(), np.uint8)
    _chk(lib().fpohm_voxel_sign(ctx.h, mesh.h, _p(grid.origin), C.c_double(grid.spacing), _p(grid.dims), _p(out)))
    grid.data = out.reshape(grid.dims[2], grid.dims[1], grid.dims[0])
    return grid.data


def voxel_sign_dev(ctx: Context, mesh: TriMesh, grid: VoxelGrid, out_ptr: int, stream: int = 0):
    _chk(lib().fpohm_voxel_sign_dev(ctx.h, mesh.h, _p(grid.origin), C.c_double(grid.spacing), _p(grid.dims), C.c_void_p(out_ptr), C.c_void_p(stream)))


def voxel_sign_slab_dev(ctx: Context, mesh: TriMesh, grid: VoxelGrid, z_begin: int, z_end: int, out_ptr: int, stream: int = 0):
    _chk(lib().fpohm_voxel_sign_slab_dev(ctx.h, mesh.h, _p(grid.origin), C.c_double(grid.spacing), _p(grid.dims), C.c_int32(z_begin),
                                         C.c_int32(z_end), C.c_void_p(out_ptr), C.c_void_p(stream)))


def voxel_occupancy(ctx: Context, mesh: TriMesh, grid: VoxelGrid):
    out = np.zeros(grid.num_voxels(), np.uint8)
    _chk(lib().fpohm_voxel_occupancy(ctx.h, mesh.h, _p(grid.origin), C.c_double(grid.spacing), _p(grid.dims), _p(out)))
    return out.reshape(grid.dims[2], grid.dims[1], grid.dims[0])


# compute_sign(M, aabb, DexelGrid&), voxelization.h:275-331 -> CSR (offsets x-fastest, values)
def compute_sign_dexels(ctx: Context, mesh: TriMesh, grid: VoxelGrid):
    d2 = np.ascontiguousarray(grid.dims[:2])
    tot = C.c_int64()
    off = np.zeros(int(d2[0]) * int(d2[1]) + 1, np.int64)
    _chk(lib().fpohm_dexel_sign(ctx.h, mesh.h, _p(grid.origin), C.c_double(grid.spacing), _p(d2), _p(off), None, C.byref(tot)))
    val = np.zeros(tot.value)
    _chk(lib().fpohm_dexel_sign(ctx.h, mesh.h, _p(grid.origin), C.c_double(grid.spacing), _p(d2), _p(off), _p(val), C.byref(tot)))
    return off, val


# LINE branch of dirty_graph_projection, ghm.cpp:3967-3994
def polyline_project(ctx: Context, Vc, curve_off, curve_vs, circle, P, curve_id):
    Vc, P = _f64(Vc), _f64(P).reshape(-1, 3)
    co = np.ascontiguousarray(curve_off, np.int64); cv = _i32(curve_vs); ci = np.ascontiguousarray(circle, np.uint8); cid = _i32(curve_id)
    n = len(P)
    oL = np.zeros((n, 3)); aL = np.zeros((n, 3))
    _chk(lib().fpohm_polyline_project(ctx.h, _p(Vc), C.c_int64(len(Vc)), _p(co), _p(cv), _p(ci), C.c_int64(len(ci)), _p(P), _p(cid),
                                      C.c_int64(n), _p(oL), _p(aL)))
    return oL, aL


# metro compute(...), metro_hausdorff.cpp:12,196,358
def hausdorff(ctx: Context, A: TriMesh, B: TriMesh, extra_face_samples: int = 0):
    out = np.zeros(7); ns = np.zeros(2, np.int64)
    _chk(lib().fpohm_hausdorff(ctx.h, A.h, B.h, C.c_int64(extra_face_samples), _p(out), _p(ns)))
    d = dict(diag=out[0], max_ab=out[1], max_ba=out[2], mean_ab=out[3], mean_ba=out[4], rms_ab=out[5], rms_ba=out[6],
             n_ab=int(ns[0]), n_ba=int(ns[1]))
    d["max"] = max(out[1], out[2]); d["mean"] = max(out[3], out[4])
    d["ratio"] = float(np.float32(d["max"])) / out[0]        # `(float)mesh_dist_max / bbox.Diag()`, metro_hausdorff.cpp:186
    return d


# hausdorff_dis(mesh0, mesh1, outlierVs, thr), gf.cpp:3590-3628
def hausdorff_outliers(ctx: Context, A: TriMesh, B: TriMesh, dis_threshold: float):
    out = np.zeros(len(B.V), np.int32); n = C.c_int64()
    _chk(lib().fpohm_hausdorff_outliers(ctx.h, A.h, B.h, C.c_double(dis_threshold), _p(out), C.byref(n)))
    return out[:n.value].copy()


def voxel_lattice(ctx: Context, bb_min, bb_max, num_voxels: int):
    mn, mx = _f64(bb_min), _f64(bb_max)
    dim = np.zeros(3, np.int32)
    _chk(lib().fpohm_voxel_lattice_dims(_p(mn), _p(mx), C.c_int32(num_voxels), _p(dim)))
    nv = int(dim[0]) * int(dim[1]) * int(dim[2]); nh = int(dim[0] - 1) * int(dim[1] - 1) * int(dim[2] - 1)
    Vp = np.zeros((nv, 3)); hexa = np.zeros((nh, 8), np.uint32)
    _chk(lib().fpohm_voxel_lattice(ctx.h, _p(mn), _p(mx), C.c_int32(num_voxels), _p(Vp), _p(hexa)))
    return Vp, hexa, dim


# ---- wire formats: h_io::write_hybrid_mesh_MESH / _VTK, read / write_feature_Graph_FGRAPH (io.cpp) ----------------
MESH_TYPES = {"Tri": 0, "Qua": 1, "HSur": 2, "Tet": 3, "Hyb": 4, "Hex": 5}


def write_mesh(path, V, mesh_type: str, elems):
    V = _f64(V); el = np.ascontiguousarray(elems, np.uint32)
    _chk(lib().fpohm_io_write_mesh(str(path).encode(), _p(V), C.c_int64(len(V)), C.c_int32(MESH_TYPES[mesh_type]), _p(el), C.c_int64(len(el))))


def write_vtk(path, V, mesh_type: str, elems, V_boundary=None, elem_off=None):
    V = _f64(V)
    vb = np.ascontiguousarray(V_boundary, np.uint8) if V_boundary is not None else np.zeros(len(V), np.uint8)
    if mesh_type == "Hyb":
        off = np.ascontiguousarray(elem_off, np.int64); el = np.ascontiguousarray(elems, np.uint32).reshape(-1)
        _chk(lib().fpohm_io_write_vtk(str(path).encode(), _p(V), C.c_int64(len(V)), C.c_int32(4), _p(off), _p(el), C.c_int64(len(off) - 1), C.c_int32(0),
                                      _p(vb), C.c_int64(len(vb))))
    else:
        el = np.ascontiguousarray(elems, np.uint32)
        _chk(lib().fpohm_io_write_vtk(str(path).encode(), _p(V), C.c_int64(len(V)), C.c_int32(MESH_TYPES[mesh_type]), None, _p(el), C.c_int64(len(el)),
                                      C.c_int32(el.shape[1]), _p(vb), C.c_int64(len(vb))))


def read_fgraph(path):
    ang = C.c_double(); oc = C.c_int32(); ocs = C.c_int32(); nc = C.c_int64(0); npairs = C.c_int64(0)
    _chk(lib().fpohm_io_read_fgraph(str(path).encode(), C.byref(ang), C.byref(oc), C.byref(ocs), None, C.byref(nc), None, C.byref(npairs)))
    corners = np.zeros(max(nc.value, 1), np.int32); pairs = np.zeros((max(npairs.value, 1), 2), np.int32)
    _chk(lib().fpohm_io_read_fgraph(str(path).encode(), C.byref(ang), C.byref(oc), C.byref(ocs), _p(corners), C.byref(nc), _p(pairs), C.byref(npairs)))
    return dict(angle_threshold=ang.value, orphan_curve=oc.value, orphan_curve_single=ocs.value, corners=corners[:nc.value].copy(), pairs=pairs[:npairs.value].copy())


def write_fgraph(path, angle_threshold, orphan_curve, orphan_curve_single, corners, pairs):
    c = np.ascontiguousarray(corners, np.int32); p = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    _chk(lib().fpohm_io_write_fgraph(str(path).encode(), C.c_double(angle_threshold), C.c_int32(orphan_curve), C.c_int32(orphan_curve_single),
                                     _p(c), C.c_int64(len(c)), _p(p), C.c_int64(len(p))))

"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libfpohm_ref.so — the reference's own
sources (octree.cpp, voxelization.cpp, global_functions.cpp, metro_hausdorff.cpp + vendored geogram /
libigl / VCG) compiled in place by oracle/ref/Makefile, behind the extern "C" driver oracle/ref/ref_driver.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libfpohm_ref.so"

_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} missing: run `make -C oracle/ref` where /root/reference exists")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ref_octree_build.restype = C.c_void_p
        _lib.ref_octree_random.restype = C.c_void_p
        _lib.ref_octree_from_marks.restype = C.c_void_p
        _lib.ref_compute_octree.restype = C.c_void_p
        _lib.ref_dexel_sign.restype = C.c_void_p
        _lib.ref_tree_build.restype = C.c_void_p
        _lib.ref_hex_connectivity.restype = C.c_void_p
        _lib.ref_tree_num_edges.restype = C.c_int64
        _lib.ref_tree_num_nodes.restype = C.c_int64
        _lib.ref_conn_csr.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_cores() -> int:
    return lib().ref_num_cores()


def octree_grid_setup(V, F, num_voxels: int):
    V, F = _f64(V), _i32(F)
    gs = np.zeros(3, np.int32); o = np.zeros(3); mt = np.zeros(3); vs = C.c_double()
    lib().ref_octree_grid_setup(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), C.c_int(num_voxels),
                                _p(gs), _p(o), _p(mt), C.byref(vs))
    return gs, o, mt, vs.value


class RefOctree:
    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    @classmethod
    def build(cls, V, F, grid_size, origin, mesh_transform, voxel_size, stop_extent, graded=True, paired=True):
        V, F = _f64(V), _i32(F)
        gs, o, mt = _i32(grid_size), _f64(origin), _f64(mesh_transform)
        h = lib().ref_octree_build(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), _p(gs), _p(o), _p(mt),
                                   C.c_double(voxel_size), C.c_int(stop_extent), C.c_int(graded), C.c_int(paired))
        return cls(h)

    @classmethod
    def random(cls, grid_size, graded=True, paired=True):
        gs = _i32(grid_size)
        return cls(lib().ref_octree_random(_p(gs), C.c_int(graded), C.c_int(paired)))

    @classmethod
    def from_marks(cls, grid_size, marks, graded=True, paired=True):
        gs, m = _i32(grid_size), _i32(marks).reshape(-1, 4)
        return cls(lib().ref_octree_from_marks(_p(gs), _p(m), C.c_int64(len(m)), C.c_int(graded), C.c_int(paired)))

    def subdivide(self, stop_extent):
        """a further OctreeGrid::subdivide pass over this tree (ghm.cpp:495-500,523-524)"""
        lib().ref_octree_subdivide(self.h, C.c_int(stop_extent))

    def refine(self, cells, stop_extent):
        c = _i32(cells)
        lib().ref_octree_refine(self.h, _p(c), C.c_int64(len(c)), C.c_int(stop_extent))

    def sizes(self):
        nn, nc, nl = C.c_int64(), C.c_int64(), C.c_int64(); nr, md = C.c_int32(), C.c_int32()
        lib().ref_octree_sizes(self.h, C.byref(nn), C.byref(nc), C.byref(nl), C.byref(nr), C.byref(md))
        return dict(nodes=nn.value, cells=nc.value, leaves=nl.value, roots=nr.value, max_depth=md.value)

    def export(self):
        s = self.sizes()
        node_pos = np.zeros((s["nodes"], 3), np.int32); node_neigh = np.zeros((s["nodes"], 6), np.int32)
        first_child = np.zeros(s["cells"], np.int32); corner = np.zeros((s["cells"], 8), np.int32)
        neigh = np.zeros((s["cells"], 6), np.int32)
        lib().ref_octree_export(self.h, _p(node_pos), _p(node_neigh), _p(first_child), _p(corner), _p(neigh))
        return dict(node_pos=node_pos, node_neigh=node_neigh, first_child=first_child, corner=corner, neigh=neigh, **s)

    def flags(self):
        f = lib().ref_octree_flags(self.h)
        return bool(f & 1), bool(f & 2)

    def cell_sign(self, origin, spacing):
        s = self.sizes(); o = _f64(origin)
        inside = np.zeros(s["cells"], np.float32)
        lib().ref_octree_cell_sign(self.h, _p(o), C.c_double(spacing), _p(inside))
        return inside

    def hexes(self):
        s = self.sizes()
        Vp = np.zeros((s["nodes"], 3)); hexa = np.zeros((s["leaves"], 8), np.uint32); h2c = np.zeros(s["leaves"], np.int32)
        lib().ref_octree_hexes(self.h, _p(Vp), _p(hexa), _p(h2c))
        return Vp, hexa, h2c

    def __del__(self):
        if self.h and _lib is not None:
            _lib.ref_octree_free(self.h)
            self.h = None


def compute_octree(V, F, min_corner, extent, spacing, padding=0, graded=True, paired=True):
    V, F = _f64(V), _i32(F); mc, ex = _f64(min_corner), _f64(extent)
    nv, nh = C.c_int64(), C.c_int64()
    h = C.c_void_p(lib().ref_compute_octree(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), _p(mc), _p(ex),
                                            C.c_double(spacing), C.c_int(padding), C.c_int(graded), C.c_int(paired),
                                            C.byref(nv), C.byref(nh)))
    Vp = np.zeros((nv.value, 3)); hexa = np.zeros((nh.value, 8), np.uint32); inside = np.zeros(nh.value, np.float32)
    lib().ref_compute_octree_export(h, _p(Vp), _p(hexa), _p(inside))
    lib().ref_compute_octree_free(h)
    return Vp, hexa, inside


def voxel_sign(V, F, origin, extent, spacing, padding=0):
    V, F = _f64(V), _i32(F); o, ex = _f64(origin), _f64(extent)
    dims = np.zeros(3, np.int32)
    lib().ref_voxel_dims(_p(ex), C.c_double(spacing), C.c_int(padding), _p(dims))
    out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.uint8)
    lib().ref_voxel_sign(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), _p(o), _p(ex), C.c_double(spacing),
                         C.c_int(padding), _p(out))
    return out.reshape(dims[2], dims[1], dims[0]), dims


def dexel_sign(V, F, origin, extent, spacing, padding=0):
    V, F = _f64(V), _i32(F); o, ex = _f64(origin), _f64(extent)
    dims = np.zeros(2, np.int32); tot = C.c_int64()
    h = C.c_void_p(lib().ref_dexel_sign(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), _p(o), _p(ex),
                                        C.c_double(spacing), C.c_int(padding), _p(dims), C.byref(tot)))
    off = np.zeros(int(dims[0]) * int(dims[1]) + 1, np.int64); val = np.zeros(tot.value)
    lib().ref_dexel_export(h, _p(off), _p(val))
    lib().ref_dexel_free(h)
    return off, val, dims


class RefTree:
    def __init__(self, V, F):
        self.V, self.F = _f64(V), _i32(F)
        self.h = C.c_void_p(lib().ref_tree_build(_p(self.V), C.c_int64(len(self.V)), _p(self.F), C.c_int64(len(self.F))))

    def normals(self):
        nE = lib().ref_tree_num_edges(self.h)
        FN = np.zeros((len(self.F), 3)); VN = np.zeros((len(self.V), 3)); EN = np.zeros((nE, 3))
        E = np.zeros((nE, 2), np.int32); EMAP = np.zeros(3 * len(self.F), np.int32)
        lib().ref_tree_normals(self.h, _p(FN), _p(VN), _p(EN), _p(E), _p(EMAP))
        return FN, VN, EN, E, EMAP

    def flatten(self):
        n = lib().ref_tree_num_nodes(self.h)
        box = np.zeros((n, 6)); prim = np.zeros(n, np.int32); lr = np.zeros((n, 2), np.int32)
        lib().ref_tree_flatten(self.h, _p(box), _p(prim), _p(lr))
        return box, prim, lr

    def signed_distance(self, P):
        P = _f64(P); n = len(P)
        S = np.zeros(n); I = np.zeros(n, np.int32); Cc = np.zeros((n, 3)); N = np.zeros((n, 3))
        lib().ref_signed_distance(self.h, _p(P), C.c_int64(n), _p(S), _p(I), _p(Cc), _p(N))
        return S, I, Cc, N

    def __del__(self):
        if self.h and _lib is not None:
            _lib.ref_tree_free(self.h)
            self.h = None


def point_mesh_sqdist(V, F, P):
    V, F, P = _f64(V), _i32(F), _f64(P); n = len(P)
    D = np.zeros(n); I = np.zeros(n, np.int32); Cc = np.zeros((n, 3))
    lib().ref_point_mesh_sqdist(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), _p(P), C.c_int64(n), _p(D), _p(I), _p(Cc))
    return D, I, Cc


def points_inside_mesh(V, F, P):
    V, F, P = _f64(V), _i32(F), _f64(P); n = len(P)
    S = np.zeros(n)
    lib().ref_points_inside_mesh(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), _p(P), C.c_int64(n), _p(S))
    return S


def polyline_project(V, curve_off, curve_vs, circle, P, curve_id):
    V, P = _f64(V), _f64(P)
    co = np.ascontiguousarray(curve_off, np.int64); cv = _i32(curve_vs); ci = np.ascontiguousarray(circle, np.uint8)
    cid = _i32(curve_id); n = len(P)
    oL = np.zeros((n, 3)); aL = np.zeros((n, 3))
    lib().ref_polyline_project(_p(V), C.c_int64(len(V)), _p(co), _p(cv), _p(ci), _p(P), _p(cid), C.c_int64(n), _p(oL), _p(aL))
    return oL, aL


def scaled_jacobian(V, hexa):
    V = _f64(V); hexa = np.ascontiguousarray(hexa, np.uint32); H = len(hexa)
    VJ = np.zeros(8 * H); HJ = np.zeros(H); mad = np.zeros(3); fl = C.c_int64()
    lib().ref_scaled_jacobian(_p(V), C.c_int64(len(V)), _p(hexa), C.c_int64(H), _p(VJ), _p(HJ), _p(mad), C.byref(fl))
    return VJ, HJ, mad, fl.value


def hex_connectivity(hexa, nV):
    hexa = np.ascontiguousarray(hexa, np.uint32); H = len(hexa)
    nF, nE = C.c_int64(), C.c_int64()
    h = C.c_void_p(lib().ref_hex_connectivity(_p(hexa), C.c_int64(H), C.c_int64(nV), C.byref(nF), C.byref(nE)))
    nF, nE = nF.value, nE.value
    out = dict(F_vs=np.zeros((nF, 4), np.uint32), F_es=np.zeros((nF, 4), np.uint32), F_boundary=np.zeros(nF, np.uint8),
               E_vs=np.zeros((nE, 2), np.uint32), E_boundary=np.zeros(nE, np.uint8), V_boundary=np.zeros(nV, np.uint8),
               H_fs=np.zeros((H, 6), np.uint32))
    lib().ref_conn_fixed(h, _p(out["F_vs"]), _p(out["F_es"]), _p(out["F_boundary"]), _p(out["E_vs"]), _p(out["E_boundary"]),
                         _p(out["V_boundary"]), _p(out["H_fs"]))
    names = ["F_nhs", "E_nfs", "E_nhs", "V_nvs", "V_nes", "V_nfs", "V_nhs"]
    sizes = [nF, nE, nE, nV, nV, nV, nV]
    for which, (nm, n) in enumerate(zip(names, sizes)):
        tot = lib().ref_conn_csr(h, C.c_int(which), None, None)
        off = np.zeros(n + 1, np.int64); val = np.zeros(tot, np.uint32)
        lib().ref_conn_csr(h, C.c_int(which), _p(off), _p(val))
        out[nm] = (off, val)
    lib().ref_conn_free(h)
    return out


def hausdorff(VA, FA, VB, FB):
    VA, FA, VB, FB = _f64(VA), _i32(FA), _f64(VB), _i32(FB)
    out = np.zeros(3)
    lib().ref_hausdorff(_p(VA), C.c_int64(len(VA)), _p(FA), C.c_int64(len(FA)), _p(VB), C.c_int64(len(VB)), _p(FB),
                        C.c_int64(len(FB)), _p(out))
    return dict(diag=out[0], max=out[1], mean=out[2])


def hausdorff_ratio(VA, FA, VB, FB, thr):
    VA, FA, VB, FB = _f64(VA), _i32(FA), _f64(VB), _i32(FB)
    r = C.c_double()
    ok = lib().ref_hausdorff_ratio(_p(VA), C.c_int64(len(VA)), _p(FA), C.c_int64(len(FA)), _p(VB), C.c_int64(len(VB)),
                                   _p(FB), C.c_int64(len(FB)), C.c_double(thr), C.byref(r))
    return bool(ok), r.value


def hausdorff_face_sampled(VA, FA, VB, FB, n_target_ab, n_target_ba):
    """vcg::Sampling with VERTEX | FACE | SIMILAR sampling (extern/vcg/sampling.h:513-602) and the given total sample targets."""
    VA, FA, VB, FB = _f64(VA), _i32(FA), _f64(VB), _i32(FB)
    out = np.zeros(5); ns = np.zeros(2, np.uint64)
    lib().ref_hausdorff_face_sampled(_p(VA), C.c_int64(len(VA)), _p(FA), C.c_int64(len(FA)), _p(VB), C.c_int64(len(VB)), _p(FB),
                                     C.c_int64(len(FB)), C.c_uint64(n_target_ab), C.c_uint64(n_target_ba), _p(out), _p(ns))
    return dict(diag=out[0], max_ab=out[1], max_ba=out[2], mean_ab=out[3], mean_ba=out[4], n_ab=int(ns[0]), n_ba=int(ns[1]))


def hausdorff_dis_outliers(VA, FA, VB, FB, thr):
    """hausdorff_dis(mesh0, mesh1, outlierVs, thr), global_functions.cpp:3590-3628 — the compiled function; vertex ids of mesh1
    in the reference's own push order."""
    VA, FA, VB, FB = _f64(VA), _i32(FA), _f64(VB), _i32(FB)
    f = lib().ref_hausdorff_dis_outliers
    f.restype = C.c_int64
    out = np.zeros(max(len(VB), 1), np.int32)
    n = f(_p(VA), C.c_int64(len(VA)), _p(FA), C.c_int64(len(FA)), _p(VB), C.c_int64(len(VB)), _p(FB), C.c_int64(len(FB)),
          C.c_double(thr), _p(out), C.c_int64(len(out)))
    return out[:n].copy()


def voxel_meshing(V, F, num_voxels):
    """grid_hex_meshing_bijective::voxel_meshing, grid_hex_meshing.cpp:215-296 — the compiled member function on a GEO::Mesh
    built from (V, F).  Returns (lattice vertex positions [nV,3], hexes [nH,8] uint32)."""
    V, F = _f64(V), _i32(F)
    f = lib().ref_voxel_meshing
    f.restype = C.c_void_p
    sizes = np.zeros(2, np.int64)
    h = C.c_void_p(f(_p(V), C.c_int64(len(V)), _p(F), C.c_int64(len(F)), C.c_int(num_voxels), _p(sizes)))
    Vp = np.zeros((int(sizes[0]), 3)); H = np.zeros((int(sizes[1]), 8), np.uint32)
    lib().ref_voxel_meshing_export(h, _p(Vp), _p(H))
    lib().ref_voxel_meshing_free(h)
    return Vp, H


# ---- §8(f)-1: conforming_mesh (grid_hex_meshing.cpp:568-696) -----------------------------------------------------------
def _export_hybrid(h, sizes):
    nV, nF, nH, nE, fv, hf, hv, fn = [int(x) for x in sizes]
    out = dict(nV=nV, nF=nF, nH=nH, nE=nE,
               F_off=np.zeros(nF + 1, np.int64), F_vs=np.zeros(fv, np.uint32), F_es=np.zeros(fv, np.uint32), F_boundary=np.zeros(nF, np.uint8),
               E_vs=np.zeros((nE, 2), np.uint32), E_boundary=np.zeros(nE, np.uint8), V_boundary=np.zeros(nV, np.uint8),
               H_foff=np.zeros(nH + 1, np.int64), H_fs=np.zeros(hf, np.uint32), H_voff=np.zeros(nH + 1, np.int64), H_vs=np.zeros(hv, np.uint32),
               F_nhoff=np.zeros(nF + 1, np.int64), F_nhs=np.zeros(fn, np.uint32))
    lib().ref_hybrid_export(h, _p(out["F_off"]), _p(out["F_vs"]), _p(out["F_es"]), _p(out["F_boundary"]), _p(out["E_vs"]), _p(out["E_boundary"]),
                            _p(out["V_boundary"]), _p(out["H_foff"]), _p(out["H_fs"]), _p(out["H_voff"]), _p(out["H_vs"]), _p(out["F_nhoff"]), _p(out["F_nhs"]))
    return out


def _alloc_hybrid(sizes):
    nV, nF, nH, nE, fv, hf, hv, fn = [int(x) for x in sizes]
    return dict(nV=nV, nF=nF, nH=nH, nE=nE,
                F_off=np.zeros(nF + 1, np.int64), F_vs=np.zeros(fv, np.uint32), F_es=np.zeros(fv, np.uint32), F_boundary=np.zeros(nF, np.uint8),
                E_vs=np.zeros((nE, 2), np.uint32), E_boundary=np.zeros(nE, np.uint8), V_boundary=np.zeros(nV, np.uint8),
                H_foff=np.zeros(nH + 1, np.int64), H_fs=np.zeros(hf, np.uint32), H_voff=np.zeros(nH + 1, np.int64), H_vs=np.zeros(hv, np.uint32),
                F_nhoff=np.zeros(nF + 1, np.int64), F_nhs=np.zeros(fn, np.uint32))


def conforming_and_dual_tables(node_pos, node_neigh, Vpos, hexa, grid_size):
    """conforming_mesh followed by dual_conforming_mesh (ghm.cpp:697-872): (hybrid, dual) with dual["V"], dual["h_type"]."""
    npos, nn, Vp, hx, gs = _i32(node_pos), _i32(node_neigh), _f64(Vpos), np.ascontiguousarray(hexa, np.uint32), _i32(grid_size)
    sizes = (C.c_int64 * 8)()
    lib().ref_conforming_mesh_tables.restype = C.c_void_p
    h = C.c_void_p(lib().ref_conforming_mesh_tables(_p(npos), _p(nn), C.c_int64(len(npos)), _p(Vp), _p(hx), C.c_int64(len(hx)), _p(gs), sizes))
    hyb = _export_hybrid(h, list(sizes))
    lib().ref_dual_conforming_mesh(h, sizes)
    d = _alloc_hybrid(list(sizes))
    d["V"] = np.zeros((d["nV"], 3)); d["h_type"] = np.zeros(d["nH"], np.int32)
    lib().ref_dual_export(h, _p(d["V"]), _p(d["h_type"]), _p(d["F_off"]), _p(d["F_vs"]), _p(d["F_es"]), _p(d["F_boundary"]), _p(d["E_vs"]), _p(d["E_boundary"]),
                          _p(d["V_boundary"]), _p(d["H_foff"]), _p(d["H_fs"]), _p(d["H_voff"]), _p(d["H_vs"]), _p(d["F_nhoff"]), _p(d["F_nhs"]))
    lib().ref_hybrid_free(h)
    return hyb, d


def conforming_mesh_tables(node_pos, node_neigh, Vpos, hexa, grid_size):
    """The reference's conforming_mesh on an octree given as tables (vertex i = node i), any numbering."""
    npos, nn, Vp, hx, gs = _i32(node_pos), _i32(node_neigh), _f64(Vpos), np.ascontiguousarray(hexa, np.uint32), _i32(grid_size)
    sizes = (C.c_int64 * 8)()
    lib().ref_conforming_mesh_tables.restype = C.c_void_p
    h = C.c_void_p(lib().ref_conforming_mesh_tables(_p(npos), _p(nn), C.c_int64(len(npos)), _p(Vp), _p(hx), C.c_int64(len(hx)), _p(gs), sizes))
    out = _export_hybrid(h, list(sizes))
    lib().ref_hybrid_free(h)
    return out


# ---- §8(f)-2: clean_hex_mesh and its stages (grid_hex_meshing.cpp:1932-2126, global_functions.cpp:664-698,2199-2229) -----
class RefClean:
    """The reference's own clean_hex_mesh stages on a hex mesh (Vpos, hexa); build_connectivity runs in the constructor."""

    def __init__(self, Vpos, hexa):
        self.Vp, self.hx = _f64(Vpos), np.ascontiguousarray(hexa, np.uint32)
        self.nV, self.H = len(self.Vp), len(self.hx)
        lib().ref_clean_new.restype = C.c_void_p
        self.h = C.c_void_p(lib().ref_clean_new(_p(self.Vp), C.c_int64(self.nV), _p(self.hx), C.c_int64(self.H)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_clean_free(self.h); self.h = None

    def reorder(self):
        out = np.zeros((self.H, 8), np.uint32)
        lib().ref_clean_reorder(self.h, _p(out))
        return out

    def set_flags(self, f):
        f = np.ascontiguousarray(f, np.uint8); assert len(f) == self.H
        lib().ref_clean_set_flags(self.h, _p(f))

    def flags(self):
        f = np.zeros(self.H, np.uint8)
        lib().ref_clean_get_flags(self.h, _p(f))
        return f

    def tagging(self):
        lib().ref_clean_tagging(self.h); return self.flags()

    def reindex(self):
        sizes = (C.c_int64 * 4)()
        lib().ref_clean_reindex(self.h, sizes)
        return self.sub(list(sizes))

    def sub(self, sizes):
        nv, nh = int(sizes[0]), int(sizes[1])
        out = dict(nV=nv, nH=nh, nF=int(sizes[2]), nE=int(sizes[3]), V_map=np.zeros(self.nV, np.int32), V_map_reverse=np.zeros(nv, np.int32),
                   H_map_reverse=np.zeros(nh, np.int32), hex=np.zeros((nh, 8), np.uint32), V=np.zeros((nv, 3)))
        lib().ref_clean_sub_export(self.h, _p(out["V_map"]), _p(out["V_map_reverse"]), _p(out["H_map_reverse"]), _p(out["hex"]), _p(out["V"]))
        return out

    def non_manifold(self):
        lib().ref_clean_non_manifold(self.h); return self.flags()

    def drop_small(self):
        lib().ref_clean_drop_small(self.h); return self.flags()

    def full(self, tV, tF):
        tV, tF = _f64(tV), _i32(tF)
        lib().ref_clean_full(self.h, _p(tV), C.c_int64(len(tV)), _p(tF), C.c_int64(len(tF)))
        return self.flags()

    def medial(self):
        sizes = (C.c_int64 * 4)()
        lib().ref_clean_entire_sizes(self.h, sizes)
        Fm, Vm = np.zeros(int(sizes[2]), np.uint8), np.zeros(int(sizes[0]), np.uint8)
        lib().ref_clean_medial(self.h, _p(Fm), _p(Vm))
        return Fm, Vm


def extract_surface(Vpos, hexa, as_triangles=False):
    """extract_surface_conforming_mesh (global_functions.cpp:1021-1072) incl. orient_surface_mesh, on the hex mesh (Vpos, hexa)."""
    Vp, hx = _f64(Vpos), np.ascontiguousarray(hexa, np.uint32)
    sizes = (C.c_int64 * 6)()
    lib().ref_extract_surface.restype = C.c_void_p
    h = C.c_void_p(lib().ref_extract_surface(_p(Vp), C.c_int64(len(Vp)), _p(hx), C.c_int64(len(hx)), C.c_int(1 if as_triangles else 0), sizes))
    nv, nf, ne, nfs, nvf, nF_hex = [int(x) for x in sizes]
    vn = 3 if as_triangles else 4
    out = dict(V=np.zeros((nv, 3)), F_vs=np.zeros((nf, vn), np.uint32), F_es=np.zeros((nf, vn), np.uint32), E_vs=np.zeros((ne, 2), np.uint32),
               E_boundary=np.zeros(ne, np.uint8), V_boundary=np.zeros(nv, np.uint8), V_map=np.zeros(len(Vp), np.int32),
               V_map_reverse=np.zeros(nv, np.int32), F_map=np.zeros(nF_hex, np.int32), F_map_reverse=np.zeros(nf, np.int32))
    lib().ref_surface_export(h, _p(out["V"]), _p(out["F_vs"]), _p(out["F_es"]), _p(out["E_vs"]), _p(out["E_boundary"]), _p(out["V_boundary"]),
                             _p(out["V_map"]), _p(out["V_map_reverse"]), _p(out["F_map"]), _p(out["F_map_reverse"]))
    for which, (nm, n, tot) in enumerate((("E_nfs", ne, nfs), ("V_nvs", nv, 2 * ne), ("V_nes", nv, 2 * ne), ("V_nfs", nv, nvf))):
        off = np.zeros(n + 1, np.int64); val = np.zeros(tot, np.uint32)
        lib().ref_surface_csr(h, C.c_int(which), _p(off), _p(val))
        out[nm] = (off, val)
    lib().ref_surface_free(h)
    return out


# ---- §8(f)-3: SLIM per-element stages (slim_m.cpp:84-381, 792-916), tet branch ------------------------------------------------
SLIM_ENERGIES = {"ARAP": 0, "LOG_ARAP": 1, "SYMMETRIC_DIRICHLET": 2, "CONFORMAL": 3, "EXP_CONFORMAL": 4, "EXP_SYMMETRIC_DIRICHLET": 5}


def slim_jacobians(off, col, vx, vy, vz, uv, nv):
    """compute_jacobians (slim_m.cpp:84-106): Dx, Dy, Dz share one CSR pattern (n rows, nv columns); uv is nv x 3."""
    off = np.ascontiguousarray(off, np.int64); col = _i32(col); uv = _f64(uv); n = len(off) - 1
    Ji = np.zeros((n, 9))
    lib().ref_slim_jacobians(C.c_int64(n), C.c_int64(nv), _p(off), _p(col), _p(_f64(vx)), _p(_f64(vy)), _p(_f64(vz)), _p(uv), _p(Ji))
    return Ji


def slim_weights_rotations(J, energy, exp_factor=1.0):
    """update_weights_and_closest_rotations (slim_m.cpp:108-381, tet branch) on given Jacobians (n x 9, rows as s.Ji): (W, Ri)."""
    J = _f64(J).reshape(-1, 9); n = len(J)
    W = np.zeros((n, 9)); Ri = np.zeros((n, 9))
    lib().ref_slim_weights_rotations(_p(J), C.c_int64(n), C.c_int(SLIM_ENERGIES[energy]), C.c_double(exp_factor), _p(W), _p(Ri))
    return W, Ri


def slim_energy(J, areas, energy, exp_factor=1.0):
    """compute_energy_with_jacobians (slim_m.cpp:792-916, tet branch)."""
    J = _f64(J).reshape(-1, 9); a = _f64(areas)
    lib().ref_slim_energy.restype = C.c_double
    return float(lib().ref_slim_energy(_p(J), C.c_int64(len(J)), _p(a), C.c_int(SLIM_ENERGIES[energy]), C.c_double(exp_factor)))


def slim_max_step(uv, T, d):
    """compute_max_step_from_singularities (igl/flip_avoiding_line_search.cpp:273-299, tets): (max_step, per-tet smallest positive roots)."""
    uv, d = _f64(uv), _f64(d); T = _i32(T)
    roots = np.zeros(len(T))
    lib().ref_slim_max_step.restype = C.c_double
    m = lib().ref_slim_max_step(_p(uv), C.c_int64(len(uv)), _p(T), C.c_int64(len(T)), _p(d), _p(roots))
    return float(m), roots


def slim_rhs_terms(W, Ri):
    """per-element part of buildRhs (slim_m.cpp:1061-1083), through the reference function with At = I."""
    W = _f64(W).reshape(-1, 9); Ri = _f64(Ri).reshape(-1, 9); f = np.zeros(9 * len(W))
    lib().ref_slim_rhs_terms(_p(W), _p(Ri), C.c_int64(len(W)), _p(f))
    return f


# ---- §8(f)-4: the reference's own h_io (io.cpp) -----------------------------------------------------------------------------
def io_write_mesh(path, V, mesh_type: int, elems):
    V = _f64(V); el = np.ascontiguousarray(elems, np.uint32)
    lib().ref_io_write_mesh(str(path).encode(), _p(V), C.c_int64(len(V)), C.c_int(mesh_type), _p(el), C.c_int64(len(el)), C.c_int(el.shape[1]))


def io_write_vtk(path, V, mesh_type: int, elems, V_boundary, elem_off=None):
    V = _f64(V); vb = np.ascontiguousarray(V_boundary, np.uint8)
    if elem_off is not None:
        off = np.ascontiguousarray(elem_off, np.int64); el = np.ascontiguousarray(elems, np.uint32).reshape(-1)
        lib().ref_io_write_vtk(str(path).encode(), _p(V), C.c_int64(len(V)), C.c_int(mesh_type), _p(off), _p(el), C.c_int64(len(off) - 1), C.c_int(0), _p(vb))
    else:
        el = np.ascontiguousarray(elems, np.uint32)
        lib().ref_io_write_vtk(str(path).encode(), _p(V), C.c_int64(len(V)), C.c_int(mesh_type), None, _p(el), C.c_int64(len(el)), C.c_int(el.shape[1]), _p(vb))


def io_read_fgraph(path):
    ang = C.c_double(); oc = C.c_int(); ocs = C.c_int(); nc = C.c_int64(1 << 20); npairs = C.c_int64(1 << 20)
    corners = np.zeros(1 << 20, np.int32); pairs = np.zeros((1 << 20, 2), np.int32)
    ok = lib().ref_io_read_fgraph(str(path).encode(), C.byref(ang), C.byref(oc), C.byref(ocs), _p(corners), C.byref(nc), _p(pairs), C.byref(npairs))
    if not ok:
        return None
    return dict(angle_threshold=ang.value, orphan_curve=oc.value, orphan_curve_single=ocs.value, corners=corners[:nc.value].copy(), pairs=pairs[:npairs.value].copy())

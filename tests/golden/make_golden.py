#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz from the REFERENCE'S OWN compiled code (oracle/_ref/libfpohm_ref.so, built by
oracle/ref/Makefile from the unmodified sources under /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors, so these fixtures are what pins the oracle port
(oracle/port) and the host logic on machines where /root/reference — and thus oracle/_ref — may be absent.
Inputs are stored next to the outputs so the fixtures do not depend on the procedural generators staying unchanged.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fpohm_b200 as fp            # procedural generators only
from oracle import ref_oracle as R
from canon import canon_octree

pm = fp.procedural
G = {}
rng = np.random.default_rng(20261017)

# --- scaled Jacobian (gf.cpp:2309-2358)
V, H = pm.warped_hex_block(5, 1.3)
VJ, HJ, mad, fl = R.scaled_jacobian(V, H)
G.update(jac_V=V, jac_H=H, jac_VJ=VJ, jac_HJ=HJ, jac_mad=mad, jac_flipped=np.int64(fl))

# --- igl tree + normals + signed distance on a mesh WITHOUT barycentre ties (irrational offsets) and one WITH ties
V, F = pm.torus(24, 14)
V = V @ np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1.0]]) + np.array([0.01234, -0.0321, 0.00777])
V = np.ascontiguousarray(V @ np.array([[1, 0, 0], [0, 0.96, -0.28], [0, 0.28, 0.96]]))
rt = R.RefTree(V, F)
box, prim, lr = rt.flatten()
FN, VN, EN, E, EMAP = rt.normals()
P = np.concatenate([rng.uniform(-0.7, 0.7, (600, 3)), V[:40], V[F[:40]].mean(1), V[F[:40]][:, :2].mean(1)])
S, I, C, N = rt.signed_distance(P)
D2, I2, C2 = R.point_mesh_sqdist(V, F, P)
G.update(sd_V=V, sd_F=F, sd_box=box, sd_prim=prim, sd_lr=lr, sd_FN=FN, sd_VN=VN, sd_EN=EN, sd_EMAP=EMAP, sd_P=P, sd_S=S, sd_I=I, sd_C=C,
         sd_N=N, sd_D2=D2)
Vt, Ft, _ = pm.gear(teeth=6, n_radial=2, n_axial=3, n_arc=2)      # structured: many tied barycentre coordinates
rt2 = R.RefTree(Vt, Ft)
box2, prim2, lr2 = rt2.flatten()
Pt = rng.uniform(-0.6, 0.6, (400, 3))
St, It, Ct, Nt = rt2.signed_distance(Pt)
G.update(tie_V=Vt, tie_F=Ft, tie_box=box2, tie_prim=prim2, tie_lr=lr2, tie_P=Pt, tie_S=St, tie_I=It, tie_C=Ct, tie_N=Nt)

# --- octree: grid set-up, predicate build at two stop extents, incremental refine, random marks in all 4 modes
gs, org, mt, vs = R.octree_grid_setup(V, F, 1 << 20)
G.update(oct_gs=gs, oct_origin=org, oct_mt=mt, oct_vs=np.float64(vs))
for E_ in (17, 16):
    ro = R.RefOctree.build(V, F, gs, org, mt, vs, 1 << E_)
    for k, v in canon_octree(ro.export()).items():
        G[f"oct{E_}_{k}"] = v
    Vh, Hh, h2c = ro.hexes()
    hp = Vh[Hh.astype(np.int64)].reshape(len(Hh), 24)
    G[f"oct{E_}_hexpos"] = hp[np.lexsort(hp.T[::-1])]              # 8 corner positions per hex, rows sorted (ghm.cpp:531-562)
    G[f"oct{E_}_Vpos_sorted"] = Vh[np.lexsort(Vh.T[::-1])]
mark_cases = []
for ci, (g, gr, pa) in enumerate([([16, 16, 16], 1, 1), ([32, 16, 8], 1, 1), ([16, 8, 8], 1, 0), ([8, 8, 16], 0, 1), ([16, 16, 8], 0, 0)]):
    g = np.array(g, np.int32)
    marks = []
    for e in (2, 4, 8, 16):
        if e > g.min():
            continue
        n = g // e
        k = max(1, int(np.prod(n) * (0.2 if e > 2 else 0.04)))
        xyz = np.stack([rng.integers(0, n[d], k) for d in range(3)], -1) * e
        marks.append(np.concatenate([xyz, np.full((k, 1), e)], 1))
    marks = np.concatenate(marks).astype(np.int32)
    ro = R.RefOctree.from_marks(g, marks, gr, pa)
    G[f"marks{ci}_gs"] = g; G[f"marks{ci}_marks"] = marks; G[f"marks{ci}_mode"] = np.array([gr, pa], np.int32)
    G[f"marks{ci}_flags"] = np.array(ro.flags(), np.int32)
    for k, v in canon_octree(ro.export()).items():
        G[f"marks{ci}_{k}"] = v
G["marks_n"] = np.int32(5)

# --- ray parity: voxel grid, dexels, octree cells (compute_octree flavour)
mn, ext = V.min(0), V.max(0) - V.min(0)
vox, dims = R.voxel_sign(V, F, mn, ext, 1 / 14.5, 1)
doff, dval, d2 = R.dexel_sign(V, F, mn, ext, 1 / 14.5, 1)
G.update(vox_spacing=np.float64(1 / 14.5), vox_pad=np.int32(1), vox_out=vox, vox_dims=dims, dex_off=doff, dex_val=dval)
sp = 1 / 16
gs2 = np.array([1 << int(np.ceil(np.log2(np.ceil(e / sp)))) for e in ext], np.int32)
ro = R.RefOctree.build(V, F, gs2, mn, [0, 0, 0], sp, 1)
ins = ro.cell_sign(mn, sp)
ex = ro.export()
c0 = ex["node_pos"][ex["corner"][:, 0]]; e_ = ex["node_pos"][ex["corner"][:, 1]][:, 0] - c0[:, 0]
key = np.concatenate([c0, e_[:, None], ins[:, None].astype(np.int64)], 1)
G.update(cs_gs=gs2, cs_spacing=np.float64(sp), cs_keyed=key[np.lexsort(key.T[::-1])])

# --- connectivity (gf.cpp:121-264)
Vc, Hc = pm.warped_hex_block(3)
rc = R.hex_connectivity(Hc, len(Vc))
G.update(conn_H=Hc, conn_nV=np.int64(len(Vc)))
for k, v in rc.items():
    if isinstance(v, tuple):
        G[f"conn_{k}_off"], G[f"conn_{k}_val"] = v
    else:
        G[f"conn_{k}"] = v

# --- polyline projection (ghm.cpp:3967-3994)
loops = [np.arange(0, 30, dtype=np.int32), np.array([40, 44, 47, 52, 60, 61], np.int32)]
off = np.concatenate([[0], np.cumsum([len(l) for l in loops])]).astype(np.int64)
Pp = rng.uniform(-0.6, 0.6, (300, 3)); cid = rng.integers(0, 2, 300).astype(np.int32)
oL, aL = R.polyline_project(V, off, np.concatenate(loops), np.array([1, 0], np.uint8), Pp, cid)
G.update(pl_off=off, pl_vs=np.concatenate(loops), pl_circle=np.array([1, 0], np.uint8), pl_P=Pp, pl_cid=cid, pl_origin=oL, pl_axis=aL)

# --- metro Hausdorff (metro_hausdorff.cpp:358-505)
VB, FB = pm.torus(17, 11)
VB = VB * 1.02 + 0.004
hd = R.hausdorff(V, F, VB, FB)
G.update(hd_VB=VB, hd_FB=FB, hd_out=np.array([hd["diag"], hd["max"], hd["mean"]]))

out = Path(__file__).resolve().parent / "golden_v1.npz"
np.savez_compressed(out, **G)
print(f"wrote {out} ({out.stat().st_size / 1e3:.0f} kB, {len(G)} arrays)")

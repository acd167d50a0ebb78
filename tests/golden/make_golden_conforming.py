#!/usr/bin/env python
"""Generate tests/golden/golden_conforming_v1.npz from the REFERENCE'S OWN conforming_mesh and dual_conforming_mesh
(grid_meshing/grid_hex_meshing.cpp:568-696, :697-872 compiled unmodified into oracle/_ref/libfpohm_ref.so) on octrees built by the
reference's own OctreeGrid.  Inputs (node tables, hexes, grid size) are stored next to the outputs.

    python tests/golden/make_golden_conforming.py        # in the build container only
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_oracle as R

G = {}
rng = np.random.default_rng(20261017)
for name, gs, paired in (("a", [32, 16, 16], False), ("b", [32, 16, 16], True), ("c", [16, 16, 32], False)):
    gs = np.array(gs, np.int32)
    marks = [[x, y, z, 16] for x in range(0, gs[0], 16) for y in range(0, gs[1], 16) for z in range(0, gs[2], 16)][:2]
    for e, k in ((8, 6), (4, 10), (2, 14)):
        n = gs // e
        for _ in range(k):
            marks.append([int(rng.integers(0, n[0])) * e, int(rng.integers(0, n[1])) * e, int(rng.integers(0, n[2])) * e, e])
    marks = np.array(marks, np.int32)
    ro = R.RefOctree.from_marks(gs, marks, True, paired)
    ex = ro.export(); Vp, H, _ = ro.hexes()
    hy, dual = R.conforming_and_dual_tables(ex["node_pos"], ex["node_neigh"], Vp, H, gs)      # ghm.cpp:568-696 and :697-872
    G[f"{name}_grid"] = gs; G[f"{name}_marks"] = marks; G[f"{name}_paired"] = np.int64(paired)
    G[f"{name}_node_pos"] = ex["node_pos"]; G[f"{name}_node_neigh"] = ex["node_neigh"]; G[f"{name}_hex"] = H
    G[f"{name}_Vpos"] = Vp
    for k, v in hy.items():
        G[f"{name}_out_{k}"] = np.asarray(v)
    for k, v in dual.items():
        G[f"{name}_dual_{k}"] = np.asarray(v)
    print(name, "hexes", len(H), "faces", hy["nF"], "loops with mid vertices", int((np.diff(hy["F_off"]) > 4).sum()))
out = Path(__file__).resolve().parent / "golden_conforming_v1.npz"
np.savez_compressed(out, **G)
print(f"wrote {out} ({out.stat().st_size / 1e3:.0f} kB, {len(G)} arrays)")

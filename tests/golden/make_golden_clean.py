#!/usr/bin/env python
"""Generate tests/golden/golden_clean_v1.npz from the REFERENCE'S OWN clean_hex_mesh stages (reorder_hex_mesh,
tagging_uneven_element, re_indexing_connectivity, clean_non_manifold_ve, drop_small_pieces and the whole clean_hex_mesh with
its medial-surface flags: grid_meshing/grid_hex_meshing.cpp:1932-2124, global_functions.cpp:664-698,2199-2229, compiled
unmodified into oracle/_ref/libfpohm_ref.so).  Inputs are stored next to the outputs.

    python tests/golden/make_golden_clean.py        # in the build container only
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import ref_oracle as R
from clean_cases import carved_block, lattice_around  # noqa: E402
import fpohm_b200 as fp  # procedural meshes only (no device needed)

G = {}
for name, dims, p, seed in (("a", (6, 5, 4), 0.6, 1), ("b", (8, 8, 8), 0.5, 2), ("c", (10, 6, 7), 0.75, 3), ("d", (9, 9, 9), 0.35, 4)):
    V, H, flag, Hm = carved_block(dims, p, seed)
    rc = R.RefClean(V, Hm)
    G[f"{name}_V"] = V; G[f"{name}_hex"] = H; G[f"{name}_hex_mirrored"] = Hm; G[f"{name}_flag"] = flag
    G[f"{name}_reordered"] = rc.reorder()
    rc = R.RefClean(V, H)
    rc.set_flags(flag)
    t = rc.tagging(); G[f"{name}_tagged"] = t
    s = rc.reindex()
    for k in ("V_map", "V_map_reverse", "H_map_reverse", "hex"):
        G[f"{name}_sub_{k}"] = s[k]
    n = rc.non_manifold(); G[f"{name}_manifold"] = n
    d = rc.drop_small(); G[f"{name}_dropped"] = d
    print(name, dims, "tag", int((t != flag).sum()), "non-manifold", int((n != t).sum()), "dropped", int((d != n).sum()), "left", int(d.sum()))

# whole clean_hex_mesh: a lattice around a torus / two linked tori (two pieces -> drop_small_pieces has work)
for name, (tV, tF), n in (("torus", fp.procedural.torus(40, 24), 14), ("tori", fp.procedural.linked_tori(2, 16, 8), 18)):
    V, H = lattice_around(tV, n)
    rc = R.RefClean(V, H)
    flag = rc.full(tV, tF)
    Fm, Vm = rc.medial()
    G[f"{name}_tV"] = tV; G[f"{name}_tF"] = tF; G[f"{name}_V"] = V; G[f"{name}_hex"] = H
    G[f"{name}_flag"] = flag; G[f"{name}_F_medial"] = Fm; G[f"{name}_V_medial"] = Vm
    print(name, "hexes", len(H), "inside", int(flag.sum()), "medial faces", int(Fm.sum()))
# extract_surface_conforming_mesh + orient_surface_mesh (gf.cpp:1021-1112): a block with a cavity, and the cleaned sub-mesh of case b
# with every third hex mirrored
from clean_cases import block  # noqa: E402
V, H = block(3, 3, 3); keep = np.ones(len(H), bool); keep[13] = False
surf_cases = [("surf_cavity", V, H[keep])]
rc = R.RefClean(G["b_V"], G["b_hex"]); rc.set_flags(G["b_dropped"]); s = rc.reindex()
sh = s["hex"].copy(); sh[::3] = sh[::3][:, [3, 2, 1, 0, 7, 6, 5, 4]]
surf_cases.append(("surf_carved", s["V"], sh))
for name, V, H in surf_cases:
    G[f"{name}_V"] = V; G[f"{name}_hex"] = H
    for tri in (0, 1):
        r = R.extract_surface(V, H, bool(tri))
        for k, v in r.items():
            if isinstance(v, tuple):
                G[f"{name}_{tri}_{k}_off"] = v[0]; G[f"{name}_{tri}_{k}_val"] = v[1]
            else:
                G[f"{name}_{tri}_{k}"] = v
    print(name, "hexes", len(H), "surface faces", len(r["F_vs"]))
out = Path(__file__).resolve().parent / "golden_clean_v1.npz"
np.savez_compressed(out, **G)
print(f"wrote {out} ({out.stat().st_size / 1e3:.0f} kB, {len(G)} arrays)")

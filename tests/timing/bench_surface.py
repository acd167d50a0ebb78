#!/usr/bin/env python
"""Timing of extract_surface_conforming_mesh (+ orient_surface_mesh) on the cleaned lattice around the bench gear."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import fpohm_b200 as fp
ctx = fp.Context(0)
tV, tF = fp.procedural.gear()[:2]
m = fp.TriMesh(ctx, tV, tF); m.build_aabb_tree()
for n in [int(a) for a in sys.argv[1:]] or [64, 160, 256]:
    V, H = fp.procedural.hex_lattice_around(tV, n)
    r = fp.clean_hex_mesh(ctx, m, V, H)
    s = fp.reindex_submesh(ctx, r["hex"], len(V), r["H_flag"])
    Vs = V[s["V_map_reverse"]]
    conn = fp.HexConnectivity(ctx, s["hex"], len(Vs), keep=True)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); q = fp.extract_surface(ctx, conn, Vs, False); best = min(best, time.perf_counter() - t0)
    rec = dict(hexes=len(s["hex"]), quads=len(q["F_vs"]), bfs_levels=q["bfs_levels"], ms=best * 1e3, kernel_ms=ctx.last_kernel_ms())
    if len(s["hex"]) <= 300000:
        from oracle import ref_oracle as R
        if R.available():
            t0 = time.perf_counter(); w = R.extract_surface(Vs, s["hex"], False); rec["reference_ms_1core_incl_hex_connectivity"] = (time.perf_counter() - t0) * 1e3
            rec["equal"] = bool(np.array_equal(w["F_vs"], q["F_vs"]))
    print(json.dumps({n: rec}), flush=True)
    conn.close()

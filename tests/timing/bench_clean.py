#!/usr/bin/env python
"""Timing of clean_hex_mesh on lattices around the bench gear: product (B200, wall clock of the C-ABI call with
host buffers) vs the compiled reference on one host core at the size the reference finishes in seconds."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fpohm_b200 as fp
from clean_cases import lattice_around

ctx = fp.Context(0)
tV, tF = fp.procedural.gear()[:2]
m = fp.TriMesh(ctx, tV, tF); m.build_aabb_tree()
out = {}
for n in [int(a) for a in sys.argv[1:]] or [64, 160, 256]:
    V, H = lattice_around(tV, n)
    conn = fp.HexConnectivity(ctx, H, len(V), keep=True)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); r = fp.clean_hex_mesh(ctx, m, V, H, conn); best = min(best, time.perf_counter() - t0)
    rec = dict(hexes=len(H), verts=len(V), ms=best * 1e3, stats=r["stats"])
    if len(H) <= 400000:
        from oracle import ref_oracle as R
        if R.available():
            rc = R.RefClean(V, H)
            t0 = time.perf_counter(); want = rc.full(tV, tF); rec["reference_ms_1core"] = (time.perf_counter() - t0) * 1e3
            rec["equal"] = bool(np.array_equal(want, r["H_flag"]))
    out[f"lattice_{n}"] = rec
    print(json.dumps({n: rec}), flush=True)
    conn.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_clean.json").write_text(json.dumps(out, indent=1))

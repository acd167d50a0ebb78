#!/usr/bin/env python
"""Randomised comparison of the cleaning stages and the surface extraction against the compiled reference (more seeds and
densities than the unit tests; run on a B200: python tests/timing/fuzz_clean.py 60)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fpohm_b200 as fp
from oracle import ref_oracle as R
from clean_cases import carved_block

ctx = fp.Context(0)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 777)
bad = 0; rounds_hist = {}; surf = 0
for case in range(n_cases):
    dims = tuple(int(x) for x in rng.integers(3, 15, 3)); p = float(rng.uniform(0.15, 0.95)); seed = int(rng.integers(1 << 30))
    V, H, flag, Hm = carved_block(dims, p, seed)
    nV = len(V)
    rc = R.RefClean(V, H); rc.set_flags(flag)
    conn = fp.HexConnectivity(ctx, H, nV, keep=True)
    t, _ = fp.tag_uneven_elements(ctx, conn, flag); conn.close()
    ok = np.array_equal(t, rc.tagging())
    if t.any():
        rc.reindex()
        n, rounds = fp.clean_non_manifold(ctx, H, nV, t)
        ok &= np.array_equal(n, rc.non_manifold())
        rounds_hist[rounds] = rounds_hist.get(rounds, 0) + 1
        if n.any():
            d, pieces = fp.drop_small_pieces(ctx, H, nV, n)
            ok &= np.array_equal(d, rc.drop_small())
            s = fp.reindex_submesh(ctx, H, nV, d)
            Vs = V[s["V_map_reverse"]]; sh = s["hex"].copy()
            sh[::2] = sh[::2][:, [3, 2, 1, 0, 7, 6, 5, 4]]
            sc = fp.HexConnectivity(ctx, sh, len(Vs), keep=True)
            for tri in (False, True):
                got = fp.extract_surface(ctx, sc, Vs, tri); want = R.extract_surface(Vs, sh, tri)
                for k, v in want.items():
                    same = (np.array_equal(got[k][0], v[0]) and np.array_equal(got[k][1], v[1])) if isinstance(v, tuple) else np.array_equal(np.asarray(got[k]), np.asarray(v))
                    ok &= bool(same)
                surf += 1
            sc.close()
    if not ok:
        bad += 1; print("MISMATCH", dims, p, seed, flush=True)
print(f"{n_cases} cases, {bad} mismatches, non-manifold rounds histogram {dict(sorted(rounds_hist.items()))}, {surf} surfaces compared")
sys.exit(1 if bad else 0)

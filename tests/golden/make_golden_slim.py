#!/usr/bin/env python
"""Generate tests/golden/golden_slim_v1.npz from the REFERENCE'S OWN SLIM per-element functions (compute_jacobians,
update_weights_and_closest_rotations, compute_energy_with_jacobians of slim_m.cpp, compiled unmodified into oracle/_ref).
Inputs are stored next to the outputs.

    python tests/golden/make_golden_slim.py        # in the build container only
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_oracle as R

rng = np.random.default_rng(20261017)
n = 600
J = np.eye(3)[None] + rng.normal(0, 0.35, (n, 3, 3))
J[::7] = J[::7] @ np.diag([1, 1, -1.0])
J[1] = np.eye(3); J[2] = np.diag([2.0, 1.0, 0.5]); J[3] = np.diag([1.0 + 5e-9, 1.0, 1.0 - 5e-9])
J = J.reshape(n, 9)
areas = rng.uniform(0.5, 2.0, n)
G = {"J": J, "areas": areas, "exp_factor": np.float64(0.7)}
for en in R.SLIM_ENERGIES:
    W, Ri = R.slim_weights_rotations(J, en, 0.7)
    G[f"{en}_W"] = W; G[f"{en}_Ri"] = Ri; G[f"{en}_energy"] = np.float64(R.slim_energy(J, areas, en, 0.7))
nv, nt = 200, 500
col = rng.integers(0, nv, (nt, 4)).astype(np.int32); off = np.arange(0, 4 * nt + 1, 4)
vx, vy, vz = rng.normal(size=(3, 4 * nt)); uv = rng.normal(size=(nv, 3))
G.update(jac_off=off, jac_col=col.reshape(-1), jac_vx=vx, jac_vy=vy, jac_vz=vz, jac_uv=uv, jac_Ji=R.slim_jacobians(off, col.reshape(-1), vx, vy, vz, uv, nv))
# flip-avoiding step bound (igl/flip_avoiding_line_search.cpp:177-299) on the 8-tets-per-hex split of a warped block
import fpohm_b200 as fp  # procedural meshes only
Vb, Hb = fp.procedural.warped_hex_block(5, 0.3)
Tb = Hb[:, [[0, 1, 3, 4], [1, 2, 0, 5], [2, 3, 1, 6], [3, 0, 2, 7], [4, 7, 5, 0], [5, 4, 6, 1], [6, 5, 7, 2], [7, 6, 4, 3]]].reshape(-1, 4).astype(np.int32)
G["step_uv"] = Vb; G["step_T"] = Tb
for k, sc in enumerate((0.02, 0.2, 2.0)):
    d = rng.normal(0, sc, Vb.shape)
    m, r = R.slim_max_step(Vb, Tb, d)
    G[f"step{k}_d"] = d; G[f"step{k}_max"] = np.float64(m); G[f"step{k}_roots"] = r
G["rhs_terms"] = R.slim_rhs_terms(G["SYMMETRIC_DIRICHLET_W"], G["SYMMETRIC_DIRICHLET_Ri"])      # buildRhs with At = I (slim_m.cpp:1044-1093)
out = Path(__file__).resolve().parent / "golden_slim_v1.npz"
np.savez_compressed(out, **G)
print(f"wrote {out} ({out.stat().st_size / 1e3:.0f} kB, {len(G)} arrays)")

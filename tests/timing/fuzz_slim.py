#!/usr/bin/env python
"""Worst-case agreement of the SLIM weights / rotations with the compiled reference over condition numbers 1 .. 1e10."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import fpohm_b200 as fp
from oracle import ref_oracle as R
ctx = fp.Context(0)
rng = np.random.default_rng(99)
n = 20000
worst = 0
for cond in (1e0, 1e2, 1e4, 1e6, 1e8, 1e10):
    U = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]; Vt = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    s = np.stack([np.ones(n) * rng.uniform(0.5, 2, n), rng.uniform(1 / np.sqrt(cond), 1, n), np.ones(n) / cond], 1)
    J = ((U * s[:, None, :]) @ Vt).reshape(n, 9)
    for en in ("ARAP", "SYMMETRIC_DIRICHLET", "LOG_ARAP", "CONFORMAL"):
        W, Ri = fp.slim_weights_rotations(ctx, J, en, 1.0); rW, rRi = R.slim_weights_rotations(J, en, 1.0)
        fin = np.isfinite(rW).all(1) & np.isfinite(W).all(1)
        dw = (np.abs(W[fin] - rW[fin]).max(1) / np.abs(rW[fin]).max(1)).max(); dr = np.abs(Ri[fin] - rRi[fin]).max()
        worst = max(worst, dw, dr)
        print(f"cond {cond:7.0e} {en:20s} W rel {dw:.2e}  Ri abs {dr:.2e}  non-finite rows ref/ours {int((~np.isfinite(rW).all(1)).sum())}/{int((~np.isfinite(W).all(1)).sum())}", flush=True)
print("worst", worst)

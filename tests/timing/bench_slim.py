#!/usr/bin/env python
"""Timing of the SLIM per-element stages at config C4 scale (10.08 M hexes x 8 tets), device resident, CUDA events; the compiled
reference on a sample of the same Jacobians on one host core."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch
import fpohm_b200 as fp
PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]) if (ROOT / "MEASURED_PEAKS.json").exists() else 6553.6
ctx = fp.Context(0)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * 10_077_696
g = torch.Generator(device=dev); g.manual_seed(5)
J = (torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9) + 0.3 * torch.randn(n, 9, dtype=torch.float64, device=dev, generator=g)).contiguous()
areas = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
W = torch.empty(n, 9, dtype=torch.float64, device=dev); Ri = torch.empty(n, 9, dtype=torch.float64, device=dev); E = torch.empty(1, dtype=torch.float64, device=dev)
def timed(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): f()
    b.record(st); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {"tets": n}
for en in ("SYMMETRIC_DIRICHLET", "ARAP", "EXP_CONFORMAL"):
    ms = timed(lambda: fp.slim_weights_rotations_dev(ctx, J.data_ptr(), n, en, 1.0, W.data_ptr(), Ri.data_ptr(), st.cuda_stream))
    ms_e = timed(lambda: fp.slim_energy_dev(ctx, J.data_ptr(), n, areas.data_ptr(), en, 1.0, E.data_ptr(), st.cuda_stream))
    out[en] = {"weights_rotations_ms": ms, "tets_per_s": n / ms * 1e3, "GBs": 216 * n / ms / 1e6, "frac": 216 * n / ms / 1e6 / PEAK,
               "energy_ms": ms_e, "energy_GBs": 80 * n / ms_e / 1e6, "energy_frac": 80 * n / ms_e / 1e6 / PEAK}
try:
    from oracle import ref_oracle as R
    if R.available():
        m = 400_000
        Jh = J[:m].cpu().numpy(); ah = areas[:m].cpu().numpy()
        t0 = time.perf_counter(); R.slim_weights_rotations(Jh, "SYMMETRIC_DIRICHLET", 1.0); dt = time.perf_counter() - t0
        t0 = time.perf_counter(); R.slim_energy(Jh, ah, "SYMMETRIC_DIRICHLET", 1.0); de = time.perf_counter() - t0
        out["reference_1core"] = {"sample_tets": m, "weights_rotations_tets_per_s": m / dt, "energy_tets_per_s": m / de}
except Exception as e:
    out["reference_1core"] = {"error": str(e)}
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_slim.json").write_text(json.dumps(out, indent=1))

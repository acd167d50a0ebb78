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
# line-search step bound on the 8-tets-per-hex split of a 128^3 block (16.8 M tets), resident
Vb, Hb = fp.procedural.warped_hex_block(128, 0.3)
Tb = np.ascontiguousarray(Hb[:, [[0, 1, 3, 4], [1, 2, 0, 5], [2, 3, 1, 6], [3, 0, 2, 7], [4, 7, 5, 0], [5, 4, 6, 1], [6, 5, 7, 2], [7, 6, 4, 3]]].reshape(-1, 4).astype(np.int32))
dV = torch.from_numpy(Vb).to(dev); dT = torch.from_numpy(Tb).to(dev); dD = 0.01 * torch.randn(len(Vb), 3, dtype=torch.float64, device=dev, generator=g)
dM = torch.empty(1, dtype=torch.float64, device=dev)
import ctypes as C
ms = timed(lambda: fp.api._chk(fp.lib().fpohm_slim_max_step_dev(ctx.h, C.c_void_p(dV.data_ptr()), C.c_void_p(dT.data_ptr()), C.c_int64(len(Tb)), C.c_void_p(dD.data_ptr()),
                                                                 None, C.c_void_p(dM.data_ptr()), C.c_void_p(st.cuda_stream))))
out["max_step"] = {"tets": len(Tb), "ms": ms, "tets_per_s": len(Tb) / ms * 1e3, "bound": float(dM.item())}
try:
    from oracle import ref_oracle as R
    if R.available():
        m = 400_000
        Jh = J[:m].cpu().numpy(); ah = areas[:m].cpu().numpy()
        t0 = time.perf_counter(); R.slim_weights_rotations(Jh, "SYMMETRIC_DIRICHLET", 1.0); dt = time.perf_counter() - t0
        t0 = time.perf_counter(); R.slim_energy(Jh, ah, "SYMMETRIC_DIRICHLET", 1.0); de = time.perf_counter() - t0
        mt = 400_000
        t0 = time.perf_counter(); R.slim_max_step(Vb, Tb[:mt], dD.cpu().numpy()); dstep = time.perf_counter() - t0
        out["reference_1core_max_step_tets_per_s"] = mt / dstep / 2      # the driver evaluates every tet twice (roots + bound)
        out["reference_1core"] = {"sample_tets": m, "weights_rotations_tets_per_s": m / dt, "energy_tets_per_s": m / de}
except Exception as e:
    out["reference_1core"] = {"error": str(e)}
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_slim.json").write_text(json.dumps(out, indent=1))

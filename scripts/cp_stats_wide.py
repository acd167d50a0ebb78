"""Debug: wide-packet counters (FPOHM_CP_STATS=1, FPOHM_CP_MODE=2): per packet phase-A steps, candidates, staged clusters,
exact-loop iterations, per-lane exact evaluations, cycles."""
import os, sys
from pathlib import Path
os.environ["FPOHM_CP_STATS"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import fpohm_b200 as fp
import bench
ctx = fp.Context(0)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
which = set(sys.argv[1:]) or {"gear", "c3"}


def report(mesh, name, Q):
    Q = np.ascontiguousarray(Q); n = len(Q)
    dP = torch.from_numpy(Q).to(dev)
    dS = torch.empty(n, dtype=torch.float64, device=dev); dI = torch.empty(n, dtype=torch.int32, device=dev)
    dC = torch.empty(n, 3, dtype=torch.float64, device=dev); dN = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    mesh.signed_distance_dev(dP.data_ptr(), n, dS.data_ptr(), dI.data_ptr(), dC.data_ptr(), dN.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    N = dN.cpu().numpy()
    a = N[:, 0].astype(np.int64); b = N[:, 1].astype(np.int64); c = N[:, 2].astype(np.int64)
    steps, cand, code = a & 0xffff, (a >> 16) & 0xffff, a >> 32
    p1, iters, evals = b & 0xffff, (b >> 16) & 0xffff, b >> 32
    p2, cyc = c & 0xfffff, (c >> 20) << 4
    f = lambda x: f"{x.mean():.1f} (p50 {np.percentile(x, 50):.0f} p90 {np.percentile(x, 90):.0f} p99 {np.percentile(x, 99):.0f} max {x.max():.0f})"
    print(f"{name}: n={n}\n  per packet: A-steps {f(steps)}\n  candidates {f(cand)}\n  (lane,cluster) pairs {f(p1)}\n  (lane,facet) pairs {f(p2)}\n"
          f"  C iterations {f(iters)}\n  exact evals per lane {f(evals)}\n  cycles {f(cyc)}\n"
          f"  code0 {np.mean(code == 0):.4f} walk {np.mean(code == 1):.4f} search {np.mean(code == 2):.5f} ({int(np.sum(code == 2))}) heavy {np.mean(code == 3):.5f} ({int(np.sum(code == 3))})", flush=True)
    pk = cyc.reshape(-1)[: n // 32 * 32].reshape(-1, 32)[:, 0]
    srt = np.sort(pk)[::-1]
    print(f"  packets {len(pk)}: total cycles {pk.sum():.3g}; top 10 packets {srt[:10].astype(int).tolist()}; share of top 1% packets {srt[: len(pk) // 100].sum() / pk.sum():.3f}")


if "gear" in which:
    V, F, _ = fp.procedural.gear()
    mesh = fp.TriMesh(ctx, V, F)
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << 12
    o = fp.Octree.build(ctx, mesh, prm)
    Vh, H, _ = o.hexes()
    ext = Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0]
    P = bench.make_queries(Vh, H, ext)
    report(mesh, "gear:bench", P)
    report(mesh, "gear:hexverts", Vh[: 1 << 20])
    report(mesh, "gear:onverts", np.repeat(V, 2, 0)[: 1 << 18])
if "c3" in which:
    V, F = fp.procedural.c3_mesh()
    mesh = fp.TriMesh(ctx, V, F)
    proj, cls = fp.procedural.c4_queries(V, F)
    report(mesh, "c3:project", proj)
    report(mesh, "c3:classify", cls)

"""Per-kernel measurements on the other BASELINE.json configs (C3-C5), device-resident, CUDA events on the launching
stream.  Not the driver's bench (that is bench.py); results feed DESIGN.md / profiles/.

    python scripts/bench_kernels.py [jacobian] [voxel] [octree] [hausdorff] [connectivity] [query]
"""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import fpohm_b200 as fp

PEAK = 6553.6
try:
    PEAK = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
which = set(sys.argv[1:]) or {"jacobian", "voxel", "occupancy", "octree", "hausdorff", "connectivity", "query"}
pm = fp.procedural
dev = torch.device("cuda", 0)
ctx = fp.Context(0)
stream = torch.cuda.current_stream()
out = {}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def c3_mesh():
    V, F = pm.linked_tori(4, 90, 44)
    return pm.midpoint_subdivide(V, F, 1)      # 2 027 520 triangles, genus 64


if "jacobian" in which:   # C4: 216^3 = 10 077 696 hexes
    V, H = pm.warped_hex_block(216)
    dV = torch.from_numpy(V).to(dev); dH = torch.from_numpy(H.view(np.int32)).to(dev)
    nH, nV = len(H), len(V)
    dVJ = torch.empty(8 * nH, dtype=torch.float64, device=dev); dHJ = torch.empty(nH, dtype=torch.float64, device=dev)
    st = torch.empty(3, dtype=torch.float64, device=dev); fl = torch.empty(1, dtype=torch.int64, device=dev)
    ms = timed(lambda: fp.scaled_jacobian_dev(ctx, dV.data_ptr(), nV, dH.data_ptr(), nH, dVJ.data_ptr(), dHJ.data_ptr(), st.data_ptr(), fl.data_ptr(), stream.cuda_stream))
    by = 104 * nH + 24 * nV
    out["jacobian_C4"] = dict(hexes=nH, ms=ms, hexes_per_s=nH / ms * 1e3, GBs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK, stats=st.tolist())
    del dV, dH, dVJ, dHJ

if which & {"voxel", "occupancy", "octree", "hausdorff", "query"}:
    t = time.perf_counter(); V, F = c3_mesh(); gen_s = time.perf_counter() - t
    mesh = fp.TriMesh(ctx, V, F)
    out["C3_mesh"] = dict(tris=len(F), verts=len(V), gen_s=gen_s)

if "voxel" in which:
    mn, ext = V.min(0), V.max(0) - V.min(0)
    for n in (512, 1024):
        g = fp.VoxelGrid(mn, ext, 1.0 / n, 0)
        buf = torch.empty(g.num_voxels(), dtype=torch.uint8, device=dev)
        ms = timed(lambda: fp.voxel_sign_dev(ctx, mesh, g, buf.data_ptr(), stream.cuda_stream), reps=5, warm=2)
        by = g.num_voxels() + 72 * len(F)
        out[f"voxel_sign_{n}"] = dict(dims=g.dims.tolist(), voxels=g.num_voxels(), ms=ms, GBs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK,
                                      inside=int(buf.sum().item()))
        del buf

if "occupancy" in which:
    mn, ext = V.min(0), V.max(0) - V.min(0)
    for n in (512, 1024):
        g = fp.VoxelGrid(mn, ext, 1.0 / n, 0)
        out_h = np.zeros(g.num_voxels(), np.uint8)
        import ctypes as C
        for _ in range(3):
            rc = fp.lib().fpohm_voxel_occupancy(ctx.h, mesh.h, g.origin.ctypes.data_as(C.c_void_p), C.c_double(g.spacing), g.dims.ctypes.data_as(C.c_void_p),
                                                out_h.ctypes.data_as(C.c_void_p))
            assert rc == 0
        ms = ctx.last_kernel_ms()
        by = g.num_voxels() + 72 * len(F)
        out[f"voxel_occupancy_{n}"] = dict(voxels=g.num_voxels(), kernel_ms=ms, GBs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK, occupied=int(out_h.sum()))

if "octree" in which:
    p = fp.octree_grid_setup(V, 1 << 20)
    for E in (12, 11, 10):
        p.c.stop_extent = 1 << E
        ts = []
        for i in range(4):
            ctx.sync(); t = time.perf_counter(); o = fp.Octree.build(ctx, mesh, p); ctx.sync(); ts.append((time.perf_counter() - t) * 1e3)
            sz = o.sizes()
            if i < 3:
                o.close()
        out[f"octree_C3_e{E}"] = dict(ms=min(ts[1:]), first_ms=ts[0], **sz)
        if E != 10:
            o.close()
    oct10 = o

if "query" in which:
    t = time.perf_counter(); mesh.build_aabb_tree(); out["C3_tree_build_host_s"] = time.perf_counter() - t
    Vh, Hh, _ = oct10.hexes() if "octree" in which else (None, None, None)
    if Vh is not None:
        P = Vh[Hh.astype(np.int64)].mean(1)
        Q = len(P)
        dP = torch.from_numpy(P).to(dev); dS = torch.empty(Q, dtype=torch.float64, device=dev)
        ms = timed(lambda: mesh.signed_distance_dev(dP.data_ptr(), Q, dS.data_ptr(), 0, 0, 0, stream.cuda_stream), reps=3, warm=1)
        out["classify_hex_centres_C3"] = dict(queries=Q, ms=ms, qps=Q / ms * 1e3, inside=int((dS < 0).sum().item()))

if "hausdorff" in which:
    VB, FB = pm.linked_tori(4, 64, 32)
    VB = VB * 1.001
    B = fp.TriMesh(ctx, VB, FB)
    mesh.build_aabb_tree(); B.build_aabb_tree()
    for extra in (0, 25_000_000):
        ctx.sync(); t = time.perf_counter(); h = fp.hausdorff(ctx, mesh, B, extra); dt = time.perf_counter() - t
        out[f"hausdorff_extra{extra}"] = dict(samples=h["n_ab"] + h["n_ba"], s=dt, samples_per_s=(h["n_ab"] + h["n_ba"]) / dt, max=h["max"], mean=h["mean"],
                                              kernel_ms=ctx.last_kernel_ms())

if "connectivity" in which:
    Vc, Hc = pm.warped_hex_block(128)   # 2.1 M hexes
    ctx.sync(); t = time.perf_counter(); c = fp.lib()
    import ctypes as C
    h = C.c_void_p()
    Hc = np.ascontiguousarray(Hc)
    rc = c.fpohm_hex_connectivity(ctx.h, Hc.ctypes.data_as(C.c_void_p), C.c_int64(len(Hc)), C.c_int64(len(Vc)), C.byref(h))
    dt = time.perf_counter() - t
    assert rc == 0
    out["connectivity_2M"] = dict(hexes=len(Hc), s=dt, kernel_ms=ctx.last_kernel_ms())
    c.fpohm_conn_free(h)

print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""bench.py — headline measurement of the hot path (BASELINE.json: "voxel+octree build ms and closest-point queries/s at
1/2/4/8 B200 vs CPU TBB").

Workload = BASELINE config C4 ON THE C3 MESH: the 2 027 520-facet genus-64 surface (igl tree 260 MB + wide
tree 74 MB + triangles 146 MB: larger than the 126 MB L2) queried by
  * the projection set — 1.5 M points jittered +-2h around the surface in facet order plus the 279 938 boundary vertices of
    the 216^3 hex block (what projection_smooth / dirty_graph_projection ask, ghm.cpp:3760-3781,4034-4081), and
  * the classification set — the 10 077 696 hex centres of that block in the lattice's own z-fastest order (what
    clean_hex_mesh's points_inside_mesh asks, ghm.cpp:1937-1951),
11 857 634 queries per step, 84 B/query algorithmic = 996 MB per step (inputs + outputs > L2).  One STEP = both sets through
igl::signed_distance_pseudonormal's replacement.  `value` = queries/s with inputs resident in HBM; `e2e` = the same through
the host-pointer C-ABI call (`fpohm_signed_distance`) with pinned host buffers, H2D + kernels + D2H inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (torchrun, one rank per GPU): STRONG scaling — rank r takes `shard_range` of each of the two query sets of the ONE job,
mesh + trees replicated, and the results are all-gathered over NCCL into every rank's (hence rank 0's) HBM inside the timed
region.  Timing = CUDA events on the launching stream, max over ranks.  `also` carries the other configs, each with its own
roofline: C3 1024^3 z-ray voxelization + `--e 10` octree build (median AND max of 8), C4 scaled Jacobian (10.08 M hexes),
C5 metro Hausdorff (50 M samples, hex boundary surface vs the C3 mesh, VCG similar-triangle rule), the C2 gear numbers of
round 1, and for N > 1 the z-slab sharded octree on the C3 mesh.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

STOP_E = 12            # C2: --e 12  => depth 8 on the 2^20 grid
BYTES_PER_QUERY = 84   # 24 B read + 8 (S) + 4 (I) + 24 (C) + 24 (N) written
BYTES_PER_HEX = 104    # J1: 32 B ids + 72 B written ; + 24 B per vertex once
BYTES_PER_SAMPLE = 32  # H1: 24 B materialised sample + 8 B distance
WORKLOAD = ("C4 on the C3 mesh: 2 027 520-facet genus-64 surface; signed-distance queries = projection set (1.5 M points +-2h around "
            "the surface + 279 938 block boundary vertices) + classification set (10 077 696 lattice hex centres, z-fastest order)")


def make_queries(Vh, H, extent_of_leaf, copies=3, seed=1234):
    """C2 query set of round 1 (kept for continuity and for the full-size parity test): leaf-hex centres + jittered copies."""
    centres = Vh[H.astype(np.int64)].mean(1)
    rng = np.random.Generator(np.random.PCG64(seed))
    out = [centres]
    for _ in range(copies):
        out.append(centres + (rng.random(centres.shape) - 0.5) * extent_of_leaf[:, None])
    # keep the spatial (Morton-ish) order of the leaves: copy k of leaf i sits next to leaf i
    return np.ascontiguousarray(np.stack(out, 1).reshape(-1, 3))


def shard_range(n, rank, world):
    q, r = divmod(int(n), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.samples, self.stop_flag, self.index = [], False, index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5)
                p = [x.strip() for x in r.stdout.strip().split(",")]
                if len(p) >= 7:
                    self.samples.append(p)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(local):
    """e2e at N > 1 moved all 8 ranks' pinned traffic through whichever NUMA node the processes happened to start on (round 1:
    efficiency 0.32 at N = 8).  Pin this process (and therefore its pinned allocations, first touch) to the CPUs of the node its
    GPU hangs off.  Best effort: silently does nothing where sysfs does not say."""
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True, timeout=5)
        bus = r.stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def c4_workload(fp):
    V, F = fp.procedural.c3_mesh()
    proj, cls = fp.procedural.c4_queries(V, F)
    return V, F, proj, cls


def ref_signed_distance_mt(rt, P, threads):
    out = [None] * threads
    chunks = np.array_split(np.arange(len(P)), threads)

    def work(k):
        out[k] = rt.signed_distance(np.ascontiguousarray(P[chunks[k]]))
    th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    [t.start() for t in th]; [t.join() for t in th]
    return [np.concatenate([o[j] for o in out]) for j in range(4)]


def cpu_baseline_and_parity(V, F, P, gpu_results, seconds_target=12.0):
    """oracle/_ref (the reference's own igl code) on a bounded every-k-th sample of the step's queries, ONE thread — how the
    reference runs on Linux (no -fopenmp) — and every answer it gives compared bit for bit with what the
    GPU path returned for the same queries."""
    from oracle import ref_oracle as R
    t = time.perf_counter(); rt = R.RefTree(V, F); tree_s = time.perf_counter() - t
    n = min(len(P), 20000)
    idx = np.linspace(0, len(P) - 1, n).astype(np.int64)
    t = time.perf_counter(); rt.signed_distance(P[idx]); dt = time.perf_counter() - t
    n2 = int(min(len(P), max(n, n * seconds_target / max(dt, 1e-6))))
    idx = np.unique(np.linspace(0, len(P) - 1, n2).astype(np.int64))
    t = time.perf_counter(); rS, rI, rC, rN = rt.signed_distance(P[idx]); dt = time.perf_counter() - t
    S, I, C, N = gpu_results
    par = {"checked": int(len(idx)), "mismatch_I": int((I[idx] != rI).sum()), "mismatch_S": int((S[idx] != rS).sum()),
           "mismatch_C": int((C[idx] != rC).any(1).sum()), "mismatch_N": int((N[idx] != rN).any(1).sum()),
           "against": "igl::signed_distance_pseudonormal compiled from /root/reference (oracle/_ref), bitwise"}
    base = {"value": len(idx) / dt, "unit": "queries/s", "cores": 1, "kind": "reference",
            "sample": f"{len(idx)} of the {len(P)} queries of a step (every k-th), igl::signed_distance_pseudonormal via oracle/_ref, {dt:.1f} s; "
                      f"igl::AABB::init + normals of the 2 027 520-facet mesh {tree_s:.1f} s (not in the rate)"}
    return base, par


def run_reference(args):
    """--impl reference: the reference's own CPU code (oracle/_ref) on the same workload, bounded sample per step, every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import fpohm_b200 as fp  # procedural generators only; no kernels are launched on this arm
    from oracle import ref_oracle as R
    V, F, proj, cls = c4_workload(fp)
    P = np.concatenate([proj, cls])
    t = time.perf_counter(); rt = R.RefTree(V, F); tree_s = time.perf_counter() - t
    cores = os.cpu_count() or 1
    n_step = min(len(P), 20000 * cores)
    idx = np.linspace(0, len(P) - 1, n_step).astype(np.int64)
    sample = np.ascontiguousarray(P[idx])

    def step():
        ref_signed_distance_mt(rt, sample, cores)
    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t) / args.steps
    val = len(sample) / dt
    # the other halves of the metric on the reference: C2 octree build (serial: no parallel form exists), Jacobian
    also = {"query_tree_build_s": tree_s}
    try:
        gV, gF, _ = fp.procedural.gear()
        gs, org, mt, vs = R.octree_grid_setup(gV, gF, 1 << 20)
        t = time.perf_counter(); ro = R.RefOctree.build(gV, gF, gs, org, mt, vs, 1 << STOP_E); Vh, H, _ = ro.hexes()
        also["c2_octree_build_ms"] = (time.perf_counter() - t) * 1e3
        also["c2_octree_note"] = "OctreeGrid::subdivide + hex export on the C2 gear, serial (no parallel form exists)"
        t = time.perf_counter(); R.scaled_jacobian(Vh, H); also["jacobian_hexes_per_s"] = len(H) / (time.perf_counter() - t)
    except Exception as e:
        also["error"] = str(e)
    line = {"impl": "reference", "metric": "closest_point_queries_per_s", "value": val, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "queries_total": int(len(P)), "queries_per_step": int(len(sample)), "tris": int(len(F))},
            "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "reference",
                             "sample": f"{len(sample)} queries per step (every k-th of the {len(P)} of a step), {cores} threads over query slices "
                                       "(what igl's inert `#pragma omp parallel for` would do); oracle/_ref"},
            "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "also": also}
    emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's "NCCL version ..." banner when NCCL_DEBUG is set by the
    environment or /etc/nccl.conf), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="headline only (skip the C2/C3/C5 entries of 'also')")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import fpohm_b200 as fp

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = fp.Context(local)
    stream = torch.cuda.current_stream()
    peak, peak_src = measured_peak_gbs()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # ---- workload: one job, every rank generates it (deterministic), rank r works on its ranges ---------------------------
    V, F, proj, cls = c4_workload(fp)
    mesh = fp.TriMesh(ctx, V, F)
    t0 = time.perf_counter(); mesh.build_aabb_tree(); tree_build_s = time.perf_counter() - t0
    sets = [proj, cls]
    Q = sum(len(s) for s in sets)
    ranges = [shard_range(len(s), rank, world) for s in sets]
    nmax = [max(shard_range(len(s), r, world)[1] - shard_range(len(s), r, world)[0] for r in range(world)) for s in sets]
    my = [np.ascontiguousarray(s[lo:hi]) for s, (lo, hi) in zip(sets, ranges)]
    dP = [torch.from_numpy(p).to(dev) for p in my]
    # per set one packed result block [n, 8] f64 would need a repack kernel; four plain arrays, padded to the largest slice
    dS = [torch.zeros(n, dtype=torch.float64, device=dev) for n in nmax]; dI = [torch.zeros(n, dtype=torch.int32, device=dev) for n in nmax]
    dC = [torch.zeros(n, 3, dtype=torch.float64, device=dev) for n in nmax]; dN = [torch.zeros(n, 3, dtype=torch.float64, device=dev) for n in nmax]
    if world > 1:
        gS = [torch.empty(world * n, dtype=torch.float64, device=dev) for n in nmax]; gI = [torch.empty(world * n, dtype=torch.int32, device=dev) for n in nmax]
        gC = [torch.empty(world * n, 3, dtype=torch.float64, device=dev) for n in nmax]; gN = [torch.empty(world * n, 3, dtype=torch.float64, device=dev) for n in nmax]

    def step_dev():
        # the large set first: its results travel (NCCL's own stream, async_op) while the small set is computed
        works = []
        for k in (1, 0):
            n = len(my[k])
            mesh.signed_distance_dev(dP[k].data_ptr(), n, dS[k].data_ptr(), dI[k].data_ptr(), dC[k].data_ptr(), dN[k].data_ptr(), stream.cuda_stream)
            if world > 1:   # the job's result lands in every rank's HBM (rank 0's included): NCCL all-gather of the padded slices
                works += [dist.all_gather_into_tensor(g, d, async_op=True) for g, d in ((gS[k], dS[k]), (gI[k], dI[k]), (gC[k], dC[k]), (gN[k], dN[k]))]
        for w in works:
            w.wait()            # stream-side wait: the step's closing event comes after the gathers

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    l0 = ctx.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for k in range(args.steps):
        step_dev()
        evs[k + 1].record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    k1_ms_step = 2.0 * ctx.query_kernel_ms(min(2 * args.steps, 32))   # CUDA events around the dominant kernel; two launches per step
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms_local = total_ms / args.steps
    t = torch.tensor([total_ms, k1_ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0].item()) / args.steps
    k1_ms_step = float(t[1].item())
    value = Q / (ms_per_step * 1e-3)
    if world > 1:   # the gathered job result equals the single-rank layout: spot check on rank 0 (own slice inside the gathered block)
        lo, hi = ranges[0]
        assert torch.equal(gI[0][rank * nmax[0]: rank * nmax[0] + (hi - lo)], dI[0][: hi - lo])

    # ---- e2e through the host-pointer C-ABI call, pinned host buffers (rank-local slices) ------------------------------------
    import ctypes as C
    hP = [torch.from_numpy(p).pin_memory() for p in my]
    hS = [torch.empty(len(p), dtype=torch.float64).pin_memory() for p in my]; hI = [torch.empty(len(p), dtype=torch.int32).pin_memory() for p in my]
    hC = [torch.empty(len(p), 3, dtype=torch.float64).pin_memory() for p in my]; hN = [torch.empty(len(p), 3, dtype=torch.float64).pin_memory() for p in my]

    def step_e2e():
        for k in range(2):
            rc = fp.lib().fpohm_signed_distance(ctx.h, mesh.h, C.c_void_p(hP[k].data_ptr()), C.c_int64(len(my[k])), C.c_void_p(hS[k].data_ptr()),
                                                C.c_void_p(hI[k].data_ptr()), C.c_void_p(hC[k].data_ptr()), C.c_void_p(hN[k].data_ptr()))
            assert rc == 0, fp.lib().fpohm_last_error()
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, args.steps // 2)
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = Q / float(t.item())
    for k in range(2):      # both paths produce the same bits
        n = len(my[k])
        assert torch.equal(hS[k].to(dev), dS[k][:n]) and torch.equal(hI[k].to(dev), dI[k][:n])
    # cold call: upload + build of every query structure + the same step (a fresh mesh handle)
    e2e_cold_s = None
    if rank == 0:
        try:
            t0 = time.perf_counter()
            m2 = fp.TriMesh(ctx, V, F)
            for k in range(2):
                rc = fp.lib().fpohm_signed_distance(ctx.h, m2.h, C.c_void_p(hP[k].data_ptr()), C.c_int64(len(my[k])), C.c_void_p(hS[k].data_ptr()),
                                                    C.c_void_p(hI[k].data_ptr()), C.c_void_p(hC[k].data_ptr()), C.c_void_p(hN[k].data_ptr()))
                assert rc == 0
            e2e_cold_s = time.perf_counter() - t0
            m2.close()
        except Exception:
            e2e_cold_s = None

    also = {}
    if not args.no_also:
        # ---- C4: scaled Jacobian over the 216^3 warped block (resident), hexes sharded by range ------------------------------------
        try:
            Vb, Hb = fp.procedural.warped_hex_block(216)
            lo, hi = shard_range(len(Hb), rank, world)
            dVb = torch.from_numpy(Vb).to(dev); dHb = torch.from_numpy(np.ascontiguousarray(Hb[lo:hi]).view(np.int32)).to(dev)
            nH, nVb = hi - lo, len(Vb)
            dVJ = torch.empty(8 * nH, dtype=torch.float64, device=dev); dHJ = torch.empty(nH, dtype=torch.float64, device=dev)
            dst = torch.empty(3, dtype=torch.float64, device=dev); dfl = torch.empty(1, dtype=torch.int64, device=dev)
            jac_ms = timed(lambda: fp.scaled_jacobian_dev(ctx, dVb.data_ptr(), nVb, dHb.data_ptr(), nH, dVJ.data_ptr(), dHJ.data_ptr(), dst.data_ptr(),
                                                          dfl.data_ptr(), stream.cuda_stream), 20, 3)
            tj = torch.tensor([jac_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tj, op=dist.ReduceOp.MAX)
            jac_ms = float(tj.item())
            jb = BYTES_PER_HEX * nH + 24 * nVb
            also["c4_jacobian"] = {"hexes": int(len(Hb)), "hexes_per_rank": int(nH), "ms": jac_ms, "hexes_per_s": len(Hb) / (jac_ms * 1e-3),
                                   "roofline": {"bound": "hbm", "achieved": jb / (jac_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": jb / (jac_ms * 1e-3) / 1e9 / peak,
                                                "bytes": "104 B/hex + 24 B/vertex (per rank)"}}
            del dVb, dHb, dVJ, dHJ, Vb, Hb
        except Exception as e:
            also["c4_jacobian"] = {"error": str(e)}
        # ---- C3: 1024^3 z-ray parity voxelization (resident output), dense predicate occupancy, `--e 10` octree build --------------------
        try:
            g3 = fp.VoxelGrid(V.min(0), V.max(0) - V.min(0), 1.0 / 1024, 0)
            d3 = torch.empty(g3.num_voxels(), dtype=torch.uint8, device=dev)
            v3_ms = timed(lambda: fp.voxel_sign_dev(ctx, mesh, g3, d3.data_ptr(), stream.cuda_stream), 5, 2)
            v3_bytes = g3.num_voxels() + 72 * len(F)
            del d3
            occ_ms = None
            try:    # dense predicate occupancy of the same grid (host output buffer: only the library's CUDA-event kernel time is reported)
                occ = []
                for _ in range(4):      # 1 warm-up + 3: median (a single sample after the 1 GB pageable download of the previous call wobbles by 15 %)
                    fp.voxel_occupancy(ctx, mesh, g3)
                    occ.append(ctx.last_kernel_ms())
                occ_ms = float(np.median(occ[1:]))
            except Exception:
                pass
            p3 = fp.octree_grid_setup(V, 1 << 20); p3.c.stop_extent = 1 << 10
            ts3 = []
            for _ in range(3 + 8):        # 3 warm-up builds (the stream-ordered pool grows to the job's size), then 8 timed
                ctx.sync(); t0 = time.perf_counter(); o3 = fp.Octree.build(ctx, mesh, p3); ctx.sync(); ts3.append((time.perf_counter() - t0) * 1e3)
                sz3 = o3.sizes(); o3.close()
            tb = ts3[3:]
            o1_bytes = 70 * sz3["cells"] + 36 * sz3["nodes"]
            also["c3_1024"] = {"tris": int(len(F)), "voxel_sign_1024_ms": v3_ms, "voxel_sign_dims": g3.dims.tolist(),
                               "voxel_sign_roofline": {"bound": "hbm", "achieved": v3_bytes / (v3_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                       "frac": v3_bytes / (v3_ms * 1e-3) / 1e9 / peak},
                               "voxel_occupancy_1024_kernel_ms": occ_ms,
                               "voxel_occupancy_roofline_frac": (v3_bytes / (occ_ms * 1e-3) / 1e9 / peak) if occ_ms else None,
                               "octree_e10_build_ms": float(np.median(tb)), "octree_e10_build_ms_max": float(np.max(tb)), "octree_e10_build_ms_all": [round(x, 1) for x in ts3],
                               "octree_e10_cells": int(sz3["cells"]), "octree_e10_leaves": int(sz3["leaves"]), "octree_e10_nodes": int(sz3["nodes"]),
                               "octree_roofline": {"bound": "hbm", "achieved": o1_bytes / (np.median(tb) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                   "frac": o1_bytes / (np.median(tb) * 1e-3) / 1e9 / peak, "bytes": "70 B/cell + 36 B/node (SURVEY O1)"}}
        except Exception as e:
            also["c3_1024"] = {"error": str(e)}
        # ---- C5: metro Hausdorff, 50 M samples: boundary surface of the inside hexes of the C3 octree (--e 12) vs the C3 mesh -----------
        if rank == 0:
            try:
                p5 = fp.octree_grid_setup(V, 1 << 20); p5.c.stop_extent = 1 << 12
                o5 = fp.Octree.build(ctx, mesh, p5)
                Vh5, H5, _ = o5.hexes(); o5.close()
                S5 = mesh.signed_distance_pseudonormal(np.ascontiguousarray(Vh5[H5.astype(np.int64)].mean(1)), want=("S",))[0]
                keep = np.ascontiguousarray(H5[S5 < 0])
                conn = fp.HexConnectivity(ctx, keep, len(Vh5))
                bq = conn.F_vs[conn.F_boundary != 0].astype(np.int64)        # boundary quads, (0,1,2),(2,3,0) as ghm.cpp:4257-4272
                del conn
                used, tri5 = np.unique(np.concatenate([bq[:, [0, 1, 2]], bq[:, [2, 3, 0]]]), return_inverse=True)
                sur = {"V": np.ascontiguousarray(Vh5[used]), "F_vs": tri5.reshape(-1, 3).astype(np.int32)}
                B = fp.TriMesh(ctx, sur["V"], sur["F_vs"].astype(np.int32))
                B.build_aabb_tree()
                nA, nB = len(np.unique(F)), len(sur["V"])
                extra = 25_000_000 - max(nA, nB)
                hd = None
                ts5 = []
                for _ in range(3):
                    t0 = time.perf_counter(); hd = fp.hausdorff(ctx, mesh, B, extra_face_samples=extra); ts5.append((time.perf_counter() - t0) * 1e3)
                n5 = hd["n_ab"] + hd["n_ba"]
                k5 = ctx.last_kernel_ms()
                also["c5_hausdorff"] = {"samples": int(n5), "n_ab": hd["n_ab"], "n_ba": hd["n_ba"], "surface_tris": int(len(sur["F_vs"])), "wall_ms": float(min(ts5)),
                                        "device_ms": k5, "samples_per_s": n5 / (min(ts5) * 1e-3), "max": hd["max"], "mean": hd["mean"], "ratio": hd["ratio"],
                                        "sampling": "all referenced vertices + vcg similar-triangle face samples (sampling.h:496-540, sequential carry reproduced), 25 M per direction",
                                        "roofline": {"bound": "hbm", "achieved": BYTES_PER_SAMPLE * n5 / (k5 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                     "frac": BYTES_PER_SAMPLE * n5 / (k5 * 1e-3) / 1e9 / peak, "bytes": "32 B/sample materialised (SURVEY H1)",
                                                     "note": "tree search: latency / issue bound, see DESIGN.md"}}
                B.close(); del sur, Vh5, H5
            except Exception as e:
                also["c5_hausdorff"] = {"error": str(e)}
        # ---- C2 (round-1 headline, kept for continuity): gear 199 680 tris, octree depth 8, leaf-centre queries x4 ---------------
        try:
            gV, gF, _ = fp.procedural.gear()
            gm = fp.TriMesh(ctx, gV, gF)
            gp = fp.octree_grid_setup(gV, 1 << 20); gp.c.stop_extent = 1 << STOP_E
            t0 = time.perf_counter(); gm.build_aabb_tree(); g_tree_s = time.perf_counter() - t0
            go = fp.Octree.build(ctx, gm, gp)
            gVh, gH, _ = go.hexes(); gsz = go.sizes()
            gb = []
            for i in range(3 + 8):
                ctx.sync(); t0 = time.perf_counter(); o2 = fp.Octree.build(ctx, gm, gp); ctx.sync(); gb.append((time.perf_counter() - t0) * 1e3); o2.close()
            gext = gVh[gH[:, 1].astype(np.int64), 0] - gVh[gH[:, 0].astype(np.int64), 0]
            gP = make_queries(gVh, gH, gext); gQ = len(gP)
            dgP = torch.from_numpy(gP).to(dev)
            dgS = torch.empty(gQ, dtype=torch.float64, device=dev); dgI = torch.empty(gQ, dtype=torch.int32, device=dev)
            dgC = torch.empty(gQ, 3, dtype=torch.float64, device=dev); dgN = torch.empty(gQ, 3, dtype=torch.float64, device=dev)
            g_ms = timed(lambda: gm.signed_distance_dev(dgP.data_ptr(), gQ, dgS.data_ptr(), dgI.data_ptr(), dgC.data_ptr(), dgN.data_ptr(), stream.cuda_stream), 10, 3)
            also["c2_gear"] = {"tris": int(len(gF)), "cells": int(gsz["cells"]), "leaves": int(gsz["leaves"]), "octree_build_ms": float(np.median(gb[3:])),
                               "octree_build_ms_max": float(np.max(gb[3:])), "queries": int(gQ), "query_ms": g_ms, "queries_per_s": gQ / (g_ms * 1e-3),
                               "query_tree_build_s": g_tree_s, "round1_queries_per_s": 794.0e6}
            go.close(); gm.close(); del dgP, dgS, dgI, dgC, dgN
        except Exception as e:
            also["c2_gear"] = {"error": str(e)}
        # ---- z-slab sharded octree build on the C3 mesh (N > 1) ---------------------------------------------------------------------
        if world > 1:
            try:
                from fpohm_b200 import sharding
                comm = sharding.TorchComm()
                sh = {}
                for e in (10, 11):
                    prm_e = fp.octree_grid_setup(V, 1 << 20); prm_e.c.stop_extent = 1 << e
                    single, multi = [], []
                    for i in range(2 + 3):
                        barrier(); t0 = time.perf_counter()
                        o1 = fp.Octree.build(ctx, mesh, prm_e)
                        barrier(); single.append((time.perf_counter() - t0) * 1e3)
                        st = {}
                        barrier(); t0 = time.perf_counter()
                        oN = sharding.build_octree_sharded(fp, ctx, mesh, prm_e, comm, device=dev, stats=st)
                        barrier(); multi.append((time.perf_counter() - t0) * 1e3)
                        same = o1.sizes() == oN.sizes()
                        if i == 0 and e == 11:
                            a, b = o1.export(), oN.export()
                            same = same and all(np.array_equal(a[k], b[k]) for k in ("node_pos", "node_neigh", "first_child", "corner", "neigh"))
                        assert same, "sharded octree differs from the single-GPU octree"
                        cells = o1.sizes()["cells"]
                        o1.close(); oN.close()
                    sh[f"e{e}"] = {"cells": int(cells), "single_gpu_ms": float(np.median(single[2:])), "sharded_ms": float(np.median(multi[2:])),
                                   "replicated_levels": st.get("replicated_levels"), "phase_ms": st.get("phase_ms"), "bit_identical": True}
                also["c3_octree_zslab_sharded"] = sh
            except Exception as e:
                also["c3_octree_zslab_sharded"] = {"error": str(e)}
    sampler.stop_flag = True; sampler.join(timeout=2)

    if rank == 0:
        achieved = BYTES_PER_QUERY * (Q / world) / (k1_ms_step * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tf = ROOT / "profiles" / "r02_closest_point_traffic.json"
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                traffic, traffic_src = tj.get("dram_bytes_per_step"), tj.get("source")
            except Exception:
                traffic = None
        line = {"metric": "closest_point_queries_per_s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "queries_per_step": int(Q), "queries_this_rank": int(sum(len(p) for p in my)), "tris": int(len(F)),
                           "l2_policy": "inputs+outputs per step (%.0f MB) and the query structures (480 MB) larger than L2 (126 MB)" % (BYTES_PER_QUERY * Q / 1e6),
                           "sharding": "strong: rank r takes shard_range of each query set of the one job; mesh + trees replicated; results NCCL all-gathered into "
                                       "every rank's HBM inside the timed region" if world > 1 else "single GPU",
                           "numa_node_bound": numa},
                "roofline": {"bound": "hbm", "kernel": "cp_pair_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                             "kernel_ms_per_step": k1_ms_step, "kernel_share_of_step": k1_ms_step / step_ms_local,
                             "issue_slots_used": {"frac": 0.70, "source": "ncu smsp__issue_active.avg.pct_of_peak_sustained_active of this kernel on the "
                                                  "classification launch (not measured live): profiles/r02_ncu_summary.md section 1"},
                             "note": "84 B/query algorithmic x this rank's queries over the packet kernel's own duration (CUDA events inside the library, two "
                                     "launches per step); a tree search is issue/latency bound, not HBM bound: see profiles/r02_ncu_summary.md"},
                "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": 24 * Q, "d2h_bytes_per_step": 60 * Q,
                        "note": "host-pointer C-ABI call per query set, pinned rank-local buffers" + ("; each rank moves its own slice" if world > 1 else "")},
                "e2e_cold": {"seconds": e2e_cold_s, "queries_per_s": (Q / world / e2e_cold_s) if e2e_cold_s else None,
                             "note": "fresh mesh handle: upload + igl-identical tree, normals and wide tree built on the device (host std::sort only for axes with tied barycentres, host acos) + one step (rank 0's slice)"},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "also": also}
        line["also"]["query_tree_build_s"] = tree_build_s
        if not args.no_cpu_baseline and world == 1:
            try:
                res = (np.concatenate([hS[0].numpy(), hS[1].numpy()]), np.concatenate([hI[0].numpy(), hI[1].numpy()]),
                       np.concatenate([hC[0].numpy(), hC[1].numpy()]), np.concatenate([hN[0].numpy(), hN[1].numpy()]))
                line["cpu_baseline"], line["parity"] = cpu_baseline_and_parity(V, F, np.concatenate(sets), res)
            except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 1, "kind": "reference", "sample": f"unavailable: {e}"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

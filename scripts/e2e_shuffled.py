"""Host-pointer query call on a SHUFFLED copy of the bench queries: auto Morton order (default) vs FPOHM_CP_NOSORT=1."""
import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
import bench
ctx = fp.Context(0)
V, F = bench.workload(fp)
mesh = fp.TriMesh(ctx, V, F)
prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << bench.STOP_E
mesh.build_aabb_tree()
o = fp.Octree.build(ctx, mesh, prm)
Vh, H, _ = o.hexes()
ext = (Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0])
P0 = bench.make_queries(Vh, H, ext)
for name, P in (("bench order", P0), ("shuffled", np.ascontiguousarray(P0[np.random.default_rng(0).permutation(len(P0))]))):
    Q = len(P)
    hP = torch.from_numpy(P).pin_memory()
    hS = torch.empty(Q, dtype=torch.float64).pin_memory(); hI = torch.empty(Q, dtype=torch.int32).pin_memory()
    hC = torch.empty(Q, 3, dtype=torch.float64).pin_memory(); hN = torch.empty(Q, 3, dtype=torch.float64).pin_memory()
    def step():
        rc = fp.lib().fpohm_signed_distance(ctx.h, mesh.h, C.c_void_p(hP.data_ptr()), C.c_int64(Q), C.c_void_p(hS.data_ptr()),
                                            C.c_void_p(hI.data_ptr()), C.c_void_p(hC.data_ptr()), C.c_void_p(hN.data_ptr()))
        assert rc == 0, fp.lib().fpohm_last_error()
    for _ in range(2): step()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 4
    print(f"{name:12s} nosort={os.environ.get('FPOHM_CP_NOSORT','0')} e2e {dt*1e3:.3f} ms  {Q/dt/1e6:.1f} Mq/s", flush=True)
    np.save(f"/tmp/e2e_{name.split()[0]}_{os.environ.get('FPOHM_CP_NOSORT','0')}.npy", np.concatenate([hS.numpy(), hI.numpy().astype(np.float64), hC.numpy().reshape(-1), hN.numpy().reshape(-1)]))

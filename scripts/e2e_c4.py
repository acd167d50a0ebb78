"""e2e (host-pointer C-ABI call, pinned buffers) of the C4 step on one GPU, per query set; env switches are read by the library
(FPOHM_CP_CHUNK_SORT = chunk size of the sorted pipeline).  argv[1] = 'shuffle' also times the classification set in random order."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
mesh = fp.TriMesh(ctx, V, F); mesh.build_aabb_tree()
sets = dict(zip(("project", "classify"), fp.procedural.c4_queries(V, F)))
if len(sys.argv) > 1 and sys.argv[1] == "shuffle":
    sets["classify shuffled"] = np.ascontiguousarray(sets["classify"][np.random.default_rng(3).permutation(len(sets["classify"]))])
tot = 0.0
for name, P in sets.items():
    Q = len(P)
    hP = torch.from_numpy(P).pin_memory()
    hS = torch.empty(Q, dtype=torch.float64).pin_memory(); hI = torch.empty(Q, dtype=torch.int32).pin_memory()
    hC = torch.empty(Q, 3, dtype=torch.float64).pin_memory(); hN = torch.empty(Q, 3, dtype=torch.float64).pin_memory()
    def step():
        rc = fp.lib().fpohm_signed_distance(ctx.h, mesh.h, C.c_void_p(hP.data_ptr()), C.c_int64(Q), C.c_void_p(hS.data_ptr()),
                                            C.c_void_p(hI.data_ptr()), C.c_void_p(hC.data_ptr()), C.c_void_p(hN.data_ptr()))
        assert rc == 0
    for _ in range(2): step()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    if "shuffled" not in name: tot += dt
    print(f"{name:18s} n={Q:9d} e2e {dt*1e3:8.3f} ms  {Q/dt/1e6:7.1f} Mq/s  checksum I {int(hI.sum())} S {float(hS.sum()):.12e}", flush=True)
print(f"C4 step e2e {tot*1e3:.2f} ms = {11857634/tot/1e6:.1f} Mq/s")

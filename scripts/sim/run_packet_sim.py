"""Design-space study for the closest-point kernel (CPU only; uses the oracle to get the bench queries)."""
import ctypes as C, sys, numpy as np
sys.path.insert(0, ".")
import fpohm_b200 as fp
from oracle import ref_oracle as R
import bench
V, F = bench.workload(fp)
gs, org, mt, vs = R.octree_grid_setup(V, F, 1 << 20)
ro = R.RefOctree.build(V, F, gs, org, mt, vs, 1 << bench.STOP_E)
Vh, H, _ = ro.hexes()
ext = (Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0])
P = bench.make_queries(Vh, H, ext)
box, prim, lr = fp.host_igl_tree(V, F)
tri = np.ascontiguousarray(V[F].reshape(-1, 9))
lib = C.CDLL("/tmp/packet_sim.so")
p = lambda a: a.ctypes.data_as(C.c_void_p)
n = 32 * 4000
start = (len(P) // 2) // 32 * 32
for name, Q in (("bench order", P[start:start + n]), ("random order", P[np.random.default_rng(0).permutation(len(P))[:n]])):
    Q = np.ascontiguousarray(Q)
    for W in (32, 8):
        out = np.zeros(8)
        lib.simulate(p(box), p(prim), p(lr), p(tri), p(Q), C.c_int64(len(Q)), C.c_int(W), p(out))
        g = len(Q) // W
        print(f"{name:13s} W={W:2d}: single nodes/q {out[0]/len(Q):6.1f} leaves/q {out[1]/len(Q):5.1f} max-lane nodes/warp {out[6]/g:6.1f} | "
              f"packet nodes/warp {out[2]/g:7.1f} (want {out[5]/max(out[2],1):4.1f}/{W}) leaf steps/warp {out[3]/g:6.1f} (active {out[4]/max(out[3],1):4.1f}/{W})")

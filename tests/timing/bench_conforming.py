"""conforming_mesh timing on the bench octree (gear, depth 8) and a deeper one, device kernels only (KernelTimer)."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F, _ = fp.procedural.gear()
m = fp.TriMesh(ctx, V, F)
for e in (12, 11, 10):
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << e
    o = fp.Octree.build(ctx, m, prm)
    H = o.hexes()[1]
    for rep in range(3):
        tm = {}
        t = time.perf_counter(); hy = fp.conforming_mesh(ctx, o, H, keep_timing=tm); wall = time.perf_counter() - t
    by = 4 * (hy["F_vs"].size * 2 + hy["H_fs"].size + hy["H_vs"].size + hy["E_vs"].size + hy["F_nhs"].size) + 8 * (hy["nF"] * 2 + hy["nH"] * 2)
    print(f"e={e}: hexes {len(H)} faces {hy['nF']} replaced {hy['n_replaced']} edges {hy['nE']}  hex connectivity {tm['connectivity_ms']:.2f} ms  "
          f"conforming {tm['conforming_ms']:.2f} ms ({len(H)/tm['conforming_ms']/1e3:.1f} M hexes/s, {by/tm['conforming_ms']/1e6:.0f} GB/s of output)  wall incl. export {wall*1e3:.0f} ms")
    if e == 12:
        sys.path.insert(0, "."); 
        try:
            from oracle import ref_oracle as R
            ex = o.export(); Vp = o.hexes()[0]
            t = time.perf_counter(); R.conforming_mesh_tables(ex["node_pos"], ex["node_neigh"], Vp, H, prm.grid_size); print(f"   reference (build_connectivity + conforming_mesh, 1 thread): {(time.perf_counter()-t)*1e3:.0f} ms")
        except Exception as ex_:
            print("   reference unavailable:", ex_)

print("--- conforming + dual ---")
for e in (12, 11):
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << e
    o = fp.Octree.build(ctx, m, prm)
    for rep in range(3):
        tm = {}
        hyb, d = fp.conforming_and_dual(ctx, o, keep_timing=tm)
    print(f"e={e}: conforming {tm['conforming_ms']:.2f} ms, dual {tm['dual_ms']:.2f} ms -> dual cells {d['nH']} faces {d['nF']} edges {d['nE']} census {d['census'].tolist()}")
    if e == 12:
        try:
            ex = o.export(); Vp, H, _ = o.hexes()
            t = time.perf_counter(); R.conforming_and_dual_tables(ex["node_pos"], ex["node_neigh"], Vp, H, prm.grid_size)
            print(f"   reference (build_connectivity + conforming_mesh + dual_conforming_mesh, 1 thread): {(time.perf_counter()-t)*1e3:.0f} ms")
        except Exception as ex_:
            print("   reference unavailable:", ex_)

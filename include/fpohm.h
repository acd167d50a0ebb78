/*
 * fpohm.h — C-ABI of the B200-native geometry core (libfpohm.so).
 *
 * Drop-in boundary for the data-parallel hot path of
 * gaoxifeng/Feature-Preserving-Octree-Hex-Meshing.  The reference has no FFI layer (plain C++
 * headers), so every entry point below names the reference function(s) whose body it replaces
 * (paths relative to the reference root; ghm.cpp = grid_meshing/grid_hex_meshing.cpp,
 * gf.cpp = global_functions.cpp, igl/ = extern/libigl/include/igl/).  The C++ shim in
 * `feature-preserving-octree-hex-meshing_b200/host/fpohm_shim.hpp` re-creates the reference
 * signatures on top of this header; INTEGRATION.md shows how a maintainer wires it in.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FPOHM_E* code otherwise; it never throws,
 *     aborts or prints.  fpohm_last_error() returns the message of the calling thread's last failure.
 *   - plain pointers + sizes only.  Unless a name ends in `_dev`, pointers are HOST pointers and the
 *     call performs its own H2D/D2H copies (this is what a drop-in caller uses).  `_dev` variants take
 *     device pointers plus a cudaStream_t (passed as void*) and never synchronise: they are what a
 *     caller with HBM-resident data (and bench.py's device-timed `value`) uses.
 *   - vertex arrays are "xyz per vertex", i.e. the memory of the reference's `Mesh::V` (3 x n,
 *     column-major Eigen, global_types.h:516) or of a row-major n x 3 array — they are the same bytes.
 *   - there is NO CPU fallback: every compute entry point fails with FPOHM_ENODEV without a CUDA device.
 */
#ifndef FPOHM_H
#define FPOHM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPOHM_OK        0
#define FPOHM_EINVAL   -1   /* bad argument */
#define FPOHM_ENODEV   -2   /* no CUDA device / wrong architecture */
#define FPOHM_ECUDA    -3   /* CUDA runtime error (message in fpohm_last_error) */
#define FPOHM_ENOMEM   -4
#define FPOHM_ERANGE   -5   /* integer range exceeded (e.g. > 2^21 finest cells per axis after shift) */
#define FPOHM_ESTATE   -6   /* object not in the state the call needs */

typedef struct fpohm_ctx    fpohm_ctx;     /* one per GPU: device id, streams, scratch arena            */
typedef struct fpohm_mesh   fpohm_mesh;    /* device-resident triangle soup + facet bboxes (+ trees)     */
typedef struct fpohm_octree fpohm_octree;  /* device-resident graded/paired octree                       */
typedef struct fpohm_surface fpohm_surface; /* result of fpohm_extract_surface                          */
typedef struct fpohm_conn   fpohm_conn;    /* result of fpohm_hex_connectivity                           */

const char *fpohm_last_error(void);
const char *fpohm_version(void);
/* number of visible CUDA devices (0 when none); never fails */
int fpohm_device_count(void);

/* ------------------------------------------------------------------------------------------------ */
/* context                                                                                            */
int  fpohm_ctx_create(int device, fpohm_ctx **out);
void fpohm_ctx_destroy(fpohm_ctx *ctx);
int  fpohm_ctx_sync(fpohm_ctx *ctx);
/* Device memory of a context lives in its own arena (big cudaMalloc chunks, reused across calls; the device's default
 * stream-ordered pool is not touched).  _trim returns the chunks nothing lives in to the driver, _memory reports bytes. */
int  fpohm_ctx_trim(fpohm_ctx *ctx, int64_t *bytes_released);
int  fpohm_ctx_memory(fpohm_ctx *ctx, int64_t *bytes_reserved, int64_t *bytes_in_use);
/* milliseconds of the kernels launched by the last host-pointer call on this ctx (CUDA events) */
int  fpohm_ctx_last_kernel_ms(fpohm_ctx *ctx, double *ms);
/* mean duration (CUDA events on the launching stream) of the dominant query kernel — the packet walk — over the last
 * `last_n` (<= 32) query launches on this ctx; synchronises on those launches.  bench.py's roofline numerator. */
int  fpohm_ctx_query_kernel_ms(fpohm_ctx *ctx, int32_t last_n, double *mean_ms);
/* number of kernels this library launched on ctx since creation (bench.py "gpu_launches") */
int  fpohm_ctx_launch_count(fpohm_ctx *ctx, int64_t *n);

/* ------------------------------------------------------------------------------------------------ */
/* triangle mesh (the reference's GEO::Mesh M_i / Mesh mf.tri)                                        */
int  fpohm_mesh_upload(fpohm_ctx *ctx, const double *V, int64_t nV, const int32_t *F, int64_t nF, fpohm_mesh **out);
void fpohm_mesh_free(fpohm_mesh *mesh);
/* Same as fpohm_mesh_upload, but keyed on the CONTENT of (V, F): a surface the context has seen before comes back with its
 * trees and normals already built (the reference's points_inside_mesh rebuilds an igl::AABB on every call, gf.cpp:4038, and the
 * outer loop passes the same surface again and again).  Handles are reference counted — release them with fpohm_mesh_free as
 * usual; the context keeps up to 8 surfaces (least recently used unreferenced one evicted) until fpohm_ctx_mesh_cache_clear /
 * fpohm_ctx_destroy. */
int  fpohm_mesh_upload_cached(fpohm_ctx *ctx, const double *V, int64_t nV, const int32_t *F, int64_t nF, fpohm_mesh **out);
int  fpohm_ctx_mesh_cache_clear(fpohm_ctx *ctx);

/* ------------------------------------------------------------------------------------------------ */
/* octree (OctreeGrid, grid_meshing/octree.h:62-270, octree.cpp; octree_mesh, ghm.cpp:460-567)        */
typedef struct fpohm_octree_params {
	int32_t grid_size[3];      /* finest-cell grid, each a power of two (octree.cpp:26-28)               */
	double  origin[3];         /* ghm.cpp:484                                                            */
	double  mesh_transform[3]; /* ghm.cpp:493                                                            */
	double  voxel_size;        /* ghm.cpp:467-470                                                        */
	int32_t stop_extent;       /* cells with extent <= stop_extent are never split (ghm.cpp:503)        */
	int32_t graded;            /* 2:1 over faces and edges (octree.cpp:598-627)                          */
	int32_t paired;            /* sibling/root pairing (octree.cpp:574-590,632-643)                      */
	int32_t reserved;
} fpohm_octree_params;

/* ghm.cpp:463-493: bbox -> voxel_size, grid_size = next_pow2(ceil(extent/voxel_size)), origin, mesh_transform.
 * Pure host arithmetic in the reference's exact expression order; fills p->grid_size/origin/mesh_transform/voxel_size. */
int fpohm_octree_grid_setup(const double *V, int64_t nV, int32_t num_voxels, fpohm_octree_params *p);

/* OctreeGrid_initialize + subdivide(should_subdivide, graded, paired): octree.cpp:39-117,648-690 with the
 * bbox-overlap predicate of ghm.cpp:502-517 (== voxelization.cpp:367-380 when mesh_transform = 0). */
int  fpohm_octree_build(fpohm_ctx *ctx, const fpohm_mesh *mesh, const fpohm_octree_params *p, fpohm_octree **out);
/* Same closure, but the set of cells whose predicate is true is given explicitly as (x,y,z,extent) rows:
 * the floating-point-free form of subdivide(), used for predicate-independent topology parity
 * (and the equivalent of OctreeGrid::testSubdivideRandom, octree.cpp:900-961). */
int  fpohm_octree_build_from_marks(fpohm_ctx *ctx, const int32_t grid_size[3], const int32_t *marks, int64_t n_marks,
                                   int32_t graded, int32_t paired, fpohm_octree **out);
/* subdivide(pred, graded, paired) on an EXISTING octree with a (smaller) stop_extent — what octree_mesh does on the
 * later passes of the pipeline's outer loop (ghm.cpp:495-500,523-524): every current leaf is tested, children of
 * predicate-true cells recursively.  The octree is re-numbered afterwards. */
int  fpohm_octree_subdivide(fpohm_octree *oct, const fpohm_mesh *mesh, int32_t stop_extent);
/* subdivide(pred, cells, graded, paired): octree.cpp:691-729 via ghm.cpp:518-521.  cell_ids index the cell
 * numbering of this octree (fpohm_octree_export); only listed LEAF cells are tested, no recursion into children.
 * The octree is re-numbered afterwards. */
int  fpohm_octree_refine(fpohm_octree *oct, const fpohm_mesh *mesh, const int32_t *cell_ids, int64_t n, int32_t stop_extent);
void fpohm_octree_free(fpohm_octree *oct);

/* ---- z-slab sharded octree build (multi-GPU).  The reference has no distributed form; these entry
 * points split fpohm_octree_build (OctreeGrid::subdivide, octree.cpp:648-690 + makeCellGraded :598-627 + pairing
 * :574-590) so that the 2:1 grading constraints crossing a slab face can be exchanged by the HOST's collective
 * (NCCL all-gather of Morton codes; this library links no communication layer).  Protocol, identical on every rank:
 *     shard_create -> shard_refine(&lmax) -> G = allreduce_max(lmax)
 *     for l = G .. 0:  level_outgoing(G, l, &n) ; outgoing_copy(buf) ; gathered = allgather(buf) ; level_close(l, gathered)
 *     for each l:      level_result(l, buf, &n) ; gathered[l] = allgather(buf)
 *     shard_finish(gathered, counts, &octree)       -- every rank holds the complete canonical octree
 * The result is bit-identical to fpohm_octree_build for every world size.  All *_dev pointers are device memory of
 * the shard's context; codes are per-level Morton codes (x bit 0, y bit 1, z bit 2 interleaved). */
typedef struct fpohm_octree_shard fpohm_octree_shard;
int  fpohm_octree_shard_create(fpohm_ctx *ctx, const fpohm_mesh *mesh, const fpohm_octree_params *p, int32_t rank, int32_t world,
                               fpohm_octree_shard **out);
void fpohm_octree_shard_free(fpohm_octree_shard *sh);
int  fpohm_octree_shard_refine(fpohm_octree_shard *sh, int32_t *local_max_level);
int  fpohm_octree_shard_info(const fpohm_octree_shard *sh, int32_t *replicated_levels, int32_t *slab_bounds, int64_t *owned_true_cells);
int  fpohm_octree_shard_level_outgoing(fpohm_octree_shard *sh, int32_t global_max_level, int32_t level, int64_t *n_out);
int  fpohm_octree_shard_outgoing_copy(fpohm_octree_shard *sh, uint64_t *dst_dev);
int  fpohm_octree_shard_level_close(fpohm_octree_shard *sh, int32_t level, const uint64_t *gathered_dev, int64_t n_gathered, int64_t *n_closed);
int  fpohm_octree_shard_level_result(const fpohm_octree_shard *sh, int32_t level, uint64_t *dst_dev, int64_t *n);
int  fpohm_octree_shard_finish(fpohm_octree_shard *sh, const uint64_t *const *gathered_dev, const int64_t *counts, fpohm_octree **out);


int  fpohm_octree_sizes(const fpohm_octree *oct, int64_t *n_nodes, int64_t *n_cells, int64_t *n_leaves,
                        int32_t *n_roots, int32_t *max_depth);
/* m_Nodes / m_Cells (octree.h:16-61,113-114) as SoA: node_pos 3/node, node_neigh 6/node (prev/next per axis),
 * cell_first_child 1/cell, cell_corner 8/cell (Cube::delta order, common.h:147), cell_neigh 6/cell.
 * Any pointer may be NULL.  Numbering is canonical (DESIGN.md "octree numbering"), not the reference's
 * split-order artefact; every structural invariant the pipeline relies on holds. */
int  fpohm_octree_export(const fpohm_octree *oct, int32_t *node_pos, int32_t *node_neigh, int32_t *cell_first_child,
                         int32_t *cell_corner, int32_t *cell_neigh);
/* octree_mesh export, ghm.cpp:527-562: Vpos[3*node] = mesh_transform + origin + nodePos*voxel_size;
 * hex[8*leaf + lv] = cellCornerId(leaf cell, lv) for leaves in cell order; hex2cell = hex2Octree_map. */
int  fpohm_octree_hexes(const fpohm_octree *oct, double *Vpos, uint32_t *hex, int32_t *hex2cell);
/* is2to1Graded / isPaired (octree.cpp:148-202) evaluated on the device; bit0 graded, bit1 paired */
int  fpohm_octree_check(const fpohm_octree *oct, int32_t *flags);
/* compute_sign(M, aabb, octree, origin, spacing), voxelization.cpp:101-163: z-ray parity per cell (all cells) */
int  fpohm_octree_cell_sign(const fpohm_octree *oct, const fpohm_mesh *mesh, const double origin[3], double spacing, float *inside);

/* ------------------------------------------------------------------------------------------------ */
/* dense voxel grid (VoxelGrid<num_t>, voxelization.h:41-91) + compute_sign (voxelization.h:220-272)  */
/* dims[d] = ceil(extent[d]/spacing) + 2*padding; origin_out = origin - padding*spacing (voxelization.h:71-82) */
int fpohm_voxel_grid_setup(const double origin[3], const double extent[3], double spacing, int32_t padding,
                           int32_t dims[3], double origin_out[3]);
/* out: dims[0]*dims[1]*dims[2] bytes, x fastest (voxelization.cpp:26-28); 1 = inside */
int fpohm_voxel_sign(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                     const int32_t dims[3], uint8_t *out);
/* device output on the caller's stream.  Returns once the hit stage has reported that every hit list fitted (the pass is repeated with
 * more room otherwise); the fill may still be running: out_dev is complete in stream order, like the output of any kernel on `stream`. */
int fpohm_voxel_sign_dev(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                         const int32_t dims[3], uint8_t *out_dev, void *stream);
/* z-slab of the same grid (multi-GPU sharding): writes layers [z_begin, z_end) only, the first at out_dev.
 * z_begin must be a multiple of 32 and z_end a multiple of 32 or dims[2]; concatenating the slabs is bit-identical to the
 * unsliced call (hits and parity are evaluated on the whole column, only the fill is sliced). */
int fpohm_voxel_sign_slab_dev(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                              const int32_t dims[3], int32_t z_begin, int32_t z_end, uint8_t *out_dev, void *stream);
/* subdivision predicate on a dense grid: out[x + nx*(y + ny*z)] = 1 iff some facet AABB overlaps the closed
 * cell box (geo/basic/geometry.h:612-622 with the box of voxelization.cpp:370-375, extent = 1) */
int fpohm_voxel_occupancy(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                          const int32_t dims[3], uint8_t *out);
/* DexelGrid<double> + compute_sign (voxelization.h:95-138,275-331).  Two-phase: call with values == NULL to get
 * *total; offsets has dims2[0]*dims2[1]+1 entries (x fastest). */
int fpohm_dexel_sign(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                     const int32_t dims2[2], int64_t *offsets, double *values, int64_t *total);

/* ------------------------------------------------------------------------------------------------ */
/* query layer                                                                                        */
/* build_aabb_tree (ghm.cpp:4231-4248): igl::AABB::init (igl/AABB.cpp:94-200, identical tree) + per_face /
 * per_vertex(ANGLE) / per_edge(UNIFORM) normals.  Idempotent; called lazily by the query entry points. */
int fpohm_mesh_build_query_tree(fpohm_ctx *ctx, fpohm_mesh *mesh);
/* Treestr normals (global_types.h:681-691).  n_edges via fpohm_mesh_num_edges.  Any pointer may be NULL. */
int fpohm_mesh_num_edges(const fpohm_mesh *mesh, int64_t *n_edges);
int fpohm_mesh_normals(const fpohm_mesh *mesh, double *FN, double *VN, double *EN, int32_t *E, int32_t *EMAP);
/* pre-order flattening of the igl-identical tree for inspection: box 6/node, primitive, left/right (-1 leaf) */
int fpohm_mesh_tree_nodes(const fpohm_mesh *mesh, int64_t *n_nodes);
int fpohm_mesh_tree_export(const fpohm_mesh *mesh, double *box, int32_t *prim, int32_t *lr);

/* The two builders above are host code (they must call the same std:: algorithms as igl); these
 * device-free variants expose them for CPU-side verification.  Two-phase: pass the capacity in *n_nodes / *n_edges
 * with output pointers, or NULL outputs to get the sizes. */
int fpohm_host_igl_tree(const double *V, int64_t nV, const int32_t *F, int64_t nF, int64_t *n_nodes,
                        double *box, int32_t *prim, int32_t *lr);
int fpohm_host_igl_normals(const double *V, int64_t nV, const int32_t *F, int64_t nF, int64_t *n_edges,
                           double *FN, double *VN, double *EN, int32_t *E, int32_t *EMAP);

/* igl::signed_distance_pseudonormal batch (igl/signed_distance.cpp:186-218): P np x 3;
 * S signed distance, I closest facet, C closest point np x 3, N pseudonormal np x 3.  Any output may be NULL.
 * Callers: points_inside_mesh gf.cpp:4024-4048 (S only), projection_smooth ghm.cpp:3760-3781,
 * dirty_graph_projection ghm.cpp:4034-4081, node_mapping ghm.cpp:2273, curve_mapping ghm.cpp:2603. */
int fpohm_signed_distance(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P, int64_t np,
                          double *S, int32_t *I, double *C, double *N);
int fpohm_signed_distance_dev(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P_dev, int64_t np,
                              double *S_dev, int32_t *I_dev, double *C_dev, double *N_dev, void *stream);
/* igl::point_mesh_squared_distance (igl/point_mesh_squared_distance.cpp:17-46; hausdorff_dis gf.cpp:3590-3604) */
int fpohm_point_mesh_sqdist(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P, int64_t np,
                            double *sqrD, int32_t *I, double *C);
/* hausdorff_dis(mesh0, mesh1, outlierVs, thr), gf.cpp:3590-3628: vertices of A vs B and back; outliers written to
 * outlier_vs (capacity nV of B), *n_outliers set; the threshold decays x0.9 until the list is non-empty. */
int fpohm_hausdorff_outliers(fpohm_ctx *ctx, fpohm_mesh *A, fpohm_mesh *B, double dis_threshold,
                             int32_t *outlier_vs, int64_t *n_outliers);

/* feature-curve projection, LINE branch of dirty_graph_projection (ghm.cpp:3967-3994) + point_line_projection
 * (gf.cpp:3454-3465).  Curves in CSR over vertex ids of Vc; circle[c] != 0 closes the polyline. */
int fpohm_polyline_project(fpohm_ctx *ctx, const double *Vc, int64_t nVc, const int64_t *curve_off, const int32_t *curve_vs,
                           const uint8_t *circle, int64_t n_curves, const double *P, const int32_t *curve_id, int64_t np,
                           double *origin_L, double *axis_L);

/* scaled_jacobian, Hex branch (gf.cpp:2309-2358) + a_jacobian (gf.cpp:2422-2442), table global_types.h:163-173.
 * V_Js 8/hex, H_Js 1/hex (either may be NULL), min_ave_dev = {min_Jacobian, ave_Jacobian, deviation_Jacobian}. */
int fpohm_scaled_jacobian(fpohm_ctx *ctx, const double *V, int64_t nV, const uint32_t *hex, int64_t H,
                          double *V_Js, double *H_Js, double min_ave_dev[3], int64_t *flipped);
int fpohm_scaled_jacobian_dev(fpohm_ctx *ctx, const double *V_dev, int64_t nV, const uint32_t *hex_dev, int64_t H,
                              double *V_Js_dev, double *H_Js_dev, double *min_ave_dev_dev /*3*/, int64_t *flipped_dev,
                              void *stream);

/* metro two-sided sampled Hausdorff, compute(...) x3 (metro_hausdorff.cpp:12,196,358) with the reference's flags
 * (vertex sampling only, metro_hausdorff.cpp:39-48).  out = {bbox_diag, max_ab, max_ba, mean_ab, mean_ba,
 * rms_ab, rms_ba}; n_samples = {n_ab, n_ba}.  extra_face_samples > 0 adds that many similar-triangle face
 * samples per direction (sampling.h:496-540 rule) — 0 reproduces the reference. */
int fpohm_hausdorff(fpohm_ctx *ctx, fpohm_mesh *A, fpohm_mesh *B, int64_t extra_face_samples,
                    double out[7], int64_t n_samples[2]);

/* The data-parallel head of clean_hex_mesh (grid_hex_meshing.cpp:1937-1951): per hex the centre of the bounding box of its 8
 * corners ((max + min) / 2 per axis), points_inside_mesh against `surface`, H_flag = signed_dis < 0.  Either output may be NULL. */
int  fpohm_classify_hexes(fpohm_ctx *ctx, fpohm_mesh *surface, const double *V, int64_t nV, const uint32_t *hex, int64_t H,
                          double *signed_dis, uint8_t *H_flag);

/* build_connectivity, Hex branch (gf.cpp:121-186) + adjacency (gf.cpp:226-264) */
int  fpohm_hex_connectivity(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, fpohm_conn **out);
int  fpohm_conn_sizes(const fpohm_conn *c, int64_t *nF, int64_t *nE);
/* F_vs 4/F, F_es 4/F, F_boundary 1/F, E_vs 2/E, E_boundary 1/E, V_boundary 1/V, H_fs 6/H; any may be NULL */
int  fpohm_conn_fixed(const fpohm_conn *c, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs,
                      uint8_t *E_boundary, uint8_t *V_boundary, uint32_t *H_fs);
/* CSR relations; which: 0 F.neighbor_hs 1 E.neighbor_fs 2 E.neighbor_hs 3 V.neighbor_vs 4 V.neighbor_es
 * 5 V.neighbor_fs 6 V.neighbor_hs.  Call with val == NULL for the total. */
int  fpohm_conn_csr(const fpohm_conn *c, int32_t which, int64_t *off, uint32_t *val, int64_t *total);
void fpohm_conn_free(fpohm_conn *c);

/* ---- clean_hex_mesh and its stages (grid_meshing/grid_hex_meshing.cpp:1932-2124).  H_flag is the
 * reference's md.H_flag as one byte per hex of the ENTIRE mesh (1 = inside / kept); stages update it in place and give the
 * flags the reference's sequential loops give (see csrc/cleaning.cu for why the data-parallel forms are equivalent).
 *
 * fpohm_reorder_hexes        reorder_hex_mesh (global_functions.cpp:2199-2229): a hex whose 8 corner determinants sum to a
 *                            negative volume gets its vertex list mirrored (3,2,1,0,7,6,5,4), in place.
 * fpohm_tag_uneven_elements  tagging_uneven_element (ghm.cpp:1983-2005) on the connectivity of the entire mesh.
 * fpohm_reindex_submesh      re_indexing_connectivity (global_functions.cpp:664-698): V_map (nV, -1 = dropped), V_map_reverse
 *                            and H_map_reverse (capacity nV / H, lengths returned), sub_hex = hexes of the sub-mesh in its
 *                            own vertex numbering (8 x n_sub_h, capacity 8 H).  Any output may be NULL.  The reference's
 *                            H_map stays empty (it is cleared and never filled).  The sub-mesh's connectivity is
 *                            fpohm_hex_connectivity(sub_hex).
 * fpohm_clean_non_manifold   clean_non_manifold_ve (ghm.cpp:2006-2080): the whole loop, re-indexing between rounds.
 * fpohm_drop_small_pieces    drop_small_pieces (ghm.cpp:2081-2124); n_pieces = number of face-connected pieces found.
 * fpohm_medial_surface_flags tail of clean_hex_mesh (ghm.cpp:1970-1981): Fs[].on_medial_surface / Vs[].on_medial_surface.
 * fpohm_clean_hex_mesh       the composition (ghm.cpp:1932-1981) with args.scaffold_type 1 (the default, no scaffold
 *                            layers): reorder (hex is rewritten) -> bbox centres + points_inside_mesh against `surface`
 *                            -> tagging -> non-manifold loop -> small pieces -> medial flags.  `conn` = connectivity of
 *                            the entire mesh (NULL: built inside; F_medial then has the face count of that build).
 *                            stats = {mirrored hexes, tagging sweeps, non-manifold rounds, pieces, hexes kept, vertices
 *                            kept}; signed_dis / F_medial / V_medial / stats may be NULL. */
int  fpohm_reorder_hexes(fpohm_ctx *ctx, const double *V, int64_t nV, uint32_t *hex, int64_t H, int64_t *n_mirrored);
int  fpohm_tag_uneven_elements(fpohm_ctx *ctx, const fpohm_conn *conn, uint8_t *H_flag, int32_t *n_sweeps);
int  fpohm_reindex_submesh(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, const uint8_t *H_flag, int32_t *V_map,
                           int32_t *V_map_reverse, int64_t *n_sub_v, int32_t *H_map_reverse, int64_t *n_sub_h, uint32_t *sub_hex);
int  fpohm_clean_non_manifold(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, uint8_t *H_flag, int32_t *n_rounds);
int  fpohm_drop_small_pieces(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, uint8_t *H_flag, int64_t *n_pieces);
int  fpohm_medial_surface_flags(fpohm_ctx *ctx, const fpohm_conn *conn, const uint8_t *H_flag, uint8_t *F_medial, uint8_t *V_medial);
int  fpohm_clean_hex_mesh(fpohm_ctx *ctx, fpohm_mesh *surface, const double *V, int64_t nV, uint32_t *hex, int64_t H,
                          const fpohm_conn *conn, double *signed_dis, uint8_t *H_flag, uint8_t *F_medial, uint8_t *V_medial,
                          int64_t stats[6]);

/* ---- extract_surface_conforming_mesh (global_functions.cpp:1021-1072) with orient_surface_mesh (:1073-1112): the boundary
 * faces of the hex mesh behind `conn` as a quad surface (as_triangles = 0; Mesh_type::Qua) or a triangle surface (two
 * triangles (0,1,2), (2,3,0) per quad; Mesh_type::Tri), vertices renumbered, consistently oriented from face 0 and flipped
 * as a whole when the signed volume is positive, with the tables build_connectivity leaves on a surface mesh.  V = the
 * hex mesh's vertex positions (xyz per vertex).  The component of face 0 must be an orientable 2-manifold (what
 * clean_hex_mesh leaves behind); otherwise FPOHM_EINVAL.
 * export: V 3/vertex, F_vs and F_es vn/face, E_vs 2/edge, flags, V_map (per hex-mesh vertex, -1 = not on the boundary),
 * V_map_reverse, F_map (per hex-mesh face; first triangle for a triangle surface), F_map_reverse.  Any may be NULL.
 * csr which: 0 E.neighbor_fs 1 V.neighbor_vs 2 V.neighbor_es 3 V.neighbor_fs. */
int  fpohm_extract_surface(fpohm_ctx *ctx, const fpohm_conn *conn, const double *V, int32_t as_triangles, fpohm_surface **out);
int  fpohm_surface_sizes(const fpohm_surface *s, int64_t *nV, int64_t *nF, int64_t *nE, int32_t *vn, int64_t *bfs_levels);
int  fpohm_surface_export(const fpohm_surface *s, double *V, uint32_t *F_vs, uint32_t *F_es, uint32_t *E_vs, uint8_t *E_boundary,
                          uint8_t *V_boundary, int32_t *V_map, int32_t *V_map_reverse, int32_t *F_map, int32_t *F_map_reverse);
int  fpohm_surface_csr(const fpohm_surface *s, int32_t which, int64_t *off, uint32_t *val, int64_t *total);
void fpohm_surface_free(fpohm_surface *s);

/* ---- SLIM per-element stages, tet branch (slim_m.cpp).  Jacobians travel as the reference's s.Ji: n x 9
 * row-major, entry 3r + c = ji(r, c).  energy = SLIM_ENERGY (global_types.h:56-64: 0 ARAP, 1 LOG_ARAP, 2 SYMMETRIC_DIRICHLET,
 * 3 CONFORMAL, 4 EXP_CONFORMAL, 5 EXP_SYMMETRIC_DIRICHLET), exp_factor = s.exp_factor.  Floating point: 1e-5 relative.
 *   fpohm_slim_jacobians          compute_jacobians (:84-106): Dx, Dy, Dz as ONE CSR pattern (off n+1, col nnz) with three value
 *                                 arrays, uv = nv x 3 row-major; Ji = [Dx u, Dy u, Dz u, Dx v, Dy v, Dz v, Dx w, Dy w, Dz w].
 *   fpohm_slim_weights_rotations  update_weights_and_closest_rotations (:229-381) after its compute_jacobians call: W n x 9
 *                                 = (W_11, W_12, W_13, W_21, ... W_33), Ri n x 9 as s.Ri (ri stored column by column).
 *   fpohm_slim_energy             compute_energy_with_jacobians (:861-913): sum_i areas[i] * f(singular values of ji).
 * The _dev forms take device pointers and a cudaStream_t and do not synchronise (the optimisation loop keeps uv, Ji, W, Ri
 * resident between the solve and the line search). */
int  fpohm_slim_jacobians(fpohm_ctx *ctx, int64_t n, int64_t nv, const int64_t *off, const int32_t *col, const double *vx,
                          const double *vy, const double *vz, const double *uv, double *Ji);
int  fpohm_slim_weights_rotations(fpohm_ctx *ctx, const double *Ji, int64_t n, int32_t energy, double exp_factor, double *W, double *Ri);
int  fpohm_slim_energy(fpohm_ctx *ctx, const double *Ji, int64_t n, const double *areas, int32_t energy, double exp_factor, double *energy_out);
/* igl::flip_avoiding::compute_max_step_from_singularities (igl/flip_avoiding_line_search.cpp:177-299, tets; called by the line
 * search of slim_solve, slim_m.cpp:1224-1387): the smallest positive t at which some tet (uv + t d) has zero volume, +inf if
 * none.  uv, d = nv x 3 row-major, T = n x 4 vertex ids; roots (n, optional) = the per-tet values of get_min_pos_root_3D. */
int  fpohm_slim_max_step(fpohm_ctx *ctx, const double *uv, int64_t nv, const int32_t *T, int64_t n, const double *d, double *roots,
                         double *max_step);
int  fpohm_slim_max_step_dev(fpohm_ctx *ctx, const double *uv_dev, const int32_t *T_dev, int64_t n, const double *d_dev, double *roots_dev,
                             double *max_step_dev, void *stream);
/* per-element part of buildRhs (slim_m.cpp:1061-1083): f_rhs (9 n, block layout i + (3a + b) n) = rows of W times columns of ri, the
 * reference's left-to-right sums (bit-exact given W and Ri); the product with A^T and the proximal term stay with the caller's solve. */
int  fpohm_slim_rhs_terms(fpohm_ctx *ctx, const double *W, const double *Ri, int64_t n, double *f_rhs);
int  fpohm_slim_rhs_terms_dev(fpohm_ctx *ctx, const double *W_dev, const double *Ri_dev, int64_t n, double *f_rhs_dev, void *stream);
int  fpohm_slim_jacobians_dev(fpohm_ctx *ctx, int64_t n, const int64_t *off_dev, const int32_t *col_dev, const double *vx_dev,
                              const double *vy_dev, const double *vz_dev, const double *uv_dev, double *Ji_dev, void *stream);
int  fpohm_slim_weights_rotations_dev(fpohm_ctx *ctx, const double *Ji_dev, int64_t n, int32_t energy, double exp_factor, double *W_dev,
                                      double *Ri_dev, void *stream);
int  fpohm_slim_energy_dev(fpohm_ctx *ctx, const double *Ji_dev, int64_t n, const double *areas_dev, int32_t energy, double exp_factor,
                           double *energy_dev, void *stream);

/* ---- conforming_mesh (grid_meshing/grid_hex_meshing.cpp:568-696): the octree hex mesh with every big
 * face at a T-junction replaced by the 4 small faces of the other side and the mid vertices inserted into the loops of
 * the faces around it, as a polyhedral ("Hyb") mesh with the connectivity build_connectivity gives it (gf.cpp:187-264):
 * faces = vertex loops (4..8), cells = face lists + sorted vertex sets, edges, boundary flags, F.neighbor_hs.
 * `conn` must be fpohm_hex_connectivity of fpohm_octree_hexes(oct) (vertex i = octree node i).  Ids and orders are the
 * reference's for the same input numbering.  sizes = {nV, nF, nH, nE, sum |F.vs|, sum |H.fs|, sum |H.vs|, sum |F.nhs|}. */
typedef struct fpohm_hybrid fpohm_hybrid;
int  fpohm_conforming_mesh(fpohm_ctx *ctx, const fpohm_octree *oct, const fpohm_conn *conn, fpohm_hybrid **out);
/* the same for an octree given as host tables in ANY numbering (e.g. the reference's own OctreeGrid: m_Nodes[i].position /
 * .neighNodeId); vertex i of `conn` = node i */
int  fpohm_conforming_mesh_tables(fpohm_ctx *ctx, const int32_t *node_pos, const int32_t *node_neigh, int64_t n_nodes,
                                  const int32_t grid_size[3], const fpohm_conn *conn, fpohm_hybrid **out);
int  fpohm_hybrid_sizes(const fpohm_hybrid *hy, int64_t sizes[8], int64_t *n_replaced_faces);
int  fpohm_hybrid_export(const fpohm_hybrid *hy, int64_t *F_off, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs,
                         uint8_t *E_boundary, uint8_t *V_boundary, int64_t *H_foff, uint32_t *H_fs, int64_t *H_voff, uint32_t *H_vs,
                         int64_t *F_nhoff, uint32_t *F_nhs);
void fpohm_hybrid_free(fpohm_hybrid *hy);
/* dual_conforming_mesh (grid_hex_meshing.cpp:697-872): the dual polyhedral mesh of `hy` (vertices = centres of the hexes
 * of the octree mesh `Vpos`/`hex`, one face per interior edge, one cell per interior vertex), its connectivity, and the
 * element-type census with each cell's vertex list re-ordered for its template (Element_Type, global_types.h:40-48:
 * 0 tetrahedral (also: unrecognised, empty list) 1 slab 2 pyramid 3 prism 4 pyramid-combine 5 tet-combine 6 hexahedral).
 * The result is read with fpohm_hybrid_sizes / _export plus fpohm_hybrid_dual_extra. */
int  fpohm_dual_conforming_mesh(fpohm_ctx *ctx, const fpohm_hybrid *hy, const double *Vpos, int64_t nV, const uint32_t *hex, int64_t n_hex,
                                fpohm_hybrid **out);
int  fpohm_hybrid_dual_extra(const fpohm_hybrid *dual, double *V, int32_t *h_type, int64_t census[7]);


/* voxel_meshing lattice (ghm.cpp:215-296, the `--o 0` path): dim[d] = ceil(extent/len), float grid_length,
 * vertex (i,j,k) id = i*dimY*dimZ + j*dimZ + k; hexes in hex_ref_shape corner order. Two-phase via dims. */
int fpohm_voxel_lattice_dims(const double bb_min[3], const double bb_max[3], int32_t num_voxels, int32_t dim[3]);
int fpohm_voxel_lattice(fpohm_ctx *ctx, const double bb_min[3], const double bb_max[3], int32_t num_voxels,
                        double *Vpos, uint32_t *hex);

/* ---- wire formats, host code: ASCII files byte-identical to the reference's writers, rows formatted by
 * all host threads.  V is xyz-interleaved (Mesh::V, 3 x n column-major).  mesh_type = the reference's Mesh_type
 * (global_types.h:457-465: 0 Tri, 1 Qua, 2 HSur, 3 Tet, 4 Hyb, 5 Hex).
 * _write_mesh:  h_io::write_hybrid_mesh_MESH (io.cpp:295-325): "Vertices", then "Triangles" (Tri / HSur, elems = 3 ids per
 *               face) or "Hexahedra" (Hex, 8 ids per hex), ids written 1-based; other types get no element block, as there.
 * _write_vtk:   h_io::write_hybrid_mesh_VTK (io.cpp:101-181): Tri / Qua faces, Hyb polygons (elem_off = CSR offsets,
 *               n_elems + 1 entries), otherwise cells of `arity` vertices (Tet -> type 10, else 12); POINT_DATA "fixed" =
 *               V_boundary for n_point_data entries (the reference writes hmi.Vs.size() of them).
 * _read_fgraph: h_io::read_feature_Graph_FGRAPH (io.cpp:412-434), two-phase (NULL arrays: header and counts only);
 *               pairs = 2 ids per feature edge.  A missing file is FPOHM_EINVAL (the reference returns false).
 * _write_fgraph: h_io::write_feature_Graph_FGRAPH (io.cpp:435-446) for already selected corners / feature edges. */
int fpohm_io_write_mesh(const char *path, const double *V, int64_t nV, int32_t mesh_type, const uint32_t *elems, int64_t n_elems);
int fpohm_io_write_vtk(const char *path, const double *V, int64_t nV, int32_t mesh_type, const int64_t *elem_off, const uint32_t *elems,
                       int64_t n_elems, int32_t arity, const uint8_t *V_boundary, int64_t n_point_data);
int fpohm_io_read_fgraph(const char *path, double *angle_threshold, int32_t *orphan_curve, int32_t *orphan_curve_single,
                         int32_t *corners, int64_t *n_corners, int32_t *pairs, int64_t *n_pairs);
int fpohm_io_write_fgraph(const char *path, double angle_threshold, int32_t orphan_curve, int32_t orphan_curve_single,
                          const int32_t *corners, int64_t n_corners, const int32_t *pairs, int64_t n_pairs);

#ifdef __cplusplus
}
#endif
#endif /* FPOHM_H */

// fpohm_mesh: device-resident triangle soup, facet-bbox tree (subdivision predicate) and the
// igl-identical closest-point tree with its normals (query layer).
#pragma once
#include "internal.h"

#include <functional>
#include <vector>

namespace fpohm {

// One internal node of the flattened closest-point tree.  Both child boxes live in the parent so a
// node visit is ONE 128-byte line.  child >= 0: internal node index; child < 0: leaf, primitive = ~child.
struct alignas(128) QNode {
	double lmin[3], lmax[3];
	double rmin[3], rmax[3];
	int32_t left, right;
	int32_t parent;   // internal node index of the parent, -1 for the root
	int32_t depth;    // root = 0
	int32_t pad[2];
};
static_assert(sizeof(QNode) == 128, "QNode must be one cache line");

// The same node for the packet search's conservative fp32 filter: child boxes rounded OUTWARDS to float.
struct alignas(64) QNodeF {
	float lmin[3], lmax[3];
	float rmin[3], rmax[3];
	int32_t left, right;
	int32_t pad[2];
};
static_assert(sizeof(QNodeF) == 64, "QNodeF must be half a cache line");

// 8-wide collapse of the same tree for the box-parallel packet search (closest_point.cu, cp_wide_kernel): a node is
// eight 32-byte child entries — one 256-byte line pair, lane (node, child) of a warp reads ONE entry with two 16-byte
// loads.  Boxes are the igl boxes rounded OUTWARDS to float (a conservative filter only; exact arithmetic stays fp64).
// child >= 0: wide node index; child < 0: facet ~child; child == WCHILD_EMPTY: unused slot (box = +inf/-inf).
// flags bit 0 (children that are wide nodes only): every child of that node is a facet ("cluster").
#define WCHILD_EMPTY ((int32_t)0x80000000)
struct alignas(32) WChild {
	float lo[3], hi[3];
	int32_t child;
	int32_t flags;
};
static_assert(sizeof(WChild) == 32, "WChild must be 32 bytes");
struct alignas(128) WNode { WChild c[8]; };
static_assert(sizeof(WNode) == 256, "WNode must be 256 bytes");

// host-side result of the igl::AABB::init restatement, in DFS pre-order (node 0 = root)
struct HostTree {
	std::vector<double> box;    // 6 per node: min xyz, max xyz
	std::vector<int32_t> prim;  // -1 for internal nodes
	std::vector<int32_t> lr;    // 2 per node, -1 for leaves
};

void build_igl_tree(const double *V, int64_t nV, const int32_t *F, int64_t nF, HostTree &out);
void host_rank_axis(const double *V, const int32_t *F, int64_t nF, int d, int32_t *rank);
void build_igl_normals(const double *V, int64_t nV, const int32_t *F, int64_t nF,
                       std::vector<double> &FN, std::vector<double> &VN, std::vector<double> &EN,
                       std::vector<int32_t> &E, std::vector<int32_t> &EMAP);

} // namespace fpohm

struct fpohm_mesh {
	fpohm_ctx *ctx = nullptr;
	int64_t nV = 0, nF = 0;
	std::vector<double> hV;
	std::vector<int32_t> hF;
	fpohm::DevBuf<double> V;       // 3 per vertex
	fpohm::DevBuf<int32_t> F;      // 3 per facet
	fpohm::DevBuf<double> tri;     // 9 per facet: A, B, C
	double bbox[6] = {0, 0, 0, 0, 0, 0};

	// subdivision-predicate structure: implicit balanced tree over Morton-sorted facet boxes
	// (heap layout, node 1 = root, children 2n / 2n+1, leaves = single facet boxes), 6 doubles per node
	bool has_pred = false;
	int64_t pred_nodes = 0;
	fpohm::DevBuf<double> pred_box;
	fpohm::DevBuf<int32_t> pred_order; // leaf j of the tree holds facet pred_order[j]

	// query structure
	bool cached = false;                 // owned by the context's cache: fpohm_mesh_free only drops a reference
	bool has_tree = false;
	int tree_ties_host[3] = {0, 0, 0};   // axes whose barycentre ranks came from the host sort (equal coordinates)
	fpohm::DevBuf<double> t_box;         // igl tree in DFS pre-order (node 0 = root): box 6 per node, prim -1 for internal nodes
	fpohm::DevBuf<int32_t> t_prim;
	bool htree_valid = false;            // host copy (fpohm_mesh_tree_export) is made on demand
	fpohm::HostTree htree;
	int64_t n_qnodes = 0;
	fpohm::DevBuf<fpohm::QNode> qnodes;
	int32_t qroot = 0;             // >= 0 internal node, < 0 leaf (~prim) when nF == 1
	int32_t qdepth = 0;            // deepest internal node
	fpohm::DevBuf<fpohm::QNodeF> qfnodes;
	fpohm::DevBuf<int32_t> prim_parent; // internal node holding facet f as a child
	fpohm::DevBuf<int2> node_pd;        // (parent, depth) per internal node: 8 B instead of a 128-byte QNode line for the tie-break's walks to a common ancestor
	int64_t n_wnodes = 0;
	fpohm::DevBuf<fpohm::WNode> wnodes; // 8-wide collapse (node 0 = root); empty when nF < 2
	fpohm::DevBuf<float4> trif;         // 3 float4 per facet: vertices rounded to nearest float (fp32 refine filter)
	float eps_v = 0.f;                  // >= max |v - float(v)| over the vertices
	float slack_q = 0.f;                // >= the error of an fp32 barycentric combination of three vertices (36 u max|coordinate|)
	std::vector<double> hFN, hVN, hEN;   // host copies for fpohm_mesh_normals, made on demand (mesh_host_normals)
	std::vector<int32_t> hE, hEMAP;
	bool hnormals_valid = false;
	int64_t nE = 0;
	fpohm::DevBuf<int32_t> dE;           // 2 per unique undirected edge
	fpohm::DevBuf<double> FN, VN, EN;
	fpohm::DevBuf<int32_t> EMAP;
};

namespace fpohm {
// igl::AABB::init on the device (tree_device.cu): fills `out` exactly as build_igl_tree does.  ties_host[d] reports the axes whose
// ranks had to come from the host sort.
// `while_sorting` runs on the calling thread after the host sorts of the tied axes have been started and before they are joined
void build_igl_tree_device(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s, int ties_host[3], const std::function<void()> &while_sorting);      // fills m->t_box / m->t_prim
struct WideKid { int32_t dfs, cnt, wide; };      // binary node behind a wide child, its facet count (0: empty slot), wide node it becomes
// shape of the 8-wide collapse: a function of the facet count alone (host, data-free) -> 8 entries per wide node
void wide_shape_host(int64_t nF, std::vector<WideKid> &kids, int32_t &n_wide);
void flatten_tree_device(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s, const double *box, const int32_t *prim, const std::vector<WideKid> &kids, int32_t n_wide);   // tree_flatten.cu
void mesh_host_tree(fpohm_mesh *m);
void build_normals_device(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s);                           // normals_device.cu: FN / VN / EN / E / EMAP
void mesh_host_normals(fpohm_mesh *m);                                                              // host copies, on demand                                                                 // host copy of the DFS arrays, on demand
void mesh_ensure_pred(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s);
void mesh_ensure_tree(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s);
// head of clean_hex_mesh on device arrays: bbox centres -> signed distance -> flag = S < 0 (closest_point.cu)
void classify_hexes_dev(fpohm_ctx *ctx, fpohm_mesh *surface, const double *V_dev, const uint32_t *hex_dev, int64_t H,
                        double *S_dev, uint8_t *flag_dev, cudaStream_t s);
// closest point (+ pseudonormal sign when with_sign) of np device-resident points; S = signed distance, or the
// SQUARED distance when !with_sign.  Any output may be null.
// sort_policy: 0 probe the batch on the device and walk it in Morton order if its packets are not compact, 1 never, 2 always
void launch_closest_point(fpohm_ctx *ctx, fpohm_mesh *m, bool with_sign, const double *P_dev, int64_t np,
                          double *S, int32_t *I, double *C, double *N, cudaStream_t s, int sort_policy = 0);
} // namespace fpohm

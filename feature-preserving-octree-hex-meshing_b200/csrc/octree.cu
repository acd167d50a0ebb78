// Level-synchronous construction of the graded (2:1 over faces AND edges) and paired octree.
//
// Replaces OctreeGrid (grid_meshing/octree.cpp) as driven by octree_mesh (ghm.cpp:460-567):
//   createRootCells            octree.cpp:64-117
//   subdivide(pred, ...)       octree.cpp:648-690   BFS: a cell is TESTED iff it is a root or its parent's
//                                                   predicate was true (children pushed only then, :676-678)
//   splitCell                  octree.cpp:501-593
//   makeCellGraded             octree.cpp:598-627   internal cell => its 6 face + 12 edge neighbours of the same
//                                                   size exist (their parents are internal)
//   makeCellPaired + siblings  octree.cpp:574-590,632-643   internal cell => all 7 siblings internal
//                                                   (root cell => all roots internal)
//   should_subdivide           ghm.cpp:502-517 -> geo/mesh/mesh_AABB.h:214-239 -> geo/basic/geometry.h:612-622
//
// The sequential code only ever splits a cell when one of these rules forces it, and enforces every rule
// after every split, so its result is the LEAST fix-point of the rules over the predicate-true set P and is
// independent of split order (DESIGN.md §octree proves the two directions).  Rules only push constraints
// from level l+1 to level l (grading, tree) or sideways inside one sibling family (pairing), hence:
//   phase 1 (top-down)   T_0 = roots, P_l = {c in T_l : pred(c)}, T_{l+1} = children(P_l)
//   phase 2 (bottom-up)  I_l = family_close(P_l ∪ parents(I_{l+1}) ∪ parents(N18(I_{l+1})))
// with every set a sorted array of per-level Morton codes; union = radix sort + unique, family_close = unique on
// code>>3 then x8 expansion.  Phase 3 numbers cells/nodes canonically and derives every link table with
// binary searches instead of the reference's linked-list walks.
#include "octree.h"
#include <chrono>

#include <algorithm>
#include <map>
#include <set>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>

using namespace fpohm;

namespace {

constexpr uint64_t INVALID = ~0ull;

// Cube::delta / invDelta, common.h:147-158
__host__ __device__ __forceinline__ int corner_to_morton(int k) { return ((k & 1) ^ ((k >> 1) & 1)) | (k & 2) | (k & 4); }
__host__ __device__ __forceinline__ int morton_to_corner(int m) { const int dx = m & 1, dy = (m >> 1) & 1, dz = (m >> 2) & 1; return dy ? 4 * dz + 3 - dx : 4 * dz + dx; }

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *a, int64_t lo, int64_t hi, uint64_t key) {
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if (a[mid] < key) lo = mid + 1; else hi = mid;
	}
	return lo;
}
// rank of `code` among internal cells of `level` (global rank, i.e. including lower levels), or -1
__device__ __forceinline__ int64_t internal_rank(const LevelTable &t, int level, uint64_t code) {
	if (level < 0 || level >= t.n_levels) return -1;
	const int64_t lo = t.off[level], hi = t.off[level + 1];
	const int64_t i = lower_bound_u64(t.code, lo, hi, code);
	return (i < hi && t.code[i] == code) ? i : -1;
}

// ---------------------------------------------------------------------------------------------------
// phase 1
__global__ void roots_kernel(uint64_t *__restrict__ code, int rx, int ry, int rz) {
	// roots are listed in MORTON order here (sets are sorted arrays); their cell ids are Layout3D ids (phase 3)
	const int n = rx * ry * rz;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const int x = i % rx, y = (i / rx) % ry, z = (i / rx) / ry;
		code[i] = morton3(x, y, z);
	}
}

// should_subdivide (ghm.cpp:502-517) for every cell of T_l over the complete binary facet-box tree (the reference
// visits every overlapping leaf, mesh_AABB.h:214-239, but only ORs a flag — any-hit is the same result).
//
// Phase A: one thread per cell, depth-first with early exit, at most PRED_BUDGET node visits (swept 48..512 on a 200 k and
// a 2 M facet mesh: 256 is within 1 % of the best on both; 48 left too many lanes to phase B on the deeper tree).  Nearly
// every cell finishes here.  Phase B: the few cells that overlap many internal union boxes without touching a facet box (cells in
// concavities: the serial chain was ~2 300 dependent L2 loads = 350 us for ONE thread on the first ncu capture) are
// finished by the whole warp: the frontier lives in shared memory and 32 nodes are tested per step.
#define PRED_BUDGET 256
#define PRED_WCAP 1024

__device__ __forceinline__ bool box_overlap(const double *__restrict__ b, double mn0, double mn1, double mn2, double mx0, double mx1, double mx2) {
	// bboxes_overlap, geo/basic/geometry.h:612-622 (closed intervals)
	return !(mx0 < b[0] || mn0 > b[3] || mx1 < b[1] || mn1 > b[4] || mx2 < b[2] || mn2 > b[5]);
}

__device__ bool coop_any_overlap(const double *__restrict__ box, uint32_t P, double mn0, double mn1, double mn2, double mx0, double mx1,
                                 double mx2, uint32_t *stk, int lane)
{
	int top = 1;
	if (lane == 0) stk[0] = 1;
	__syncwarp();
	while (top > 0) {
		int take = top < 32 ? top : 32;
		if (top > PRED_WCAP - 64) take = 1;          // nearly full: pure DFS, growth bounded by the tree depth
		uint32_t nd = 0;
		bool ov = false;
		if (lane < take) {
			nd = stk[top - 1 - lane];
			ov = box_overlap(box + 6 * (int64_t)nd, mn0, mn1, mn2, mx0, mx1, mx2);
		}
		__syncwarp();
		top -= take;
		if (__any_sync(0xffffffffu, ov && nd >= P)) return true;
		const unsigned m = __ballot_sync(0xffffffffu, ov);
		if (ov) {
			const int r = __popc(m & ((1u << lane) - 1));
			stk[top + 2 * r] = 2 * nd + 1;
			stk[top + 2 * r + 1] = 2 * nd;
		}
		top += 2 * __popc(m);
		__syncwarp();
	}
	return false;
}

__global__ void __launch_bounds__(256)
predicate_kernel(const uint64_t *__restrict__ cells, int64_t n, int shift /*depth - level*/, double bx, double by, double bz,
                 double vs, const double *__restrict__ box, int64_t P64, uint8_t *__restrict__ flag, int budget)
{
	__shared__ uint32_t wstack[8][PRED_WCAP];
	const uint32_t P = (uint32_t)P64;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {   // warp-uniform trip count
		const int64_t i = base + lane;
		const bool valid = i < n;
		double mn0 = 0, mn1 = 0, mn2 = 0, mx0 = 0, mx1 = 0, mx2 = 0;
		if (valid) {
			const uint64_t c = cells[i];
			const int x = (int)(compact1by2(c) << shift), y = (int)(compact1by2(c >> 1) << shift), z = (int)(compact1by2(c >> 2) << shift);
			const int extent = 1 << shift;
			// box.xyz_min = (mesh_transform + origin) + voxel_size * x ; box.xyz_max = box.xyz_min + voxel_size * extent
			mn0 = bx + vs * x; mn1 = by + vs * y; mn2 = bz + vs * z;
			mx0 = mn0 + vs * extent; mx1 = mn1 + vs * extent; mx2 = mn2 + vs * extent;
		}
		uint32_t stack[34];
		int sp = 0, visits = 0;
		bool hit = false, done = !valid;
		if (valid) stack[sp++] = 1;
		while (!done) {
			if (sp == 0) { done = true; break; }
			if (visits >= budget) break;
			const uint32_t nd = stack[--sp];
			++visits;
			if (!box_overlap(box + 6 * (int64_t)nd, mn0, mn1, mn2, mx0, mx1, mx2)) continue;
			if (nd >= P) { hit = true; done = true; break; }
			stack[sp++] = 2 * nd + 1;
			stack[sp++] = 2 * nd;
		}
		unsigned todo = __ballot_sync(0xffffffffu, !done);
		while (todo) {
			const int src = __ffs(todo) - 1;
			const double a0 = __shfl_sync(0xffffffffu, mn0, src), a1 = __shfl_sync(0xffffffffu, mn1, src), a2 = __shfl_sync(0xffffffffu, mn2, src);
			const double b0 = __shfl_sync(0xffffffffu, mx0, src), b1 = __shfl_sync(0xffffffffu, mx1, src), b2 = __shfl_sync(0xffffffffu, mx2, src);
			const bool h = coop_any_overlap(box, P, a0, a1, a2, b0, b1, b2, wstack[warp], lane);
			if (lane == src) hit = h;
			todo &= todo - 1;
		}
		if (valid) flag[i] = hit;
	}
}

// Coarse levels have few cells (16 ... a few thousand) but each cell is large: one WARP per cell, spread over the whole
// chip, so a level costs the latency of ONE cooperative descent instead of 32 serial ones per warp.
__global__ void __launch_bounds__(256)
predicate_warp_kernel(const uint64_t *__restrict__ cells, int64_t n, int shift, double bx, double by, double bz, double vs,
                      const double *__restrict__ box, int64_t P64, uint8_t *__restrict__ flag)
{
	__shared__ uint32_t wstack[8][PRED_WCAP];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t nwarps = (int64_t)gridDim.x * 8;
	for (int64_t i = blockIdx.x * 8 + warp; i < n; i += nwarps) {
		const uint64_t c = cells[i];
		const int x = (int)(compact1by2(c) << shift), y = (int)(compact1by2(c >> 1) << shift), z = (int)(compact1by2(c >> 2) << shift);
		const int extent = 1 << shift;
		const double mn0 = bx + vs * x, mn1 = by + vs * y, mn2 = bz + vs * z;
		const double mx0 = mn0 + vs * extent, mx1 = mn1 + vs * extent, mx2 = mn2 + vs * extent;
		const bool h = coop_any_overlap(box, (uint32_t)P64, mn0, mn1, mn2, mx0, mx1, mx2, wstack[warp], lane);
		if (lane == 0) flag[i] = h;
		__syncwarp();
	}
}

__global__ void children_kernel(const uint64_t *__restrict__ parents, int64_t n, uint64_t *__restrict__ out) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 8 * n; i += (int64_t)gridDim.x * blockDim.x)
		out[i] = (parents[i >> 3] << 3) | (uint64_t)(i & 7);
}

// ---------------------------------------------------------------------------------------------------
// phase 2: cells forced at level l by the internal cells of level l+1.
// For a cell c with parent p and octant o, parents(c ∪ N18(c)) = { p + (a_x s_x, a_y s_y, a_z s_z) : a in {0,1}^3,
// |a| <= 2 } with s = -1/+1 for o = 0/1: 7 candidates, a = 0 is the tree rule.
__global__ void forced_kernel(const uint64_t *__restrict__ cells, int64_t n, int graded, int lx, int ly, int lz /*level-l bounds*/,
                              uint64_t *__restrict__ out)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const uint64_t c = cells[i];
		const uint64_t p = c >> 3;
		const int px = (int)compact1by2(p), py = (int)compact1by2(p >> 1), pz = (int)compact1by2(p >> 2);
		const int sx = (c & 1) ? 1 : -1, sy = (c & 2) ? 1 : -1, sz = (c & 4) ? 1 : -1;
		uint64_t *o = out + 7 * i;
		o[0] = p;
#pragma unroll
		for (int a = 1; a < 7; ++a) { // a = bit mask over axes, 7 (=vertex neighbour) excluded
			uint64_t v = INVALID;
			if (graded) {
				const int qx = px + ((a & 1) ? sx : 0), qy = py + ((a & 2) ? sy : 0), qz = pz + ((a & 4) ? sz : 0);
				if (qx >= 0 && qy >= 0 && qz >= 0 && qx < lx && qy < ly && qz < lz) v = morton3(qx, qy, qz);
			}
			o[a] = v;
		}
	}
}

__global__ void family_kernel(const uint64_t *__restrict__ in, int64_t n, uint64_t *__restrict__ out) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] >> 3;
}

// ---------------------------------------------------------------------------------------------------
// phase 3
__global__ void root_cells_kernel(uint8_t *__restrict__ lvl, uint64_t *__restrict__ code, int rx, int ry, int rz) {
	const int n = rx * ry * rz;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const int x = i % rx, y = (i / rx) % ry, z = (i / rx) / ry; // Layout3D::toGrid, common.h:76-82
		lvl[i] = 0; code[i] = morton3(x, y, z);
	}
}

// children of the g-th internal cell (global rank over levels) get ids n_roots + 8 g + corner
__global__ void child_cells_kernel(LevelTable t, int64_t n_internal, uint8_t *__restrict__ lvl, uint64_t *__restrict__ code) {
	for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_internal; g += (int64_t)gridDim.x * blockDim.x) {
		int l = 0;
		while (g >= t.off[l + 1]) ++l;
		const uint64_t pc = t.code[g];
		const int64_t base = t.n_roots + 8 * g;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			lvl[base + k] = (uint8_t)(l + 1);
			code[base + k] = (pc << 3) | (uint64_t)corner_to_morton(k);
		}
	}
}

// Top-down form of the link tables (replaced a per-cell search kernel in the build: 12.2 of 73 ms at 59 M cells were its ~8
// binary searches per cell).  (a) per INTERNAL cell g: its own cell id (one search of the parent) gives firstChild;
// (b) per level, coarse to fine: a child's neighbour is a sibling, or hangs off the parent's neighbour q in that
// direction: the mirrored child of q if q is internal (then q has the parent's size), else q itself — the larger (or
// equal) leaf that contains the position, exactly what updateSubcellLinks leaves behind (octree.cpp:255-281).
__global__ void __launch_bounds__(256)
first_child_kernel(LevelTable t, int64_t n_internal, int32_t *__restrict__ first_child, uint8_t *__restrict__ leaf_flag,
                   int32_t *__restrict__ icell)
{
	for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_internal; g += (int64_t)gridDim.x * blockDim.x) {
		int l = 0;
		while (g >= t.off[l + 1]) ++l;
		const uint64_t c = t.code[g];
		int32_t id;
		if (l == 0) {
			id = (int32_t)(compact1by2(c) + t.roots[0] * (compact1by2(c >> 1) + t.roots[1] * compact1by2(c >> 2)));
		} else {
			const int64_t r = internal_rank(t, l - 1, c >> 3);       // exists: the sets are closed under "parent"
			id = (int32_t)(t.n_roots + 8 * r + morton_to_corner((int)(c & 7)));
		}
		first_child[id] = (int32_t)(t.n_roots + 8 * g);
		leaf_flag[id] = 0;
		icell[g] = id;
	}
}
__global__ void root_links_kernel(LevelTable t, int32_t *__restrict__ neigh) {
	const int n = t.n_roots;
	for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const int x = id % t.roots[0], y = (id / t.roots[0]) % t.roots[1], z = (id / t.roots[0]) / t.roots[1];
		const int q[3] = {x, y, z};
		for (int ax = 0; ax < 3; ++ax)
			for (int dir = 0; dir < 2; ++dir) {
				const int v = q[ax] + (dir ? 1 : -1);
				int p[3] = {x, y, z};
				p[ax] = v;
				neigh[6 * id + 2 * ax + dir] = (v >= 0 && v < t.roots[ax]) ? p[0] + t.roots[0] * (p[1] + t.roots[1] * p[2]) : -1;
			}
	}
}
// cells [id0, id0 + n) = one level >= 1; the parents' rows of `neigh` are complete
__global__ void __launch_bounds__(256)
child_links_kernel(int64_t id0, int64_t n, int32_t n_roots, const int32_t *__restrict__ icell, const int32_t *__restrict__ first_child,
                   int32_t *__restrict__ neigh)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t id = id0 + t;
		const int64_t g = (id - n_roots) >> 3;
		const int m = corner_to_morton((int)((id - n_roots) & 7));
		const int64_t base = id - ((id - n_roots) & 7);
		const int32_t P = icell[g];
#pragma unroll
		for (int ax = 0; ax < 3; ++ax) {
#pragma unroll
			for (int dir = 0; dir < 2; ++dir) {
				const int bit = (m >> ax) & 1;
				const int mirrored = morton_to_corner(m ^ (1 << ax));
				int32_t res;
				if (bit != dir) {
					res = (int32_t)(base + mirrored);                       // sibling
				} else {
					const int32_t q = neigh[6 * (int64_t)P + 2 * ax + dir];
					if (q < 0) res = -1;
					else { const int32_t fc = first_child[q]; res = fc >= 0 ? fc + mirrored : q; }
				}
				neigh[6 * id + 2 * ax + dir] = res;
			}
		}
	}
}

// Family-level node keys.  Most leaves sit in families of 8 sibling leaves, whose 64 corners are only 27 distinct lattice
// points: such a family emits 27 keys, every other leaf its 8 corners.  Halves the sort, which was the largest item of
// the numbering (15 of 57 ms at 59 M cells).  The payload of a key is its own index t in the unsorted array (t < 27 *
// n_fam: point t % 27 of family t / 27; else corner (t - 27 n_fam) & 7 of loose leaf (t - 27 n_fam) >> 3) and travels
// as the value array of a (key, value) sort.  FPOHM_OCTREE_PACK=1 instead sorts ONE 8-byte word (key << 31 | index) on
// its upper bits when both fit (node lattices up to 2^11 per axis): a third less traffic on paper, but measured 1.5 ms
// SLOWER at 59 M cells (cub's keys-only onesweep at this size; profiles/r02_octree.md), so it is not the default.
#define NODE_PACK_BITS 31
__global__ void full_family_flags_kernel(const int32_t *__restrict__ icell, int64_t n_internal, const int32_t *__restrict__ first_child,
                                         uint8_t *__restrict__ fam_flag)
{
	for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_internal; g += (int64_t)gridDim.x * blockDim.x) {
		const int32_t fc = first_child[icell[g]];
		bool all = true;
#pragma unroll
		for (int k = 0; k < 8; ++k) all &= first_child[fc + k] < 0;
		fam_flag[g] = all;
	}
}
// leaves whose family is not full (or that are roots) keep the per-leaf path
__global__ void loose_leaf_flags_kernel(const int32_t *__restrict__ leaf_cell, int64_t n_leaves, int32_t n_roots,
                                        const uint8_t *__restrict__ fam_flag, uint8_t *__restrict__ loose)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_leaves; i += (int64_t)gridDim.x * blockDim.x) {
		const int32_t id = leaf_cell[i];
		loose[i] = id < n_roots ? 1 : !fam_flag[(id - n_roots) >> 3];
	}
}
__global__ void family_keys_kernel(const int32_t *__restrict__ fam /* internal ranks g */, int64_t n_fam, const int32_t *__restrict__ icell,
                                   const uint8_t *__restrict__ lvl, const uint64_t *__restrict__ code, int depth, int node_shift,
                                   int pack, uint64_t *__restrict__ keys, uint32_t *__restrict__ payload)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 27 * n_fam; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t fi = t / 27; const int j = (int)(t % 27);
		const int32_t id = icell[fam[fi]];                      // the internal cell whose 8 children are leaves
		const int sh = depth - lvl[id] - 1 - node_shift;       // child extent in key units = 1 << sh
		const uint64_t c = code[id];
		const uint32_t x = (compact1by2(c) << (sh + 1)) + ((uint32_t)(j % 3) << sh);
		const uint32_t y = (compact1by2(c >> 1) << (sh + 1)) + ((uint32_t)((j / 3) % 3) << sh);
		const uint32_t z = (compact1by2(c >> 2) << (sh + 1)) + ((uint32_t)(j / 9) << sh);
		if (pack) keys[t] = (morton3(x, y, z) << pack) | (uint64_t)t;
		else { keys[t] = morton3(x, y, z); payload[t] = (uint32_t)t; }
	}
}
__global__ void loose_leaf_keys_kernel(const int32_t *__restrict__ loose_leaf /* cell ids */, int64_t n_loose, const uint8_t *__restrict__ lvl,
                                       const uint64_t *__restrict__ code, int depth, int node_shift, int pack, int64_t t0,
                                       uint64_t *__restrict__ keys, uint32_t *__restrict__ payload)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_loose; i += (int64_t)gridDim.x * blockDim.x) {
		const int32_t id = loose_leaf[i];
		const int l = lvl[id];
		const uint64_t c = code[id];
		const int sh = depth - l - node_shift; // extent in key units = 1 << sh
		const uint32_t x = compact1by2(c) << sh, y = compact1by2(c >> 1) << sh, z = compact1by2(c >> 2) << sh;
		const uint32_t e = 1u << sh;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const int m = corner_to_morton(k);
			const uint64_t key = morton3(x + ((m & 1) ? e : 0), y + ((m & 2) ? e : 0), z + ((m & 4) ? e : 0));
			const int64_t t = t0 + 8 * i + k;
			if (pack) keys[t] = (key << pack) | (uint64_t)t;
			else { keys[t] = key; payload[t] = (uint32_t)t; }
		}
	}
}
// After the sort: node id = index of the key's run; leaf corners are written through the payload, no search.
__global__ void node_scatter2_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ payload, int pack,
                                     const int32_t *__restrict__ head,
                                     const int32_t *__restrict__ nid_incl, int64_t n, const int32_t *__restrict__ fam, int64_t n_fam_keys,
                                     const int32_t *__restrict__ icell, const int32_t *__restrict__ first_child,
                                     const int32_t *__restrict__ loose_leaf, int node_shift,
                                     uint64_t *__restrict__ node_key, int32_t *__restrict__ node_pos, int32_t *__restrict__ corner)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const int32_t nid = nid_incl[t] - 1;
		const uint64_t kt = key[t];
		const uint32_t q = pack ? (uint32_t)(kt & ((1ull << NODE_PACK_BITS) - 1)) : payload[t];
		if ((int64_t)q < n_fam_keys) {
			const int j = (int)(q % 27);
			const int32_t fc = first_child[icell[fam[q / 27]]];
			const int a = j % 3, b = (j / 3) % 3, c = j / 9;       // lattice point; child (mx,my,mz) has it as local corner (a-mx, b-my, c-mz)
			for (int mz = max(c - 1, 0); mz <= min(c, 1); ++mz)
				for (int my = max(b - 1, 0); my <= min(b, 1); ++my)
					for (int mx = max(a - 1, 0); mx <= min(a, 1); ++mx) {
						const int child = morton_to_corner(mx | (my << 1) | (mz << 2));
						const int loc = morton_to_corner((a - mx) | ((b - my) << 1) | ((c - mz) << 2));
						corner[8 * (int64_t)(fc + child) + loc] = nid;
					}
		} else {
			const uint32_t pl = q - (uint32_t)n_fam_keys;
			corner[8 * (int64_t)loose_leaf[pl >> 3] + (pl & 7)] = nid;
		}
		if (head[t]) {
			const uint64_t k = kt >> pack;
			node_key[nid] = k;
			node_pos[3 * (int64_t)nid] = (int32_t)(compact1by2(k) << node_shift);
			node_pos[3 * (int64_t)nid + 1] = (int32_t)(compact1by2(k >> 1) << node_shift);
			node_pos[3 * (int64_t)nid + 2] = (int32_t)(compact1by2(k >> 2) << node_shift);
		}
	}
}

__global__ void key_heads_kernel(const uint64_t *__restrict__ k, int64_t n, int pack, int32_t *__restrict__ head) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		head[t] = (t == 0 || (k[t] >> pack) != (k[t - 1] >> pack)) ? 1 : 0;
}

// corner k of an internal cell = corner k of its child k, recursively down to a leaf (octree.cpp:549-556)
__global__ void internal_corners_kernel(const int32_t *__restrict__ first_child, int64_t n_cells, int32_t *__restrict__ corner) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 8 * n_cells; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t id = t >> 3; const int k = (int)(t & 7);
		int32_t c = first_child[id];
		if (c < 0) continue;
		c += k;
		for (int32_t nx = first_child[c]; nx >= 0; nx = first_child[c]) c = nx + k;
		corner[t] = corner[8 * (int64_t)c + k];
	}
}

// Node::neighNodeId = other end of the SHORTEST leaf edge leaving the node in each direction (octree.h:85-94 names the 12
// edges as lower corner, upper corner, axis).  Levels are processed coarse to fine with plain 4-byte stores, so the edge
// of the finest (= shortest) leaf touching a node in a direction is the one that stays; leaves of one level that share an
// edge store identical values.  (An atomicMin on (length << 32 | node) was 14.2 of 73 ms at 59 M cells.)
__global__ void __launch_bounds__(256)
loose_edges_kernel(const int32_t *__restrict__ loose_leaf, int64_t n, const int32_t *__restrict__ corner, int32_t *__restrict__ node_neigh)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 12 * n; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t id = loose_leaf[t / 12]; const int k = (int)(t % 12);
		const int ea[12] = {0, 3, 4, 7, 0, 1, 4, 5, 0, 1, 3, 2};
		const int eb[12] = {1, 2, 5, 6, 3, 2, 7, 6, 4, 5, 7, 6};
		const int ax = k >> 2;
		const int32_t a = corner[8 * id + ea[k]], b = corner[8 * id + eb[k]];
		node_neigh[6 * (int64_t)a + 2 * ax + 1] = b;
		node_neigh[6 * (int64_t)b + 2 * ax] = a;
	}
}
// A family of 8 sibling leaves has 96 leaf edges but only 54 distinct ones between its 27 lattice points: one thread per
// lattice point writes the (up to 6) links of its own 24-byte row — 108 stores per family instead of 192.
__global__ void __launch_bounds__(256)
family_edges_kernel(const int32_t *__restrict__ fam, int64_t n_fam, const int32_t *__restrict__ icell, const int32_t *__restrict__ first_child,
                    const int32_t *__restrict__ corner, int32_t *__restrict__ node_neigh)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 27 * n_fam; t += (int64_t)gridDim.x * blockDim.x) {
		const int j = (int)(t % 27);
		const int64_t fc = first_child[icell[fam[t / 27]]];
		const int p[3] = {j % 3, (j / 3) % 3, j / 9};
		auto node_at = [&](int a, int b, int c) {   // child (mx,my,mz) holds lattice point (a,b,c) as its local corner (a-mx, b-my, c-mz)
			const int mx = min(a, 1), my = min(b, 1), mz = min(c, 1);
			return corner[8 * (fc + morton_to_corner(mx | (my << 1) | (mz << 2))) + morton_to_corner((a - mx) | ((b - my) << 1) | ((c - mz) << 2))];
		};
		const int64_t me = node_at(p[0], p[1], p[2]);
#pragma unroll
		for (int ax = 0; ax < 3; ++ax) {
			int q[3] = {p[0], p[1], p[2]};
			if (p[ax] < 2) { q[ax] = p[ax] + 1; node_neigh[6 * me + 2 * ax + 1] = node_at(q[0], q[1], q[2]); }
			if (p[ax] > 0) { q[ax] = p[ax] - 1; node_neigh[6 * me + 2 * ax] = node_at(q[0], q[1], q[2]); }
		}
	}
}
// where each level's full families (ascending internal ranks) and loose leaves (ascending cell ids) begin in their lists
struct LevelBounds { int n; int64_t g[26], id[26]; };
__global__ void level_ranges_kernel(LevelBounds lb, const int32_t *__restrict__ fam, int64_t n_fam, const int32_t *__restrict__ loose,
                                    int64_t n_loose, int64_t *__restrict__ out /* [2][26] */)
{
	const int l = threadIdx.x;
	if (l >= lb.n) return;
	auto lower = [](const int32_t *a, int64_t n, int64_t v) { int64_t lo = 0, hi = n; while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; } return lo; };
	out[l] = lower(fam, n_fam, lb.g[l]);
	out[26 + l] = lower(loose, n_loose, lb.id[l]);
}

__global__ void iota_i32_kernel(int32_t *p, int64_t n) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

// octree_mesh export, ghm.cpp:531-562
__global__ void hex_vertices_kernel(const int32_t *__restrict__ pos, int64_t n, double bx, double by, double bz, double vs,
                                    double *__restrict__ out)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 3 * n; i += (int64_t)gridDim.x * blockDim.x) {
		const int c = (int)(i % 3);
		const double b = c == 0 ? bx : (c == 1 ? by : bz);
		out[i] = b + (double)pos[i] * vs; // (mesh_transform + o) + nodePos * s
	}
}
__global__ void hex_cells_kernel(const int32_t *__restrict__ leaf_cell, int64_t n_leaves, const int32_t *__restrict__ corner,
                                 uint32_t *__restrict__ hex)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 8 * n_leaves; i += (int64_t)gridDim.x * blockDim.x)
		hex[i] = (uint32_t)corner[8 * (int64_t)leaf_cell[i >> 3] + (i & 7)];
}

// is2to1Graded / isPaired, octree.cpp:148-202
__global__ void check_kernel(const int32_t *__restrict__ first_child, const int32_t *__restrict__ corner,
                             const int32_t *__restrict__ nneigh, int64_t n_cells, int32_t *__restrict__ bad /*[2]*/)
{
	for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < n_cells; id += (int64_t)gridDim.x * blockDim.x) {
		const int32_t fc = first_child[id];
		if (fc >= 0) {
			const bool all_leaf = first_child[fc] < 0;
			for (int k = 1; k < 8; ++k) if ((first_child[fc + k] < 0) != all_leaf) atomicAdd(&bad[1], 1);
		} else {
			const int32_t *v = corner + 8 * id;
			const int ea[12] = {0, 3, 4, 7, 0, 1, 4, 5, 0, 1, 3, 2};
			const int eb[12] = {1, 2, 5, 6, 3, 2, 7, 6, 4, 5, 7, 6};
			for (int k = 0; k < 12; ++k) {
				const int ax = k >> 2;
				const int32_t a = v[ea[k]], b = v[eb[k]];
				const int32_t na = nneigh[6 * (int64_t)a + 2 * ax + 1], pb = nneigh[6 * (int64_t)b + 2 * ax];
				if (!(na == b || na == pb)) atomicAdd(&bad[0], 1);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------
struct Sorter {
	fpohm_ctx *ctx; cudaStream_t s;
	// sort + unique, drop INVALID; returns count; result in `out` (allocated here)
	int64_t sort_unique(DevBuf<uint64_t> &in, int64_t n, int bits, DevBuf<uint64_t> &out) {
		if (n == 0) { out.alloc(0, s); return 0; }
		DevBuf<uint64_t> sorted(n, s);
		size_t tb = 0;
		const int end_bit = bits + 1 < 64 ? bits + 1 : 64;   // bit `bits` is 0 in every valid key and 1 in INVALID
		FPOHM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, in.p, sorted.p, n, 0, end_bit, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, in.p, sorted.p, n, 0, end_bit, s));
		DevBuf<uint64_t> uniq(n, s);
		DevBuf<int64_t> cnt(1, s);
		size_t tb2 = 0;
		FPOHM_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, sorted.p, uniq.p, cnt.p, n, s));
		DevBuf<uint8_t> tmp2((int64_t)tb2, s);
		FPOHM_CUDA(cub::DeviceSelect::Unique(tmp2.p, tb2, sorted.p, uniq.p, cnt.p, n, s));
		ctx->launches += 6;
		int64_t m = 0;
		cnt.download(&m, 1);
		uint64_t last = 0;
		FPOHM_CUDA(cudaStreamSynchronize(s));
		if (m > 0) {
			FPOHM_CUDA(cudaMemcpyAsync(&last, uniq.p + (m - 1), 8, cudaMemcpyDeviceToHost, s));
			FPOHM_CUDA(cudaStreamSynchronize(s));
			if (last == INVALID) --m;
		}
		out.alloc(m, s);
		if (m) FPOHM_CUDA(cudaMemcpyAsync(out.p, uniq.p, 8 * (size_t)m, cudaMemcpyDeviceToDevice, s));
		return m;
	}
};

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
// number of Morton key bits for per-axis coordinates in [0, max_coord]
int key_bits(int64_t max_coord) { int b = 1; while ((1ll << b) <= max_coord) ++b; return 3 * b; }

void setup_geometry(fpohm_octree *o, const int32_t gs[3]) {
	for (int d = 0; d < 3; ++d) {
		FPOHM_REQUIRE(gs[d] >= 1 && (gs[d] & (gs[d] - 1)) == 0, FPOHM_EINVAL, "octree: grid_size[%d]=%d is not a power of two (octree.cpp:26-28)", d, gs[d]);
		FPOHM_REQUIRE(gs[d] <= (1 << 21), FPOHM_ERANGE, "octree: grid_size[%d]=%d exceeds 2^21", d, gs[d]);
		o->prm.grid_size[d] = gs[d];
	}
	const int mn = std::min(gs[0], std::min(gs[1], gs[2]));
	o->depth = ilog2(mn);
	o->n_roots = 1;
	for (int d = 0; d < 3; ++d) { o->roots[d] = gs[d] / mn; o->n_roots *= o->roots[d]; }
	FPOHM_REQUIRE(o->depth < FPOHM_MAX_LEVELS, FPOHM_ERANGE, "octree: depth %d too large", o->depth);
}

LevelTable make_table(const fpohm_octree *o) {
	LevelTable t;
	t.code = o->icode.p;
	for (int l = 0; l < FPOHM_MAX_LEVELS + 2; ++l) t.off[l] = o->lvl_off[std::min(l, o->n_levels)];
	t.n_levels = o->n_levels;
	for (int d = 0; d < 3; ++d) t.roots[d] = o->roots[d];
	t.n_roots = o->n_roots;
	t.depth = o->depth;
	return t;
}

// one level of phase 2: I_l = family_close(sort_unique(cand)); `cand` may hold INVALID entries
int64_t close_level(fpohm_octree *o, int l, DevBuf<uint64_t> &cand, int64_t n_cand, DevBuf<uint64_t> &uq) {
	fpohm_ctx *ctx = o->ctx;
	cudaStream_t s = ctx->stream;
	Sorter sorter{ctx, s};
	const int blk = 256;
	const bool paired = o->prm.paired != 0;
	const int lbits = key_bits(((int64_t)std::max(o->roots[0], std::max(o->roots[1], o->roots[2])) << l) - 1);
	int64_t m = sorter.sort_unique(cand, n_cand, lbits, uq);
	if (paired && m > 0) {
		if (l == 0) {
			// root rule, octree.cpp:577-581: one root split => all roots split
			uq.alloc(o->n_roots, s);
			roots_kernel<<<grid_for(ctx, o->n_roots, blk), blk, 0, s>>>(uq.p, o->roots[0], o->roots[1], o->roots[2]);
			FPOHM_LAUNCH_CHECK(ctx);
			DevBuf<uint64_t> sorted_roots;
			m = sorter.sort_unique(uq, o->n_roots, lbits, sorted_roots);
			uq = std::move(sorted_roots);
		} else {
			// sibling rule, octree.cpp:583-587 (+ makeCellPaired :632-643): whole families
			DevBuf<uint64_t> fam(m, s), famu;
			family_kernel<<<grid_for(ctx, m, blk), blk, 0, s>>>(uq.p, m, fam.p);
			FPOHM_LAUNCH_CHECK(ctx);
			const int64_t nf = sorter.sort_unique(fam, m, lbits, famu);
			uq.alloc(8 * nf, s);
			children_kernel<<<grid_for(ctx, 8 * nf, blk), blk, 0, s>>>(famu.p, nf, uq.p);
			FPOHM_LAUNCH_CHECK(ctx);
			m = 8 * nf;
		}
	}
	return m;
}

// candidates of level l: P_l ∪ forced(I_{l+1})
int64_t level_candidates(fpohm_octree *o, int l, const uint64_t *Pl, int64_t nPl, const uint64_t *Iup, int64_t nIup, DevBuf<uint64_t> &cand) {
	fpohm_ctx *ctx = o->ctx;
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	const int64_t n_forced = 7 * nIup;
	const int64_t n_cand = nPl + n_forced;
	cand.alloc(n_cand, s);
	if (nPl) FPOHM_CUDA(cudaMemcpyAsync(cand.p, Pl, 8 * (size_t)nPl, cudaMemcpyDeviceToDevice, s));
	if (n_forced) {
		forced_kernel<<<grid_for(ctx, nIup, blk), blk, 0, s>>>(Iup, nIup, o->prm.graded ? 1 : 0,
			o->roots[0] << l, o->roots[1] << l, o->roots[2] << l, cand.p + nPl);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	return n_cand;
}

void number_levels(fpohm_octree *o, std::vector<DevBuf<uint64_t>> &I, std::vector<int64_t> &nI);

// phase 2 + 3 given the predicate-true sets P[l] (device arrays) — shared by build / from_marks / refine
void close_and_number(fpohm_octree *o, std::vector<DevBuf<uint64_t>> &P, std::vector<int64_t> &nP) {
	int Lmax = -1;
	for (int l = 0; l < (int)P.size(); ++l) if (nP[l] > 0) Lmax = l;
	std::vector<DevBuf<uint64_t>> I((size_t)std::max(Lmax + 1, 0));
	std::vector<int64_t> nI((size_t)std::max(Lmax + 1, 0), 0);
	static const bool timeline = getenv("FPOHM_OCTREE_TIMELINE") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	for (int l = Lmax; l >= 0; --l) {
		DevBuf<uint64_t> cand;
		const int64_t n_cand = level_candidates(o, l, P[l].p, nP[l], l < Lmax ? I[l + 1].p : nullptr, l < Lmax ? nI[l + 1] : 0, cand);
		nI[l] = close_level(o, l, cand, n_cand, I[l]);
	}
	if (timeline) cudaStreamSynchronize(o->ctx->stream);
	const auto t1 = std::chrono::steady_clock::now();
	number_levels(o, I, nI);
	if (timeline) {
		cudaStreamSynchronize(o->ctx->stream);
		o->dbg_close_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		o->dbg_number_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
	}
}

// phase 3 given the closed internal sets I[l] (sorted device arrays)
void number_levels(fpohm_octree *o, std::vector<DevBuf<uint64_t>> &I, std::vector<int64_t> &nI) {
	fpohm_ctx *ctx = o->ctx;
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	int Lmax = -1;
	for (int l = 0; l < (int)I.size(); ++l) if (nI[l] > 0) Lmax = l;
	// concatenate levels
	o->n_levels = Lmax + 1;
	int64_t tot = 0;
	for (int l = 0; l <= Lmax; ++l) { o->lvl_off[l] = tot; tot += nI[l]; }
	for (int l = Lmax + 1; l < FPOHM_MAX_LEVELS + 2; ++l) o->lvl_off[l] = tot;
	o->icode.alloc(std::max<int64_t>(tot, 1), s);
	for (int l = 0; l <= Lmax; ++l)
		if (nI[l]) FPOHM_CUDA(cudaMemcpyAsync(o->icode.p + o->lvl_off[l], I[l].p, 8 * (size_t)nI[l], cudaMemcpyDeviceToDevice, s));

	// ---- phase 3: numbering -------------------------------------------------------------------
	const int64_t n_internal = tot;
	const int64_t n_cells = o->n_roots + 8 * n_internal;
	FPOHM_REQUIRE(n_cells < (1ll << 31), FPOHM_ERANGE, "octree: %lld cells exceed int32 ids (octree.h:140-143)", (long long)n_cells);
	o->n_cells = n_cells;
	o->cell_level.alloc(n_cells, s);
	o->cell_code.alloc(n_cells, s);
	root_cells_kernel<<<grid_for(ctx, o->n_roots, blk), blk, 0, s>>>(o->cell_level.p, o->cell_code.p, o->roots[0], o->roots[1], o->roots[2]);
	FPOHM_LAUNCH_CHECK(ctx);
	const LevelTable t = make_table(o);
	if (n_internal) {
		child_cells_kernel<<<grid_for(ctx, n_internal, blk), blk, 0, s>>>(t, n_internal, o->cell_level.p, o->cell_code.p);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	o->cell_first_child.alloc(n_cells, s);
	o->cell_neigh.alloc(6 * n_cells, s);
	DevBuf<uint8_t> leaf_flag(n_cells, s);
	DevBuf<int32_t> icell(std::max<int64_t>(n_internal, 1), s);      // cell id of the g-th internal cell
	{
		FPOHM_CUDA(cudaMemsetAsync(o->cell_first_child.p, 0xff, 4 * (size_t)n_cells, s));
		FPOHM_CUDA(cudaMemsetAsync(leaf_flag.p, 1, (size_t)n_cells, s));
		if (n_internal) {
			first_child_kernel<<<grid_for(ctx, n_internal, blk), blk, 0, s>>>(t, n_internal, o->cell_first_child.p, leaf_flag.p, icell.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		root_links_kernel<<<grid_for(ctx, o->n_roots, blk), blk, 0, s>>>(t, o->cell_neigh.p);
		FPOHM_LAUNCH_CHECK(ctx);
		for (int l = 1; l <= o->n_levels; ++l) {
			const int64_t id0 = o->n_roots + 8 * o->lvl_off[l - 1], n = 8 * (o->lvl_off[l] - o->lvl_off[l - 1]);
			if (n == 0) continue;
			child_links_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(id0, n, o->n_roots, icell.p, o->cell_first_child.p, o->cell_neigh.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
	}
	// leaves in cell order (hex2Octree_map, ghm.cpp:551) and, in the same pass, the families of 8 sibling leaves
	DevBuf<uint8_t> fam_flag(std::max<int64_t>(n_internal, 1), s);
	DevBuf<int32_t> fam(std::max<int64_t>(n_internal, 1), s);
	int64_t n_fam = 0;
	{
		DevBuf<int32_t> ids(n_cells, s), sel(n_cells, s);
		DevBuf<int64_t> cnt(2, s);
		FPOHM_CUDA(cudaMemsetAsync(cnt.p, 0, 16, s));
		iota_i32_kernel<<<grid_for(ctx, n_cells, blk), blk, 0, s>>>(ids.p, n_cells);
		FPOHM_LAUNCH_CHECK(ctx);
		if (n_internal) {
			full_family_flags_kernel<<<grid_for(ctx, n_internal, blk), blk, 0, s>>>(icell.p, n_internal, o->cell_first_child.p, fam_flag.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		size_t tb = 0, tb1 = 0;
		FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids.p, leaf_flag.p, sel.p, cnt.p, n_cells, s));
		FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb1, ids.p, fam_flag.p, fam.p, cnt.p + 1, std::max<int64_t>(n_internal, 1), s));
		DevBuf<uint8_t> tmp((int64_t)std::max(tb, tb1), s);
		FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, ids.p, leaf_flag.p, sel.p, cnt.p, n_cells, s));
		if (n_internal) FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb1, ids.p, fam_flag.p, fam.p, cnt.p + 1, n_internal, s));
		ctx->launches += 4;
		int64_t hc[2] = {0, 0};
		cnt.download(hc, 2);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		o->n_leaves = hc[0];
		n_fam = hc[1];
		o->leaf_cell.alloc(o->n_leaves, s);
		FPOHM_CUDA(cudaMemcpyAsync(o->leaf_cell.p, sel.p, 4 * (size_t)o->n_leaves, cudaMemcpyDeviceToDevice, s));
	}
	// nodes: unique corners of leaves, Morton order of (pos >> node_shift)
	const int finest_level = o->n_levels; // deepest cells live one level below the deepest internal level
	o->node_shift = o->depth - finest_level;
	FPOHM_REQUIRE(o->node_shift >= 0, FPOHM_ESTATE, "octree: internal cell below extent 1");
	for (int d = 0; d < 3; ++d)
		FPOHM_REQUIRE((o->prm.grid_size[d] >> o->node_shift) < (1 << 21), FPOHM_ERANGE,
		              "octree: %d node positions per axis after shift do not fit 21-bit Morton keys", (o->prm.grid_size[d] >> o->node_shift) + 1);
	o->cell_corner.alloc(8 * n_cells, s);
	// the leaves outside full families keep the per-leaf path; their number needs no read-back
	const int64_t n_loose = o->n_leaves - 8 * n_fam;
	DevBuf<int32_t> loose(std::max<int64_t>(n_loose, 1), s);
	int64_t ranges[2][26];
	FPOHM_REQUIRE(o->n_levels + 2 <= 26, FPOHM_ERANGE, "octree: %d levels", o->n_levels);
	{
		DevBuf<uint8_t> loose_flag(std::max<int64_t>(o->n_leaves, 1), s);
		DevBuf<int64_t> d_ranges(2 * 26, s);
		if (n_loose > 0) {
			DevBuf<int64_t> cnt2(1, s);
			loose_leaf_flags_kernel<<<grid_for(ctx, o->n_leaves, blk), blk, 0, s>>>(o->leaf_cell.p, o->n_leaves, o->n_roots, fam_flag.p, loose_flag.p);
			FPOHM_LAUNCH_CHECK(ctx);
			size_t tb1 = 0;
			FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb1, o->leaf_cell.p, loose_flag.p, loose.p, cnt2.p, o->n_leaves, s));
			DevBuf<uint8_t> tmp1((int64_t)tb1, s);
			FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp1.p, tb1, o->leaf_cell.p, loose_flag.p, loose.p, cnt2.p, o->n_leaves, s));
			ctx->launches += 2;
		}
		{
			// level l = 0 .. n_levels holds the leaves whose parents have internal ranks [lvl_off[l-1], lvl_off[l]) and whose own
			// ids are [id0(l), id0(l+1)); entry n_levels + 1 closes the last range
			LevelBounds lb; lb.n = o->n_levels + 2;
			for (int l = 0; l <= o->n_levels + 1; ++l) {
				lb.g[l] = l == 0 ? 0 : o->lvl_off[l - 1];
				lb.id[l] = l == 0 ? 0 : o->n_roots + 8 * o->lvl_off[l - 1];
			}
			level_ranges_kernel<<<1, 32, 0, s>>>(lb, fam.p, n_fam, loose.p, n_loose, d_ranges.p);
			FPOHM_LAUNCH_CHECK(ctx);
			d_ranges.download(&ranges[0][0], 2 * 26);   // read after the synchronize that follows the scan below
		}
		const int64_t nk = 27 * n_fam + 8 * n_loose;
		FPOHM_REQUIRE(nk < (1ll << 31), FPOHM_ERANGE, "octree: %lld node keys exceed the 31-bit payload", (long long)nk);
		const int64_t gmax = std::max(o->prm.grid_size[0], std::max(o->prm.grid_size[1], o->prm.grid_size[2]));
		const int bits = std::min(64, key_bits(gmax >> o->node_shift));
		static const bool want_pack = getenv("FPOHM_OCTREE_PACK") != nullptr;
		const int pack = (bits + NODE_PACK_BITS <= 64 && want_pack) ? NODE_PACK_BITS : 0;
		DevBuf<uint64_t> keys(nk, s), skeys(nk, s);
		DevBuf<uint32_t> pay(pack ? 0 : nk, s), spay(pack ? 0 : nk, s);
		if (n_fam) {
			family_keys_kernel<<<grid_for(ctx, 27 * n_fam, blk), blk, 0, s>>>(fam.p, n_fam, icell.p, o->cell_level.p, o->cell_code.p,
				o->depth, o->node_shift, pack, keys.p, pay.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		if (n_loose) {
			loose_leaf_keys_kernel<<<grid_for(ctx, n_loose, blk), blk, 0, s>>>(loose.p, n_loose, o->cell_level.p, o->cell_code.p, o->depth,
				o->node_shift, pack, 27 * n_fam, keys.p, pay.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		size_t tb = 0;
		if (pack) {
			FPOHM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, skeys.p, nk, pack, pack + bits, s));
			DevBuf<uint8_t> tmp((int64_t)tb, s);
			FPOHM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, skeys.p, nk, pack, pack + bits, s));
		} else {
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, skeys.p, pay.p, spay.p, nk, 0, bits, s));
			DevBuf<uint8_t> tmp((int64_t)tb, s);
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, skeys.p, pay.p, spay.p, nk, 0, bits, s));
		}
		keys.release(); pay.release();
		DevBuf<int32_t> head(nk, s), nid(nk, s);
		key_heads_kernel<<<grid_for(ctx, nk, blk), blk, 0, s>>>(skeys.p, nk, pack, head.p);
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tb2 = 0;
		FPOHM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, head.p, nid.p, nk, s));
		DevBuf<uint8_t> tmp2((int64_t)tb2, s);
		FPOHM_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, tb2, head.p, nid.p, nk, s));
		ctx->launches += 4;
		int32_t last = 0;
		FPOHM_CUDA(cudaMemcpyAsync(&last, nid.p + (nk - 1), 4, cudaMemcpyDeviceToHost, s));
		FPOHM_CUDA(cudaStreamSynchronize(s));
		o->n_nodes = last;
		o->node_key.alloc(o->n_nodes, s);
		o->node_pos.alloc(3 * o->n_nodes, s);
		node_scatter2_kernel<<<grid_for(ctx, nk, blk), blk, 0, s>>>(skeys.p, spay.p, pack, head.p, nid.p, nk, fam.p, 27 * n_fam, icell.p,
			o->cell_first_child.p, loose.p, o->node_shift, o->node_key.p, o->node_pos.p, o->cell_corner.p);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	internal_corners_kernel<<<grid_for(ctx, 8 * n_cells, blk), blk, 0, s>>>(o->cell_first_child.p, n_cells, o->cell_corner.p);
	FPOHM_LAUNCH_CHECK(ctx);
	{
		o->node_neigh.alloc(6 * o->n_nodes, s);
		FPOHM_CUDA(cudaMemsetAsync(o->node_neigh.p, 0xff, 4 * (size_t)(6 * o->n_nodes), s));
		for (int l = 0; l <= o->n_levels; ++l) {
			const int64_t f0 = ranges[0][l], nf = ranges[0][l + 1] - f0, l0 = ranges[1][l], nl = ranges[1][l + 1] - l0;
			if (nl > 0) {
				loose_edges_kernel<<<grid_for(ctx, 12 * nl, blk), blk, 0, s>>>(loose.p + l0, nl, o->cell_corner.p, o->node_neigh.p);
				FPOHM_LAUNCH_CHECK(ctx);
			}
			if (nf > 0) {
				family_edges_kernel<<<grid_for(ctx, 27 * nf, blk), blk, 0, s>>>(fam.p + f0, nf, icell.p, o->cell_first_child.p, o->cell_corner.p,
					o->node_neigh.p);
				FPOHM_LAUNCH_CHECK(ctx);
			}
		}
	}
	FPOHM_CUDA(cudaStreamSynchronize(s));
}

__global__ void gather_cells_kernel(const int32_t *__restrict__ ids, int64_t n, const uint8_t *__restrict__ level, const uint64_t *__restrict__ code,
                                    const int32_t *__restrict__ first_child, uint8_t *__restrict__ l, uint64_t *__restrict__ c, int32_t *__restrict__ f)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const int32_t id = ids[i];
		l[i] = level[id]; c[i] = code[id]; f[i] = first_child[id];
	}
}

// leaf codes of the level-l cells of an already numbered octree (cells of one level are a contiguous id range)
__global__ void old_leaf_codes_kernel(const uint64_t *__restrict__ code, const int32_t *__restrict__ first_child, int64_t id0, int64_t n,
                                      uint64_t *__restrict__ out)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		out[i] = first_child[id0 + i] < 0 ? code[id0 + i] : INVALID;
}

// P = {c in T : should_subdivide(c)} for cells of level l, compacted in input order
int64_t test_cells(fpohm_octree *o, const fpohm_mesh *mesh, int l, const uint64_t *T, int64_t nT, DevBuf<uint64_t> &sel) {
	fpohm_ctx *ctx = o->ctx;
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	const double bx = o->prm.mesh_transform[0] + o->prm.origin[0], by = o->prm.mesh_transform[1] + o->prm.origin[1],
	             bz = o->prm.mesh_transform[2] + o->prm.origin[2];
	static const int budget_env = getenv("FPOHM_PRED_BUDGET") ? atoi(getenv("FPOHM_PRED_BUDGET")) : 0;
	const int pred_budget = budget_env > 0 ? budget_env : PRED_BUDGET;
	DevBuf<uint8_t> flag(nT, s);
	if (nT <= 32768)
		predicate_warp_kernel<<<(int)std::min<int64_t>((nT + 7) / 8, (int64_t)ctx->sm_count * 8), blk, 0, s>>>(T, nT, o->depth - l,
			bx, by, bz, o->prm.voxel_size, mesh->pred_box.p, mesh->pred_nodes / 2, flag.p);
	else
		predicate_kernel<<<grid_for(ctx, nT, blk, 8), blk, 0, s>>>(T, nT, o->depth - l, bx, by, bz, o->prm.voxel_size,
			mesh->pred_box.p, mesh->pred_nodes / 2, flag.p, pred_budget);
	FPOHM_LAUNCH_CHECK(ctx);
	sel.alloc(nT, s);
	DevBuf<int64_t> cnt(1, s);
	size_t tb = 0;
	FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, T, flag.p, sel.p, cnt.p, nT, s));
	DevBuf<uint8_t> tmp((int64_t)tb, s);
	FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, T, flag.p, sel.p, cnt.p, nT, s));
	ctx->launches += 2;
	int64_t np = 0;
	cnt.download(&np, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	return np;
}

// phase 1 with the bbox predicate.  BFS of octree.cpp:648-690: the queue starts with every current leaf
// (roots on a fresh tree) and a cell's children are queued iff its predicate is true, so
//   T_l = old_leaves(l) ∪ children(P_{l-1}),  P_l = {c in T_l : pred(c)}.
// On return P[l] = old internal cells of level l ∪ P_l (unsorted; phase 2 sorts).
void predicate_sets(fpohm_octree *o, const fpohm_mesh *mesh, int stop_extent, bool fresh,
                    std::vector<DevBuf<uint64_t>> &P, std::vector<int64_t> &nP)
{
	fpohm_ctx *ctx = o->ctx;
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	Sorter sorter{ctx, s};
	mesh_ensure_pred(ctx, const_cast<fpohm_mesh *>(mesh), s);
	DevBuf<uint64_t> kids;   // children(P_{l-1})
	int64_t n_kids = 0;
	for (int l = 0; l <= o->depth; ++l) {
		// old leaves / old internal cells of this level
		DevBuf<uint64_t> old_leaves;
		int64_t n_old_leaves = 0;
		const int64_t n_old_int = (!fresh && l < o->n_levels) ? o->lvl_off[l + 1] - o->lvl_off[l] : 0;
		if (fresh) {
			if (l == 0) {
				old_leaves.alloc(o->n_roots, s);
				roots_kernel<<<grid_for(ctx, o->n_roots, blk), blk, 0, s>>>(old_leaves.p, o->roots[0], o->roots[1], o->roots[2]);
				FPOHM_LAUNCH_CHECK(ctx);
				n_old_leaves = o->n_roots;
			}
		} else if (l <= o->n_levels) {
			const int64_t id0 = l == 0 ? 0 : o->n_roots + 8 * o->lvl_off[l - 1];
			const int64_t cnt = l == 0 ? o->n_roots : 8 * (o->lvl_off[l] - o->lvl_off[l - 1]);
			if (cnt > 0) {
				DevBuf<uint64_t> raw(cnt, s);
				old_leaf_codes_kernel<<<grid_for(ctx, cnt, blk), blk, 0, s>>>(o->cell_code.p, o->cell_first_child.p, id0, cnt, raw.p);
				FPOHM_LAUNCH_CHECK(ctx);
				n_old_leaves = sorter.sort_unique(raw, cnt, key_bits(((int64_t)std::max(o->roots[0], std::max(o->roots[1], o->roots[2])) << l) - 1), old_leaves);
			}
		}
		const int64_t nT = n_old_leaves + n_kids;
		const int extent = 1 << (o->depth - l);
		const bool testable = extent > stop_extent && extent > 1; // ghm.cpp:503 ; octree.cpp:670-671
		int64_t np = 0;
		DevBuf<uint64_t> sel;
		if (testable && nT > 0) {
			DevBuf<uint64_t> T(nT, s);
			if (n_old_leaves) FPOHM_CUDA(cudaMemcpyAsync(T.p, old_leaves.p, 8 * (size_t)n_old_leaves, cudaMemcpyDeviceToDevice, s));
			if (n_kids) FPOHM_CUDA(cudaMemcpyAsync(T.p + n_old_leaves, kids.p, 8 * (size_t)n_kids, cudaMemcpyDeviceToDevice, s));
			np = test_cells(o, mesh, l, T.p, nT, sel);
		}
		P.emplace_back(np + n_old_int, s);
		nP.push_back(np + n_old_int);
		if (np) FPOHM_CUDA(cudaMemcpyAsync(P.back().p, sel.p, 8 * (size_t)np, cudaMemcpyDeviceToDevice, s));
		if (n_old_int) FPOHM_CUDA(cudaMemcpyAsync(P.back().p + np, o->icode.p + o->lvl_off[l], 8 * (size_t)n_old_int, cudaMemcpyDeviceToDevice, s));
		n_kids = 8 * np;
		FPOHM_REQUIRE(n_kids < (1ll << 31), FPOHM_ERANGE, "octree: level %d has %lld cells to test", l + 1, (long long)n_kids);
		if (n_kids) {
			kids.alloc(n_kids, s);
			children_kernel<<<grid_for(ctx, n_kids, blk), blk, 0, s>>>(sel.p, np, kids.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		FPOHM_CUDA(cudaStreamSynchronize(s)); // sel / T die here
		if (fresh && n_kids == 0) break;
		if (!fresh && n_kids == 0 && l >= o->n_levels) break;
	}
}


// ---------------------------------------------------------------------------------------------------
// z-slab sharded build.
//
// Rank r of W owns the cells whose z range lies in slab r of the level-`ls` grid; slabs are cut on level-ls cell
// boundaries, so for l > ls a whole sibling family has one owner and pairing never crosses ranks.  Levels <= ls are
// REPLICATED: every rank computes them in full (at most (8 W)^3 cells).  The only constraints that cross a slab face
// are the grading candidates parents(N18(c)) of boundary cells — the halo.  Per level of phase 2 each rank emits the
// candidates it does not own (`outgoing`), the host all-gathers them (NCCL; this library does not link it), and each
// rank keeps what it owns from everybody's emissions.  The fix-point is a set, so the union over ranks of the closed
// sets is bit-identical to the single-GPU result; phase 3 then numbers the gathered sets canonically.
struct SlabBounds { int32_t W; int32_t b[65]; };   // slab r = level-ls z layers [b[r], b[r+1])

__global__ void owner_flag_kernel(const uint64_t *__restrict__ code, int64_t n, int dl /* l - ls */, SlabBounds sb, int rank,
                                  uint8_t *__restrict__ own)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const uint64_t c = code[i];
		uint8_t f = 0;
		if (c != INVALID) {
			const int zs = (int)(compact1by2(c >> 2) >> dl);
			f = (zs >= sb.b[rank] && zs < sb.b[rank + 1]) ? 1 : 0;
		}
		own[i] = f;
	}
}
__global__ void invalid_flag_kernel(const uint64_t *__restrict__ code, int64_t n, const uint8_t *__restrict__ own, uint8_t *__restrict__ foreign) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		foreign[i] = (code[i] != INVALID && !own[i]) ? 1 : 0;
}

int64_t select_flagged(fpohm_ctx *ctx, cudaStream_t s, const uint64_t *in, const uint8_t *flag, int64_t n, DevBuf<uint64_t> &out) {
	out.alloc(n, s);
	if (n == 0) return 0;
	DevBuf<int64_t> cnt(1, s);
	size_t tb = 0;
	FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, in, flag, out.p, cnt.p, n, s));
	DevBuf<uint8_t> tmp((int64_t)tb, s);
	FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, in, flag, out.p, cnt.p, n, s));
	ctx->launches += 2;
	int64_t m = 0;
	cnt.download(&m, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	return m;
}

} // namespace

struct fpohm_octree_shard {
	fpohm_octree *o = nullptr;          // geometry + (after finish) nothing else: the numbered tree is a fresh object
	const fpohm_mesh *mesh = nullptr;
	int rank = 0, world = 1;
	int ls = 0;                          // last replicated level
	SlabBounds sb{};
	std::vector<DevBuf<uint64_t>> P, I;  // per level: P = predicate-true (owned, or all for l <= ls); I = closed (same)
	std::vector<int64_t> nP, nI;
	std::vector<uint8_t> closed;
	int gmax = -1;                       // global deepest level with predicate-true cells (set by the first level_outgoing)
	// per-level scratch between level_outgoing and level_close
	int pending_level = -1;
	DevBuf<uint64_t> own_cand, out_cand;
	int64_t n_own = 0, n_out = 0;
	~fpohm_octree_shard() { delete o; }
};

namespace {

void shard_split(fpohm_octree_shard *sh, int l, const uint64_t *cand, int64_t n, DevBuf<uint64_t> &own, int64_t &n_own,
                 DevBuf<uint64_t> *foreign, int64_t *n_foreign)
{
	fpohm_ctx *ctx = sh->o->ctx;
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	DevBuf<uint8_t> f(n, s);
	if (n) {
		owner_flag_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(cand, n, l - sh->ls, sh->sb, sh->rank, f.p);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	n_own = select_flagged(ctx, s, cand, f.p, n, own);
	if (foreign) {
		DevBuf<uint8_t> g(n, s);
		if (n) {
			invalid_flag_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(cand, n, f.p, g.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		*n_foreign = select_flagged(ctx, s, cand, g.p, n, *foreign);
	}
}

} // namespace

extern "C" {

int fpohm_octree_grid_setup(const double *V, int64_t nV, int32_t num_voxels, fpohm_octree_params *p) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(V && p && nV > 0 && num_voxels > 0, FPOHM_EINVAL, "fpohm_octree_grid_setup: bad argument");
	// GEO::get_bbox + ghm.cpp:463-493, expression order preserved
	double mn[3] = {V[0], V[1], V[2]}, mx[3] = {V[0], V[1], V[2]};
	for (int64_t i = 1; i < nV; ++i)
		for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], V[3 * i + c]); mx[c] = std::max(mx[c], V[3 * i + c]); }
	double center[3], extent[3];
	for (int c = 0; c < 3; ++c) { center[c] = (mn[c] + mx[c]) / 2; extent[c] = mx[c] - mn[c]; }
	const double max_extent = std::max(extent[0], std::max(extent[1], extent[2]));
	const double voxel_size = max_extent / num_voxels;
	auto next_pow2 = [](unsigned x) { x -= 1; x |= (x >> 1); x |= (x >> 2); x |= (x >> 4); x |= (x >> 8); x |= (x >> 16); return x + 1; };
	for (int c = 0; c < 3; ++c) {
		const double origin = mn[c] - 0 * voxel_size * 1.0; // padding = 0
		const unsigned gs = next_pow2((unsigned)(std::ceil(extent[c] / voxel_size) + 2 * 0));
		FPOHM_REQUIRE(gs >= 1 && gs <= (1u << 21), FPOHM_ERANGE, "fpohm_octree_grid_setup: grid size %u out of range", gs);
		const double origin_max = origin + voxel_size * (int)gs;
		const double origin_center = (origin_max + origin) * 0.5;
		p->grid_size[c] = (int32_t)gs;
		p->origin[c] = origin;
		p->mesh_transform[c] = center[c] - origin_center;
	}
	p->voxel_size = voxel_size;
	FPOHM_API_END
}

int fpohm_octree_build(fpohm_ctx *ctx, const fpohm_mesh *mesh, const fpohm_octree_params *p, fpohm_octree **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && p && out, FPOHM_EINVAL, "fpohm_octree_build: null argument");
	FPOHM_REQUIRE(p->voxel_size > 0, FPOHM_EINVAL, "fpohm_octree_build: voxel_size must be positive");
	DeviceGuard g(ctx->device);
	fpohm_octree *o = new fpohm_octree;
	try {
		o->ctx = ctx; o->prm = *p;
		setup_geometry(o, p->grid_size);
		KernelTimer t(ctx, ctx->stream);
		static const bool timeline = getenv("FPOHM_OCTREE_TIMELINE") != nullptr;     // debug: host wall clock per phase on stderr
		const auto t0 = std::chrono::steady_clock::now();
		std::vector<DevBuf<uint64_t>> P; std::vector<int64_t> nP;
		predicate_sets(o, mesh, p->stop_extent, true, P, nP);
		if (timeline) cudaStreamSynchronize(ctx->stream);
		const auto t1 = std::chrono::steady_clock::now();
		close_and_number(o, P, nP);
		if (timeline) cudaStreamSynchronize(ctx->stream);
		const auto t2 = std::chrono::steady_clock::now();
		t.stop();
		if (timeline) {
			fprintf(stderr, "[fpohm octree] phase 1 %.2f ms, phases 2+3 %.2f ms (closure %.2f, numbering %.2f), launches so far %lld; cudaMallocAsync %.2f ms in %lld calls\n",
			        std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(),
			        o->dbg_close_ms, o->dbg_number_ms, (long long)ctx->launches, g_alloc_ms, g_alloc_calls);
			g_alloc_ms = 0; g_alloc_calls = 0;
		}
	} catch (...) { delete o; throw; }
	*out = o;
	FPOHM_API_END
}

int fpohm_octree_build_from_marks(fpohm_ctx *ctx, const int32_t grid_size[3], const int32_t *marks, int64_t n_marks,
                                  int32_t graded, int32_t paired, fpohm_octree **out)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && grid_size && out && (marks || n_marks == 0), FPOHM_EINVAL, "fpohm_octree_build_from_marks: null argument");
	DeviceGuard g(ctx->device);
	fpohm_octree *o = new fpohm_octree;
	try {
		o->ctx = ctx;
		o->prm.graded = graded; o->prm.paired = paired; o->prm.voxel_size = 1; o->prm.stop_extent = 1;
		setup_geometry(o, grid_size);
		// host: the BFS only tests roots and children of predicate-true cells (octree.cpp:669-678)
		std::vector<std::set<uint64_t>> M((size_t)o->depth + 1);
		for (int64_t i = 0; i < n_marks; ++i) {
			const int32_t x = marks[4 * i], y = marks[4 * i + 1], z = marks[4 * i + 2], e = marks[4 * i + 3];
			if (e <= 1 || (e & (e - 1)) || e > (1 << o->depth)) continue;
			const int sh = ilog2(e), l = o->depth - sh;
			if (x < 0 || y < 0 || z < 0 || x >= grid_size[0] || y >= grid_size[1] || z >= grid_size[2]) continue;
			if ((x & (e - 1)) || (y & (e - 1)) || (z & (e - 1))) continue;
			M[(size_t)l].insert(morton3(x >> sh, y >> sh, z >> sh));
		}
		std::vector<std::vector<uint64_t>> Ph;
		for (int l = 0; l <= o->depth; ++l) {
			std::vector<uint64_t> cur;
			for (uint64_t c : M[(size_t)l])
				if (l == 0 || std::binary_search(Ph[(size_t)l - 1].begin(), Ph[(size_t)l - 1].end(), c >> 3)) cur.push_back(c);
			if (cur.empty()) break;
			Ph.push_back(std::move(cur));
		}
		std::vector<DevBuf<uint64_t>> P; std::vector<int64_t> nP;
		for (auto &v : Ph) {
			P.emplace_back((int64_t)v.size(), ctx->stream);
			P.back().upload(v.data(), (int64_t)v.size());
			nP.push_back((int64_t)v.size());
		}
		FPOHM_CUDA(cudaStreamSynchronize(ctx->stream));
		KernelTimer t(ctx, ctx->stream);
		close_and_number(o, P, nP);
		t.stop();
	} catch (...) { delete o; throw; }
	*out = o;
	FPOHM_API_END
}

int fpohm_octree_subdivide(fpohm_octree *o, const fpohm_mesh *mesh, int32_t stop_extent) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o && mesh, FPOHM_EINVAL, "fpohm_octree_subdivide: null argument");
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	KernelTimer t(ctx, ctx->stream);
	std::vector<DevBuf<uint64_t>> P; std::vector<int64_t> nP;
	predicate_sets(o, mesh, stop_extent, false, P, nP);
	close_and_number(o, P, nP);
	t.stop();
	FPOHM_API_END
}

int fpohm_octree_refine(fpohm_octree *o, const fpohm_mesh *mesh, const int32_t *cell_ids, int64_t n, int32_t stop_extent) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o && mesh && (cell_ids || n == 0), FPOHM_EINVAL, "fpohm_octree_refine: null argument");
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	// listed LEAF cells, per level.  The list is the pipeline's tb_subdivided_cells (short); its (level, code, leaf) triples are
	// gathered on the device — the first version downloaded the whole cell table for this (13 B x n_cells: 740 MB at the C3 size,
	// once per pass of the outer loop, ghm.cpp:518-521).
	for (int64_t i = 0; i < n; ++i)
		FPOHM_REQUIRE(cell_ids[i] >= 0 && cell_ids[i] < o->n_cells, FPOHM_EINVAL, "fpohm_octree_refine: cell id %d out of range", cell_ids[i]);
	std::vector<uint8_t> lvl((size_t)n);
	std::vector<uint64_t> code((size_t)n);
	std::vector<int32_t> fc((size_t)n);
	if (n > 0) {
		DevBuf<int32_t> dids(n, s), gfc(n, s);
		DevBuf<uint8_t> glvl(n, s);
		DevBuf<uint64_t> gcode(n, s);
		dids.upload(cell_ids, n);
		gather_cells_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(dids.p, n, o->cell_level.p, o->cell_code.p, o->cell_first_child.p, glvl.p, gcode.p, gfc.p);
		FPOHM_LAUNCH_CHECK(ctx);
		glvl.download(lvl.data(), n); gcode.download(code.data(), n); gfc.download(fc.data(), n);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	std::vector<std::set<uint64_t>> L((size_t)o->depth + 1);
	for (int64_t i = 0; i < n; ++i) {
		if (fc[(size_t)i] >= 0) continue; // octree.cpp:694: only leaves are queued
		const int extent = 1 << (o->depth - lvl[(size_t)i]);
		if (extent <= stop_extent || extent <= 1) continue;
		L[lvl[(size_t)i]].insert(code[(size_t)i]);
	}
	mesh_ensure_pred(ctx, const_cast<fpohm_mesh *>(mesh), s);
	// P_l = old internal cells of level l  ∪  listed leaves whose predicate is true
	std::vector<DevBuf<uint64_t>> P; std::vector<int64_t> nP;
	const int top = std::max(o->n_levels, o->depth + 1);
	for (int l = 0; l < top; ++l) {
		const int64_t n_old = l < o->n_levels ? o->lvl_off[l + 1] - o->lvl_off[l] : 0;
		std::vector<uint64_t> cand;
		if (l <= o->depth) cand.assign(L[(size_t)l].begin(), L[(size_t)l].end());
		DevBuf<uint64_t> merged(n_old + (int64_t)cand.size(), s);
		if (n_old) FPOHM_CUDA(cudaMemcpyAsync(merged.p, o->icode.p + o->lvl_off[l], 8 * (size_t)n_old, cudaMemcpyDeviceToDevice, s));
		int64_t n_new = 0;
		if (!cand.empty()) {
			DevBuf<uint64_t> dc((int64_t)cand.size(), s), sel;
			dc.upload(cand.data(), (int64_t)cand.size());
			n_new = test_cells(o, mesh, l, dc.p, (int64_t)cand.size(), sel);      // warp-per-cell kernel for short lists, FPOHM_PRED_BUDGET honoured
			if (n_new) FPOHM_CUDA(cudaMemcpyAsync(merged.p + n_old, sel.p, 8 * (size_t)n_new, cudaMemcpyDeviceToDevice, s));
			FPOHM_CUDA(cudaStreamSynchronize(s));
		}
		merged.n = n_old + n_new;
		nP.push_back(n_old + n_new);
		P.push_back(std::move(merged));
	}
	KernelTimer t(ctx, s);
	close_and_number(o, P, nP);
	t.stop();
	FPOHM_API_END
}


// ---- z-slab sharded build ---------------------------------------------------------------------------------------------
int fpohm_octree_shard_create(fpohm_ctx *ctx, const fpohm_mesh *mesh, const fpohm_octree_params *p, int32_t rank, int32_t world,
                              fpohm_octree_shard **out)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && p && out, FPOHM_EINVAL, "fpohm_octree_shard_create: null argument");
	FPOHM_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, FPOHM_EINVAL, "fpohm_octree_shard_create: rank %d / world %d", rank, world);
	FPOHM_REQUIRE(p->voxel_size > 0, FPOHM_EINVAL, "fpohm_octree_shard_create: voxel_size must be positive");
	DeviceGuard g(ctx->device);
	fpohm_octree_shard *sh = new fpohm_octree_shard;
	try {
		sh->o = new fpohm_octree;
		sh->o->ctx = ctx; sh->o->prm = *p;
		setup_geometry(sh->o, p->grid_size);
		sh->mesh = mesh; sh->rank = rank; sh->world = world;
		// last replicated level.  Balance wants >= 8 z layers per rank at level ls (the cut is weighted by layer), but the
		// two deepest predicate levels hold ~94 % of the work and must stay sharded; at least one layer per rank if possible.
		const int depth = sh->o->depth;
		auto first_level_with = [&](int64_t layers) { int l = 0; while (((int64_t)sh->o->roots[2] << l) < layers && l < depth) ++l; return l; };
		int ltest = -1;                                  // deepest level the predicate is evaluated on (ghm.cpp:503, octree.cpp:670)
		for (int l = 0; l <= depth; ++l) { const int e = 1 << (depth - l); if (e > p->stop_extent && e > 1) ltest = l; }
		int ls = std::min(std::max(ltest - 2, first_level_with(world)), first_level_with(8ll * world));
		ls = std::max(0, std::min(ls, depth - 1));
		sh->ls = ls;
	} catch (...) { delete sh; throw; }
	*out = sh;
	FPOHM_API_END
}

void fpohm_octree_shard_free(fpohm_octree_shard *sh) {
	if (!sh) return;
	DeviceGuard g(sh->o->ctx->device);
	cudaStreamSynchronize(sh->o->ctx->stream);
	delete sh;
}

// phase 1: replicated down to level ls, then only the owned slab.  No communication.
int fpohm_octree_shard_refine(fpohm_octree_shard *sh, int32_t *local_max_level) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh && local_max_level, FPOHM_EINVAL, "fpohm_octree_shard_refine: null argument");
	fpohm_octree *o = sh->o;
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	mesh_ensure_pred(ctx, const_cast<fpohm_mesh *>(sh->mesh), s);
	sh->P.clear(); sh->nP.clear();
	DevBuf<uint64_t> T(o->n_roots, s);
	int64_t nT = o->n_roots;
	roots_kernel<<<grid_for(ctx, o->n_roots, blk), blk, 0, s>>>(T.p, o->roots[0], o->roots[1], o->roots[2]);
	FPOHM_LAUNCH_CHECK(ctx);
	const int nz_s = o->roots[2] << sh->ls;
	// default bounds: equal layer counts (replaced below by the weighted cut when level ls has predicate-true cells)
	sh->sb.W = sh->world;
	for (int r = 0; r <= sh->world; ++r) sh->sb.b[r] = (int32_t)((int64_t)nz_s * r / sh->world);
	int lmax = -1;
	for (int l = 0; l <= o->depth; ++l) {
		const int extent = 1 << (o->depth - l);
		const bool testable = extent > o->prm.stop_extent && extent > 1;
		DevBuf<uint64_t> sel;
		int64_t np = 0;
		if (testable && nT > 0) np = test_cells(o, sh->mesh, l, T.p, nT, sel);
		sh->P.emplace_back(np, s);
		sh->nP.push_back(np);
		if (np) { FPOHM_CUDA(cudaMemcpyAsync(sh->P.back().p, sel.p, 8 * (size_t)np, cudaMemcpyDeviceToDevice, s)); lmax = l; }
		if (l == sh->ls && np > 0) {
			// weighted slab cut: balance the number of predicate-true level-ls cells per rank (identical on every rank:
			// level ls is replicated).  The surface is where all the deeper work is.
			std::vector<uint64_t> h((size_t)np);
			sel.download(h.data(), np);
			FPOHM_CUDA(cudaStreamSynchronize(s));
			std::vector<int64_t> hist((size_t)nz_s + 1, 0);
			for (uint64_t c : h) hist[(size_t)compact1by2(c >> 2) + 1]++;
			for (int z = 0; z < nz_s; ++z) hist[(size_t)z + 1] += hist[(size_t)z];
			for (int r = 1; r < sh->world; ++r) {
				const int64_t want = np * r / sh->world;
				int z = (int)(std::lower_bound(hist.begin(), hist.end(), want) - hist.begin());
				if (z > 0 && want - hist[(size_t)z - 1] < hist[(size_t)std::min(z, nz_s)] - want) --z;   // nearest layer boundary
				z = std::max(z, sh->sb.b[r - 1]);
				sh->sb.b[r] = std::min(z, nz_s);
			}
			sh->sb.b[0] = 0; sh->sb.b[sh->world] = nz_s;
		}
		int64_t n_kids = 8 * np;
		FPOHM_REQUIRE(n_kids < (1ll << 31), FPOHM_ERANGE, "octree shard: level %d has %lld cells to test", l + 1, (long long)n_kids);
		if (n_kids == 0) { nT = 0; if (l >= sh->ls) break; else continue; }
		DevBuf<uint64_t> kids(n_kids, s);
		children_kernel<<<grid_for(ctx, n_kids, blk), blk, 0, s>>>(sel.p, np, kids.p);
		FPOHM_LAUNCH_CHECK(ctx);
		if (l == sh->ls) {
			// entering the sharded levels: keep the children this rank owns
			int64_t n_own = 0;
			shard_split(sh, l + 1, kids.p, n_kids, T, n_own, nullptr, nullptr);
			nT = n_own;
		} else {
			T = std::move(kids);
			nT = n_kids;
		}
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	while ((int)sh->P.size() <= o->depth) { sh->P.emplace_back((int64_t)0, s); sh->nP.push_back(0); }
	*local_max_level = lmax;
	sh->gmax = -1;
	FPOHM_API_END
}

int fpohm_octree_shard_info(const fpohm_octree_shard *sh, int32_t *replicated_levels, int32_t *slab_bounds /*world+1, level-ls z layers*/,
                            int64_t *owned_true_cells)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh, FPOHM_EINVAL, "fpohm_octree_shard_info: null shard");
	if (replicated_levels) *replicated_levels = sh->ls + 1;
	if (slab_bounds) for (int r = 0; r <= sh->world; ++r) slab_bounds[r] = sh->sb.b[r];
	if (owned_true_cells) {
		int64_t t = 0;
		for (int l = sh->ls + 1; l < (int)sh->nP.size(); ++l) t += sh->nP[l];
		*owned_true_cells = t;
	}
	FPOHM_API_END
}

// phase 2, level l (call for l = global_max_level ... 0 in order): candidates of level l that other ranks must see.
//   l >  ls : the forced cells this rank does not own (halo)            -> all-gather -> level_close
//   l == ls : every forced cell of the owned level ls+1 set              -> all-gather -> level_close (replicated close)
//   l <  ls : nothing (level l+1 is replicated, every rank derives the same candidates)
int fpohm_octree_shard_level_outgoing(fpohm_octree_shard *sh, int32_t global_max_level, int32_t level, int64_t *n_out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh && n_out, FPOHM_EINVAL, "fpohm_octree_shard_level_outgoing: null argument");
	fpohm_octree *o = sh->o;
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	FPOHM_REQUIRE(level >= 0 && level <= global_max_level && global_max_level <= o->depth, FPOHM_EINVAL,
	              "fpohm_octree_shard_level_outgoing: level %d / max %d", level, global_max_level);
	if (sh->gmax < 0) {
		sh->gmax = global_max_level;
		sh->I.clear(); sh->nI.clear();
		for (int l = 0; l <= global_max_level; ++l) { sh->I.emplace_back((int64_t)0, s); sh->nI.push_back(0); }
		sh->closed.assign((size_t)global_max_level + 1, 0);
	}
	FPOHM_REQUIRE(global_max_level == sh->gmax, FPOHM_ESTATE, "fpohm_octree_shard_level_outgoing: max level changed");
	FPOHM_REQUIRE(level == sh->gmax || sh->closed[(size_t)level + 1], FPOHM_ESTATE,
	              "fpohm_octree_shard_level_outgoing: level %d before level %d was closed", level, level + 1);
	const bool has_up = level < sh->gmax;
	DevBuf<uint64_t> cand;
	const int64_t n_cand = level_candidates(o, level, sh->P[(size_t)level].p, sh->nP[(size_t)level],
	                                        has_up ? sh->I[(size_t)level + 1].p : nullptr, has_up ? sh->nI[(size_t)level + 1] : 0, cand);
	Sorter sorter{ctx, s};
	const int lbits = key_bits(((int64_t)std::max(o->roots[0], std::max(o->roots[1], o->roots[2])) << level) - 1);
	DevBuf<uint64_t> uq;
	const int64_t m = sorter.sort_unique(cand, n_cand, lbits, uq);
	if (level > sh->ls) {
		shard_split(sh, level, uq.p, m, sh->own_cand, sh->n_own, &sh->out_cand, &sh->n_out);
	} else if (level == sh->ls) {
		// P_ls is replicated, the forced part is not: ship everything (duplicates of P_ls across ranks are harmless)
		sh->own_cand.alloc(0, s); sh->n_own = 0;
		sh->out_cand = std::move(uq); sh->n_out = m;
	} else {
		sh->own_cand = std::move(uq); sh->n_own = m;
		sh->out_cand.alloc(0, s); sh->n_out = 0;
	}
	sh->pending_level = level;
	*n_out = sh->n_out;
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_octree_shard_outgoing_copy(fpohm_octree_shard *sh, uint64_t *dst_dev) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh && (dst_dev || sh->n_out == 0), FPOHM_EINVAL, "fpohm_octree_shard_outgoing_copy: null argument");
	FPOHM_REQUIRE(sh->pending_level >= 0, FPOHM_ESTATE, "fpohm_octree_shard_outgoing_copy: no pending level");
	DeviceGuard g(sh->o->ctx->device);
	if (sh->n_out) FPOHM_CUDA(cudaMemcpyAsync(dst_dev, sh->out_cand.p, 8 * (size_t)sh->n_out, cudaMemcpyDeviceToDevice, sh->o->ctx->stream));
	FPOHM_CUDA(cudaStreamSynchronize(sh->o->ctx->stream));
	FPOHM_API_END
}

// close level l with everybody's outgoing candidates (the concatenation of all ranks' buffers, this rank's included)
int fpohm_octree_shard_level_close(fpohm_octree_shard *sh, int32_t level, const uint64_t *gathered_dev, int64_t n_gathered, int64_t *n_closed) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh && (gathered_dev || n_gathered == 0), FPOHM_EINVAL, "fpohm_octree_shard_level_close: null argument");
	FPOHM_REQUIRE(level == sh->pending_level, FPOHM_ESTATE, "fpohm_octree_shard_level_close: level %d is not the pending level %d", level, sh->pending_level);
	fpohm_octree *o = sh->o;
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<uint64_t> mine;
	int64_t n_mine = 0;
	if (level > sh->ls) {
		if (n_gathered) shard_split(sh, level, gathered_dev, n_gathered, mine, n_mine, nullptr, nullptr);
	} else if (level == sh->ls) {
		mine.alloc(n_gathered, s);
		if (n_gathered) FPOHM_CUDA(cudaMemcpyAsync(mine.p, gathered_dev, 8 * (size_t)n_gathered, cudaMemcpyDeviceToDevice, s));
		n_mine = n_gathered;
	}
	DevBuf<uint64_t> cand(sh->n_own + n_mine, s);
	if (sh->n_own) FPOHM_CUDA(cudaMemcpyAsync(cand.p, sh->own_cand.p, 8 * (size_t)sh->n_own, cudaMemcpyDeviceToDevice, s));
	if (n_mine) FPOHM_CUDA(cudaMemcpyAsync(cand.p + sh->n_own, mine.p, 8 * (size_t)n_mine, cudaMemcpyDeviceToDevice, s));
	sh->nI[(size_t)level] = close_level(o, level, cand, sh->n_own + n_mine, sh->I[(size_t)level]);
	sh->closed[(size_t)level] = 1;
	sh->pending_level = -1;
	sh->own_cand.release(); sh->out_cand.release();
	sh->n_own = sh->n_out = 0;
	if (n_closed) *n_closed = sh->nI[(size_t)level];
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

// owned closed cells of a sharded level (l > ls), for the final gather; replicated levels report 0 (every rank has them)
int fpohm_octree_shard_level_result(const fpohm_octree_shard *sh, int32_t level, uint64_t *dst_dev, int64_t *n) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh && n, FPOHM_EINVAL, "fpohm_octree_shard_level_result: null argument");
	FPOHM_REQUIRE(level >= 0 && level <= sh->gmax && sh->closed[(size_t)level], FPOHM_ESTATE, "fpohm_octree_shard_level_result: level %d not closed", level);
	const int64_t m = level > sh->ls ? sh->nI[(size_t)level] : 0;
	*n = m;
	if (dst_dev && m) {
		DeviceGuard g(sh->o->ctx->device);
		FPOHM_CUDA(cudaMemcpyAsync(dst_dev, sh->I[(size_t)level].p, 8 * (size_t)m, cudaMemcpyDeviceToDevice, sh->o->ctx->stream));
		FPOHM_CUDA(cudaStreamSynchronize(sh->o->ctx->stream));
	}
	FPOHM_API_END
}

// phase 3 on this rank: `gathered[l]` = concatenation over ranks of level_result(l) for every sharded level
// (entries for replicated levels are ignored and may be NULL).  Returns the complete, canonically numbered octree.
int fpohm_octree_shard_finish(fpohm_octree_shard *sh, const uint64_t *const *gathered_dev, const int64_t *counts, fpohm_octree **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(sh && out, FPOHM_EINVAL, "fpohm_octree_shard_finish: null argument");
	fpohm_ctx *ctx = sh->o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	for (int l = 0; l <= sh->gmax; ++l)
		FPOHM_REQUIRE(sh->closed[(size_t)l], FPOHM_ESTATE, "fpohm_octree_shard_finish: level %d not closed", l);
	fpohm_octree *o = new fpohm_octree;
	try {
		o->ctx = ctx; o->prm = sh->o->prm;
		setup_geometry(o, o->prm.grid_size);
		std::vector<DevBuf<uint64_t>> I((size_t)sh->gmax + 1);
		std::vector<int64_t> nI((size_t)sh->gmax + 1, 0);
		Sorter sorter{ctx, s};
		for (int l = 0; l <= sh->gmax; ++l) {
			if (l <= sh->ls) {
				nI[(size_t)l] = sh->nI[(size_t)l];
				I[(size_t)l].alloc(nI[(size_t)l], s);
				if (nI[(size_t)l]) FPOHM_CUDA(cudaMemcpyAsync(I[(size_t)l].p, sh->I[(size_t)l].p, 8 * (size_t)nI[(size_t)l], cudaMemcpyDeviceToDevice, s));
			} else {
				FPOHM_REQUIRE(gathered_dev && counts && (counts[l] == 0 || gathered_dev[l]), FPOHM_EINVAL, "fpohm_octree_shard_finish: level %d missing", l);
				DevBuf<uint64_t> in(counts[l], s);
				if (counts[l]) FPOHM_CUDA(cudaMemcpyAsync(in.p, gathered_dev[l], 8 * (size_t)counts[l], cudaMemcpyDeviceToDevice, s));
				const int lbits = key_bits(((int64_t)std::max(o->roots[0], std::max(o->roots[1], o->roots[2])) << l) - 1);
				nI[(size_t)l] = sorter.sort_unique(in, counts[l], lbits, I[(size_t)l]);   // slab order -> Morton order
			}
		}
		KernelTimer t(ctx, s);
		number_levels(o, I, nI);
		t.stop();
	} catch (...) { delete o; throw; }
	*out = o;
	FPOHM_API_END
}

void fpohm_octree_free(fpohm_octree *o) {
	if (!o) return;
	DeviceGuard g(o->ctx->device);
	cudaStreamSynchronize(o->ctx->stream);
	delete o;
}

int fpohm_octree_sizes(const fpohm_octree *o, int64_t *n_nodes, int64_t *n_cells, int64_t *n_leaves, int32_t *n_roots, int32_t *max_depth) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o, FPOHM_EINVAL, "fpohm_octree_sizes: null octree");
	if (n_nodes) *n_nodes = o->n_nodes;
	if (n_cells) *n_cells = o->n_cells;
	if (n_leaves) *n_leaves = o->n_leaves;
	if (n_roots) *n_roots = o->n_roots;
	if (max_depth) *max_depth = o->depth;
	FPOHM_API_END
}

int fpohm_octree_export(const fpohm_octree *o, int32_t *node_pos, int32_t *node_neigh, int32_t *cell_first_child,
                        int32_t *cell_corner, int32_t *cell_neigh)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o, FPOHM_EINVAL, "fpohm_octree_export: null octree");
	DeviceGuard g(o->ctx->device);
	if (node_pos) o->node_pos.download(node_pos, 3 * o->n_nodes);
	if (node_neigh) o->node_neigh.download(node_neigh, 6 * o->n_nodes);
	if (cell_first_child) o->cell_first_child.download(cell_first_child, o->n_cells);
	if (cell_corner) o->cell_corner.download(cell_corner, 8 * o->n_cells);
	if (cell_neigh) o->cell_neigh.download(cell_neigh, 6 * o->n_cells);
	FPOHM_CUDA(cudaStreamSynchronize(o->ctx->stream));
	FPOHM_API_END
}

int fpohm_octree_hexes(const fpohm_octree *o, double *Vpos, uint32_t *hex, int32_t *hex2cell) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o, FPOHM_EINVAL, "fpohm_octree_hexes: null octree");
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	const double bx = o->prm.mesh_transform[0] + o->prm.origin[0], by = o->prm.mesh_transform[1] + o->prm.origin[1],
	             bz = o->prm.mesh_transform[2] + o->prm.origin[2];
	if (Vpos) {
		DevBuf<double> dV(3 * o->n_nodes, s);
		hex_vertices_kernel<<<grid_for(ctx, 3 * o->n_nodes, blk), blk, 0, s>>>(o->node_pos.p, o->n_nodes, bx, by, bz, o->prm.voxel_size, dV.p);
		FPOHM_LAUNCH_CHECK(ctx);
		dV.download(Vpos, 3 * o->n_nodes);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	if (hex) {
		DevBuf<uint32_t> dH(8 * o->n_leaves, s);
		hex_cells_kernel<<<grid_for(ctx, 8 * o->n_leaves, blk), blk, 0, s>>>(o->leaf_cell.p, o->n_leaves, o->cell_corner.p, dH.p);
		FPOHM_LAUNCH_CHECK(ctx);
		dH.download(hex, 8 * o->n_leaves);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	if (hex2cell) { o->leaf_cell.download(hex2cell, o->n_leaves); FPOHM_CUDA(cudaStreamSynchronize(s)); }
	FPOHM_API_END
}

int fpohm_octree_check(const fpohm_octree *o, int32_t *flags) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o && flags, FPOHM_EINVAL, "fpohm_octree_check: null argument");
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<int32_t> bad(2, s);
	bad.zero();
	check_kernel<<<grid_for(ctx, o->n_cells, 256), 256, 0, s>>>(o->cell_first_child.p, o->cell_corner.p, o->node_neigh.p, o->n_cells, bad.p);
	FPOHM_LAUNCH_CHECK(ctx);
	int32_t h[2] = {0, 0};
	bad.download(h, 2);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	*flags = (h[0] == 0 ? 1 : 0) | (h[1] == 0 ? 2 : 0);
	FPOHM_API_END
}

} // extern "C"

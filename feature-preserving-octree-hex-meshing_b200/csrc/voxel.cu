// z-ray parity inside/outside classification: compute_sign x3.
//   VoxelGrid   voxelization.h:220-272     DexelGrid  voxelization.h:275-331     OctreeGrid cells  voxelization.cpp:101-163
//   intersect_ray_z  voxelization.h:194-217      point_in_triangle_2d / orientation (SoS)  voxelization.cpp:57-93
//   orient_2d_inexact  voxelization.h:169-182    VoxelGrid ctor / voxel_center  voxelization.h:71-91
//
// Two stages.
//   (1) hits.  Regular grids (VoxelGrid, DexelGrid): TRIANGLE-parallel — facet_rect_kernel / pair_hits_kernel below enumerate
//       exactly the (facet, column) pairs the reference's per-column box query would test, run the SoS point-in-triangle
//       at the column centre and append (z, sign) to a fixed-capacity per-column list (overflow is reported, never
//       truncated silently).  Octree cells: one thread per cell, stack descent of the facet-box tree with the cell
//       footprint as query box (cells of all sizes have their own centres, there is no shared column structure).
//   (2) a consumer per flavour.  The VoxelGrid rule "voxel = [ sum of signs of hits with z < centre_z ] < 0" is
//       order independent, so nothing is sorted: each hit becomes an event (first layer k0 above it, sign) and
//       voxel_fill_kernel writes every voxel exactly once (1 B/voxel).  Dexel / octree-cell rules need (z, sign)
//       order: tiny in-register insertion sort.
#include "mesh.h"
#include "octree.h"

#include <cstdlib>
#include <cub/device/device_scan.cuh>

using namespace fpohm;

namespace {

// Hit lists have a per-call capacity: 32 per column / cell to begin with (a handful is typical).  The kernels keep counting past
// the capacity and report the largest count; a call that overflowed is repeated with the next capacity that fits (128, 512,
// 2048) — the reference collects hits in a std::vector (voxelization.h:248-256) and stacked sheets or finned parts do
// cross a column more than 32 times.  Beyond 2048 hits in one column the call fails with FPOHM_ERANGE.
constexpr int HIT_CAP = 32;
constexpr int HIT_CAP_MAX = 2048;
inline int next_hit_cap(int need) { for (int c : {32, 128, 512, 2048}) if (need <= c) return c; return 0; }

// voxelization.cpp:57-68
__device__ __forceinline__ int orientation(double x1, double y1, double x2, double y2, double &twice_signed_area) {
	twice_signed_area = y1 * x2 - x1 * y2;
	if (twice_signed_area > 0) return 1;
	else if (twice_signed_area < 0) return -1;
	else if (y2 > y1) return 1;
	else if (y2 < y1) return -1;
	else if (x1 > x2) return 1;
	else if (x1 < x2) return -1;
	else return 0;
}

// voxelization.cpp:74-93
__device__ __forceinline__ bool point_in_triangle_2d(double x0, double y0, double x1, double y1, double x2, double y2,
                                                     double x3, double y3, double &a, double &b, double &c)
{
	x1 -= x0; x2 -= x0; x3 -= x0;
	y1 -= y0; y2 -= y0; y3 -= y0;
	const int signa = orientation(x2, y2, x3, y3, a);
	if (signa == 0) return false;
	const int signb = orientation(x3, y3, x1, y1, b);
	if (signb != signa) return false;
	const int signc = orientation(x1, y1, x2, y2, c);
	if (signc != signa) return false;
	const double sum = a + b + c;
	a /= sum; b /= sum; c /= sum;
	return true;
}

// intersect_ray_z, voxelization.h:194-217
__device__ __forceinline__ int intersect_ray_z(const double *__restrict__ t, double qx, double qy, double &z) {
	double u, v, w;
	if (point_in_triangle_2d(qx, qy, t[0], t[1], t[3], t[4], t[6], t[7], u, v, w)) {
		z = u * t[2] + v * t[5] + w * t[8];
		const double a11 = t[3] - t[0], a12 = t[4] - t[1], a21 = t[6] - t[0], a22 = t[7] - t[1];
		const double delta = a11 * a22 - a12 * a21; // GEO::det2x2
		return delta > 0 ? 1 : (delta < 0 ? -1 : 0);
	}
	return 0;
}

struct ColumnGrid {      // VoxelGrid / DexelGrid columns
	double ox, oy, spacing;
	int nx, ny;
	int cap;             // capacity of a column's hit list in this pass
};

// gather hits of the vertical ray through (qx,qy); facets are pre-filtered by the reference's xy box [bx0,bx1]x[by0,by1]
__device__ __forceinline__ int gather_hits(const double *__restrict__ box, int64_t P, const int32_t *__restrict__ order,
                                           const double *__restrict__ tri, double bx0, double bx1, double by0, double by1,
                                           double qx, double qy, double *hz, int8_t *hs, int cap, bool &overflow)
{
	int64_t stack[40];
	int sp = 0, n = 0;
	stack[sp++] = 1;
	while (sp > 0) {
		const int64_t nd = stack[--sp];
		const double *b = box + 6 * nd;
		if (bx1 < b[0] || bx0 > b[3] || by1 < b[1] || by0 > b[4]) continue; // bboxes_overlap, z always overlaps
		if (nd >= P) {
			double z;
			const int s = intersect_ray_z(tri + 9 * (int64_t)order[nd - P], qx, qy, z);
			if (s) {
				if (n < cap) { hz[n] = z; hs[n] = (int8_t)s; } else overflow = true;
				++n;      // keeps counting: the caller learns how much room the retry needs
			}
			continue;
		}
		stack[sp++] = 2 * nd + 1;
		stack[sp++] = 2 * nd;
	}
	return n;
}

// Programmatic dependent launch (sm_90+): a kernel that opens with pdl_release() lets the NEXT kernel of the stream be scheduled as
// soon as every CTA of this one has started; the next kernel's CTAs then sit in pdl_wait() until this grid has completed and its
// writes are visible.  Launch latency and CTA ramp-up of the VoxelGrid pass's four kernels hide behind each other's tails.
// Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// std::sort of pair<double,int>: ascending z, then ascending sign
__device__ __forceinline__ void sort_hits(double *hz, int8_t *hs, int n) {
	for (int i = 1; i < n; ++i) {
		const double z = hz[i]; const int8_t s = hs[i];
		int j = i - 1;
		while (j >= 0 && (hz[j] > z || (hz[j] == z && hs[j] > s))) { hz[j + 1] = hz[j]; hs[j + 1] = hs[j]; --j; }
		hz[j + 1] = z; hs[j + 1] = s;
	}
}

// ---- hits of all columns of a regular grid, TRIANGLE-parallel ------------------------------------------------------
// The reference gathers, per column, the facets whose box contains the column's (x,y) (a degenerate query box) and
// tests those.  Seen from the facet: facet f is tested by exactly the columns whose centre lies inside its closed xy
// box — an index rectangle [x0,x1] x [y0,y1] found with the reference's own centre expression (x + 0.5) * spacing + ox.
// (1) rect kernel: rectangle + pair count per facet; (2) exclusive scan; (3) pair kernel: thread t finds its facet by
// binary search in the scan and its column inside the rectangle, runs intersect_ray_z, and appends a hit to the
// column's fixed-capacity list with one atomicAdd.  Work is proportional to the number of (facet, column) pairs the
// reference tests, instead of one tree descent per column (first version: 6.9 of 7.1 ms at 1024^2 columns, 2 M facets).
// first layer index k in [0, nz] with  hit_z < (k + 0.5) * spacing + oz   (exactly the comparison of voxelization.h:259-261)
// The three index searches below decide with the reference's own floating-point expression.  t = (value - origin) / spacing - 0.5
// is the same quantity in real numbers; unless t lies within INDEX_MARGIN of an integer (or is huge) the rounding of either
// expression (a few ulps of its operands, i.e. < 1e-6 cells for |t| < 1e9) cannot change the outcome, and the answer follows from
// floor(t) alone.  Only the near-integer cases run the exact comparison loops (ncu: the loops were 21 % of facet_rect_kernel).
#define INDEX_MARGIN 1e-3
__device__ __forceinline__ bool index_safe(double t, double ft) { return t - ft > INDEX_MARGIN && t - ft < 1.0 - INDEX_MARGIN && fabs(t) < 1e9; }
__device__ __forceinline__ int first_layer_above(double z, double oz, double spacing, int nz) {
	const double t = (z - oz) * __drcp_rn(spacing) - 0.5, ft = floor(t);
	if (index_safe(t, ft)) return max(0, min(nz, (int)ft + 1));        // smallest k with t < k
	int k = (int)fmax(-1.0, fmin(ft, 2147483000.0));                   // guess only: the two loops below decide with the exact expression
	if (k < 0) k = 0;
	if (k > nz) k = nz;
	while (k > 0 && z < ((k - 1) + 0.5) * spacing + oz) --k;
	while (k < nz && !(z < (k + 0.5) * spacing + oz)) ++k;
	return k;
}

__device__ __forceinline__ int first_center_ge(double lo, double o, double sp, int n) {
	// smallest index i in [0, n] with (i + 0.5) * sp + o >= lo
	const double t = (lo - o) * __drcp_rn(sp) - 0.5, ft = floor(t);
	if (index_safe(t, ft)) return max(0, min(n, (int)ft + 1));         // smallest i with i >= t
	int i = (int)fmax(-1.0, fmin(ft, 2147483000.0));                   // guess only
	if (i < 0) i = 0;
	if (i > n) i = n;
	while (i > 0 && ((i - 1) + 0.5) * sp + o >= lo) --i;
	while (i < n && !((i + 0.5) * sp + o >= lo)) ++i;
	return i;
}
__device__ __forceinline__ int last_center_le(double hi, double o, double sp, int n) {
	// largest index i in [-1, n-1] with (i + 0.5) * sp + o <= hi
	const double t = (hi - o) * __drcp_rn(sp) - 0.5, ft = floor(t);
	if (index_safe(t, ft)) return max(-1, min(n - 1, (int)ft));        // largest i with i <= t
	int i = (int)fmax(-2.0, fmin(ft, 2147483000.0));                   // guess only
	if (i < -1) i = -1;
	if (i > n - 1) i = n - 1;
	while (i < n - 1 && ((i + 1) + 0.5) * sp + o <= hi) ++i;
	while (i >= 0 && !((i + 0.5) * sp + o <= hi)) --i;
	return i;
}

// append one hit to its column (VoxelGrid: packed (k0, sign) event; otherwise (z, sign))
__device__ __forceinline__ void append_hit(const ColumnGrid &g, int x, int y, double z, int s, double *__restrict__ hit_z, int8_t *__restrict__ hit_s,
                                           int32_t *__restrict__ hit_n, int32_t *__restrict__ overflow_flag, int32_t *__restrict__ hit_ev, double oz, int nz)
{
	const int64_t col = (int64_t)y * g.nx + x;
	const int cap = g.cap;
	const int slot = atomicAdd(&hit_n[col], 1);
	if (slot >= cap) atomicMax(overflow_flag, slot + 1);      // largest count seen: what the retry must hold
	else if (hit_ev) hit_ev[(int64_t)slot * ((int64_t)g.nx * g.ny) + col] = (first_layer_above(z, oz, g.spacing, nz) << 2) | (s + 1);   // [slot][column]: the summary reads a slot of neighbouring columns in one line
	else { hit_z[col * cap + slot] = z; hit_s[col * cap + slot] = (int8_t)s; }
}

// Rectangle of columns per facet.  Facets covering at most RECT_INLINE columns (the common case on fine meshes) are
// finished right here; larger ones are left to the load-balanced pair kernel through a compact list.
// The small rectangles of a warp's 32 facets are pooled: a warp prefix sum over the pair counts, then the lanes take
// the (facet, column) pairs 32 at a time whatever facet they belong to.  (ncu on the first version, where each lane
// looped over its own facet's columns: 8.8 of 32 lanes active, 154 M warp instructions, 0.29 of the 0.77 ms at 1024^3.)
// Big facets: a warp reserves list slots AND pair offsets for its big facets with ONE 64-bit atomicAdd on
// (count << BIG_PAIR_BITS | pairs), so the list is ordered by pair offset whatever order the warps arrive in and the pair
// kernel can binary-search it — no scan over all facets, no per-facet rect / count arrays (round 1 wrote 32 B per facet and
// ran a cub scan of nF + 1 counts: two more launches in front of a pair kernel that has nothing to do on fine meshes).
// Limits, both checked: nF < 2^28, and < 2^36 big pairs in total (ctl[4..5] holds the exact sum).
#define RECT_INLINE 16
#define BIG_PAIR_BITS 36
#define HIT_CTL_WORDS 8     // ctl[0] largest hit count past the capacity, ctl[2..3] packed big counter, ctl[4..5] exact big pair sum
__global__ void __launch_bounds__(256)
facet_rect_kernel(ColumnGrid g, const double *__restrict__ tri, int64_t nF, int4 *__restrict__ big_rect, int32_t *__restrict__ big_f,
                  int64_t *__restrict__ big_off, int32_t *__restrict__ ctl,
                  double *__restrict__ hit_z, int8_t *__restrict__ hit_s, int32_t *__restrict__ hit_n,
                  int32_t *__restrict__ hit_ev, double oz, int nz)
{
	pdl_release();
	const int lane = threadIdx.x & 31;
	int32_t *overflow_flag = ctl;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < nF; base += stride) {   // warp-uniform trip count
		const int64_t f = base + lane;
		int x0 = 0, y0 = 0, w = 0, h = 0, n_small = 0;
		int64_t n_big = 0;
		if (f < nF) {
			const double *t = tri + 9 * f;
			const double xmin = fmin(t[0], fmin(t[3], t[6])), xmax = fmax(t[0], fmax(t[3], t[6]));
			const double ymin = fmin(t[1], fmin(t[4], t[7])), ymax = fmax(t[1], fmax(t[4], t[7]));
			x0 = first_center_ge(xmin, g.ox, g.spacing, g.nx);
			y0 = first_center_ge(ymin, g.oy, g.spacing, g.ny);
			const int x1 = last_center_le(xmax, g.ox, g.spacing, g.nx), y1 = last_center_le(ymax, g.oy, g.spacing, g.ny);
			w = x1 - x0 + 1;
			h = y1 - y0 + 1;
			const int64_t n = (w > 0 && h > 0) ? (int64_t)w * h : 0;
			if (n <= RECT_INLINE) n_small = (int)n; else n_big = n;
		}
		const unsigned big_mask = __ballot_sync(0xffffffffu, n_big > 0);
		if (big_mask) {                                   // rare on fine meshes; warp-uniform
			long long incl64 = n_big;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, incl64, o); if (lane >= o) incl64 += v; }
			const long long total64 = __shfl_sync(0xffffffffu, incl64, 31);
			unsigned long long old = 0;
			if (lane == 0) {
				old = atomicAdd(reinterpret_cast<unsigned long long *>(ctl + 2), ((unsigned long long)__popc(big_mask) << BIG_PAIR_BITS) + (unsigned long long)total64);
				atomicAdd(reinterpret_cast<unsigned long long *>(ctl + 4), (unsigned long long)total64);
			}
			old = __shfl_sync(0xffffffffu, old, 0);
			if (n_big > 0) {
				const int64_t pos = (int64_t)(old >> BIG_PAIR_BITS) + __popc(big_mask & ((1u << lane) - 1u));
				big_rect[pos] = make_int4(x0, y0, w, h);
				big_f[pos] = (int32_t)f;
				big_off[pos] = (int64_t)(old & ((1ull << BIG_PAIR_BITS) - 1ull)) + (incl64 - n_big);
			}
		}
		// exclusive prefix sum of the small pair counts over the warp
		int incl = n_small;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
		const int excl = incl - n_small;
		const int total = __shfl_sync(0xffffffffu, incl, 31);
		const float inv_w = w > 0 ? __frcp_rn((float)w) : 0.0f;
		for (int p0 = 0; p0 < total; p0 += 32) {
			const int p = p0 + lane;
			int src = 0;                                    // largest lane with excl <= p
#pragma unroll
			for (int step = 16; step > 0; step >>= 1) {
				const int e = __shfl_sync(0xffffffffu, excl, (src + step) & 31);
				if (src + step < 32 && e <= p) src += step;
			}
			const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
			const int sw = __shfl_sync(0xffffffffu, w, src), se = __shfl_sync(0xffffffffu, excl, src);
			const float siw = __shfl_sync(0xffffffffu, inv_w, src);
			if (p < total) {
				const int k = p - se;                         // 0 <= k < RECT_INLINE, 1 <= sw <= RECT_INLINE: (k + 0.5) / sw is at least
				const int row = __float2int_rz(((float)k + 0.5f) * siw);   // 0.5 / 16 away from an integer, far beyond the float error
				const int x = sx0 + (k - row * sw), y = sy0 + row;
				const double cx = (x + 0.5) * g.spacing + g.ox, cy = (y + 0.5) * g.spacing + g.oy; // voxel_center, voxelization.h:85-91
				double z;
				const int sgn = intersect_ray_z(tri + 9 * (base + src), cx, cy, z);
				if (sgn) append_hit(g, x, y, z, sgn, hit_z, hit_s, hit_n, overflow_flag, hit_ev, oz, nz);
			}
		}
	}
}

__global__ void __launch_bounds__(256)
pair_hits_kernel(ColumnGrid g, const double *__restrict__ tri, const int4 *__restrict__ big_rect, const int32_t *__restrict__ big_f,
                 const int64_t *__restrict__ big_off, int32_t *__restrict__ ctl,
                 double *__restrict__ hit_z, int8_t *__restrict__ hit_s, int32_t *__restrict__ hit_n,
                 int32_t *__restrict__ hit_ev, double oz, int nz)
{
	pdl_release();
	pdl_wait();
	// counts are read on the device: no host round trip in front of this launch
	const unsigned long long packed = *reinterpret_cast<const unsigned long long *>(ctl + 2);
	const int64_t n_big = (int64_t)(packed >> BIG_PAIR_BITS), n_pairs = (int64_t)(packed & ((1ull << BIG_PAIR_BITS) - 1ull));
	if (*reinterpret_cast<const unsigned long long *>(ctl + 4) >> BIG_PAIR_BITS) return;      // offsets wrapped: the host reports FPOHM_ERANGE
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_pairs; t += (int64_t)gridDim.x * blockDim.x) {
		int64_t lo = 0, hi = n_big;                    // largest i with big_off[i] <= t
		while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (big_off[mid] <= t) lo = mid; else hi = mid; }
		const int4 r = big_rect[lo];
		const int64_t k = t - big_off[lo];
		const int x = r.x + (int)(k % r.z), y = r.y + (int)(k / r.z);
		const double cx = (x + 0.5) * g.spacing + g.ox, cy = (y + 0.5) * g.spacing + g.oy; // voxel_center, voxelization.h:85-91
		double z;
		const int s = intersect_ray_z(tri + 9 * (int64_t)big_f[lo], cx, cy, z);
		if (s) append_hit(g, x, y, z, s, hit_z, hit_s, hit_n, ctl, hit_ev, oz, nz);
	}
}

// VoxelGrid fill.  voxel(x,y,z) = [ sum of sign_i over the column's hits with k0_i <= z ] < 0 — order independent, so no
// sort.  The z range is cut into chunks of 32 layers.
//   column_summary_kernel  one thread per column: decodes the column's events ONCE and writes two bits per chunk —
//                          "inside at chunk start" and "an event fires inside the chunk" (8 B per column for nz <= 1024:
//                          L2 resident).
//   voxel_fill_kernel      one thread = 4 adjacent columns x one chunk; reads the two summary words of its columns; a
//                          clean chunk needs no event at all (mask = all ones / zero), a dirty one (a few % of the tiles)
//                          re-decodes its events.  Every voxel is written exactly once as part of a 4-byte store: a warp
//                          stores 128 contiguous bytes per layer — the "columns4 zchunk=32" pattern of
//                          scripts/micro/write_patterns.cu, which sustains 5.97 TB/s on B200.
// History (ncu, 1024^3, 2 M facets, profiles/r01_ncu_summary.md): per-column z stack 3.7 ms; per-tile event decode
// 1.08 ms with 1.0 GB of DRAM reads and ~2 400 instructions per thread; events in registers per column 1.66 ms (128 regs).
#define FILL_Z 32
__device__ __forceinline__ uint32_t bit_range(int a, int b) {          // bits [a, b), 0 <= a <= b <= 32
	const int len = b - a;
	return len <= 0 ? 0u : ((len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << a);
}
// summary layout: word w of column col at sum[(2 * w) * ncol + col] (inside bits) and sum[(2 * w + 1) * ncol + col] (dirty bits).
// A dirty chunk's 32 layer bits are worked out here as well, once, and stored at dmask[chunk * ncol + col] (coalesced over
// neighbouring columns): the fill then reads 4 bytes per dirty chunk instead of re-walking the column's 128-byte event list
// (ncu: 0.24 GB of reads and a third of the fill's instructions at 1024^3 before).
// Event driven (round 2; the first version walked all chunks of every column: 1 217 instructions per column, 78 us at 1024^2
// columns).  Events in ascending k; s = sum of the signs so far.  The state ENTERING chunk b is the state after every event with
// k <= 32 b, so a stretch with s < 0 between events at k_prev and k sets the inside bits [ceil(k_prev / 32), ceil(k / 32)); an
// event strictly inside a chunk (k % 32 != 0) makes that chunk dirty and its 32 layer bits are accumulated while the events of
// the chunk go by.  `get(i)` = i-th event in ascending k.
template <class Get>
__device__ __forceinline__ void summarize_column(Get get, int n, int64_t col, int64_t ncol, int nz, int n_words, uint32_t *__restrict__ sum,
                                                 uint32_t *__restrict__ dmask)
{
	const int n_chunks = (nz + FILL_Z - 1) / FILL_Z;
	for (int w = 0; w < n_words; ++w) {
		const int c0 = 32 * w, c1 = min(c0 + 32, n_chunks);           // chunks of this summary word
		uint32_t inside = 0, dirty = 0, mask = 0;
		int s = 0, from = 0 /* ceil(k_prev / 32) */, cur = -1 /* open dirty chunk */, t = 0, prev = 0;
		for (int i = 0; i <= n; ++i) {
			const int32_t ei = i < n ? get(i) : 0;
			const int k = i < n ? (ei >> 2) : nz + FILL_Z * 32;          // sentinel: closes the last stretch
			const int to = min((k + FILL_Z - 1) / FILL_Z, n_chunks);
			if (s < 0) inside |= bit_range(max(from, c0) - c0, min(to, c1) - c0);
			from = to;
			if (i == n) break;
			const int b = k / FILL_Z, r = k % FILL_Z;
			if (r != 0 && b >= c0 && b < c1) {
				if (b != cur) {
					if (cur >= 0) { if (t < 0) mask |= bit_range(prev, FILL_Z); dmask[(int64_t)cur * ncol + col] = mask; }
					cur = b; dirty |= 1u << (b - c0); mask = 0; t = s; prev = 0;
				}
				if (t < 0) mask |= bit_range(prev, r);
				prev = r;
				t = s + (ei & 3) - 1;
			}
			s += (ei & 3) - 1;
		}
		if (cur >= 0) { if (t < 0) mask |= bit_range(prev, FILL_Z); dmask[(int64_t)cur * ncol + col] = mask; }
		sum[(int64_t)(2 * w) * ncol + col] = inside;
		sum[(int64_t)(2 * w + 1) * ncol + col] = dirty;
	}
}
template <int CAP>
__global__ void __launch_bounds__(256)
column_summary_kernel(int64_t ncol, int nz, int n_words, const int32_t *__restrict__ hit_ev, const int32_t *__restrict__ hit_n,
                      uint32_t *__restrict__ sum, uint32_t *__restrict__ dmask, const int32_t *__restrict__ ctl, volatile int32_t *host_post, int32_t seq)
{
	pdl_release();
	pdl_wait();
	// The control words of the hit stage are final now: post them to the host (pinned, mapped) so that the caller learns whether
	// a list overflowed while summary and fill are still running — no event, no copy, nothing between the kernels of the pass.
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		for (int j = 0; j < HIT_CTL_WORDS; ++j) host_post[j] = ctl[j];
		__threadfence_system();
		host_post[HIT_CTL_WORDS] = seq;
	}
	for (int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; col < ncol; col += (int64_t)gridDim.x * blockDim.x) {
		const int n = min(hit_n[col], CAP);
		const int32_t *ev = hit_ev + col;                    // event i of this column at ev[i * ncol]
		if (n == 0) {
			for (int w = 0; w < 2 * n_words; ++w) sum[(int64_t)w * ncol + col] = 0u;
			continue;
		}
		// (sorting networks in registers for n <= 4 / n <= 8 were measured: no gain, the kernel is bound by its divergent event loop)
		int32_t e[CAP];
		for (int i = 0; i < n; ++i) {                       // insertion sort by k0 (order among equal k0 is irrelevant)
			const int32_t v = ev[(int64_t)i * ncol];
			int j = i - 1;
			while (j >= 0 && (e[j] >> 2) > (v >> 2)) { e[j + 1] = e[j]; --j; }
			e[j + 1] = v;
		}
		summarize_column([&](int i) { return e[i]; }, n, col, ncol, nz, n_words, sum, dmask);
	}
}

#define FILL_CH_DEFAULT 1      // chunks of 32 layers per thread (1 / 2 / 4 with the grid uncapped: 0.357 / 0.364 / 0.383 ms per pass)
template <int FILL_CH>
__global__ void __launch_bounds__(256)
voxel_fill_kernel(int nx, int ny, int nz, const uint32_t *__restrict__ dmask, const uint32_t *__restrict__ sum, uint8_t *__restrict__ out,
                  int zc_begin, int zc_end)
{
	pdl_wait();
	// layers [zc_begin * 32, min(zc_end * 32, nz)) are written, the first of them at `out` (z-slab sharding)
	const int gx = (nx + 3) / 4, gzc = zc_end, gz = (zc_end - zc_begin + FILL_CH - 1) / FILL_CH;
	const int64_t nthreads = (int64_t)gx * ny * gz;
	const int64_t layer = (int64_t)nx * ny;
	const bool aligned = (nx & 3) == 0;               // rows are 4-byte aligned, no ragged tail
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nthreads; t += (int64_t)gridDim.x * blockDim.x) {
		const int x0 = (int)(t % gx) * 4, y = (int)((t / gx) % ny), zg = (int)(t / ((int64_t)gx * ny));
		const int64_t col0 = (int64_t)y * nx + x0;
		uint32_t si[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0};
		int cur_word = -1;
#pragma unroll
		for (int ch = 0; ch < FILL_CH; ++ch) {
			const int zc = zc_begin + zg * FILL_CH + ch;
			if (zc >= gzc) break;
			if ((zc >> 5) != cur_word) {
				cur_word = zc >> 5;
				const uint32_t *pi = sum + (int64_t)(2 * cur_word) * layer + col0, *pd = pi + layer;
				if (aligned) {
					const uint4 a = __ldg(reinterpret_cast<const uint4 *>(pi)), b = __ldg(reinterpret_cast<const uint4 *>(pd));
					si[0] = a.x; si[1] = a.y; si[2] = a.z; si[3] = a.w; sd[0] = b.x; sd[1] = b.y; sd[2] = b.z; sd[3] = b.w;
				} else {
					for (int c = 0; c < 4; ++c) if (x0 + c < nx) { si[c] = pi[c]; sd[c] = pd[c]; }
				}
			}
			const int z0 = zc * FILL_Z, z1 = min(z0 + FILL_Z, nz);
			const int bit = zc & 31;
			uint32_t m[4];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				if ((sd[c] >> bit) & 1u) m[c] = __ldg(dmask + (int64_t)zc * layer + col0 + c);
				else m[c] = ((si[c] >> bit) & 1u) ? 0xffffffffu : 0u;
			}
			uint8_t *o = out + (int64_t)(z0 - zc_begin * FILL_Z) * layer + col0;     // index_from_index3, voxelization.cpp:26-28
			const uint32_t any = m[0] | m[1] | m[2] | m[3], all = m[0] & m[1] & m[2] & m[3];
			// the whole warp takes ONE store path: lanes split between the two loops would turn every 128-byte row into two
			// partial-line stores (ncu: 3.37 of 4 sectors per store request, 22 of 32 lanes active)
			const bool tile_uniform = any == 0u || all == 0xffffffffu;
			const bool warp_uniform = __ballot_sync(__activemask(), !tile_uniform) == 0u;
			if (aligned && warp_uniform) {                                 // uniform tiles: store only
				const uint32_t w = any ? 0x01010101u : 0u;
				for (int z = z0; z < z1; ++z, o += layer) __stcs(reinterpret_cast<uint32_t *>(o), w);
			} else if (aligned) {
				uint32_t m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
				for (int z = z0; z < z1; ++z, o += layer) {
					__stcs(reinterpret_cast<uint32_t *>(o), (m0 & 1u) | ((m1 & 1u) << 8) | ((m2 & 1u) << 16) | ((m3 & 1u) << 24));
					m0 >>= 1; m1 >>= 1; m2 >>= 1; m3 >>= 1;
				}
			} else {
				for (int z = z0; z < z1; ++z, o += layer)
					for (int c = 0; c < 4; ++c) if (x0 + c < nx) o[c] = (m[c] >> (z - z0)) & 1u;
			}
		}
	}
}

// DexelGrid: sorted hits reduced to entry/exit events, voxelization.h:312-321
template <int CAP>
__global__ void dexel_reduce_kernel(int64_t ncol, double *__restrict__ hit_z, int8_t *__restrict__ hit_s, int32_t *__restrict__ hit_n,
                                    int64_t *__restrict__ count)
{
	for (int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; col < ncol; col += (int64_t)gridDim.x * blockDim.x) {
		const int n = min(hit_n[col], CAP);
		double hz[CAP]; int8_t hs[CAP];
		for (int i = 0; i < n; ++i) { hz[i] = hit_z[col * CAP + i]; hs[i] = hit_s[col * CAP + i]; }
		sort_hits(hz, hs, n);
		int m = 0;
		for (int i = 0, s = 0; i < n; ++i) {
			const int ds = hs[i];
			s += ds;
			if ((s == -1 && ds < 0) || (s == 0 && ds > 0)) hit_z[col * CAP + m++] = hz[i];
		}
		hit_n[col] = m;
		count[col] = m;
	}
}
__global__ void dexel_emit_kernel(int64_t ncol, int cap, const double *__restrict__ hit_z, const int32_t *__restrict__ hit_n,
                                  const int64_t *__restrict__ off, double *__restrict__ values)
{
	for (int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; col < ncol; col += (int64_t)gridDim.x * blockDim.x) {
		const int n = hit_n[col];
		for (int i = 0; i < n; ++i) values[off[col] + i] = hit_z[col * cap + i];
	}
}

// compute_sign(OctreeGrid), voxelization.cpp:114-158: one thread per cell (ALL cells, leaves and internal)
template <int CAP>
__global__ void __launch_bounds__(128)
cell_sign_kernel(const uint8_t *__restrict__ lvl, const uint64_t *__restrict__ code, int64_t n_cells, int depth,
                 double ox, double oy, double oz, double spacing, const double *__restrict__ box, int64_t P,
                 const int32_t *__restrict__ order, const double *__restrict__ tri, float *__restrict__ inside,
                 int32_t *__restrict__ overflow_flag)
{
	for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < n_cells; id += (int64_t)gridDim.x * blockDim.x) {
		const int sh = depth - lvl[id];
		const uint64_t c = code[id];
		const int x = (int)(compact1by2(c) << sh), y = (int)(compact1by2(c >> 1) << sh), z = (int)(compact1by2(c >> 2) << sh);
		const int extent = 1 << sh;
		const double bx0 = ox + spacing * x, by0 = oy + spacing * y;
		const double bx1 = bx0 + spacing * extent, by1 = by0 + spacing * extent;
		const double cx = bx0 + 0.5 * spacing * extent, cy = by0 + 0.5 * spacing * extent;
		const double cz = oz + spacing * z + 0.5 * spacing * extent;
		double hz[CAP]; int8_t hs[CAP];
		bool ov = false;
		int n = gather_hits(box, P, order, tri, bx0, bx1, by0, by1, cx, cy, hz, hs, CAP, ov);
		if (ov) { atomicMax(overflow_flag, n); n = CAP; }
		sort_hits(hz, hs, n);
		int num_before = 0;
		for (int i = 0, s = 0; i < n; ++i) {
			const int ds = hs[i];
			s += ds;
			if ((s == -1 && ds < 0) || (s == 0 && ds > 0)) { if (hz[i] < cz) ++num_before; }
		}
		inside[id] = (num_before % 2 == 1) ? 1.0f : 0.0f;
	}
}

// launch with the programmatic-stream-serialization attribute (see pdl_release / pdl_wait)
template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
	static const bool off = getenv("FPOHM_NO_PDL") != nullptr;      // A/B switch
	if (off) { kernel<<<grid, block, 0, s>>>(args...); return; }
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	FPOHM_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

struct HitScratch {
	DevBuf<double> z; DevBuf<int8_t> s; DevBuf<int32_t> ctl /* HIT_CTL_WORDS control words, then one count per column */, ev;
	int32_t *n() const { return ctl.p + HIT_CTL_WORDS; }
};
struct HitPtrs {        // the same buffers as raw pointers (the VoxelGrid pass carves them out of one allocation)
	double *z; int8_t *s; int32_t *ctl /* zeroed: HIT_CTL_WORDS control words, then one count per column */, *ev;
	int4 *big_rect; int32_t *big_f; int64_t *big_off;      // nF entries each; touched only where a facet covers more than RECT_INLINE columns
};

void launch_column_hits(fpohm_ctx *ctx, const fpohm_mesh *mesh, const ColumnGrid &g, const HitPtrs &h, cudaStream_t s, double oz, int nz) {
	const int64_t nF = mesh->nF;
	static const int rect_ctas = getenv("FPOHM_RECT_CTAS") ? atoi(getenv("FPOHM_RECT_CTAS")) : 8;      // CTAs per SM in the grid (debug sweep)
	facet_rect_kernel<<<grid_for(ctx, nF, 256, rect_ctas), 256, 0, s>>>(g, mesh->tri.p, nF, h.big_rect, h.big_f, h.big_off, h.ctl, h.z, h.s, h.ctl + HIT_CTL_WORDS, h.ev, oz, nz);
	FPOHM_LAUNCH_CHECK(ctx);
	// on fine meshes (no facet over RECT_INLINE columns) the launch finds nothing to do
	launch_pdl(pair_hits_kernel, ctx->sm_count * 8, 256, s, g, mesh->tri.p, h.big_rect, h.big_f, h.big_off, h.ctl, h.z, h.s, h.ctl + HIT_CTL_WORDS, h.ev, oz, nz);
	FPOHM_LAUNCH_CHECK(ctx);
}

// (z, sign) hit lists of all columns (DexelGrid)
void run_column_hits(fpohm_ctx *ctx, fpohm_mesh *mesh, const ColumnGrid &g, HitScratch &h, cudaStream_t s) {
	const int64_t ncol = (int64_t)g.nx * g.ny, nF = mesh->nF;
	FPOHM_REQUIRE(nF < (1ll << (64 - BIG_PAIR_BITS)), FPOHM_ERANGE, "ray parity: %lld facets (the limit is 2^28)", (long long)nF);
	h.z.alloc(ncol * g.cap, s); h.s.alloc(ncol * g.cap, s);
	h.ctl.alloc(HIT_CTL_WORDS + ncol, s);
	h.ctl.zero();                                      // counts and control words in one memset
	DevBuf<int4> big_rect(nF, s);
	DevBuf<int32_t> big_f(nF, s);
	DevBuf<int64_t> big_off(nF, s);
	launch_column_hits(ctx, mesh, g, HitPtrs{h.z.p, h.s.p, h.ctl.p, nullptr, big_rect.p, big_f.p, big_off.p}, s, 0, 0);
}

// control words of a pass -> 0 if every list fitted, else the capacity the pass has to be repeated with (FPOHM_ERANGE beyond HIT_CAP_MAX)
int retry_cap_from(const int32_t *ctl, const char *who) {
	unsigned long long big_pairs; memcpy(&big_pairs, ctl + 4, 8);
	FPOHM_REQUIRE((big_pairs >> BIG_PAIR_BITS) == 0, FPOHM_ERANGE, "%s: %llu (facet, column) pairs (the limit is 2^%d)", who, big_pairs, BIG_PAIR_BITS);
	const int32_t ov = ctl[0];
	if (ov == 0) return 0;
	const int cap = next_hit_cap(ov);
	FPOHM_REQUIRE(cap > 0, FPOHM_ERANGE, "%s: %d ray/facet hits in one column (the limit is %d)", who, ov, HIT_CAP_MAX);
	return cap;
}

// 0 if every list fitted, else the capacity the pass has to be repeated with (FPOHM_ERANGE beyond HIT_CAP_MAX)
int overflow_retry_cap(DevBuf<int32_t> &ctl_dev /* HIT_CTL_WORDS control words first */, cudaStream_t s, const char *who) {
	int32_t ctl[HIT_CTL_WORDS] = {0};
	ctl_dev.download(ctl, HIT_CTL_WORDS);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	return retry_cap_from(ctl, who);
}

void check_dims(const int32_t *dims, int nd, const char *who) {
	int64_t prod = 1;
	for (int d = 0; d < nd; ++d) {
		FPOHM_REQUIRE(dims[d] > 0, FPOHM_EINVAL, "%s: dims[%d]=%d", who, d, dims[d]);
		prod *= dims[d];
	}
	FPOHM_REQUIRE(prod < (1ll << 31), FPOHM_ERANGE, "%s: %lld voxels overflow the reference's int indexing (voxelization.h:55)", who, (long long)prod);
}

} // namespace

extern "C" {

int fpohm_voxel_grid_setup(const double origin[3], const double extent[3], double spacing, int32_t padding,
                           int32_t dims[3], double origin_out[3])
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(origin && extent && dims && origin_out && spacing > 0 && padding >= 0, FPOHM_EINVAL, "fpohm_voxel_grid_setup: bad argument");
	for (int d = 0; d < 3; ++d) {
		origin_out[d] = origin[d] - padding * spacing * 1.0;             // m_origin -= padding * spacing * vec3(1,1,1)
		dims[d] = (int)std::ceil(extent[d] / spacing) + 2 * padding;     // voxelization.h:76-78
	}
	FPOHM_API_END
}

int fpohm_voxel_sign_slab_dev(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                              const int32_t dims[3], int32_t z_begin, int32_t z_end, uint8_t *out_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && grid_origin && dims && out_dev && spacing > 0, FPOHM_EINVAL, "fpohm_voxel_sign_slab_dev: bad argument");
	check_dims(dims, 3, "fpohm_voxel_sign_slab_dev");
	FPOHM_REQUIRE(z_begin >= 0 && z_begin < z_end && z_end <= dims[2] && z_begin % FILL_Z == 0 && (z_end % FILL_Z == 0 || z_end == dims[2]), FPOHM_EINVAL,
	              "fpohm_voxel_sign_slab_dev: slab [%d,%d) must be non-empty, inside [0,%d) and aligned to %d layers", z_begin, z_end, dims[2], FILL_Z);
	DeviceGuard g(ctx->device);
	cudaStream_t s = (cudaStream_t)stream;
	const int64_t nF = mesh->nF;
	FPOHM_REQUIRE(nF < (1ll << (64 - BIG_PAIR_BITS)), FPOHM_ERANGE, "fpohm_voxel_sign: %lld facets (the limit is 2^28)", (long long)nF);
	if (!ctx->pinned_words) {
		FPOHM_CUDA(cudaHostAlloc((void **)&ctx->pinned_words, 16 * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
		memset(ctx->pinned_words, 0, 16 * sizeof(int32_t));
	}
	for (int cap = HIT_CAP;;) {      // optimistic pass; repeated with more room only if a column overflowed
	const ColumnGrid cg{grid_origin[0], grid_origin[1], spacing, dims[0], dims[1], cap};
	const int gz = (dims[2] + FILL_Z - 1) / FILL_Z, n_words = (gz + 31) / 32;
	const int64_t ncol = (int64_t)dims[0] * dims[1];
	// ONE allocation for the pass (seven stream-ordered allocations cost ~20 us of host time in front of the first kernel)
	auto up = [](int64_t b) { return (b + 255) & ~(int64_t)255; };
	const int64_t b_ev = up(4 * ncol * cap), b_ctl = up(4 * (HIT_CTL_WORDS + ncol)), b_rect = up(16 * nF), b_f = up(4 * nF), b_off = up(8 * nF);
	const int64_t b_sum = up(4 * 2 * (int64_t)n_words * ncol), b_dm = up(4 * (int64_t)gz * ncol);      // dmask is only written / read where a chunk is dirty
	DevBuf<uint8_t> slab(b_ev + b_ctl + b_rect + b_f + b_off + b_sum + b_dm, s);
	uint8_t *q = slab.p;
	HitPtrs h{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	h.ev = (int32_t *)q; q += b_ev; h.ctl = (int32_t *)q; q += b_ctl; h.big_rect = (int4 *)q; q += b_rect; h.big_f = (int32_t *)q; q += b_f; h.big_off = (int64_t *)q; q += b_off;
	uint32_t *summary = (uint32_t *)q; q += b_sum;
	uint32_t *dmask = (uint32_t *)q;
	FPOHM_CUDA(cudaMemsetAsync(h.ctl, 0, 4 * (size_t)(HIT_CTL_WORDS + ncol), s));      // counts and control words in one memset
	// hits and per-column summaries are global (every slab needs the parity of everything below it); only the fill is sliced
	launch_column_hits(ctx, mesh, cg, h, s, grid_origin[2], dims[2]);
	int32_t *h_n = h.ctl + HIT_CTL_WORDS;
	static const int sum_ctas = getenv("FPOHM_SUM_CTAS") ? atoi(getenv("FPOHM_SUM_CTAS")) : 32;         // CTAs per SM in the grid: 4 / 8 / 16 / 32 / 64 -> 0.383 / 0.375 / 0.368 / 0.363 / 0.366 ms per pass
	const int32_t seq = ++ctx->post_seq;
	volatile int32_t *post = ctx->pinned_words;
	switch (cap) {
	case 32: launch_pdl(column_summary_kernel<32>, grid_for(ctx, ncol, 256, sum_ctas), 256, s, ncol, dims[2], n_words, h.ev, h_n, summary, dmask, h.ctl, post, seq); break;
	case 128: launch_pdl(column_summary_kernel<128>, grid_for(ctx, ncol, 256, sum_ctas), 256, s, ncol, dims[2], n_words, h.ev, h_n, summary, dmask, h.ctl, post, seq); break;
	case 512: launch_pdl(column_summary_kernel<512>, grid_for(ctx, ncol, 256, 4), 256, s, ncol, dims[2], n_words, h.ev, h_n, summary, dmask, h.ctl, post, seq); break;
	default: launch_pdl(column_summary_kernel<2048>, grid_for(ctx, ncol, 256, 2), 256, s, ncol, dims[2], n_words, h.ev, h_n, summary, dmask, h.ctl, post, seq); break;
	}
	FPOHM_LAUNCH_CHECK(ctx);
	const int zc0 = z_begin / FILL_Z, zc1 = (z_end + FILL_Z - 1) / FILL_Z;
	static const int fill_ch = getenv("FPOHM_FILL_CH") ? atoi(getenv("FPOHM_FILL_CH")) : FILL_CH_DEFAULT;
	static const int fill_ctas = getenv("FPOHM_FILL_CTAS") ? atoi(getenv("FPOHM_FILL_CTAS")) : 256;
	// (a 16-columns-per-thread variant with 16-byte stores was measured as well: the bare pattern is faster, 6.7 vs 5.9 TB/s, the fill
	// is not — 182 us either way at 1024^3, profiles/r02_ncu_summary.md)
	const int64_t nthreads = (int64_t)((dims[0] + 3) / 4) * dims[1] * ((zc1 - zc0 + fill_ch - 1) / fill_ch);
	const int fgrid = grid_for(ctx, nthreads, 256, fill_ctas);
	switch (fill_ch) {
	case 1: launch_pdl(voxel_fill_kernel<1>, fgrid, 256, s, dims[0], dims[1], z_end, dmask, summary, out_dev, zc0, zc1); break;
	case 4: launch_pdl(voxel_fill_kernel<4>, fgrid, 256, s, dims[0], dims[1], z_end, dmask, summary, out_dev, zc0, zc1); break;
	case 8: launch_pdl(voxel_fill_kernel<8>, fgrid, 256, s, dims[0], dims[1], z_end, dmask, summary, out_dev, zc0, zc1); break;
	default: launch_pdl(voxel_fill_kernel<2>, fgrid, 256, s, dims[0], dims[1], z_end, dmask, summary, out_dev, zc0, zc1); break;
	}
	FPOHM_LAUNCH_CHECK(ctx);
	// wait for the post of THIS pass (the call returns stream-ordered, like every _dev entry point, without waiting for the fill)
	for (int64_t spin = 1; post[HIT_CTL_WORDS] != seq; ++spin) {
		if ((spin & 4095) == 0) { const cudaError_t q = cudaStreamQuery(s); if (q != cudaSuccess && q != cudaErrorNotReady) FPOHM_CUDA(q); if (q == cudaSuccess && post[HIT_CTL_WORDS] != seq) FPOHM_REQUIRE(false, FPOHM_ESTATE, "fpohm_voxel_sign: the pass finished without posting its control words"); }
	}
	int32_t h_ctl[HIT_CTL_WORDS];
	for (int j = 0; j < HIT_CTL_WORDS; ++j) h_ctl[j] = post[j];
	const int retry = retry_cap_from(h_ctl, "fpohm_voxel_sign");
	if (!retry) break;
	cap = retry;
	}
	FPOHM_API_END
}

int fpohm_voxel_sign_dev(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                         const int32_t dims[3], uint8_t *out_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(dims, FPOHM_EINVAL, "fpohm_voxel_sign_dev: bad argument");
	return fpohm_voxel_sign_slab_dev(ctx, mesh, grid_origin, spacing, dims, 0, dims[2], out_dev, stream);
	FPOHM_API_END
}

int fpohm_voxel_sign(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                     const int32_t dims[3], uint8_t *out)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && grid_origin && dims && out && spacing > 0, FPOHM_EINVAL, "fpohm_voxel_sign: bad argument");
	check_dims(dims, 3, "fpohm_voxel_sign");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int64_t n = (int64_t)dims[0] * dims[1] * dims[2];
	DevBuf<uint8_t> d(n, s);
	KernelTimer t(ctx, s);
	const int rc = fpohm_voxel_sign_dev(ctx, mesh, grid_origin, spacing, dims, d.p, s);
	t.stop();
	if (rc != FPOHM_OK) return rc;
	d.download(out, n);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_dexel_sign(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing,
                     const int32_t dims2[2], int64_t *offsets, double *values, int64_t *total)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && grid_origin && dims2 && total && spacing > 0, FPOHM_EINVAL, "fpohm_dexel_sign: bad argument");
	check_dims(dims2, 2, "fpohm_dexel_sign");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int64_t ncol = (int64_t)dims2[0] * dims2[1];
	KernelTimer t(ctx, s);
	DevBuf<int64_t> cnt(ncol + 1, s), off(ncol + 1, s);
	for (int cap = HIT_CAP;;) {
		HitScratch h;
		const ColumnGrid cg{grid_origin[0], grid_origin[1], spacing, dims2[0], dims2[1], cap};
		run_column_hits(ctx, const_cast<fpohm_mesh *>(mesh), cg, h, s);
		cnt.zero();
		switch (cap) {
		case 32: dexel_reduce_kernel<32><<<grid_for(ctx, ncol, 128), 128, 0, s>>>(ncol, h.z.p, h.s.p, h.n(), cnt.p); break;
		case 128: dexel_reduce_kernel<128><<<grid_for(ctx, ncol, 128), 128, 0, s>>>(ncol, h.z.p, h.s.p, h.n(), cnt.p); break;
		case 512: dexel_reduce_kernel<512><<<grid_for(ctx, ncol, 128, 4), 128, 0, s>>>(ncol, h.z.p, h.s.p, h.n(), cnt.p); break;
		default: dexel_reduce_kernel<2048><<<grid_for(ctx, ncol, 128, 2), 128, 0, s>>>(ncol, h.z.p, h.s.p, h.n(), cnt.p); break;
		}
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p, ncol + 1, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, off.p, ncol + 1, s));
		ctx->launches += 2;
		int64_t tot = 0;
		FPOHM_CUDA(cudaMemcpyAsync(&tot, off.p + ncol, 8, cudaMemcpyDeviceToHost, s));
		const int retry = overflow_retry_cap(h.ctl, s, "fpohm_dexel_sign");
		if (retry) { cap = retry; continue; }
		*total = tot;
		if (values) {
			DevBuf<double> dv(tot, s);
			dexel_emit_kernel<<<grid_for(ctx, ncol, 128), 128, 0, s>>>(ncol, cap, h.z.p, h.n(), off.p, dv.p);
			FPOHM_LAUNCH_CHECK(ctx);
			dv.download(values, tot);
			FPOHM_CUDA(cudaStreamSynchronize(s));
		}
		break;
	}
	t.stop();
	if (offsets) off.download(offsets, ncol + 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_octree_cell_sign(const fpohm_octree *o, const fpohm_mesh *mesh, const double origin[3], double spacing, float *inside) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(o && mesh && origin && inside && spacing > 0, FPOHM_EINVAL, "fpohm_octree_cell_sign: bad argument");
	fpohm_ctx *ctx = o->ctx;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	fpohm_mesh *m = const_cast<fpohm_mesh *>(mesh);
	mesh_ensure_pred(ctx, m, s);
	DevBuf<float> d(o->n_cells, s);
	DevBuf<int32_t> ov(HIT_CTL_WORDS, s);
	KernelTimer t(ctx, s);
	for (int cap = HIT_CAP;;) {
		ov.zero();
#define FPOHM_CELL_SIGN(C, CTAS) cell_sign_kernel<C><<<grid_for(ctx, o->n_cells, 128, CTAS), 128, 0, s>>>(o->cell_level.p, o->cell_code.p, o->n_cells, o->depth, \
			origin[0], origin[1], origin[2], spacing, m->pred_box.p, m->pred_nodes / 2, m->pred_order.p, m->tri.p, d.p, ov.p)
		switch (cap) {
		case 32: FPOHM_CELL_SIGN(32, 16); break;
		case 128: FPOHM_CELL_SIGN(128, 8); break;
		case 512: FPOHM_CELL_SIGN(512, 4); break;
		default: FPOHM_CELL_SIGN(2048, 1); break;
		}
#undef FPOHM_CELL_SIGN
		FPOHM_LAUNCH_CHECK(ctx);
		const int retry = overflow_retry_cap(ov, s, "fpohm_octree_cell_sign");
		if (!retry) break;
		cap = retry;
	}
	t.stop();
	d.download(inside, o->n_cells);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

} // extern "C"

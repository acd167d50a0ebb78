// Remaining query-layer entry points: feature-curve projection, metro Hausdorff, Hausdorff outliers, voxel lattice,
// dense subdivision-predicate occupancy.
#include "mesh.h"
#include <math_constants.h>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>

using namespace fpohm;

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 sub(const V3 &a, const V3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot3(const V3 &a, const V3 &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); } // fixed-size Eigen redux
__device__ __forceinline__ double sqnorm(const V3 &a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 ld3(const double *p) { return {p[0], p[1], p[2]}; }

// LINE branch of dirty_graph_projection (ghm.cpp:3967-3994) + point_line_projection (gf.cpp:3454-3465)
__global__ void polyline_kernel(const double *__restrict__ Vc, const int64_t *__restrict__ off, const int32_t *__restrict__ cvs,
                                const uint8_t *__restrict__ circle, const double *__restrict__ P, const int32_t *__restrict__ cid,
                                int64_t np, double *__restrict__ origin_L, double *__restrict__ axis_L)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
		const V3 v = ld3(P + 3 * i);
		const int c = cid[i];
		const int32_t *curve = cvs + off[c];
		const uint32_t size = (uint32_t)(off[c + 1] - off[c]);
		uint32_t curve_len = size;
		if (!circle[c]) curve_len--;
		V3 pv = {0, 0, 0};                 // the reference's pv survives iterations (only matters for NaN t)
		V3 best_pv = {0, 0, 0}, best_t = {1, 0, 0};
		double best_d = CUDART_INF;
		bool have = false;
		for (uint32_t j = 0; j < curve_len; ++j) {
			const V3 a = ld3(Vc + 3 * (int64_t)curve[j]), b = ld3(Vc + 3 * (int64_t)curve[(j + 1) % size]);
			const V3 vv1 = sub(v, a), v21 = sub(b, a);
			const double nv21_2 = sqnorm(v21);
			double t;
			if (nv21_2 >= 1.e-7) t = dot3(vv1, v21) / nv21_2; else t = 0;
			if (t >= 0.0 && t <= 1.0) pv = {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z)};
			else if (t < 0.0) pv = a;
			else if (t > 1.0) pv = b;
			const double d = sqrt(sqnorm(sub(v, pv)));
			// std::sort of pair<dist, idx>: smallest dist, ties -> smallest idx; a NaN dist never wins unless first
			if (!have || d < best_d) {
				have = true; best_d = d; best_pv = pv;
				const double n = sqrt(sqnorm(v21));
				best_t = {v21.x / n, v21.y / n, v21.z / n}; // (b - a).normalized(): true division in Eigen 3.2
			}
		}
		origin_L[3 * i] = best_pv.x; origin_L[3 * i + 1] = best_pv.y; origin_L[3 * i + 2] = best_pv.z;
		axis_L[3 * i] = best_t.x; axis_L[3 * i + 1] = best_t.y; axis_L[3 * i + 2] = best_t.z;
	}
}

// ---- metro -------------------------------------------------------------------------------------------------------
__global__ void mark_referenced_kernel(const int32_t *__restrict__ F, int64_t n3, uint8_t *__restrict__ flag) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n3; t += (int64_t)gridDim.x * blockDim.x) flag[F[t]] = 1;
}
__global__ void gather_points_kernel(const double *__restrict__ V, const int32_t *__restrict__ idx, int64_t n, double *__restrict__ P) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 3 * n; t += (int64_t)gridDim.x * blockDim.x) P[t] = V[3 * (int64_t)idx[t / 3] + t % 3];
}
__global__ void iota_kernel(int32_t *p, int64_t n) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = (int32_t)t;
}
// similar-triangle face samples: the lattice of Sampling::SimilarTriangles (extern/vcg/sampling.h:496-511) for the per-face edge
// counts the host recurrence produced
__global__ void face_samples_kernel(const double *__restrict__ tri, int64_t nF, const int64_t *__restrict__ off, const int32_t *__restrict__ per_edge,
                                    double *__restrict__ P)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const int m = per_edge[f];
		if (m < 4) continue;
		const double *t = tri + 9 * f;
		const V3 v0 = ld3(t), v1 = ld3(t + 3), v2 = ld3(t + 6);
		const double inv = (double)(m - 1);
		const V3 V1 = {(v1.x - v0.x) / inv, (v1.y - v0.y) / inv, (v1.z - v0.z) / inv};
		const V3 V2 = {(v2.x - v0.x) / inv, (v2.y - v0.y) / inv, (v2.z - v0.z) / inv};
		int64_t o = off[f];
		for (int i = 1; i < m - 1; ++i)
			for (int j = 1; j < m - 1 - i; ++j) {
				P[3 * o] = v0.x + (V1.x * (double)i + V2.x * (double)j);
				P[3 * o + 1] = v0.y + (V1.y * (double)i + V2.y * (double)j);
				P[3 * o + 2] = v0.z + (V1.z * (double)i + V2.z * (double)j);
				++o;
			}
	}
}

struct DistPartial { double mx, sum, sumsq; long long n; };
// AddSample statistics (extern/vcg/sampling.h:244-251) from squared distances
__global__ void __launch_bounds__(256)
dist_stats_kernel(const double *__restrict__ sq, int64_t n, double upper, DistPartial *__restrict__ partials) {
	__shared__ DistPartial sm[8];
	DistPartial a{-CUDART_INF, 0.0, 0.0, 0};
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const double d = sqrt(sq[i]);
		if (d >= upper) continue;          // dist == dist_upper_bound: nothing found inside the search radius
		a.mx = fmax(a.mx, d); a.sum += d; a.sumsq += d * d; a.n += 1;
	}
	for (int o = 16; o > 0; o >>= 1) {
		a.mx = fmax(a.mx, __shfl_down_sync(0xffffffffu, a.mx, o));
		a.sum += __shfl_down_sync(0xffffffffu, a.sum, o);
		a.sumsq += __shfl_down_sync(0xffffffffu, a.sumsq, o);
		a.n += __shfl_down_sync(0xffffffffu, a.n, o);
	}
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a.mx = fmax(a.mx, sm[w].mx); a.sum += sm[w].sum; a.sumsq += sm[w].sumsq; a.n += sm[w].n; }
		partials[blockIdx.x] = a;
	}
}

// hausdorff_dis outliers (gf.cpp:3606-3626), one decay round
__global__ void outlier_flag_kernel(const double *__restrict__ sqAB, const int32_t *__restrict__ I0, int64_t nA, const int32_t *__restrict__ FB,
                                    const double *__restrict__ sqBA, int64_t nB, double thr, uint8_t *__restrict__ flag)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nA + nB; i += (int64_t)gridDim.x * blockDim.x) {
		if (i < nA) {
			if (sqAB[i] > thr) { const int64_t f = I0[i]; flag[FB[3 * f]] = 1; flag[FB[3 * f + 1]] = 1; flag[FB[3 * f + 2]] = 1; }
		} else {
			const int64_t j = i - nA;
			if (sqBA[j] > thr) flag[j] = 1;
		}
	}
}

// voxel_meshing lattice, ghm.cpp:226-291
__global__ void lattice_vertices_kernel(double mx, double my, double mz, float gx, float gy, float gz, int d0, int d1, int d2,
                                        double *__restrict__ V)
{
	const int64_t n = (int64_t)d0 * d1 * d2;
	for (int64_t vn = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; vn < n; vn += (int64_t)gridDim.x * blockDim.x) {
		const int i = (int)(vn / ((int64_t)d1 * d2)), j = (int)((vn % ((int64_t)d1 * d2)) / d2), k = (int)(vn % d2);
		V[3 * vn] = mx + (double)__fmul_rn(gx, (float)i);       // `grid_length[0] * i` is a float product (Vector3f)
		V[3 * vn + 1] = my + (double)__fmul_rn(gy, (float)j);
		V[3 * vn + 2] = mz + (double)__fmul_rn(gz, (float)k);
	}
}
__global__ void lattice_hexes_kernel(int d0, int d1, int d2, uint32_t *__restrict__ hex) {
	const int64_t n = (int64_t)(d0 - 1) * (d1 - 1) * (d2 - 1);
	for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < n; h += (int64_t)gridDim.x * blockDim.x) {
		const int a = (int)(h / ((int64_t)(d1 - 1) * (d2 - 1))), b = (int)((h % ((int64_t)(d1 - 1) * (d2 - 1))) / (d2 - 1)), c = (int)(h % (d2 - 1));
		const int da[8] = {0, 1, 1, 0, 0, 1, 1, 0}, db[8] = {0, 0, 1, 1, 0, 0, 1, 1}, dc[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
		for (int j = 0; j < 8; ++j) hex[8 * h + j] = (uint32_t)(((int64_t)(a + da[j]) * d1 + (b + db[j])) * d2 + (c + dc[j]));
	}
}

// Dense subdivision predicate, FACET-parallel.  Cell (x,y,z) has the box [o + sp*x, (o + sp*x) + sp*1] per axis
// (voxelization.cpp:370-375 with extent 1) and is occupied iff some facet box overlaps it in closed intervals
// (geo/basic/geometry.h:612-622).  Per axis the overlap condition is monotone in the cell index, so every facet owns an
// index BOX [x0,x1] x [y0,y1] x [z0,z1] whose limits are found with the reference's own expressions; the grid is zeroed
// once and each facet stamps its box.  Work ~ number of (facet, cell) overlaps instead of one tree descent per voxel.
// As in voxel.cu: t = (value - origin) / spacing is the same quantity in real numbers; unless t lies within 1e-3 of an integer (or is
// huge) the rounding of the reference's expression cannot change the outcome and floor(t) decides; only the near-integer cases run
// the exact comparison loops.
__device__ __forceinline__ bool occ_index_safe(double t, double ft) { return t - ft > 1e-3 && t - ft < 1.0 - 1e-3 && fabs(t) < 1e9; }
__device__ __forceinline__ int first_cell_max_ge(double lo, double o, double sp, int n) {
	// smallest i in [0, n] with (o + sp*i) + sp*1 >= lo      (cell.max < tri.min fails)
	const double t = (lo - o) * __drcp_rn(sp), ft = floor(t);
	if (occ_index_safe(t, ft)) return max(0, min(n, (int)ft));          // smallest i with i + 1 >= t
	int i = (int)fmax(-2.0, fmin(ft - 1.0, 2147483000.0));              // guess only
	if (i < 0) i = 0;
	if (i > n) i = n;
	while (i > 0 && (o + sp * (i - 1)) + sp * 1 >= lo) --i;
	while (i < n && !((o + sp * i) + sp * 1 >= lo)) ++i;
	return i;
}
__device__ __forceinline__ int last_cell_min_le(double hi, double o, double sp, int n) {
	// largest i in [-1, n-1] with o + sp*i <= hi             (cell.min > tri.max fails)
	const double t = (hi - o) * __drcp_rn(sp), ft = floor(t);
	if (occ_index_safe(t, ft)) return max(-1, min(n - 1, (int)ft));     // largest i with i <= t
	int i = (int)fmax(-2.0, fmin(ft, 2147483000.0));                    // guess only
	if (i < -1) i = -1;
	if (i > n - 1) i = n - 1;
	while (i < n - 1 && o + sp * (i + 1) <= hi) ++i;
	while (i >= 0 && !(o + sp * i <= hi)) --i;
	return i;
}

struct OccGrid { int nx, ny, nz; double ox, oy, oz, sp; };
#define OCC_INLINE 64
// Facets stamp BITS, a streaming kernel then writes every output byte exactly once.  (The first version zeroed the byte grid and
// scattered single byte stores into it: every stamped voxel cost a 32-byte sector read-modify-write on top of the 1 GiB memset,
// 0.79 ms at 1024^3, 23 % of HBM peak.)  The bits are TILED: one 32-bit word = a 4 x 4 x 2 block of voxels (bit = (z&1)*16 +
// (y&3)*4 + (x&3)), words x-fastest over the tile grid, so a 32-byte sector holds 32 x 4 x 2 voxels and the box of a facet — a few
// voxels in every direction — lies in 1 - 4 sectors.  With one bit ROW per x-run (round-2 first form) a 3 x 3 x 3 box touched 9
// different sectors; ncu: 34 M sector loads = 1.1 GB of L2 traffic for 2 M facets, 203 us, L2-bandwidth bound.
__device__ __forceinline__ uint32_t occ_tile_mask(int x0, int x1, int y0, int y1, int z0, int z1, int tx, int ty, int tz) {
	// voxels of the closed box [x0,x1] x [y0,y1] x [z0,z1] that lie in tile (tx, ty, tz); the box meets the tile
	const int xa = max(x0 - 4 * tx, 0), xb = min(x1 - 4 * tx, 3), ya = max(y0 - 4 * ty, 0), yb = min(y1 - 4 * ty, 3);
	const int za = max(z0 - 2 * tz, 0), zb = min(z1 - 2 * tz, 1);
	const uint32_t mx = ((1u << (xb - xa + 1)) - 1u) << xa, yb4 = ((1u << (yb - ya + 1)) - 1u) << ya;
	const uint32_t spread = (yb4 & 1u) | ((yb4 & 2u) << 3) | ((yb4 & 4u) << 6) | ((yb4 & 8u) << 9);      // bit i -> bit 4 i
	const uint32_t layer = mx * spread;                                                                    // 16 bits, no carries (mx < 16)
	return (za == 0 ? layer : 0u) | (zb == 1 ? layer << 16 : 0u);
}
// One warp = 32 facets.  The tiles of the small boxes of a warp are POOLED: a warp prefix sum over the tile counts, then the lanes
// take (facet, tile) pairs OCC_MLP x 32 at a time whatever facet they belong to; all words of a step are LOADED first (independent
// loads in flight), then compared and OR-ed (an atomic only where a bit is missing).  Boxes over OCC_INLINE voxels go to a compact
// list: slots AND tile offsets reserved with one packed 64-bit atomicAdd per warp (count << OCC_ROW_BITS | tiles), so the list is
// ordered by offset and occupancy_tiles_kernel binary-searches it — no scan over all facets.
#define OCC_ROW_BITS 36
#define OCC_MLP 4
__global__ void __launch_bounds__(256)
occupancy_boxes_kernel(OccGrid g, const double *__restrict__ tri, int64_t nF, int *__restrict__ big_box6, int64_t *__restrict__ big_off,
                       unsigned long long *__restrict__ ctl /* [0] packed counter, [1] exact tile sum */, uint32_t *__restrict__ bits)
{
	const int lane = threadIdx.x & 31;
	const int ntx = (g.nx + 3) >> 2, nty = (g.ny + 3) >> 2;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < nF; base += stride) {   // warp-uniform trip count
		const int64_t f = base + lane;
		int lo[3] = {0, 0, 0}, hi[3] = {-1, -1, -1}, n_small = 0, nwx = 0, nwy = 0;
		int64_t n_big = 0;
		if (f < nF) {
			const double *t = tri + 9 * f;
			const int n[3] = {g.nx, g.ny, g.nz};
			const double o[3] = {g.ox, g.oy, g.oz};
			int64_t vol = 1;
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const double mn = fmin(t[c], fmin(t[3 + c], t[6 + c])), mx = fmax(t[c], fmax(t[3 + c], t[6 + c]));
				lo[c] = first_cell_max_ge(mn, o[c], g.sp, n[c]);
				hi[c] = last_cell_min_le(mx, o[c], g.sp, n[c]);
				vol *= hi[c] >= lo[c] ? hi[c] - lo[c] + 1 : 0;
			}
			if (vol > 0) {
				nwx = (hi[0] >> 2) - (lo[0] >> 2) + 1; nwy = (hi[1] >> 2) - (lo[1] >> 2) + 1;
				const int64_t tiles = (int64_t)nwx * nwy * ((hi[2] >> 1) - (lo[2] >> 1) + 1);
				if (vol <= OCC_INLINE) n_small = (int)tiles; else n_big = tiles;
			}
		}
		const unsigned big_mask = __ballot_sync(0xffffffffu, n_big > 0);
		if (big_mask) {                                   // rare on fine meshes; warp-uniform
			long long incl64 = n_big;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, incl64, o); if (lane >= o) incl64 += v; }
			const long long total64 = __shfl_sync(0xffffffffu, incl64, 31);
			unsigned long long old = 0;
			if (lane == 0) {
				old = atomicAdd(ctl, ((unsigned long long)__popc(big_mask) << OCC_ROW_BITS) + (unsigned long long)total64);
				atomicAdd(ctl + 1, (unsigned long long)total64);
			}
			old = __shfl_sync(0xffffffffu, old, 0);
			if (n_big > 0) {
				const int64_t pos = (int64_t)(old >> OCC_ROW_BITS) + __popc(big_mask & ((1u << lane) - 1u));
				for (int c = 0; c < 3; ++c) { big_box6[6 * pos + c] = lo[c]; big_box6[6 * pos + 3 + c] = hi[c]; }
				big_off[pos] = (int64_t)(old & ((1ull << OCC_ROW_BITS) - 1ull)) + (incl64 - n_big);
			}
		}
		int incl = n_small;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
		const int excl = incl - n_small;
		const int total = __shfl_sync(0xffffffffu, incl, 31);
		for (int p0 = 0; p0 < total; p0 += 32 * OCC_MLP) {
			int64_t w[OCC_MLP];
			uint32_t m[OCC_MLP], v[OCC_MLP];
#pragma unroll
			for (int u = 0; u < OCC_MLP; ++u) {
				const int p = p0 + 32 * u + lane;
				int src = 0;                                    // largest lane with excl <= p
#pragma unroll
				for (int step = 16; step > 0; step >>= 1) {
					const int e = __shfl_sync(0xffffffffu, excl, (src + step) & 31);
					if (src + step < 32 && e <= p) src += step;
				}
				const int x0 = __shfl_sync(0xffffffffu, lo[0], src), y0 = __shfl_sync(0xffffffffu, lo[1], src), z0 = __shfl_sync(0xffffffffu, lo[2], src);
				const int x1 = __shfl_sync(0xffffffffu, hi[0], src), y1 = __shfl_sync(0xffffffffu, hi[1], src), z1 = __shfl_sync(0xffffffffu, hi[2], src);
				const int swx = __shfl_sync(0xffffffffu, nwx, src), swy = __shfl_sync(0xffffffffu, nwy, src), se = __shfl_sync(0xffffffffu, excl, src);
				w[u] = 0; m[u] = 0;
				if (p < total) {
					// tile number inside the box, x fastest.  A small box (<= 64 voxels) has fewer than 64 tiles and no side over 17: the
					// quotients (k + 0.5) / d stay at least 0.5 / 64 away from an integer, far beyond the float error of the reciprocal
					const int k = p - se, sxy = swx * swy;
					const int kz = __float2int_rz(((float)k + 0.5f) * __frcp_rn((float)sxy)), kr = k - kz * sxy;
					const int ky = __float2int_rz(((float)kr + 0.5f) * __frcp_rn((float)swx)), kx = kr - ky * swx;
					const int tx = (x0 >> 2) + kx, ty = (y0 >> 2) + ky, tz = (z0 >> 1) + kz;
					w[u] = ((int64_t)tz * nty + ty) * ntx + tx;
					m[u] = occ_tile_mask(x0, x1, y0, y1, z0, z1, tx, ty, tz);
				}
			}
#pragma unroll
			for (int u = 0; u < OCC_MLP; ++u) v[u] = m[u] ? bits[w[u]] : 0xffffffffu;
#pragma unroll
			for (int u = 0; u < OCC_MLP; ++u) if ((v[u] & m[u]) != m[u]) atomicOr(bits + w[u], m[u]);
		}
	}
}
__global__ void __launch_bounds__(256)
occupancy_tiles_kernel(OccGrid g, const int *__restrict__ big_box6, const int64_t *__restrict__ big_off, const unsigned long long *__restrict__ ctl,
                       uint32_t *__restrict__ bits)
{
	// counts are read on the device: no host round trip in front of this launch
	const int64_t n_big = (int64_t)(ctl[0] >> OCC_ROW_BITS), n_tiles = (int64_t)(ctl[0] & ((1ull << OCC_ROW_BITS) - 1ull));
	if (ctl[1] >> OCC_ROW_BITS) return;                // offsets wrapped: the host reports FPOHM_ERANGE
	const int ntx = (g.nx + 3) >> 2, nty = (g.ny + 3) >> 2;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_tiles; t += (int64_t)gridDim.x * blockDim.x) {
		int64_t lo = 0, hi = n_big;
		while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (big_off[mid] <= t) lo = mid; else hi = mid; }
		const int *b = big_box6 + 6 * lo;                 // lo[3], hi[3]
		const int64_t k = t - big_off[lo];
		const int nwx = (b[3] >> 2) - (b[0] >> 2) + 1, nwy = (b[4] >> 2) - (b[1] >> 2) + 1;
		const int kz = (int)(k / ((int64_t)nwx * nwy)), kr = (int)(k - (int64_t)kz * nwx * nwy), ky = kr / nwx, kx = kr - ky * nwx;
		const int tx = (b[0] >> 2) + kx, ty = (b[1] >> 2) + ky, tz = (b[2] >> 1) + kz;
		const uint32_t m = occ_tile_mask(b[0], b[3], b[1], b[4], b[2], b[5], tx, ty, tz);
		uint32_t *wp = bits + ((int64_t)tz * nty + ty) * ntx + tx;
		if ((*wp & m) != m) atomicOr(wp, m);
	}
}
// tiled bits -> bytes.  One thread = one tile word: 4 bytes (4 voxels along x) into each of its 4 x 2 (y, z) rows, so a warp
// (32 consecutive tiles along x) stores 128 contiguous bytes per row — the store pattern of the voxel fill.
__device__ __forceinline__ uint32_t occ_nibble_bytes(uint32_t q) { return (q & 1u) | ((q & 2u) << 7) | ((q & 4u) << 14) | ((q & 8u) << 21); }
// Fast path (nx a multiple of 16): one thread = FOUR tile words (one 16-byte load) -> 16 bytes into each of their 4 x 2 (y, z) rows,
// a warp stores 512 contiguous bytes per row.  (One word and 4-byte stores per thread was latency bound: 335 us against 211 us.)
__global__ void __launch_bounds__(256)
occupancy_expand4_kernel(OccGrid g, const uint32_t *__restrict__ bits, uint8_t *__restrict__ out) {
	const unsigned ntx4 = (unsigned)(g.nx >> 4), nty = (unsigned)((g.ny + 3) >> 2), ntz = (unsigned)((g.nz + 1) >> 1);
	const unsigned n_items = ntx4 * nty * ntz;         // < 2^31 / 128
	const int64_t layer = (int64_t)g.nx * g.ny;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
		const unsigned row = it / ntx4, i = it - row * ntx4, tz = row / nty, ty = row - tz * nty;
		const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(bits) + it);
#pragma unroll
		for (int dz = 0; dz < 2; ++dz) {
			const int z = 2 * (int)tz + dz;
			if (z >= g.nz) break;
#pragma unroll
			for (int dy = 0; dy < 4; ++dy) {
				const int y = 4 * (int)ty + dy;
				if (y >= g.ny) break;
				const int sh = 16 * dz + 4 * dy;
				uint4 v;
				v.x = occ_nibble_bytes((wv.x >> sh) & 15u); v.y = occ_nibble_bytes((wv.y >> sh) & 15u);
				v.z = occ_nibble_bytes((wv.z >> sh) & 15u); v.w = occ_nibble_bytes((wv.w >> sh) & 15u);
				__stcs(reinterpret_cast<uint4 *>(out + (int64_t)z * layer + (int64_t)y * g.nx + 16 * (int64_t)i), v);   // index_from_index3, voxelization.cpp:26-28
			}
		}
	}
}
// any dims: one thread = one tile word
__global__ void __launch_bounds__(256)
occupancy_expand_kernel(OccGrid g, const uint32_t *__restrict__ bits, uint8_t *__restrict__ out) {
	const int ntx = (g.nx + 3) >> 2, nty = (g.ny + 3) >> 2, ntz = (g.nz + 1) >> 1;
	const int n_rows = nty * ntz;                      // rows of tiles
	const int64_t layer = (int64_t)g.nx * g.ny;
	const bool aligned = (g.nx & 3) == 0;
	for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {      // block-uniform: the two divisions are per row, not per tile
		const int ty = row % nty, tz = row / nty;
		for (int tx = threadIdx.x; tx < ntx; tx += blockDim.x) {
			const uint32_t wv = __ldg(bits + (int64_t)row * ntx + tx);
#pragma unroll
			for (int dz = 0; dz < 2; ++dz) {
				const int z = 2 * tz + dz;
				if (z >= g.nz) break;
#pragma unroll
				for (int dy = 0; dy < 4; ++dy) {
					const int y = 4 * ty + dy;
					if (y >= g.ny) break;
					const uint32_t q = (wv >> (16 * dz + 4 * dy)) & 15u;
					uint8_t *o = out + (int64_t)z * layer + (int64_t)y * g.nx + 4 * tx;
					if (aligned) __stcs(reinterpret_cast<uint32_t *>(o), occ_nibble_bytes(q));
					else for (int dx = 0; dx < 4 && 4 * tx + dx < g.nx; ++dx) o[dx] = (q >> dx) & 1u;
				}
			}
		}
	}
}

struct Stats { double mx, mean, rms; int64_t n; };

// one direction of Sampling<>::Hausdorff (sampling.h:547-602): samples = referenced vertices of A (+ optional face samples)
Stats directed(fpohm_ctx *ctx, fpohm_mesh *A, fpohm_mesh *B, int64_t extra, double upper, cudaStream_t s) {
	const int blk = 256;
	mesh_ensure_tree(ctx, B, s);
	DevBuf<uint8_t> flag(A->nV, s);
	flag.zero();
	mark_referenced_kernel<<<grid_for(ctx, 3 * A->nF, blk), blk, 0, s>>>(A->F.p, 3 * A->nF, flag.p);
	FPOHM_LAUNCH_CHECK(ctx);
	DevBuf<int32_t> ids(A->nV, s), sel(A->nV, s);
	DevBuf<int64_t> cnt(1, s);
	iota_kernel<<<grid_for(ctx, A->nV, blk), blk, 0, s>>>(ids.p, A->nV);
	FPOHM_LAUNCH_CHECK(ctx);
	size_t tb = 0;
	FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids.p, flag.p, sel.p, cnt.p, A->nV, s));
	DevBuf<uint8_t> tmp((int64_t)tb, s);
	FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, ids.p, flag.p, sel.p, cnt.p, A->nV, s));
	ctx->launches += 2;
	int64_t nv = 0;
	cnt.download(&nv, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	// optional face samples
	int64_t nf_samples = 0;
	DevBuf<int64_t> foff;
	DevBuf<int32_t> per_edge;
	if (extra > 0) {
		// VCG's similar-triangle rule is a SEQUENTIAL recurrence over the faces (extern/vcg/sampling.h:513-540): a running
		// decimal `n_samples_decimal += 0.5 * DoubleArea(f) * density`, n = (int) of it, m = samples per edge from n, and
		// what the lattice really produced ((m-2)(m-3)/2) is taken off again.  Identical sample sets need the identical
		// recurrence in the identical fp64 order, so the counts are made on the host from the mesh's own arrays (2 M faces:
		// ~10 ms); positions are then generated on the device.  DoubleArea = Norm((v1-v0)^(v2-v0)), sums left to right
		// (vcg/space/triangle3.h:241, deprecated_point3.h:282-290,321-324); area_S1 = sum / 2 (sampling.h:212-222);
		// density = (n_samples_target - n_vertex_samples) / area_S1 (sampling.h:573-586).
		const double *hV = A->hV.data();
		const int32_t *hF = A->hF.data();
		std::vector<double> da((size_t)A->nF);
		double area2 = 0.0;
		for (int64_t f = 0; f < A->nF; ++f) {
			const double *v0 = hV + 3 * (int64_t)hF[3 * f], *v1 = hV + 3 * (int64_t)hF[3 * f + 1], *v2 = hV + 3 * (int64_t)hF[3 * f + 2];
			const double ax = v1[0] - v0[0], ay = v1[1] - v0[1], az = v1[2] - v0[2];
			const double bx = v2[0] - v0[0], by = v2[1] - v0[1], bz = v2[2] - v0[2];
			const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
			da[(size_t)f] = std::sqrt(cx * cx + cy * cy + cz * cz);
			area2 += da[(size_t)f];
		}
		const double area_S1 = area2 / 2.0;
		const double density = (double)(unsigned long)extra / area_S1;
		std::vector<int64_t> hoff((size_t)A->nF + 1);
		std::vector<int32_t> hper((size_t)A->nF);
		double dec = 0.0;
		int64_t tot = 0;
		for (int64_t f = 0; f < A->nF; ++f) {
			dec += 0.5 * da[(size_t)f] * density;
			const int n = (int)dec;
			int m = 0; int64_t got = 0;
			if (n) {
				m = (int)((std::sqrt(1.0 + 8.0 * (double)n) + 5.0) / 2.0);
				got = (int64_t)(m - 2) * (m - 3) / 2;      // i = 1..m-2, j = 1..m-2-i
			}
			hoff[(size_t)f] = tot; hper[(size_t)f] = m >= 4 ? m : 0;
			tot += got;
			dec -= (double)got;
		}
		hoff[(size_t)A->nF] = tot;
		nf_samples = tot;
		foff.alloc(A->nF + 1, s); per_edge.alloc(A->nF, s);
		foff.upload(hoff.data(), A->nF + 1); per_edge.upload(hper.data(), A->nF);
		FPOHM_CUDA(cudaStreamSynchronize(s));      // the staging vectors are locals
	}
	const int64_t n = nv + nf_samples;
	DevBuf<double> P(3 * n, s), sq(n, s);
	gather_points_kernel<<<grid_for(ctx, 3 * nv, blk), blk, 0, s>>>(A->V.p, sel.p, nv, P.p);
	FPOHM_LAUNCH_CHECK(ctx);
	if (nf_samples) {
		face_samples_kernel<<<grid_for(ctx, A->nF, blk), blk, 0, s>>>(A->tri.p, A->nF, foff.p, per_edge.p, P.p + 3 * nv);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	launch_closest_point(ctx, B, false, P.p, n, sq.p, nullptr, nullptr, nullptr, s);
	const int grid = grid_for(ctx, n, blk, 4);
	DevBuf<DistPartial> part(grid, s);
	dist_stats_kernel<<<grid, blk, 0, s>>>(sq.p, n, upper, part.p);
	FPOHM_LAUNCH_CHECK(ctx);
	std::vector<DistPartial> hp((size_t)grid);
	part.download(hp.data(), grid);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	Stats st{-HUGE_VAL, 0, 0, 0};
	double sum = 0, sumsq = 0;
	for (auto &p : hp) { st.mx = std::max(st.mx, p.mx); sum += p.sum; sumsq += p.sumsq; st.n += p.n; } // fixed order
	st.mean = sum / (double)st.n;
	st.rms = std::sqrt(sumsq / (double)st.n);
	return st;
}

} // namespace

extern "C" {

int fpohm_polyline_project(fpohm_ctx *ctx, const double *Vc, int64_t nVc, const int64_t *curve_off, const int32_t *curve_vs,
                           const uint8_t *circle, int64_t n_curves, const double *P, const int32_t *curve_id, int64_t np,
                           double *origin_L, double *axis_L)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && Vc && curve_off && curve_vs && circle && P && curve_id && origin_L && axis_L && nVc > 0 && n_curves > 0 && np >= 0,
	              FPOHM_EINVAL, "fpohm_polyline_project: bad argument");
	const int64_t tot = curve_off[n_curves];
	for (int64_t c = 0; c < n_curves; ++c)
		FPOHM_REQUIRE(curve_off[c + 1] - curve_off[c] >= 2, FPOHM_EINVAL, "fpohm_polyline_project: curve %lld has fewer than 2 vertices", (long long)c);
	for (int64_t i = 0; i < tot; ++i) FPOHM_REQUIRE(curve_vs[i] >= 0 && curve_vs[i] < nVc, FPOHM_EINVAL, "fpohm_polyline_project: vertex id out of range");
	for (int64_t i = 0; i < np; ++i) FPOHM_REQUIRE(curve_id[i] >= 0 && curve_id[i] < n_curves, FPOHM_EINVAL, "fpohm_polyline_project: curve id out of range");
	if (np == 0) return FPOHM_OK;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dV(3 * nVc, s), dP(3 * np, s), dO(3 * np, s), dA(3 * np, s);
	DevBuf<int64_t> doff(n_curves + 1, s);
	DevBuf<int32_t> dcv(tot, s), dcid(np, s);
	DevBuf<uint8_t> dci(n_curves, s);
	dV.upload(Vc, 3 * nVc); dP.upload(P, 3 * np); doff.upload(curve_off, n_curves + 1); dcv.upload(curve_vs, tot);
	dcid.upload(curve_id, np); dci.upload(circle, n_curves);
	KernelTimer t(ctx, s);
	polyline_kernel<<<grid_for(ctx, np, 128), 128, 0, s>>>(dV.p, doff.p, dcv.p, dci.p, dP.p, dcid.p, np, dO.p, dA.p);
	FPOHM_LAUNCH_CHECK(ctx);
	t.stop();
	dO.download(origin_L, 3 * np); dA.download(axis_L, 3 * np);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_hausdorff(fpohm_ctx *ctx, fpohm_mesh *A, fpohm_mesh *B, int64_t extra_face_samples, double out[7], int64_t n_samples[2]) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && A && B && out && extra_face_samples >= 0, FPOHM_EINVAL, "fpohm_hausdorff: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	// joint bbox over ALL vertices, inflated by 2 % of its diagonal (metro_hausdorff.cpp:113-118; vcg Box3::Diag, Offset)
	double mn[3], mx[3];
	for (int c = 0; c < 3; ++c) { mn[c] = std::min(A->bbox[c], B->bbox[c]); mx[c] = std::max(A->bbox[3 + c], B->bbox[3 + c]); }
	auto diag = [&]() { const double dx = mn[0] - mx[0], dy = mn[1] - mx[1], dz = mn[2] - mx[2]; return std::sqrt(dx * dx + dy * dy + dz * dz); };
	const double off = diag() * 0.02;
	for (int c = 0; c < 3; ++c) { mn[c] -= off; mx[c] += off; }
	const double D = diag();
	KernelTimer t(ctx, s);
	const Stats ab = directed(ctx, A, B, extra_face_samples, D, s);
	const Stats ba = directed(ctx, B, A, extra_face_samples, D, s);
	t.stop();
	out[0] = D; out[1] = ab.mx; out[2] = ba.mx; out[3] = ab.mean; out[4] = ba.mean; out[5] = ab.rms; out[6] = ba.rms;
	if (n_samples) { n_samples[0] = ab.n; n_samples[1] = ba.n; }
	FPOHM_API_END
}

int fpohm_hausdorff_outliers(fpohm_ctx *ctx, fpohm_mesh *A, fpohm_mesh *B, double dis_threshold, int32_t *outlier_vs, int64_t *n_outliers) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && A && B && outlier_vs && n_outliers, FPOHM_EINVAL, "fpohm_hausdorff_outliers: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	mesh_ensure_tree(ctx, A, s); mesh_ensure_tree(ctx, B, s);
	DevBuf<double> sqAB(A->nV, s), sqBA(B->nV, s);
	DevBuf<int32_t> I0(A->nV, s);
	KernelTimer t(ctx, s);
	launch_closest_point(ctx, B, false, A->V.p, A->nV, sqAB.p, I0.p, nullptr, nullptr, s);   // point_mesh_squared_distance(A, B, FB)
	launch_closest_point(ctx, A, false, B->V.p, B->nV, sqBA.p, nullptr, nullptr, nullptr, s); // point_mesh_squared_distance(B, A, FA)
	double thr = dis_threshold * dis_threshold;
	DevBuf<uint8_t> flag(B->nV, s);
	DevBuf<int32_t> ids(B->nV, s), sel(B->nV, s);
	DevBuf<int64_t> cnt(1, s);
	iota_kernel<<<grid_for(ctx, B->nV, blk), blk, 0, s>>>(ids.p, B->nV);
	FPOHM_LAUNCH_CHECK(ctx);
	flag.zero();
	int64_t n = 0;
	for (int round = 0; round < 4096 && n == 0; ++round) {      // `while (!outlierVs.size())`, flags persist across rounds
		outlier_flag_kernel<<<grid_for(ctx, A->nV + B->nV, blk), blk, 0, s>>>(sqAB.p, I0.p, A->nV, B->F.p, sqBA.p, B->nV, thr, flag.p);
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids.p, flag.p, sel.p, cnt.p, B->nV, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, ids.p, flag.p, sel.p, cnt.p, B->nV, s));
		ctx->launches += 2;
		cnt.download(&n, 1);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		thr *= 0.9;
	}
	t.stop();
	*n_outliers = n;
	sel.download(outlier_vs, n);   // ascending vertex id; the reference's push order differs, the SET is identical
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_voxel_lattice_dims(const double bb_min[3], const double bb_max[3], int32_t num_voxels, int32_t dim[3]) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(bb_min && bb_max && dim && num_voxels > 0, FPOHM_EINVAL, "fpohm_voxel_lattice_dims: bad argument");
	double extent[3];
	for (int c = 0; c < 3; ++c) extent[c] = bb_max[c] - bb_min[c];
	const double max_extent = std::max(extent[0], std::max(extent[1], extent[2]));
	const double len = max_extent / num_voxels;
	for (int c = 0; c < 3; ++c) dim[c] = (int32_t)std::ceil(extent[c] / len);
	FPOHM_API_END
}

int fpohm_voxel_lattice(fpohm_ctx *ctx, const double bb_min[3], const double bb_max[3], int32_t num_voxels, double *Vpos, uint32_t *hex) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && bb_min && bb_max && num_voxels > 0, FPOHM_EINVAL, "fpohm_voxel_lattice: bad argument");
	int32_t d[3];
	int rc = fpohm_voxel_lattice_dims(bb_min, bb_max, num_voxels, d);
	if (rc) return rc;
	FPOHM_REQUIRE(d[0] >= 2 && d[1] >= 2 && d[2] >= 2, FPOHM_EINVAL, "fpohm_voxel_lattice: degenerate lattice %d x %d x %d", d[0], d[1], d[2]);
	float gl[3];
	for (int c = 0; c < 3; ++c) gl[c] = (float)((bb_max[c] - bb_min[c]) / d[c]);
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int64_t nv = (int64_t)d[0] * d[1] * d[2], nh = (int64_t)(d[0] - 1) * (d[1] - 1) * (d[2] - 1);
	FPOHM_REQUIRE(nv < (1ll << 31), FPOHM_ERANGE, "fpohm_voxel_lattice: too many vertices");
	KernelTimer t(ctx, s);
	if (Vpos) {
		DevBuf<double> dV(3 * nv, s);
		lattice_vertices_kernel<<<grid_for(ctx, nv, 256), 256, 0, s>>>(bb_min[0], bb_min[1], bb_min[2], gl[0], gl[1], gl[2], d[0], d[1], d[2], dV.p);
		FPOHM_LAUNCH_CHECK(ctx);
		dV.download(Vpos, 3 * nv);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	if (hex) {
		DevBuf<uint32_t> dH(8 * nh, s);
		lattice_hexes_kernel<<<grid_for(ctx, nh, 256), 256, 0, s>>>(d[0], d[1], d[2], dH.p);
		FPOHM_LAUNCH_CHECK(ctx);
		dH.download(hex, 8 * nh);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	t.stop();
	FPOHM_API_END
}

int fpohm_voxel_occupancy(fpohm_ctx *ctx, const fpohm_mesh *mesh, const double grid_origin[3], double spacing, const int32_t dims[3], uint8_t *out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && grid_origin && dims && out && spacing > 0, FPOHM_EINVAL, "fpohm_voxel_occupancy: bad argument");
	const int64_t n = (int64_t)dims[0] * dims[1] * dims[2];
	FPOHM_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && n < (1ll << 31), FPOHM_ERANGE, "fpohm_voxel_occupancy: bad dims");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int64_t nF = mesh->nF;
	FPOHM_REQUIRE(nF < (1ll << (64 - OCC_ROW_BITS)), FPOHM_ERANGE, "fpohm_voxel_occupancy: %lld facets (the limit is 2^28)", (long long)nF);
	DevBuf<uint8_t> d(n, s);
	const int64_t n_words = ((int64_t)((dims[0] + 3) / 4) * ((dims[1] + 3) / 4) * ((dims[2] + 1) / 2) + 1) & ~(int64_t)1;   // tiles; even: the 64-bit control words behind them are aligned
	DevBuf<uint32_t> bits(n_words + 4, s);
	unsigned long long *ctl = reinterpret_cast<unsigned long long *>(bits.p + n_words);
	KernelTimer t(ctx, s);
	bits.zero();                                       // bits and control words in one memset
	const OccGrid og{dims[0], dims[1], dims[2], grid_origin[0], grid_origin[1], grid_origin[2], spacing};
	DevBuf<int> box6(6 * nF, s);                        // touched only where a facet's box holds more than OCC_INLINE voxels
	DevBuf<int64_t> off(nF, s);
	occupancy_boxes_kernel<<<grid_for(ctx, nF, 256), 256, 0, s>>>(og, mesh->tri.p, nF, box6.p, off.p, ctl, bits.p);
	FPOHM_LAUNCH_CHECK(ctx);
	occupancy_tiles_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(og, box6.p, off.p, ctl, bits.p);
	FPOHM_LAUNCH_CHECK(ctx);
	if ((dims[0] & 15) == 0) {
		const int64_t items = (int64_t)(dims[0] / 16) * ((dims[1] + 3) / 4) * ((dims[2] + 1) / 2);
		static const int ex_ctas = getenv("FPOHM_OCC_CTAS") ? atoi(getenv("FPOHM_OCC_CTAS")) : 256;     // CTAs per SM in the grid (1024^3: one item per thread).  16 / 32 / 64 / 128 / 256: 0.367 / 0.352 / 0.347 / 0.343 / 0.334 ms; the bare store pattern goes 6.5 -> 7.1 TB/s from 16 to 64 (scripts/micro/write_patterns.cu, rows16x8)
		occupancy_expand4_kernel<<<grid_for(ctx, items, 256, ex_ctas), 256, 0, s>>>(og, bits.p, d.p);
	} else {
		occupancy_expand_kernel<<<(int)std::min<int64_t>((int64_t)((dims[1] + 3) / 4) * ((dims[2] + 1) / 2), (int64_t)ctx->sm_count * 32), 256, 0, s>>>(og, bits.p, d.p);
	}
	FPOHM_LAUNCH_CHECK(ctx);
	t.stop();
	unsigned long long h_ctl[2] = {0, 0};
	FPOHM_CUDA(cudaMemcpyAsync(h_ctl, ctl, 16, cudaMemcpyDeviceToHost, s));
	d.download(out, n);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_REQUIRE((h_ctl[1] >> OCC_ROW_BITS) == 0, FPOHM_ERANGE, "fpohm_voxel_occupancy: %llu box tiles (the limit is 2^%d)", h_ctl[1], OCC_ROW_BITS);
	FPOHM_API_END
}

} // extern "C"

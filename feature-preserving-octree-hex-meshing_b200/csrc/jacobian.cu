// Per-hex scaled Jacobian: scaled_jacobian Hex branch (gf.cpp:2309-2358) + a_jacobian (gf.cpp:2422-2442).
//
// HBM-bound map + reduce.  Algorithmic bytes per hex: 32 B corner ids read + 64 B V_Js + 8 B H_Js written,
// plus every vertex (24 B) once (SURVEY.md §8d "J1": 104 B/hex + 24 B/vertex).
//   pass 1  jacobian_kernel      one thread per hex: 2 x 16 B id loads, 8 vertex gathers (L2-resident reuse x8
//                                on structured meshes), 8 corner determinants, 4 x 16 B + 8 B stores,
//                                block partials of {min, sum, flipped}
//   pass 2  finalize_kernel<0>   fixed-order reduction of the partials -> min, ave, flipped
//   pass 3  deviation_kernel     sum (H_J - ave)^2, block partials (8 B/hex re-read, L2 hits for < 126 MB)
//   pass 4  finalize_kernel<1>   -> deviation
// All reductions are fixed-order trees, so results are run-to-run and shard-count stable.
#include "internal.h"
#include <math_constants.h>

using namespace fpohm;

namespace {

struct V3 { double x, y, z; };

// The 8 corner frames of hex_tetra_table use the 12 hex edges twice each, once per end point and with opposite
// orientation.  IEEE subtraction is exactly antisymmetric and *0.5 is exact, so (a - b)*.5 == -((b - a)*.5) bit for bit,
// the norms of the two are identical, and negating a column of the 3x3 negates every product and every partial sum of
// Eigen's cofactor expansion exactly.  Hence: 12 half-edge vectors + 12 norms per hex (instead of 24 + 24), each corner
// picks +-e and evaluates the SAME expression tree as a_jacobian (gf.cpp:2422-2442) — results are bit-identical to the
// reference while the fp64 pipe, which bounds this kernel (ncu: 44 % fp64 pipe vs 19 % DRAM on the first version),
// does ~1/3 less work.
//   edge k = (lo, hi): 0:(0,1) 1:(1,2) 2:(3,2) 3:(0,3) 4:(4,5) 5:(5,6) 6:(7,6) 7:(4,7) 8:(0,4) 9:(1,5) 10:(2,6) 11:(3,7)
//   corner j of hex_tetra_table (global_types.h:163-173) -> columns (edge, sign): sign +1 means e = hi - lo is used as is

// a_jacobian(Vector3d...), gf.cpp:2422-2442 on precomputed halved columns; Eigen 3.2 determinant (bruteforce_det3_helper)
__device__ __forceinline__ double a_jacobian_cols(const V3 &c0, const V3 &c1, const V3 &c2, double norm1, double norm2, double norm3) {
	const double m00 = c0.x, m10 = c0.y, m20 = c0.z, m01 = c1.x, m11 = c1.y, m21 = c1.z, m02 = c2.x, m12 = c2.y, m22 = c2.z;
	const double det = m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
	if (norm1 < 1.e-7 || norm2 < 1.e-7 || norm3 < 1.e-7) return det; // "Potential Bug" branch, gf.cpp:2436-2439
	return det / (norm1 * norm2 * norm3);
}

struct Partial { double mn, sum; long long flipped; };

__device__ __forceinline__ void block_reduce(Partial &p, Partial *smem) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		p.mn = fmin(p.mn, __shfl_down_sync(0xffffffffu, p.mn, o));
		p.sum += __shfl_down_sync(0xffffffffu, p.sum, o);
		p.flipped += __shfl_down_sync(0xffffffffu, p.flipped, o);
	}
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) smem[w] = p;
	__syncthreads();
	if (w == 0) {
		const int nw = blockDim.x >> 5;
		p = l < nw ? smem[l] : Partial{CUDART_INF, 0.0, 0};
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			p.mn = fmin(p.mn, __shfl_down_sync(0xffffffffu, p.mn, o));
			p.sum += __shfl_down_sync(0xffffffffu, p.sum, o);
			p.flipped += __shfl_down_sync(0xffffffffu, p.flipped, o);
		}
	}
}

__global__ void __launch_bounds__(128)
jacobian_kernel(const double *__restrict__ V, const uint32_t *__restrict__ hex, int64_t H,
                double *__restrict__ V_Js, double *__restrict__ H_Js, Partial *__restrict__ partials)
{
	__shared__ Partial smem[32];
	Partial acc{CUDART_INF, 0.0, 0};
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < H; i += (int64_t)gridDim.x * blockDim.x) {
		const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(hex + 8 * i));
		const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(hex + 8 * i) + 1);
		const uint32_t id[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
		V3 p[8];
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const double *v = V + 3 * (int64_t)id[k];
			p[k] = {__ldg(v), __ldg(v + 1), __ldg(v + 2)};
		}
		// 12 halved edge vectors (hi - lo) * .5 and their norms (Eigen: sqrt(x^2 + (y^2 + z^2)))
		constexpr int elo[12] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3}, ehi[12] = {1, 2, 2, 3, 5, 6, 6, 7, 4, 5, 6, 7};
		V3 e[12]; double en[12];
#pragma unroll
		for (int k = 0; k < 12; ++k) {
			e[k] = {(p[ehi[k]].x - p[elo[k]].x) * .5, (p[ehi[k]].y - p[elo[k]].y) * .5, (p[ehi[k]].z - p[elo[k]].z) * .5};
			en[k] = sqrt(e[k].x * e[k].x + (e[k].y * e[k].y + e[k].z * e[k].z));
		}
		constexpr int ce[8][3] = {{3, 8, 0}, {0, 9, 1}, {1, 10, 2}, {2, 11, 3}, {7, 4, 8}, {4, 5, 9}, {5, 6, 10}, {6, 7, 11}};
		constexpr int cs[8][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, 1, -1}, {1, 1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, -1}};
		double j[8];
		double hex_min = 1;
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			V3 col[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const V3 &v = e[ce[c][k]];
				col[k] = cs[c][k] > 0 ? v : V3{-v.x, -v.y, -v.z};
			}
			j[c] = a_jacobian_cols(col[0], col[1], col[2], en[ce[c][0]], en[ce[c][1]], en[ce[c][2]]);
			if (hex_min > j[c]) hex_min = j[c];
		}
		if (V_Js) {
			double2 *o = reinterpret_cast<double2 *>(V_Js + 8 * i);
			o[0] = make_double2(j[0], j[1]); o[1] = make_double2(j[2], j[3]);
			o[2] = make_double2(j[4], j[5]); o[3] = make_double2(j[6], j[7]);
		}
		H_Js[i] = hex_min;
		acc.mn = fmin(acc.mn, hex_min);
		acc.sum += hex_min;
		acc.flipped += hex_min < 0;
	}
	block_reduce(acc, smem);
	if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256)
deviation_kernel(const double *__restrict__ H_Js, int64_t H, const double *__restrict__ stats, Partial *__restrict__ partials) {
	__shared__ Partial smem[32];
	const double ave = stats[1];
	Partial acc{CUDART_INF, 0.0, 0};
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < H; i += (int64_t)gridDim.x * blockDim.x) {
		const double d = H_Js[i] - ave;
		acc.sum += d * d;
	}
	block_reduce(acc, smem);
	if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// MODE 0: partials -> stats = {min(1, min), sum / H, .}, flipped.  MODE 1: partials -> stats[2] = sum / H
template <int MODE>
__global__ void __launch_bounds__(256)
finalize_kernel(const Partial *__restrict__ partials, int n, int64_t H, double *__restrict__ stats, long long *__restrict__ flipped) {
	__shared__ Partial smem[32];
	Partial acc{CUDART_INF, 0.0, 0};
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		acc.mn = fmin(acc.mn, partials[i].mn);
		acc.sum += partials[i].sum;
		acc.flipped += partials[i].flipped;
	}
	block_reduce(acc, smem);
	if (threadIdx.x == 0) {
		if (MODE == 0) {
			stats[0] = fmin(1.0, acc.mn);         // mq.min_Jacobian starts at 1 (gf.cpp:2313)
			stats[1] = acc.sum / (double)H;
			if (flipped) *flipped = acc.flipped;
		} else {
			stats[2] = acc.sum / (double)H;
		}
	}
}

} // namespace

namespace fpohm {

void launch_scaled_jacobian(fpohm_ctx *ctx, const double *V, const uint32_t *hex, int64_t H, double *V_Js, double *H_Js,
                            double *stats3, long long *flipped, cudaStream_t s)
{
	const int blk = 256;
	const int jblk = 128;                        // 144 registers/thread: 3 CTAs of 128 per SM keep 12 warps resident
	const int grid = grid_for(ctx, H, jblk, 6);
	DevBuf<Partial> partials(grid, s);
	DevBuf<double> tmpH;
	if (!H_Js) { tmpH.alloc(H, s); H_Js = tmpH.p; }
	jacobian_kernel<<<grid, jblk, 0, s>>>(V, hex, H, V_Js, H_Js, partials.p);
	FPOHM_LAUNCH_CHECK(ctx);
	finalize_kernel<0><<<1, blk, 0, s>>>(partials.p, grid, H, stats3, flipped);
	FPOHM_LAUNCH_CHECK(ctx);
	deviation_kernel<<<grid, blk, 0, s>>>(H_Js, H, stats3, partials.p);
	FPOHM_LAUNCH_CHECK(ctx);
	finalize_kernel<1><<<1, blk, 0, s>>>(partials.p, grid, H, stats3, nullptr);
	FPOHM_LAUNCH_CHECK(ctx);
}

} // namespace fpohm

extern "C" {

int fpohm_scaled_jacobian_dev(fpohm_ctx *ctx, const double *V_dev, int64_t nV, const uint32_t *hex_dev, int64_t H,
                              double *V_Js_dev, double *H_Js_dev, double *min_ave_dev_dev, int64_t *flipped_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && V_dev && hex_dev && min_ave_dev_dev, FPOHM_EINVAL, "fpohm_scaled_jacobian_dev: null argument");
	FPOHM_REQUIRE(nV > 0 && H > 0, FPOHM_EINVAL, "fpohm_scaled_jacobian_dev: empty mesh");
	FPOHM_REQUIRE(((uintptr_t)hex_dev & 15) == 0 && (!V_Js_dev || ((uintptr_t)V_Js_dev & 15) == 0), FPOHM_EINVAL,
	              "fpohm_scaled_jacobian_dev: hex / V_Js must be 16-byte aligned");
	DeviceGuard g(ctx->device);
	launch_scaled_jacobian(ctx, V_dev, hex_dev, H, V_Js_dev, H_Js_dev, min_ave_dev_dev, (long long *)flipped_dev, (cudaStream_t)stream);
	FPOHM_API_END
}

int fpohm_scaled_jacobian(fpohm_ctx *ctx, const double *V, int64_t nV, const uint32_t *hex, int64_t H,
                          double *V_Js, double *H_Js, double min_ave_dev[3], int64_t *flipped)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && V && hex && min_ave_dev, FPOHM_EINVAL, "fpohm_scaled_jacobian: null argument");
	FPOHM_REQUIRE(nV > 0 && H > 0, FPOHM_EINVAL, "fpohm_scaled_jacobian: empty mesh (nV=%lld H=%lld)", (long long)nV, (long long)H);
	for (int64_t i = 0; i < 8 * H; ++i)
		FPOHM_REQUIRE((int64_t)hex[i] < nV, FPOHM_EINVAL, "fpohm_scaled_jacobian: corner id %u out of range at %lld", hex[i], (long long)i);
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dV(3 * nV, s), dVJ(V_Js ? 8 * H : 0, s), dHJ(H, s), dstats(3, s);
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<long long> dfl(1, s);
	dV.upload(V, 3 * nV);
	dhex.upload(hex, 8 * H);
	KernelTimer t(ctx, s);
	launch_scaled_jacobian(ctx, dV.p, dhex.p, H, dVJ.p, dHJ.p, dstats.p, dfl.p, s);
	t.stop();
	if (V_Js) dVJ.download(V_Js, 8 * H);
	if (H_Js) dHJ.download(H_Js, H);
	dstats.download(min_ave_dev, 3);
	long long fl = 0;
	dfl.download(&fl, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (flipped) *flipped = fl;
	FPOHM_API_END
}

} // extern "C"

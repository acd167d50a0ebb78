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

// global_types.h:163-173
__constant__ int c_hex_tetra[8][4] = {
	{0, 3, 4, 1}, {1, 0, 5, 2}, {2, 1, 6, 3}, {3, 2, 7, 0}, {4, 7, 5, 0}, {5, 4, 6, 1}, {6, 5, 7, 2}, {7, 6, 4, 3}};

struct V3 { double x, y, z; };

// a_jacobian(Vector3d...), gf.cpp:2422-2442; Eigen 3.2 determinant (bruteforce_det3_helper) and norm association
__device__ __forceinline__ double a_jacobian(const V3 &v0, const V3 &v1, const V3 &v2, const V3 &v3) {
	const double m00 = (v1.x - v0.x) * .5, m10 = (v1.y - v0.y) * .5, m20 = (v1.z - v0.z) * .5;
	const double m01 = (v2.x - v0.x) * .5, m11 = (v2.y - v0.y) * .5, m21 = (v2.z - v0.z) * .5;
	const double m02 = (v3.x - v0.x) * .5, m12 = (v3.y - v0.y) * .5, m22 = (v3.z - v0.z) * .5;
	const double norm1 = sqrt(m00 * m00 + (m10 * m10 + m20 * m20));
	const double norm2 = sqrt(m01 * m01 + (m11 * m11 + m21 * m21));
	const double norm3 = sqrt(m02 * m02 + (m12 * m12 + m22 * m22));
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

__global__ void __launch_bounds__(256)
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
		double j[8];
		double hex_min = 1;
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			j[c] = a_jacobian(p[c_hex_tetra[c][0]], p[c_hex_tetra[c][1]], p[c_hex_tetra[c][2]], p[c_hex_tetra[c][3]]);
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
	const int grid = grid_for(ctx, H, blk, 8);
	DevBuf<Partial> partials(grid, s);
	DevBuf<double> tmpH;
	if (!H_Js) { tmpH.alloc(H, s); H_Js = tmpH.p; }
	jacobian_kernel<<<grid, blk, 0, s>>>(V, hex, H, V_Js, H_Js, partials.p);
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

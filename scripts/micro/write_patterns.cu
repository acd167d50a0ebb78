// Microbenchmark: achievable write bandwidth of the access patterns considered for voxel_fill_kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void linear16(uint4 *o, int64_t n16) { for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) o[i] = make_uint4(0, 0, 0, 0); }
__global__ void linear4(uint32_t *o, int64_t n4) { for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) o[i] = 0; }
// thread = 4 columns (uint32), loops nzc layers with stride = layer bytes
template <bool CS> __global__ void columns4(uint8_t *out, int nx, int ny, int nz, int zchunk) {
	const int gx = nx / 4; const int gz = nz / zchunk; const int64_t nth = (int64_t)gx * ny * gz; const int64_t layer = (int64_t)nx * ny;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nth; t += (int64_t)gridDim.x * blockDim.x) {
		const int x0 = (int)(t % gx) * 4, y = (int)((t / gx) % ny), z0 = (int)(t / ((int64_t)gx * ny)) * zchunk;
		uint8_t *o = out + (int64_t)z0 * layer + (int64_t)y * nx + x0;
		for (int z = 0; z < zchunk; ++z, o += layer) { if (CS) __stcs((uint32_t *)o, 0u); else *(uint32_t *)o = 0u; }
	}
}
// thread = 16 columns (uint4)
__global__ void columns16(uint8_t *out, int nx, int ny, int nz, int zchunk) {
	const int gx = nx / 16; const int gz = nz / zchunk; const int64_t nth = (int64_t)gx * ny * gz; const int64_t layer = (int64_t)nx * ny;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nth; t += (int64_t)gridDim.x * blockDim.x) {
		const int x0 = (int)(t % gx) * 16, y = (int)((t / gx) % ny), z0 = (int)(t / ((int64_t)gx * ny)) * zchunk;
		uint8_t *o = out + (int64_t)z0 * layer + (int64_t)y * nx + x0;
		for (int z = 0; z < zchunk; ++z, o += layer) *(uint4 *)o = make_uint4(0, 0, 0, 0);
	}
}
// TMA 1-D bulk stores: one elected thread per CTA streams its shared-memory buffer to consecutive chunks of global memory
// (cp.async.bulk.global.shared::cta, SASS UBLKCP).  `inflight` commit groups are kept open.
__global__ void bulk_store(uint8_t *out, int64_t bytes, int chunk, int inflight) {
	extern __shared__ __align__(128) uint8_t buf[];
	for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4 *>(buf)[i] = make_uint4(0, 0, 0, 0);
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint32_t sa = (uint32_t)__cvta_generic_to_shared(buf);
		const int64_t n_chunks = bytes / chunk;
		int open = 0;
		for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(sa), "r"(chunk) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			if (++open >= inflight) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); open = 0; }
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}
// every warp's lane 0 issues bulk stores of `chunk` bytes from a per-warp slice of shared memory (more issuers per SM)
__global__ void bulk_store_warps(uint8_t *out, int64_t bytes, int chunk, int inflight) {
	extern __shared__ __align__(128) uint8_t buf[];
	const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
	uint8_t *mine = buf + (int64_t)warp * chunk;
	for (int i = lane; i < chunk / 16; i += 32) reinterpret_cast<uint4 *>(mine)[i] = make_uint4(0, 0, 0, 0);
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncwarp();
	if (lane == 0) {
		const uint32_t sa = (uint32_t)__cvta_generic_to_shared(mine);
		const int64_t n_chunks = bytes / chunk;
		int open = 0;
		for (int64_t c = (int64_t)blockIdx.x * nw + warp; c < n_chunks; c += (int64_t)gridDim.x * nw) {
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(sa), "r"(chunk) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			if (++open >= inflight) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); open = 0; }
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}
// 256-bit stores (st.global.v8.b32, sm_100+)
__device__ __forceinline__ void st32(void *p, uint32_t v) { asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory"); }
__global__ void linear32(uint8_t *o, int64_t n32) { for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) st32(o + 32 * i, 0u); }
__global__ void columns32(uint8_t *out, int nx, int ny, int nz, int zchunk) {
	const int gx = nx / 32; const int gz = nz / zchunk; const int64_t nth = (int64_t)gx * ny * gz; const int64_t layer = (int64_t)nx * ny;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nth; t += (int64_t)gridDim.x * blockDim.x) {
		const int x0 = (int)(t % gx) * 32, y = (int)((t / gx) % ny), z0 = (int)(t / ((int64_t)gx * ny)) * zchunk;
		uint8_t *o = out + (int64_t)z0 * layer + (int64_t)y * nx + x0;
		for (int z = 0; z < zchunk; ++z, o += layer) st32(o, 0u);
	}
}
// rows: thread = 16 B of one (y, z) row, 8 rows (4 y x 2 z) per thread: the occupancy expansion's pattern
__global__ void rows16x8(uint8_t *out, int nx, int ny, int nz) {
	const int gx = nx / 16; const int64_t nth = (int64_t)gx * (ny / 4) * (nz / 2); const int64_t layer = (int64_t)nx * ny;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nth; t += (int64_t)gridDim.x * blockDim.x) {
		const int x0 = (int)(t % gx) * 16, ty = (int)((t / gx) % (ny / 4)), tz = (int)(t / ((int64_t)gx * (ny / 4)));
		for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 4; ++dy) __stcs((uint4 *)(out + (int64_t)(2 * tz + dz) * layer + (int64_t)(4 * ty + dy) * nx + x0), make_uint4(0, 0, 0, 0));
	}
}
template <class F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); f(); cudaDeviceSynchronize(); cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5; }
int main() {
	for (int n : {1024}) {
		const int64_t bytes = (int64_t)n * n * n; uint8_t *d; cudaMalloc(&d, bytes + 64);
		const int grid = 148 * 8;
		auto rep = [&](const char *name, float ms) { printf("n=%4d %-28s %8.3f ms %8.1f GB/s\n", n, name, ms, bytes / ms / 1e6); };
		rep("cudaMemset", timeit([&] { cudaMemsetAsync(d, 0, bytes); }));
		rep("linear16", timeit([&] { linear16<<<grid, 256>>>((uint4 *)d, bytes / 16); }));
		for (int chunk : {4096, 16384}) for (int fl : {4}) {
			char nm[64]; snprintf(nm, sizeof nm, "bulk_store %5d B x%d", chunk, fl);
			cudaFuncSetAttribute(bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
			rep(nm, timeit([&] { bulk_store<<<148 * 4, 128, chunk>>>(d, bytes, chunk, fl); }));
		}
		for (int chunk : {4096}) for (int fl : {4}) {
			char nm[64]; snprintf(nm, sizeof nm, "bulk_store_warps %5d B x%d", chunk, fl);
			cudaFuncSetAttribute(bulk_store_warps, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * chunk);
			rep(nm, timeit([&] { bulk_store_warps<<<148 * 4, 256, 8 * chunk>>>(d, bytes, chunk, fl); }));
		}
		rep("linear32 (v8.b32)", timeit([&] { linear32<<<grid, 256>>>(d, bytes / 32); }));
		rep("columns32 zchunk=32", timeit([&] { columns32<<<grid * 2, 256>>>(d, n, n, n, 32); }));
		rep("columns32 zchunk=8", timeit([&] { columns32<<<grid * 2, 256>>>(d, n, n, n, 8); }));
		rep("rows16x8 stcs", timeit([&] { rows16x8<<<grid * 2, 256>>>(d, n, n, n); }));
		rep("rows16x8 stcs grid x8", timeit([&] { rows16x8<<<grid * 8, 256>>>(d, n, n, n); }));
		rep("linear4", timeit([&] { linear4<<<grid, 256>>>((uint32_t *)d, bytes / 4); }));
		if (n % 16 == 0) {
			rep("columns4 zchunk=32", timeit([&] { columns4<false><<<grid * 2, 256>>>(d, n, n, n, 32); }));
			rep("columns4 zchunk=32 stcs", timeit([&] { columns4<true><<<grid * 2, 256>>>(d, n, n, n, 32); }));
			rep("columns4 zchunk=n", timeit([&] { columns4<false><<<grid * 2, 256>>>(d, n, n, n, n); }));
			rep("columns16 zchunk=32", timeit([&] { columns16<<<grid * 2, 256>>>(d, n, n, n, 32); }));
			rep("columns16 zchunk=8", timeit([&] { columns16<<<grid * 2, 256>>>(d, n, n, n, 8); }));
			rep("columns4 zchunk=8", timeit([&] { columns4<false><<<grid * 2, 256>>>(d, n, n, n, 8); }));
		} else {
			rep("columns4 zchunk=8 (n=1000)", timeit([&] { columns4<false><<<grid * 2, 256>>>(d, n, n, n, 8); }));
			rep("columns4 zchunk=40 (n=1000)", timeit([&] { columns4<false><<<grid * 2, 256>>>(d, n, n, n, 40); }));
		}
		cudaFree(d);
	}
	return 0;
}

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
template <class F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); f(); cudaDeviceSynchronize(); cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5; }
int main() {
	for (int n : {512, 1024, 1000}) {
		const int64_t bytes = (int64_t)n * n * n; uint8_t *d; cudaMalloc(&d, bytes + 64);
		const int grid = 148 * 8;
		auto rep = [&](const char *name, float ms) { printf("n=%4d %-28s %8.3f ms %8.1f GB/s\n", n, name, ms, bytes / ms / 1e6); };
		rep("cudaMemset", timeit([&] { cudaMemsetAsync(d, 0, bytes); }));
		rep("linear16", timeit([&] { linear16<<<grid, 256>>>((uint4 *)d, bytes / 16); }));
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

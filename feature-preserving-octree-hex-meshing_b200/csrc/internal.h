// Internal plumbing of libfpohm.so: error reporting, device buffers, context.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <new>
#include <chrono>

#include "../../include/fpohm.h"

namespace fpohm {

void set_error(const char *fmt, ...);

struct Failure { int code; };

#define FPOHM_CUDA(expr)                                                                     \
	do {                                                                                     \
		cudaError_t e__ = (expr);                                                            \
		if (e__ != cudaSuccess) {                                                            \
			::fpohm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
			throw ::fpohm::Failure{e__ == cudaErrorMemoryAllocation ? FPOHM_ENOMEM : FPOHM_ECUDA};   \
		}                                                                                    \
	} while (0)

#define FPOHM_REQUIRE(cond, code, ...)            \
	do {                                          \
		if (!(cond)) {                            \
			::fpohm::set_error(__VA_ARGS__);      \
			throw ::fpohm::Failure{code};         \
		}                                         \
	} while (0)

// every extern "C" body is wrapped so no exception crosses the C-ABI
#define FPOHM_API_BEGIN try {
#define FPOHM_API_END                                                        \
	return FPOHM_OK;                                                         \
	} catch (const ::fpohm::Failure &f) { return f.code; }                   \
	catch (const std::bad_alloc &) { ::fpohm::set_error("host out of memory"); return FPOHM_ENOMEM; } \
	catch (...) { ::fpohm::set_error("unexpected exception"); return FPOHM_ECUDA; }

// debug probe of the allocator (FPOHM_OCTREE_TIMELINE): host time spent allocating device memory
extern bool g_alloc_probe;
extern double g_alloc_ms;
extern long long g_alloc_calls;

// Device arena owned by a context.  Round 1 took every buffer from the device's default stream-ordered pool with the release
// threshold raised to "never": a process-wide side effect (memory invisible to the caller's own allocator), and a 1024^3-
// equivalent octree build spent 5 - 54 ms (once 227 ms) of its 40 - 150 ms inside cudaMallocAsync, whatever the pool had to
// re-map for the build's ~250 buffers of up to 1.8 GB.  The arena is a handful of big cudaMalloc chunks (doubling up to 4 GB)
// with a host-side best-fit free list; a block freed by a buffer is handed to the next request at once, which is safe for
// work ordered on ONE stream — so only buffers of the context's main stream live here, everything allocated on another
// stream (the caller's stream of the _dev entry points, the compute lanes of the host pipelines) stays with cudaMallocAsync.
// fpohm_ctx_trim returns unused chunks to the driver; fpohm_ctx_destroy returns everything.
struct Arena;
Arena *arena_for_stream(cudaStream_t s);                 // nullptr: not a context's main stream
void *arena_alloc(Arena *a, size_t bytes);               // throws Failure{FPOHM_ENOMEM}
void arena_free(Arena *a, void *p);
Arena *arena_create(int device, cudaStream_t main_stream);
void arena_destroy(Arena *a);
size_t arena_trim(Arena *a);                             // bytes returned to the driver
void arena_stats(Arena *a, size_t *reserved, size_t *in_use);
// Buffers on any OTHER stream (scratch of the _dev entry points on the caller's stream, lanes of the host pipelines) are
// stream-ordered allocations from a PRIVATE pool of the library, one per device, which keeps what is freed (the entry points
// are called in loops) — not from the device's default pool, whose configuration belongs to the process.
cudaMemPool_t side_pool();                               // pool of the current device; nullptr before the first context

// Plain device buffer: arena memory on a context's main stream, stream-ordered allocation elsewhere.
template <class T>
struct DevBuf {
	T *p = nullptr;
	int64_t n = 0;
	cudaStream_t s = nullptr;
	Arena *arena = nullptr;
	DevBuf() = default;
	DevBuf(int64_t count, cudaStream_t stream) { alloc(count, stream); }
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s), arena(o.arena) { o.p = nullptr; o.n = 0; }
	DevBuf &operator=(DevBuf &&o) noexcept {
		if (this != &o) { release(); p = o.p; n = o.n; s = o.s; arena = o.arena; o.p = nullptr; o.n = 0; }
		return *this;
	}
	~DevBuf() { release(); }
	void alloc(int64_t count, cudaStream_t stream) {
		release();
		n = count; s = stream;
		if (count > 0) {
			const auto t0 = g_alloc_probe ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
			arena = arena_for_stream(stream);
			if (arena) p = (T *)arena_alloc(arena, sizeof(T) * (size_t)count);
			else if (cudaMemPool_t pool = side_pool()) FPOHM_CUDA(cudaMallocFromPoolAsync((void **)&p, sizeof(T) * (size_t)count, pool, stream));
			else FPOHM_CUDA(cudaMallocAsync((void **)&p, sizeof(T) * (size_t)count, stream));
			if (g_alloc_probe) { g_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ++g_alloc_calls; }
		}
	}
	void release() {
		if (p) { if (arena) arena_free(arena, p); else cudaFreeAsync(p, s); }
		p = nullptr; n = 0; arena = nullptr;
	}
	void zero() { if (n) FPOHM_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, s)); }
	void upload(const T *h, int64_t count) {
		if (count) FPOHM_CUDA(cudaMemcpyAsync(p, h, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice, s));
	}
	void download(T *h, int64_t count) const {
		if (count) FPOHM_CUDA(cudaMemcpyAsync(h, p, sizeof(T) * (size_t)count, cudaMemcpyDeviceToHost, s));
	}
};

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

} // namespace fpohm

struct fpohm_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t aux[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // upload / download / extra compute lanes of the host-pointer entry points
	std::vector<cudaEvent_t> ev_pool;                     // per-chunk ordering events of those pipelines (grown on demand)
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr;
	int32_t *pinned_words = nullptr;              // 16 pinned host words a kernel posts control words to (voxel.cu), created on first use
	int32_t post_seq = 0;                         // sequence number of the last post
	double last_ms = 0;
	// CUDA-event ring around the dominant query kernel (packet walk) of the last resident/host query launches
	static constexpr int QRING = 32;
	cudaEvent_t q_ev0[QRING] = {}, q_ev1[QRING] = {};
	int64_t q_launches = 0;
	int64_t launches = 0;
	fpohm::Arena *arena = nullptr;
	// content-keyed cache of uploaded surfaces (fpohm_mesh_upload_cached): points_inside_mesh rebuilds its tree on EVERY call in the
	// reference (gf.cpp:4038) and the outer loop hands the same surface in again and again
	struct CachedMesh { uint64_t h0, h1; int64_t nV, nF; struct ::fpohm_mesh *mesh; int refs; uint64_t stamp; };
	std::vector<CachedMesh> mesh_cache;
	uint64_t cache_clock = 0;
};

namespace fpohm {
// grid for a grid-stride kernel: a multiple of the SM count (148 on B200), capped by the work
inline int grid_for(const fpohm_ctx *ctx, int64_t work_items, int block, int ctas_per_sm = 8) {
	int64_t need = (work_items + block - 1) / block;
	int64_t cap = (int64_t)ctx->sm_count * ctas_per_sm;
	if (need < 1) need = 1;
	return (int)(need < cap ? need : cap);
}
struct DeviceGuard {
	int prev = 0;
	explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
	~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct KernelTimer {
	fpohm_ctx *c; cudaStream_t s;
	KernelTimer(fpohm_ctx *ctx, cudaStream_t st) : c(ctx), s(st) { cudaEventRecord(c->ev0, s); }
	void stop() {
		cudaEventRecord(c->ev1, s);
		cudaEventSynchronize(c->ev1);
		float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
		c->last_ms = ms;
	}
};
#define FPOHM_LAUNCH_CHECK(ctx)                      \
	do {                                             \
		(ctx)->launches++;                           \
		FPOHM_CUDA(cudaGetLastError());              \
	} while (0)
} // namespace fpohm

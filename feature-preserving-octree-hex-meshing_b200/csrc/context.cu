// Context, error reporting, library identity.
#include "internal.h"
#include <algorithm>
#include <map>
#include <mutex>
#include <unordered_map>

namespace fpohm {
bool g_alloc_probe = getenv("FPOHM_OCTREE_TIMELINE") != nullptr;
double g_alloc_ms = 0;
long long g_alloc_calls = 0;

// ---- device arena (internal.h) ----------------------------------------------------------------------------------------
struct Arena {
	struct Chunk { char *base = nullptr; size_t size = 0; std::map<size_t, size_t> free_blocks; /* offset -> size */ };
	int device = 0;
	cudaStream_t stream = nullptr;
	std::vector<Chunk> chunks;
	std::unordered_map<void *, std::pair<int, size_t>> live;   // pointer -> (chunk, size)
	std::mutex mu;
	size_t reserved = 0, in_use = 0;
};
static std::mutex g_arena_mu;
static std::vector<Arena *> g_arenas;

Arena *arena_for_stream(cudaStream_t s) {
	std::lock_guard<std::mutex> l(g_arena_mu);
	for (Arena *a : g_arenas) if (a->stream == s) return a;
	return nullptr;
}
Arena *arena_create(int device, cudaStream_t main_stream) {
	Arena *a = new Arena;
	a->device = device; a->stream = main_stream;
	std::lock_guard<std::mutex> l(g_arena_mu);
	g_arenas.push_back(a);
	return a;
}
void arena_destroy(Arena *a) {
	if (!a) return;
	{
		std::lock_guard<std::mutex> l(g_arena_mu);
		g_arenas.erase(std::remove(g_arenas.begin(), g_arenas.end(), a), g_arenas.end());
	}
	for (auto &c : a->chunks) cudaFree(c.base);
	delete a;
}
void *arena_alloc(Arena *a, size_t bytes) {
	const size_t need = (bytes + 511) & ~(size_t)511;
	std::lock_guard<std::mutex> l(a->mu);
	for (int attempt = 0; attempt < 2; ++attempt) {
		int bc = -1; size_t bo = 0, bs = ~(size_t)0;
		for (int c = 0; c < (int)a->chunks.size(); ++c)
			for (auto &fb : a->chunks[(size_t)c].free_blocks)
				if (fb.second >= need && fb.second < bs) { bc = c; bo = fb.first; bs = fb.second; }
		if (bc >= 0) {
			Arena::Chunk &ch = a->chunks[(size_t)bc];
			ch.free_blocks.erase(bo);
			if (bs > need) ch.free_blocks[bo + need] = bs - need;
			void *p = ch.base + bo;
			a->live[p] = {bc, need};
			a->in_use += need;
			return p;
		}
		// grow: doubling chunks, 256 MB .. 4 GB, never smaller than the request
		size_t grow = std::min<size_t>(std::max<size_t>(a->reserved, (size_t)256 << 20), (size_t)4 << 30);
		grow = std::max(grow, (need + (((size_t)256 << 20) - 1)) & ~(((size_t)256 << 20) - 1));
		char *base = nullptr;
		cudaError_t e = cudaMalloc((void **)&base, grow);
		if (e != cudaSuccess && grow > need) { cudaGetLastError(); grow = need; e = cudaMalloc((void **)&base, grow); }
		if (e != cudaSuccess) {
			cudaGetLastError();
			set_error("device arena: cannot allocate %zu bytes (%zu reserved, %zu in use): %s", need, a->reserved, a->in_use, cudaGetErrorString(e));
			throw Failure{FPOHM_ENOMEM};
		}
		Arena::Chunk ch;
		ch.base = base; ch.size = grow; ch.free_blocks[0] = grow;
		a->chunks.push_back(std::move(ch));
		a->reserved += grow;
	}
	set_error("device arena: internal error");
	throw Failure{FPOHM_ECUDA};
}
void arena_free(Arena *a, void *p) {
	{   // a buffer that outlives its context (teardown order of a garbage-collected caller): the chunks are gone already
		std::lock_guard<std::mutex> g(g_arena_mu);
		if (std::find(g_arenas.begin(), g_arenas.end(), a) == g_arenas.end()) return;
	}
	std::lock_guard<std::mutex> l(a->mu);
	auto it = a->live.find(p);
	if (it == a->live.end()) return;
	const int c = it->second.first; size_t sz = it->second.second;
	a->live.erase(it);
	a->in_use -= sz;
	Arena::Chunk &ch = a->chunks[(size_t)c];
	size_t off = (size_t)((char *)p - ch.base);
	auto nx = ch.free_blocks.lower_bound(off);
	if (nx != ch.free_blocks.end() && off + sz == nx->first) { sz += nx->second; nx = ch.free_blocks.erase(nx); }
	if (nx != ch.free_blocks.begin()) {
		auto pv = std::prev(nx);
		if (pv->first + pv->second == off) { off = pv->first; sz += pv->second; ch.free_blocks.erase(pv); }
	}
	ch.free_blocks[off] = sz;
}
size_t arena_trim(Arena *a) {
	std::lock_guard<std::mutex> l(a->mu);
	size_t freed = 0;
	// a chunk can go only if nothing lives in it; chunk indices of live blocks must stay valid, so emptied chunks keep their slot
	for (auto &c : a->chunks) {
		if (c.base && c.free_blocks.size() == 1 && c.free_blocks.begin()->second == c.size) {
			cudaFree(c.base);
			freed += c.size; a->reserved -= c.size;
			c.base = nullptr; c.size = 0; c.free_blocks.clear();
		}
	}
	return freed;
}
void arena_stats(Arena *a, size_t *reserved, size_t *in_use) {
	std::lock_guard<std::mutex> l(a->mu);
	if (reserved) *reserved = a->reserved;
	if (in_use) *in_use = a->in_use;
}

static std::mutex g_pool_mu;
static cudaMemPool_t g_side_pool[64] = {};
cudaMemPool_t side_pool() {
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
	return g_side_pool[dev];
}
static void side_pool_ensure(int device) {
	std::lock_guard<std::mutex> l(g_pool_mu);
	if (device < 0 || device >= 64 || g_side_pool[device]) return;
	cudaMemPoolProps props = {};
	props.allocType = cudaMemAllocationTypePinned;
	props.handleTypes = cudaMemHandleTypeNone;
	props.location.type = cudaMemLocationTypeDevice;
	props.location.id = device;
	cudaMemPool_t pool = nullptr;
	if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return; }
	uint64_t thr = UINT64_MAX;
	cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
	g_side_pool[device] = pool;
}

static thread_local std::string g_err;
void set_error(const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
}
} // namespace fpohm

using namespace fpohm;

extern "C" {

const char *fpohm_last_error(void) { return g_err.c_str(); }
const char *fpohm_version(void) { return "fpohm-b200 0.1 (sm_100a, fp64, -fmad=false)"; }

int fpohm_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int fpohm_ctx_create(int device, fpohm_ctx **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(out, FPOHM_EINVAL, "fpohm_ctx_create: null out");
	int n = fpohm_device_count();
	FPOHM_REQUIRE(n > 0, FPOHM_ENODEV, "fpohm_ctx_create: no CUDA device visible — this library has no CPU fallback");
	FPOHM_REQUIRE(device >= 0 && device < n, FPOHM_EINVAL, "fpohm_ctx_create: device %d out of range [0,%d)", device, n);
	cudaDeviceProp prop;
	FPOHM_CUDA(cudaGetDeviceProperties(&prop, device));
	FPOHM_REQUIRE(prop.major == 10, FPOHM_ENODEV,
	              "fpohm_ctx_create: device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major, prop.minor);
	DeviceGuard g(device);
	fpohm_ctx *c = new fpohm_ctx;
	struct Guard { fpohm_ctx *c; ~Guard() { if (c) fpohm_ctx_destroy(c); } } guard{c};      // a failing call below must not leak the context
	c->device = device;
	c->sm_count = prop.multiProcessorCount;
	FPOHM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	FPOHM_CUDA(cudaStreamCreateWithFlags(&c->aux[0], cudaStreamNonBlocking));
	FPOHM_CUDA(cudaStreamCreateWithFlags(&c->aux[1], cudaStreamNonBlocking));
	for (int k = 2; k < 5; ++k) FPOHM_CUDA(cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking));
	FPOHM_CUDA(cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming));
	for (int k = 0; k < fpohm_ctx::QRING; ++k) { FPOHM_CUDA(cudaEventCreate(&c->q_ev0[k])); FPOHM_CUDA(cudaEventCreate(&c->q_ev1[k])); }
	FPOHM_CUDA(cudaEventCreate(&c->ev0));
	FPOHM_CUDA(cudaEventCreate(&c->ev1));
	// (the device's default stream-ordered pool is left as the process configured it: this context's buffers live in its own arena)
	c->arena = arena_create(device, c->stream);
	side_pool_ensure(device);
	guard.c = nullptr;
	*out = c;
	FPOHM_API_END
}

void fpohm_ctx_destroy(fpohm_ctx *ctx) {
	if (!ctx) return;
	DeviceGuard g(ctx->device);
	// (also the failure path of fpohm_ctx_create: members may still be null)
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	fpohm_ctx_mesh_cache_clear(ctx);
	for (int k = 0; k < fpohm_ctx::QRING; ++k) { if (ctx->q_ev0[k]) cudaEventDestroy(ctx->q_ev0[k]); if (ctx->q_ev1[k]) cudaEventDestroy(ctx->q_ev1[k]); }
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	for (int k = 0; k < 5; ++k) if (ctx->aux[k]) cudaStreamDestroy(ctx->aux[k]);
	for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
	if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
	if (ctx->pinned_words) cudaFreeHost(ctx->pinned_words);
	if (ctx->arena) arena_destroy(ctx->arena);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	cudaGetLastError();
	delete ctx;
}

int fpohm_ctx_trim(fpohm_ctx *ctx, int64_t *bytes_released) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx, FPOHM_EINVAL, "fpohm_ctx_trim: null ctx");
	DeviceGuard g(ctx->device);
	FPOHM_CUDA(cudaStreamSynchronize(ctx->stream));
	const size_t f = arena_trim(ctx->arena);
	if (cudaMemPool_t pool = side_pool()) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(pool, 0); }
	if (bytes_released) *bytes_released = (int64_t)f;
	FPOHM_API_END
}

int fpohm_ctx_memory(fpohm_ctx *ctx, int64_t *bytes_reserved, int64_t *bytes_in_use) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx, FPOHM_EINVAL, "fpohm_ctx_memory: null ctx");
	size_t r = 0, u = 0;
	arena_stats(ctx->arena, &r, &u);
	if (bytes_reserved) *bytes_reserved = (int64_t)r;
	if (bytes_in_use) *bytes_in_use = (int64_t)u;
	FPOHM_API_END
}

int fpohm_ctx_sync(fpohm_ctx *ctx) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx, FPOHM_EINVAL, "fpohm_ctx_sync: null ctx");
	DeviceGuard g(ctx->device);
	FPOHM_CUDA(cudaStreamSynchronize(ctx->stream));
	FPOHM_API_END
}

int fpohm_ctx_last_kernel_ms(fpohm_ctx *ctx, double *ms) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && ms, FPOHM_EINVAL, "fpohm_ctx_last_kernel_ms: null argument");
	*ms = ctx->last_ms;
	FPOHM_API_END
}

int fpohm_ctx_query_kernel_ms(fpohm_ctx *ctx, int32_t last_n, double *mean_ms) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mean_ms && last_n > 0, FPOHM_EINVAL, "fpohm_ctx_query_kernel_ms: bad argument");
	DeviceGuard g(ctx->device);
	const int64_t have = std::min<int64_t>(std::min<int64_t>(ctx->q_launches, fpohm_ctx::QRING), last_n);
	FPOHM_REQUIRE(have > 0, FPOHM_ESTATE, "fpohm_ctx_query_kernel_ms: no query launched on this context yet");
	double sum = 0;
	for (int64_t k = 0; k < have; ++k) {
		const int slot = (int)((ctx->q_launches - 1 - k) % fpohm_ctx::QRING);
		FPOHM_CUDA(cudaEventSynchronize(ctx->q_ev1[slot]));
		float ms = 0;
		FPOHM_CUDA(cudaEventElapsedTime(&ms, ctx->q_ev0[slot], ctx->q_ev1[slot]));
		sum += ms;
	}
	*mean_ms = sum / (double)have;
	FPOHM_API_END
}

int fpohm_ctx_launch_count(fpohm_ctx *ctx, int64_t *n) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n, FPOHM_EINVAL, "fpohm_ctx_launch_count: null argument");
	*n = ctx->launches;
	FPOHM_API_END
}

} // extern "C"

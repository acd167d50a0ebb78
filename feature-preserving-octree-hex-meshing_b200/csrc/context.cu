// Context, error reporting, library identity.
#include "internal.h"
#include <algorithm>

namespace fpohm {
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
	// keep freed blocks in the stream-ordered pool: the pipeline calls these entry points in loops
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		uint64_t thr = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
	}
	*out = c;
	FPOHM_API_END
}

void fpohm_ctx_destroy(fpohm_ctx *ctx) {
	if (!ctx) return;
	DeviceGuard g(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (int k = 0; k < fpohm_ctx::QRING; ++k) { cudaEventDestroy(ctx->q_ev0[k]); cudaEventDestroy(ctx->q_ev1[k]); }
	cudaEventDestroy(ctx->ev0);
	cudaEventDestroy(ctx->ev1);
	cudaStreamDestroy(ctx->aux[0]);
	cudaStreamDestroy(ctx->aux[1]);
	for (int k = 2; k < 5; ++k) cudaStreamDestroy(ctx->aux[k]);
	for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
	cudaEventDestroy(ctx->ev_sync);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
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

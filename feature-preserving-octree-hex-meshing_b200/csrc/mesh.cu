// fpohm_mesh: upload, lazily built acceleration structures.
#include "mesh.h"

#include <algorithm>
#include <chrono>
#include <thread>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>
#include <math_constants.h>

using namespace fpohm;

namespace {

__global__ void gather_triangles_kernel(const double *__restrict__ V, const int32_t *__restrict__ F, int64_t nF,
                                        double *__restrict__ tri)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const int64_t v = F[3 * f + k];
			tri[9 * f + 3 * k + 0] = V[3 * v + 0];
			tri[9 * f + 3 * k + 1] = V[3 * v + 1];
			tri[9 * f + 3 * k + 2] = V[3 * v + 2];
		}
	}
}

// ---- facet-bbox tree for the subdivision predicate --------------------------------------------------
// geogram sorts facets along a Morton curve and builds an implicit balanced tree (mesh_AABB.cpp:166-189,
// 325-348).  The predicate "does ANY facet box overlap" does not depend on facet order or tree shape
// (SURVEY.md Appendix B), so we are free to use our own order: 30-bit Morton code of the bbox centre.
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

__global__ void facet_morton_kernel(const double *__restrict__ tri, int64_t nF, double bx, double by, double bz,
                                    double sx, double sy, double sz, uint32_t *__restrict__ code, int32_t *__restrict__ idx)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const double *t = tri + 9 * f;
		const double cx = (t[0] + t[3] + t[6]) * (1.0 / 3.0), cy = (t[1] + t[4] + t[7]) * (1.0 / 3.0), cz = (t[2] + t[5] + t[8]) * (1.0 / 3.0);
		const uint32_t ix = (uint32_t)fmin(fmax((cx - bx) * sx, 0.0), 1023.0);
		const uint32_t iy = (uint32_t)fmin(fmax((cy - by) * sy, 0.0), 1023.0);
		const uint32_t iz = (uint32_t)fmin(fmax((cz - bz) * sz, 0.0), 1023.0);
		code[f] = (expand10(iz) << 2) | (expand10(iy) << 1) | expand10(ix);
		idx[f] = (int32_t)f;
	}
}

// leaves of the implicit tree: leaf j (heap index nLeafBase + j) holds the box of sorted facet j; the tree
// is a complete binary tree over P = next_pow2(nF) leaves, padding leaves are empty boxes (+inf, -inf).
__global__ void pred_leaf_kernel(const double *__restrict__ tri, const int32_t *__restrict__ order, int64_t nF, int64_t P,
                                 double *__restrict__ box)
{
	for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < P; j += (int64_t)gridDim.x * blockDim.x) {
		double *b = box + 6 * (P + j);
		if (j < nF) {
			const double *t = tri + 9 * (int64_t)order[j];
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				b[c] = fmin(t[c], fmin(t[3 + c], t[6 + c]));
				b[3 + c] = fmax(t[c], fmax(t[3 + c], t[6 + c]));
			}
		} else {
			b[0] = b[1] = b[2] = CUDART_INF; b[3] = b[4] = b[5] = -CUDART_INF; // empty: never overlaps, neutral for union
		}
	}
}

__global__ void pred_level_kernel(double *__restrict__ box, int64_t first, int64_t count) {
	for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < count; j += (int64_t)gridDim.x * blockDim.x) {
		const int64_t n = first + j;
		const double *l = box + 6 * (2 * n), *r = box + 6 * (2 * n + 1);
		double *b = box + 6 * n;
#pragma unroll
		for (int c = 0; c < 3; ++c) { b[c] = fmin(l[c], r[c]); b[3 + c] = fmax(l[3 + c], r[3 + c]); }
	}
}

} // namespace

namespace fpohm {

void mesh_ensure_pred(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s) {
	if (m->has_pred) return;
	const int64_t nF = m->nF;
	int64_t P = 1;
	while (P < nF) P <<= 1;
	DevBuf<uint32_t> code(nF, s), code2(nF, s);
	DevBuf<int32_t> idx(nF, s);
	DevBuf<int32_t> &order = m->pred_order;
	order.alloc(nF, s);
	const double ex = m->bbox[3] - m->bbox[0], ey = m->bbox[4] - m->bbox[1], ez = m->bbox[5] - m->bbox[2];
	const int blk = 256;
	facet_morton_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(m->tri.p, nF, m->bbox[0], m->bbox[1], m->bbox[2],
		ex > 0 ? 1024.0 / ex : 0.0, ey > 0 ? 1024.0 / ey : 0.0, ez > 0 ? 1024.0 / ez : 0.0, code.p, idx.p);
	FPOHM_LAUNCH_CHECK(ctx);
	size_t tmp_bytes = 0;
	FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, code.p, code2.p, idx.p, order.p, (int)nF, 0, 30, s));
	DevBuf<uint8_t> tmp((int64_t)tmp_bytes, s);
	FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, code.p, code2.p, idx.p, order.p, (int)nF, 0, 30, s));
	ctx->launches += 4;
	m->pred_box.alloc(6 * 2 * P, s);
	m->pred_nodes = 2 * P;
	pred_leaf_kernel<<<grid_for(ctx, P, blk), blk, 0, s>>>(m->tri.p, order.p, nF, P, m->pred_box.p);
	FPOHM_LAUNCH_CHECK(ctx);
	for (int64_t cnt = P / 2; cnt >= 1; cnt >>= 1) {
		pred_level_kernel<<<grid_for(ctx, cnt, blk), blk, 0, s>>>(m->pred_box.p, cnt, cnt);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	m->has_pred = true;
}

void mesh_ensure_tree(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s) {
	if (m->has_tree) return;
	// igl::AABB::init: on the device (tree_device.cu; FPOHM_TREE_HOST=1 keeps the host builder for A/B), host for tiny meshes
	static const bool host_tree = getenv("FPOHM_TREE_HOST") != nullptr;
	static const bool timeline = getenv("FPOHM_TREE_TIMELINE") != nullptr;      // debug: stage times on stderr
	auto now = []() { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
	const auto t0 = now();
	const bool on_host = host_tree || m->nF < 64;
	// independent of the tree: the three normal sets and the SHAPE of the 8-wide collapse (a function of nF alone).  In the device
	// build they run while the host threads sort the tied barycentre axes (the longest single item of a build).
	std::vector<WideKid> kids;
	int32_t n_wide = 0;
	double ms_normals = 0, ms_shape = 0;
	auto independent = [&]() {
		const auto a = now();
		// the shape enumeration is pure host work on its own thread, beside the normals (device kernels + host acos threads)
		std::thread shape_thread([&]() { const auto b0 = now(); wide_shape_host(m->nF, kids, n_wide); ms_shape = ms(b0, now()); });
		struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_shape{shape_thread};
		static const bool host_normals = getenv("FPOHM_NORMALS_HOST") != nullptr;      // A/B: the host builder of round 1
		if (host_normals) {
			build_igl_normals(m->hV.data(), m->nV, m->hF.data(), m->nF, m->hFN, m->hVN, m->hEN, m->hE, m->hEMAP);
			m->hnormals_valid = true; m->nE = (int64_t)m->hE.size() / 2;
			m->FN.alloc((int64_t)m->hFN.size(), s); m->FN.upload(m->hFN.data(), (int64_t)m->hFN.size());
			m->VN.alloc((int64_t)m->hVN.size(), s); m->VN.upload(m->hVN.data(), (int64_t)m->hVN.size());
			m->EN.alloc((int64_t)m->hEN.size(), s); m->EN.upload(m->hEN.data(), (int64_t)m->hEN.size());
			m->EMAP.alloc((int64_t)m->hEMAP.size(), s); m->EMAP.upload(m->hEMAP.data(), (int64_t)m->hEMAP.size());
			m->dE.alloc((int64_t)m->hE.size(), s); m->dE.upload(m->hE.data(), (int64_t)m->hE.size());
		} else {
			build_normals_device(ctx, m, s);
		}
		ms_normals = ms(a, now());
		shape_thread.join();
	};
	if (on_host) {
		build_igl_tree(m->hV.data(), m->nV, m->hF.data(), m->nF, m->htree);
		m->htree_valid = true;
		m->t_box.alloc((int64_t)m->htree.box.size(), s); m->t_box.upload(m->htree.box.data(), (int64_t)m->htree.box.size());
		m->t_prim.alloc((int64_t)m->htree.prim.size(), s); m->t_prim.upload(m->htree.prim.data(), (int64_t)m->htree.prim.size());
		independent();
	} else {
		build_igl_tree_device(ctx, m, s, m->tree_ties_host, independent);
	}
	const auto t2 = now();
	// QNode / QNodeF / prim_parent / (parent, depth) / 8-wide collapse / float triangles: on the device (tree_flatten.cu)
	flatten_tree_device(ctx, m, s, m->t_box.p, m->t_prim.p, kids, n_wide);
	m->qroot = m->nF == 1 ? ~0 : 0;      // a single facet: the root is the leaf ~0
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (timeline) fprintf(stderr, "[fpohm tree] %lld facets: igl tree + normals + wide shape %.1f ms (%s; host-sorted axes %d%d%d; normals %.1f ms and wide shape %.1f ms under the host sorts), flattening kernels + uploads %.1f ms\n",
	                      (long long)m->nF, ms(t0, t2), on_host ? "host" : "device", m->tree_ties_host[0], m->tree_ties_host[1], m->tree_ties_host[2],
	                      ms_normals, ms_shape, ms(t2, now()));
	m->has_tree = true;
}

// host copy of the DFS pre-order arrays for fpohm_mesh_tree_export: box and prim come down from the device, the child ids follow
// from the facet count (left = me + 1, right = me + 2 * ceil(n / 2))
void mesh_host_tree(fpohm_mesh *m) {
	if (m->htree_valid) return;
	const int64_t nF = m->nF, nn = 2 * nF - 1;
	HostTree &t = m->htree;
	t.box.assign(6 * (size_t)nn, 0.0); t.prim.assign((size_t)nn, -1); t.lr.assign(2 * (size_t)nn, -1);
	m->t_box.download(t.box.data(), 6 * nn);
	m->t_prim.download(t.prim.data(), nn);
	cudaStreamSynchronize(m->t_box.s);
	std::vector<std::pair<int32_t, int32_t>> st;
	st.push_back({0, (int32_t)nF});
	while (!st.empty()) {
		const auto [id, cnt] = st.back();
		st.pop_back();
		if (cnt <= 1) continue;
		const int32_t nl = (cnt + 1) / 2;
		t.lr[2 * (size_t)id] = id + 1; t.lr[2 * (size_t)id + 1] = id + 2 * nl;
		st.push_back({id + 1, nl}); st.push_back({id + 2 * nl, cnt - nl});
	}
	m->htree_valid = true;
}

} // namespace fpohm

extern "C" {

int fpohm_mesh_upload(fpohm_ctx *ctx, const double *V, int64_t nV, const int32_t *F, int64_t nF, fpohm_mesh **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && V && F && out, FPOHM_EINVAL, "fpohm_mesh_upload: null argument");
	FPOHM_REQUIRE(nV > 0 && nF > 0 && nF < (1ll << 30), FPOHM_EINVAL, "fpohm_mesh_upload: empty or oversized mesh (nV=%lld nF=%lld)", (long long)nV, (long long)nF);
	for (int64_t i = 0; i < 3 * nF; ++i)
		FPOHM_REQUIRE(F[i] >= 0 && F[i] < nV, FPOHM_EINVAL, "fpohm_mesh_upload: facet index %d out of range at %lld", F[i], (long long)i);
	DeviceGuard g(ctx->device);
	fpohm_mesh *m = new fpohm_mesh;
	try {
		m->ctx = ctx; m->nV = nV; m->nF = nF;
		m->hV.assign(V, V + 3 * nV);
		m->hF.assign(F, F + 3 * nF);
		for (int c = 0; c < 3; ++c) { m->bbox[c] = V[c]; m->bbox[3 + c] = V[c]; }
		for (int64_t i = 0; i < nV; ++i)
			for (int c = 0; c < 3; ++c) {
				m->bbox[c] = std::min(m->bbox[c], V[3 * i + c]);
				m->bbox[3 + c] = std::max(m->bbox[3 + c], V[3 * i + c]);
			}
		cudaStream_t s = ctx->stream;
		m->V.alloc(3 * nV, s); m->V.upload(V, 3 * nV);
		m->F.alloc(3 * nF, s); m->F.upload(F, 3 * nF);
		m->tri.alloc(9 * nF, s);
		gather_triangles_kernel<<<grid_for(ctx, nF, 256), 256, 0, s>>>(m->V.p, m->F.p, nF, m->tri.p);
		FPOHM_LAUNCH_CHECK(ctx);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	} catch (...) { delete m; throw; }
	*out = m;
	FPOHM_API_END
}

void fpohm_mesh_free(fpohm_mesh *mesh) {
	if (!mesh) return;
	if (mesh->cached) {      // a cached surface stays with its context; it goes when the cache evicts it or the context dies
		for (auto &c : mesh->ctx->mesh_cache) if (c.mesh == mesh && c.refs > 0) --c.refs;
		return;
	}
	DeviceGuard g(mesh->ctx->device);
	cudaStreamSynchronize(mesh->ctx->stream);
	delete mesh;
}

// 128 bits over the bytes of V and F (two multiply-xorshift lanes, 8 bytes at a time): a key, not a secret
static void content_hash(const double *V, int64_t nV, const int32_t *F, int64_t nF, uint64_t &h0, uint64_t &h1) {
	uint64_t a = 0x9e3779b97f4a7c15ull ^ (uint64_t)nV, b = 0xc2b2ae3d27d4eb4full ^ (uint64_t)nF;
	auto mix = [&](uint64_t x) {
		a = (a ^ x) * 0xff51afd7ed558ccdull; a ^= a >> 29;
		b = (b + x) * 0xc4ceb9fe1a85ec53ull; b ^= b >> 32;
	};
	const uint64_t *pv = reinterpret_cast<const uint64_t *>(V);
	for (int64_t i = 0; i < 3 * nV; ++i) mix(pv[i]);
	for (int64_t i = 0; i + 1 < 3 * nF; i += 2) mix((uint64_t)(uint32_t)F[i] | ((uint64_t)(uint32_t)F[i + 1] << 32));
	if ((3 * nF) & 1) mix((uint64_t)(uint32_t)F[3 * nF - 1]);
	h0 = a; h1 = b;
}

int fpohm_mesh_upload_cached(fpohm_ctx *ctx, const double *V, int64_t nV, const int32_t *F, int64_t nF, fpohm_mesh **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && V && F && out && nV > 0 && nF > 0, FPOHM_EINVAL, "fpohm_mesh_upload_cached: bad argument");
	uint64_t h0, h1;
	content_hash(V, nV, F, nF, h0, h1);
	for (auto &c : ctx->mesh_cache)
		if (c.h0 == h0 && c.h1 == h1 && c.nV == nV && c.nF == nF) { ++c.refs; c.stamp = ++ctx->cache_clock; *out = c.mesh; return FPOHM_OK; }
	fpohm_mesh *m = nullptr;
	const int rc = fpohm_mesh_upload(ctx, V, nV, F, nF, &m);
	if (rc != FPOHM_OK) return rc;
	m->cached = true;
	// at most 8 surfaces: evict the least recently used one nobody holds
	if (ctx->mesh_cache.size() >= 8) {
		int victim = -1;
		for (int i = 0; i < (int)ctx->mesh_cache.size(); ++i)
			if (ctx->mesh_cache[(size_t)i].refs == 0 && (victim < 0 || ctx->mesh_cache[(size_t)i].stamp < ctx->mesh_cache[(size_t)victim].stamp)) victim = i;
		if (victim >= 0) {
			fpohm_mesh *old = ctx->mesh_cache[(size_t)victim].mesh;
			old->cached = false;
			fpohm_mesh_free(old);
			ctx->mesh_cache.erase(ctx->mesh_cache.begin() + victim);
		}
	}
	ctx->mesh_cache.push_back({h0, h1, nV, nF, m, 1, ++ctx->cache_clock});
	*out = m;
	FPOHM_API_END
}

int fpohm_ctx_mesh_cache_clear(fpohm_ctx *ctx) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx, FPOHM_EINVAL, "fpohm_ctx_mesh_cache_clear: null ctx");
	for (auto &c : ctx->mesh_cache) { c.mesh->cached = false; fpohm_mesh_free(c.mesh); }
	ctx->mesh_cache.clear();
	FPOHM_API_END
}

int fpohm_mesh_build_query_tree(fpohm_ctx *ctx, fpohm_mesh *mesh) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh, FPOHM_EINVAL, "fpohm_mesh_build_query_tree: null argument");
	DeviceGuard g(ctx->device);
	mesh_ensure_tree(ctx, mesh, ctx->stream);
	FPOHM_API_END
}

int fpohm_mesh_num_edges(const fpohm_mesh *mesh, int64_t *n_edges) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(mesh && n_edges, FPOHM_EINVAL, "fpohm_mesh_num_edges: null argument");
	FPOHM_REQUIRE(mesh->has_tree, FPOHM_ESTATE, "fpohm_mesh_num_edges: call fpohm_mesh_build_query_tree first");
	*n_edges = mesh->nE;
	FPOHM_API_END
}

int fpohm_mesh_normals(const fpohm_mesh *mesh, double *FN, double *VN, double *EN, int32_t *E, int32_t *EMAP) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(mesh, FPOHM_EINVAL, "fpohm_mesh_normals: null argument");
	FPOHM_REQUIRE(mesh->has_tree, FPOHM_ESTATE, "fpohm_mesh_normals: call fpohm_mesh_build_query_tree first");
	fpohm::mesh_host_normals(const_cast<fpohm_mesh *>(mesh));
	if (FN) std::copy(mesh->hFN.begin(), mesh->hFN.end(), FN);
	if (VN) std::copy(mesh->hVN.begin(), mesh->hVN.end(), VN);
	if (EN) std::copy(mesh->hEN.begin(), mesh->hEN.end(), EN);
	if (E) std::copy(mesh->hE.begin(), mesh->hE.end(), E);
	if (EMAP) std::copy(mesh->hEMAP.begin(), mesh->hEMAP.end(), EMAP);
	FPOHM_API_END
}

int fpohm_mesh_tree_nodes(const fpohm_mesh *mesh, int64_t *n_nodes) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(mesh && n_nodes, FPOHM_EINVAL, "fpohm_mesh_tree_nodes: null argument");
	FPOHM_REQUIRE(mesh->has_tree, FPOHM_ESTATE, "fpohm_mesh_tree_nodes: call fpohm_mesh_build_query_tree first");
	*n_nodes = 2 * mesh->nF - 1;
	FPOHM_API_END
}

int fpohm_mesh_tree_export(const fpohm_mesh *mesh, double *box, int32_t *prim, int32_t *lr) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(mesh, FPOHM_EINVAL, "fpohm_mesh_tree_export: null argument");
	FPOHM_REQUIRE(mesh->has_tree, FPOHM_ESTATE, "fpohm_mesh_tree_export: call fpohm_mesh_build_query_tree first");
	fpohm::mesh_host_tree(const_cast<fpohm_mesh *>(mesh));
	const fpohm::HostTree &t = mesh->htree;
	if (box) std::copy(t.box.begin(), t.box.end(), box);
	if (prim) std::copy(t.prim.begin(), t.prim.end(), prim);
	if (lr) std::copy(t.lr.begin(), t.lr.end(), lr);
	FPOHM_API_END
}

// Host-only restatement entry points (no device needed): the tree / normals builders are pure host code and
// this is how the CPU test-suite pins them against the reference.
int fpohm_host_igl_tree(const double *V, int64_t nV, const int32_t *F, int64_t nF, int64_t *n_nodes,
                        double *box, int32_t *prim, int32_t *lr)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(V && F && n_nodes && nV > 0 && nF > 0, FPOHM_EINVAL, "fpohm_host_igl_tree: bad argument");
	fpohm::HostTree t;
	fpohm::build_igl_tree(V, nV, F, nF, t);
	const int64_t cap = *n_nodes;
	*n_nodes = (int64_t)t.prim.size();
	if (box || prim || lr) FPOHM_REQUIRE(cap >= *n_nodes, FPOHM_EINVAL, "fpohm_host_igl_tree: capacity %lld < %lld nodes", (long long)cap, (long long)*n_nodes);
	if (box) std::copy(t.box.begin(), t.box.end(), box);
	if (prim) std::copy(t.prim.begin(), t.prim.end(), prim);
	if (lr) std::copy(t.lr.begin(), t.lr.end(), lr);
	FPOHM_API_END
}

int fpohm_host_igl_normals(const double *V, int64_t nV, const int32_t *F, int64_t nF, int64_t *n_edges,
                           double *FN, double *VN, double *EN, int32_t *E, int32_t *EMAP)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(V && F && n_edges && nV > 0 && nF > 0, FPOHM_EINVAL, "fpohm_host_igl_normals: bad argument");
	std::vector<double> fn, vn, en; std::vector<int32_t> e, emap;
	fpohm::build_igl_normals(V, nV, F, nF, fn, vn, en, e, emap);
	const int64_t cap = *n_edges;
	*n_edges = (int64_t)e.size() / 2;
	if (EN || E) FPOHM_REQUIRE(cap >= *n_edges, FPOHM_EINVAL, "fpohm_host_igl_normals: capacity %lld < %lld edges", (long long)cap, (long long)*n_edges);
	if (FN) std::copy(fn.begin(), fn.end(), FN);
	if (VN) std::copy(vn.begin(), vn.end(), VN);
	if (EN) std::copy(en.begin(), en.end(), EN);
	if (E) std::copy(e.begin(), e.end(), E);
	if (EMAP) std::copy(emap.begin(), emap.end(), EMAP);
	FPOHM_API_END
}

} // extern "C"

// fpohm_mesh: upload, lazily built acceleration structures.
#include "mesh.h"

#include <algorithm>
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

__global__ void float_triangles_kernel(const double *__restrict__ tri, int64_t nF, float4 *__restrict__ out) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 3 * nF; t += (int64_t)gridDim.x * blockDim.x) {
		const double *v = tri + 3 * t;
		out[t] = make_float4((float)v[0], (float)v[1], (float)v[2], 0.f);
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
	build_igl_tree(m->hV.data(), m->nV, m->hF.data(), m->nF, m->htree);
	build_igl_normals(m->hV.data(), m->nV, m->hF.data(), m->nF, m->hFN, m->hVN, m->hEN, m->hE, m->hEMAP);
	const HostTree &t = m->htree;
	const size_t nn = t.prim.size();
	std::vector<int32_t> internal_id(nn, -1);
	int32_t ni = 0;
	for (size_t i = 0; i < nn; ++i) if (t.prim[i] < 0) internal_id[i] = ni++;
	std::vector<QNode> q((size_t)ni);
	auto child_ref = [&](int32_t c) -> int32_t { return t.prim[c] >= 0 ? ~t.prim[c] : internal_id[c]; };
	for (size_t i = 0; i < nn; ++i) {
		if (t.prim[i] >= 0) continue;
		QNode &n = q[(size_t)internal_id[i]];
		const int32_t l = t.lr[2 * i], r = t.lr[2 * i + 1];
		for (int c = 0; c < 3; ++c) {
			n.lmin[c] = t.box[6 * (size_t)l + c]; n.lmax[c] = t.box[6 * (size_t)l + 3 + c];
			n.rmin[c] = t.box[6 * (size_t)r + c]; n.rmax[c] = t.box[6 * (size_t)r + 3 + c];
		}
		n.left = child_ref(l); n.right = child_ref(r);
		n.pad[0] = n.pad[1] = 0;
	}
	// parents / depths (pre-order: a parent precedes its children)
	int32_t maxd = 0;
	if (ni) { q[0].parent = -1; q[0].depth = 0; }
	for (int32_t i = 0; i < ni; ++i)
		for (int32_t c : {q[(size_t)i].left, q[(size_t)i].right})
			if (c >= 0) { q[(size_t)c].parent = i; q[(size_t)c].depth = q[(size_t)i].depth + 1; maxd = std::max(maxd, q[(size_t)c].depth); }
	m->qdepth = maxd;
	std::vector<int32_t> pp((size_t)std::max<int64_t>(m->nF, 1), -1);
	for (int32_t i = 0; i < ni; ++i) {
		if (q[(size_t)i].left < 0) pp[(size_t)~q[(size_t)i].left] = i;
		if (q[(size_t)i].right < 0) pp[(size_t)~q[(size_t)i].right] = i;
	}
	m->prim_parent.alloc((int64_t)pp.size(), s);
	m->prim_parent.upload(pp.data(), (int64_t)pp.size());
	std::vector<int2> pd((size_t)std::max<int32_t>(ni, 1));
	for (int32_t i = 0; i < ni; ++i) pd[(size_t)i] = make_int2(q[(size_t)i].parent, q[(size_t)i].depth);
	m->node_pd.alloc((int64_t)pd.size(), s);
	m->node_pd.upload(pd.data(), (int64_t)pd.size());
	FPOHM_CUDA(cudaStreamSynchronize(s));      // pd is a local
	// fp32 filter copy, boxes rounded outwards
	auto f_dn = [](double x) { float f = (float)x; if ((double)f > x) f = std::nextafterf(f, -INFINITY); return f; };
	auto f_up = [](double x) { float f = (float)x; if ((double)f < x) f = std::nextafterf(f, INFINITY); return f; };
	std::vector<QNodeF> qf((size_t)std::max<int32_t>(ni, 1));
	for (int32_t i = 0; i < ni; ++i) {
		const QNode &n = q[(size_t)i];
		QNodeF &f = qf[(size_t)i];
		for (int c = 0; c < 3; ++c) {
			f.lmin[c] = f_dn(n.lmin[c]); f.lmax[c] = f_up(n.lmax[c]);
			f.rmin[c] = f_dn(n.rmin[c]); f.rmax[c] = f_up(n.rmax[c]);
		}
		f.left = n.left; f.right = n.right; f.pad[0] = f.pad[1] = 0;
	}
	m->qfnodes.alloc((int64_t)qf.size(), s);
	m->qfnodes.upload(qf.data(), (int64_t)qf.size());
	// 8-wide collapse for the box-parallel packet search: starting from a binary node's two children, the child with the
	// most facets is opened until there are eight (igl's median splits are balanced by count, so this regroups three
	// binary levels; a subtree of <= 8 facets becomes one all-facet node, a "cluster").
	if (ni > 0) {
		std::vector<int32_t> leaves(nn, 1);
		for (size_t i = nn; i-- > 0;) if (t.prim[i] < 0) leaves[i] = leaves[(size_t)t.lr[2 * i]] + leaves[(size_t)t.lr[2 * i + 1]];
		std::vector<WNode> w;
		w.reserve((size_t)ni / 3 + 8);
		std::vector<std::pair<int32_t, int32_t>> work;   // (binary node, wide node index)
		w.emplace_back();
		work.push_back({0, 0});
		while (!work.empty()) {
			const auto [b, wi] = work.back();
			work.pop_back();
			int32_t kids[8];
			int nk = 2;
			kids[0] = t.lr[2 * (size_t)b]; kids[1] = t.lr[2 * (size_t)b + 1];
			while (nk < 8) {
				int best = -1;
				for (int k = 0; k < nk; ++k)
					if (t.prim[(size_t)kids[k]] < 0 && (best < 0 || leaves[(size_t)kids[k]] > leaves[(size_t)kids[best]])) best = k;
				if (best < 0) break;
				const int32_t o = kids[best];
				for (int k = nk; k > best + 1; --k) kids[k] = kids[k - 1];      // keep the binary tree's left-to-right order
				kids[best] = t.lr[2 * (size_t)o]; kids[best + 1] = t.lr[2 * (size_t)o + 1];
				++nk;
			}
			for (int k = 0; k < 8; ++k) {
				WChild e;
				if (k < nk) {
					const size_t c = (size_t)kids[k];
					for (int a = 0; a < 3; ++a) { e.lo[a] = f_dn(t.box[6 * c + a]); e.hi[a] = f_up(t.box[6 * c + 3 + a]); }
					if (t.prim[c] >= 0) { e.child = ~t.prim[c]; e.flags = 0; }
					else {
						e.child = (int32_t)w.size();
						e.flags = leaves[c] <= 8 ? 1 : 0;
						w.emplace_back();
						work.push_back({kids[k], e.child});
					}
				} else {
					for (int a = 0; a < 3; ++a) { e.lo[a] = INFINITY; e.hi[a] = -INFINITY; }
					e.child = WCHILD_EMPTY; e.flags = 0;
				}
				w[(size_t)wi].c[k] = e;
			}
		}
		m->n_wnodes = (int64_t)w.size();
		m->wnodes.alloc(m->n_wnodes, s);
		m->wnodes.upload(w.data(), m->n_wnodes);
		m->trif.alloc(3 * m->nF, s);
		float_triangles_kernel<<<grid_for(ctx, 3 * m->nF, 256), 256, 0, s>>>(m->tri.p, m->nF, m->trif.p);
		FPOHM_LAUNCH_CHECK(ctx);
		double ev = 0, mc = 0;
		for (int64_t v = 0; v < m->nV; ++v) {
			double e2 = 0;
			for (int a = 0; a < 3; ++a) {
				const double x = m->hV[(size_t)(3 * v + a)], e = x - (double)(float)x;
				e2 += e * e; mc = std::max(mc, std::fabs(x));
			}
			ev = std::max(ev, e2);
		}
		m->eps_v = f_up(std::sqrt(ev) * 1.000001);
		m->slack_q = f_up(36.0 * 5.9604644775390625e-8 * mc * 1.01);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	}
	m->n_qnodes = ni;
	m->qroot = nn ? (t.prim[0] >= 0 ? ~t.prim[0] : 0) : 0;
	m->qnodes.alloc(std::max<int64_t>(ni, 1), s);
	m->qnodes.upload(q.data(), ni);
	m->FN.alloc((int64_t)m->hFN.size(), s); m->FN.upload(m->hFN.data(), (int64_t)m->hFN.size());
	m->VN.alloc((int64_t)m->hVN.size(), s); m->VN.upload(m->hVN.data(), (int64_t)m->hVN.size());
	m->EN.alloc((int64_t)m->hEN.size(), s); m->EN.upload(m->hEN.data(), (int64_t)m->hEN.size());
	m->EMAP.alloc((int64_t)m->hEMAP.size(), s); m->EMAP.upload(m->hEMAP.data(), (int64_t)m->hEMAP.size());
	FPOHM_CUDA(cudaStreamSynchronize(s)); // q is a local: the upload must finish before it dies
	m->has_tree = true;
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
	DeviceGuard g(mesh->ctx->device);
	cudaStreamSynchronize(mesh->ctx->stream);
	delete mesh;
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
	*n_edges = (int64_t)mesh->hE.size() / 2;
	FPOHM_API_END
}

int fpohm_mesh_normals(const fpohm_mesh *mesh, double *FN, double *VN, double *EN, int32_t *E, int32_t *EMAP) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(mesh, FPOHM_EINVAL, "fpohm_mesh_normals: null argument");
	FPOHM_REQUIRE(mesh->has_tree, FPOHM_ESTATE, "fpohm_mesh_normals: call fpohm_mesh_build_query_tree first");
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
	*n_nodes = (int64_t)mesh->htree.prim.size();
	FPOHM_API_END
}

int fpohm_mesh_tree_export(const fpohm_mesh *mesh, double *box, int32_t *prim, int32_t *lr) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(mesh, FPOHM_EINVAL, "fpohm_mesh_tree_export: null argument");
	FPOHM_REQUIRE(mesh->has_tree, FPOHM_ESTATE, "fpohm_mesh_tree_export: call fpohm_mesh_build_query_tree first");
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

// Batched closest point + pseudonormal signed distance over the flattened igl-identical tree.
//
// Replaces igl::signed_distance_pseudonormal (igl/signed_distance.cpp:186-257):
//   AABB::squared_distance   igl/AABB.cpp:352-441     -> traverse()
//   leaf_squared_distance    igl/AABB.cpp:735-750
//   set_min (strict '<')     igl/AABB.cpp:754-779
//   point_simplex_squared_distance (Ericson) igl/point_simplex_squared_distance.cpp:44-124 -> closest_on_triangle()
//   pseudonormal_test        igl/pseudonormal_test.cpp:14-119 -> pseudonormal()
//
// One thread per query; the visiting order inside a query is exactly igl's (it decides which facet wins a
// distance tie), parallelism is across queries only.  Every fp64 expression keeps igl/Eigen's association
// (3-term sums are a0 + (a1 + a2), Eigen 3.2 unrolled redux) and the library is built with -fmad=false, so
// the device evaluates the same IEEE operations as the reference's SSE2 build.
#include "mesh.h"
#include <math_constants.h>
#include <cstdlib>
#include <algorithm>

using namespace fpohm;

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 sub(const V3 &a, const V3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot3(const V3 &a, const V3 &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ double sqnorm(const V3 &a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 ld3(const double *p) { return {p[0], p[1], p[2]}; }

// igl/point_simplex_squared_distance.cpp:44-109
__device__ __forceinline__ V3 closest_on_triangle(const V3 &p, const V3 &a, const V3 &b, const V3 &c) {
	const V3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
	const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
	if (d1 <= 0.0 && d2 <= 0.0) return a;
	const V3 bp = sub(p, b);
	const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
	if (d3 >= 0.0 && d4 <= d3) return b;
	const double vc = d1 * d4 - d3 * d2;
	if (a.x != b.x || a.y != b.y || a.z != b.z) {
		if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
			const double v = d1 / (d1 - d3);
			return {a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z};
		}
	}
	const V3 cp = sub(p, c);
	const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
	if (d6 >= 0.0 && d5 <= d6) return c;
	const double vb = d5 * d2 - d1 * d6;
	if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
		const double w = d2 / (d2 - d6);
		return {a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z};
	}
	const double va = d3 * d6 - d5 * d4;
	if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
		const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
		return {b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z)};
	}
	const double denom = 1.0 / (va + vb + vc);
	const double v = vb * denom, w = vc * denom;
	return {(a.x + ab.x * v) + ac.x * w, (a.y + ab.y * v) + ac.y * w, (a.z + ab.z * v) + ac.z * w};
}

// Eigen AlignedBox::squaredExteriorDistance: sequential accumulation over the three axes
__device__ __forceinline__ double box_ext_sqdist(const double *mn, const double *mx, const V3 &p) {
	double d2 = 0.0, aux;
	if (mn[0] > p.x) { aux = mn[0] - p.x; d2 += aux * aux; } else if (p.x > mx[0]) { aux = p.x - mx[0]; d2 += aux * aux; }
	if (mn[1] > p.y) { aux = mn[1] - p.y; d2 += aux * aux; } else if (p.y > mx[1]) { aux = p.y - mx[1]; d2 += aux * aux; }
	if (mn[2] > p.z) { aux = mn[2] - p.z; d2 += aux * aux; } else if (p.z > mx[2]) { aux = p.z - mx[2]; d2 += aux * aux; }
	return d2;
}
__device__ __forceinline__ bool box_contains(const double *mn, const double *mx, const V3 &p) {
	return mn[0] <= p.x && mn[1] <= p.y && mn[2] <= p.z && p.x <= mx[0] && p.y <= mx[1] && p.z <= mx[2];
}

struct Hit { double sqr_d; int32_t f; V3 c; int nodes, leaves; };

__device__ __forceinline__ void test_leaf(const double *__restrict__ tri, int32_t prim, const V3 &p, Hit &h) {
	const double *t = tri + 9 * (int64_t)prim;
	const V3 c = closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6));
	const double d = sqnorm(sub(p, c));
	if (d < h.sqr_d) { h.sqr_d = d; h.f = prim; h.c = c; }
	++h.leaves;
}

// igl/AABB.cpp:364-441 with the recursion unrolled onto an explicit stack.  Per node igl does:
//   look at children whose box CONTAINS p (left first), then the remaining ones by ascending exterior
//   distance, each only if that distance is < the best so far AT THAT MOMENT.
// "contains" <=> exterior distance 0, so the order is: first = left if (p in left box or dl < dr) else right,
// and both visits are gated by (d < best) evaluated when the visit is due (the deferred child is re-tested
// when it is popped).  A contained child visited with best == 0 cannot change the result (strict '<').
// NOTE (round 1, measured): a "while-while" restructuring (inner loop over internal nodes only, leaf tests batched per
// warp) ran 2.6x SLOWER on B200 (37.3 ms vs 14.2 ms for 4.6 M queries) — igl's order visits few leaves per query and the
// forced reconvergence serialises the short box steps; an explicit warp-voted variant (leaf tests batched once >= 20 lanes
// hold one) took 25.0 ms, and an order-free search seeded with the previous query's facet (+ an igl visiting-order
// comparator at ties, bit-identical results) did not reduce the work (65 -> 61 node visits, 14.5 -> 12.8 leaf tests per query:
// intrinsic to the tree and geometry).  The plain form below is kept.  profiles/r01_ncu_summary.md.
#define FPOHM_STACK 48
__device__ __forceinline__ void traverse(const QNode *__restrict__ nodes, int32_t root, const double *__restrict__ tri,
                                         const V3 &p, Hit &h)
{
	h.sqr_d = CUDART_INF; h.f = -1; h.c = {0, 0, 0}; h.nodes = 0; h.leaves = 0;
	if (root < 0) { test_leaf(tri, ~root, p, h); return; }
	int32_t st_node[FPOHM_STACK];
	double st_d[FPOHM_STACK];
	int sp = 0;
	int32_t cur = root;
	for (;;) {
		const QNode *n = nodes + cur;
		++h.nodes;
		const double dl = box_ext_sqdist(n->lmin, n->lmax, p);
		const double dr = box_ext_sqdist(n->rmin, n->rmax, p);
		const bool in_l = box_contains(n->lmin, n->lmax, p);
		const bool left_first = in_l || dl < dr;
		const int32_t c1 = left_first ? n->left : n->right, c2 = left_first ? n->right : n->left;
		const double d1 = left_first ? dl : dr, d2 = left_first ? dr : dl;
		if (d2 < h.sqr_d && sp < FPOHM_STACK) { st_node[sp] = c2; st_d[sp] = d2; ++sp; }
		int32_t next = -1;
		bool have_next = false;
		if (d1 < h.sqr_d) {
			if (c1 < 0) test_leaf(tri, ~c1, p, h); else { next = c1; have_next = true; }
		}
		while (!have_next && sp > 0) {
			--sp;
			if (st_d[sp] < h.sqr_d) {
				const int32_t c = st_node[sp];
				if (c < 0) test_leaf(tri, ~c, p, h); else { next = c; have_next = true; }
			}
		}
		if (!have_next) break;
		cur = next;
	}
}

// igl/pseudonormal_test.cpp:14-119
__device__ __forceinline__ double pseudonormal(const double *__restrict__ V, const int32_t *__restrict__ F, int64_t nF,
                                               const double *__restrict__ FN, const double *__restrict__ VN,
                                               const double *__restrict__ EN, const int32_t *__restrict__ EMAP,
                                               const V3 &q, int32_t f, const V3 &c, V3 &n)
{
	const int32_t fv[3] = {F[3 * (int64_t)f], F[3 * (int64_t)f + 1], F[3 * (int64_t)f + 2]};
	const V3 A = ld3(V + 3 * (int64_t)fv[0]), B = ld3(V + 3 * (int64_t)fv[1]), C = ld3(V + 3 * (int64_t)fv[2]);
	// doublearea(A,B,C): igl/doublearea.cpp:101-109,142-166 (Kahan's Heron on edge lengths sorted descending)
	double l0 = sqrt(sqnorm(sub(B, C))), l1 = sqrt(sqnorm(sub(C, A))), l2 = sqrt(sqnorm(sub(A, B)));
	{ // igl::sort3 descending, igl/sort.cpp:249-266
		double t;
		if (l0 < l1) { t = l0; l0 = l1; l1 = t; }
		if (l1 < l2) { t = l1; l1 = l2; l2 = t; if (l0 < l1) { t = l0; l0 = l1; l1 = t; } }
	}
	const double arg = (l0 + (l1 + l2)) * (l2 - (l0 - l1)) * (l2 + (l0 - l1)) * (l0 + (l1 - l2));
	const double area = 2.0 * 0.25 * sqrt(arg);
	const double MIN_DOUBLE_AREA = 1e-4, epsilon = 1e-12;
	const V3 *P[3] = {&A, &B, &C};
	bool set = false;
	if (area > MIN_DOUBLE_AREA) {
		// barycentric_coordinates, igl/barycentric_coordinates.cpp:89-101
		const V3 v0 = sub(B, A), v1 = sub(C, A), v2 = sub(c, A);
		const double d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
		const double denom = d00 * d11 - d01 * d01;
		double b[3];
		b[1] = (d11 * d20 - d01 * d21) / denom;
		b[2] = (d00 * d21 - d01 * d20) / denom;
		b[0] = 1.0 - (b[1] + b[2]);
		const int type = (b[0] <= epsilon) + (b[1] <= epsilon) + (b[2] <= epsilon);
		if (type == 2) {
			for (int x = 0; x < 3; ++x) if (b[x] > epsilon) { n = ld3(VN + 3 * (int64_t)fv[x]); set = true; break; }
		} else if (type == 1) {
			for (int x = 0; x < 3; ++x) if (b[x] <= epsilon) { n = ld3(EN + 3 * (int64_t)EMAP[nF * x + f]); set = true; break; }
		} else {
			n = ld3(FN + 3 * (int64_t)f); set = true;
		}
		// type == 2 with no b(x) > epsilon cannot happen (exactly one is > epsilon); keep igl's "n untouched" otherwise
	} else {
		for (int v = 0; v < 3 && !set; ++v) {
			if (sqrt(sqnorm(sub(c, *P[v]))) < epsilon) { set = true; n = ld3(VN + 3 * (int64_t)fv[v]); }
		}
		for (int e = 0; e < 3 && !set; ++e) {
			// project_to_line_segment(c, s, d): igl/project_to_line.cpp:36-56, igl/project_to_line_segment.cpp:25-41
			const V3 &s = *P[(e + 1) % 3], &d = *P[(e + 2) % 3];
			const V3 dms = sub(d, s);
			const double v_sqrlen = sqnorm(dms);
			const V3 smp = sub(s, c);
			double t = -(dms.x * smp.x + (dms.y * smp.y + dms.z * smp.z)) / v_sqrlen;
			const V3 proj = {(1 - t) * s.x + t * d.x, (1 - t) * s.y + t * d.y, (1 - t) * s.z + t * d.z};
			double sq = sqnorm(sub(c, proj));
			if (t < 0) sq = sqnorm(sub(c, s)); else if (t > 1) sq = sqnorm(sub(c, d));
			if (sqrt(sq) < epsilon) { n = ld3(EN + 3 * (int64_t)EMAP[nF * e + f]); set = true; }
		}
		if (!set) { n = ld3(FN + 3 * (int64_t)f); set = true; }
	}
	const V3 qc = sub(q, c);
	return dot3(qc, n) >= 0 ? 1. : -1.;
}

// Pass 1: traversal only.  Writes facet, closest point and SQUARED distance (into S).  Keeping the pseudonormal code out
// of this kernel keeps it at <= 64 registers (8 CTAs of 128 threads per SM) and removes a long divergent tail per query.
template <bool STATS>
__global__ void __launch_bounds__(128, 6)
closest_point_kernel(const QNode *__restrict__ nodes, int32_t root, const double *__restrict__ tri,
                     const double *__restrict__ P, int64_t np,
                     double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C, double *__restrict__ N)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
		const V3 p = ld3(P + 3 * i);
		Hit h;
		traverse(nodes, root, tri, p, h);
		if (I) I[i] = h.f;
		if (C) { C[3 * i] = h.c.x; C[3 * i + 1] = h.c.y; C[3 * i + 2] = h.c.z; }
		if (S) S[i] = h.sqr_d;
		if (STATS && N) { N[3 * i] = h.nodes; N[3 * i + 1] = h.leaves; }   // debug: traversal counters
	}
}

// Pass 2 (signed queries only): pseudonormal_test on (q, facet, closest point); fully coherent, one thread per query.
__global__ void __launch_bounds__(128)
pseudonormal_kernel(const double *__restrict__ V, const int32_t *__restrict__ F, int64_t nF,
                    const double *__restrict__ FN, const double *__restrict__ VN, const double *__restrict__ EN,
                    const int32_t *__restrict__ EMAP, const double *__restrict__ P, int64_t np,
                    const int32_t *__restrict__ I, const double *__restrict__ C, double *__restrict__ S, double *__restrict__ N, bool keep_n)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
		const V3 p = ld3(P + 3 * i), c = ld3(C + 3 * i);
		V3 n = {0, 0, 0};
		const double s = pseudonormal(V, F, nF, FN, VN, EN, EMAP, p, I[i], c, n);
		if (S) S[i] = s * sqrt(S[i]);                       // S held the squared distance
		if (N && keep_n) { N[3 * i] = n.x; N[3 * i + 1] = n.y; N[3 * i + 2] = n.z; }
	}
}

} // namespace

namespace fpohm {

void launch_closest_point(fpohm_ctx *ctx, fpohm_mesh *m, bool with_sign, const double *P_dev, int64_t np,
                          double *S, int32_t *I, double *C, double *N, cudaStream_t s)
{
	if (np <= 0) return;
	const int blk = 128;
	const int grid = grid_for(ctx, np, blk, 16);
	static const bool stats = getenv("FPOHM_CP_STATS") != nullptr;   // debug only: N[:,0:2] := (node visits, leaf tests)
	// the sign pass needs facet + closest point even when the caller did not ask for them
	DevBuf<int32_t> tmpI; DevBuf<double> tmpC, tmpS;
	if (with_sign) {
		if (!I) { tmpI.alloc(np, s); I = tmpI.p; }
		if (!C) { tmpC.alloc(3 * np, s); C = tmpC.p; }
		if (!S && N) { tmpS.alloc(np, s); S = tmpS.p; }
	}
	if (stats) closest_point_kernel<true><<<grid, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, P_dev, np, S, I, C, N);
	else closest_point_kernel<false><<<grid, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, P_dev, np, S, I, C, N);
	FPOHM_LAUNCH_CHECK(ctx);
	if (with_sign && (S || N)) {
		pseudonormal_kernel<<<grid, blk, 0, s>>>(m->V.p, m->F.p, m->nF, m->FN.p, m->VN.p, m->EN.p, m->EMAP.p, P_dev, np, I, C, S, N, !stats);
		FPOHM_LAUNCH_CHECK(ctx);
	}
}

} // namespace fpohm

extern "C" {

int fpohm_signed_distance_dev(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P_dev, int64_t np,
                              double *S_dev, int32_t *I_dev, double *C_dev, double *N_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && (P_dev || np == 0), FPOHM_EINVAL, "fpohm_signed_distance_dev: null argument");
	FPOHM_REQUIRE(np >= 0, FPOHM_EINVAL, "fpohm_signed_distance_dev: negative count");
	DeviceGuard g(ctx->device);
	mesh_ensure_tree(ctx, mesh, ctx->stream);
	launch_closest_point(ctx, mesh, true, P_dev, np, S_dev, I_dev, C_dev, N_dev, (cudaStream_t)stream);
	FPOHM_API_END
}

// Host-pointer path.  Queries are independent, so the batch is cut into chunks that rotate over three streams:
// chunk k+1 is on its way up (H2D) while chunk k computes and chunk k-1 is on its way down (D2H).  With pinned caller
// buffers the PCIe copies (84 B/query) hide behind the traversal; with pageable memory CUDA stages the copies and the
// pipeline degrades gracefully to the serial order.
static int host_query(fpohm_ctx *ctx, fpohm_mesh *mesh, bool with_sign, const double *P, int64_t np,
                      double *S, int32_t *I, double *C, double *N, const char *who)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && mesh && (P || np == 0), FPOHM_EINVAL, "%s: null argument", who);
	FPOHM_REQUIRE(np >= 0, FPOHM_EINVAL, "%s: negative count", who);
	if (np == 0) return FPOHM_OK;
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	mesh_ensure_tree(ctx, mesh, s);
	DevBuf<double> dP(3 * np, s), dS(S ? np : 0, s), dC(C ? 3 * np : 0, s), dN(N ? 3 * np : 0, s);
	DevBuf<int32_t> dI(I ? np : 0, s);
	KernelTimer t(ctx, s);
	FPOHM_CUDA(cudaEventRecord(ctx->ev_sync, s));              // allocations are ordered on s
	cudaStream_t lanes[3] = {s, ctx->aux[0], ctx->aux[1]};
	for (int k = 1; k < 3; ++k) FPOHM_CUDA(cudaStreamWaitEvent(lanes[k], ctx->ev_sync, 0));
	const int64_t chunk = 1 << 19;
	int k = 0;
	for (int64_t o = 0; o < np; o += chunk, ++k) {
		const int64_t n = std::min(chunk, np - o);
		cudaStream_t ls = lanes[k % 3];
		FPOHM_CUDA(cudaMemcpyAsync(dP.p + 3 * o, P + 3 * o, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, ls));
		launch_closest_point(ctx, mesh, with_sign, dP.p + 3 * o, n, S ? dS.p + o : nullptr, I ? dI.p + o : nullptr,
		                     C ? dC.p + 3 * o : nullptr, N ? dN.p + 3 * o : nullptr, ls);
		if (S) FPOHM_CUDA(cudaMemcpyAsync(S + o, dS.p + o, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ls));
		if (I) FPOHM_CUDA(cudaMemcpyAsync(I + o, dI.p + o, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ls));
		if (C) FPOHM_CUDA(cudaMemcpyAsync(C + 3 * o, dC.p + 3 * o, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ls));
		if (N) FPOHM_CUDA(cudaMemcpyAsync(N + 3 * o, dN.p + 3 * o, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ls));
	}
	for (int j = 1; j < 3; ++j) {                             // join the side lanes back into s before the buffers die
		FPOHM_CUDA(cudaEventRecord(ctx->ev_sync, lanes[j]));
		FPOHM_CUDA(cudaStreamWaitEvent(s, ctx->ev_sync, 0));
	}
	t.stop();
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_signed_distance(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P, int64_t np,
                          double *S, int32_t *I, double *C, double *N)
{
	return host_query(ctx, mesh, true, P, np, S, I, C, N, "fpohm_signed_distance");
}

int fpohm_point_mesh_sqdist(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P, int64_t np,
                            double *sqrD, int32_t *I, double *C)
{
	return host_query(ctx, mesh, false, P, np, sqrD, I, C, nullptr, "fpohm_point_mesh_sqdist");
}

} // extern "C"

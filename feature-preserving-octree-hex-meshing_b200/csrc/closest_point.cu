// Batched closest point + pseudonormal signed distance over the flattened igl-identical tree.
//
// Replaces igl::signed_distance_pseudonormal (igl/signed_distance.cpp:186-257):
//   AABB::squared_distance   igl/AABB.cpp:352-441     -> traverse()
//   leaf_squared_distance    igl/AABB.cpp:735-750
//   set_min (strict '<')     igl/AABB.cpp:754-779
//   point_simplex_squared_distance (Ericson) igl/point_simplex_squared_distance.cpp:44-124 -> closest_on_triangle()
//   pseudonormal_test        igl/pseudonormal_test.cpp:14-119 -> pseudonormal()
//
// Two paths produce the same bits.  The reference path (`closest_point_kernel`, FPOHM_CP_MODE=0, and the fallback
// of every other kernel) is one thread per query walking in exactly igl's order.  The shipped path ("Packet search"
// below) finds the minimum distance with warp packets in any order and then reproduces igl's choice among equidistant
// facets.  Every fp64 expression keeps igl/Eigen's association (3-term sums are a0 + (a1 + a2), Eigen 3.2 unrolled
// redux) and the library is built with -fmad=false, so the device evaluates the same IEEE operations as the
// reference's SSE2 build.
#include "mesh.h"
#include "octree.h"
#include <cub/device/device_radix_sort.cuh>
#include <cmath>
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

// ---------------------------------------------------------------------------------------------------------------------
// Packet search (default path).  Four kernels, each handing on only what it could not finish:
//
//  K1 packet   (cp_packet_kernel) one WARP walks the tree once for its 32 consecutive queries.  The per-lane igl-order kernel above is
//              bound by instruction issue, a third of it fp64 box arithmetic, at 10 of 32 lanes active
//              (profiles/r01_ncu_summary.md).  Here the walk, the stack and every node / triangle load are
//              warp-uniform and the box tests are a CONSERVATIVE fp32 FILTER: child boxes rounded outwards to float
//              (QNodeF, 64 B per node), the query rounded down/up, every operation rounded down and the result shrunk
//              by 2^-20, so a lane's bound is always below the exact distance to anything in the box.  A lane wants a
//              child iff bound <= its best distance (rounded up).  The only fp64 work left is the exact igl
//              closest-point evaluation of the leaves that pass the filter, so a lane that sees the walk through holds
//              the exact minimum distance over ALL facets, whatever the visiting order was.
//  K2 search   (cp_search_kernel) queries of a warp that stopped sharing its search (checked every 32 visits) restart an
//              order-free per-lane search seeded with the bound they already have: persistent lanes, one node visit per
//              loop iteration, a lane that finishes fetches the next query at once.  Bounded work per query.
//  K2 tie      (cp_tie_kernel) one THREAD per query with near-ties: igl's tie-break (below).
//  K3 heavy    (cp_heavy_kernel) one WARP per query that exceeded K2's budgets, 32 tree nodes per step (0.2 % of the bench queries hold
//              the tail: one near the gear's bore axis needs 6 590 node visits, milliseconds for a lone thread).
//
// Ties.  What the visiting order decides in igl is which facet wins when several are at (nearly) the same distance:
// igl keeps the first one its depth-first order reaches — left child first if it contains p or is nearer
// (AABB.cpp:392-437), independent of the running minimum, so for a given p the order is a fixed total order of the
// leaves — and its strict pruning on the COMPUTED box distances can even skip a facet that is an ulp closer.  A lane
// therefore counts the leaves within (1 + 2^-40) of its minimum and keeps the first three at EXACTLY the minimum.  With
// more than one near-tie, K2 names the winner from that list when one box test proves igl reaches it (`igl_tie_winner`);
// otherwise it re-walks the tree in exact igl order (`traverse_limited`) restricted to boxes within (1 + 2^-38) of the
// minimum.  Restricting is exact: leaves
// outside those boxes are farther than every near-minimal candidate, so they can neither win nor make igl prune an
// ancestor of a candidate (its box distance is below their distance); by induction over igl's order the restricted
// walk evaluates the same candidates with the same running minimum as the full walk.  The walk stops at the first
// facet at exactly the minimum distance (igl replaces its candidate only by a strictly closer one).  K3 finds that
// facet in parallel instead: every stack entry carries its rank in igl's order as a bit string (0 = the child igl looks
// at first), the exact-minimum leaf with the smallest rank is igl's winner PROVIDED igl reaches it, which holds when
// no box on its path is farther than the minimum (checked; otherwise one lane does the rigorous re-walk).
#define PK_STACK 64
#define PK_WINDOW 32
#define PK_MIN_WANT 3
#define PK_A_BUDGET 1024
#define PK_EPS_TIE 9.094947017729282e-13      /* 2^-40 */
#define PK_EPS_WALK 3.637978807091713e-12     /* 2^-38 */
#define K2_STACK 64
#define K2_SEARCH_BUDGET 320
#define K2_WALK_BUDGET 96
#define HV_STACK 768

__device__ __forceinline__ float box_low(const float *__restrict__ mn, const float *__restrict__ mx,
                                         float plx, float ply, float plz, float phx, float phy, float phz)
{
	const float gx = fmaxf(fmaxf(__fsub_rd(mn[0], phx), __fsub_rd(plx, mx[0])), 0.f);
	const float gy = fmaxf(fmaxf(__fsub_rd(mn[1], phy), __fsub_rd(ply, mx[1])), 0.f);
	const float gz = fmaxf(fmaxf(__fsub_rd(mn[2], phz), __fsub_rd(plz, mx[2])), 0.f);
	return __fmul_rd(__fmaf_rd(gz, gz, __fmaf_rd(gy, gy, __fmul_rd(gx, gx))), 0.99999905f);
}

#define PK_TIES 7          /* exact ties kept besides the winner: a vertex of valence 8 (the first version kept 3 and sent every
                              closest-to-a-vertex query — 18 % of the tie list of far-away lattice points — into the igl-order re-walk) */
#define TT_STRIDE (PK_TIES + 1)
struct Packet {
	V3 p;
	double best;
	float best_hi;
	int32_t bf;
	V3 bc;
	int near;       // leaves within (1 + PK_EPS_TIE) of best, the best one included
	int nt;         // leaves at EXACTLY best besides bf; the first PK_TIES are kept in the tie slots
};

template <bool KEEP_C, int STRIDE>
__device__ __forceinline__ void lane_update(Packet &k, int32_t *__restrict__ ties, int32_t prim, const V3 &q, double d) {
	if (d < k.best) {
		k.near = (d + d * PK_EPS_TIE < k.best) ? 1 : k.near + 1;
		k.nt = 0;
		k.best = d; k.bf = prim;
		if (KEEP_C) k.bc = q;
		k.best_hi = __double2float_ru(d);
	} else if (d <= k.best + k.best * PK_EPS_TIE) {
		++k.near;
		if (d == k.best) { if (k.nt < PK_TIES) ties[k.nt * STRIDE] = prim; ++k.nt; }
	}
}
__device__ __forceinline__ void lane_leaf(const double *__restrict__ tri, int32_t prim, Packet &k, int32_t *__restrict__ ties) {
	const double *t = tri + 9 * (int64_t)prim;
	const V3 q = closest_on_triangle(k.p, ld3(t), ld3(t + 3), ld3(t + 6));
	lane_update<true, 1>(k, ties, prim, q, sqnorm(sub(k.p, q)));
}

// true iff igl's depth-first order for query p reaches facet fa before facet fb: decided at their lowest common
// ancestor, where igl looks first at the child that contains p or is nearer (AABB.cpp:392-437).
__device__ __forceinline__ bool igl_visits_first(const QNode *__restrict__ nodes, const int32_t *__restrict__ prim_parent,
                                                 const V3 &p, int32_t fa, int32_t fb, const int2 *__restrict__ pd = nullptr)
{
	int32_t na = prim_parent[fa], nb = prim_parent[fb];
	int32_t ca = ~fa;                                // child reference through which a's side enters the ancestor (b's is the other one)
	if (pd) {
		// the climb reads 8 bytes per node from a table that stays in L2 (16 MB at 2 M facets) instead of a 128-byte node line
		// from DRAM (ncu launch list, C4 step: the tie-break was 16 % of the step, 6.8 ms for 7.6 M queries)
		int2 a = __ldg(pd + na), b = __ldg(pd + nb);
		while (a.y > b.y) { ca = na; na = a.x; a = __ldg(pd + na); }
		while (b.y > a.y) { nb = b.x; b = __ldg(pd + nb); }
		while (na != nb) { ca = na; na = a.x; a = __ldg(pd + na); nb = b.x; b = __ldg(pd + nb); }
		const QNode *n = nodes + na;
		const double dl = box_ext_sqdist(n->lmin, n->lmax, p), dr = box_ext_sqdist(n->rmin, n->rmax, p);
		const bool left_first = box_contains(n->lmin, n->lmax, p) || dl < dr;
		return (n->left == ca) == left_first;
	}
	int da = nodes[na].depth, db = nodes[nb].depth;
	while (da > db) { ca = na; na = nodes[na].parent; --da; }
	while (db > da) { nb = nodes[nb].parent; --db; }
	while (na != nb) { ca = na; na = nodes[na].parent; nb = nodes[nb].parent; }
	const QNode *n = nodes + na;
	const double dl = box_ext_sqdist(n->lmin, n->lmax, p), dr = box_ext_sqdist(n->rmin, n->rmax, p);
	const bool left_first = box_contains(n->lmin, n->lmax, p) || dl < dr;
	return (n->left == ca) == left_first;
}

// igl's winner among the facets at exactly the minimum distance, when it can be named without walking the tree:
// a* = the first of them in igl's order.  igl reaches a* if no box on its path is farther than the minimum; boxes nest,
// so every per-axis gap of an ancestor is <= the gap of a*'s own box and, the fp operations being monotone, the computed
// distance of every ancestor box is <= the computed distance of a*'s box: ONE box test proves the whole path.  Until
// a* is evaluated igl's running minimum is above the true minimum (a* is the first facet at that distance), so the
// strict test on those boxes passes, a* is evaluated, and nothing later can replace it.  Returns -1 if the box test
// fails (then only the walk in igl's order can tell).
__device__ __forceinline__ int32_t igl_tie_winner(const QNode *__restrict__ nodes, const int32_t *__restrict__ prim_parent,
                                                  const double *__restrict__ tri, const V3 &p, double dmin, int32_t bf,
                                                  const int32_t *__restrict__ ties, int stride, int nt, const int2 *__restrict__ pd = nullptr)
{
	int32_t win = bf;
	for (int j = 0; j < nt; ++j) {
		const int32_t f = ties[j * stride];
		if (igl_visits_first(nodes, prim_parent, p, f, win, pd)) win = f;
	}
	const double *t = tri + 9 * (int64_t)win;
	const V3 a = ld3(t), b = ld3(t + 3), c = ld3(t + 6);
	const double mn[3] = {fmin(a.x, fmin(b.x, c.x)), fmin(a.y, fmin(b.y, c.y)), fmin(a.z, fmin(b.z, c.z))};
	const double mx[3] = {fmax(a.x, fmax(b.x, c.x)), fmax(a.y, fmax(b.y, c.y)), fmax(a.z, fmax(b.z, c.z))};
	return box_ext_sqdist(mn, mx, p) <= dmin ? win : -1;
}

// Sign finalisation shared by every kernel of the wide-packet pipeline (FPOHM_CP_MODE=2): whoever settles a query also runs
// pseudonormal_test on it, so P / I / C are not read back by a separate pass.  S holds the SQUARED distance until then.
struct SignArgs {
	const double *V; const int32_t *F; int64_t nF;
	const double *FN, *VN, *EN; const int32_t *EMAP;
	double *N;          // may be null
	int enabled;        // 0: unsigned query (S stays the squared distance) or the separate pseudonormal pass does it
};
__device__ __forceinline__ void finalize_sign(const SignArgs &sa, int64_t i, const V3 &p, int32_t f, const V3 &c, double d2,
                                              double *__restrict__ S)
{
	if (!sa.enabled) return;
	V3 n = {0, 0, 0};
	if (f < 0) {
		if (S) S[i] = CUDART_NAN;
	} else {
		const double s = pseudonormal(sa.V, sa.F, sa.nF, sa.FN, sa.VN, sa.EN, sa.EMAP, p, f, c, n);
		if (S) S[i] = s * sqrt(d2);
	}
	if (sa.N) { sa.N[3 * i] = n.x; sa.N[3 * i + 1] = n.y; sa.N[3 * i + 2] = n.z; }
}

// work-list entry codes
#define TODO_WALK 1      /* minimum is exact, near-ties need igl's tie-break */
#define TODO_SEARCH 2    /* S holds an upper bound (or +inf): search not finished */
#define TODO_HEAVY 3     /* as TODO_SEARCH, straight to the warp-per-query kernel */

// ---- K1 ---------------------------------------------------------------------------------------------------------
// Register diet (ncu: at 72 registers the first version re-loaded five spilled query coordinates from local memory in
// EVERY node step and fetched the node with twelve 4-byte loads): the node step only needs the query as three floats
// and one threshold, so the fp64 query and the running minimum live in global / shared memory and are read in the leaf
// steps only (2.5x rarer).  Filter for a query p, pf = float(p), delta >= |p - pf|: the exact distance D to anything in
// the box obeys D >= D_f - delta with D_f = dist(pf, box_f); so D^2 <= best implies D_f^2 <= (sqrt(best) + delta)^2 =: thr.
// D_f^2 is evaluated rounded down, thr rounded up and widened by 2^-19.
struct NodeBoxes { float4 a, b, c; int32_t left, right; };
__device__ __forceinline__ NodeBoxes load_node(const QNodeF *__restrict__ n) {
	const float4 *q = reinterpret_cast<const float4 *>(n);
	NodeBoxes r;
	r.a = __ldg(q); r.b = __ldg(q + 1); r.c = __ldg(q + 2);
	const int2 lr = __ldg(reinterpret_cast<const int2 *>(q + 3));
	r.left = lr.x; r.right = lr.y;
	return r;
}
__device__ __forceinline__ float gap2_low(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float px, float py, float pz) {
	const float gx = fmaxf(fmaxf(__fsub_rd(mnx, px), __fsub_rd(px, mxx)), 0.f);
	const float gy = fmaxf(fmaxf(__fsub_rd(mny, py), __fsub_rd(py, mxy)), 0.f);
	const float gz = fmaxf(fmaxf(__fsub_rd(mnz, pz), __fsub_rd(pz, mxz)), 0.f);
	return __fmaf_rd(gz, gz, __fmaf_rd(gy, gy, __fmul_rd(gx, gx)));
}
__device__ __forceinline__ float filter_threshold(double best, float delta) {
	const float r = __fadd_ru(__fsqrt_ru(__double2float_ru(best)), delta);
	return __fmul_ru(__fmul_ru(r, r), 1.0000020f);      // (sqrt(best) + delta)^2, up, widened by 2^-19
}

template <bool STATS, int MINB>
__global__ void __launch_bounds__(128, MINB)
cp_packet_kernel(const QNodeF *__restrict__ fnodes, int32_t root, const double *__restrict__ tri,
                 const double *__restrict__ P, int64_t np,
                 double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C, double *__restrict__ N,
                 int32_t *__restrict__ todo, int32_t *__restrict__ todo_ties, int32_t *__restrict__ todo_count, int min_want_sum, int a_budget)
{
	__shared__ int32_t s_stack[4][PK_STACK];
	__shared__ int32_t s_tie[4][PK_TIES][32];
	__shared__ double s_best[4][32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int32_t *stk = s_stack[warp];
	int32_t *ties = &s_tie[warp][0][lane];
	double *bestp = &s_best[warp][lane];
	const int64_t nwarps = (int64_t)gridDim.x * 4;
	for (int64_t w = blockIdx.x * 4ll + warp; w * 32 < np; w += nwarps) {
		const int64_t i = w * 32 + lane;
		const bool valid = i < np;
		const double *pp = P + 3 * (valid ? i : 0);
		float pfx = 0, pfy = 0, pfz = 0, delta = 0;
		{
			const V3 p = ld3(pp);
			pfx = (float)p.x; pfy = (float)p.y; pfz = (float)p.z;
			const double ex = p.x - (double)pfx, ey = p.y - (double)pfy, ez = p.z - (double)pfz;
			delta = __double2float_ru(sqrt(ex * ex + ey * ey + ez * ez) * 1.0000001);
		}
		*bestp = CUDART_INF;
		float thr = valid ? CUDART_INF_F : -1.f;        // an idle lane never wants a node (bounds are >= 0)
		int32_t bf = -1;
		int near = 0, nt = 0;
		int visits = 0, window = 0, leaf_steps = 0;
		int top = 0;
		bool bail = false;
		stk[top++] = root;                              // every lane writes the same value: no synchronisation needed
		long long t0 = 0;
		if (STATS) t0 = clock64();
		// exact evaluation of one (warp-uniform) leaf by the lanes whose filter passed
		auto leaf = [&](int32_t prim, bool want) {
			const double *t = tri + 9 * (int64_t)prim;
			const V3 a = ld3(t), b = ld3(t + 3), c = ld3(t + 6);
			if (want) {
				const V3 p = ld3(pp);
				const V3 q = closest_on_triangle(p, a, b, c);
				const double d = sqnorm(sub(p, q));
				const double best = *bestp;
				if (d < best) {
					near = (d + d * PK_EPS_TIE < best) ? 1 : near + 1;
					nt = 0; bf = prim;
					*bestp = d;
					thr = filter_threshold(d, delta);
				} else if (d <= best + best * PK_EPS_TIE) {
					++near;
					if (d == best) { if (nt < PK_TIES) ties[32 * nt] = prim; ++nt; }
				}
			}
			if (STATS) ++leaf_steps;
		};
		while (top > 0) {
			// the per-warp stack is shared memory every lane pushes the same values to: vote intrinsics are not memory barriers, so
			// the pushes of the previous iteration are ordered against this pop explicitly
			__syncwarp();
			const int32_t cur = stk[--top];
			__syncwarp();                                    // nobody re-uses the slot before every lane has read it
			const NodeBoxes nb = load_node(fnodes + cur);
			const float dl = gap2_low(nb.a.x, nb.a.y, nb.a.z, nb.a.w, nb.b.x, nb.b.y, pfx, pfy, pfz);
			const float dr = gap2_low(nb.b.z, nb.b.w, nb.c.x, nb.c.y, nb.c.z, nb.c.w, pfx, pfy, pfz);
			const bool wl = dl <= thr, wr = dr <= thr;
			const unsigned ml = __ballot_sync(0xffffffffu, wl), mr = __ballot_sync(0xffffffffu, wr);
			if (!(ml | mr)) continue;
			if ((visits & (PK_WINDOW - 1)) == PK_WINDOW - 1) {
				if (window < min_want_sum) { stk[top++] = cur; bail = true; break; }     // the lanes stopped sharing their search
				window = 0;
			}
			if (visits >= a_budget || top + 2 > PK_STACK) { stk[top++] = cur; bail = true; break; }
			++visits;
			window += __popc(ml | mr);
			// nearer child first: majority vote of the lanes that still want this node
			const unsigned near_l = __ballot_sync(0xffffffffu, (wl || wr) && (dl < dr || dl == 0.f));
			const bool left_first = 2 * __popc(near_l) >= __popc(ml | mr);
			const int32_t c1 = left_first ? nb.left : nb.right, c2 = left_first ? nb.right : nb.left;
			const float d2 = left_first ? dr : dl;
			const bool w1 = left_first ? wl : wr;
			const unsigned m1 = left_first ? ml : mr, m2 = left_first ? mr : ml;
			if (c1 < 0) {
				if (m1) leaf(~c1, w1);
				const unsigned m2b = __ballot_sync(0xffffffffu, d2 <= thr);       // the bound may have dropped
				if (m2b) {
					if (c2 < 0) leaf(~c2, d2 <= thr);
					else stk[top++] = c2;
				}
			} else {
				if (m2) {
					if (c2 < 0) leaf(~c2, d2 <= thr);
					else stk[top++] = c2;
				}
				if (m1) stk[top++] = c1;
			}
		}
		// a warp that gave up: only the lanes that still want something on the stack have an unfinished search
		bool unfinished = false;
		if (bail) {
			for (int j = 0; j < top; ++j) {
				const NodeBoxes nb = load_node(fnodes + stk[j]);
				unfinished |= gap2_low(nb.a.x, nb.a.y, nb.a.z, nb.a.w, nb.b.x, nb.b.y, pfx, pfy, pfz) <= thr ||
				              gap2_low(nb.b.z, nb.b.w, nb.c.x, nb.c.y, nb.c.z, nb.c.w, pfx, pfy, pfz) <= thr;
			}
		}
		const int code = !valid ? 0 : (unfinished ? TODO_SEARCH : (near > 1 ? TODO_WALK : 0));
		if (valid) {
			V3 bc = {0, 0, 0};
			if (bf >= 0) { const double *t = tri + 9 * (int64_t)bf; bc = closest_on_triangle(ld3(pp), ld3(t), ld3(t + 3), ld3(t + 6)); }
			I[i] = bf;
			C[3 * i] = bc.x; C[3 * i + 1] = bc.y; C[3 * i + 2] = bc.z;
			S[i] = *bestp;
		}
		// warp-aggregated append to the work lists (tie-breaks grow from the front of `todo`, unfinished searches from
		// the back, so that each completion kernel gets warps full of one kind of work)
		const unsigned mw = __ballot_sync(0xffffffffu, code == TODO_WALK), ms = __ballot_sync(0xffffffffu, code == TODO_SEARCH);
		if (mw | ms) {
			int bw = 0, bs = 0;
			if (lane == 0) { if (mw) bw = atomicAdd(todo_count, __popc(mw)); if (ms) bs = atomicAdd(todo_count + 2, __popc(ms)); }
			bw = __shfl_sync(0xffffffffu, bw, 0); bs = __shfl_sync(0xffffffffu, bs, 0);
			const unsigned lt = (1u << lane) - 1;
			if (code == TODO_WALK) {
				const int slot = bw + __popc(mw & lt);
				todo[slot] = (int32_t)i;
				int32_t *tr = todo_ties + TT_STRIDE * (int64_t)slot;
				tr[0] = nt;
				for (int j = 0; j < PK_TIES; ++j) tr[1 + j] = j < nt ? ties[32 * j] : -1;
			} else if (code == TODO_SEARCH) {
				todo[2 * np - 1 - (bs + __popc(ms & lt))] = (int32_t)i;
			}
		}
		if (STATS && N && valid) { N[3 * i] = visits + 65536.0 * leaf_steps; N[3 * i + 1] = (double)(clock64() - t0); N[3 * i + 2] = code; }
		__syncwarp();
	}
}

// ---- K1, wide form (default, FPOHM_CP_MODE=2) --------------------------------------------------------------------
// ncu on the binary packet walk above: 70 % of the issue slots busy, 2.93 G warp instructions = 101.6 node visits x 112
// instructions + 40.7 warp-uniform fp64 leaf evaluations per packet, at 25 of 32 lanes.  A packet's lanes agree on almost every
// box of the upper tree, so testing 2 boxes per step with 32 lanes each spends 32 lanes on one bit of information.  This
// kernel turns the upper tree around:
//   phase A  BOX-parallel: lane (j, c) of the warp tests child c of the j-th of the four wide nodes on top of the stack
//            (32 boxes per step, one 32-byte WChild entry per lane) against the PACKET — the box of the lanes' float queries
//            and T = the largest lane threshold.  box-box gap <= every lane's point-box gap and T >= every lane's thr, so
//            nothing a lane could want is dropped.  Wanted inner nodes go back on the stack, wanted clusters (nodes of <= 8
//            facets) and stray facets into a candidate list.
//   phase B  QUERY-parallel: per candidate cluster one box test per lane with the lane's OWN threshold (this is what keeps
//            far-away packets cheap: the union of the lanes' balls is far smaller than any ball around the packet), then the
//            8 facet boxes; the facets a lane wants become bits of a per-lane mask over up to 4 staged clusters.
//   phase C  exact fp64 evaluation, each lane walking its OWN mask (nearest box first, filter re-checked as the bound drops):
//            the step count is the longest lane's list, not the union of all lists as in the warp-uniform leaf steps.
// A greedy descent for the packet centre finds one facet; every lane's exact distance to it seeds thr / T.
// Bookkeeping of the minimum, the near-tie count and the exact ties is the old one (`near`, `nt`, tie slots), so the
// completion kernels and igl's tie-break are untouched.  Results are bit-identical to the binary walk (scripts/cp_ab.py).
#define WK_STACK 160
#define WK_CAND 64
#define PK_CAND 48        /* pair form: candidates per round */
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ float boxgap2_low(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                             float qlx, float qly, float qlz, float qhx, float qhy, float qhz) {
	const float gx = fmaxf(fmaxf(__fsub_rd(lox, qhx), __fsub_rd(qlx, hix)), 0.f);
	const float gy = fmaxf(fmaxf(__fsub_rd(loy, qhy), __fsub_rd(qly, hiy)), 0.f);
	const float gz = fmaxf(fmaxf(__fsub_rd(loz, qhz), __fsub_rd(qlz, hiz)), 0.f);
	return __fmaf_rd(gz, gz, __fmaf_rd(gy, gy, __fmul_rd(gx, gx)));
}

struct WideStage { uint4 e[4][16]; };      // four staged clusters, 8 entries of two uint4 each

template <bool STATS, int MINB>
__global__ void __launch_bounds__(128, MINB)
cp_wide_kernel(const WNode *__restrict__ wnodes, const double *__restrict__ tri,
               const double *__restrict__ P, int64_t np, const uint32_t *__restrict__ perm,
               double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C, SignArgs sa,
               int32_t *__restrict__ todo, int32_t *__restrict__ todo_ties, int32_t *__restrict__ todo_count, double *__restrict__ stats)
{
	__shared__ int32_t s_stack[4][WK_STACK];
	__shared__ int32_t s_cid[4][WK_CAND];
	__shared__ float s_cbox[4][6][WK_CAND];
	__shared__ WideStage s_stage[4];
	__shared__ int32_t s_tie[4][PK_TIES][32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1;
	int32_t *stk = s_stack[warp];
	int32_t *cid = s_cid[warp];
	float (*cbox)[WK_CAND] = s_cbox[warp];
	uint4 (*stage)[16] = s_stage[warp].e;
	int32_t *ties = &s_tie[warp][0][lane];
	const uint4 *wn4 = reinterpret_cast<const uint4 *>(wnodes);
	const int64_t nwarps = (int64_t)gridDim.x * 4;
	for (int64_t w = blockIdx.x * 4ll + warp; w * 32 < np; w += nwarps) {
		const int64_t slot = w * 32 + lane;
		const bool in_range = slot < np;
		const int64_t i = in_range ? (perm ? (int64_t)perm[slot] : slot) : 0;
		const V3 p = ld3(P + 3 * i);
		const bool valid = in_range && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
		const float pfx = (float)p.x, pfy = (float)p.y, pfz = (float)p.z;
		float delta;
		{
			const double ex = p.x - (double)pfx, ey = p.y - (double)pfy, ez = p.z - (double)pfz;
			delta = __double2float_ru(sqrt(ex * ex + ey * ey + ez * ez) * 1.0000001);
		}
		// packet box over the searching lanes
		const float qlx = ord2f(__reduce_min_sync(FULL, valid ? f2ord(pfx) : 0x7fffffff)), qhx = ord2f(__reduce_max_sync(FULL, valid ? f2ord(pfx) : (int)0x80000000));
		const float qly = ord2f(__reduce_min_sync(FULL, valid ? f2ord(pfy) : 0x7fffffff)), qhy = ord2f(__reduce_max_sync(FULL, valid ? f2ord(pfy) : (int)0x80000000));
		const float qlz = ord2f(__reduce_min_sync(FULL, valid ? f2ord(pfz) : 0x7fffffff)), qhz = ord2f(__reduce_max_sync(FULL, valid ? f2ord(pfz) : (int)0x80000000));
		double best = CUDART_INF;
		float thr = valid ? CUDART_INF_F : -1.f;        // an idle lane never wants anything (bounds are >= 0)
		int32_t bf = -1;
		V3 bc = {0, 0, 0};
		int near = 0, nt = 0;
		int st_steps = 0, st_cand = 0, st_staged = 0, st_iters = 0, st_evals = 0;
		long long t0 = 0;
		if (STATS) t0 = clock64();
		auto exact = [&](int32_t prim) {
			const double *t = tri + 9 * (int64_t)prim;
			const V3 q = closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6));
			const double d = sqnorm(sub(p, q));
			if (d < best) {
				near = (d + d * PK_EPS_TIE < best) ? 1 : near + 1;
				nt = 0; bf = prim; bc = q;
				best = d;
				thr = filter_threshold(d, delta);
			} else if (d <= best + best * PK_EPS_TIE) {
				++near;
				if (d == best) { if (nt < PK_TIES) ties[32 * nt] = prim; ++nt; }
			}
			if (STATS) ++st_evals;
		};
		bool bail = false;
		int32_t seed = -1;
		if (__any_sync(FULL, valid)) {
			// ---- seed: greedy descent for the packet centre (all four lane groups do the same work) ----
			{
				const float cx = 0.5f * qlx + 0.5f * qhx, cy = 0.5f * qly + 0.5f * qhy, cz = 0.5f * qlz + 0.5f * qhz;
				const int c = lane & 7;
				int32_t nd = 0;
				for (int guard = 0; guard < 64; ++guard) {
					const uint4 a = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2), b = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2 + 1);
					const float lox = __uint_as_float(a.x), loy = __uint_as_float(a.y), loz = __uint_as_float(a.z);
					const float hix = __uint_as_float(a.w), hiy = __uint_as_float(b.x), hiz = __uint_as_float(b.y);
					const int32_t ch = (int32_t)b.z;
					const bool ok = ch != WCHILD_EMPTY;
					const float gap = gap2_low(lox, loy, loz, hix, hiy, hiz, cx, cy, cz);
					const float mx = 0.5f * lox + 0.5f * hix - cx, my = 0.5f * loy + 0.5f * hiy - cy, mz = 0.5f * loz + 0.5f * hiz - cz;
					const float cd = mx * mx + my * my + mz * mz;
					const unsigned m0 = __ballot_sync(FULL, ok && gap == 0.f) & 0xffu;
					float key = !ok ? CUDART_INF_F : (m0 ? (gap == 0.f ? cd : CUDART_INF_F) : gap);
					int kc = c;
#pragma unroll
					for (int o = 1; o < 8; o <<= 1) {
						const float ok2 = __shfl_xor_sync(FULL, key, o);
						const int oc = __shfl_xor_sync(FULL, kc, o);
						if (ok2 < key || (ok2 == key && oc < kc)) { key = ok2; kc = oc; }
					}
					const int32_t wch = __shfl_sync(FULL, ch, kc);
					if (wch < 0) { seed = wch == WCHILD_EMPTY ? -1 : ~wch; break; }
					nd = wch;
				}
			}
			if (seed >= 0 && valid) exact(seed);
			float T = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(fmaxf(thr, 0.f))));
			// ---- main loop ----
			int top = 1, ncand = 0;
			if (lane == 0) stk[0] = 0;
			__syncwarp();
			for (;;) {
				// phase A
				while (top > 0 && ncand <= WK_CAND - 32) {
					const int j = lane >> 3, c = lane & 7;
					const int take = top < 4 ? top : 4;
					const bool act = j < take;
					const int32_t nd = act ? stk[top - 1 - j] : 0;
					__syncwarp();
					top -= take;
					const uint4 a = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2), b = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2 + 1);
					const float g = boxgap2_low(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z),
					                            __uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y), qlx, qly, qlz, qhx, qhy, qhz);
					const int32_t ch = (int32_t)b.z;
					const bool want = act && ch != WCHILD_EMPTY && g <= T;
					const bool to_stack = want && ch >= 0 && !(b.w & 1u);
					const bool to_cand = want && !to_stack;
					const unsigned mi = __ballot_sync(FULL, to_stack), mc = __ballot_sync(FULL, to_cand);
					if (top + __popc(mi) > WK_STACK) { bail = true; break; }
					if (to_stack) stk[top + __popc(mi & lt)] = ch;
					top += __popc(mi);
					if (to_cand) {
						const int k = ncand + __popc(mc & lt);
						cid[k] = ch;
						cbox[0][k] = __uint_as_float(a.x); cbox[1][k] = __uint_as_float(a.y); cbox[2][k] = __uint_as_float(a.z);
						cbox[3][k] = __uint_as_float(a.w); cbox[4][k] = __uint_as_float(b.x); cbox[5][k] = __uint_as_float(b.y);
					}
					ncand += __popc(mc);
					if (STATS) ++st_steps;
					__syncwarp();
				}
				if (bail || ncand == 0) break;
				if (STATS) st_cand += ncand;
				// phase B + C
				int nst = 0;
				unsigned mask = 0;
				float bestg = CUDART_INF_F;
				int bestbit = -1;
				auto run_exact = [&]() {
					bool first = true;
					while (__any_sync(FULL, mask != 0)) {
						if (mask) {
							const int bit = (first && bestbit >= 0) ? bestbit : __ffs(mask) - 1;
							mask &= ~(1u << bit);
							const uint4 a = stage[bit >> 3][2 * (bit & 7)], b = stage[bit >> 3][2 * (bit & 7) + 1];
							const int32_t prim = ~(int32_t)b.z;
							if (prim != seed &&
							    (first || gap2_low(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
							                       __uint_as_float(b.x), __uint_as_float(b.y), pfx, pfy, pfz) <= thr))
								exact(prim);
						}
						first = false;
						if (STATS) ++st_iters;
					}
					nst = 0; bestg = CUDART_INF_F; bestbit = -1;
					__syncwarp();
				};
				for (int k = 0; k < ncand; ++k) {
					const bool wk = gap2_low(cbox[0][k], cbox[1][k], cbox[2][k], cbox[3][k], cbox[4][k], cbox[5][k], pfx, pfy, pfz) <= thr;
					if (!__any_sync(FULL, wk)) continue;
					const int32_t id = cid[k];
					if (id >= 0) {
						reinterpret_cast<uint2 *>(stage[nst])[lane] = __ldg(reinterpret_cast<const uint2 *>(wnodes + id) + lane);
					} else if (lane < 8) {       // a stray facet: a cluster of one
						uint4 a, b;
						if (lane == 0) {
							a = make_uint4(__float_as_uint(cbox[0][k]), __float_as_uint(cbox[1][k]), __float_as_uint(cbox[2][k]), __float_as_uint(cbox[3][k]));
							b = make_uint4(__float_as_uint(cbox[4][k]), __float_as_uint(cbox[5][k]), (uint32_t)id, 0u);
						} else {
							a = make_uint4(0x7f800000u, 0x7f800000u, 0x7f800000u, 0xff800000u);
							b = make_uint4(0xff800000u, 0xff800000u, (uint32_t)WCHILD_EMPTY, 0u);
						}
						stage[nst][2 * lane] = a; stage[nst][2 * lane + 1] = b;
					}
					__syncwarp();
					if (wk) {
#pragma unroll
						for (int c = 0; c < 8; ++c) {
							const uint4 a = stage[nst][2 * c], b = stage[nst][2 * c + 1];
							const float gc = gap2_low(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
							                          __uint_as_float(b.x), __uint_as_float(b.y), pfx, pfy, pfz);
							if ((int32_t)b.z != WCHILD_EMPTY && gc <= thr) {
								mask |= 1u << (nst * 8 + c);
								if (gc < bestg) { bestg = gc; bestbit = nst * 8 + c; }
							}
						}
					}
					if (STATS) ++st_staged;
					if (++nst == 4) run_exact();
				}
				if (nst) run_exact();
				ncand = 0;
				T = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(fmaxf(thr, 0.f))));
			}
		}
		// ---- epilogue ----
		const int code = !valid ? 0 : (bail ? TODO_SEARCH : (near > 1 ? TODO_WALK : 0));
		if (in_range) {
			I[i] = bf;
			C[3 * i] = bc.x; C[3 * i + 1] = bc.y; C[3 * i + 2] = bc.z;
			if (code == 0 && sa.enabled) finalize_sign(sa, i, p, bf, bc, best, S);
			else S[i] = best;
		}
		const unsigned mw = __ballot_sync(FULL, code == TODO_WALK), ms = __ballot_sync(FULL, code == TODO_SEARCH);
		if (mw | ms) {
			int bw = 0, bs = 0;
			if (lane == 0) { if (mw) bw = atomicAdd(todo_count, __popc(mw)); if (ms) bs = atomicAdd(todo_count + 2, __popc(ms)); }
			bw = __shfl_sync(FULL, bw, 0); bs = __shfl_sync(FULL, bs, 0);
			if (code == TODO_WALK) {
				const int sl = bw + __popc(mw & lt);
				todo[sl] = (int32_t)i;
				int32_t *tr = todo_ties + TT_STRIDE * (int64_t)sl;
				tr[0] = nt;
				for (int j = 0; j < PK_TIES; ++j) tr[1 + j] = j < nt ? ties[32 * j] : -1;
			} else if (code == TODO_SEARCH) {
				todo[2 * np - 1 - (bs + __popc(ms & lt))] = (int32_t)i;
			}
		}
		if (STATS && stats && in_range) {
			double *o = stats + 3 * i;
			o[0] = st_steps + 65536.0 * st_cand + 4294967296.0 * code;
			o[1] = st_staged + 65536.0 * st_iters + 4294967296.0 * st_evals;
			o[2] = (double)(clock64() - t0);
		}
		__syncwarp();
	}
}

// ---- K1, pair form (default, FPOHM_CP_MODE=3) -----------------------------------------------------------------------
// Counters of the wide form above on the bench workload: 9.5 box-parallel steps, 76 candidate clusters, 23 of them staged
// and 61 exact-loop iterations per packet at 26 % lane occupancy — the fp64 evaluations (15.7 per lane, which is what a
// bounding-box filter leaves over in ANY order) cost more than the whole tree walk, and a lane tests the 8 facet boxes of
// every staged cluster although it wants a third of the clusters.  This form keeps phase A and changes the rest:
//   * everything below the candidate list is PAIR-parallel: (lane, cluster) pairs are queued in shared memory and each
//     lane of the warp takes one pair, whatever query it belongs to (Q1 -> 8 facet-box tests); the (lane, facet) pairs
//     that pass are queued again (Q2) and refined the same way.  Occupancy no longer depends on the lanes agreeing.
//   * an fp32 REFINE between the box test and the fp64 evaluation: Ericson's closest point in float on the float-rounded
//     triangle gives a point q on it, hence an UPPER bound U >= D (distance to a point of the triangle, rounding slack
//     added) that tightens the owner's threshold without any fp64 work, and a LOWER bound L <= D from the supporting
//     plane through the triangle's vertices normal to p - q: for every x in the triangle n.(p - x) >= min_v n.(p - v).
//     Both are rigorous whatever q is: the dot products are rounded in the safe direction and the subtraction error of
//     p - v is bounded term by term, so a facet is only dropped when it is provably farther than the threshold.
//     Facets with L^2 <= thr go to the owner's survivor list; only those reach igl's exact fp64 evaluation — the facets
//     within ~1e-6 of the minimum instead of everything whose box is near.
//   * the survivors are evaluated by their owner (nearest bound first), with the old bookkeeping of minimum, near-tie
//     count and exact ties, so the completion kernels and igl's tie-break are untouched and results stay bit-identical.
#define PQ_CAP 64
#define PQ2_CAP 352        /* < 64 pending + 8 x 32 from one B2 iteration */
#define SL_CAP 8
#define PB_SLOTS 6
struct RefineOut { float L2, U; };
__device__ __forceinline__ RefineOut refine_f32(float px, float py, float pz, const float4 A, const float4 B, const float4 Cv, float slack_q)
{
	const float abx = B.x - A.x, aby = B.y - A.y, abz = B.z - A.z;
	const float acx = Cv.x - A.x, acy = Cv.y - A.y, acz = Cv.z - A.z;
	const float apx = px - A.x, apy = py - A.y, apz = pz - A.z;
	const float bpx = px - B.x, bpy = py - B.y, bpz = pz - B.z;
	const float cpx = px - Cv.x, cpy = py - Cv.y, cpz = pz - Cv.z;
	const float d1 = abx * apx + aby * apy + abz * apz, d2 = acx * apx + acy * apy + acz * apz;
	const float d3 = abx * bpx + aby * bpy + abz * bpz, d4 = acx * bpx + acy * bpy + acz * bpz;
	const float d5 = abx * cpx + aby * cpy + abz * cpz, d6 = acx * cpx + acy * cpy + acz * cpz;
	const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
	float v, w;                                          // q = A + v ab + w ac — a heuristic: the bounds below hold for any v, w
	if (d1 <= 0.f && d2 <= 0.f) { v = 0.f; w = 0.f; }
	else if (d3 >= 0.f && d4 <= d3) { v = 1.f; w = 0.f; }
	else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { v = __fdividef(d1, d1 - d3); w = 0.f; }
	else if (d6 >= 0.f && d5 <= d6) { v = 0.f; w = 1.f; }
	else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { v = 0.f; w = __fdividef(d2, d2 - d6); }
	else if (va <= 0.f && d4 - d3 >= 0.f && d5 - d6 >= 0.f) { w = __fdividef(d4 - d3, (d4 - d3) + (d5 - d6)); v = 1.f - w; }
	else { const float den = __fdividef(1.f, va + vb + vc); v = vb * den; w = vc * den; }
	v = fminf(fmaxf(v, 0.f), 1.f);                       // (fmaxf drops a NaN)
	w = fminf(fmaxf(w, 0.f), __fsub_rd(1.f, v));         // v + w <= 1 exactly: A + v ab + w ac is a point of the float triangle
	const float qx = fmaf(w, acx, fmaf(v, abx, A.x)), qy = fmaf(w, acy, fmaf(v, aby, A.y)), qz = fmaf(w, acz, fmaf(v, abz, A.z));
	const float nx = px - qx, ny = py - qy, nz = pz - qz;
	const float nn = __fmaf_ru(nz, nz, __fmaf_ru(ny, ny, __fmul_ru(nx, nx)));
	RefineOut r;
	// |p - x*| <= |p - q| + |q - x*|, x* the exact combination: slack_q bounds the rounding of q, (1 + 2^-22) that of n
	r.U = __fmaf_ru(__fsqrt_ru(nn), 1.0000005f, slack_q);
	// min over the vertices of n.(p - v), each rounded down; 2^-23 |n|.|p - v| covers the rounding of p - v
	const float anx = fabsf(nx), any = fabsf(ny), anz = fabsf(nz);
	const float sa = __fmaf_rd(__fmaf_ru(anz, fabsf(apz), __fmaf_ru(any, fabsf(apy), __fmul_ru(anx, fabsf(apx)))), -1.1920929e-7f,
	                           __fmaf_rd(nz, apz, __fmaf_rd(ny, apy, __fmul_rd(nx, apx))));
	const float sb = __fmaf_rd(__fmaf_ru(anz, fabsf(bpz), __fmaf_ru(any, fabsf(bpy), __fmul_ru(anx, fabsf(bpx)))), -1.1920929e-7f,
	                           __fmaf_rd(nz, bpz, __fmaf_rd(ny, bpy, __fmul_rd(nx, bpx))));
	const float sc = __fmaf_rd(__fmaf_ru(anz, fabsf(cpz), __fmaf_ru(any, fabsf(cpy), __fmul_ru(anx, fabsf(cpx)))), -1.1920929e-7f,
	                           __fmaf_rd(nz, cpz, __fmaf_rd(ny, cpy, __fmul_rd(nx, cpx))));
	const float sm = fminf(sa, fminf(sb, sc));
	r.L2 = sm > 0.f ? __fdiv_rd(__fmul_rd(sm, sm), nn) : 0.f;
	return r;
}
// threshold on squared float-side lower bounds from an upper bound ub of D(pf, T_f): (ub + 2 dq)^2, up, widened by 2^-19
__device__ __forceinline__ float threshold_from_upper(float ub, float dq) {
	const float x = __fadd_ru(ub, __fadd_ru(dq, dq));
	return __fmul_ru(__fmul_ru(x, x), 1.0000020f);
}

// per-warp shared state of cp_pair_kernel.  Everything is indexed as s.x[warp][...] on the __shared__ object itself: pointer
// variables into it decayed to generic addresses (ncu: an S2R SR_CgaCtaId + LEA window computation in front of every access).
struct PairShared {
	int32_t stack[4][WK_STACK];
	float4 cand[4][PK_CAND][2];                  // candidate: (lo.xyz, hi.x) (hi.y, hi.z, id bits, -)
	float qx[4][32], qy[4][32], qz[4][32], dq[4][32];
	unsigned thr[4][32];                         // float bits of the lane's threshold (>= 0, so unsigned order = float order)
	int32_t q1[4][PQ_CAP];                       // (slot << 8) | owner
	int32_t q2[4][PQ2_CAP];                      // (owner << 27) | facet
	uint4 stage[4][PB_SLOTS][16];                // clusters of the (lane, cluster) pairs in flight: 8 entries of two uint4
	int32_t slot_id[4][PB_SLOTS];
	int32_t sprim[4][SL_CAP][32];
	float sl2[4][SL_CAP][32];
	int scnt[4][32];
	int32_t tie[4][PK_TIES][32];
	double best[4][32];                          // running exact minimum (phase C only: kept out of the registers of the hot loops)
	float pbox[4][8];                            // packet box (phase A only)
	unsigned dead[4];                            // lanes whose survivor list overflowed: they leave the packet for the heavy kernel
};

template <bool STATS, int MINB>
__global__ void __launch_bounds__(128, MINB)
cp_pair_kernel(const WNode *__restrict__ wnodes, const float4 *__restrict__ trif, const double *__restrict__ tri,
               const double *__restrict__ P, int64_t np, const uint32_t *__restrict__ perm,
               double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C, SignArgs sa, float eps_v, float slack_q,
               int32_t *__restrict__ todo, int32_t *__restrict__ todo_ties, int32_t *__restrict__ todo_count, int32_t *__restrict__ heavy,
               int b3_budget, int c_budget, double *__restrict__ stats)
{
	__shared__ PairShared s;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1;
	const uint4 *wn4 = reinterpret_cast<const uint4 *>(wnodes);
	// Persistent warps: packet costs spread over two orders of magnitude (p50 85 k cycles, p99 455 k), and a CTA that owns four
	// fixed packets keeps its slots until the slowest one is done.  Every warp fetches its next packet from a global counter.
	for (;;) {
		int64_t w = 0;
		if (lane == 0) w = atomicAdd(todo_count + 4, 1);
		w = __shfl_sync(FULL, w, 0);
		if (w * 32 >= np) break;
		const int64_t slot = w * 32 + lane;
		const bool in_range = slot < np;
		const int64_t i = in_range ? (perm ? (int64_t)perm[slot] : slot) : 0;
		// (the fp64 query is re-read where igl's arithmetic needs it; between those places only its float image is live)
		float pfx, pfy, pfz, dq;
		bool valid;
		{
			const V3 p = ld3(P + 3 * i);
			valid = in_range && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
			pfx = (float)p.x; pfy = (float)p.y; pfz = (float)p.z;
			const double ex = p.x - (double)pfx, ey = p.y - (double)pfy, ez = p.z - (double)pfz;
			dq = __fadd_ru(__double2float_ru(sqrt(ex * ex + ey * ey + ez * ez) * 1.0000001), eps_v);
		}
		// (opaque to the optimiser: at 80 registers ptxas kept the fp64 query alive and re-converted it in front of every box test)
		asm volatile("" : "+f"(pfx), "+f"(pfy), "+f"(pfz), "+f"(dq));
		s.qx[warp][lane] = pfx; s.qy[warp][lane] = pfy; s.qz[warp][lane] = pfz; s.dq[warp][lane] = dq;
		s.scnt[warp][lane] = 0;
		if (lane == 0) s.dead[warp] = 0u;
		{
			const float qlx = ord2f(__reduce_min_sync(FULL, valid ? f2ord(pfx) : 0x7fffffff)), qhx = ord2f(__reduce_max_sync(FULL, valid ? f2ord(pfx) : (int)0x80000000));
			const float qly = ord2f(__reduce_min_sync(FULL, valid ? f2ord(pfy) : 0x7fffffff)), qhy = ord2f(__reduce_max_sync(FULL, valid ? f2ord(pfy) : (int)0x80000000));
			const float qlz = ord2f(__reduce_min_sync(FULL, valid ? f2ord(pfz) : 0x7fffffff)), qhz = ord2f(__reduce_max_sync(FULL, valid ? f2ord(pfz) : (int)0x80000000));
			if (lane == 0) { s.pbox[warp][0] = qlx; s.pbox[warp][1] = qly; s.pbox[warp][2] = qlz; s.pbox[warp][3] = qhx; s.pbox[warp][4] = qhy; s.pbox[warp][5] = qhz; }
		}
		s.best[warp][lane] = CUDART_INF;
		int32_t bf = -1;
		int near = 0, nt = 0;
		int st_steps = 0, st_cand = 0, st_p1 = 0, st_p2 = 0, st_iters = 0, st_evals = 0;
		long long t0 = 0;
		if (STATS) t0 = clock64();
		bool bail = false;
		// ---- seed: every lane walks down the wide tree for ITS OWN query, always into the child whose box is nearest
		// (ties: nearest box centre), and takes the fp32 upper bound to the facet it ends at as its first threshold.  One facet
		// for the whole packet left the far lanes with thresholds of the packet's diameter, and the candidate funnel below is
		// quadratic in that (counters on the C3 mesh: 264 candidate clusters per packet before, against ~40 facets it spans). ----
		{
			float thr0 = 0.f;
			if (valid) {
				int32_t nd = 0;
				thr0 = CUDART_INF_F;
				for (int guard = 0; guard < 48; ++guard) {
					float bk = CUDART_INF_F, bcd = CUDART_INF_F;
					int32_t bch = WCHILD_EMPTY;
#pragma unroll
					for (int c = 0; c < 8; ++c) {
						const uint4 a = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2), b = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2 + 1);
						const float lox = __uint_as_float(a.x), loy = __uint_as_float(a.y), loz = __uint_as_float(a.z);
						const float hix = __uint_as_float(a.w), hiy = __uint_as_float(b.x), hiz = __uint_as_float(b.y);
						const float gap = gap2_low(lox, loy, loz, hix, hiy, hiz, pfx, pfy, pfz);
						const float mx = 0.5f * lox + 0.5f * hix - pfx, my = 0.5f * loy + 0.5f * hiy - pfy, mz = 0.5f * loz + 0.5f * hiz - pfz;
						const float cd = mx * mx + my * my + mz * mz;
						if ((int32_t)b.z != WCHILD_EMPTY && (gap < bk || (gap == bk && cd < bcd))) { bk = gap; bcd = cd; bch = (int32_t)b.z; }
					}
					if (bch < 0) {
						if (bch != WCHILD_EMPTY) {
							const int64_t f = ~bch;
							const RefineOut r = refine_f32(pfx, pfy, pfz, __ldg(trif + 3 * f), __ldg(trif + 3 * f + 1), __ldg(trif + 3 * f + 2), slack_q);
							thr0 = threshold_from_upper(r.U, dq);
						}
						break;
					}
					nd = bch;
				}
			}
			s.thr[warp][lane] = __float_as_uint(thr0);
		}
		__syncwarp();
		if (__any_sync(FULL, valid)) {
			float T = __uint_as_float(__reduce_max_sync(FULL, valid ? s.thr[warp][lane] : 0u));
			int top = 1, ncand = 0, n1 = 0, n2 = 0, nslots = 0, b3_left = b3_budget, c_left = c_budget, cpass_left = 2 * c_budget;
			unsigned dead = 0u;
			if (lane == 0) s.stack[warp][0] = 0;
			__syncwarp();
			// Every stage below exists ONCE in the code (the first version inlined them at each call site: 9 200 instructions,
			// spills, instruction-cache misses); the loops are arranged so that one site serves every trigger.
			for (;;) {
				// ---- phase A: box-parallel expansion of the upper tree against the packet box and the largest threshold ----
				__syncwarp();                                      // lane 0's pushes in phase B1 are ordered against the reads below (racecheck)
				while (top > 0 && ncand <= PK_CAND - 32) {
					const int j = lane >> 3, c = lane & 7;
					const int take = top < 4 ? top : 4;
					const bool act = j < take;
					const int32_t nd = s.stack[warp][act ? top - 1 - j : 0];
					__syncwarp();
					top -= take;
					const uint4 a = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2), b = __ldg(wn4 + ((int64_t)nd * 8 + c) * 2 + 1);
					const float g = boxgap2_low(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z),
					                            __uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y),
					                            s.pbox[warp][0], s.pbox[warp][1], s.pbox[warp][2], s.pbox[warp][3], s.pbox[warp][4], s.pbox[warp][5]);
					const int32_t ch = (int32_t)b.z;
					// Everything that passes goes through the per-lane test of phase B1, inner nodes included: the packet-level
					// test (box against box, largest threshold) is loose when the packet is large or its lanes are at different
					// distances — far-away packets kept whole shells of the mesh alive (55 ns per far query before, 15 in the old walk).
					const bool to_cand = act && ch != WCHILD_EMPTY && g <= T;
					const unsigned mc = __ballot_sync(FULL, to_cand);
					if (to_cand) {
						const int k = ncand + __popc(mc & lt);
						s.cand[warp][k][0] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
						s.cand[warp][k][1] = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __int_as_float(ch), (ch >= 0 && !(b.w & 1u)) ? 1.f : 0.f);
					}
					ncand += __popc(mc);
					if (STATS) ++st_steps;
					__syncwarp();
				}
				if (bail) break;
				const bool last = ncand == 0;                      // the stack is empty too: one more pass drains the queues
				if (STATS) st_cand += ncand;
				for (int k = 0; k <= ncand; ++k) {
					const bool fin = k == ncand;
					// ---- phase B1: candidate box against every lane's own threshold -> (lane, cluster) pairs ----
					if (!fin) {
						const float4 ca = s.cand[warp][k][0], cb = s.cand[warp][k][1];
						const bool wk = valid && gap2_low(ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, pfx, pfy, pfz) <= __uint_as_float(s.thr[warp][lane]) && !((dead >> lane) & 1u);
						const unsigned m = __ballot_sync(FULL, wk);
						if (!m) continue;
						const int32_t id = __float_as_int(cb.z);
						if (cb.w != 0.f) {                         // an inner node some lane wants: back on the stack
							if (top >= WK_STACK) { bail = true; break; }
							if (lane == 0) s.stack[warp][top] = id;
							++top;
							continue;
						}
						if (id < 0) {                              // a stray facet: its box has just been tested
							if (wk) s.q2[warp][n2 + __popc(m & lt)] = (lane << 27) | ~id;
							n2 += __popc(m);
						} else {
							if (wk) s.q1[warp][n1 + __popc(m & lt)] = (nslots << 8) | lane;
							if (lane == 0) s.slot_id[warp][nslots] = id;
							n1 += __popc(m);
							++nslots;
						}
						__syncwarp();
						if (n1 < 32 && nslots < PB_SLOTS && n2 < 32) continue;
					}
					// ---- flush: bring the queued pairs' clusters into shared memory, every load in flight at once ----
					if (n1 > 0) {
						uint2 r[PB_SLOTS];
#pragma unroll
						for (int sl = 0; sl < PB_SLOTS; ++sl)
							if (sl < nslots) r[sl] = __ldg(reinterpret_cast<const uint2 *>(wnodes + s.slot_id[warp][sl]) + lane);
#pragma unroll
						for (int sl = 0; sl < PB_SLOTS; ++sl)
							if (sl < nslots) reinterpret_cast<uint2 *>(s.stage[warp][sl])[lane] = r[sl];
						__syncwarp();
					}
					for (;;) {
						// ---- phase B2: one (owner, cluster) pair per lane, the 8 facet boxes against the owner's threshold ----
						if (n1 > 0 && n2 < 64) {
							const int cnt = n1 < 32 ? n1 : 32;
							const bool has = lane < cnt;
							const int32_t e = s.q1[warp][has ? n1 - cnt + lane : 0];
							n1 -= cnt;
							const int owner = e & 31, sl = e >> 8;
							const float ox = s.qx[warp][owner], oy = s.qy[warp][owner], oz = s.qz[warp][owner];
							const float othr = (has && !((dead >> owner) & 1u)) ? __uint_as_float(s.thr[warp][owner]) : -1.f;   // an idle lane passes nothing (bounds are >= 0)
							const int32_t obits = owner << 27;
							if (STATS) st_p1 += cnt;
#pragma unroll
							for (int c = 0; c < 8; ++c) {
								const uint4 a = s.stage[warp][sl][2 * c], b = s.stage[warp][sl][2 * c + 1];
								// (an empty slot has box (+inf, -inf): its bound is +inf and never passes)
								const bool pass = gap2_low(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
								                           __uint_as_float(b.x), __uint_as_float(b.y), ox, oy, oz) <= othr;
								const unsigned m = __ballot_sync(FULL, pass);
								if (pass) s.q2[warp][n2 + __popc(m & lt)] = obits | ~(int32_t)b.z;
								n2 += __popc(m);
							}
							__syncwarp();
						}
						// ---- phase B3: one (owner, facet) pair per lane — fp32 refine, tighten the owner's threshold, keep survivors ----
						const bool more = n2 >= 32 || (fin && n1 == 0 && n2 > 0);
						if (more) {
							const int cnt = n2 < 32 ? n2 : 32;
							const bool has = lane < cnt;
							const int32_t ent = s.q2[warp][has ? n2 - cnt + lane : 0];
							n2 -= cnt;
							if (has && !((dead >> (int)((uint32_t)ent >> 27)) & 1u)) {
								const int owner = (int)((uint32_t)ent >> 27);
								const int32_t prim = ent & 0x07ffffff;
								const float4 A = __ldg(trif + 3 * (int64_t)prim), B = __ldg(trif + 3 * (int64_t)prim + 1), Cv = __ldg(trif + 3 * (int64_t)prim + 2);
								const RefineOut r = refine_f32(s.qx[warp][owner], s.qy[warp][owner], s.qz[warp][owner], A, B, Cv, slack_q);
								const float tu = threshold_from_upper(r.U, s.dq[warp][owner]);
								const float told = __uint_as_float(atomicMin(&s.thr[warp][owner], __float_as_uint(tu)));
								if (r.L2 <= fminf(told, tu)) {
									const int pos = atomicAdd(&s.scnt[warp][owner], 1);
									if (pos < SL_CAP) { s.sprim[warp][pos][owner] = prim; s.sl2[warp][pos][owner] = r.L2; }
									else atomicOr(&s.dead[warp], 1u << owner);
								}
							}
							if (STATS) st_p2 += cnt;
							__syncwarp();
							// More near-minimal facets than a list holds (a query on the axis of a bore sees thousands): that lane is
							// finished by the warp-per-query heavy kernel; here it stops wanting anything, so the packet's other lanes
							// are not held up (one such packet ran for 2.8 ms of a 4.4 ms launch).
							dead = s.dead[warp];
							if (--b3_left < 0) bail = true;
						}
						// ---- phase C: every lane walks its own survivor list, smallest lower bound first: igl's exact fp64 evaluation ----
						const bool idle = !more && n1 == 0;
						int mycnt = s.scnt[warp][lane];
						const bool final_c = idle && fin && last;
						if (!final_c && __any_sync(FULL, mycnt >= SL_CAP / 2)) {
							// a list that fills up mostly holds facets that passed an EARLIER threshold: drop what the current one
							// excludes before spending fp64 on it
							const float th = __uint_as_float(s.thr[warp][lane]);
							const int c0 = min(mycnt, SL_CAP);
							int c1 = 0;
							for (int j = 0; j < c0; ++j) {
								const float l = s.sl2[warp][j][lane];
								if (l <= th) { if (c1 != j) { s.sl2[warp][c1][lane] = l; s.sprim[warp][c1][lane] = s.sprim[warp][j][lane]; } ++c1; }
							}
							if (mycnt <= SL_CAP) { mycnt = c1; s.scnt[warp][lane] = c1; }
							__syncwarp();
						}
						if (__any_sync(FULL, mycnt >= SL_CAP / 2) || (final_c && __any_sync(FULL, mycnt > 0))) {
							const int cnt = min(mycnt, SL_CAP);
							if (cnt > 1) {
								int jm = 0; float lm = s.sl2[warp][0][lane];
								for (int j = 1; j < cnt; ++j) { const float l = s.sl2[warp][j][lane]; if (l < lm) { lm = l; jm = j; } }
								if (jm) {
									const int32_t tp = s.sprim[warp][0][lane]; s.sprim[warp][0][lane] = s.sprim[warp][jm][lane]; s.sprim[warp][jm][lane] = tp;
									s.sl2[warp][jm][lane] = s.sl2[warp][0][lane]; s.sl2[warp][0][lane] = lm;
								}
							}
							float mythr = __uint_as_float(s.thr[warp][lane]);
							const V3 p = ld3(P + 3 * i);
							double best = s.best[warp][lane];
							for (int j = 0; __any_sync(FULL, j < cnt); ++j) {
								if (j < cnt && s.sl2[warp][j][lane] <= mythr) {
									const int32_t prim = s.sprim[warp][j][lane];
									const double *t = tri + 9 * (int64_t)prim;
									const V3 q = closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6));
									const double d = sqnorm(sub(p, q));
									if (d < best) {
										near = (d + d * PK_EPS_TIE < best) ? 1 : near + 1;
										nt = 0; bf = prim;
										best = d;
										mythr = fminf(mythr, filter_threshold(d, dq));
									} else if (d <= best + best * PK_EPS_TIE) {
										++near;
										if (d == best) { if (nt < PK_TIES) s.tie[warp][nt][lane] = prim; ++nt; }
									}
									if (STATS) ++st_evals;
								}
								if (STATS) ++st_iters;
							}
							// a lane that has needed this many exact evaluations sits among a crowd of near-minimal facets (p99 is 11)
							c_left -= cnt;
							if (c_left < 0) atomicOr(&s.dead[warp], 1u << lane);
							if (--cpass_left < 0) bail = true;
							s.thr[warp][lane] = __float_as_uint(mythr);
							s.best[warp][lane] = best;
							s.scnt[warp][lane] = 0;
							__syncwarp();
							dead = s.dead[warp];
						}
						if (idle || bail) break;
					}
					nslots = 0;
					if (bail) break;
				}
				if (last || bail) break;
				ncand = 0;
				T = __uint_as_float(__reduce_max_sync(FULL, valid ? s.thr[warp][lane] : 0u));
			}
			if ((dead >> lane) & 1u) near = -1;
		}
		// ---- epilogue ----
		// a packet that ran out of budget holds queries with very many near-minimal facets: the per-lane search kernel would chew on
		// each of them for a millisecond (its leaf evaluations are not budgeted), the warp-per-query kernel takes them in its stride
		const int code = !valid ? 0 : ((near < 0 || bail) ? TODO_HEAVY : (near > 1 ? TODO_WALK : 0));
		if (in_range) {
			const V3 p = ld3(P + 3 * i);
			const double best = s.best[warp][lane];
			V3 bc = {0, 0, 0};
			if (bf >= 0) { const double *t = tri + 9 * (int64_t)bf; bc = closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6)); }
			I[i] = bf;
			C[3 * i] = bc.x; C[3 * i + 1] = bc.y; C[3 * i + 2] = bc.z;
			if (code == 0 && sa.enabled) finalize_sign(sa, i, p, bf, bc, best, S);
			else S[i] = best;
		}
		const unsigned mw = __ballot_sync(FULL, code == TODO_WALK), ms = __ballot_sync(FULL, code == TODO_SEARCH), mh = __ballot_sync(FULL, code == TODO_HEAVY);
		if (mw | ms | mh) {
			int bw = 0, bs = 0, bh = 0;
			if (lane == 0) {
				if (mw) bw = atomicAdd(todo_count, __popc(mw));
				if (ms) bs = atomicAdd(todo_count + 2, __popc(ms));
				if (mh) bh = atomicAdd(todo_count + 1, __popc(mh));
			}
			bw = __shfl_sync(FULL, bw, 0); bs = __shfl_sync(FULL, bs, 0); bh = __shfl_sync(FULL, bh, 0);
			if (code == TODO_WALK) {
				const int sl = bw + __popc(mw & lt);
				todo[sl] = (int32_t)i;
				int32_t *tr = todo_ties + TT_STRIDE * (int64_t)sl;
				tr[0] = nt;
				for (int j = 0; j < PK_TIES; ++j) tr[1 + j] = j < nt ? s.tie[warp][j][lane] : -1;
			} else if (code == TODO_SEARCH) {
				todo[2 * np - 1 - (bs + __popc(ms & lt))] = (int32_t)i;
			} else if (code == TODO_HEAVY) {
				heavy[bh + __popc(mh & lt)] = (int32_t)i;
			}
		}
		if (STATS && stats && in_range) {
			double *o = stats + 3 * i;
			o[0] = st_steps + 65536.0 * st_cand + 4294967296.0 * code;
			o[1] = st_p1 + 65536.0 * st_iters + 4294967296.0 * st_evals;
			o[2] = st_p2 + 1048576.0 * (double)((clock64() - t0) >> 4);
		}
		__syncwarp();
	}
}


// igl-order walk (as `traverse`) that additionally skips every box farther than `limit`, stops as soon as it holds a
// facet at distance `dmin` (the exact minimum over all facets) and gives up after `budget` node visits (returns false).
__device__ __forceinline__ bool traverse_limited(const QNode *__restrict__ nodes, int32_t root, const double *__restrict__ tri,
                                                 const V3 &p, double limit, double dmin, int budget, Hit &h)
{
	h.sqr_d = CUDART_INF; h.f = -1; h.c = {0, 0, 0};
	int32_t st_node[FPOHM_STACK];
	double st_d[FPOHM_STACK];
	int sp = 0;
	int32_t cur = root;
	for (;;) {
		if (--budget < 0) return false;
		const QNode *n = nodes + cur;
		const double dl = box_ext_sqdist(n->lmin, n->lmax, p);
		const double dr = box_ext_sqdist(n->rmin, n->rmax, p);
		const bool in_l = box_contains(n->lmin, n->lmax, p);
		const bool left_first = in_l || dl < dr;
		const int32_t c1 = left_first ? n->left : n->right, c2 = left_first ? n->right : n->left;
		const double d1 = left_first ? dl : dr, d2 = left_first ? dr : dl;
		if (d2 < h.sqr_d && d2 <= limit && sp < FPOHM_STACK) { st_node[sp] = c2; st_d[sp] = d2; ++sp; }
		int32_t next = -1;
		bool have_next = false;
		if (d1 < h.sqr_d && d1 <= limit) {
			if (c1 < 0) { test_leaf(tri, ~c1, p, h); if (h.sqr_d == dmin) return true; } else { next = c1; have_next = true; }
		}
		while (!have_next && sp > 0) {
			--sp;
			if (st_d[sp] < h.sqr_d) {
				const int32_t c = st_node[sp];
				if (c < 0) { test_leaf(tri, ~c, p, h); if (h.sqr_d == dmin) return true; } else { next = c; have_next = true; }
			}
		}
		if (!have_next) break;
		cur = next;
	}
	return true;
}

// ---- K2 (tie-break) -----------------------------------------------------------------------------------------
// One thread per query whose minimum is exact but shared (nearly) by several facets: name igl's winner from the exact-tie
// list when one box test proves it (`igl_tie_winner`), else re-walk in igl order; a walk that exceeds its budget goes on
// to the heavy kernel.
__global__ void __launch_bounds__(128)
cp_tie_kernel(const QNode *__restrict__ nodes, int32_t root, const double *__restrict__ tri, const int32_t *__restrict__ prim_parent,
              const double *__restrict__ P, const int32_t *__restrict__ todo, const int32_t *__restrict__ todo_ties,
              int32_t *__restrict__ counters, double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C,
              int32_t *__restrict__ heavy, int walk_budget, SignArgs sa, const int2 *__restrict__ pd, bool count_walks = false)
{
	const int n_todo = counters[0];
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_todo; t += gridDim.x * blockDim.x) {
		const int64_t i = todo[t];
		const V3 p = ld3(P + 3 * i);
		const double best = S[i];
		const int32_t bf = I[i];
		const int nt = todo_ties[TT_STRIDE * (int64_t)t];
		int32_t ties[PK_TIES];
		for (int j = 0; j < PK_TIES; ++j) ties[j] = todo_ties[TT_STRIDE * (int64_t)t + 1 + j];
		int32_t win = -1;
		if (nt <= PK_TIES) win = igl_tie_winner(nodes, prim_parent, tri, p, best, bf, ties, 1, nt, pd);
		if (win >= 0) {
			V3 c;
			if (win != bf) {
				const double *tt = tri + 9 * (int64_t)win;
				c = closest_on_triangle(p, ld3(tt), ld3(tt + 3), ld3(tt + 6));
				I[i] = win;
				C[3 * i] = c.x; C[3 * i + 1] = c.y; C[3 * i + 2] = c.z;
			} else if (sa.enabled) c = ld3(C + 3 * i);
			finalize_sign(sa, i, p, win, c, best, S);
		} else {
			Hit h;
			if (count_walks) atomicAdd(counters + 5, 1);
			if (traverse_limited(nodes, root, tri, p, best + best * PK_EPS_WALK, best, walk_budget, h)) {
				I[i] = h.f;
				C[3 * i] = h.c.x; C[3 * i + 1] = h.c.y; C[3 * i + 2] = h.c.z;
				S[i] = h.sqr_d;
				finalize_sign(sa, i, p, h.f, h.c, h.sqr_d, S);
			} else {
				if (count_walks) atomicAdd(counters + 6, 1);
				heavy[atomicAdd(counters + 1, 1)] = (int32_t)i;      // S[i] stays a valid upper bound for K3
			}
		}
	}
}

// ---- K2 (search) --------------------------------------------------------------------------------------------
// Unfinished searches have very uneven lengths (tens to thousands of node visits), and a thread per query makes every
// warp wait for its longest lane.  Persistent lanes instead: every loop iteration is ONE node visit for every lane that
// has a query, and a lane that finishes fetches the next query from the list at once (warp-aggregated counter), so the
// warp's lanes stay busy and its time follows the mean search length, not the maximum.
__global__ void __launch_bounds__(128, 5)
cp_search_kernel(const QNodeF *__restrict__ fnodes, int32_t root, const double *__restrict__ tri,
                 const double *__restrict__ P, int64_t np, int32_t *__restrict__ todo, int32_t *__restrict__ todo_ties,
                 int32_t *__restrict__ counters /* [0] walk entries, [1] heavy, [2] search entries, [3] next search entry */,
                 double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C,
                 int32_t *__restrict__ heavy, int search_budget, SignArgs sa)
{
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1;
	const int n_todo = counters[2];
	Packet k;
	int32_t ties[PK_TIES];
	int32_t lst[K2_STACK];
	int sp = 0, budget = 0;
	int64_t i = -1;
	int32_t bf0 = -1;
	float plx = 0, ply = 0, plz = 0, phx = 0, phy = 0, phz = 0;
	bool active = false, exhausted = false;
	k.p = {0, 0, 0}; k.best = 0; k.best_hi = 0; k.bf = -1; k.bc = {0, 0, 0}; k.near = 0; k.nt = 0;
	for (;;) {
		// ---- fetch ----
		const unsigned need = __ballot_sync(0xffffffffu, !active && !exhausted);
		if (need) {
			int base = 0;
			if (lane == __ffs(need) - 1) base = atomicAdd(counters + 3, __popc(need));
			base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
			if (!active && !exhausted) {
				const int t = base + __popc(need & lt);
				if (t < n_todo) {
					i = todo[2 * np - 1 - t];
					k.p = ld3(P + 3 * i);
					bf0 = I[i];
					k.best = S[i];
					if (bf0 >= 0) k.best = k.best + k.best * PK_EPS_TIE * 4;   // still an upper bound; lets bf0 itself pass 'd < best'
					k.best_hi = __double2float_ru(k.best);
					k.bf = -1; k.near = 0; k.nt = 0;
					plx = __double2float_rd(k.p.x); ply = __double2float_rd(k.p.y); plz = __double2float_rd(k.p.z);
					phx = __double2float_ru(k.p.x); phy = __double2float_ru(k.p.y); phz = __double2float_ru(k.p.z);
					sp = 0; lst[sp++] = root; budget = search_budget;
					active = true;
				} else {
					exhausted = true;
				}
			}
		}
		if (!__any_sync(0xffffffffu, active)) break;
		// ---- one node visit ----
		bool give_up = false;
		if (active && sp > 0) {
			const QNodeF *n = fnodes + lst[--sp];
			const float dl = box_low(n->lmin, n->lmax, plx, ply, plz, phx, phy, phz);
			const float dr = box_low(n->rmin, n->rmax, plx, ply, plz, phx, phy, phz);
			if (dl <= k.best_hi || dr <= k.best_hi) {
				if (--budget < 0 || sp + 2 > K2_STACK) give_up = true;
				else {
					const bool left_first = dl < dr || dl == 0.f;
					const int32_t c1 = left_first ? n->left : n->right, c2 = left_first ? n->right : n->left;
					const float d1 = left_first ? dl : dr, d2 = left_first ? dr : dl;
					if (c1 < 0) {
						if (d1 <= k.best_hi) lane_leaf(tri, ~c1, k, ties);
						if (d2 <= k.best_hi) { if (c2 < 0) lane_leaf(tri, ~c2, k, ties); else lst[sp++] = c2; }
					} else {
						if (d2 <= k.best_hi) { if (c2 < 0) lane_leaf(tri, ~c2, k, ties); else lst[sp++] = c2; }
						if (d1 <= k.best_hi) lst[sp++] = c1;
					}
				}
			}
		}
		// ---- retire ----
		const bool finished = active && (give_up || sp == 0);
		const unsigned mh = __ballot_sync(0xffffffffu, finished && give_up);
		const unsigned mw = __ballot_sync(0xffffffffu, finished && !give_up && k.near > 1);
		if (mh | mw) {
			int bh = 0, bw = 0;
			if (lane == 0) { if (mh) bh = atomicAdd(counters + 1, __popc(mh)); if (mw) bw = atomicAdd(counters, __popc(mw)); }
			bh = __shfl_sync(0xffffffffu, bh, 0); bw = __shfl_sync(0xffffffffu, bw, 0);
			if (finished && give_up) heavy[bh + __popc(mh & lt)] = (int32_t)i;
			else if (finished && k.near > 1) {
				const int slot = bw + __popc(mw & lt);          // the walk list grows from the front, we consume from the back
				todo[slot] = (int32_t)i;
				int32_t *tr = todo_ties + TT_STRIDE * (int64_t)slot;
				tr[0] = k.nt;
				for (int j = 0; j < PK_TIES; ++j) tr[1 + j] = j < k.nt ? ties[j] : -1;
			}
		}
		if (finished) {
			if (k.bf >= 0) {                                 // (a give-up before any facet was met keeps K1's upper bound)
				I[i] = k.bf;
				C[3 * i] = k.bc.x; C[3 * i + 1] = k.bc.y; C[3 * i + 2] = k.bc.z;
				S[i] = k.best;
			}
			if (sa.enabled && !give_up && k.near <= 1) {     // settled here: nobody else will look at this query again
				if (k.bf >= 0) finalize_sign(sa, i, k.p, k.bf, k.bc, k.best, S);
				else finalize_sign(sa, i, k.p, bf0, ld3(C + 3 * i), S[i], S);
			}
			active = false;
		}
	}
}

// ---- K3 ---------------------------------------------------------------------------------------------------------
// One WARP per query, 32 tree nodes per step: the form for LONG lists (at least one query per resident warp).
__global__ void __launch_bounds__(128)
cp_heavy_warp_kernel(const QNode *__restrict__ nodes, int32_t root, const double *__restrict__ tri, const int32_t *__restrict__ prim_parent,
                const double *__restrict__ P, const int32_t *__restrict__ heavy, const int32_t *__restrict__ heavy_count,
                double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C, SignArgs sa, int min_count)
{
	__shared__ int32_t s_node[4][HV_STACK];
	__shared__ unsigned long long s_key[4][HV_STACK];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int32_t *stk = s_node[warp];
	unsigned long long *stkk = s_key[warp];
	const int n_heavy = *heavy_count;
	if (n_heavy < min_count) return;                   // short lists: cp_heavy_kernel (one CTA per query)
	const unsigned lt = (1u << lane) - 1;
	for (int hq = blockIdx.x * 4 + warp; hq < n_heavy; hq += gridDim.x * 4) {
		const int64_t i = heavy[hq];
		const V3 p = ld3(P + 3 * i);
		double best = S[i];                               // exact distance of some facet (or +inf): a valid upper bound
		unsigned long long bkey = ~0ull;
		int32_t bfac = -1;
		bool deep = false;                                // a rank did not fit 62 bits
		int top = 1;
		if (lane == 0) { stk[0] = root; stkk[0] = 0ull; }
		__syncwarp();
		while (top > 0) {
			int take = top < 32 ? top : 32;
			if (top > HV_STACK - 64) take = 1;              // nearly full: pure depth-first, growth bounded by the tree depth
			int32_t cl = 0, cr = 0;
			unsigned long long kl = 0, kr = 0;
			bool wl = false, wr = false;
			const double lim = best + best * PK_EPS_TIE;
			if (lane < take) {
				const QNode *n = nodes + stk[top - 1 - lane];
				const unsigned long long key = stkk[top - 1 - lane];
				const double dl = box_ext_sqdist(n->lmin, n->lmax, p), dr = box_ext_sqdist(n->rmin, n->rmax, p);
				const bool left_first = box_contains(n->lmin, n->lmax, p) || dl < dr;
				const int depth = n->depth;
				if (depth > 61) deep = true;
				const unsigned long long bit = depth > 61 ? 0ull : 1ull << (61 - depth);
				kl = left_first ? key : key | bit; kr = left_first ? key | bit : key;
				wl = dl <= lim; wr = dr <= lim;
				cl = n->left; cr = n->right;
			}
			__syncwarp();
			top -= take;
			// leaves: exact distance; candidate ordered by (distance, igl rank)
			double d = CUDART_INF; unsigned long long dk = ~0ull; int32_t df = -1;
			if (wl && cl < 0) { const double *t = tri + 9 * (int64_t)~cl; d = sqnorm(sub(p, closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6)))); dk = kl; df = ~cl; }
			if (wr && cr < 0) {
				const double *t = tri + 9 * (int64_t)~cr;
				const double d2 = sqnorm(sub(p, closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6))));
				if (d2 < d || (d2 == d && kr < dk)) { d = d2; dk = kr; df = ~cr; }
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const double od = __shfl_xor_sync(0xffffffffu, d, o);
				const unsigned long long ok = __shfl_xor_sync(0xffffffffu, dk, o);
				const int32_t of = __shfl_xor_sync(0xffffffffu, df, o);
				if (od < d || (od == d && ok < dk)) { d = od; dk = ok; df = of; }
			}
			if (df >= 0 && (d < best || (d == best && dk < bkey))) { best = d; bkey = dk; bfac = df; }
			const bool pl = wl && cl >= 0, pr = wr && cr >= 0;
			const unsigned mpl = __ballot_sync(0xffffffffu, pl), mpr = __ballot_sync(0xffffffffu, pr);
			if (pl) { const int q = top + __popc(mpl & lt); stk[q] = cl; stkk[q] = kl; }
			if (pr) { const int q = top + __popc(mpl) + __popc(mpr & lt); stk[q] = cr; stkk[q] = kr; }
			top += __popc(mpl) + __popc(mpr);
			__syncwarp();
		}
		deep = __any_sync(0xffffffffu, deep);
		if (lane == 0) {
			// igl reaches the first-ranked exact-minimum leaf if every box on its path is at most `best` away
			bool ok = !deep && bfac >= 0;
			if (ok) {
				int32_t child = ~bfac, nd = prim_parent[bfac];
				while (nd >= 0 && ok) {
					const QNode *n = nodes + nd;
					const double bd = n->left == child ? box_ext_sqdist(n->lmin, n->lmax, p) : box_ext_sqdist(n->rmin, n->rmax, p);
					ok = bd <= best;
					child = nd; nd = n->parent;
				}
			}
			Hit h;
			if (ok) {
				const double *t = tri + 9 * (int64_t)bfac;
				h.f = bfac; h.sqr_d = best; h.c = closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6));
			} else {
				traverse_limited(nodes, root, tri, p, best + best * PK_EPS_WALK, best, 0x7fffffff, h);
			}
			if (I) I[i] = h.f;
			if (C) { C[3 * i] = h.c.x; C[3 * i + 1] = h.c.y; C[3 * i + 2] = h.c.z; }
			if (S) S[i] = h.sqr_d;
			finalize_sign(sa, i, p, h.f, h.c, h.sqr_d, S);
		}
		__syncwarp();
	}
}

__global__ void __launch_bounds__(128)
cp_heavy_kernel(const QNode *__restrict__ nodes, int32_t root, const double *__restrict__ tri, const int32_t *__restrict__ prim_parent,
                const double *__restrict__ P, const int32_t *__restrict__ heavy, const int32_t *__restrict__ heavy_count,
                double *__restrict__ S, int32_t *__restrict__ I, double *__restrict__ C, SignArgs sa, int max_count)
{
	// One CTA (4 warps) per query, 128 tree nodes per step: the form for SHORT lists (fewer queries than resident warps).  (One WARP per query until late in round 2: a launch could not be shorter
	// than its longest query, ~0.36 ms — 14 launches per host-pointer C4 step, two per rank and step under strong scaling.)  The
	// winner is the minimum over (distance, igl rank) of every leaf within the shrinking bound, so the visiting order is free.
	constexpr int CAP = 4 * HV_STACK;
	__shared__ int32_t stk[CAP];
	__shared__ unsigned long long stkk[CAP];
	__shared__ double r_d[4];
	__shared__ unsigned long long r_k[4];
	__shared__ int32_t r_f[4];
	__shared__ int s_cnt[4][2];
	__shared__ int s_deep;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n_heavy = *heavy_count;
	if (n_heavy >= max_count) return;                  // long lists: cp_heavy_warp_kernel
	const unsigned lt = (1u << lane) - 1;
	for (int hq = blockIdx.x; hq < n_heavy; hq += gridDim.x) {
		const int64_t i = heavy[hq];
		const V3 p = ld3(P + 3 * i);
		double best = S[i];                               // exact distance of some facet (or +inf): a valid upper bound
		unsigned long long bkey = ~0ull;
		int32_t bfac = -1;
		int top = 1;
		if (tid == 0) { stk[0] = root; stkk[0] = 0ull; s_deep = 0; }
		__syncthreads();
		while (top > 0) {
			int take = top < 128 ? top : 128;
			if (top > CAP - 320) take = 1;                  // nearly full: pure depth-first, growth bounded by the tree depth
			int32_t cl = 0, cr = 0;
			unsigned long long kl = 0, kr = 0;
			bool wl = false, wr = false;
			const double lim = best + best * PK_EPS_TIE;
			if (tid < take) {
				const QNode *n = nodes + stk[top - 1 - tid];
				const unsigned long long key = stkk[top - 1 - tid];
				const double dl = box_ext_sqdist(n->lmin, n->lmax, p), dr = box_ext_sqdist(n->rmin, n->rmax, p);
				const bool left_first = box_contains(n->lmin, n->lmax, p) || dl < dr;
				const int depth = n->depth;
				if (depth > 61) s_deep = 1;                  // a rank does not fit 62 bits (benign race: every writer stores 1)
				const unsigned long long bit = depth > 61 ? 0ull : 1ull << (61 - depth);
				kl = left_first ? key : key | bit; kr = left_first ? key | bit : key;
				wl = dl <= lim; wr = dr <= lim;
				cl = n->left; cr = n->right;
			}
			// leaves: exact distance; candidate ordered by (distance, igl rank)
			double d = CUDART_INF; unsigned long long dk = ~0ull; int32_t df = -1;
			if (wl && cl < 0) { const double *t = tri + 9 * (int64_t)~cl; d = sqnorm(sub(p, closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6)))); dk = kl; df = ~cl; }
			if (wr && cr < 0) {
				const double *t = tri + 9 * (int64_t)~cr;
				const double d2 = sqnorm(sub(p, closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6))));
				if (d2 < d || (d2 == d && kr < dk)) { d = d2; dk = kr; df = ~cr; }
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const double od = __shfl_xor_sync(0xffffffffu, d, o);
				const unsigned long long ok = __shfl_xor_sync(0xffffffffu, dk, o);
				const int32_t of = __shfl_xor_sync(0xffffffffu, df, o);
				if (od < d || (od == d && ok < dk)) { d = od; dk = ok; df = of; }
			}
			const bool pl = wl && cl >= 0, pr = wr && cr >= 0;
			const unsigned mpl = __ballot_sync(0xffffffffu, pl), mpr = __ballot_sync(0xffffffffu, pr);
			__syncthreads();                                 // every pop of this step has been read: the slots may be overwritten
			if (lane == 0) { r_d[warp] = d; r_k[warp] = dk; r_f[warp] = df; s_cnt[warp][0] = __popc(mpl); s_cnt[warp][1] = __popc(mpr); }
			__syncthreads();
			top -= take;
#pragma unroll
			for (int w = 0; w < 4; ++w) {
				const double od = r_d[w]; const unsigned long long ok = r_k[w]; const int32_t of = r_f[w];
				if (of >= 0 && (od < best || (od == best && ok < bkey))) { best = od; bkey = ok; bfac = of; }
			}
			int base_l = top, total_l = 0, before_r = 0, total_r = 0;
#pragma unroll
			for (int w = 0; w < 4; ++w) {
				if (w < warp) { base_l += s_cnt[w][0]; before_r += s_cnt[w][1]; }
				total_l += s_cnt[w][0]; total_r += s_cnt[w][1];
			}
			if (pl) { const int q = base_l + __popc(mpl & lt); stk[q] = cl; stkk[q] = kl; }
			if (pr) { const int q = top + total_l + before_r + __popc(mpr & lt); stk[q] = cr; stkk[q] = kr; }
			top += total_l + total_r;
			__syncthreads();
		}
		if (tid == 0) {
			// igl reaches the first-ranked exact-minimum leaf if every box on its path is at most `best` away
			bool ok = !s_deep && bfac >= 0;
			if (ok) {
				int32_t child = ~bfac, nd = prim_parent[bfac];
				while (nd >= 0 && ok) {
					const QNode *n = nodes + nd;
					const double bd = n->left == child ? box_ext_sqdist(n->lmin, n->lmax, p) : box_ext_sqdist(n->rmin, n->rmax, p);
					ok = bd <= best;
					child = nd; nd = n->parent;
				}
			}
			Hit h;
			if (ok) {
				const double *t = tri + 9 * (int64_t)bfac;
				h.f = bfac; h.sqr_d = best; h.c = closest_on_triangle(p, ld3(t), ld3(t + 3), ld3(t + 6));
			} else {
				traverse_limited(nodes, root, tri, p, best + best * PK_EPS_WALK, best, 0x7fffffff, h);
			}
			if (I) I[i] = h.f;
			if (C) { C[3 * i] = h.c.x; C[3 * i + 1] = h.c.y; C[3 * i + 2] = h.c.z; }
			if (S) S[i] = h.sqr_d;
			finalize_sign(sa, i, p, h.f, h.c, h.sqr_d, S);
		}
		__syncthreads();
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
		if (I[i] < 0) {                                     // no facet: NaN query coordinates (igl leaves its outputs untouched)
			if (S) S[i] = CUDART_NAN;
			if (N && keep_n) { N[3 * i] = 0; N[3 * i + 1] = 0; N[3 * i + 2] = 0; }
			continue;
		}
		const double s = pseudonormal(V, F, nF, FN, VN, EN, EMAP, p, I[i], c, n);
		if (S) S[i] = s * sqrt(S[i]);                       // S held the squared distance
		if (N && keep_n) { N[3 * i] = n.x; N[3 * i + 1] = n.y; N[3 * i + 2] = n.z; }
	}
}

// clean_hex_mesh, ghm.cpp:1937-1951: per hex the centre of the bounding box of its 8 corners, then points_inside_mesh
__global__ void hex_box_centres_kernel(const double *__restrict__ V, const uint32_t *__restrict__ hex, int64_t H, double *__restrict__ P) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 3 * H; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t h = t / 3; const int c = (int)(t % 3);
		double mn = V[3 * (int64_t)hex[8 * h] + c], mx = mn;
		for (int k = 1; k < 8; ++k) { const double x = V[3 * (int64_t)hex[8 * h + k] + c]; mn = fmin(mn, x); mx = fmax(mx, x); }
		P[t] = (mx + mn) / 2;
	}
}
__global__ void inside_flags_kernel(const double *__restrict__ S, int64_t n, uint8_t *__restrict__ flag) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) flag[t] = S[t] < 0 ? 1 : 0;
}

// ---- incoherent batches: Morton order for the packets, original order for the caller ---------------------------------
__global__ void query_keys_kernel(const double *__restrict__ P, int64_t np, double ox, double oy, double oz, double scale,
                                  uint32_t *__restrict__ key, uint32_t *__restrict__ idx)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
		const double fx = (P[3 * i] - ox) * scale, fy = (P[3 * i + 1] - oy) * scale, fz = (P[3 * i + 2] - oz) * scale;
		const uint32_t x = (uint32_t)fmin(fmax(fx, 0.0), 1023.0), y = (uint32_t)fmin(fmax(fy, 0.0), 1023.0), z = (uint32_t)fmin(fmax(fz, 0.0), 1023.0);
		key[i] = (uint32_t)morton3(x, y, z);
		idx[i] = (uint32_t)i;
	}
}
// device-resident batches: bounding box of the finite queries (ordered-int min/max), then the same 30-bit keys
__global__ void query_bbox_kernel(const double *__restrict__ P, int64_t np, int *__restrict__ box /* 6 ordered ints, pre-set */) {
	int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
		const double x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
		if (isfinite(x) && isfinite(y) && isfinite(z)) {
			lo[0] = min(lo[0], f2ord(__double2float_rd(x))); hi[0] = max(hi[0], f2ord(__double2float_ru(x)));
			lo[1] = min(lo[1], f2ord(__double2float_rd(y))); hi[1] = max(hi[1], f2ord(__double2float_ru(y)));
			lo[2] = min(lo[2], f2ord(__double2float_rd(z))); hi[2] = max(hi[2], f2ord(__double2float_ru(z)));
		}
	}
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		lo[c] = __reduce_min_sync(0xffffffffu, lo[c]); hi[c] = __reduce_max_sync(0xffffffffu, hi[c]);
	}
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int c = 0; c < 3; ++c) { atomicMin(box + c, lo[c]); atomicMax(box + 3 + c, hi[c]); }
	}
}
// Is the batch already coherent — do 32 CONSECUTIVE queries (one packet) sit in a compact clump?  4 096 sampled packets: the
// squared diagonal of the box of four of their members (first, last and two in between) against the diagonal a clump of 32
// points would have if the np points were spread evenly over the batch's box.  box[6] := 1 if NOT coherent.  Neighbouring
// queries being close is not enough: a lattice row (z-fastest ids) or a strip of facets makes packets 30 cells long, and the
// candidate funnel of the packet kernel grows with the packet's box.  Packets cut from an already ordered batch (octree
// leaves) are left alone: they are aligned with the caller's hierarchy, while windows of a Morton curve straddle its jumps
// (measured on octree-ordered leaf centres: +37 % candidate clusters after re-sorting).
__global__ void query_probe_kernel(const double *__restrict__ P, int64_t np, int *__restrict__ box, int force) {
	__shared__ double s_sum[32];
	const int64_t m = 4096, npk = np / 32, stride = npk > m ? npk / m : 1;
	double acc = 0;
	int cnt = 0;
	for (int64_t k = threadIdx.x; k < m && k * stride < npk; k += blockDim.x) {
		const int64_t base = k * stride * 32;
		double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
		const int pick[4] = {0, 10, 21, 31};
		for (int j = 0; j < 4; ++j) {
			const double *a = P + 3 * (base + pick[j]);
			for (int c = 0; c < 3; ++c) if (isfinite(a[c])) { lo[c] = fmin(lo[c], a[c]); hi[c] = fmax(hi[c], a[c]); }
		}
		double d2 = 0;
		for (int c = 0; c < 3; ++c) if (hi[c] >= lo[c]) d2 += (hi[c] - lo[c]) * (hi[c] - lo[c]);
		acc += d2; ++cnt;
	}
	for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
	__shared__ int s_cnt[32];
	if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_cnt[threadIdx.x >> 5] = cnt; }
	__syncthreads();
	if (threadIdx.x == 0) {
		double t = 0; int n = 0;
		for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t += s_sum[k]; n += s_cnt[k]; }
		t /= (double)(n > 0 ? n : 1);
		const double ex = (double)ord2f(box[3]) - (double)ord2f(box[0]), ey = (double)ord2f(box[4]) - (double)ord2f(box[1]), ez = (double)ord2f(box[5]) - (double)ord2f(box[2]);
		const double em = fmax(ex, fmax(ey, ez));
		// volume of the box, flat directions counted as 1 % of the longest one
		const double vol = fmax(ex, 0.01 * em) * fmax(ey, 0.01 * em) * fmax(ez, 0.01 * em);
		const double clump2 = 3.0 * pow(32.0 * vol / (double)np, 2.0 / 3.0);      // squared diagonal of a cube holding 32 evenly spread points
		box[6] = force ? 1 : ((em > 0 && n > 0 && t > 16.0 * clump2) ? 1 : 0);
	}
}
__global__ void query_keys_dev_kernel(const double *__restrict__ P, int64_t np, const int *__restrict__ box,
                                      uint32_t *__restrict__ key, uint32_t *__restrict__ idx)
{
	if (box[6] == 0) {      // coherent batch: keys in the caller's order (the stable sort then returns the identity)
		for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) { key[i] = (uint32_t)(i >> 2); idx[i] = (uint32_t)i; }
		return;
	}
	const float lx = ord2f(box[0]), ly = ord2f(box[1]), lz = ord2f(box[2]);
	const float ext = fmaxf(fmaxf(ord2f(box[3]) - lx, ord2f(box[4]) - ly), ord2f(box[5]) - lz);
	const double scale = ext > 0.f ? 1024.0 / (double)ext : 0.0;
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
		const double fx = (P[3 * i] - lx) * scale, fy = (P[3 * i + 1] - ly) * scale, fz = (P[3 * i + 2] - lz) * scale;
		// (NaN compares false: a non-finite query lands in cell 0, it never searches anyway)
		const uint32_t x = fx > 0.0 ? (uint32_t)fmin(fx, 1023.0) : 0u, y = fy > 0.0 ? (uint32_t)fmin(fy, 1023.0) : 0u, z = fz > 0.0 ? (uint32_t)fmin(fz, 1023.0) : 0u;
		key[i] = (uint32_t)morton3(x, y, z);
		idx[i] = (uint32_t)i;
	}
}
__global__ void gather_points_kernel(const double *__restrict__ P, const uint32_t *__restrict__ perm, int64_t np, double *__restrict__ out) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 3 * np; t += (int64_t)gridDim.x * blockDim.x)
		out[t] = P[3 * (int64_t)perm[t / 3] + t % 3];
}
__global__ void scatter_results_kernel(const uint32_t *__restrict__ perm, int64_t np, const double *__restrict__ S, const int32_t *__restrict__ I,
                                       const double *__restrict__ C, const double *__restrict__ N, double *__restrict__ So, int32_t *__restrict__ Io,
                                       double *__restrict__ Co, double *__restrict__ No)
{
	for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < np; j += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = perm[j];
		if (So) So[i] = S[j];
		if (Io) Io[i] = I[j];
		if (Co) { Co[3 * i] = C[3 * j]; Co[3 * i + 1] = C[3 * j + 1]; Co[3 * i + 2] = C[3 * j + 2]; }
		if (No) { No[3 * i] = N[3 * j]; No[3 * i + 1] = N[3 * j + 1]; No[3 * i + 2] = N[3 * j + 2]; }
	}
}

} // namespace

namespace fpohm {

// Scratch of one query launch (stream-ordered allocations, released when the launch has been queued).
struct QueryScratch {
	DevBuf<int32_t> tmpI, todo, todo_ties, heavy, cnt;
	DevBuf<double> tmpC, tmpS;
	DevBuf<uint32_t> key, key2, idx, perm;
	DevBuf<uint8_t> sort_tmp;
	DevBuf<int> qbox;
};
enum { CP_SORT_AUTO = 0, CP_SORT_NEVER = 1, CP_SORT_ALWAYS = 2 };
static int cp_mode() {
	static const int mode = getenv("FPOHM_CP_MODE") ? atoi(getenv("FPOHM_CP_MODE")) : 3;   // debug A/B: 0 per-lane igl order, 1 binary packets, 2 wide packets, 3 pair queues
	return mode;
}

static void launch_closest_point_ex(fpohm_ctx *ctx, fpohm_mesh *m, bool with_sign, const double *P_dev, int64_t np,
                                    double *S, int32_t *I, double *C, double *N, cudaStream_t s, QueryScratch &q, int sort_policy = CP_SORT_AUTO)
{
	if (np <= 0) return;
	const int mode = cp_mode();
	if (mode >= 2 && m->qroot >= 0 && m->n_wnodes > 0) {
		static const bool stats = getenv("FPOHM_CP_STATS") != nullptr;   // debug only: N := traversal counters
		static const int sort_env = getenv("FPOHM_CP_SORT") ? atoi(getenv("FPOHM_CP_SORT")) : -1;   // 0 never, 1 always
		static const int64_t sort_min = getenv("FPOHM_CP_SORT_MIN") ? atoll(getenv("FPOHM_CP_SORT_MIN")) : (1 << 16);
		static const int k2_search = getenv("FPOHM_K2_SEARCH") ? atoi(getenv("FPOHM_K2_SEARCH")) : K2_SEARCH_BUDGET;
		static const int k2_walk = getenv("FPOHM_K2_WALK") ? atoi(getenv("FPOHM_K2_WALK")) : K2_WALK_BUDGET;
		static const int minb = getenv("FPOHM_CP_MINB") ? atoi(getenv("FPOHM_CP_MINB")) : 6;      // debug A/B: CTAs per SM the packet kernel is compiled for
		static const int b3_budget = getenv("FPOHM_CP_B3") ? atoi(getenv("FPOHM_CP_B3")) : 1024;   // swept 256 .. 2048 on the nine A/B sets: 1024 takes 1 ms off the C4 lattice set, nothing moves elsewhere
		static const int c_budget = getenv("FPOHM_CP_CB") ? atoi(getenv("FPOHM_CP_CB")) : 32;
		FPOHM_REQUIRE(np < (1ll << 30), FPOHM_ERANGE, "closest point: %lld queries in one launch", (long long)np);
		const int blk = 128;
		// the kernels hand unfinished queries on through S/I/C, so all three must exist
		if (!I) { q.tmpI.alloc(np, s); I = q.tmpI.p; }
		if (!C) { q.tmpC.alloc(3 * np, s); C = q.tmpC.p; }
		if (!S) { q.tmpS.alloc(np, s); S = q.tmpS.p; }
		// packets want 32 neighbours: walk the batch in Morton order of a 1024^3 grid over its own bounding box.  Only keys and
		// indices are sorted (8 B/query); the kernels read and write the caller's arrays through the permutation.
		const uint32_t *perm = nullptr;
		const bool do_sort = sort_env == 0 ? false : (sort_env == 1 || sort_policy == CP_SORT_ALWAYS || (sort_policy == CP_SORT_AUTO && np >= sort_min));
		if (do_sort) {
			q.qbox.alloc(8, s); q.key.alloc(np, s); q.key2.alloc(np, s); q.idx.alloc(np, s); q.perm.alloc(np, s);
			static const int h_init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
			FPOHM_CUDA(cudaMemcpyAsync(q.qbox.p, h_init, sizeof(h_init), cudaMemcpyHostToDevice, s));
			query_bbox_kernel<<<grid_for(ctx, np, 256, 4), 256, 0, s>>>(P_dev, np, q.qbox.p);
			FPOHM_LAUNCH_CHECK(ctx);
			query_probe_kernel<<<1, 1024, 0, s>>>(P_dev, np, q.qbox.p, (sort_env == 1 || sort_policy == CP_SORT_ALWAYS) ? 1 : 0);
			FPOHM_LAUNCH_CHECK(ctx);
			query_keys_dev_kernel<<<grid_for(ctx, np, 256), 256, 0, s>>>(P_dev, np, q.qbox.p, q.key.p, q.idx.p);
			FPOHM_LAUNCH_CHECK(ctx);
			size_t tb = 0;
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, q.key.p, q.key2.p, q.idx.p, q.perm.p, (int)np, 0, 30, s));
			q.sort_tmp.alloc((int64_t)tb, s);
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(q.sort_tmp.p, tb, q.key.p, q.key2.p, q.idx.p, q.perm.p, (int)np, 0, 30, s));
			ctx->launches += 1;
			perm = q.perm.p;
		}
		const int pgrid = (int)((np + 127) / 128);        // one CTA per 128 queries: the block scheduler balances uneven packets
		q.todo.alloc(2 * np, s); q.todo_ties.alloc(TT_STRIDE * np, s); q.heavy.alloc(np, s);
		q.cnt.alloc(8, s);                                 // walk entries, heavy entries, search entries, next search entry, next packet
		FPOHM_CUDA(cudaMemsetAsync(q.cnt.p, 0, 8 * sizeof(int32_t), s));
		SignArgs sa = {m->V.p, m->F.p, m->nF, m->FN.p, m->VN.p, m->EN.p, m->EMAP.p, stats ? nullptr : N, (with_sign && (S || N)) ? 1 : 0};
		const int qslot = (int)(ctx->q_launches % fpohm_ctx::QRING);
		FPOHM_CUDA(cudaEventRecord(ctx->q_ev0[qslot], s));
		if (mode == 3 && m->nF < (1ll << 27)) {
			const int pers = (int)std::min<int64_t>(pgrid, (int64_t)ctx->sm_count * 6);
			if (stats && N) cp_pair_kernel<true, 3><<<pers, blk, 0, s>>>(m->wnodes.p, m->trif.p, m->tri.p, P_dev, np, perm, S, I, C, sa, m->eps_v, m->slack_q, q.todo.p, q.todo_ties.p, q.cnt.p, q.heavy.p, b3_budget, c_budget, N);
			else if (minb == 5) cp_pair_kernel<false, 5><<<(int)std::min<int64_t>(pgrid, (int64_t)ctx->sm_count * 5), blk, 0, s>>>(m->wnodes.p, m->trif.p, m->tri.p, P_dev, np, perm, S, I, C, sa, m->eps_v, m->slack_q, q.todo.p, q.todo_ties.p, q.cnt.p, q.heavy.p, b3_budget, c_budget, nullptr);
			else cp_pair_kernel<false, 6><<<pers, blk, 0, s>>>(m->wnodes.p, m->trif.p, m->tri.p, P_dev, np, perm, S, I, C, sa, m->eps_v, m->slack_q, q.todo.p, q.todo_ties.p, q.cnt.p, q.heavy.p, b3_budget, c_budget, nullptr);
		} else if (stats && N) cp_wide_kernel<true, 4><<<pgrid, blk, 0, s>>>(m->wnodes.p, m->tri.p, P_dev, np, perm, S, I, C, sa, q.todo.p, q.todo_ties.p, q.cnt.p, N);
		else cp_wide_kernel<false, 5><<<pgrid, blk, 0, s>>>(m->wnodes.p, m->tri.p, P_dev, np, perm, S, I, C, sa, q.todo.p, q.todo_ties.p, q.cnt.p, nullptr);
		FPOHM_CUDA(cudaEventRecord(ctx->q_ev1[qslot], s));
		ctx->q_launches++;
		FPOHM_LAUNCH_CHECK(ctx);
		cp_search_kernel<<<ctx->sm_count * 5, blk, 0, s>>>(m->qfnodes.p, m->qroot, m->tri.p, P_dev, np, q.todo.p, q.todo_ties.p, q.cnt.p, S, I, C, q.heavy.p, k2_search, sa);
		FPOHM_LAUNCH_CHECK(ctx);
		cp_tie_kernel<<<pgrid, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, m->prim_parent.p, P_dev, q.todo.p, q.todo_ties.p, q.cnt.p, S, I, C, q.heavy.p, k2_walk, sa, m->node_pd.p, getenv("FPOHM_CP_COUNTERS") != nullptr);
		FPOHM_LAUNCH_CHECK(ctx);
		// the list length is only known on the device: both forms are launched, one of them returns at once
		const int heavy_switch = ctx->sm_count * 6 * 4;    // = warps of the grid
		cp_heavy_warp_kernel<<<ctx->sm_count * 6, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, m->prim_parent.p, P_dev, q.heavy.p, q.cnt.p + 1, S, I, C, sa, heavy_switch);
		FPOHM_LAUNCH_CHECK(ctx);
		cp_heavy_kernel<<<ctx->sm_count * 6, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, m->prim_parent.p, P_dev, q.heavy.p, q.cnt.p + 1, S, I, C, sa, heavy_switch);
		FPOHM_LAUNCH_CHECK(ctx);
		static const bool counters = getenv("FPOHM_CP_COUNTERS") != nullptr;      // debug: work-list sizes on stderr (synchronises)
		if (counters) {
			int32_t hc[8] = {};
			FPOHM_CUDA(cudaMemcpyAsync(hc, q.cnt.p, sizeof(hc), cudaMemcpyDeviceToHost, s));
			FPOHM_CUDA(cudaStreamSynchronize(s));
			fprintf(stderr, "[fpohm cp] %lld queries: tie-break list %d, heavy list %d, search list %d, tie walks %d, walk give-ups %d\n", (long long)np, hc[0], hc[1], hc[2], hc[5], hc[6]);
		}
		return;
	}
	const int blk = 128;
	const int grid = grid_for(ctx, np, blk, 16);
	static const bool stats = getenv("FPOHM_CP_STATS") != nullptr;   // debug only: N := traversal counters
	// the sign pass needs facet + closest point even when the caller did not ask for them
	if (with_sign) {
		if (!I) { q.tmpI.alloc(np, s); I = q.tmpI.p; }
		if (!C) { q.tmpC.alloc(3 * np, s); C = q.tmpC.p; }
		if (!S && N) { q.tmpS.alloc(np, s); S = q.tmpS.p; }
	}
	cudaStream_t sc = s;   // (completion kernels run on the same stream)
	const SignArgs no_sign = {nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
	if (mode >= 1 && m->qroot >= 0) {
		FPOHM_REQUIRE(np < (1ll << 30), FPOHM_ERANGE, "closest point: %lld queries in one launch", (long long)np);
		// the kernels hand unfinished queries on through S/I/C, so all three must exist
		if (!I) { q.tmpI.alloc(np, s); I = q.tmpI.p; }
		if (!C) { q.tmpC.alloc(3 * np, s); C = q.tmpC.p; }
		if (!S) { q.tmpS.alloc(np, s); S = q.tmpS.p; }
		const int pgrid = (int)((np + 127) / 128);        // one CTA per 128 queries: the block scheduler balances uneven packets
		q.todo.alloc(2 * np, s); q.todo_ties.alloc(TT_STRIDE * np, s); q.heavy.alloc(np, s);
		q.cnt.alloc(4, s);                                 // walk entries, heavy entries, search entries, next search entry
		FPOHM_CUDA(cudaMemsetAsync(q.cnt.p, 0, 4 * sizeof(int32_t), s));
		static const int k2_search = getenv("FPOHM_K2_SEARCH") ? atoi(getenv("FPOHM_K2_SEARCH")) : K2_SEARCH_BUDGET;
		static const int k2_walk = getenv("FPOHM_K2_WALK") ? atoi(getenv("FPOHM_K2_WALK")) : K2_WALK_BUDGET;
		const int min_want = PK_WINDOW * PK_MIN_WANT, a_budget = PK_A_BUDGET;
		const int qslot = (int)(ctx->q_launches % fpohm_ctx::QRING);
		FPOHM_CUDA(cudaEventRecord(ctx->q_ev0[qslot], s));
		if (stats) cp_packet_kernel<true, 8><<<pgrid, blk, 0, s>>>(m->qfnodes.p, m->qroot, m->tri.p, P_dev, np, S, I, C, N, q.todo.p, q.todo_ties.p, q.cnt.p, min_want, a_budget);
		else cp_packet_kernel<false, 7><<<pgrid, blk, 0, s>>>(m->qfnodes.p, m->qroot, m->tri.p, P_dev, np, S, I, C, N, q.todo.p, q.todo_ties.p, q.cnt.p, min_want, a_budget);
		FPOHM_CUDA(cudaEventRecord(ctx->q_ev1[qslot], s));
		ctx->q_launches++;
		FPOHM_LAUNCH_CHECK(ctx);
		cp_search_kernel<<<ctx->sm_count * 5, blk, 0, sc>>>(m->qfnodes.p, m->qroot, m->tri.p, P_dev, np, q.todo.p, q.todo_ties.p, q.cnt.p, S, I, C, q.heavy.p, k2_search, no_sign);
		FPOHM_LAUNCH_CHECK(ctx);
		cp_tie_kernel<<<pgrid, blk, 0, sc>>>(m->qnodes.p, m->qroot, m->tri.p, m->prim_parent.p, P_dev, q.todo.p, q.todo_ties.p, q.cnt.p, S, I, C, q.heavy.p, k2_walk, no_sign, nullptr);
		FPOHM_LAUNCH_CHECK(ctx);
		cp_heavy_warp_kernel<<<ctx->sm_count * 6, blk, 0, sc>>>(m->qnodes.p, m->qroot, m->tri.p, m->prim_parent.p, P_dev, q.heavy.p, q.cnt.p + 1, S, I, C, no_sign, 0);
	} else {
		if (stats) closest_point_kernel<true><<<grid, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, P_dev, np, S, I, C, N);
		else closest_point_kernel<false><<<grid, blk, 0, s>>>(m->qnodes.p, m->qroot, m->tri.p, P_dev, np, S, I, C, N);
	}
	FPOHM_LAUNCH_CHECK(ctx);
	if (with_sign && (S || N)) {
		pseudonormal_kernel<<<grid, blk, 0, sc>>>(m->V.p, m->F.p, m->nF, m->FN.p, m->VN.p, m->EN.p, m->EMAP.p, P_dev, np, I, C, S, N, !stats);
		FPOHM_LAUNCH_CHECK(ctx);
	}
}

void launch_closest_point(fpohm_ctx *ctx, fpohm_mesh *m, bool with_sign, const double *P_dev, int64_t np,
                          double *S, int32_t *I, double *C, double *N, cudaStream_t s, int sort_policy)
{
	QueryScratch q;
	launch_closest_point_ex(ctx, m, with_sign, P_dev, np, S, I, C, N, s, q, sort_policy);
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

// Host-pointer path.  Queries are independent, so the batch is cut into chunks: an upload stream runs ahead with every
// H2D copy, three compute lanes rotate over the chunks (the latency-bound tails of one chunk's completion kernels run under
// the next chunks' packet walks), and each chunk's results go down on its own lane as soon as they exist — in completion
// order: on the bench workload the first two chunks hold the heavy queries and finish after the four that follow them
// (FPOHM_CP_TIMELINE=1 prints the per-chunk event times).  With pinned caller buffers the
// PCIe copies (84 B/query, 55 GB/s each way) hide behind the search; with pageable memory CUDA stages the copies and
// the pipeline degrades gracefully to the serial order.
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
	// The packet search lives on consecutive queries being close to each other.  Probe that on the host (4 096 sampled
	// neighbours against the spacing np points would have if spread evenly over their box); a batch that fails is walked
	// in Morton order and its results are scattered back to the caller's order (a shuffled 4.6 M batch: 3.6 x faster).
	bool incoherent = false;
	if (np >= (1 << 16)) {
		// same criterion as query_probe_kernel: sampled packets of 32 consecutive queries against an evenly spread clump of 32
		const int64_t m = 4096, npk = np / 32, stride = npk > m ? npk / m : 1;
		double mn[3] = {P[0], P[1], P[2]}, mx[3] = {P[0], P[1], P[2]}, adj = 0;
		int64_t cnt = 0;
		for (int64_t k = 0; k < m && k * stride < npk; ++k) {
			const int64_t base = k * stride * 32;
			double lo[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL}, hi[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
			for (int j : {0, 10, 21, 31}) {
				const double *a = P + 3 * (base + j);
				for (int c = 0; c < 3; ++c) if (std::isfinite(a[c])) { lo[c] = std::min(lo[c], a[c]); hi[c] = std::max(hi[c], a[c]); mn[c] = std::min(mn[c], a[c]); mx[c] = std::max(mx[c], a[c]); }
			}
			for (int c = 0; c < 3; ++c) if (hi[c] >= lo[c]) adj += (hi[c] - lo[c]) * (hi[c] - lo[c]);
			++cnt;
		}
		adj /= (double)std::max<int64_t>(cnt, 1);
		const double em = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
		const double vol = std::max(mx[0] - mn[0], 0.01 * em) * std::max(mx[1] - mn[1], 0.01 * em) * std::max(mx[2] - mn[2], 0.01 * em);
		const double D2 = em > 0 ? 16.0 * 3.0 * std::pow(32.0 * vol / (double)np, 2.0 / 3.0) : 0.0;
		static const bool never = getenv("FPOHM_CP_NOSORT") != nullptr;
		if (!never && D2 > 0 && std::isfinite(adj) && adj > D2) {
			FPOHM_REQUIRE(np < (1ll << 30), FPOHM_ERANGE, "%s: %lld queries in one call", who, (long long)np);
			const double ext = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
			if (cp_mode() >= 2 && mesh->n_wnodes > 0) {
				// the wide-packet kernels walk a permutation themselves (no gather / scatter passes): the batch goes through the same
				// upload / compute / download pipeline as a coherent one, in a few large chunks that are each walked in Morton order
				incoherent = true;
				goto pipeline;
			}
			dP.upload(P, 3 * np);
			DevBuf<uint32_t> key(np, s), key2(np, s), idx(np, s), perm(np, s);
			query_keys_kernel<<<grid_for(ctx, np, 256), 256, 0, s>>>(dP.p, np, mn[0], mn[1], mn[2], ext > 0 ? 1024.0 / ext : 0.0, key.p, idx.p);
			FPOHM_LAUNCH_CHECK(ctx);
			size_t tb = 0;
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, perm.p, (int)np, 0, 30, s));
			DevBuf<uint8_t> tmp((int64_t)tb, s);
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, perm.p, (int)np, 0, 30, s));
			ctx->launches += 1;
			DevBuf<double> sP(3 * np, s), sS(np, s), sC(3 * np, s), sN(N ? 3 * np : 0, s);
			DevBuf<int32_t> sI(np, s);
			gather_points_kernel<<<grid_for(ctx, 3 * np, 256), 256, 0, s>>>(dP.p, perm.p, np, sP.p);
			FPOHM_LAUNCH_CHECK(ctx);
			launch_closest_point(ctx, mesh, with_sign, sP.p, np, sS.p, sI.p, sC.p, N ? sN.p : nullptr, s);
			scatter_results_kernel<<<grid_for(ctx, np, 256), 256, 0, s>>>(perm.p, np, sS.p, sI.p, sC.p, N ? sN.p : nullptr, S ? dS.p : nullptr,
				I ? dI.p : nullptr, C ? dC.p : nullptr, N ? dN.p : nullptr);
			FPOHM_LAUNCH_CHECK(ctx);
			if (S) dS.download(S, np);
			if (I) dI.download(I, np);
			if (C) dC.download(C, 3 * np);
			if (N) dN.download(N, 3 * np);
			t.stop();
			FPOHM_CUDA(cudaStreamSynchronize(s));
			return FPOHM_OK;
		}
	}
pipeline:
	static const int64_t chunk_coherent = getenv("FPOHM_CP_CHUNK") ? atoll(getenv("FPOHM_CP_CHUNK")) : (1 << 19);
	static const int64_t chunk_sorted_env = getenv("FPOHM_CP_CHUNK_SORT") ? atoll(getenv("FPOHM_CP_CHUNK_SORT")) : 0;
	static const int n_comp_env = getenv("FPOHM_CP_LANES") ? atoi(getenv("FPOHM_CP_LANES")) : 0;
	// A chunk of an incoherent batch is ordered on its own, so the fewer points it has the wider its packets are: chunks are large
	// and of equal size.  If a stretch of 2^20 consecutive queries is a compact part of the batch (lattice rows: 35.4 ms for the
	// 10 M classification set against 45.2 ms in one piece) chunks of 2^20 overlap best; a batch in random order is best cut at
	// 2^21 (43.6 against 45.6 ms; 2^20: 46.9, 2^19: 55 ms).  scripts/e2e_c4.py.
	int64_t chunk = chunk_coherent;
	if (incoherent) {
		int64_t target = chunk_sorted_env;
		if (target <= 0) {
			target = 1 << 21;
			if (np > (1 << 20)) {
				double lo[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL}, hi[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL}, all_lo[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL}, all_hi[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
				for (int64_t k = 0; k < 2048; ++k) {
					const double *a = P + 3 * (k * ((1 << 20) / 2048)), *b = P + 3 * (k * (np / 2048));
					for (int c = 0; c < 3; ++c) {
						if (std::isfinite(a[c])) { lo[c] = std::min(lo[c], a[c]); hi[c] = std::max(hi[c], a[c]); }
						if (std::isfinite(b[c])) { all_lo[c] = std::min(all_lo[c], b[c]); all_hi[c] = std::max(all_hi[c], b[c]); }
					}
				}
				double v = 1, va = 1;
				for (int c = 0; c < 3; ++c) { v *= std::max(hi[c] - lo[c], 0.0); va *= std::max(all_hi[c] - all_lo[c], 0.0); }
				if (va > 0 && v <= 0.3 * va) target = 1 << 20;
			}
			// a batch of only a few such chunks would barely overlap its transfers with its compute: at least four chunks, none under 2^19
			// (the 1.78 M projection set: 6.98 ms in one piece, 5.5 ms in four)
			if (np < 4 * target) target = std::max<int64_t>(1 << 19, (np + 3) / 4);
		}
		const int64_t k = (np + target - 1) / target;
		chunk = (((np + k - 1) / k) + 31) & ~(int64_t)31;
	}
	const int n_comp = n_comp_env > 0 ? n_comp_env : (incoherent ? 4 : 3);
	const int sort_policy = incoherent ? CP_SORT_ALWAYS : CP_SORT_NEVER;
	const int64_t n_chunks = (np + chunk - 1) / chunk;
	while ((int64_t)ctx->ev_pool.size() < 2 * n_chunks) {
		cudaEvent_t e;
		FPOHM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		ctx->ev_pool.push_back(e);
	}
	cudaStream_t up = ctx->aux[0], down0 = ctx->aux[1];
	cudaStream_t comp[4] = {s, ctx->aux[2], ctx->aux[3], ctx->aux[4]};
	const int lanes_n = std::max(1, std::min(n_comp, 4));
	FPOHM_CUDA(cudaEventRecord(ctx->ev_sync, s));              // allocations are ordered on s
	FPOHM_CUDA(cudaStreamWaitEvent(up, ctx->ev_sync, 0));
	FPOHM_CUDA(cudaStreamWaitEvent(down0, ctx->ev_sync, 0));
	for (int k = 1; k < lanes_n; ++k) FPOHM_CUDA(cudaStreamWaitEvent(comp[k], ctx->ev_sync, 0));
	// every upload is queued at once on its own stream: the copy engine runs ahead of the kernels
	for (int64_t k = 0; k < n_chunks; ++k) {
		const int64_t o = k * chunk, n = std::min(chunk, np - o);
		FPOHM_CUDA(cudaMemcpyAsync(dP.p + 3 * o, P + 3 * o, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, up));
		FPOHM_CUDA(cudaEventRecord(ctx->ev_pool[(size_t)(2 * k)], up));
	}
	static const bool timeline = getenv("FPOHM_CP_TIMELINE") != nullptr;      // debug: per-chunk event timestamps on stderr
	std::vector<cudaEvent_t> tl;
	if (timeline) { tl.resize((size_t)(4 * n_chunks + 1)); for (auto &e : tl) cudaEventCreate(&e); cudaEventRecord(tl[(size_t)(4 * n_chunks)], s); }
	for (int64_t k = 0; k < n_chunks; ++k) {
		const int64_t o = k * chunk, n = std::min(chunk, np - o);
		cudaStream_t cs = comp[k % lanes_n];
		FPOHM_CUDA(cudaStreamWaitEvent(cs, ctx->ev_pool[(size_t)(2 * k)], 0));
		if (timeline) cudaEventRecord(tl[(size_t)(4 * k)], cs);
		launch_closest_point(ctx, mesh, with_sign, dP.p + 3 * o, n, S ? dS.p + o : nullptr, I ? dI.p + o : nullptr,
		                     C ? dC.p + 3 * o : nullptr, N ? dN.p + 3 * o : nullptr, cs, sort_policy);   // decided by the host probe above
		FPOHM_CUDA(cudaEventRecord(ctx->ev_pool[(size_t)(2 * k + 1)], cs));
		if (timeline) cudaEventRecord(tl[(size_t)(4 * k + 1)], cs);
		static const bool own_down = getenv("FPOHM_CP_DOWN") ? atoi(getenv("FPOHM_CP_DOWN")) != 0 : true;
		cudaStream_t down = own_down ? cs : down0;      // results go down on the chunk's own lane: in COMPLETION order, not chunk order
		if (!own_down) FPOHM_CUDA(cudaStreamWaitEvent(down, ctx->ev_pool[(size_t)(2 * k + 1)], 0));
		if (timeline) cudaEventRecord(tl[(size_t)(4 * k + 2)], down);
		if (S) FPOHM_CUDA(cudaMemcpyAsync(S + o, dS.p + o, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, down));
		if (I) FPOHM_CUDA(cudaMemcpyAsync(I + o, dI.p + o, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, down));
		if (C) FPOHM_CUDA(cudaMemcpyAsync(C + 3 * o, dC.p + 3 * o, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, down));
		if (N) FPOHM_CUDA(cudaMemcpyAsync(N + 3 * o, dN.p + 3 * o, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, down));
		if (timeline) cudaEventRecord(tl[(size_t)(4 * k + 3)], down);
	}
	if (timeline) {
		cudaDeviceSynchronize();
		for (int64_t k = 0; k < n_chunks; ++k) {
			float a, b, c, d;
			cudaEventElapsedTime(&a, tl[(size_t)(4 * n_chunks)], tl[(size_t)(4 * k)]); cudaEventElapsedTime(&b, tl[(size_t)(4 * n_chunks)], tl[(size_t)(4 * k + 1)]);
			cudaEventElapsedTime(&c, tl[(size_t)(4 * n_chunks)], tl[(size_t)(4 * k + 2)]); cudaEventElapsedTime(&d, tl[(size_t)(4 * n_chunks)], tl[(size_t)(4 * k + 3)]);
			fprintf(stderr, "[fpohm timeline] chunk %2lld compute %.3f .. %.3f ms   download %.3f .. %.3f ms\n", (long long)k, a, b, c, d);
		}
		for (auto &e : tl) cudaEventDestroy(e);
	}
	// join everything back into s before the buffers die
	FPOHM_CUDA(cudaEventRecord(ctx->ev_sync, down0));
	FPOHM_CUDA(cudaStreamWaitEvent(s, ctx->ev_sync, 0));
	for (int k = 1; k < lanes_n; ++k) {
		FPOHM_CUDA(cudaEventRecord(ctx->ev_sync, comp[k]));
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

} // extern "C"

// device form of fpohm_classify_hexes, shared with cleaning.cu
void fpohm::classify_hexes_dev(fpohm_ctx *ctx, fpohm_mesh *surface, const double *V_dev, const uint32_t *hex_dev, int64_t H,
                               double *S_dev, uint8_t *flag_dev, cudaStream_t s)
{
	mesh_ensure_tree(ctx, surface, s);
	DevBuf<double> dP(3 * H, s);
	hex_box_centres_kernel<<<grid_for(ctx, 3 * H, 256), 256, 0, s>>>(V_dev, hex_dev, H, dP.p);
	FPOHM_LAUNCH_CHECK(ctx);
	launch_closest_point(ctx, surface, true, dP.p, H, S_dev, nullptr, nullptr, nullptr, s);
	inside_flags_kernel<<<grid_for(ctx, H, 256), 256, 0, s>>>(S_dev, H, flag_dev);
	FPOHM_LAUNCH_CHECK(ctx);
}

extern "C" {

int fpohm_classify_hexes(fpohm_ctx *ctx, fpohm_mesh *surface, const double *V, int64_t nV, const uint32_t *hex, int64_t H,
                         double *signed_dis, uint8_t *H_flag)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && surface && V && hex && nV > 0 && H >= 0, FPOHM_EINVAL, "fpohm_classify_hexes: bad argument");
	if (H == 0) return FPOHM_OK;
	for (int64_t i = 0; i < 8 * H; ++i) FPOHM_REQUIRE((int64_t)hex[i] < nV, FPOHM_EINVAL, "fpohm_classify_hexes: corner id %u out of range at %lld", hex[i], (long long)i);
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dV(3 * nV, s), dS(H, s);
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<uint8_t> dF(H, s);
	dV.upload(V, 3 * nV); dhex.upload(hex, 8 * H);
	KernelTimer t(ctx, s);
	classify_hexes_dev(ctx, surface, dV.p, dhex.p, H, dS.p, dF.p, s);
	t.stop();
	if (signed_dis) dS.download(signed_dis, H);
	if (H_flag) dF.download(H_flag, H);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_point_mesh_sqdist(fpohm_ctx *ctx, fpohm_mesh *mesh, const double *P, int64_t np,
                            double *sqrD, int32_t *I, double *C)
{
	return host_query(ctx, mesh, false, P, np, sqrD, I, C, nullptr, "fpohm_point_mesh_sqdist");
}

} // extern "C"

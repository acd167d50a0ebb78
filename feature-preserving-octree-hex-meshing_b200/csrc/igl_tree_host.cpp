// Host-side construction of the closest-point tree and the pseudonormal tables.
//
// The tree must be IDENTICAL to igl::AABB<MatrixXd,3>::init (igl/AABB.cpp:30-200): the winner among
// equidistant facets is the first one visited (strict '<' in set_min, igl/AABB.cpp:773), and the visiting
// order is a function of the tree shape.  The shape depends on (a) the tie order std::sort gives equal
// barycentre coordinates (igl/sort.cpp:281-300, IndexLessThan) and (b) what std::nth_element leaves at
// position n after the second call of igl's median lambda (igl/AABB.cpp:154-167) — both are properties of
// the C++ standard library, so this file calls the same std:: algorithms on the same data in the same
// order rather than "an equivalent" algorithm.  It is compiled by the host compiler (no CUDA in here).
#include "mesh.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace fpohm {

namespace {

struct IndexLess {
	const std::vector<double> &arr;
	bool operator()(const size_t a, const size_t b) const { return arr[a] < arr[b]; }
};

// igl::sort(BC,1,true,_,IS) for one column; result: order[i] = index of the i-th smallest
void igl_sort_column(const std::vector<double> &data, std::vector<size_t> &order) {
	const size_t n = data.size();
	order.resize(n);
	for (size_t i = 0; i < n; ++i) order[i] = i;
	if (n == 2) { // igl::sort2
		if (data[0] > data[1]) std::swap(order[0], order[1]);
		return;
	}
	if (n == 3) { // igl::sort3, igl/sort.cpp:226-248
		double a = data[0], b = data[1], c = data[2];
		size_t ai = 0, bi = 1, ci = 2;
		if (a > b) { std::swap(a, b); std::swap(ai, bi); }
		if (b > c) {
			std::swap(b, c); std::swap(bi, ci);
			if (a > b) { std::swap(a, b); std::swap(ai, bi); }
		}
		order[0] = ai; order[1] = bi; order[2] = ci;
		return;
	}
	std::sort(order.begin(), order.end(), IndexLess{data});
}

struct Builder {
	const int32_t *F;
	const std::vector<double> &tbox; // 6 per facet
	const std::vector<int32_t> &SI;  // 3 per facet: rank on each axis
	HostTree &out;

	int32_t build(std::vector<int32_t> &I) {
		const int32_t me = (int32_t)out.prim.size();
		out.prim.push_back(-1);
		out.lr.push_back(-1); out.lr.push_back(-1);
		out.box.resize(out.box.size() + 6);
		double mn[3] = {std::numeric_limits<double>::max(), std::numeric_limits<double>::max(), std::numeric_limits<double>::max()};
		double mx[3] = {-mn[0], -mn[1], -mn[2]};
		for (int32_t f : I) {
			const double *b = &tbox[6 * (size_t)f];
			for (int c = 0; c < 3; ++c) {
				if (b[c] < mn[c]) mn[c] = b[c];
				if (b[3 + c] > mx[c]) mx[c] = b[3 + c];
			}
		}
		for (int c = 0; c < 3; ++c) { out.box[6 * (size_t)me + c] = mn[c]; out.box[6 * (size_t)me + 3 + c] = mx[c]; }
		const size_t n = I.size();
		if (n == 1) {
			out.prim[me] = I[0];
			return me;
		}
		// longest direction: Eigen maxCoeff keeps the first strict maximum
		int max_d = 0;
		double best = mx[0] - mn[0];
		for (int c = 1; c < 3; ++c) {
			const double d = mx[c] - mn[c];
			if (d > best) { best = d; max_d = c; }
		}
		std::vector<int> SIdI(n);
		for (size_t i = 0; i < n; ++i) SIdI[i] = SI[3 * (size_t)I[i] + max_d];
		double med;
		{
			std::vector<int> A(SIdI);
			const size_t h = A.size() / 2;
			std::nth_element(A.data(), A.data() + h, A.data() + A.size());
			if (A.size() % 2 == 1) {
				med = A[h];
			} else {
				std::nth_element(A.data(), A.data() + h - 1, A.data() + A.size());
				med = 0.5 * (A[h] + A[h - 1]);
			}
		}
		std::vector<int32_t> LI, RI;
		LI.reserve((n + 1) / 2); RI.reserve(n / 2 + 1);
		for (size_t i = 0; i < n; ++i) {
			if (SIdI[i] <= med) LI.push_back(I[i]); else RI.push_back(I[i]);
		}
		std::vector<int32_t>().swap(I);
		std::vector<int>().swap(SIdI);
		int32_t l = -1, r = -1;
		if (!LI.empty()) l = build(LI);
		if (!RI.empty()) r = build(RI);
		out.lr[2 * (size_t)me] = l;
		out.lr[2 * (size_t)me + 1] = r;
		return me;
	}
};

} // namespace

void build_igl_tree(const double *V, int64_t nV, const int32_t *F, int64_t nF, HostTree &out) {
	(void)nV;
	out.box.clear(); out.prim.clear(); out.lr.clear();
	if (nF <= 0) return;
	out.prim.reserve(2 * (size_t)nF); out.lr.reserve(4 * (size_t)nF); out.box.reserve(12 * (size_t)nF);
	// barycentres, igl/barycenter.cpp: ((0 + v0) + v1) + v2, then `/= 3.0`.  NB Eigen 3.2's operator/=(scalar)
	// multiplies by Scalar(1)/other (Eigen/src/Core/SelfCwiseBinaryOp.h:181-193) — also for per_face_normals' `N.row(i) /= r` below.
	std::vector<double> col[3];
	for (int d = 0; d < 3; ++d) col[d].resize((size_t)nF);
	std::vector<double> tbox(6 * (size_t)nF);
	const double third = 1.0 / 3.0;
	for (int64_t f = 0; f < nF; ++f) {
		const double *a = V + 3 * (int64_t)F[3 * f], *b = V + 3 * (int64_t)F[3 * f + 1], *c = V + 3 * (int64_t)F[3 * f + 2];
		for (int d = 0; d < 3; ++d) {
			double s = 0.0;
			s += a[d]; s += b[d]; s += c[d];
			col[d][(size_t)f] = s * third;
			tbox[6 * (size_t)f + d] = std::min(a[d], std::min(b[d], c[d]));
			tbox[6 * (size_t)f + 3 + d] = std::max(a[d], std::max(b[d], c[d]));
		}
	}
	std::vector<int32_t> SI(3 * (size_t)nF);
	{
		std::vector<size_t> order;
		for (int d = 0; d < 3; ++d) {
			igl_sort_column(col[d], order);
			for (size_t i = 0; i < (size_t)nF; ++i) SI[3 * order[i] + d] = (int32_t)i;
			std::vector<double>().swap(col[d]);
		}
	}
	std::vector<int32_t> I((size_t)nF);
	for (int64_t f = 0; f < nF; ++f) I[(size_t)f] = (int32_t)f;
	Builder b{F, tbox, SI, out};
	b.build(I);
}

// per_face_normals (igl/per_face_normals.cpp:13-36), per_vertex_normals ANGLE (igl/per_vertex_normals.cpp:38-108,
// igl/internal_angles.cpp:64-87, igl/squared_edge_lengths.cpp:30-44), per_edge_normals UNIFORM
// (igl/per_edge_normals.cpp:20-77, igl/all_edges.cpp:35-42, igl/unique_simplices.cpp:16-33).
void build_igl_normals(const double *V, int64_t nV, const int32_t *F, int64_t nF,
                       std::vector<double> &FN, std::vector<double> &VN, std::vector<double> &EN,
                       std::vector<int32_t> &E, std::vector<int32_t> &EMAP)
{
	FN.assign(3 * (size_t)nF, 0.0);
	VN.assign(3 * (size_t)nV, 0.0);
	// rows of dynamic-size Eigen matrices reduce sequentially: (x^2 + y^2) + z^2 (Eigen 3.2 DefaultTraversal/NoUnrolling)
	auto sqn = [](double x, double y, double z) { return (x * x + y * y) + z * z; };
	for (int64_t f = 0; f < nF; ++f) {
		const double *p0 = V + 3 * (int64_t)F[3 * f], *p1 = V + 3 * (int64_t)F[3 * f + 1], *p2 = V + 3 * (int64_t)F[3 * f + 2];
		const double a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
		const double b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
		double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
		const double r = std::sqrt(sqn(n[0], n[1], n[2]));
		if (r == 0) { n[0] = n[1] = n[2] = 0; } else { const double ir = 1.0 / r; n[0] *= ir; n[1] *= ir; n[2] *= ir; }
		for (int c = 0; c < 3; ++c) FN[3 * (size_t)f + c] = n[c];
	}
	for (int64_t f = 0; f < nF; ++f) {
		const double *p0 = V + 3 * (int64_t)F[3 * f], *p1 = V + 3 * (int64_t)F[3 * f + 1], *p2 = V + 3 * (int64_t)F[3 * f + 2];
		double L[3];
		L[0] = sqn(p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]);
		L[1] = sqn(p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]);
		L[2] = sqn(p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]);
		for (int d = 0; d < 3; ++d) {
			const double s1 = L[d], s2 = L[(d + 1) % 3], s3 = L[(d + 2) % 3];
			const double w = std::acos((s3 + s2 - s1) / (2. * std::sqrt(s3 * s2)));
			double *vn = &VN[3 * (size_t)F[3 * f + d]];
			for (int c = 0; c < 3; ++c) vn[c] += w * FN[3 * (size_t)f + c];
		}
	}
	for (int64_t v = 0; v < nV; ++v) {
		double *vn = &VN[3 * (size_t)v];
		// N.rowwise().normalize() is a cwiseQuotient by the row norm (Eigen/src/Core/VectorwiseOp.h:539-550): true division
		const double r = std::sqrt(sqn(vn[0], vn[1], vn[2]));
		vn[0] /= r; vn[1] /= r; vn[2] /= r;
	}
	// undirected edges: directed edge (f,c) = (F[(c+1)%3], F[(c+2)%3]), stored at row f + c*nF
	const size_t m = (size_t)nF;
	std::vector<uint64_t> key(3 * m);
	for (size_t f = 0; f < m; ++f)
		for (int c = 0; c < 3; ++c) {
			uint32_t a = (uint32_t)F[3 * f + (c + 1) % 3], b = (uint32_t)F[3 * f + (c + 2) % 3];
			if (a > b) std::swap(a, b);
			key[f + (size_t)c * m] = ((uint64_t)a << 32) | b;
		}
	std::vector<uint64_t> uniq(key);
	std::sort(uniq.begin(), uniq.end());
	uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
	E.resize(2 * uniq.size());
	for (size_t e = 0; e < uniq.size(); ++e) { E[2 * e] = (int32_t)(uniq[e] >> 32); E[2 * e + 1] = (int32_t)(uniq[e] & 0xffffffffu); }
	EMAP.resize(3 * m);
	for (size_t i = 0; i < 3 * m; ++i) EMAP[i] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), key[i]) - uniq.begin());
	EN.assign(3 * uniq.size(), 0.0);
	for (size_t f = 0; f < m; ++f)
		for (int c = 0; c < 3; ++c) {
			double *en = &EN[3 * (size_t)EMAP[f + (size_t)c * m]];
			for (int k = 0; k < 3; ++k) en[k] += FN[3 * f + k];
		}
}

} // namespace fpohm

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
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <thread>

namespace fpohm {

namespace {

struct IndexLess {
	const std::vector<double> &arr;
	bool operator()(const size_t a, const size_t b) const { return arr[a] < arr[b]; }
};

// std::sort's RESULT, order of equal keys included, on several threads.  libstdc++'s std::sort is
//     __introsort_loop(first, last, 2 * lg(n));  __final_insertion_sort(first, last);
// where the loop partitions [first, last) around a median-of-three pivot, recurses into the right part and iterates on
// the left one, and the closing insertion pass is a STABLE sort of whatever the loop left behind.  What happens inside a
// sub-range depends only on that sub-range's content and on the depth budget it is entered with, and a cut separates
// "everything <= " from "everything >=".  So the right parts can run on other threads, and the closing pass can be run per
// range between cuts, with the same outcome element for element.  The pieces are the library's OWN internals
// (std::__unguarded_partition_pivot, std::__introsort_loop, std::__partial_sort, std::__insertion_sort): nothing of the
// algorithm is restated here except the ten-line loop that strings them together, so the tie order stays libstdc++'s by
// construction.  Other standard libraries (or FPOHM_SORT_SERIAL=1) take the plain std::sort call.
// The 2 M-facet C3 mesh spent 170 ms per axis in this sort (all three axes have tied barycentres); ~25 ms with this.
template <class T, class Less>
void sort_like_std(T *first, T *last, Less less) {
#if defined(__GLIBCXX__)
	static const bool serial = getenv("FPOHM_SORT_SERIAL") != nullptr;
	const std::ptrdiff_t n = last - first, min_par = 1 << 15;
	if (serial || n < 4 * min_par) { std::sort(first, last, less); return; }
	auto comp = __gnu_cxx::__ops::__iter_comp_iter(less);
	using Comp = decltype(comp);
	std::mutex mu;
	std::vector<T *> cuts;                       // starts of the ranges finished by one serial call
	std::atomic<int> budget((int)std::max(2u, std::thread::hardware_concurrency()));
	struct Run {
		Comp &comp; std::mutex &mu; std::vector<T *> &cuts; std::atomic<int> &budget; std::ptrdiff_t min_par;
		void operator()(T *lo, T *hi, long depth) {
			std::vector<std::thread> kids;
			while (hi - lo > min_par) {
				if (depth == 0) break;               // the serial call below does the heap-sort fallback of this range
				--depth;
				T *cut = std::__unguarded_partition_pivot(lo, hi, comp);
				if (hi - cut > min_par && budget.fetch_sub(1) > 0) {
					kids.emplace_back([this, cut, hi, depth]() { (*this)(cut, hi, depth); budget.fetch_add(1); });   // the slot is free as soon as the range is done
				} else {
					if (hi - cut > min_par) budget.fetch_add(1);
					(*this)(cut, hi, depth);
				}
				hi = cut;
			}
			std::__introsort_loop(lo, hi, depth, comp);
			{ std::lock_guard<std::mutex> g(mu); cuts.push_back(lo); }
			for (auto &t : kids) t.join();
		}
	} run{comp, mu, cuts, budget, min_par};
	run(first, last, (long)std::__lg(n) * 2);
	std::sort(cuts.begin(), cuts.end());
	cuts.push_back(last);
	// closing pass: a stable insertion sort per range (ranges are ordered among themselves)
	const int nt = (int)std::min<size_t>(std::max(2u, std::thread::hardware_concurrency()), cuts.size() - 1);
	std::atomic<size_t> next(0);
	std::vector<std::thread> th;
	for (int t = 0; t < nt; ++t)
		th.emplace_back([&]() { for (size_t i = next.fetch_add(1); i + 1 < cuts.size(); i = next.fetch_add(1)) std::__insertion_sort(cuts[i], cuts[i + 1], comp); });
	for (auto &t : th) t.join();
#else
	std::sort(first, last, less);
#endif
}

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
	// std::sort is comparison based: sorting (key, index) records by key performs exactly the swaps that sorting the indices
	// through igl's indirect IndexLessThan does, without the two cache misses per comparison
	struct Rec { double k; size_t i; };
	std::vector<Rec> rec(n);
	for (size_t i = 0; i < n; ++i) rec[i] = {data[i], i};
	sort_like_std(rec.data(), rec.data() + n, [](const Rec &a, const Rec &b) { return a.k < b.k; });
	for (size_t i = 0; i < n; ++i) order[i] = rec[i].i;
}

// A tree over n facets is a full binary tree with 2n-1 nodes, so DFS pre-order positions are known up front
// (left child = me + 1, right child = me + 2 * n_left): subtrees can be filled independently and the top levels are
// built by parallel tasks.  Every std:: call still sees exactly the data igl would pass it.
struct Builder {
	const std::vector<double> &tbox; // 6 per facet
	const std::vector<int32_t> &SI;  // 3 per facet: rank on each axis
	HostTree &out;

	void build(std::vector<int32_t> &I, int32_t me, int depth) {
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
			return;
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
		// ranks are distinct, so both sides are non-empty (igl sizes them (n+1)/2 and n/2, AABB.cpp:168)
		const int32_t l = me + 1, r = me + 2 * (int32_t)LI.size();
		out.lr[2 * (size_t)me] = l;
		out.lr[2 * (size_t)me + 1] = r;
		if (depth < 4 && n > 4096) {
			std::thread t([&]() { build(LI, l, depth + 1); });
			build(RI, r, depth + 1);
			t.join();
		} else {
			build(LI, l, depth + 1);
			build(RI, r, depth + 1);
		}
	}
};

} // namespace

// rank of every facet's barycentre on axis d in igl::sort's order (igl/AABB.cpp:63-88): the host half of the device tree build,
// needed only for axes on which two barycentre coordinates are EQUAL (their order is libstdc++'s introsort's, tree_device.cu)
void host_rank_axis(const double *V, const int32_t *F, int64_t nF, int d, int32_t *rank) {
	std::vector<double> col((size_t)nF);
	const double third = 1.0 / 3.0;
	for (int64_t f = 0; f < nF; ++f) {
		double s = 0.0;
		s += V[3 * (int64_t)F[3 * f] + d]; s += V[3 * (int64_t)F[3 * f + 1] + d]; s += V[3 * (int64_t)F[3 * f + 2] + d];
		col[(size_t)f] = s * third;
	}
	std::vector<size_t> order;
	igl_sort_column(col, order);
	for (size_t i = 0; i < (size_t)nF; ++i) rank[order[i]] = (int32_t)i;
}

void build_igl_tree(const double *V, int64_t nV, const int32_t *F, int64_t nF, HostTree &out) {
	(void)nV;
	out.box.clear(); out.prim.clear(); out.lr.clear();
	if (nF <= 0) return;
	const size_t nn = 2 * (size_t)nF - 1;
	out.prim.assign(nn, -1); out.lr.assign(2 * nn, -1); out.box.assign(6 * nn, 0.0);
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
		// the three per-axis sorts are independent (igl/sort.cpp:50-79 loops over columns): one thread each
		auto sort_axis = [&](int d) {
			std::vector<size_t> order;
			igl_sort_column(col[d], order);
			for (size_t i = 0; i < (size_t)nF; ++i) SI[3 * order[i] + d] = (int32_t)i;
			std::vector<double>().swap(col[d]);
		};
		if (nF > 4096) {
			std::thread t0(sort_axis, 0), t1(sort_axis, 1);
			sort_axis(2);
			t0.join(); t1.join();
		} else {
			for (int d = 0; d < 3; ++d) sort_axis(d);
		}
	}
	std::vector<int32_t> I((size_t)nF);
	for (int64_t f = 0; f < nF; ++f) I[(size_t)f] = (int32_t)f;
	Builder b{tbox, SI, out};
	b.build(I, 0, 0);
}

// per_face_normals (igl/per_face_normals.cpp:13-36), per_vertex_normals ANGLE (igl/per_vertex_normals.cpp:38-108,
// igl/internal_angles.cpp:64-87, igl/squared_edge_lengths.cpp:30-44), per_edge_normals UNIFORM
// (igl/per_edge_normals.cpp:20-77, igl/all_edges.cpp:35-42, igl/unique_simplices.cpp:16-33).
// static range split over host threads (results do not depend on the thread count: every output element is produced
// by exactly one thread with the reference's sequential order of operations)
template <class Fn>
static void parallel_ranges(int64_t n, const Fn &fn) {
	unsigned T = std::thread::hardware_concurrency();
	if (T == 0) T = 1;
	if (T > 16) T = 16;
	if (n < 50000) T = 1;
	if (T == 1) { fn((int64_t)0, n); return; }
	std::vector<std::thread> th;
	for (unsigned t = 0; t < T; ++t) {
		const int64_t lo = n * t / T, hi = n * (t + 1) / T;
		th.emplace_back([&fn, lo, hi]() { fn(lo, hi); });
	}
	for (auto &x : th) x.join();
}

void build_igl_normals(const double *V, int64_t nV, const int32_t *F, int64_t nF,
                       std::vector<double> &FN, std::vector<double> &VN, std::vector<double> &EN,
                       std::vector<int32_t> &E, std::vector<int32_t> &EMAP)
{
	FN.assign(3 * (size_t)nF, 0.0);
	VN.assign(3 * (size_t)nV, 0.0);
	// rows of dynamic-size Eigen matrices reduce sequentially: (x^2 + y^2) + z^2 (Eigen 3.2 DefaultTraversal/NoUnrolling)
	auto sqn = [](double x, double y, double z) { return (x * x + y * y) + z * z; };
	std::vector<double> W(3 * (size_t)nF);       // internal angles, igl/internal_angles.cpp:64-87
	parallel_ranges(nF, [&](int64_t lo, int64_t hi) {
		for (int64_t f = lo; f < hi; ++f) {
			const double *p0 = V + 3 * (int64_t)F[3 * f], *p1 = V + 3 * (int64_t)F[3 * f + 1], *p2 = V + 3 * (int64_t)F[3 * f + 2];
			const double a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
			const double b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
			double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
			const double r = std::sqrt(sqn(n[0], n[1], n[2]));
			if (r == 0) { n[0] = n[1] = n[2] = 0; } else { const double ir = 1.0 / r; n[0] *= ir; n[1] *= ir; n[2] *= ir; }
			for (int c = 0; c < 3; ++c) FN[3 * (size_t)f + c] = n[c];
			double L[3];
			L[0] = sqn(p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]);
			L[1] = sqn(p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]);
			L[2] = sqn(p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]);
			for (int d = 0; d < 3; ++d) {
				const double s1 = L[d], s2 = L[(d + 1) % 3], s3 = L[(d + 2) % 3];
				W[3 * (size_t)f + d] = std::acos((s3 + s2 - s1) / (2. * std::sqrt(s3 * s2)));
			}
		}
	});
	// N.row(F(i,j)) += W(i,j) * FN.row(i) for i ascending (igl/per_vertex_normals.cpp:66-73): each thread owns a vertex
	// range and walks ALL faces in order, so every vertex sees its contributions in the reference's order
	parallel_ranges(nV, [&](int64_t vlo, int64_t vhi) {
		for (int64_t f = 0; f < nF; ++f)
			for (int d = 0; d < 3; ++d) {
				const int64_t v = F[3 * f + d];
				if (v < vlo || v >= vhi) continue;
				const double w = W[3 * (size_t)f + d];
				double *vn = &VN[3 * (size_t)v];
				for (int c = 0; c < 3; ++c) vn[c] += w * FN[3 * (size_t)f + c];
			}
		for (int64_t v = vlo; v < vhi; ++v) {
			double *vn = &VN[3 * (size_t)v];
			// N.rowwise().normalize() is a cwiseQuotient by the row norm (Eigen/src/Core/VectorwiseOp.h:539-550): true division
			const double r = std::sqrt(sqn(vn[0], vn[1], vn[2]));
			vn[0] /= r; vn[1] /= r; vn[2] /= r;
		}
	});
	std::vector<double>().swap(W);
	// undirected edges: directed edge (f,c) = (F[(c+1)%3], F[(c+2)%3]), stored at row f + c*nF; unique rows in
	// lexicographic (min, max) order (igl/unique_simplices.cpp:16-33).  LSD radix sort (16-bit digits) on (min << b) | max.
	const size_t m = (size_t)nF;
	int vb = 1;
	while ((1ll << vb) < nV) ++vb;
	std::vector<uint64_t> key(3 * m);
	for (size_t f = 0; f < m; ++f)
		for (int c = 0; c < 3; ++c) {
			uint64_t a = (uint32_t)F[3 * f + (c + 1) % 3], b = (uint32_t)F[3 * f + (c + 2) % 3];
			if (a > b) std::swap(a, b);
			key[f + (size_t)c * m] = (a << vb) | b;
		}
	std::vector<uint64_t> uniq(key), tmp(3 * m);
	for (int shift = 0; shift < 2 * vb; shift += 16) {
		size_t cnt[65537] = {0};
		for (uint64_t k : uniq) ++cnt[((k >> shift) & 0xffff) + 1];
		for (int i = 0; i < 65536; ++i) cnt[i + 1] += cnt[i];
		for (uint64_t k : uniq) tmp[cnt[(k >> shift) & 0xffff]++] = k;
		uniq.swap(tmp);
	}
	uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
	const uint64_t lowmask = (1ull << vb) - 1;
	E.resize(2 * uniq.size());
	for (size_t e = 0; e < uniq.size(); ++e) { E[2 * e] = (int32_t)(uniq[e] >> vb); E[2 * e + 1] = (int32_t)(uniq[e] & lowmask); }
	EMAP.resize(3 * m);
	parallel_ranges((int64_t)(3 * m), [&](int64_t lo, int64_t hi) {
		for (int64_t i = lo; i < hi; ++i) EMAP[(size_t)i] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), key[(size_t)i]) - uniq.begin());
	});
	EN.assign(3 * uniq.size(), 0.0);
	for (size_t f = 0; f < m; ++f)                      // N.row(EMAP(f + c*m)) += FN.row(f), f then c (igl/per_edge_normals.cpp:63-75)
		for (int c = 0; c < 3; ++c) {
			double *en = &EN[3 * (size_t)EMAP[f + (size_t)c * m]];
			for (int k = 0; k < 3; ++k) en[k] += FN[3 * f + k];
		}
}

} // namespace fpohm

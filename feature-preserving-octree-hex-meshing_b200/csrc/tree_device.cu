// igl::AABB<MatrixXd,3>::init on the device (igl/AABB.cpp:30-200), node for node identical to the host builder
// (igl_tree_host.cpp) and therefore to igl.
//
// What makes that possible: igl splits at the MEDIAN RANK of the barycentres on the longest box axis, left = ranks <= median
// (AABB.cpp:154-183).  Ranks on an axis are distinct integers, so the left child holds exactly the ceil(n/2) elements of
// smallest rank — the SHAPE of the tree (sizes, DFS pre-order ids: left = me + 1, right = me + 2 * n_left) depends on the facet
// count alone and is laid out on the host in microseconds per level; only WHICH facet sits where depends on the data.  The
// build is level-synchronous: every level (a) reduces the node boxes over the elements of each segment (exact: min / max of
// the facets' vertex coordinates), (b) picks the node's axis (first strict maximum of the extents, Eigen maxCoeff), (c) sorts
// the elements of every segment by their rank on the node's axis — one radix sort of (segment, rank) keys — after which the
// children are the two halves of the segment.  21 levels at 2 M facets, ~1 ms each.
//
// What still needs the host: the rank of EQUAL barycentre coordinates is whatever libstdc++'s introsort leaves
// (igl/sort.cpp:281-300) and decides on which side of a median such facets fall.  The device sorts each axis with a stable
// radix sort and counts equal neighbours; only an axis that has ties takes its ranks from the host's std::sort
// (host_rank_axis).  Meshes without repeated coordinates (scans) never touch the host; the procedural benchmark meshes
// (translated copies of one torus, a symmetric gear) do, on all three axes.
#include "mesh.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <math_constants.h>
#include <algorithm>
#include <map>
#include <thread>

using namespace fpohm;

namespace {

__device__ __forceinline__ unsigned long long ord64(double x) {
	const unsigned long long u = (unsigned long long)__double_as_longlong(x);
	return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord64(unsigned long long u) {
	return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

// facet boxes (min / max of the three vertices, AABB.cpp:119-128 via the element's vertices) and barycentre columns
// (igl/barycenter.cpp: ((0 + a) + b) + c, then * (1.0 / 3.0): Eigen's operator/= multiplies by the reciprocal)
__global__ void facet_box_bary_kernel(const double *__restrict__ tri, int64_t nF, double *__restrict__ tbox,
                                      unsigned long long *__restrict__ kx, unsigned long long *__restrict__ ky, unsigned long long *__restrict__ kz,
                                      int32_t *__restrict__ idx)
{
	const double third = 1.0 / 3.0;
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const double *t = tri + 9 * f;
		double bc[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			const double a = t[d], b = t[3 + d], c = t[6 + d];
			tbox[6 * f + d] = fmin(a, fmin(b, c));
			tbox[6 * f + 3 + d] = fmax(a, fmax(b, c));
			double s = 0.0;
			s += a; s += b; s += c;
			bc[d] = s * third;
		}
		// -0.0 and +0.0 compare equal on the host: map both to one key so that they count as a tie
		kx[f] = ord64(bc[0] + 0.0); ky[f] = ord64(bc[1] + 0.0); kz[f] = ord64(bc[2] + 0.0);
		idx[f] = (int32_t)f;
	}
}
__global__ void rank_and_ties_kernel(const unsigned long long *__restrict__ skey, const int32_t *__restrict__ sidx, int64_t nF,
                                     int32_t *__restrict__ rank, int32_t *__restrict__ n_ties)
{
	int local = 0;
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nF; i += (int64_t)gridDim.x * blockDim.x) {
		rank[sidx[i]] = (int32_t)i;
		if (i + 1 < nF && skey[i] == skey[i + 1]) ++local;
	}
	if (local) atomicAdd(n_ties, local);
}

struct Seg { int32_t begin, count, id; };      // a node of the current level (or a leaf finished one level earlier), in position order

__device__ __forceinline__ int seg_of(const Seg *__restrict__ segs, int nseg, int32_t p) {
	int lo = 0, hi = nseg;                        // largest j with segs[j].begin <= p
	while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (segs[mid].begin <= p) lo = mid; else hi = mid; }
	return lo;
}

__global__ void seg_box_init_kernel(int nseg, unsigned long long *__restrict__ sbox) {
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nseg; j += gridDim.x * blockDim.x) {
#pragma unroll
		for (int c = 0; c < 3; ++c) { sbox[6 * (int64_t)j + c] = ~0ull; sbox[6 * (int64_t)j + 3 + c] = 0ull; }
	}
}
// union of the element boxes per segment; a warp whose 32 elements lie in one segment reduces first
__global__ void __launch_bounds__(256)
seg_box_kernel(const Seg *__restrict__ segs, int nseg, const int32_t *__restrict__ elem, const double *__restrict__ tbox, int64_t nF,
               unsigned long long *__restrict__ sbox)
{
	const int lane = threadIdx.x & 31;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < nF; base += stride) {
		const int64_t p = base + lane;
		const bool ok = p < nF;
		int j = -1;
		unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
		if (ok) {
			j = seg_of(segs, nseg, (int32_t)p);
			const double *b = tbox + 6 * (int64_t)elem[p];
#pragma unroll
			for (int c = 0; c < 3; ++c) { lo[c] = ord64(b[c]); hi[c] = ord64(b[3 + c]); }
		}
		const int j0 = __shfl_sync(0xffffffffu, j, 0);
		if (__all_sync(0xffffffffu, j == j0)) {
#pragma unroll
			for (int c = 0; c < 3; ++c) {
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) {
					const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo[c], o), b = __shfl_xor_sync(0xffffffffu, hi[c], o);
					lo[c] = a < lo[c] ? a : lo[c]; hi[c] = b > hi[c] ? b : hi[c];
				}
			}
			if (lane == 0 && j0 >= 0) {
#pragma unroll
				for (int c = 0; c < 3; ++c) { atomicMin(sbox + 6 * (int64_t)j0 + c, lo[c]); atomicMax(sbox + 6 * (int64_t)j0 + 3 + c, hi[c]); }
			}
		} else if (ok) {
#pragma unroll
			for (int c = 0; c < 3; ++c) { atomicMin(sbox + 6 * (int64_t)j + c, lo[c]); atomicMax(sbox + 6 * (int64_t)j + 3 + c, hi[c]); }
		}
	}
}
// node box out, split axis = first strict maximum of the extents (Eigen maxCoeff, AABB.cpp:143-144)
__global__ void seg_axis_kernel(const Seg *__restrict__ segs, int nseg, const unsigned long long *__restrict__ sbox, double *__restrict__ node_box,
                                int8_t *__restrict__ axis)
{
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nseg; j += gridDim.x * blockDim.x) {
		if (segs[j].id < 0) { axis[j] = 0; continue; }      // a leaf finished one level up: only keeps its place in the order
		double mn[3], mx[3];
#pragma unroll
		for (int c = 0; c < 3; ++c) { mn[c] = unord64(sbox[6 * (int64_t)j + c]); mx[c] = unord64(sbox[6 * (int64_t)j + 3 + c]); }
		double *o = node_box + 6 * (int64_t)segs[j].id;
#pragma unroll
		for (int c = 0; c < 3; ++c) { o[c] = mn[c]; o[3 + c] = mx[c]; }
		int d = 0;
		double best = mx[0] - mn[0];
		for (int c = 1; c < 3; ++c) { const double e = mx[c] - mn[c]; if (e > best) { best = e; d = c; } }
		axis[j] = (int8_t)d;
	}
}
__global__ void seg_keys_kernel(const Seg *__restrict__ segs, int nseg, const int8_t *__restrict__ axis, const int32_t *__restrict__ elem, int64_t nF,
                                const int32_t *__restrict__ r0, const int32_t *__restrict__ r1, const int32_t *__restrict__ r2,
                                unsigned long long *__restrict__ key)
{
	for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nF; p += (int64_t)gridDim.x * blockDim.x) {
		const int j = seg_of(segs, nseg, (int32_t)p);
		const int d = axis[j];
		const int32_t e = elem[p];
		const int32_t r = segs[j].count > 1 ? (d == 0 ? r0[e] : (d == 1 ? r1[e] : r2[e])) : 0;
		key[p] = ((unsigned long long)(unsigned)j << 32) | (unsigned)r;
	}
}
// next level's segments, on the device (the shape is a function of nF alone; round 2's first version enumerated every level on the
// host and uploaded it: ~50 MB of pageable copies and ~15 ms of host loops per 2 M facets, serial in front of everything else)
__global__ void seg_fanout_kernel(const Seg *__restrict__ segs, int nseg, int32_t *__restrict__ fan) {
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nseg; j += gridDim.x * blockDim.x) fan[j] = segs[j].count > 1 ? 2 : 1;
}
__global__ void seg_split_kernel(const Seg *__restrict__ segs, int nseg, const int32_t *__restrict__ off, Seg *__restrict__ out,
                                 int32_t *__restrict__ leaf_id_of_pos)
{
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nseg; j += gridDim.x * blockDim.x) {
		const Seg g = segs[j];
		const int32_t o = off[j];
		if (g.count > 1) {
			const int32_t nl = (g.count + 1) / 2;       // igl sizes the sides (n + 1) / 2 and n / 2 (AABB.cpp:168)
			out[o] = Seg{g.begin, nl, g.id + 1};
			out[o + 1] = Seg{g.begin + nl, g.count - nl, g.id + 2 * nl};
		} else {
			// a leaf stays in place as a segment of one element (it keeps its position in the order); id < 0: already written
			out[o] = Seg{g.begin, 1, g.id >= 0 ? -1 - g.id : g.id};
			if (g.id >= 0) leaf_id_of_pos[g.begin] = g.id;
		}
	}
}
__global__ void seg_leaves_kernel(const Seg *__restrict__ segs, int nseg, int32_t *__restrict__ leaf_id_of_pos) {
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nseg; j += gridDim.x * blockDim.x)
		if (segs[j].count == 1 && segs[j].id >= 0) leaf_id_of_pos[segs[j].begin] = segs[j].id;
}
__global__ void leaf_prim_kernel(const int32_t *__restrict__ leaf_id_of_pos, const int32_t *__restrict__ elem, int64_t nF, int32_t *__restrict__ prim) {
	for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nF; p += (int64_t)gridDim.x * blockDim.x) prim[leaf_id_of_pos[p]] = elem[p];
}

} // namespace

namespace fpohm {

void build_igl_tree_device(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s, int ties_host[3], const std::function<void()> &while_sorting) {
	const int64_t nF = m->nF;
	const size_t nn = 2 * (size_t)nF - 1;
	FPOHM_REQUIRE(nF < (1ll << 31), FPOHM_ERANGE, "tree build: %lld facets", (long long)nF);
	const int blk = 256;
	// ---- facet boxes, barycentre keys, ranks per axis ----
	DevBuf<double> tbox(6 * nF, s);
	DevBuf<unsigned long long> key[3] = {DevBuf<unsigned long long>(nF, s), DevBuf<unsigned long long>(nF, s), DevBuf<unsigned long long>(nF, s)};
	DevBuf<unsigned long long> skey(nF, s);
	DevBuf<int32_t> idx(nF, s), sidx(nF, s), n_ties(3, s);
	DevBuf<int32_t> rank[3] = {DevBuf<int32_t>(nF, s), DevBuf<int32_t>(nF, s), DevBuf<int32_t>(nF, s)};
	n_ties.zero();
	facet_box_bary_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(m->tri.p, nF, tbox.p, key[0].p, key[1].p, key[2].p, idx.p);
	FPOHM_LAUNCH_CHECK(ctx);
	size_t tb = 0;
	FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key[0].p, skey.p, idx.p, sidx.p, (int)nF, 0, 64, s));
	DevBuf<uint8_t> tmp((int64_t)tb, s);
	for (int d = 0; d < 3; ++d) {
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key[d].p, skey.p, idx.p, sidx.p, (int)nF, 0, 64, s));
		rank_and_ties_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(skey.p, sidx.p, nF, rank[d].p, n_ties.p + d);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	ctx->launches += 3;
	int32_t h_ties[3] = {0, 0, 0};
	n_ties.download(h_ties, 3);
	// ---- segments per level: counts only (the multiset of segment sizes of a level has two or three distinct values) ----
	std::vector<int64_t> nseg_of;                 // segments of level L, finished leaves included
	std::vector<char> splits_of;                  // does level L still hold a segment of more than one element
	{
		std::map<int32_t, int64_t> cur;           // size -> how many
		cur[(int32_t)nF] = 1;
		for (;;) {
			int64_t n = 0; bool any = false;
			std::map<int32_t, int64_t> next;
			for (const auto &kv : cur) {
				n += kv.second;
				if (kv.first > 1) { any = true; next[(kv.first + 1) / 2] += kv.second; next[kv.first / 2] += kv.second; }
				else next[1] += kv.second;
			}
			nseg_of.push_back(n); splits_of.push_back(any ? 1 : 0);
			if (!any) break;
			cur.swap(next);
		}
	}
	FPOHM_CUDA(cudaStreamSynchronize(s));
	// ---- axes with equal barycentre coordinates: ranks from the host's std::sort, one thread per axis ----
	{
		std::vector<std::vector<int32_t>> hr(3);
		std::vector<std::thread> th;
		for (int d = 0; d < 3; ++d) {
			ties_host[d] = h_ties[d] > 0 ? 1 : 0;
			if (ties_host[d]) { hr[(size_t)d].resize((size_t)nF); th.emplace_back(host_rank_axis, m->hV.data(), m->hF.data(), nF, d, hr[(size_t)d].data()); }
		}
		// the caller's independent work (normals, the wide tree's shape) runs here, under the host sorts; a throw must not leave
		// joinable threads behind
		try { if (while_sorting) while_sorting(); } catch (...) { for (auto &t : th) t.join(); throw; }
		for (auto &t : th) t.join();
		for (int d = 0; d < 3; ++d) if (ties_host[d]) rank[d].upload(hr[(size_t)d].data(), nF);
		FPOHM_CUDA(cudaStreamSynchronize(s));      // hr is a local
	}
	// ---- level-synchronous build ----
	DevBuf<int32_t> elem(nF, s), elem2(nF, s), d_leaf(nF, s);
	m->t_prim.alloc((int64_t)nn, s);
	m->t_box.alloc(6 * (int64_t)nn, s);
	DevBuf<int32_t> &prim = m->t_prim;
	DevBuf<double> &node_box = m->t_box;
	FPOHM_CUDA(cudaMemcpyAsync(elem.p, idx.p, 4 * (size_t)nF, cudaMemcpyDeviceToDevice, s));      // 0, 1, 2, ...
	FPOHM_CUDA(cudaMemsetAsync(prim.p, 0xff, 4 * nn, s));
	int64_t max_segs = 0;
	for (int64_t n : nseg_of) max_segs = std::max(max_segs, n);
	FPOHM_REQUIRE(max_segs < (1ll << 31), FPOHM_ERANGE, "tree build: %lld segments", (long long)max_segs);
	DevBuf<Seg> segs_a(max_segs, s), segs_b(max_segs, s);
	DevBuf<int32_t> fan(max_segs, s), fan_off(max_segs, s);
	DevBuf<unsigned long long> sbox(6 * max_segs, s), lkey(nF, s), lkey2(nF, s);
	DevBuf<int8_t> axis(max_segs, s);
	size_t tb2 = 0, tb3 = 0;
	FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, lkey.p, lkey2.p, elem.p, elem2.p, (int)nF, 0, 64, s));
	FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb3, fan.p, fan_off.p, (int)max_segs, s));
	DevBuf<uint8_t> tmp2((int64_t)std::max(tb2, tb3), s);
	const Seg root_seg{0, (int32_t)nF, 0};
	FPOHM_CUDA(cudaMemcpyAsync(segs_a.p, &root_seg, sizeof(Seg), cudaMemcpyHostToDevice, s));
	FPOHM_CUDA(cudaStreamSynchronize(s));              // root_seg is a local
	int32_t *e_in = elem.p, *e_out = elem2.p;
	Seg *sg_in = segs_a.p, *sg_out = segs_b.p;
	for (size_t L = 0; L < nseg_of.size(); ++L) {
		// segments of this level: nodes (id >= 0) and earlier leaves (id < 0, skipped by the node kernels through count == 1 and id)
		const int nseg = (int)nseg_of[L];
		seg_box_init_kernel<<<grid_for(ctx, nseg, blk), blk, 0, s>>>(nseg, sbox.p);
		FPOHM_LAUNCH_CHECK(ctx);
		seg_box_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(sg_in, nseg, e_in, tbox.p, nF, sbox.p);
		FPOHM_LAUNCH_CHECK(ctx);
		seg_axis_kernel<<<grid_for(ctx, nseg, blk), blk, 0, s>>>(sg_in, nseg, sbox.p, node_box.p, axis.p);
		FPOHM_LAUNCH_CHECK(ctx);
		if (!splits_of[L]) {
			seg_leaves_kernel<<<grid_for(ctx, nseg, blk), blk, 0, s>>>(sg_in, nseg, d_leaf.p);
			FPOHM_LAUNCH_CHECK(ctx);
			break;
		}
		seg_keys_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(sg_in, nseg, axis.p, e_in, nF, rank[0].p, rank[1].p, rank[2].p, lkey.p);
		FPOHM_LAUNCH_CHECK(ctx);
		int sbits = 1;
		while ((1ll << sbits) < nseg) ++sbits;
		size_t tbs = tb2;
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, tbs, lkey.p, lkey2.p, e_in, e_out, (int)nF, 0, 32 + sbits, s));
		std::swap(e_in, e_out);
		// next level's segments
		seg_fanout_kernel<<<grid_for(ctx, nseg, blk), blk, 0, s>>>(sg_in, nseg, fan.p);
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tbx = tb3;
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(tmp2.p, tbx, fan.p, fan_off.p, nseg, s));
		seg_split_kernel<<<grid_for(ctx, nseg, blk), blk, 0, s>>>(sg_in, nseg, fan_off.p, sg_out, d_leaf.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ctx->launches += 2;
		std::swap(sg_in, sg_out);
	}
	leaf_prim_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(d_leaf.p, e_in, nF, prim.p);
	FPOHM_LAUNCH_CHECK(ctx);
	FPOHM_CUDA(cudaStreamSynchronize(s));
}

} // namespace fpohm

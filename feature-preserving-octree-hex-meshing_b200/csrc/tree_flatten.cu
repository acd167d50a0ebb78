// Flattenings of the igl tree for the query kernels, on the device: QNode (fp64, both child boxes per node), QNodeF (fp32,
// rounded outwards), prim_parent, the (parent, depth) table of the tie-break and the 8-wide collapse (WNode) — from the DFS
// pre-order arrays (box, prim) the tree build leaves in HBM.  Round 1 did all of this in host loops over 4 M nodes and
// uploaded ~0.5 GB from pageable memory: 650 ms of a 1.4 s build at 2 M facets (FPOHM_TREE_TIMELINE=1).
//
// The SHAPE of igl's tree depends on the facet count alone (median splits on distinct ranks: the left child holds ceil(n/2)
// elements, DFS pre-order ids left = me + 1, right = me + 2 * n_left; igl/AABB.cpp:154-183), so a node finds its parent, depth,
// element count and its index among the internal nodes by walking down from the root with integer arithmetic — no scans, no
// pointer tables.  The same holds for the 8-wide collapse (the child with the most facets is opened until there are eight): its
// structure is laid out on the host from counts in a few ms and filled with boxes and facet ids by one kernel.
#include "mesh.h"

#include <math_constants.h>
#include <algorithm>
#include <vector>

using namespace fpohm;

namespace {

struct ShapePos { int32_t parent_int, iid, depth, cnt; };
// node `i` of the DFS pre-order of a tree over nF facets
__device__ __forceinline__ ShapePos shape_of(int64_t nF, int32_t i) {
	int32_t id = 0, cnt = (int32_t)nF, iid = 0, depth = 0, parent_int = -1;
	while (id != i) {
		const int32_t nl = (cnt + 1) / 2;
		parent_int = iid;
		if (i < id + 2 * nl) { id += 1; iid += 1; cnt = nl; }
		else { id += 2 * nl; iid += nl; cnt -= nl; }
		++depth;
	}
	return {parent_int, iid, depth, cnt};
}

__global__ void flatten_kernel(int64_t nF, const double *__restrict__ box, const int32_t *__restrict__ prim,
                               QNode *__restrict__ q, QNodeF *__restrict__ qf, int32_t *__restrict__ prim_parent, int2 *__restrict__ pd)
{
	const int64_t nn = 2 * nF - 1;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nn; t += (int64_t)gridDim.x * blockDim.x) {
		const int32_t i = (int32_t)t;
		const ShapePos sp = shape_of(nF, i);
		if (sp.cnt <= 1) continue;                                  // a leaf: lives in its parent's node
		const int32_t nl = (sp.cnt + 1) / 2, l = i + 1, r = i + 2 * nl;
		QNode n;
		QNodeF f;
#pragma unroll
		for (int c = 0; c < 3; ++c) {
			n.lmin[c] = box[6 * (int64_t)l + c]; n.lmax[c] = box[6 * (int64_t)l + 3 + c];
			n.rmin[c] = box[6 * (int64_t)r + c]; n.rmax[c] = box[6 * (int64_t)r + 3 + c];
			f.lmin[c] = __double2float_rd(n.lmin[c]); f.lmax[c] = __double2float_ru(n.lmax[c]);
			f.rmin[c] = __double2float_rd(n.rmin[c]); f.rmax[c] = __double2float_ru(n.rmax[c]);
		}
		n.left = nl == 1 ? ~prim[l] : sp.iid + 1;
		n.right = (sp.cnt - nl) == 1 ? ~prim[r] : sp.iid + nl;
		n.parent = sp.parent_int; n.depth = sp.depth; n.pad[0] = n.pad[1] = 0;
		f.left = n.left; f.right = n.right; f.pad[0] = f.pad[1] = 0;
		q[sp.iid] = n;
		qf[sp.iid] = f;
		pd[sp.iid] = make_int2(sp.parent_int, sp.depth);
		if (n.left < 0) prim_parent[~n.left] = sp.iid;
		if (n.right < 0) prim_parent[~n.right] = sp.iid;
	}
}

__global__ void wide_fill_kernel(int64_t n_wide, const WideKid *__restrict__ kids, const double *__restrict__ box, const int32_t *__restrict__ prim,
                                 WNode *__restrict__ w)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 8 * n_wide; t += (int64_t)gridDim.x * blockDim.x) {
		const WideKid k = kids[t];
		WChild e;
		if (k.cnt > 0) {
			const double *b = box + 6 * (int64_t)k.dfs;
#pragma unroll
			for (int a = 0; a < 3; ++a) { e.lo[a] = __double2float_rd(b[a]); e.hi[a] = __double2float_ru(b[3 + a]); }
			if (k.cnt == 1) { e.child = ~prim[k.dfs]; e.flags = 0; }
			else { e.child = k.wide; e.flags = k.cnt <= 8 ? 1 : 0; }
		} else {
#pragma unroll
			for (int a = 0; a < 3; ++a) { e.lo[a] = CUDART_INF_F; e.hi[a] = -CUDART_INF_F; }
			e.child = WCHILD_EMPTY; e.flags = 0;
		}
		w[t >> 3].c[t & 7] = e;
	}
}
__global__ void float_triangles_kernel2(const double *__restrict__ tri, int64_t nF, float4 *__restrict__ out) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 3 * nF; t += (int64_t)gridDim.x * blockDim.x) {
		const double *v = tri + 3 * t;
		out[t] = make_float4((float)v[0], (float)v[1], (float)v[2], 0.f);
	}
}
// max over the vertices of |v - float(v)|^2 and of |coordinate| (both >= 0: their bit patterns order like the values)
__global__ void vertex_rounding_kernel(const double *__restrict__ V, int64_t nV, unsigned long long *__restrict__ out) {
	double e2m = 0, mc = 0;
	for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
		double e2 = 0;
#pragma unroll
		for (int a = 0; a < 3; ++a) { const double x = V[3 * v + a], e = x - (double)(float)x; e2 += e * e; mc = fmax(mc, fabs(x)); }
		e2m = fmax(e2m, e2);
	}
	for (int o = 16; o > 0; o >>= 1) { e2m = fmax(e2m, __shfl_xor_sync(0xffffffffu, e2m, o)); mc = fmax(mc, __shfl_xor_sync(0xffffffffu, mc, o)); }
	if ((threadIdx.x & 31) == 0) { atomicMax(out, (unsigned long long)__double_as_longlong(e2m)); atomicMax(out + 1, (unsigned long long)__double_as_longlong(mc)); }
}

} // namespace

namespace fpohm {

// box / prim: device arrays of the 2 nF - 1 DFS pre-order nodes
void wide_shape_host(int64_t nF, std::vector<WideKid> &kids, int32_t &n_wide_out) {
	kids.clear(); n_wide_out = 0;
	if (nF < 2) return;
	const int64_t ni = nF - 1;
	kids.reserve((size_t)ni / 2 + 64);
	struct Work { int32_t dfs, cnt, wide; };
	std::vector<Work> work;
	int32_t n_wide = 1;
	kids.resize(8);
	work.push_back({0, (int32_t)nF, 0});
	while (!work.empty()) {
		const Work b = work.back();
		work.pop_back();
		int32_t kid[8], kc[8];
		int nk = 2;
		const int32_t nl = (b.cnt + 1) / 2;
		kid[0] = b.dfs + 1; kc[0] = nl; kid[1] = b.dfs + 2 * nl; kc[1] = b.cnt - nl;
		while (nk < 8) {
			int best = -1;
			for (int k = 0; k < nk; ++k) if (kc[k] > 1 && (best < 0 || kc[k] > kc[best])) best = k;
			if (best < 0) break;
			const int32_t o = kid[best], oc = kc[best], onl = (oc + 1) / 2;
			for (int k = nk; k > best + 1; --k) { kid[k] = kid[k - 1]; kc[k] = kc[k - 1]; }      // keep the binary tree's left-to-right order
			kid[best] = o + 1; kc[best] = onl; kid[best + 1] = o + 2 * onl; kc[best + 1] = oc - onl;
			++nk;
		}
		for (int k = 0; k < 8; ++k) {
			WideKid e{0, 0, 0};
			if (k < nk) {
				e.dfs = kid[k]; e.cnt = kc[k];
				if (kc[k] > 1) {
					e.wide = n_wide++;
					kids.resize(8 * (size_t)n_wide);
					work.push_back({kid[k], kc[k], e.wide});
				}
			}
			kids[8 * (size_t)b.wide + (size_t)k] = e;
		}
	}
	n_wide_out = n_wide;
}

void flatten_tree_device(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s, const double *box, const int32_t *prim, const std::vector<WideKid> &kids, int32_t n_wide) {
	const int64_t nF = m->nF, nn = 2 * nF - 1, ni = nF - 1;
	const int blk = 256;
	m->n_qnodes = ni;
	m->qnodes.alloc(std::max<int64_t>(ni, 1), s);
	m->qfnodes.alloc(std::max<int64_t>(ni, 1), s);
	m->node_pd.alloc(std::max<int64_t>(ni, 1), s);
	m->prim_parent.alloc(std::max<int64_t>(nF, 1), s);
	FPOHM_CUDA(cudaMemsetAsync(m->prim_parent.p, 0xff, 4 * (size_t)std::max<int64_t>(nF, 1), s));
	if (ni > 0) {
		flatten_kernel<<<grid_for(ctx, nn, blk), blk, 0, s>>>(nF, box, prim, m->qnodes.p, m->qfnodes.p, m->prim_parent.p, m->node_pd.p);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	// deepest internal node: the path through the larger (left) children
	{ int32_t c = (int32_t)nF, d = 0; while ((c + 1) / 2 > 1) { c = (c + 1) / 2; ++d; } m->qdepth = nF > 1 ? d : 0; }
	m->n_wnodes = 0;
	if (ni > 0) {
		m->n_wnodes = n_wide;
		m->wnodes.alloc(n_wide, s);
		DevBuf<WideKid> dk(8 * (int64_t)n_wide, s);
		dk.upload(kids.data(), 8 * (int64_t)n_wide);
		wide_fill_kernel<<<grid_for(ctx, 8 * (int64_t)n_wide, blk), blk, 0, s>>>(n_wide, dk.p, box, prim, m->wnodes.p);
		FPOHM_LAUNCH_CHECK(ctx);
		m->trif.alloc(3 * nF, s);
		float_triangles_kernel2<<<grid_for(ctx, 3 * nF, blk), blk, 0, s>>>(m->tri.p, nF, m->trif.p);
		FPOHM_LAUNCH_CHECK(ctx);
		DevBuf<unsigned long long> vr(2, s);
		vr.zero();
		vertex_rounding_kernel<<<grid_for(ctx, m->nV, blk, 4), blk, 0, s>>>(m->V.p, m->nV, vr.p);
		FPOHM_LAUNCH_CHECK(ctx);
		unsigned long long h[2] = {0, 0};
		vr.download(h, 2);
		FPOHM_CUDA(cudaStreamSynchronize(s));      // kids (host) and h
		double ev, mc;
		memcpy(&ev, &h[0], 8); memcpy(&mc, &h[1], 8);
		auto f_up = [](double x) { float f = (float)x; if ((double)f < x) f = std::nextafterf(f, INFINITY); return f; };
		m->eps_v = f_up(std::sqrt(ev) * 1.000001);
		m->slack_q = f_up(36.0 * 5.9604644775390625e-8 * mc * 1.01);
	}
}

} // namespace fpohm

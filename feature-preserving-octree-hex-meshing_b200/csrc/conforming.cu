// conforming_mesh (grid_meshing/grid_hex_meshing.cpp:568-696): the octree hex mesh with its T-junctions made conforming.
//
// The reference walks vectors-of-vectors on one thread; everything it does is local to a node, a face, an edge or a hex:
//   (1) :578-622  a node whose octree links miss exactly one axis direction pair while the other four exist is the centre
//                 of a big face: the 4 small faces around it (ascending face id = order of Vs[i].neighbor_fs) replace the
//                 big face found through the smallest of the 4 outer corners;
//   (2) :626-645  each edge of a replaced face gets the one common neighbour of its end points as mid vertex (if that is
//                 one of the 4 in-plane neighbours of the centre);
//   (3) :646-662  every face adjacent to such an edge gets the mid vertex inserted into its loop;
//   (4) :664-693  faces minus the replaced ones (ids compacted in order), hexes with [kept faces in order] + [4 small faces
//                 of each replaced face, ascending face id], sorted unique vertex set per hex;
//   (5) gf.cpp:187-264 (Hyb branch): face boundary flags, edges = unique sorted (vmin, vmax) over the loops in
//                 (v0, v1, face, j) order, per-loop-slot edge ids, boundary edges/vertices, F.neighbor_hs.
// One kernel per step, counts -> exclusive scan -> fill for the variable-length outputs.  Integer work only: bit-exact
// against the reference run on the same numbering (tests: oracle/ref/ref_driver_ghm.cpp feeds it the product's tables).
#include "conn.h"
#include "octree.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

using namespace fpohm;

struct fpohm_hybrid {
	fpohm_ctx *ctx = nullptr;
	int64_t nV = 0, nF = 0, nH = 0, nE = 0, tot_fv = 0, tot_hf = 0, tot_hv = 0, tot_fn = 0;
	int64_t n_replaced = 0;
	DevBuf<int64_t> F_off, H_foff, H_voff, F_nhoff;
	DevBuf<uint32_t> F_vs, F_es, E_vs, H_fs, H_vs, F_nhs;
	DevBuf<uint8_t> F_boundary, E_boundary, V_boundary;
};

namespace {

__device__ __forceinline__ bool in4(const int32_t *a, int32_t v) { return a[0] == v || a[1] == v || a[2] == v || a[3] == v; }

// (1) T-node -> (big face, 4 small faces, 5 "corvs")
__global__ void __launch_bounds__(128)
tnode_kernel(const int32_t *__restrict__ node_pos, const int32_t *__restrict__ node_neigh, int64_t n_nodes, int gx, int gy, int gz,
             const uint8_t *__restrict__ V_boundary, const int64_t *__restrict__ vf_off, const uint32_t *__restrict__ vf_val,
             const uint32_t *__restrict__ F_vs, int32_t *__restrict__ rel /*4 per face, -1*/, int32_t *__restrict__ corv /*5 per face*/)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
		if (!V_boundary[i]) continue;
		const int px = node_pos[3 * i], py = node_pos[3 * i + 1], pz = node_pos[3 * i + 2];
		if (px == 0 || px == gx || py == 0 || py == gy || pz == 0 || pz == gz) continue;
		int32_t nn[6];
		int pos0 = -1;
		for (int k = 0; k < 6; ++k) { nn[k] = node_neigh[6 * i + k]; if (nn[k] < 0 && pos0 < 0) pos0 = k; }
		if (pos0 < 0) continue;
		const int pos1 = pos0 ^ 1;
		int32_t vs[4]; int m = 0; bool ok = true;
		for (int k = 0; k < 6; ++k) if (k != pos0 && k != pos1) { vs[m++] = nn[k]; ok &= nn[k] >= 0; }
		if (!ok) continue;
		for (int a = 1; a < 4; ++a) { const int32_t v = vs[a]; int b = a - 1; while (b >= 0 && vs[b] > v) { vs[b + 1] = vs[b]; --b; } vs[b + 1] = v; }
		int32_t fs4[4], corner[8]; int nf = 0, nc = 0;
		for (int64_t q = vf_off[i]; q < vf_off[i + 1]; ++q) {
			const uint32_t f = vf_val[q];
			int shared = 0;
			for (int k = 0; k < 4; ++k) shared += in4(vs, (int32_t)F_vs[4 * (int64_t)f + k]);
			if (shared != 2) continue;
			if (nf < 4) fs4[nf] = (int32_t)f;
			++nf;
			for (int k = 0; k < 4; ++k) {
				const int32_t v = (int32_t)F_vs[4 * (int64_t)f + k];
				if (v == (int32_t)i || in4(vs, v)) continue;
				bool seen = false;
				for (int c = 0; c < nc; ++c) seen |= corner[c] == v;
				if (!seen && nc < 8) corner[nc++] = v;
			}
		}
		if (nf != 4 || nc != 4) continue;              // (a different count never passes the reference's size() == 4 tests)
		int32_t cmin = corner[0];
		for (int c = 1; c < 4; ++c) cmin = min(cmin, corner[c]);
		int32_t ff = -1;
		for (int64_t q = vf_off[cmin]; q < vf_off[cmin + 1]; ++q) {
			const uint32_t f = vf_val[q];
			int shared = 0;
			for (int k = 0; k < 4; ++k) shared += in4(corner, (int32_t)F_vs[4 * (int64_t)f + k]);
			if (shared == 4) ff = (int32_t)f;
		}
		if (ff < 0) continue;
		for (int k = 0; k < 4; ++k) { rel[4 * (int64_t)ff + k] = fs4[k]; corv[5 * (int64_t)ff + k] = vs[k]; }
		corv[5 * (int64_t)ff + 4] = (int32_t)i;
	}
}

// (2) mid vertices of the edges of replaced faces
__global__ void __launch_bounds__(128)
midv_kernel(const int32_t *__restrict__ rel, const int32_t *__restrict__ corv, int64_t nF, const uint32_t *__restrict__ F_es,
            const uint32_t *__restrict__ E_vs, const int64_t *__restrict__ vv_off, const uint32_t *__restrict__ vv_val,
            int32_t *__restrict__ e_midv)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 4 * nF; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t f = t >> 2;
		if (rel[4 * f] < 0) continue;
		const uint32_t e = F_es[t];
		const uint32_t v0 = E_vs[2 * (int64_t)e], v1 = E_vs[2 * (int64_t)e + 1];
		int common = 0; uint32_t who = 0;
		for (int64_t a = vv_off[v0]; a < vv_off[v0 + 1]; ++a) {
			const uint32_t x = vv_val[a];
			for (int64_t b = vv_off[v1]; b < vv_off[v1 + 1]; ++b) if (vv_val[b] == x) { ++common; who = x; }
		}
		if (common != 1) continue;
		const int32_t *cv = corv + 5 * f;
		if (cv[0] == (int32_t)who || cv[1] == (int32_t)who || cv[2] == (int32_t)who || cv[3] == (int32_t)who || cv[4] == (int32_t)who) e_midv[e] = (int32_t)who;
	}
}

// edge of face f joining a and b (one of its 4 edges)
__device__ __forceinline__ int32_t face_edge_mid(const uint32_t *__restrict__ F_es, const uint32_t *__restrict__ E_vs, const int32_t *__restrict__ e_midv,
                                                 int64_t f, uint32_t a, uint32_t b)
{
	for (int k = 0; k < 4; ++k) {
		const uint32_t e = F_es[4 * f + k];
		const uint32_t x = E_vs[2 * (int64_t)e], y = E_vs[2 * (int64_t)e + 1];
		if ((x == a && y == b) || (x == b && y == a)) return e_midv[e];
	}
	return -1;
}

// (3)+(4a) loop sizes of the kept faces
__global__ void loop_size_kernel(const int32_t *__restrict__ rel, int64_t nF, const uint32_t *__restrict__ F_vs, const uint32_t *__restrict__ F_es,
                                 const uint32_t *__restrict__ E_vs, const int32_t *__restrict__ e_midv, int64_t *__restrict__ keep, int64_t *__restrict__ size)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f <= nF; f += (int64_t)gridDim.x * blockDim.x) {
		if (f == nF) { keep[f] = 0; size[f] = 0; continue; }
		if (rel[4 * f] >= 0) { keep[f] = 0; size[f] = 0; continue; }
		int n = 4;
		for (int k = 0; k < 4; ++k) n += face_edge_mid(F_es, E_vs, e_midv, f, F_vs[4 * f + k], F_vs[4 * f + ((k + 1) & 3)]) >= 0;
		keep[f] = 1; size[f] = n;
	}
}
__global__ void loop_fill_kernel(const int32_t *__restrict__ rel, int64_t nF, const uint32_t *__restrict__ F_vs, const uint32_t *__restrict__ F_es,
                                 const uint32_t *__restrict__ E_vs, const int32_t *__restrict__ e_midv, const int64_t *__restrict__ fmap,
                                 const int64_t *__restrict__ loop_off, int64_t *__restrict__ F_off_new, uint32_t *__restrict__ F_vs_new)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		if (rel[4 * f] >= 0) continue;
		int64_t o = loop_off[f];
		F_off_new[fmap[f]] = o;
		for (int k = 0; k < 4; ++k) {
			const uint32_t a = F_vs[4 * f + k], b = F_vs[4 * f + ((k + 1) & 3)];
			F_vs_new[o++] = a;
			const int32_t m = face_edge_mid(F_es, E_vs, e_midv, f, a, b);
			if (m >= 0) F_vs_new[o++] = (uint32_t)m;
		}
	}
}

// (4b) hexes: face lists and vertex sets
#define HYB_MAX_HV 64
__device__ __forceinline__ int hex_faces(const int32_t *__restrict__ rel, const uint32_t *__restrict__ H_fs, int64_t h, int32_t *out /*24*/) {
	int n = 0;
	uint32_t repl[6]; int nr = 0;
	for (int k = 0; k < 6; ++k) {
		const uint32_t f = H_fs[6 * h + k];
		if (rel[4 * (int64_t)f] >= 0) repl[nr++] = f; else out[n++] = (int32_t)f;
	}
	for (int a = 1; a < nr; ++a) { const uint32_t v = repl[a]; int b = a - 1; while (b >= 0 && repl[b] > v) { repl[b + 1] = repl[b]; --b; } repl[b + 1] = v; }
	for (int a = 0; a < nr; ++a) for (int k = 0; k < 4; ++k) out[n++] = rel[4 * (int64_t)repl[a] + k];
	return n;
}
template <bool FILL>
__global__ void __launch_bounds__(128)
hex_lists_kernel(const int32_t *__restrict__ rel, const uint32_t *__restrict__ H_fs, int64_t nH, const int64_t *__restrict__ fmap,
                 const int64_t *__restrict__ F_off_new, const uint32_t *__restrict__ F_vs_new,
                 int64_t *__restrict__ nf_out, int64_t *__restrict__ nv_out, const int64_t *__restrict__ foff, const int64_t *__restrict__ voff,
                 uint32_t *__restrict__ H_fs_new, uint32_t *__restrict__ H_vs_new, int32_t *__restrict__ overflow)
{
	for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h <= nH; h += (int64_t)gridDim.x * blockDim.x) {
		if (h == nH) { if (!FILL) { nf_out[h] = 0; nv_out[h] = 0; } continue; }
		int32_t fs[24];
		const int nf = hex_faces(rel, H_fs, h, fs);
		uint32_t vs[HYB_MAX_HV]; int nv = 0;
		for (int a = 0; a < nf; ++a) {
			const int64_t g = fmap[fs[a]];
			for (int64_t q = F_off_new[g]; q < F_off_new[g + 1]; ++q) {
				const uint32_t v = F_vs_new[q];
				int b = 0;
				while (b < nv && vs[b] < v) ++b;                       // sorted insert, unique
				if (b < nv && vs[b] == v) continue;
				if (nv == HYB_MAX_HV) { atomicExch(overflow, 1); continue; }
				for (int c = nv; c > b; --c) vs[c] = vs[c - 1];
				vs[b] = v; ++nv;
			}
		}
		if (!FILL) { nf_out[h] = nf; nv_out[h] = nv; }
		else {
			for (int a = 0; a < nf; ++a) H_fs_new[foff[h] + a] = (uint32_t)fmap[fs[a]];
			for (int a = 0; a < nv; ++a) H_vs_new[voff[h] + a] = vs[a];
		}
	}
}

// (5) hybrid connectivity
__global__ void face_hex_count_kernel(const uint32_t *__restrict__ H_fs_new, const int64_t *__restrict__ foff, int64_t nH, int64_t *__restrict__ cnt) {
	for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < nH; h += (int64_t)gridDim.x * blockDim.x)
		for (int64_t q = foff[h]; q < foff[h + 1]; ++q) atomicAdd((unsigned long long *)&cnt[H_fs_new[q]], 1ull);
}
__global__ void face_hex_fill_kernel(const uint32_t *__restrict__ H_fs_new, const int64_t *__restrict__ foff, int64_t nH,
                                     const int64_t *__restrict__ nh_off, int32_t *__restrict__ cursor, uint32_t *__restrict__ F_nhs)
{
	for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < nH; h += (int64_t)gridDim.x * blockDim.x)
		for (int64_t q = foff[h]; q < foff[h + 1]; ++q) {
			const uint32_t f = H_fs_new[q];
			F_nhs[nh_off[f] + atomicAdd(&cursor[f], 1)] = (uint32_t)h;
		}
}
// neighbour hexes are listed in ascending hex id (the reference pushes while looping over the hexes, gf.cpp:227-230)
__global__ void face_hex_sort_kernel(const int64_t *__restrict__ nh_off, int64_t nF, uint32_t *__restrict__ F_nhs, uint8_t *__restrict__ F_boundary) {
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const int64_t a = nh_off[f], n = nh_off[f + 1] - a;
		for (int64_t i = 1; i < n; ++i) { const uint32_t v = F_nhs[a + i]; int64_t j = i - 1; while (j >= 0 && F_nhs[a + j] > v) { F_nhs[a + j + 1] = F_nhs[a + j]; --j; } F_nhs[a + j + 1] = v; }
		F_boundary[f] = n == 2 ? 0 : 1;
	}
}
__global__ void loop_edge_keys_kernel(const int64_t *__restrict__ F_off_new, int64_t nF, const uint32_t *__restrict__ F_vs_new,
                                      uint64_t *__restrict__ key, uint32_t *__restrict__ slot)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const int64_t a = F_off_new[f], n = F_off_new[f + 1] - a;
		for (int64_t j = 0; j < n; ++j) {
			uint32_t v0 = F_vs_new[a + j], v1 = F_vs_new[a + (j + 1) % n];
			if (v0 > v1) { const uint32_t t = v0; v0 = v1; v1 = t; }
			key[a + j] = ((uint64_t)v0 << 32) | v1;
			slot[a + j] = (uint32_t)(a + j);
		}
	}
}
__global__ void edge_heads_kernel(const uint64_t *__restrict__ k, int64_t n, int64_t *__restrict__ head) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		head[t] = (t == 0 || k[t] != k[t - 1]) ? 1 : 0;
}
__global__ void edge_assign_kernel(const uint64_t *__restrict__ k, const uint32_t *__restrict__ slot, const int64_t *__restrict__ head,
                                   const int64_t *__restrict__ eid_incl, int64_t n, uint32_t *__restrict__ E_vs, uint32_t *__restrict__ F_es_new)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t e = eid_incl[t] - 1;
		F_es_new[slot[t]] = (uint32_t)e;
		if (head[t]) { E_vs[2 * e] = (uint32_t)(k[t] >> 32); E_vs[2 * e + 1] = (uint32_t)(k[t] & 0xffffffffu); }
	}
}
__global__ void boundary_marks_kernel(const int64_t *__restrict__ F_off_new, int64_t nF, const uint8_t *__restrict__ F_boundary,
                                      const uint32_t *__restrict__ F_es_new, const uint32_t *__restrict__ E_vs, uint8_t *__restrict__ E_boundary,
                                      uint8_t *__restrict__ V_boundary)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		if (!F_boundary[f]) continue;
		for (int64_t q = F_off_new[f]; q < F_off_new[f + 1]; ++q) {
			const uint32_t e = F_es_new[q];
			E_boundary[e] = 1;
			V_boundary[E_vs[2 * (int64_t)e]] = 1; V_boundary[E_vs[2 * (int64_t)e + 1]] = 1;
		}
	}
}

void exclusive_scan(fpohm_ctx *ctx, cudaStream_t s, const int64_t *in, int64_t *out, int64_t n) {
	size_t tb = 0;
	FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s));
	DevBuf<uint8_t> tmp((int64_t)tb, s);
	FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n, s));
	ctx->launches += 1;
}
int64_t last_of(const int64_t *dev, int64_t n, cudaStream_t s) {   // dev[n] after an exclusive scan over n + 1 entries
	int64_t v = 0;
	FPOHM_CUDA(cudaMemcpyAsync(&v, dev + n, 8, cudaMemcpyDeviceToHost, s));
	FPOHM_CUDA(cudaStreamSynchronize(s));
	return v;
}

} // namespace

static int conforming_impl(fpohm_ctx *ctx, const int32_t *node_pos_dev, const int32_t *node_neigh_dev, int64_t n_nodes, const int32_t gs[3],
                           const fpohm_conn *conn, fpohm_hybrid **out)
{
	FPOHM_API_BEGIN
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	const int64_t nF = conn->nF, nE = conn->nE, nH = conn->H, nV = conn->nV;
	fpohm_hybrid *hy = new fpohm_hybrid;
	try {
		hy->ctx = ctx; hy->nV = nV; hy->nH = nH;
		KernelTimer timer(ctx, s);
		DevBuf<int32_t> rel(4 * nF, s), corv(5 * nF, s), e_midv(std::max<int64_t>(nE, 1), s);
		FPOHM_CUDA(cudaMemsetAsync(rel.p, 0xff, 16 * (size_t)nF, s));
		FPOHM_CUDA(cudaMemsetAsync(e_midv.p, 0xff, 4 * (size_t)nE, s));
		tnode_kernel<<<grid_for(ctx, nV, 128), 128, 0, s>>>(node_pos_dev, node_neigh_dev, nV, gs[0], gs[1], gs[2], conn->V_boundary.p, conn->off[5].p, conn->val[5].p, conn->F_vs.p, rel.p, corv.p);
		FPOHM_LAUNCH_CHECK(ctx);
		midv_kernel<<<grid_for(ctx, 4 * nF, 128), 128, 0, s>>>(rel.p, corv.p, nF, conn->F_es.p, conn->E_vs.p, conn->off[3].p, conn->val[3].p, e_midv.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// faces
		DevBuf<int64_t> keep(nF + 1, s), size(nF + 1, s), fmap(nF + 1, s), loop_off(nF + 1, s);
		loop_size_kernel<<<grid_for(ctx, nF + 1, blk), blk, 0, s>>>(rel.p, nF, conn->F_vs.p, conn->F_es.p, conn->E_vs.p, e_midv.p, keep.p, size.p);
		FPOHM_LAUNCH_CHECK(ctx);
		exclusive_scan(ctx, s, keep.p, fmap.p, nF + 1);
		exclusive_scan(ctx, s, size.p, loop_off.p, nF + 1);
		hy->nF = last_of(fmap.p, nF, s);
		hy->tot_fv = last_of(loop_off.p, nF, s);
		hy->n_replaced = nF - hy->nF;
		hy->F_off.alloc(hy->nF + 1, s); hy->F_vs.alloc(hy->tot_fv, s); hy->F_es.alloc(hy->tot_fv, s);
		FPOHM_CUDA(cudaMemcpyAsync(hy->F_off.p + hy->nF, &hy->tot_fv, 8, cudaMemcpyHostToDevice, s));
		loop_fill_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(rel.p, nF, conn->F_vs.p, conn->F_es.p, conn->E_vs.p, e_midv.p, fmap.p, loop_off.p,
			hy->F_off.p, hy->F_vs.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// hexes
		DevBuf<int64_t> nf(nH + 1, s), nv(nH + 1, s);
		DevBuf<int32_t> ovf(1, s);
		ovf.zero();
		hy->H_foff.alloc(nH + 1, s); hy->H_voff.alloc(nH + 1, s);
		hex_lists_kernel<false><<<grid_for(ctx, nH + 1, 128), 128, 0, s>>>(rel.p, conn->H_fs.p, nH, fmap.p, hy->F_off.p, hy->F_vs.p, nf.p, nv.p,
			nullptr, nullptr, nullptr, nullptr, ovf.p);
		FPOHM_LAUNCH_CHECK(ctx);
		exclusive_scan(ctx, s, nf.p, hy->H_foff.p, nH + 1);
		exclusive_scan(ctx, s, nv.p, hy->H_voff.p, nH + 1);
		hy->tot_hf = last_of(hy->H_foff.p, nH, s);
		hy->tot_hv = last_of(hy->H_voff.p, nH, s);
		int32_t ov = 0;
		ovf.download(&ov, 1);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		FPOHM_REQUIRE(ov == 0, FPOHM_ERANGE, "fpohm_conforming_mesh: a cell has more than %d vertices", HYB_MAX_HV);
		hy->H_fs.alloc(hy->tot_hf, s); hy->H_vs.alloc(hy->tot_hv, s);
		hex_lists_kernel<true><<<grid_for(ctx, nH + 1, 128), 128, 0, s>>>(rel.p, conn->H_fs.p, nH, fmap.p, hy->F_off.p, hy->F_vs.p, nullptr, nullptr,
			hy->H_foff.p, hy->H_voff.p, hy->H_fs.p, hy->H_vs.p, ovf.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// hybrid connectivity: face <-> hex
		{
			DevBuf<int64_t> cnt(hy->nF + 1, s);
			cnt.zero();
			face_hex_count_kernel<<<grid_for(ctx, nH, blk), blk, 0, s>>>(hy->H_fs.p, hy->H_foff.p, nH, cnt.p);
			FPOHM_LAUNCH_CHECK(ctx);
			hy->F_nhoff.alloc(hy->nF + 1, s);
			exclusive_scan(ctx, s, cnt.p, hy->F_nhoff.p, hy->nF + 1);
			hy->tot_fn = hy->tot_hf;
			hy->F_nhs.alloc(hy->tot_fn, s);
			DevBuf<int32_t> cursor(std::max<int64_t>(hy->nF, 1), s);
			cursor.zero();
			face_hex_fill_kernel<<<grid_for(ctx, nH, blk), blk, 0, s>>>(hy->H_fs.p, hy->H_foff.p, nH, hy->F_nhoff.p, cursor.p, hy->F_nhs.p);
			FPOHM_LAUNCH_CHECK(ctx);
			hy->F_boundary.alloc(hy->nF, s);
			face_hex_sort_kernel<<<grid_for(ctx, hy->nF, blk), blk, 0, s>>>(hy->F_nhoff.p, hy->nF, hy->F_nhs.p, hy->F_boundary.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		// edges
		{
			const int64_t n = hy->tot_fv;
			DevBuf<uint64_t> key(n, s), skey(n, s);
			DevBuf<uint32_t> slot(n, s), sslot(n, s);
			loop_edge_keys_kernel<<<grid_for(ctx, hy->nF, blk), blk, 0, s>>>(hy->F_off.p, hy->nF, hy->F_vs.p, key.p, slot.p);
			FPOHM_LAUNCH_CHECK(ctx);
			size_t tb = 0;
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, skey.p, slot.p, sslot.p, n, 0, 64, s));
			DevBuf<uint8_t> tmp((int64_t)tb, s);
			FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, skey.p, slot.p, sslot.p, n, 0, 64, s));
			ctx->launches += 1;
			DevBuf<int64_t> head(n + 1, s), eid(n + 1, s);
			edge_heads_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(skey.p, n, head.p);
			FPOHM_LAUNCH_CHECK(ctx);
			size_t tb2 = 0;
			FPOHM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, head.p, eid.p, n, s));
			DevBuf<uint8_t> tmp2((int64_t)tb2, s);
			FPOHM_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, tb2, head.p, eid.p, n, s));
			ctx->launches += 1;
			int64_t ne = 0;
			if (n) FPOHM_CUDA(cudaMemcpyAsync(&ne, eid.p + (n - 1), 8, cudaMemcpyDeviceToHost, s));
			FPOHM_CUDA(cudaStreamSynchronize(s));
			hy->nE = ne;
			hy->E_vs.alloc(2 * ne, s); hy->E_boundary.alloc(std::max<int64_t>(ne, 1), s); hy->V_boundary.alloc(std::max<int64_t>(nV, 1), s);
			hy->E_boundary.zero(); hy->V_boundary.zero();
			edge_assign_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(skey.p, sslot.p, head.p, eid.p, n, hy->E_vs.p, hy->F_es.p);
			FPOHM_LAUNCH_CHECK(ctx);
			boundary_marks_kernel<<<grid_for(ctx, hy->nF, blk), blk, 0, s>>>(hy->F_off.p, hy->nF, hy->F_boundary.p, hy->F_es.p, hy->E_vs.p,
				hy->E_boundary.p, hy->V_boundary.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		timer.stop();
	} catch (...) { delete hy; throw; }
	*out = hy;
	(void)n_nodes;
	FPOHM_API_END
}

extern "C" {

int fpohm_conforming_mesh(fpohm_ctx *ctx, const fpohm_octree *oct, const fpohm_conn *conn, fpohm_hybrid **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && oct && conn && out, FPOHM_EINVAL, "fpohm_conforming_mesh: null argument");
	FPOHM_REQUIRE(conn->nV == oct->n_nodes && conn->H == oct->n_leaves, FPOHM_EINVAL,
	              "fpohm_conforming_mesh: the connectivity (%lld vertices, %lld hexes) is not that of this octree's hex mesh (%lld nodes, %lld leaves)",
	              (long long)conn->nV, (long long)conn->H, (long long)oct->n_nodes, (long long)oct->n_leaves);
	return conforming_impl(ctx, oct->node_pos.p, oct->node_neigh.p, oct->n_nodes, oct->prm.grid_size, conn, out);
	FPOHM_API_END
}

int fpohm_conforming_mesh_tables(fpohm_ctx *ctx, const int32_t *node_pos, const int32_t *node_neigh, int64_t n_nodes,
                                 const int32_t grid_size[3], const fpohm_conn *conn, fpohm_hybrid **out)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && node_pos && node_neigh && grid_size && conn && out, FPOHM_EINVAL, "fpohm_conforming_mesh_tables: null argument");
	FPOHM_REQUIRE(conn->nV == n_nodes, FPOHM_EINVAL, "fpohm_conforming_mesh_tables: %lld nodes but the connectivity has %lld vertices",
	              (long long)n_nodes, (long long)conn->nV);
	DeviceGuard g(ctx->device);
	DevBuf<int32_t> dpos(3 * n_nodes, ctx->stream), dnn(6 * n_nodes, ctx->stream);
	dpos.upload(node_pos, 3 * n_nodes); dnn.upload(node_neigh, 6 * n_nodes);
	return conforming_impl(ctx, dpos.p, dnn.p, n_nodes, grid_size, conn, out);     // (synchronises before returning)
	FPOHM_API_END
}

int fpohm_hybrid_sizes(const fpohm_hybrid *hy, int64_t sizes[8], int64_t *n_replaced_faces) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(hy && sizes, FPOHM_EINVAL, "fpohm_hybrid_sizes: null argument");
	sizes[0] = hy->nV; sizes[1] = hy->nF; sizes[2] = hy->nH; sizes[3] = hy->nE;
	sizes[4] = hy->tot_fv; sizes[5] = hy->tot_hf; sizes[6] = hy->tot_hv; sizes[7] = hy->tot_fn;
	if (n_replaced_faces) *n_replaced_faces = hy->n_replaced;
	FPOHM_API_END
}

int fpohm_hybrid_export(const fpohm_hybrid *hy, int64_t *F_off, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs,
                        uint8_t *E_boundary, uint8_t *V_boundary, int64_t *H_foff, uint32_t *H_fs, int64_t *H_voff, uint32_t *H_vs,
                        int64_t *F_nhoff, uint32_t *F_nhs)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(hy, FPOHM_EINVAL, "fpohm_hybrid_export: null argument");
	DeviceGuard g(hy->ctx->device);
	if (F_off) hy->F_off.download(F_off, hy->nF + 1);
	if (F_vs) hy->F_vs.download(F_vs, hy->tot_fv);
	if (F_es) hy->F_es.download(F_es, hy->tot_fv);
	if (F_boundary) hy->F_boundary.download(F_boundary, hy->nF);
	if (E_vs) hy->E_vs.download(E_vs, 2 * hy->nE);
	if (E_boundary) hy->E_boundary.download(E_boundary, hy->nE);
	if (V_boundary) hy->V_boundary.download(V_boundary, hy->nV);
	if (H_foff) hy->H_foff.download(H_foff, hy->nH + 1);
	if (H_fs) hy->H_fs.download(H_fs, hy->tot_hf);
	if (H_voff) hy->H_voff.download(H_voff, hy->nH + 1);
	if (H_vs) hy->H_vs.download(H_vs, hy->tot_hv);
	if (F_nhoff) hy->F_nhoff.download(F_nhoff, hy->nF + 1);
	if (F_nhs) hy->F_nhs.download(F_nhs, hy->tot_fn);
	FPOHM_CUDA(cudaStreamSynchronize(hy->ctx->stream));
	FPOHM_API_END
}

void fpohm_hybrid_free(fpohm_hybrid *hy) {
	if (!hy) return;
	DeviceGuard g(hy->ctx->device);
	cudaStreamSynchronize(hy->ctx->stream);
	delete hy;
}

} // extern "C"

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
	// dual meshes only (dual_conforming_mesh): vertex positions (cell centres), element types, census
	DevBuf<double> V;
	DevBuf<int32_t> h_type;
	int64_t census[7] = {0, 0, 0, 0, 0, 0, 0};
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

// build_connectivity, Hyb branch (gf.cpp:187-264) for a mesh given by its face loops (F_off/F_vs) and cell face lists
// (H_foff/H_fs): F.neighbor_hs + boundary faces, edges (ids in sorted (vmin, vmax) order) with per-slot edge ids,
// boundary edges / vertices.
static void hybrid_connectivity(fpohm_ctx *ctx, cudaStream_t s, fpohm_hybrid *hy) {
	const int blk = 256;
	// hybrid connectivity: face <-> hex
	{
		DevBuf<int64_t> cnt(hy->nF + 1, s);
		cnt.zero();
		face_hex_count_kernel<<<grid_for(ctx, hy->nH, blk), blk, 0, s>>>(hy->H_fs.p, hy->H_foff.p, hy->nH, cnt.p);
		FPOHM_LAUNCH_CHECK(ctx);
		hy->F_nhoff.alloc(hy->nF + 1, s);
		exclusive_scan(ctx, s, cnt.p, hy->F_nhoff.p, hy->nF + 1);
		hy->tot_fn = hy->tot_hf;
		hy->F_nhs.alloc(hy->tot_fn, s);
		DevBuf<int32_t> cursor(std::max<int64_t>(hy->nF, 1), s);
		cursor.zero();
		face_hex_fill_kernel<<<grid_for(ctx, hy->nH, blk), blk, 0, s>>>(hy->H_fs.p, hy->H_foff.p, hy->nH, hy->F_nhoff.p, cursor.p, hy->F_nhs.p);
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
		hy->E_vs.alloc(2 * ne, s); hy->E_boundary.alloc(std::max<int64_t>(ne, 1), s); hy->V_boundary.alloc(std::max<int64_t>(hy->nV, 1), s);
		hy->E_boundary.zero(); hy->V_boundary.zero();
		edge_assign_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(skey.p, sslot.p, head.p, eid.p, n, hy->E_vs.p, hy->F_es.p);
		FPOHM_LAUNCH_CHECK(ctx);
		boundary_marks_kernel<<<grid_for(ctx, hy->nF, blk), blk, 0, s>>>(hy->F_off.p, hy->nF, hy->F_boundary.p, hy->F_es.p, hy->E_vs.p,
			hy->E_boundary.p, hy->V_boundary.p);
		FPOHM_LAUNCH_CHECK(ctx);
	}
}

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
		hybrid_connectivity(ctx, s, hy);
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

// =====================================================================================================================
// dual_conforming_mesh (ghm.cpp:697-872): vertices = cell centres, one face per interior edge (the ring of cells around
// it, walked through shared faces), one cell per interior vertex (the faces of its edges), build_connectivity, then the
// element-type census that also rewrites each cell's vertex list in the order the templates of connectivity_modification
// expect (slab / pyramid / prism / pyramid-combine / tet-combine / hexahedron).
namespace {

__global__ void centres_kernel(const double *__restrict__ Vpos, const uint32_t *__restrict__ hex, int64_t nH, double *__restrict__ out) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 3 * nH; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t h = t / 3; const int c = (int)(t % 3);
		double acc = 0.0;
		for (int k = 0; k < 8; ++k) acc += Vpos[3 * (int64_t)hex[8 * h + k] + c];        // ghm.cpp:706
		out[t] = acc / 8;
	}
}
// generic CSR of (key, value) items with ascending values inside a key: count -> scan -> cursor fill -> per-list sort
__global__ void csr_count_kernel(const uint32_t *__restrict__ key, int64_t n, int64_t *__restrict__ cnt) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		atomicAdd((unsigned long long *)&cnt[key[t]], 1ull);
}
__global__ void csr_fill_kernel(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, int64_t n, const int64_t *__restrict__ off,
                                int32_t *__restrict__ cursor, uint32_t *__restrict__ out)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		out[off[key[t]] + atomicAdd(&cursor[key[t]], 1)] = val[t];
}
__global__ void csr_sort_kernel(const int64_t *__restrict__ off, int64_t n_keys, uint32_t *__restrict__ v) {
	for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_keys; k += (int64_t)gridDim.x * blockDim.x) {
		const int64_t a = off[k], n = off[k + 1] - a;
		for (int64_t i = 1; i < n; ++i) { const uint32_t x = v[a + i]; int64_t j = i - 1; while (j >= 0 && v[a + j] > x) { v[a + j + 1] = v[a + j]; --j; } v[a + j + 1] = x; }
	}
}
__global__ void slot_face_kernel(const int64_t *__restrict__ F_off, int64_t nF, uint32_t *__restrict__ slot_face) {
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x)
		for (int64_t q = F_off[f]; q < F_off[f + 1]; ++q) slot_face[q] = (uint32_t)f;
}
__global__ void edge_ends_kernel(int64_t nE, uint32_t *__restrict__ val) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 2 * nE; t += (int64_t)gridDim.x * blockDim.x) val[t] = (uint32_t)(t >> 1);
}
#define DUAL_RING 16
// E.neighbor_hs: sorted unique cells over the edge's faces (gf.cpp:249-258); FILL = false counts
template <bool FILL>
__global__ void edge_cells_kernel(int64_t nE, const int64_t *__restrict__ ef_off, const uint32_t *__restrict__ ef_val, const int64_t *__restrict__ fh_off,
                                  const uint32_t *__restrict__ fh_val, int64_t *__restrict__ cnt, const int64_t *__restrict__ off, uint32_t *__restrict__ out,
                                  int32_t *__restrict__ overflow)
{
	for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e <= nE; e += (int64_t)gridDim.x * blockDim.x) {
		if (e == nE) { if (!FILL) cnt[e] = 0; continue; }
		uint32_t hs[DUAL_RING]; int n = 0;
		for (int64_t a = ef_off[e]; a < ef_off[e + 1]; ++a) {
			const uint32_t f = ef_val[a];
			for (int64_t b = fh_off[f]; b < fh_off[f + 1]; ++b) {
				const uint32_t h = fh_val[b];
				int p = 0;
				while (p < n && hs[p] < h) ++p;
				if (p < n && hs[p] == h) continue;
				if (n == DUAL_RING) { atomicExch(overflow, 1); continue; }
				for (int c = n; c > p; --c) hs[c] = hs[c - 1];
				hs[p] = h; ++n;
			}
		}
		if (!FILL) cnt[e] = n; else for (int k = 0; k < n; ++k) out[off[e] + k] = hs[k];
	}
}
// dual face of interior edge e: the ring of its cells, ghm.cpp:715-738
__global__ void __launch_bounds__(128)
dual_faces_kernel(int64_t nE, const uint8_t *__restrict__ E_boundary, const int64_t *__restrict__ eh_off, const uint32_t *__restrict__ eh_val,
                  const int64_t *__restrict__ H_foff, const uint32_t *__restrict__ H_fs, const int64_t *__restrict__ fh_off,
                  const uint32_t *__restrict__ fh_val, const int64_t *__restrict__ e_tag /*exclusive rank among interior edges*/,
                  const int64_t *__restrict__ loop_off /*per dual face*/, uint32_t *__restrict__ F_vs)
{
	for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nE; e += (int64_t)gridDim.x * blockDim.x) {
		if (E_boundary[e]) continue;
		const int64_t a = eh_off[e]; const int n = (int)(eh_off[e + 1] - a);
		const uint32_t *hs = eh_val + a;
		uint32_t *out = F_vs + loop_off[e_tag[e]];
		unsigned tagged = 0;
		uint32_t sh = hs[0];
		for (int j = 0; j < n; ++j) {
			out[j] = sh;
			for (int k = 0; k < n; ++k) if (hs[k] == sh) tagged |= 1u << k;
			for (int64_t q = H_foff[sh]; q < H_foff[sh + 1]; ++q) {
				const uint32_t f = H_fs[q];
				if (fh_off[f + 1] - fh_off[f] == 1) continue;
				uint32_t hid = fh_val[fh_off[f]];
				if (hid == sh) hid = fh_val[fh_off[f] + 1];
				int idx = -1;
				for (int k = 0; k < n; ++k) if (hs[k] == hid) { idx = k; break; }
				if (idx >= 0 && !((tagged >> idx) & 1u)) { sh = hid; break; }
			}
		}
	}
}
#define DUAL_MAX_HV 64
// dual cell of interior vertex v: faces = e_tag of its edges (ascending edge id), vertex set sorted unique; FILL = false counts
template <bool FILL>
__global__ void __launch_bounds__(128)
dual_cells_kernel(int64_t nV, const uint8_t *__restrict__ V_boundary, const int64_t *__restrict__ cell_rank, const int64_t *__restrict__ ve_off,
                  const uint32_t *__restrict__ ve_val, const int64_t *__restrict__ e_tag, const int64_t *__restrict__ F_off, const uint32_t *__restrict__ F_vs,
                  int64_t *__restrict__ nf, int64_t *__restrict__ nv, const int64_t *__restrict__ foff, const int64_t *__restrict__ voff,
                  uint32_t *__restrict__ H_fs, uint32_t *__restrict__ H_vs, int32_t *__restrict__ overflow)
{
	for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
		if (V_boundary[v]) continue;
		const int64_t c = cell_rank[v];
		uint32_t vs[DUAL_MAX_HV]; int n = 0;
		const int64_t a = ve_off[v], m = ve_off[v + 1] - a;
		for (int64_t k = 0; k < m; ++k) {
			const int64_t f = e_tag[ve_val[a + k]];
			if (FILL) H_fs[foff[c] + k] = (uint32_t)f;
			for (int64_t q = F_off[f]; q < F_off[f + 1]; ++q) {
				const uint32_t x = F_vs[q];
				int p = 0;
				while (p < n && vs[p] < x) ++p;
				if (p < n && vs[p] == x) continue;
				if (n == DUAL_MAX_HV) { atomicExch(overflow, 1); continue; }
				for (int t = n; t > p; --t) vs[t] = vs[t - 1];
				vs[p] = x; ++n;
			}
		}
		if (!FILL) { nf[c] = m; nv[c] = n; } else for (int k = 0; k < n; ++k) H_vs[voff[c] + k] = vs[k];
	}
}

__device__ __forceinline__ bool edge_exists(const uint32_t *__restrict__ E_vs, int64_t nE, uint32_t a, uint32_t b) {
	if (a == b) return true;                       // (identical neighbour-edge lists always intersect)
	if (a > b) { const uint32_t t = a; a = b; b = t; }
	const uint64_t key = ((uint64_t)a << 32) | b;
	int64_t lo = 0, hi = nE;
	while (lo < hi) { const int64_t mid = (lo + hi) >> 1; const uint64_t k = ((uint64_t)E_vs[2 * mid] << 32) | E_vs[2 * mid + 1]; if (k < key) lo = mid + 1; else hi = mid; }
	return lo < nE && E_vs[2 * lo] == a && E_vs[2 * lo + 1] == b;
}
__device__ __forceinline__ bool has(const uint32_t *v, int n, uint32_t x) { for (int i = 0; i < n; ++i) if (v[i] == x) return true; return false; }
// sorted intersection of two triangles (set_intersection of the sorted vertex lists): returns count, first two in s
__device__ __forceinline__ int tri_shared(const uint32_t *a, const uint32_t *b, uint32_t s[2]) {
	uint32_t c[3]; int n = 0;
	for (int i = 0; i < 3; ++i) if (has(b, 3, a[i])) c[n++] = a[i];
	for (int i = 1; i < n; ++i) { const uint32_t x = c[i]; int j = i - 1; while (j >= 0 && c[j] > x) { c[j + 1] = c[j]; --j; } c[j + 1] = x; }
	if (n > 0) s[0] = c[0];
	if (n > 1) s[1] = c[1];
	return n;
}
// ghm.cpp:760-868: element type and ordered vertex list of one dual cell; returns the type, *n_out the list length
__device__ int classify_cell(const int64_t *__restrict__ F_off, const uint32_t *__restrict__ F_vs, const uint32_t *__restrict__ fs, int nfs,
                             const uint32_t *__restrict__ hvs, int nhvs, const uint32_t *__restrict__ E_vs, int64_t nE, uint32_t *out, int *n_out)
{
	int triN = 0, quadN = 0; uint32_t tris[8], quads[8];
	for (int k = 0; k < nfs; ++k) {
		const int sz = (int)(F_off[fs[k] + 1] - F_off[fs[k]]);
		if (sz == 3) { if (triN < 8) tris[triN] = fs[k]; ++triN; } else if (sz == 4) { if (quadN < 8) quads[quadN] = fs[k]; ++quadN; }
	}
	int n = 0, type = 0;
	auto fv = [&](uint32_t f) { return F_vs + F_off[f]; };
	if (nfs == 4) {
		if (triN == 4) type = 0;
		else if (triN == 2 && quadN == 2) {
			type = 1;
			uint32_t sh[2] = {0, 0};
			tri_shared(fv(tris[0]), fv(tris[1]), sh);
			const uint32_t *q = has(fv(quads[0]), 4, sh[0]) ? fv(quads[0]) : fv(quads[1]);
			int id = 0;
			for (int j = 0; j < 4; ++j) if (q[j] == sh[0]) { id = j; break; }
			for (int j = 0; j < 4; ++j) out[n++] = q[(id + j) & 3];
			out[n++] = sh[1];
		}
	} else if (nfs == 5) {
		if (triN == 4 && quadN == 1) {
			type = 2;
			const uint32_t *q = fv(quads[0]), *t = fv(tris[0]);
			for (int j = 0; j < 4; ++j) out[n++] = q[j];
			for (int j = 0; j < 3; ++j) if (!has(out, 4, t[j])) { out[n++] = t[j]; break; }
		} else if (triN == 2 && quadN == 3) {
			type = 3;
			const uint32_t *t0 = fv(tris[0]), *t1 = fv(tris[1]);
			for (int j = 0; j < 3; ++j) out[n++] = t0[j];
			for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) if (edge_exists(E_vs, nE, t0[j], t1[k])) { out[n++] = t1[k]; break; }
		}
	} else if (nfs == 6) {
		if (triN == 2 && quadN == 4) {
			type = 4;
			uint32_t sh[2] = {0, 0};
			tri_shared(fv(tris[0]), fv(tris[1]), sh);
			for (int a = 0; a < 4; ++a) {
				const uint32_t *q = fv(quads[a]);
				if (!has(q, 4, sh[0])) continue;
				int id = 0;
				for (int j = 0; j < 4; ++j) if (q[j] == sh[0]) { id = j; break; }
				for (int j = 0; j < 4; ++j) out[n++] = q[(id + j) & 3];
				uint32_t ordered[3]; int no = 0;
				for (int j = 1; j < 4; ++j)
					for (int k = 0; k < nhvs; ++k) {
						if (has(out, 4, hvs[k])) continue;
						if (edge_exists(E_vs, nE, out[j], hvs[k])) { ordered[no++] = hvs[k]; break; }
					}
				for (int j = 0; j < no; ++j) out[n++] = ordered[j];
				break;
			}
		} else if (quadN == 6) {
			type = 6;
			const uint32_t *q = fv(quads[0]);
			for (int j = 0; j < 4; ++j) out[n++] = q[j];
			for (int j = 0; j < 4; ++j)
				for (int k = 0; k < nhvs; ++k) {
					if (has(q, 4, hvs[k])) continue;
					if (edge_exists(E_vs, nE, q[j], hvs[k])) { out[n++] = hvs[k]; break; }
				}
		} else if (triN == 4 && quadN == 2) {
			type = 5;
			for (int k = 0; k < nhvs && k < 16; ++k) out[n++] = hvs[k];
		}
	}
	*n_out = n;
	return type;
}
template <bool FILL>
__global__ void __launch_bounds__(128)
classify_kernel(int64_t nH, const int64_t *__restrict__ F_off, const uint32_t *__restrict__ F_vs, const int64_t *__restrict__ H_foff,
                const uint32_t *__restrict__ H_fs, const int64_t *__restrict__ H_voff, const uint32_t *__restrict__ H_vs, const uint32_t *__restrict__ E_vs,
                int64_t nE, int64_t *__restrict__ cnt, const int64_t *__restrict__ new_off, uint32_t *__restrict__ new_vs, int32_t *__restrict__ h_type,
                unsigned long long *__restrict__ census)
{
	for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h <= nH; h += (int64_t)gridDim.x * blockDim.x) {
		if (h == nH) { if (!FILL) cnt[h] = 0; continue; }
		uint32_t out[16]; int n = 0;
		const int nfs = (int)(H_foff[h + 1] - H_foff[h]);
		const int type = classify_cell(F_off, F_vs, H_fs + H_foff[h], nfs, H_vs + H_voff[h], (int)(H_voff[h + 1] - H_voff[h]), E_vs, nE, out, &n);
		if (!FILL) { cnt[h] = n; continue; }
		h_type[h] = type;
		for (int k = 0; k < n; ++k) new_vs[new_off[h] + k] = out[k];
		// the census counts only the recognised shapes (a cell that matches none keeps the default type 0)
		int triN = 0;
		for (int k = 0; k < nfs; ++k) triN += (F_off[H_fs[H_foff[h] + k] + 1] - F_off[H_fs[H_foff[h] + k]]) == 3;
		if (type != 0 || (nfs == 4 && triN == 4)) atomicAdd(&census[type], 1ull);
	}
}

void build_csr(fpohm_ctx *ctx, cudaStream_t s, const uint32_t *key, const uint32_t *val, int64_t n_items, int64_t n_keys,
               DevBuf<int64_t> &off, DevBuf<uint32_t> &out)
{
	const int blk = 256;
	DevBuf<int64_t> cnt(n_keys + 1, s);
	cnt.zero();
	csr_count_kernel<<<grid_for(ctx, n_items, blk), blk, 0, s>>>(key, n_items, cnt.p);
	FPOHM_LAUNCH_CHECK(ctx);
	off.alloc(n_keys + 1, s);
	exclusive_scan(ctx, s, cnt.p, off.p, n_keys + 1);
	out.alloc(std::max<int64_t>(n_items, 1), s);
	DevBuf<int32_t> cursor(std::max<int64_t>(n_keys, 1), s);
	cursor.zero();
	csr_fill_kernel<<<grid_for(ctx, n_items, blk), blk, 0, s>>>(key, val, n_items, off.p, cursor.p, out.p);
	FPOHM_LAUNCH_CHECK(ctx);
	csr_sort_kernel<<<grid_for(ctx, n_keys, blk), blk, 0, s>>>(off.p, n_keys, out.p);
	FPOHM_LAUNCH_CHECK(ctx);
}

__global__ void not_flag_kernel(const uint8_t *__restrict__ b, int64_t n, int64_t *__restrict__ out) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= n; t += (int64_t)gridDim.x * blockDim.x) out[t] = t < n ? (b[t] ? 0 : 1) : 0;
}
__global__ void ring_size_kernel(const uint8_t *__restrict__ E_boundary, const int64_t *__restrict__ eh_off, const int64_t *__restrict__ e_tag, int64_t nE,
                                 int64_t *__restrict__ size /*per dual face*/)
{
	for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nE; e += (int64_t)gridDim.x * blockDim.x)
		if (!E_boundary[e]) size[e_tag[e]] = eh_off[e + 1] - eh_off[e];
}

} // namespace

extern "C" {

int fpohm_dual_conforming_mesh(fpohm_ctx *ctx, const fpohm_hybrid *hy, const double *Vpos, int64_t nV_mo, const uint32_t *hex, int64_t n_hex,
                               fpohm_hybrid **out)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && hy && Vpos && hex && out, FPOHM_EINVAL, "fpohm_dual_conforming_mesh: null argument");
	FPOHM_REQUIRE(n_hex == hy->nH && nV_mo == hy->nV, FPOHM_EINVAL, "fpohm_dual_conforming_mesh: %lld hexes / %lld vertices but the polyhedral mesh has %lld cells / %lld vertices",
	              (long long)n_hex, (long long)nV_mo, (long long)hy->nH, (long long)hy->nV);
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	fpohm_hybrid *d = new fpohm_hybrid;
	try {
		d->ctx = ctx;
		DevBuf<double> dV(3 * nV_mo, s);
		DevBuf<uint32_t> dhex(8 * n_hex, s);
		dV.upload(Vpos, 3 * nV_mo); dhex.upload(hex, 8 * n_hex);
		KernelTimer timer(ctx, s);
		d->nV = hy->nH;
		d->V.alloc(3 * d->nV, s);
		centres_kernel<<<grid_for(ctx, 3 * d->nV, blk), blk, 0, s>>>(dV.p, dhex.p, d->nV, d->V.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// adjacency of the polyhedral mesh that build_connectivity would have left: E.neighbor_fs, E.neighbor_hs, V.neighbor_es
		DevBuf<int64_t> ef_off, eh_off, ve_off;
		DevBuf<uint32_t> ef_val, eh_val, ve_val;
		{
			DevBuf<uint32_t> slot_face(std::max<int64_t>(hy->tot_fv, 1), s);
			slot_face_kernel<<<grid_for(ctx, hy->nF, blk), blk, 0, s>>>(hy->F_off.p, hy->nF, slot_face.p);
			FPOHM_LAUNCH_CHECK(ctx);
			build_csr(ctx, s, hy->F_es.p, slot_face.p, hy->tot_fv, hy->nE, ef_off, ef_val);
			DevBuf<uint32_t> ends(std::max<int64_t>(2 * hy->nE, 1), s);
			edge_ends_kernel<<<grid_for(ctx, 2 * hy->nE, blk), blk, 0, s>>>(hy->nE, ends.p);
			FPOHM_LAUNCH_CHECK(ctx);
			build_csr(ctx, s, hy->E_vs.p, ends.p, 2 * hy->nE, hy->nV, ve_off, ve_val);
		}
		DevBuf<int32_t> ovf(1, s);
		ovf.zero();
		{
			DevBuf<int64_t> cnt(hy->nE + 1, s);
			edge_cells_kernel<false><<<grid_for(ctx, hy->nE + 1, 128), 128, 0, s>>>(hy->nE, ef_off.p, ef_val.p, hy->F_nhoff.p, hy->F_nhs.p, cnt.p, nullptr, nullptr, ovf.p);
			FPOHM_LAUNCH_CHECK(ctx);
			eh_off.alloc(hy->nE + 1, s);
			exclusive_scan(ctx, s, cnt.p, eh_off.p, hy->nE + 1);
			const int64_t tot = last_of(eh_off.p, hy->nE, s);
			eh_val.alloc(std::max<int64_t>(tot, 1), s);
			edge_cells_kernel<true><<<grid_for(ctx, hy->nE + 1, 128), 128, 0, s>>>(hy->nE, ef_off.p, ef_val.p, hy->F_nhoff.p, hy->F_nhs.p, nullptr, eh_off.p, eh_val.p, ovf.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		// dual faces: one per interior edge
		DevBuf<int64_t> e_keep(hy->nE + 1, s), e_tag(hy->nE + 1, s);
		not_flag_kernel<<<grid_for(ctx, hy->nE + 1, blk), blk, 0, s>>>(hy->E_boundary.p, hy->nE, e_keep.p);
		FPOHM_LAUNCH_CHECK(ctx);
		exclusive_scan(ctx, s, e_keep.p, e_tag.p, hy->nE + 1);
		d->nF = last_of(e_tag.p, hy->nE, s);
		{
			DevBuf<int64_t> size(d->nF + 1, s);
			size.zero();
			ring_size_kernel<<<grid_for(ctx, hy->nE, blk), blk, 0, s>>>(hy->E_boundary.p, eh_off.p, e_tag.p, hy->nE, size.p);
			FPOHM_LAUNCH_CHECK(ctx);
			d->F_off.alloc(d->nF + 1, s);
			exclusive_scan(ctx, s, size.p, d->F_off.p, d->nF + 1);
			d->tot_fv = last_of(d->F_off.p, d->nF, s);
			d->F_vs.alloc(std::max<int64_t>(d->tot_fv, 1), s); d->F_es.alloc(std::max<int64_t>(d->tot_fv, 1), s);
			dual_faces_kernel<<<grid_for(ctx, hy->nE, 128), 128, 0, s>>>(hy->nE, hy->E_boundary.p, eh_off.p, eh_val.p, hy->H_foff.p, hy->H_fs.p,
				hy->F_nhoff.p, hy->F_nhs.p, e_tag.p, d->F_off.p, d->F_vs.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		// dual cells: one per interior vertex
		DevBuf<int64_t> v_keep(hy->nV + 1, s), v_rank(hy->nV + 1, s);
		not_flag_kernel<<<grid_for(ctx, hy->nV + 1, blk), blk, 0, s>>>(hy->V_boundary.p, hy->nV, v_keep.p);
		FPOHM_LAUNCH_CHECK(ctx);
		exclusive_scan(ctx, s, v_keep.p, v_rank.p, hy->nV + 1);
		d->nH = last_of(v_rank.p, hy->nV, s);
		DevBuf<int64_t> H_voff0;
		DevBuf<uint32_t> H_vs0;
		{
			DevBuf<int64_t> nf(d->nH + 1, s), nv(d->nH + 1, s);
			nf.zero(); nv.zero();
			dual_cells_kernel<false><<<grid_for(ctx, hy->nV, 128), 128, 0, s>>>(hy->nV, hy->V_boundary.p, v_rank.p, ve_off.p, ve_val.p, e_tag.p, d->F_off.p,
				d->F_vs.p, nf.p, nv.p, nullptr, nullptr, nullptr, nullptr, ovf.p);
			FPOHM_LAUNCH_CHECK(ctx);
			d->H_foff.alloc(d->nH + 1, s); H_voff0.alloc(d->nH + 1, s);
			exclusive_scan(ctx, s, nf.p, d->H_foff.p, d->nH + 1);
			exclusive_scan(ctx, s, nv.p, H_voff0.p, d->nH + 1);
			d->tot_hf = last_of(d->H_foff.p, d->nH, s);
			const int64_t tot_hv0 = last_of(H_voff0.p, d->nH, s);
			d->H_fs.alloc(std::max<int64_t>(d->tot_hf, 1), s); H_vs0.alloc(std::max<int64_t>(tot_hv0, 1), s);
			dual_cells_kernel<true><<<grid_for(ctx, hy->nV, 128), 128, 0, s>>>(hy->nV, hy->V_boundary.p, v_rank.p, ve_off.p, ve_val.p, e_tag.p, d->F_off.p,
				d->F_vs.p, nullptr, nullptr, d->H_foff.p, H_voff0.p, d->H_fs.p, H_vs0.p, ovf.p);
			FPOHM_LAUNCH_CHECK(ctx);
		}
		int32_t ov = 0;
		ovf.download(&ov, 1);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		FPOHM_REQUIRE(ov == 0, FPOHM_ERANGE, "fpohm_dual_conforming_mesh: more than %d cells around an edge or %d vertices in a dual cell", DUAL_RING, DUAL_MAX_HV);
		hybrid_connectivity(ctx, s, d);
		// element types + ordered vertex lists
		{
			DevBuf<int64_t> cnt(d->nH + 1, s);
			DevBuf<unsigned long long> census(7, s);
			FPOHM_CUDA(cudaMemsetAsync(census.p, 0, 56, s));
			classify_kernel<false><<<grid_for(ctx, d->nH + 1, 128), 128, 0, s>>>(d->nH, d->F_off.p, d->F_vs.p, d->H_foff.p, d->H_fs.p, H_voff0.p, H_vs0.p,
				d->E_vs.p, d->nE, cnt.p, nullptr, nullptr, nullptr, nullptr);
			FPOHM_LAUNCH_CHECK(ctx);
			d->H_voff.alloc(d->nH + 1, s);
			exclusive_scan(ctx, s, cnt.p, d->H_voff.p, d->nH + 1);
			d->tot_hv = last_of(d->H_voff.p, d->nH, s);
			d->H_vs.alloc(std::max<int64_t>(d->tot_hv, 1), s);
			d->h_type.alloc(std::max<int64_t>(d->nH, 1), s);
			classify_kernel<true><<<grid_for(ctx, d->nH + 1, 128), 128, 0, s>>>(d->nH, d->F_off.p, d->F_vs.p, d->H_foff.p, d->H_fs.p, H_voff0.p, H_vs0.p,
				d->E_vs.p, d->nE, nullptr, d->H_voff.p, d->H_vs.p, d->h_type.p, census.p);
			FPOHM_LAUNCH_CHECK(ctx);
			unsigned long long hc[7];
			census.download(hc, 7);
			FPOHM_CUDA(cudaStreamSynchronize(s));
			for (int k = 0; k < 7; ++k) d->census[k] = (int64_t)hc[k];
		}
		timer.stop();
	} catch (...) { delete d; throw; }
	*out = d;
	FPOHM_API_END
}

int fpohm_hybrid_dual_extra(const fpohm_hybrid *d, double *V, int32_t *h_type, int64_t census[7]) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(d && d->V.p, FPOHM_ESTATE, "fpohm_hybrid_dual_extra: not a dual mesh");
	DeviceGuard g(d->ctx->device);
	if (V) d->V.download(V, 3 * d->nV);
	if (h_type) d->h_type.download(h_type, d->nH);
	FPOHM_CUDA(cudaStreamSynchronize(d->ctx->stream));
	if (census) for (int k = 0; k < 7; ++k) census[k] = d->census[k];
	FPOHM_API_END
}

} // extern "C"

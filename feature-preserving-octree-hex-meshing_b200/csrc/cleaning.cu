// clean_hex_mesh and its stages: reorder_hex_mesh (gf.cpp:2199-2229), tagging_uneven_element
// (ghm.cpp:1983-2005), re_indexing_connectivity (gf.cpp:664-698), clean_non_manifold_ve (ghm.cpp:2006-2080),
// drop_small_pieces (ghm.cpp:2081-2124) and the medial-surface flags at the end of clean_hex_mesh (ghm.cpp:1970-1981).
//
// The reference runs every stage as ONE sequential loop whose later iterations see what earlier ones changed.  Each of
// them is restated here as a data-parallel computation with the same result:
//   * tagging: element i reads the NEW flags of lower-numbered neighbours and the OLD flags of higher-numbered ones.
//     Sweeps "next[i] = f(i; cur[j<i], old[j>i])" over all elements at once are repeated until nothing moves; by
//     induction on i the fix-point is the sequential result (element 0 depends on nothing new, element i only on
//     elements < i), and every sweep fixes at least one more element, in practice all of them in 2-3 sweeps.
//   * non-manifold vertices / edges: whether a vertex (edge) is non-manifold and which hexes it would drop depends only
//     on the sub-mesh of the round, not on the loop — only "was it tagged by an earlier one" does.  So all candidates are
//     found at once, and the greedy selection "candidate u counts iff no counted candidate w < u dropped a hex touching
//     u" is resolved in rounds (decided candidates never change; the lowest undecided one is decidable in every round).
//   * drop_small_pieces: face-connected pieces by lock-free union-find (the smaller root wins, so a piece is named by its
//     lowest hex, which is also the reference's order of discovery).  The reference compares the sizes of its work lists,
//     which keep duplicates: one push per interior face of the piece, hence size = 1 + #interior faces — that count is
//     what is compared here, ties to the piece found first.
// Integer work throughout (the reorder determinant repeats Eigen's expression tree): bit-exact.
#include "conn.h"
#include "mesh.h"

#include <cub/device/device_scan.cuh>
#include <cstring>

using namespace fpohm;

namespace {

#define CLEAN_MAX_BFS 64          /* boundary faces around one vertex kept in a thread's local list */

__constant__ int c_hex_tet[8][4] = {{0, 3, 4, 1}, {1, 0, 5, 2}, {2, 1, 6, 3}, {3, 2, 7, 0}, {4, 7, 5, 0}, {5, 4, 6, 1}, {6, 5, 7, 2}, {7, 6, 4, 3}}; // global_types.h:163-173

// reorder_hex_mesh: vol = sum_j a_jacobian_nonscaled(corner j) accumulated in j order; < 0 -> mirrored vertex list
__global__ void reorder_hexes_kernel(const double *__restrict__ V, uint32_t *__restrict__ hex, int64_t H, unsigned long long *__restrict__ n_mirrored) {
	for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < H; h += (int64_t)gridDim.x * blockDim.x) {
		uint32_t vs[8];
		double p[8][3];
		for (int k = 0; k < 8; ++k) { vs[k] = hex[8 * h + k]; for (int d = 0; d < 3; ++d) p[k][d] = V[3 * (int64_t)vs[k] + d]; }
		double vol = 0;
		for (int j = 0; j < 8; ++j) {
			const double *c0 = p[c_hex_tet[j][0]], *c1 = p[c_hex_tet[j][1]], *c2 = p[c_hex_tet[j][2]], *c3 = p[c_hex_tet[j][3]];
			// Jacobian.col(k) = v_{k+1} - v0 ; Eigen 3.2 determinant (bruteforce_det3_helper), as in jacobian.cu
			const double m00 = c1[0] - c0[0], m10 = c1[1] - c0[1], m20 = c1[2] - c0[2];
			const double m01 = c2[0] - c0[0], m11 = c2[1] - c0[1], m21 = c2[2] - c0[2];
			const double m02 = c3[0] - c0[0], m12 = c3[1] - c0[1], m22 = c3[2] - c0[2];
			const double det = m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
			vol += det;
		}
		if (vol < 0) {
			const int perm[8] = {3, 2, 1, 0, 7, 6, 5, 4};
			for (int k = 0; k < 8; ++k) hex[8 * h + k] = vs[perm[k]];
			atomicAdd(n_mirrored, 1ull);
		}
	}
}

// the hex on the other side of each of the 6 faces (-1 on the boundary)
__global__ void hex_neighbours_kernel(const uint32_t *__restrict__ H_fs, const int64_t *__restrict__ nh_off, const uint32_t *__restrict__ nh_val,
                                      int64_t H, int32_t *__restrict__ nb)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 6 * H; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t h = t / 6;
		const uint32_t f = H_fs[t];
		const int64_t o = nh_off[f];
		int32_t r = -1;
		if (nh_off[f + 1] - o == 2) { const uint32_t a = nh_val[o], b = nh_val[o + 1]; r = (int32_t)(a == (uint32_t)h ? b : a); }
		nb[t] = r;
	}
}

// one sweep of tagging_uneven_element: neighbours below i are read from `cur`, the others (and i itself) from `old`
__global__ void tag_sweep_kernel(const int32_t *__restrict__ nb, int64_t H, const uint8_t *__restrict__ old, const uint8_t *__restrict__ cur,
                                 uint8_t *__restrict__ next, int32_t *__restrict__ moved)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < H; i += (int64_t)gridDim.x * blockDim.x) {
		const uint8_t mine = old[i];
		int bn = 0, in = 0;
		for (int j = 0; j < 6; ++j) {
			const int32_t o = nb[6 * i + j];
			if (o < 0) { ++bn; continue; }
			const uint8_t of = o < i ? cur[o] : old[o];
			if (of != mine) ++bn; else ++in;
		}
		const uint8_t r = (bn == 5 && in == 1) ? (uint8_t)!mine : mine;
		next[i] = r;
		if (r != cur[i]) *moved = 1;
	}
}

// ---- re_indexing_connectivity
__global__ void mark_vertices_kernel(const uint32_t *__restrict__ hex, const uint8_t *__restrict__ flag, int64_t H, int32_t *__restrict__ vtag) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 8 * H; t += (int64_t)gridDim.x * blockDim.x)
		if (flag[t >> 3]) vtag[hex[t]] = 1;
}
__global__ void flags_to_i32_kernel(const uint8_t *__restrict__ flag, int64_t n, int32_t *__restrict__ out) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) out[t] = flag[t] ? 1 : 0;
}
__global__ void vertex_maps_kernel(const int32_t *__restrict__ vtag, const int32_t *__restrict__ pos, int64_t nV, int32_t *__restrict__ V_map,
                                   int32_t *__restrict__ V_rev)
{
	for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
		if (vtag[v]) { V_map[v] = pos[v]; V_rev[pos[v]] = (int32_t)v; } else V_map[v] = -1;
	}
}
__global__ void sub_hexes_kernel(const uint32_t *__restrict__ hex, const int32_t *__restrict__ hkeep, const int32_t *__restrict__ hpos, int64_t H,
                                 const int32_t *__restrict__ V_map, int32_t *__restrict__ H_rev, uint32_t *__restrict__ sub_hex)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 8 * H; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t h = t >> 3;
		if (!hkeep[h]) continue;
		const int64_t n = hpos[h];
		sub_hex[8 * n + (t & 7)] = (uint32_t)V_map[hex[t]];
		if ((t & 7) == 0) H_rev[n] = (int32_t)h;
	}
}

struct SubMesh {
	DevBuf<int32_t> V_map, V_rev, H_rev;
	DevBuf<uint32_t> hex;
	int64_t nv = 0, nh = 0;
};

void excl_scan_i32(fpohm_ctx *ctx, const int32_t *in, int32_t *out, int64_t n, cudaStream_t s) {
	size_t tb = 0;
	FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s));
	DevBuf<uint8_t> tmp((int64_t)tb, s);
	FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n, s));
	ctx->launches += 1;
}

void reindex_dev(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, const uint8_t *flag, SubMesh &out, cudaStream_t s) {
	const int blk = 256;
	DevBuf<int32_t> vtag(nV + 1, s), vpos(nV + 1, s), hkeep(H + 1, s), hpos(H + 1, s);
	vtag.zero(); hkeep.zero();
	mark_vertices_kernel<<<grid_for(ctx, 8 * H, blk), blk, 0, s>>>(hex, flag, H, vtag.p);
	FPOHM_LAUNCH_CHECK(ctx);
	flags_to_i32_kernel<<<grid_for(ctx, H, blk), blk, 0, s>>>(flag, H, hkeep.p);
	FPOHM_LAUNCH_CHECK(ctx);
	excl_scan_i32(ctx, vtag.p, vpos.p, nV + 1, s);
	excl_scan_i32(ctx, hkeep.p, hpos.p, H + 1, s);
	int32_t tot[2] = {0, 0};
	FPOHM_CUDA(cudaMemcpyAsync(&tot[0], vpos.p + nV, 4, cudaMemcpyDeviceToHost, s));
	FPOHM_CUDA(cudaMemcpyAsync(&tot[1], hpos.p + H, 4, cudaMemcpyDeviceToHost, s));
	FPOHM_CUDA(cudaStreamSynchronize(s));
	out.nv = tot[0]; out.nh = tot[1];
	out.V_map.alloc(nV, s); out.V_rev.alloc(out.nv, s); out.H_rev.alloc(out.nh, s); out.hex.alloc(8 * out.nh, s);
	vertex_maps_kernel<<<grid_for(ctx, nV, blk), blk, 0, s>>>(vtag.p, vpos.p, nV, out.V_map.p, out.V_rev.p);
	FPOHM_LAUNCH_CHECK(ctx);
	sub_hexes_kernel<<<grid_for(ctx, 8 * H, blk), blk, 0, s>>>(hex, hkeep.p, hpos.p, H, out.V_map.p, out.H_rev.p, out.hex.p);
	FPOHM_LAUNCH_CHECK(ctx);
}

// the sub-mesh of a round with the tables the cleaning stages read
struct SubConn {
	SubMesh sub;
	fpohm_conn *c = nullptr;
	~SubConn() { delete c; }
};
void build_sub(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, const uint8_t *flag, SubConn &sc, cudaStream_t s) {
	delete sc.c; sc.c = nullptr;
	reindex_dev(ctx, hex, H, nV, flag, sc.sub, s);
	if (sc.sub.nh == 0) return;
	DevBuf<uint32_t> copy(8 * sc.sub.nh, s);
	FPOHM_CUDA(cudaMemcpyAsync(copy.p, sc.sub.hex.p, 32 * (size_t)sc.sub.nh, cudaMemcpyDeviceToDevice, s));
	sc.c = conn_build_dev(ctx, std::move(copy), sc.sub.nh, sc.sub.nv, false);
}

// ---- clean_non_manifold_ve, vertex pass.  One thread per boundary vertex: flood the boundary faces around it from the
// first one through shared edges (ghm.cpp:2020-2046); the faces not reached name the hexes to drop (:2047-2056).
// rem[8h + k] = 1: corner k of hex h is a candidate vertex that drops h.
__global__ void nm_vertex_candidates_kernel(int64_t nv, const uint8_t *__restrict__ V_boundary, const int64_t *__restrict__ vf_off,
                                            const uint32_t *__restrict__ vf_val, const uint8_t *__restrict__ F_boundary,
                                            const uint32_t *__restrict__ F_es, const int64_t *__restrict__ ef_off, const uint32_t *__restrict__ ef_val,
                                            const int64_t *__restrict__ fh_off, const uint32_t *__restrict__ fh_val, const uint32_t *__restrict__ hex,
                                            uint8_t *__restrict__ rem, int32_t *__restrict__ cand_list, int32_t *__restrict__ counters /*0 n_cand 1 overflow*/)
{
	for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
		if (!V_boundary[v]) continue;
		uint32_t bfs[CLEAN_MAX_BFS];
		int n = 0;
		bool over = false;
		for (int64_t t = vf_off[v]; t < vf_off[v + 1]; ++t) {
			const uint32_t f = vf_val[t];
			if (!F_boundary[f]) continue;
			if (n < CLEAN_MAX_BFS) bfs[n++] = f; else over = true;
		}
		if (over) { counters[1] = 1; continue; }
		if (n == 0) continue;
		unsigned long long reached = 1ull, expanded = 0ull;
		while (reached != expanded) {
			const int i = __ffsll((long long)(reached & ~expanded)) - 1;
			expanded |= 1ull << i;
			const uint32_t f = bfs[i];
			for (int k = 0; k < 4; ++k) {
				const uint32_t e = F_es[4 * (int64_t)f + k];
				for (int64_t t = ef_off[e]; t < ef_off[e + 1]; ++t) {
					const uint32_t nf = ef_val[t];
					for (int j = 0; j < n; ++j) if (bfs[j] == nf) { reached |= 1ull << j; break; }
				}
			}
		}
		const unsigned long long all = n == 64 ? ~0ull : ((1ull << n) - 1ull);
		if (reached == all) continue;
		cand_list[atomicAdd(&counters[0], 1)] = (int32_t)v;
		for (int i = 0; i < n; ++i) {
			if ((reached >> i) & 1ull) continue;
			const uint32_t h = fh_val[fh_off[bfs[i]]];                       // neighbor_hs[0] of a boundary face
			for (int k = 0; k < 8; ++k) if (hex[8 * (int64_t)h + k] == (uint32_t)v) rem[8 * (int64_t)h + k] = 1;
		}
	}
}

// greedy selection in index order, one round.  state: 0 undecided, 1 counted, 2 skipped (tagged by a counted one below it)
__global__ void nm_vertex_resolve_kernel(const int32_t *__restrict__ cand_list, int n_cand, const int64_t *__restrict__ vh_off,
                                         const uint32_t *__restrict__ vh_val, const uint32_t *__restrict__ hex, const uint8_t *__restrict__ rem,
                                         uint8_t *__restrict__ state, int32_t *__restrict__ undecided)
{
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_cand; t += gridDim.x * blockDim.x) {
		const int32_t u = cand_list[t];
		if (state[u]) continue;
		bool blocked = false, wait = false;
		for (int64_t q = vh_off[u]; q < vh_off[u + 1] && !blocked; ++q) {
			const int64_t h = vh_val[q];
			for (int k = 0; k < 8; ++k) {
				if (!rem[8 * h + k]) continue;
				const uint32_t w = hex[8 * h + k];
				if (w >= (uint32_t)u) continue;
				const uint8_t sw = ((volatile const uint8_t *)state)[w];
				if (sw == 1) { blocked = true; break; }
				if (sw == 0) wait = true;
			}
		}
		if (blocked) state[u] = 2;
		else if (!wait) state[u] = 1;
		else *undecided = 1;
	}
}
__global__ void nm_vertex_apply_kernel(const int32_t *__restrict__ cand_list, int n_cand, const uint8_t *__restrict__ state,
                                       const int64_t *__restrict__ vh_off, const uint32_t *__restrict__ vh_val, const uint32_t *__restrict__ hex,
                                       const uint8_t *__restrict__ rem, const int32_t *__restrict__ H_rev, uint8_t *__restrict__ flag)
{
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_cand; t += gridDim.x * blockDim.x) {
		const int32_t u = cand_list[t];
		if (state[u] != 1) continue;
		for (int64_t q = vh_off[u]; q < vh_off[u + 1]; ++q) {
			const int64_t h = vh_val[q];
			for (int k = 0; k < 8; ++k) if (rem[8 * h + k] && hex[8 * h + k] == (uint32_t)u) flag[H_rev[h]] = 0;
		}
	}
}

// ---- edge pass (ghm.cpp:2059-2075): a boundary edge with other than 2 boundary faces drops the hex behind the first
__global__ void nm_edge_candidates_kernel(int64_t ne, const uint8_t *__restrict__ E_boundary, const int64_t *__restrict__ ef_off,
                                          const uint32_t *__restrict__ ef_val, const uint8_t *__restrict__ F_boundary,
                                          const int64_t *__restrict__ fh_off, const uint32_t *__restrict__ fh_val, int32_t *__restrict__ hsel,
                                          int32_t *__restrict__ cand_list, int32_t *__restrict__ counters)
{
	for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
		hsel[e] = -1;
		if (!E_boundary[e]) continue;
		int n = 0; uint32_t first = 0;
		for (int64_t t = ef_off[e]; t < ef_off[e + 1]; ++t) {
			const uint32_t f = ef_val[t];
			if (!F_boundary[f]) continue;
			if (n == 0) first = f;
			++n;
		}
		if (n == 2 || n == 0) continue;
		hsel[e] = (int32_t)fh_val[fh_off[first]];
		cand_list[atomicAdd(&counters[0], 1)] = (int32_t)e;
	}
}
__global__ void nm_edge_resolve_kernel(const int32_t *__restrict__ cand_list, int n_cand, const int64_t *__restrict__ ef_off,
                                       const uint32_t *__restrict__ ef_val, const int64_t *__restrict__ fh_off, const uint32_t *__restrict__ fh_val,
                                       const uint32_t *__restrict__ H_fs, const uint32_t *__restrict__ F_es, const int32_t *__restrict__ hsel,
                                       uint8_t *__restrict__ state, int32_t *__restrict__ undecided)
{
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_cand; t += gridDim.x * blockDim.x) {
		const int32_t e = cand_list[t];
		if (state[e]) continue;
		bool blocked = false, wait = false;
		// the hexes around e; an earlier candidate w tags e iff e is one of the edges of the hex w drops
		for (int64_t q = ef_off[e]; q < ef_off[e + 1] && !blocked; ++q) {
			const uint32_t f = ef_val[q];
			for (int64_t r = fh_off[f]; r < fh_off[f + 1] && !blocked; ++r) {
				const int64_t h = fh_val[r];
				for (int j = 0; j < 24; ++j) {
					const uint32_t w = F_es[4 * (int64_t)H_fs[6 * h + j / 4] + (j & 3)];
					if (w >= (uint32_t)e || hsel[w] != (int32_t)h) continue;
					const uint8_t sw = ((volatile const uint8_t *)state)[w];
					if (sw == 1) { blocked = true; break; }
					if (sw == 0) wait = true;
				}
			}
		}
		if (blocked) state[e] = 2;
		else if (!wait) state[e] = 1;
		else *undecided = 1;
	}
}
__global__ void nm_edge_apply_kernel(const int32_t *__restrict__ cand_list, int n_cand, const uint8_t *__restrict__ state,
                                     const int32_t *__restrict__ hsel, const int32_t *__restrict__ H_rev, uint8_t *__restrict__ flag)
{
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_cand; t += gridDim.x * blockDim.x) {
		const int32_t e = cand_list[t];
		if (state[e] == 1) flag[H_rev[hsel[e]]] = 0;
	}
}

// runs clean_non_manifold_ve on device flags; sc is left holding the sub-mesh of the final flags
int non_manifold_dev(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, uint8_t *flag, SubConn &sc, cudaStream_t s) {
	const int blk = 256;
	int rounds = 0;
	DevBuf<int32_t> counters(4, s);
	while (true) {
		const fpohm_conn *c = sc.c;
		if (!c) break;                                       // nothing left inside
		const int64_t nv = sc.sub.nv, nh = sc.sub.nh, ne = c->nE;
		bool changed = false;
		{
			DevBuf<uint8_t> rem(8 * nh, s), state(nv, s);
			DevBuf<int32_t> cand(nv, s);
			rem.zero(); state.zero(); counters.zero();
			nm_vertex_candidates_kernel<<<grid_for(ctx, nv, 128), 128, 0, s>>>(nv, c->V_boundary.p, c->off[5].p, c->val[5].p, c->F_boundary.p, c->F_es.p,
				c->off[1].p, c->val[1].p, c->off[0].p, c->val[0].p, c->hex.p, rem.p, cand.p, counters.p);
			FPOHM_LAUNCH_CHECK(ctx);
			int32_t hc[2] = {0, 0};
			counters.download(hc, 2);
			FPOHM_CUDA(cudaStreamSynchronize(s));
			FPOHM_REQUIRE(hc[1] == 0, FPOHM_ERANGE, "fpohm_clean_non_manifold: more than %d boundary faces around one vertex", CLEAN_MAX_BFS);
			if (hc[0] > 0) {
				changed = true;
				while (true) {
					FPOHM_CUDA(cudaMemsetAsync(counters.p + 2, 0, 4, s));
					nm_vertex_resolve_kernel<<<grid_for(ctx, hc[0], blk), blk, 0, s>>>(cand.p, hc[0], c->off[6].p, c->val[6].p, c->hex.p, rem.p, state.p, counters.p + 2);
					FPOHM_LAUNCH_CHECK(ctx);
					int32_t und = 0;
					FPOHM_CUDA(cudaMemcpyAsync(&und, counters.p + 2, 4, cudaMemcpyDeviceToHost, s));
					FPOHM_CUDA(cudaStreamSynchronize(s));
					if (!und) break;
				}
				nm_vertex_apply_kernel<<<grid_for(ctx, hc[0], blk), blk, 0, s>>>(cand.p, hc[0], state.p, c->off[6].p, c->val[6].p, c->hex.p, rem.p, sc.sub.H_rev.p, flag);
				FPOHM_LAUNCH_CHECK(ctx);
			}
		}
		if (!changed) {
			DevBuf<int32_t> hsel(ne, s), cand(ne, s);
			DevBuf<uint8_t> state(ne, s);
			state.zero(); counters.zero();
			nm_edge_candidates_kernel<<<grid_for(ctx, ne, blk), blk, 0, s>>>(ne, c->E_boundary.p, c->off[1].p, c->val[1].p, c->F_boundary.p, c->off[0].p, c->val[0].p,
				hsel.p, cand.p, counters.p);
			FPOHM_LAUNCH_CHECK(ctx);
			int32_t hc = 0;
			counters.download(&hc, 1);
			FPOHM_CUDA(cudaStreamSynchronize(s));
			if (hc > 0) {
				changed = true;
				while (true) {
					FPOHM_CUDA(cudaMemsetAsync(counters.p + 2, 0, 4, s));
					nm_edge_resolve_kernel<<<grid_for(ctx, hc, blk), blk, 0, s>>>(cand.p, hc, c->off[1].p, c->val[1].p, c->off[0].p, c->val[0].p, c->H_fs.p, c->F_es.p,
						hsel.p, state.p, counters.p + 2);
					FPOHM_LAUNCH_CHECK(ctx);
					int32_t und = 0;
					FPOHM_CUDA(cudaMemcpyAsync(&und, counters.p + 2, 4, cudaMemcpyDeviceToHost, s));
					FPOHM_CUDA(cudaStreamSynchronize(s));
					if (!und) break;
				}
				nm_edge_apply_kernel<<<grid_for(ctx, hc, blk), blk, 0, s>>>(cand.p, hc, state.p, hsel.p, sc.sub.H_rev.p, flag);
				FPOHM_LAUNCH_CHECK(ctx);
			}
		}
		if (!changed) break;
		++rounds;
		build_sub(ctx, hex, H, nV, flag, sc, s);             // re_indexing_connectivity at the end of the round (ghm.cpp:2078)
	}
	return rounds;
}

// ---- drop_small_pieces
// parent[x] <= x always (the smaller root wins), so shortening a path to any ancestor is safe under concurrency
__device__ __forceinline__ int32_t uf_find(int32_t *parent, int32_t x) {
	volatile int32_t *vp = parent;
	int32_t cur = vp[x];
	if (cur != x) {
		int32_t prev = x, next;
		while (cur > (next = vp[cur])) { vp[prev] = next; prev = cur; cur = next; }
	}
	return cur;
}
__global__ void uf_init_kernel(int32_t *__restrict__ parent, int64_t n) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) parent[i] = (int32_t)i;
}
__global__ void uf_union_faces_kernel(int64_t nF, const int64_t *__restrict__ fh_off, const uint32_t *__restrict__ fh_val, int32_t *parent) {
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const int64_t o = fh_off[f];
		if (fh_off[f + 1] - o != 2) continue;
		int32_t a = (int32_t)fh_val[o], b = (int32_t)fh_val[o + 1];
		while (true) {
			a = uf_find(parent, a); b = uf_find(parent, b);
			if (a == b) break;
			if (a < b) { const int32_t t = a; a = b; b = t; }          // the larger root goes under the smaller
			if (atomicCAS(&parent[a], a, b) == a) break;
		}
	}
}
__global__ void uf_label_kernel(int32_t *parent, int64_t n, int32_t *__restrict__ label, int32_t *__restrict__ n_roots) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const int32_t r = uf_find(parent, (int32_t)i);
		label[i] = r;
		if (r == (int32_t)i) atomicAdd(n_roots, 1);
	}
}
__global__ void piece_sizes_kernel(int64_t nF, const int64_t *__restrict__ fh_off, const uint32_t *__restrict__ fh_val, const int32_t *__restrict__ label,
                                   int32_t *__restrict__ size)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const int64_t o = fh_off[f];
		if (fh_off[f + 1] - o == 2) atomicAdd(&size[label[fh_val[o]]], 1);
	}
}
// first strict maximum in order of discovery == largest size, ties to the smallest root
__global__ void best_piece_kernel(int64_t n, const int32_t *__restrict__ label, const int32_t *__restrict__ size, unsigned long long *__restrict__ best) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		if (label[i] == (int32_t)i) atomicMax(best, ((unsigned long long)(uint32_t)size[i] << 32) | (0xffffffffu - (uint32_t)i));
}
__global__ void keep_piece_kernel(int64_t n, const int32_t *__restrict__ label, const unsigned long long *__restrict__ best, const int32_t *__restrict__ H_rev,
                                  uint8_t *__restrict__ flag)
{
	const int32_t root = (int32_t)(0xffffffffu - (uint32_t)(*best & 0xffffffffull));
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		if (label[i] == root) flag[H_rev[i]] = 1;
}

int64_t drop_small_dev(fpohm_ctx *ctx, int64_t H, uint8_t *flag, const SubConn &sc, cudaStream_t s) {
	if (!sc.c) return 0;
	const int blk = 256;
	const fpohm_conn *c = sc.c;
	const int64_t nh = sc.sub.nh;
	DevBuf<int32_t> parent(nh, s), label(nh, s), size(nh, s), nroots(1, s);
	DevBuf<unsigned long long> best(1, s);
	size.zero(); nroots.zero(); best.zero();
	uf_init_kernel<<<grid_for(ctx, nh, blk), blk, 0, s>>>(parent.p, nh);
	FPOHM_LAUNCH_CHECK(ctx);
	uf_union_faces_kernel<<<grid_for(ctx, c->nF, blk), blk, 0, s>>>(c->nF, c->off[0].p, c->val[0].p, parent.p);
	FPOHM_LAUNCH_CHECK(ctx);
	uf_label_kernel<<<grid_for(ctx, nh, blk), blk, 0, s>>>(parent.p, nh, label.p, nroots.p);
	FPOHM_LAUNCH_CHECK(ctx);
	int32_t nr = 0;
	nroots.download(&nr, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (nr > 1) {
		piece_sizes_kernel<<<grid_for(ctx, c->nF, blk), blk, 0, s>>>(c->nF, c->off[0].p, c->val[0].p, label.p, size.p);
		FPOHM_LAUNCH_CHECK(ctx);
		best_piece_kernel<<<grid_for(ctx, nh, blk), blk, 0, s>>>(nh, label.p, size.p, best.p);
		FPOHM_LAUNCH_CHECK(ctx);
		FPOHM_CUDA(cudaMemsetAsync(flag, 0, (size_t)H, s));
		keep_piece_kernel<<<grid_for(ctx, nh, blk), blk, 0, s>>>(nh, label.p, best.p, sc.sub.H_rev.p, flag);
		FPOHM_LAUNCH_CHECK(ctx);
	}
	return nr;
}

// ---- tail of clean_hex_mesh
__global__ void medial_flags_kernel(int64_t nF, const int64_t *__restrict__ fh_off, const uint32_t *__restrict__ fh_val, const uint32_t *__restrict__ F_vs,
                                    const uint8_t *__restrict__ flag, uint8_t *__restrict__ F_medial, uint8_t *__restrict__ V_medial)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const int64_t o = fh_off[f];
		const bool boundary = fh_off[f + 1] - o == 1;
		const bool m = boundary ? flag[fh_val[o]] != 0 : (flag[fh_val[o]] != flag[fh_val[o + 1]]);
		F_medial[f] = m ? 1 : 0;
		if (m) for (int k = 0; k < 4; ++k) V_medial[F_vs[4 * f + k]] = 1;
	}
}

int tag_dev(fpohm_ctx *ctx, const fpohm_conn *c, uint8_t *flag, cudaStream_t s) {
	const int blk = 256;
	const int64_t H = c->H;
	DevBuf<int32_t> nb(6 * H, s), moved(1, s);
	hex_neighbours_kernel<<<grid_for(ctx, 6 * H, blk), blk, 0, s>>>(c->H_fs.p, c->off[0].p, c->val[0].p, H, nb.p);
	FPOHM_LAUNCH_CHECK(ctx);
	DevBuf<uint8_t> old(H, s), a(H, s), b(H, s);
	FPOHM_CUDA(cudaMemcpyAsync(old.p, flag, (size_t)H, cudaMemcpyDeviceToDevice, s));
	FPOHM_CUDA(cudaMemcpyAsync(a.p, flag, (size_t)H, cudaMemcpyDeviceToDevice, s));
	uint8_t *cur = a.p, *next = b.p;
	int sweeps = 0;
	while (true) {
		moved.zero();
		tag_sweep_kernel<<<grid_for(ctx, H, blk), blk, 0, s>>>(nb.p, H, old.p, cur, next, moved.p);
		FPOHM_LAUNCH_CHECK(ctx);
		++sweeps;
		int32_t m = 0;
		moved.download(&m, 1);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		uint8_t *t = cur; cur = next; next = t;
		if (!m) break;
	}
	FPOHM_CUDA(cudaMemcpyAsync(flag, cur, (size_t)H, cudaMemcpyDeviceToDevice, s));
	return sweeps;
}

// range check of the hex list on the device (the host loop over 8 H ids was a fifth of clean_hex_mesh's wall clock)
__global__ void validate_ids_kernel(const uint32_t *__restrict__ hex, int64_t n, int64_t nV, unsigned long long *__restrict__ first_bad) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		if ((int64_t)hex[t] >= nV) atomicMin(first_bad, (unsigned long long)t);
}

void check_hex(const uint32_t *hex, int64_t H, int64_t nV, const char *who) {
	for (int64_t i = 0; i < 8 * H; ++i) FPOHM_REQUIRE((int64_t)hex[i] < nV, FPOHM_EINVAL, "%s: corner id %u out of range at %lld", who, hex[i], (long long)i);
}

} // namespace

extern "C" {

int fpohm_reorder_hexes(fpohm_ctx *ctx, const double *V, int64_t nV, uint32_t *hex, int64_t H, int64_t *n_mirrored) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && V && hex && nV > 0 && H >= 0, FPOHM_EINVAL, "fpohm_reorder_hexes: bad argument");
	if (n_mirrored) *n_mirrored = 0;
	if (H == 0) return FPOHM_OK;
	check_hex(hex, H, nV, "fpohm_reorder_hexes");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dV(3 * nV, s);
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<unsigned long long> cnt(1, s);
	dV.upload(V, 3 * nV); dhex.upload(hex, 8 * H); cnt.zero();
	KernelTimer t(ctx, s);
	reorder_hexes_kernel<<<grid_for(ctx, H, 128), 128, 0, s>>>(dV.p, dhex.p, H, cnt.p);
	FPOHM_LAUNCH_CHECK(ctx);
	t.stop();
	unsigned long long n = 0;
	cnt.download(&n, 1);
	dhex.download(hex, 8 * H);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (n_mirrored) *n_mirrored = (int64_t)n;
	FPOHM_API_END
}

int fpohm_tag_uneven_elements(fpohm_ctx *ctx, const fpohm_conn *conn, uint8_t *H_flag, int32_t *n_sweeps) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && conn && H_flag, FPOHM_EINVAL, "fpohm_tag_uneven_elements: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<uint8_t> flag(conn->H, s);
	flag.upload(H_flag, conn->H);
	KernelTimer t(ctx, s);
	const int sweeps = tag_dev(ctx, conn, flag.p, s);
	t.stop();
	flag.download(H_flag, conn->H);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (n_sweeps) *n_sweeps = sweeps;
	FPOHM_API_END
}

int fpohm_reindex_submesh(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, const uint8_t *H_flag, int32_t *V_map,
                          int32_t *V_map_reverse, int64_t *n_sub_v, int32_t *H_map_reverse, int64_t *n_sub_h, uint32_t *sub_hex)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && hex && H_flag && H > 0 && nV > 0, FPOHM_EINVAL, "fpohm_reindex_submesh: bad argument");
	check_hex(hex, H, nV, "fpohm_reindex_submesh");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<uint8_t> flag(H, s);
	dhex.upload(hex, 8 * H); flag.upload(H_flag, H);
	SubMesh sub;
	KernelTimer t(ctx, s);
	reindex_dev(ctx, dhex.p, H, nV, flag.p, sub, s);
	t.stop();
	if (n_sub_v) *n_sub_v = sub.nv;
	if (n_sub_h) *n_sub_h = sub.nh;
	if (V_map) sub.V_map.download(V_map, nV);
	if (V_map_reverse) sub.V_rev.download(V_map_reverse, sub.nv);
	if (H_map_reverse) sub.H_rev.download(H_map_reverse, sub.nh);
	if (sub_hex) sub.hex.download(sub_hex, 8 * sub.nh);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_clean_non_manifold(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, uint8_t *H_flag, int32_t *n_rounds) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && hex && H_flag && H > 0 && nV > 0, FPOHM_EINVAL, "fpohm_clean_non_manifold: bad argument");
	check_hex(hex, H, nV, "fpohm_clean_non_manifold");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<uint8_t> flag(H, s);
	dhex.upload(hex, 8 * H); flag.upload(H_flag, H);
	SubConn sc;
	build_sub(ctx, dhex.p, H, nV, flag.p, sc, s);
	const int rounds = non_manifold_dev(ctx, dhex.p, H, nV, flag.p, sc, s);
	flag.download(H_flag, H);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (n_rounds) *n_rounds = rounds;
	FPOHM_API_END
}

int fpohm_drop_small_pieces(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, uint8_t *H_flag, int64_t *n_pieces) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && hex && H_flag && H > 0 && nV > 0, FPOHM_EINVAL, "fpohm_drop_small_pieces: bad argument");
	check_hex(hex, H, nV, "fpohm_drop_small_pieces");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<uint8_t> flag(H, s);
	dhex.upload(hex, 8 * H); flag.upload(H_flag, H);
	SubConn sc;
	build_sub(ctx, dhex.p, H, nV, flag.p, sc, s);
	const int64_t np = drop_small_dev(ctx, H, flag.p, sc, s);
	flag.download(H_flag, H);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	if (n_pieces) *n_pieces = np;
	FPOHM_API_END
}

int fpohm_medial_surface_flags(fpohm_ctx *ctx, const fpohm_conn *conn, const uint8_t *H_flag, uint8_t *F_medial, uint8_t *V_medial) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && conn && H_flag && F_medial && V_medial, FPOHM_EINVAL, "fpohm_medial_surface_flags: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<uint8_t> flag(conn->H, s), fm(conn->nF, s), vm(conn->nV, s);
	flag.upload(H_flag, conn->H); vm.zero();
	medial_flags_kernel<<<grid_for(ctx, conn->nF, 256), 256, 0, s>>>(conn->nF, conn->off[0].p, conn->val[0].p, conn->F_vs.p, flag.p, fm.p, vm.p);
	FPOHM_LAUNCH_CHECK(ctx);
	fm.download(F_medial, conn->nF); vm.download(V_medial, conn->nV);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_clean_hex_mesh(fpohm_ctx *ctx, fpohm_mesh *surface, const double *V, int64_t nV, uint32_t *hex, int64_t H, const fpohm_conn *conn,
                         double *signed_dis, uint8_t *H_flag, uint8_t *F_medial, uint8_t *V_medial, int64_t stats[6])
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && surface && V && hex && H_flag && H > 0 && nV > 0, FPOHM_EINVAL, "fpohm_clean_hex_mesh: bad argument");
	FPOHM_REQUIRE(!conn || (conn->H == H && conn->nV == nV), FPOHM_EINVAL, "fpohm_clean_hex_mesh: conn belongs to another mesh");
	int64_t st[6] = {0, 0, 0, 0, 0, 0};      // mirrored hexes, tagging sweeps, non-manifold rounds, pieces, hexes kept, vertices kept
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	// one upload of V and hex; every stage below works on the device copies
	DevBuf<double> dV(3 * nV, s), dS(H, s);
	DevBuf<uint32_t> dhex(8 * H, s);
	DevBuf<unsigned long long> cnt(1, s);
	DevBuf<uint8_t> flag(H, s);
	dV.upload(V, 3 * nV); dhex.upload(hex, 8 * H); cnt.zero();
	{
		DevBuf<unsigned long long> bad(1, s);
		FPOHM_CUDA(cudaMemsetAsync(bad.p, 0xff, 8, s));
		validate_ids_kernel<<<grid_for(ctx, 8 * H, 256), 256, 0, s>>>(dhex.p, 8 * H, nV, bad.p);
		FPOHM_LAUNCH_CHECK(ctx);
		unsigned long long hb = 0;
		bad.download(&hb, 1);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		FPOHM_REQUIRE(hb == ~0ull, FPOHM_EINVAL, "fpohm_clean_hex_mesh: corner id %u out of range at %llu", hex[hb], hb);
	}
	reorder_hexes_kernel<<<grid_for(ctx, H, 128), 128, 0, s>>>(dV.p, dhex.p, H, cnt.p);              // ghm.cpp:1935
	FPOHM_LAUNCH_CHECK(ctx);
	classify_hexes_dev(ctx, surface, dV.p, dhex.p, H, dS.p, flag.p, s);                              // ghm.cpp:1937-1951
	{
		unsigned long long n = 0;
		cnt.download(&n, 1);
		dhex.download(hex, 8 * H);
		if (signed_dis) dS.download(signed_dis, H);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		st[0] = (int64_t)n;
	}
	fpohm_conn *own = nullptr;
	if (!conn) {
		DevBuf<uint32_t> copy(8 * H, s);
		FPOHM_CUDA(cudaMemcpyAsync(copy.p, dhex.p, 32 * (size_t)H, cudaMemcpyDeviceToDevice, s));
		own = conn_build_dev(ctx, std::move(copy), H, nV, false);
		conn = own;
	}
	try {
		st[1] = tag_dev(ctx, conn, flag.p, s);                                                  // ghm.cpp:1954
		SubConn sc;
		build_sub(ctx, dhex.p, H, nV, flag.p, sc, s);                                           // ghm.cpp:1955
		if (sc.sub.nh > 0) {                                                                    // "no elements inside the object" returns here (:1957-1961)
			st[2] = non_manifold_dev(ctx, dhex.p, H, nV, flag.p, sc, s);                        // ghm.cpp:1963
			st[3] = drop_small_dev(ctx, H, flag.p, sc, s);                                      // ghm.cpp:1965
			if (st[3] > 1) reindex_dev(ctx, dhex.p, H, nV, flag.p, sc.sub, s);
			st[4] = sc.sub.nh; st[5] = sc.sub.nv;
			if (F_medial || V_medial) {
				DevBuf<uint8_t> fm(conn->nF, s), vm(nV, s);
				vm.zero();
				medial_flags_kernel<<<grid_for(ctx, conn->nF, 256), 256, 0, s>>>(conn->nF, conn->off[0].p, conn->val[0].p, conn->F_vs.p, flag.p, fm.p, vm.p);
				FPOHM_LAUNCH_CHECK(ctx);
				if (F_medial) fm.download(F_medial, conn->nF);
				if (V_medial) vm.download(V_medial, nV);
			}
		} else {
			if (F_medial) memset(F_medial, 0, (size_t)conn->nF);
			if (V_medial) memset(V_medial, 0, (size_t)nV);
		}
		flag.download(H_flag, H);
		FPOHM_CUDA(cudaStreamSynchronize(s));
	} catch (...) { delete own; throw; }
	delete own;
	if (stats) for (int i = 0; i < 6; ++i) stats[i] = st[i];
	FPOHM_API_END
}

} // extern "C"

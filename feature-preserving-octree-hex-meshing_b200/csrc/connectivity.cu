// build_connectivity, Hex branch (gf.cpp:121-186) + adjacency relations (gf.cpp:226-264).
//
// The reference sorts std::tuple<v0..v3 sorted, id, hex, j> (faces) and <vmin, vmax, face, j> (edges) and walks the
// sorted lists; every adjacency list is filled by loops that visit elements in ascending id.  All of it is therefore
// reproduced by STABLE LSD radix sorts on the same composite keys (ties keep the original 6h+j / 4f+j order), run-head
// flags + inclusive scan for the ids, and CSR offsets by binary search.  Integer work only: bit-exact.
#include "conn.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

using namespace fpohm;

namespace {

__constant__ int c_hex_face[6][4] = {{0, 1, 2, 3}, {4, 7, 6, 5}, {0, 4, 5, 1}, {0, 3, 7, 4}, {3, 2, 6, 7}, {1, 5, 6, 2}}; // global_types.h:154-162

__device__ __forceinline__ void sort4(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
	uint32_t t;
	if (a > b) { t = a; a = b; b = t; }
	if (c > d) { t = c; c = d; d = t; }
	if (a > c) { t = a; a = c; c = t; }
	if (b > d) { t = b; b = d; d = t; }
	if (b > c) { t = b; b = c; c = t; }
}

__global__ void face_keys_kernel(const uint32_t *__restrict__ hex, int64_t H, uint64_t *__restrict__ hi, uint64_t *__restrict__ lo,
                                 uint32_t *__restrict__ id)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 6 * H; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = t / 6; const int j = (int)(t % 6);
		uint32_t a = hex[8 * i + c_hex_face[j][0]], b = hex[8 * i + c_hex_face[j][1]], c = hex[8 * i + c_hex_face[j][2]], d = hex[8 * i + c_hex_face[j][3]];
		sort4(a, b, c, d);
		hi[t] = ((uint64_t)a << 32) | b;
		lo[t] = ((uint64_t)c << 32) | d;
		id[t] = (uint32_t)t;
	}
}
__global__ void gather_u64_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, int64_t n, uint64_t *__restrict__ dst) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) dst[t] = src[idx[t]];
}
// run heads of the sorted (hi, lo) tuples
__global__ void face_heads_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, int64_t n, int32_t *__restrict__ head) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		head[t] = (t == 0 || hi[t] != hi[t - 1] || lo[t] != lo[t - 1]) ? 1 : 0;
}
__global__ void heads_u64_kernel(const uint64_t *__restrict__ k, int64_t n, int32_t *__restrict__ head) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		head[t] = (t == 0 || k[t] != k[t - 1]) ? 1 : 0;
}
// faces: F.vs = corner order of the FIRST tuple of the run (gf.cpp:143), boundary iff the run has one tuple (:147-150),
// H.fs[hex][j] = face id (:152), F.neighbor_hs = hexes of the run in sorted order (== ascending hex id, gf.cpp:228-230)
__global__ void faces_kernel(const uint32_t *__restrict__ hex, const uint32_t *__restrict__ id, const int32_t *__restrict__ head,
                             const int32_t *__restrict__ fid_incl, int64_t n, uint32_t *__restrict__ F_vs, uint8_t *__restrict__ F_boundary,
                             uint32_t *__restrict__ H_fs, int64_t *__restrict__ nhs_off, uint32_t *__restrict__ nhs_val, int64_t nF)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t f = fid_incl[t] - 1;
		const uint32_t tid = id[t];
		const int64_t i = tid / 6; const int j = (int)(tid % 6);
		H_fs[6 * i + j] = (uint32_t)f;
		nhs_val[t] = (uint32_t)i;
		if (head[t]) {
#pragma unroll
			for (int k = 0; k < 4; ++k) F_vs[4 * f + k] = hex[8 * i + c_hex_face[j][k]];
			F_boundary[f] = (t + 1 == n || head[t + 1]) ? 1 : 0;
			nhs_off[f] = t;
		}
		if (t == n - 1) nhs_off[nF] = n;
	}
}
__global__ void edge_keys_kernel(const uint32_t *__restrict__ F_vs, int64_t nF, uint64_t *__restrict__ key, uint32_t *__restrict__ id) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 4 * nF; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = t >> 2; const int j = (int)(t & 3);
		uint32_t v0 = F_vs[4 * i + j], v1 = F_vs[4 * i + ((j + 1) & 3)];
		if (v0 > v1) { const uint32_t x = v0; v0 = v1; v1 = x; }
		key[t] = ((uint64_t)v0 << 32) | v1;
		id[t] = (uint32_t)t;
	}
}
__global__ void edges_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ id, const int32_t *__restrict__ head,
                             const int32_t *__restrict__ eid_incl, int64_t n, uint32_t *__restrict__ E_vs, uint32_t *__restrict__ F_es,
                             int64_t *__restrict__ nfs_off, uint32_t *__restrict__ nfs_val, int64_t nE)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t e = eid_incl[t] - 1;
		const uint32_t tid = id[t];
		F_es[tid] = (uint32_t)e;              // Fs[face].es[j], tid = 4*face + j
		nfs_val[t] = tid >> 2;
		if (head[t]) { E_vs[2 * e] = (uint32_t)(key[t] >> 32); E_vs[2 * e + 1] = (uint32_t)key[t]; nfs_off[e] = t; }
		if (t == n - 1) nfs_off[nE] = n;
	}
}
// gf.cpp:179-185
__global__ void boundary_kernel(const uint8_t *__restrict__ F_boundary, const uint32_t *__restrict__ F_es, const uint32_t *__restrict__ E_vs,
                                int64_t nF, uint8_t *__restrict__ E_boundary, uint8_t *__restrict__ V_boundary)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 4 * nF; t += (int64_t)gridDim.x * blockDim.x) {
		if (!F_boundary[t >> 2]) continue;
		const uint32_t e = F_es[t];
		E_boundary[e] = 1;
		V_boundary[E_vs[2 * e]] = 1; V_boundary[E_vs[2 * e + 1]] = 1;
	}
}
// generic "for i ascending: list[key(i)].push_back(value(i))": keys + payload in visiting order
__global__ void vfs_items_kernel(const uint32_t *__restrict__ F_vs, int64_t nF, uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 4 * nF; t += (int64_t)gridDim.x * blockDim.x) { key[t] = F_vs[t]; val[t] = (uint32_t)(t >> 2); }
}
__global__ void ves_items_kernel(const uint32_t *__restrict__ E_vs, int64_t nE, uint32_t *__restrict__ key, uint32_t *__restrict__ val_e, uint32_t *__restrict__ val_v) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 2 * nE; t += (int64_t)gridDim.x * blockDim.x) {
		key[t] = E_vs[t]; val_e[t] = (uint32_t)(t >> 1); val_v[t] = E_vs[t ^ 1];
	}
}
__global__ void vhs_items_kernel(const uint32_t *__restrict__ hex, int64_t H, uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 8 * H; t += (int64_t)gridDim.x * blockDim.x) { key[t] = hex[t]; val[t] = (uint32_t)(t >> 3); }
}
// E.neighbor_hs candidates: for every (edge, face) incidence, the face's hexes (gf.cpp:250-258)
__global__ void ehs_count_kernel(const uint32_t *__restrict__ nfs_val, int64_t n, const int64_t *__restrict__ fnhs_off, int64_t *__restrict__ cnt) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const uint32_t f = nfs_val[t];
		cnt[t] = fnhs_off[f + 1] - fnhs_off[f];
	}
}
__global__ void ehs_items_kernel(const uint32_t *__restrict__ nfs_val, const int32_t *__restrict__ eid_incl, int64_t n,
                                 const int64_t *__restrict__ fnhs_off, const uint32_t *__restrict__ fnhs_val, const int64_t *__restrict__ pos,
                                 uint64_t *__restrict__ key)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const uint32_t f = nfs_val[t];
		const uint64_t e = (uint64_t)(eid_incl[t] - 1);
		int64_t o = pos[t];
		for (int64_t k = fnhs_off[f]; k < fnhs_off[f + 1]; ++k) key[o++] = (e << 32) | fnhs_val[k];
	}
}
__global__ void split_key_kernel(const uint64_t *__restrict__ key, int64_t n, uint32_t *__restrict__ hi, uint32_t *__restrict__ lo) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) { hi[t] = (uint32_t)(key[t] >> 32); lo[t] = (uint32_t)key[t]; }
}
// off[k] = lower_bound(sorted_keys, k) for k in [0, nkeys]
__global__ void csr_offsets_kernel(const uint32_t *__restrict__ sorted, int64_t n, int64_t nkeys, int64_t *__restrict__ off) {
	for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k <= nkeys; k += (int64_t)gridDim.x * blockDim.x) {
		int64_t lo = 0, hi = n;
		while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)sorted[mid] < k) lo = mid + 1; else hi = mid; }
		off[k] = lo;
	}
}

int bits_for(int64_t n) { int b = 1; while ((1ll << b) < n) ++b; return b; }

struct Ops {
	fpohm_ctx *ctx; cudaStream_t s;
	template <class K, class V>
	void sort_pairs(DevBuf<K> &k, DevBuf<V> &v, int64_t n, int begin_bit, int end_bit) {
		if (n == 0) return;
		DevBuf<K> k2(n, s); DevBuf<V> v2(n, s);
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k.p, k2.p, v.p, v2.p, n, begin_bit, end_bit, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k.p, k2.p, v.p, v2.p, n, begin_bit, end_bit, s));
		ctx->launches += 3;
		k = std::move(k2); v = std::move(v2);
	}
	void incl_scan(DevBuf<int32_t> &in, DevBuf<int32_t> &out, int64_t n) {
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, in.p, out.p, n, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, in.p, out.p, n, s));
		ctx->launches += 1;
	}
	void excl_scan64(DevBuf<int64_t> &in, DevBuf<int64_t> &out, int64_t n) {
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in.p, out.p, n, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in.p, out.p, n, s));
		ctx->launches += 1;
	}
	int32_t last_i32(DevBuf<int32_t> &b, int64_t n) {
		int32_t v = 0;
		FPOHM_CUDA(cudaMemcpyAsync(&v, b.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
		FPOHM_CUDA(cudaStreamSynchronize(s));
		return v;
	}
	// "for t ascending: list[key[t]].push_back(val[t])" -> CSR (stable sort keeps the visiting order inside a list)
	void csr_from_items(DevBuf<uint32_t> &key, DevBuf<uint32_t> &val, int64_t n, int64_t nkeys, DevBuf<int64_t> &off, DevBuf<uint32_t> &out) {
		sort_pairs(key, val, n, 0, bits_for(nkeys));
		off.alloc(nkeys + 1, s);
		csr_offsets_kernel<<<grid_for(ctx, nkeys + 1, 256), 256, 0, s>>>(key.p, n, nkeys, off.p);
		FPOHM_LAUNCH_CHECK(ctx);
		out = std::move(val);
	}
};

} // namespace

// build_connectivity on a hex list that is already on the device (takes the buffer over: it stays in the result as
// c->hex).  `full` = false skips E.neighbor_hs, V.neighbor_vs and V.neighbor_es (two sorts and two host read-backs that
// the cleaning stages of cleaning.cu never look at).
fpohm_conn *fpohm::conn_build_dev(fpohm_ctx *ctx, DevBuf<uint32_t> &&hex_in, int64_t H, int64_t nV, bool full) {
	FPOHM_REQUIRE(12 * H < (1ll << 31), FPOHM_ERANGE, "fpohm_hex_connectivity: %lld hexes overflow uint32 tuple ids", (long long)H);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	Ops ops{ctx, s};
	fpohm_conn *c = new fpohm_conn;
	try {
		c->ctx = ctx; c->H = H; c->nV = nV;
		c->hex = std::move(hex_in);
		DevBuf<uint32_t> &dhex = c->hex;
		KernelTimer timer(ctx, s);
		const int vb = bits_for(nV);
		// ---- faces: stable LSD sort on (lo = v2,v3) then (hi = v0,v1); ties keep id = 6h+j order
		const int64_t n6 = 6 * H;
		DevBuf<uint64_t> hi(n6, s), lo(n6, s);
		DevBuf<uint32_t> id(n6, s);
		face_keys_kernel<<<grid_for(ctx, n6, blk), blk, 0, s>>>(dhex.p, H, hi.p, lo.p, id.p);
		FPOHM_LAUNCH_CHECK(ctx);
		{
			DevBuf<uint64_t> lo_k(n6, s);
			FPOHM_CUDA(cudaMemcpyAsync(lo_k.p, lo.p, 8 * (size_t)n6, cudaMemcpyDeviceToDevice, s));
			ops.sort_pairs(lo_k, id, n6, 0, 32 + vb);
			DevBuf<uint64_t> hi_k(n6, s);
			gather_u64_kernel<<<grid_for(ctx, n6, blk), blk, 0, s>>>(hi.p, id.p, n6, hi_k.p);
			FPOHM_LAUNCH_CHECK(ctx);
			ops.sort_pairs(hi_k, id, n6, 0, 32 + vb);
			DevBuf<uint64_t> lo_s(n6, s);
			gather_u64_kernel<<<grid_for(ctx, n6, blk), blk, 0, s>>>(lo.p, id.p, n6, lo_s.p);
			FPOHM_LAUNCH_CHECK(ctx);
			hi = std::move(hi_k); lo = std::move(lo_s);
		}
		DevBuf<int32_t> head(n6, s), fid(n6, s);
		face_heads_kernel<<<grid_for(ctx, n6, blk), blk, 0, s>>>(hi.p, lo.p, n6, head.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ops.incl_scan(head, fid, n6);
		const int64_t nF = ops.last_i32(fid, n6);
		c->nF = nF;
		c->F_vs.alloc(4 * nF, s); c->F_boundary.alloc(nF, s); c->H_fs.alloc(6 * H, s);
		c->off[0].alloc(nF + 1, s); c->val[0].alloc(n6, s); c->tot[0] = n6;
		faces_kernel<<<grid_for(ctx, n6, blk), blk, 0, s>>>(dhex.p, id.p, head.p, fid.p, n6, c->F_vs.p, c->F_boundary.p, c->H_fs.p,
			c->off[0].p, c->val[0].p, nF);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- edges: sort (vmin, vmax) stable over 4f+j
		const int64_t n4 = 4 * nF;
		DevBuf<uint64_t> ek(n4, s);
		DevBuf<uint32_t> eid_t(n4, s);
		edge_keys_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(c->F_vs.p, nF, ek.p, eid_t.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ops.sort_pairs(ek, eid_t, n4, 0, 32 + vb);
		DevBuf<int32_t> ehead(n4, s), eid(n4, s);
		heads_u64_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(ek.p, n4, ehead.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ops.incl_scan(ehead, eid, n4);
		const int64_t nE = ops.last_i32(eid, n4);
		c->nE = nE;
		c->E_vs.alloc(2 * nE, s); c->F_es.alloc(4 * nF, s);
		c->off[1].alloc(nE + 1, s); c->val[1].alloc(n4, s); c->tot[1] = n4;
		edges_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(ek.p, eid_t.p, ehead.p, eid.p, n4, c->E_vs.p, c->F_es.p, c->off[1].p, c->val[1].p, nE);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- boundary flags
		c->E_boundary.alloc(nE, s); c->V_boundary.alloc(nV, s);
		c->E_boundary.zero(); c->V_boundary.zero();
		boundary_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(c->F_boundary.p, c->F_es.p, c->E_vs.p, nF, c->E_boundary.p, c->V_boundary.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- E.neighbor_hs: sorted unique hexes over the edge's faces
		if (full) {
			DevBuf<int64_t> cnt(n4 + 1, s), pos(n4 + 1, s);
			cnt.zero();
			ehs_count_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(c->val[1].p, n4, c->off[0].p, cnt.p);
			FPOHM_LAUNCH_CHECK(ctx);
			ops.excl_scan64(cnt, pos, n4 + 1);
			int64_t tot = 0;
			FPOHM_CUDA(cudaMemcpyAsync(&tot, pos.p + n4, 8, cudaMemcpyDeviceToHost, s));
			FPOHM_CUDA(cudaStreamSynchronize(s));
			DevBuf<uint64_t> key(tot, s), skey(tot, s), ukey(tot, s);
			ehs_items_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(c->val[1].p, eid.p, n4, c->off[0].p, c->val[0].p, pos.p, key.p);
			FPOHM_LAUNCH_CHECK(ctx);
			size_t tb = 0;
			const int eb = bits_for(nE);
			FPOHM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, key.p, skey.p, tot, 0, 32 + eb, s));
			DevBuf<uint8_t> tmp((int64_t)tb, s);
			FPOHM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, key.p, skey.p, tot, 0, 32 + eb, s));
			DevBuf<int64_t> ucnt(1, s);
			size_t tb2 = 0;
			FPOHM_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, skey.p, ukey.p, ucnt.p, tot, s));
			DevBuf<uint8_t> tmp2((int64_t)tb2, s);
			FPOHM_CUDA(cub::DeviceSelect::Unique(tmp2.p, tb2, skey.p, ukey.p, ucnt.p, tot, s));
			ctx->launches += 5;
			int64_t nu = 0;
			ucnt.download(&nu, 1);
			FPOHM_CUDA(cudaStreamSynchronize(s));
			DevBuf<uint32_t> ehi(nu, s);
			c->val[2].alloc(nu, s); c->tot[2] = nu;
			split_key_kernel<<<grid_for(ctx, nu, blk), blk, 0, s>>>(ukey.p, nu, ehi.p, c->val[2].p);
			FPOHM_LAUNCH_CHECK(ctx);
			c->off[2].alloc(nE + 1, s);
			csr_offsets_kernel<<<grid_for(ctx, nE + 1, blk), blk, 0, s>>>(ehi.p, nu, nE, c->off[2].p);
			FPOHM_LAUNCH_CHECK(ctx);
			FPOHM_CUDA(cudaStreamSynchronize(s));
		}
		// ---- V.neighbor_es / V.neighbor_vs (gf.cpp:241-248)
		if (full) {
			const int64_t n2 = 2 * nE;
			DevBuf<uint32_t> key(n2, s), ve(n2, s), vv(n2, s), key2(n2, s);
			ves_items_kernel<<<grid_for(ctx, n2, blk), blk, 0, s>>>(c->E_vs.p, nE, key.p, ve.p, vv.p);
			FPOHM_LAUNCH_CHECK(ctx);
			FPOHM_CUDA(cudaMemcpyAsync(key2.p, key.p, 4 * (size_t)n2, cudaMemcpyDeviceToDevice, s));
			ops.csr_from_items(key, ve, n2, nV, c->off[4], c->val[4]); c->tot[4] = n2;
			ops.csr_from_items(key2, vv, n2, nV, c->off[3], c->val[3]); c->tot[3] = n2;
		}
		// ---- V.neighbor_fs (gf.cpp:236-239)
		{
			DevBuf<uint32_t> key(n4, s), v(n4, s);
			vfs_items_kernel<<<grid_for(ctx, n4, blk), blk, 0, s>>>(c->F_vs.p, nF, key.p, v.p);
			FPOHM_LAUNCH_CHECK(ctx);
			ops.csr_from_items(key, v, n4, nV, c->off[5], c->val[5]); c->tot[5] = n4;
		}
		// ---- V.neighbor_hs (gf.cpp:261-263)
		{
			const int64_t n8 = 8 * H;
			DevBuf<uint32_t> key(n8, s), v(n8, s);
			vhs_items_kernel<<<grid_for(ctx, n8, blk), blk, 0, s>>>(dhex.p, H, key.p, v.p);
			FPOHM_LAUNCH_CHECK(ctx);
			ops.csr_from_items(key, v, n8, nV, c->off[6], c->val[6]); c->tot[6] = n8;
		}
		timer.stop();
		FPOHM_CUDA(cudaStreamSynchronize(s));
	} catch (...) { delete c; throw; }
	return c;
}

extern "C" {

int fpohm_hex_connectivity(fpohm_ctx *ctx, const uint32_t *hex, int64_t H, int64_t nV, fpohm_conn **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && hex && out && H > 0 && nV > 0, FPOHM_EINVAL, "fpohm_hex_connectivity: bad argument");
	FPOHM_REQUIRE(12 * H < (1ll << 31), FPOHM_ERANGE, "fpohm_hex_connectivity: %lld hexes overflow uint32 tuple ids", (long long)H);
	for (int64_t i = 0; i < 8 * H; ++i)
		FPOHM_REQUIRE((int64_t)hex[i] < nV, FPOHM_EINVAL, "fpohm_hex_connectivity: corner id %u out of range at %lld", hex[i], (long long)i);
	DeviceGuard g(ctx->device);
	DevBuf<uint32_t> dhex(8 * H, ctx->stream);
	dhex.upload(hex, 8 * H);
	*out = conn_build_dev(ctx, std::move(dhex), H, nV, true);
	FPOHM_API_END
}

int fpohm_conn_sizes(const fpohm_conn *c, int64_t *nF, int64_t *nE) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(c, FPOHM_EINVAL, "fpohm_conn_sizes: null");
	if (nF) *nF = c->nF;
	if (nE) *nE = c->nE;
	FPOHM_API_END
}

int fpohm_conn_fixed(const fpohm_conn *c, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs,
                     uint8_t *E_boundary, uint8_t *V_boundary, uint32_t *H_fs)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(c, FPOHM_EINVAL, "fpohm_conn_fixed: null");
	DeviceGuard g(c->ctx->device);
	if (F_vs) c->F_vs.download(F_vs, 4 * c->nF);
	if (F_es) c->F_es.download(F_es, 4 * c->nF);
	if (F_boundary) c->F_boundary.download(F_boundary, c->nF);
	if (E_vs) c->E_vs.download(E_vs, 2 * c->nE);
	if (E_boundary) c->E_boundary.download(E_boundary, c->nE);
	if (V_boundary) c->V_boundary.download(V_boundary, c->nV);
	if (H_fs) c->H_fs.download(H_fs, 6 * c->H);
	FPOHM_CUDA(cudaStreamSynchronize(c->ctx->stream));
	FPOHM_API_END
}

int fpohm_conn_csr(const fpohm_conn *c, int32_t which, int64_t *off, uint32_t *val, int64_t *total) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(c && which >= 0 && which < 7, FPOHM_EINVAL, "fpohm_conn_csr: bad argument");
	FPOHM_REQUIRE(c->off[which].p, FPOHM_EINVAL, "fpohm_conn_csr: relation %d was not built for this handle", which);
	DeviceGuard g(c->ctx->device);
	const int64_t n = which == 0 ? c->nF : (which <= 2 ? c->nE : c->nV);
	if (total) *total = c->tot[which];
	if (off) c->off[which].download(off, n + 1);
	if (val) c->val[which].download(val, c->tot[which]);
	FPOHM_CUDA(cudaStreamSynchronize(c->ctx->stream));
	FPOHM_API_END
}

void fpohm_conn_free(fpohm_conn *c) {
	if (!c) return;
	DeviceGuard g(c->ctx->device);
	cudaStreamSynchronize(c->ctx->stream);
	delete c;
}

} // extern "C"

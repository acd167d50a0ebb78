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
	void excl_scan32(DevBuf<int32_t> &in, DevBuf<int32_t> &out, int64_t n) {
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in.p, out.p, n, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in.p, out.p, n, s));
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

// =====================================================================================================================
// extract_surface_conforming_mesh (global_functions.cpp:1021-1072): boundary faces of a hex mesh as a quad or triangle
// surface, build_connectivity's Tri/Qua branch (:19-56 + adjacency :231-248), orient_surface_mesh (:1073-1112).
//
// orient_surface_mesh walks a queue from face 0 and reverses every newly reached face that runs along the shared edge in
// the same direction as the face it was reached from; faces of other components keep their order; then all faces are
// reversed if the signed volume is positive.  On an orientable 2-manifold (what clean_hex_mesh leaves behind) the flip
// of a face does not depend on the path it was reached by, so a level-synchronous parallel breadth-first search gives the
// reference's result; every pair of neighbouring reached faces is checked afterwards and a contradiction (non-orientable
// or non-manifold component, where the reference's result is an accident of its queue order) is reported as an error.
namespace {

__global__ void surf_mark_kernel(const uint8_t *__restrict__ F_boundary, const uint32_t *__restrict__ F_vs, int64_t nF, int32_t *__restrict__ fb,
                                 int32_t *__restrict__ vtag)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < 4 * nF; t += (int64_t)gridDim.x * blockDim.x) {
		const bool b = F_boundary[t >> 2] != 0;
		if ((t & 3) == 0) fb[t >> 2] = b ? 1 : 0;
		if (b) vtag[F_vs[t]] = 1;
	}
}
__global__ void surf_vertices_kernel(const int32_t *__restrict__ vtag, const int32_t *__restrict__ vpos, int64_t nVh, const double *__restrict__ Vh,
                                     int32_t *__restrict__ V_map, int32_t *__restrict__ V_rev, double *__restrict__ V)
{
	for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nVh; v += (int64_t)gridDim.x * blockDim.x) {
		if (!vtag[v]) { V_map[v] = -1; continue; }
		const int64_t n = vpos[v];
		V_map[v] = (int32_t)n; V_rev[n] = (int32_t)v;
		for (int d = 0; d < 3; ++d) V[3 * n + d] = Vh[3 * v + d];
	}
}
// quads copied in face order; a triangle surface gets (0,1,2) and (2,3,0) of every quad (gf.cpp:1030-1054)
__global__ void surf_faces_kernel(const int32_t *__restrict__ fb, const int32_t *__restrict__ fpos, const uint32_t *__restrict__ F_vs_hex, int64_t nFh,
                                  const int32_t *__restrict__ V_map, int tri, int32_t *__restrict__ F_map, int32_t *__restrict__ F_rev,
                                  uint32_t *__restrict__ F_vs)
{
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nFh; f += (int64_t)gridDim.x * blockDim.x) {
		if (!fb[f]) { F_map[f] = -1; continue; }
		const int64_t r = fpos[f];
		uint32_t q[4];
		for (int k = 0; k < 4; ++k) q[k] = (uint32_t)V_map[F_vs_hex[4 * f + k]];
		if (tri) {
			F_map[f] = (int32_t)(2 * r); F_rev[2 * r] = (int32_t)f; F_rev[2 * r + 1] = (int32_t)f;
			uint32_t *o = F_vs + 6 * r;
			o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[2]; o[4] = q[3]; o[5] = q[0];
		} else {
			F_map[f] = (int32_t)r; F_rev[r] = (int32_t)f;
			for (int k = 0; k < 4; ++k) F_vs[4 * r + k] = q[k];
		}
	}
}
__global__ void surf_edge_keys_kernel(const uint32_t *__restrict__ F_vs, int64_t nF, int vn, uint64_t *__restrict__ key, uint32_t *__restrict__ id) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < vn * nF; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = t / vn; const int j = (int)(t % vn);
		uint32_t v0 = F_vs[vn * i + j], v1 = F_vs[vn * i + (j + 1) % vn];
		if (v0 > v1) { const uint32_t x = v0; v0 = v1; v1 = x; }
		key[t] = ((uint64_t)v0 << 32) | v1;
		id[t] = (uint32_t)t;
	}
}
// gf.cpp:33-55: edge ids, Es.vs, boundary = the run has one tuple, Fs.es; E.neighbor_fs = faces of the run (ascending)
__global__ void surf_edges_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ id, const int32_t *__restrict__ head,
                                  const int32_t *__restrict__ eid_incl, int64_t n, int vn, uint32_t *__restrict__ E_vs, uint8_t *__restrict__ E_boundary,
                                  uint8_t *__restrict__ V_boundary, uint32_t *__restrict__ F_es, int64_t *__restrict__ nfs_off, uint32_t *__restrict__ nfs_val,
                                  int64_t nE)
{
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t e = eid_incl[t] - 1;
		const uint32_t tid = id[t];
		F_es[tid] = (uint32_t)e;
		nfs_val[t] = tid / (uint32_t)vn;
		if (head[t]) {
			const uint32_t v0 = (uint32_t)(key[t] >> 32), v1 = (uint32_t)key[t];
			E_vs[2 * e] = v0; E_vs[2 * e + 1] = v1; nfs_off[e] = t;
			const bool b = t + 1 == n || head[t + 1];
			E_boundary[e] = b ? 1 : 0;
			if (b) { V_boundary[v0] = 1; V_boundary[v1] = 1; }
		}
		if (t == n - 1) nfs_off[nE] = n;
	}
}
// does face f (in its ORIGINAL vertex order) run along edge (v0 < v1) from v0 to v1?
__device__ __forceinline__ bool runs_forward(const uint32_t *__restrict__ F_vs, int vn, int64_t f, uint32_t v0, uint32_t v1) {
	for (int k = 0; k < vn; ++k) if (F_vs[vn * f + k] == v0) return F_vs[vn * f + (k + 1) % vn] == v1;
	return false;
}
// one level of the walk from face 0.  cnt[3]: rotating frontier sizes (this level / next / the one after, zeroed here)
__global__ void surf_bfs_kernel(const int32_t *__restrict__ cur, int32_t *__restrict__ next, int32_t *__restrict__ cnt, int level, int vn,
                                const uint32_t *__restrict__ F_vs, const uint32_t *__restrict__ F_es, const uint32_t *__restrict__ E_vs,
                                const int64_t *__restrict__ nfs_off, const uint32_t *__restrict__ nfs_val, int32_t *__restrict__ visited,
                                uint8_t *__restrict__ flip)
{
	const int n = cnt[level % 3];
	if (blockIdx.x == 0 && threadIdx.x == 0) cnt[(level + 2) % 3] = 0;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
		const int64_t f = cur[t];
		const bool xf = flip[f] != 0;
		for (int j = 0; j < vn; ++j) {
			const uint32_t e = F_es[vn * f + j];
			const uint32_t v0 = E_vs[2 * e], v1 = E_vs[2 * e + 1];
			const bool a = runs_forward(F_vs, vn, f, v0, v1) != xf;
			for (int64_t q = nfs_off[e]; q < nfs_off[e + 1]; ++q) {
				const int64_t nf = nfs_val[q];
				if (nf == f || atomicCAS(&visited[nf], 0, 1) != 0) continue;
				flip[nf] = (a == runs_forward(F_vs, vn, nf, v0, v1)) ? 1 : 0;     // same direction along the edge -> reversed
				next[atomicAdd(&cnt[(level + 1) % 3], 1)] = (int32_t)nf;
			}
		}
	}
}
// neighbouring reached faces must run along their common edge in opposite directions
__global__ void surf_check_kernel(int64_t nE, int vn, const uint32_t *__restrict__ F_vs, const uint32_t *__restrict__ E_vs, const int64_t *__restrict__ nfs_off,
                                  const uint32_t *__restrict__ nfs_val, const int32_t *__restrict__ visited, const uint8_t *__restrict__ flip,
                                  int32_t *__restrict__ bad)
{
	for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nE; e += (int64_t)gridDim.x * blockDim.x) {
		const uint32_t v0 = E_vs[2 * e], v1 = E_vs[2 * e + 1];
		int fwd = 0, bwd = 0;
		for (int64_t q = nfs_off[e]; q < nfs_off[e + 1]; ++q) {
			const int64_t f = nfs_val[q];
			if (!visited[f]) continue;
			if (runs_forward(F_vs, vn, f, v0, v1) != (flip[f] != 0)) ++fwd; else ++bwd;
		}
		if (fwd > 1 || bwd > 1) *bad = 1;
	}
}
// signed-volume test of orient_surface_mesh (gf.cpp:1097-1108) on the faces as the walk left them; fixed-order partial sums
__global__ void __launch_bounds__(256)
surf_volume_kernel(int64_t nF, int vn, const uint32_t *__restrict__ F_vs, const uint8_t *__restrict__ flip, const double *__restrict__ V,
                   double *__restrict__ partial)
{
	__shared__ double sm[256];
	double acc = 0;
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		uint32_t vs[4];
		for (int k = 0; k < vn; ++k) vs[k] = F_vs[vn * f + (flip[f] ? vn - 1 - k : k)];
		double c[3] = {0, 0, 0};
		for (int k = 0; k < vn; ++k) for (int d = 0; d < 3; ++d) c[d] += V[3 * (int64_t)vs[k] + d];
		const double inv = 1.0 / vn;
		for (int d = 0; d < 3; ++d) c[d] *= inv;
		for (int j = 0; j < vn; ++j) {
			const double *x = V + 3 * (int64_t)vs[j], *y = V + 3 * (int64_t)vs[(j + 1) % vn];
			acc += -((x[0] * y[1] * c[2] + x[1] * y[2] * c[0] + x[2] * y[0] * c[1]) - (x[2] * y[1] * c[0] + x[1] * y[0] * c[2] + x[0] * y[2] * c[1]));
		}
	}
	sm[threadIdx.x] = acc;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
// final vertex / edge order of every face: reversed iff (walk flip) xor (global flip)
__global__ void surf_apply_kernel(int64_t nF, int vn, const uint8_t *__restrict__ flip, int global_flip, uint32_t *__restrict__ F_vs, uint32_t *__restrict__ F_es) {
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		if ((flip[f] != 0) == (global_flip != 0)) continue;
		uint32_t vs[4], es[4];
		for (int k = 0; k < vn; ++k) { vs[k] = F_vs[vn * f + k]; es[k] = F_es[vn * f + k]; }
		for (int k = 0; k < vn; ++k) {
			F_vs[vn * f + k] = vs[vn - 1 - k];
			F_es[vn * f + k] = es[(2 * vn - 2 - k) % vn];       // edge (vs'[k], vs'[k+1]) = old edge vn-2-k (mod vn)
		}
	}
}
__global__ void surf_vfs_items_kernel(const uint32_t *__restrict__ F_vs, int64_t n, int vn, uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) { key[t] = F_vs[t]; val[t] = (uint32_t)(t / vn); }
}

} // namespace

extern "C" {

int fpohm_extract_surface(fpohm_ctx *ctx, const fpohm_conn *conn, const double *V, int32_t as_triangles, fpohm_surface **out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && conn && V && out, FPOHM_EINVAL, "fpohm_extract_surface: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	const int blk = 256;
	Ops ops{ctx, s};
	const int64_t nVh = conn->nV, nFh = conn->nF;
	const int vn = as_triangles ? 3 : 4;
	fpohm_surface *r = new fpohm_surface;
	try {
		r->ctx = ctx; r->vn = vn; r->nV_hex = nVh; r->nF_hex = nFh;
		DevBuf<double> Vh(3 * nVh, s);
		Vh.upload(V, 3 * nVh);
		KernelTimer timer(ctx, s);
		// ---- boundary faces and their vertices, renumbered in ascending order (gf.cpp:1029-1066)
		DevBuf<int32_t> fb(nFh + 1, s), fpos(nFh + 1, s), vtag(nVh + 1, s), vpos(nVh + 1, s);
		fb.zero(); vtag.zero();
		surf_mark_kernel<<<grid_for(ctx, 4 * nFh, blk), blk, 0, s>>>(conn->F_boundary.p, conn->F_vs.p, nFh, fb.p, vtag.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ops.excl_scan32(fb, fpos, nFh + 1);
		ops.excl_scan32(vtag, vpos, nVh + 1);
		const int64_t nB = ops.last_i32(fpos, nFh + 1), nV = ops.last_i32(vpos, nVh + 1);
		const int64_t nF = as_triangles ? 2 * nB : nB;
		r->nV = nV; r->nF = nF;
		FPOHM_REQUIRE(nF > 0, FPOHM_EINVAL, "fpohm_extract_surface: the mesh has no boundary face");
		r->V_map.alloc(nVh, s); r->V_rev.alloc(nV, s); r->V.alloc(3 * nV, s);
		r->F_map.alloc(nFh, s); r->F_rev.alloc(nF, s); r->F_vs.alloc(vn * nF, s); r->F_es.alloc(vn * nF, s);
		surf_vertices_kernel<<<grid_for(ctx, nVh, blk), blk, 0, s>>>(vtag.p, vpos.p, nVh, Vh.p, r->V_map.p, r->V_rev.p, r->V.p);
		FPOHM_LAUNCH_CHECK(ctx);
		surf_faces_kernel<<<grid_for(ctx, nFh, blk), blk, 0, s>>>(fb.p, fpos.p, conn->F_vs.p, nFh, r->V_map.p, as_triangles ? 1 : 0, r->F_map.p, r->F_rev.p, r->F_vs.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- edges: stable sort of (vmin, vmax) over vn*f + j
		const int64_t n = vn * nF;
		DevBuf<uint64_t> ek(n, s);
		DevBuf<uint32_t> eid_t(n, s);
		surf_edge_keys_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(r->F_vs.p, nF, vn, ek.p, eid_t.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ops.sort_pairs(ek, eid_t, n, 0, 32 + bits_for(nV));
		DevBuf<int32_t> ehead(n, s), eid(n, s);
		heads_u64_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(ek.p, n, ehead.p);
		FPOHM_LAUNCH_CHECK(ctx);
		ops.incl_scan(ehead, eid, n);
		const int64_t nE = ops.last_i32(eid, n);
		r->nE = nE;
		r->E_vs.alloc(2 * nE, s); r->E_boundary.alloc(nE, s); r->V_boundary.alloc(nV, s);
		r->V_boundary.zero();
		r->off[0].alloc(nE + 1, s); r->val[0].alloc(n, s); r->tot[0] = n;
		surf_edges_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(ek.p, eid_t.p, ehead.p, eid.p, n, vn, r->E_vs.p, r->E_boundary.p, r->V_boundary.p, r->F_es.p,
			r->off[0].p, r->val[0].p, nE);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- orient_surface_mesh: walk from face 0, 16 levels per host look
		DevBuf<int32_t> fr_a(nF, s), fr_b(nF, s), visited(nF, s), cnt(3, s), bad(1, s);
		DevBuf<uint8_t> flip(nF, s);
		visited.zero(); flip.zero(); bad.zero();
		{
			const int32_t zero = 0, one = 1, init[3] = {1, 0, 0};
			FPOHM_CUDA(cudaMemcpyAsync(fr_a.p, &zero, 4, cudaMemcpyHostToDevice, s));
			FPOHM_CUDA(cudaMemcpyAsync(visited.p, &one, 4, cudaMemcpyHostToDevice, s));
			FPOHM_CUDA(cudaMemcpyAsync(cnt.p, init, 12, cudaMemcpyHostToDevice, s));
			FPOHM_CUDA(cudaStreamSynchronize(s));
		}
		int level = 0;
		const int bgrid = ctx->sm_count * 2;
		while (true) {
			for (int k = 0; k < 16; ++k, ++level) {
				surf_bfs_kernel<<<bgrid, 128, 0, s>>>((level & 1) ? fr_b.p : fr_a.p, (level & 1) ? fr_a.p : fr_b.p, cnt.p, level, vn, r->F_vs.p, r->F_es.p, r->E_vs.p,
					r->off[0].p, r->val[0].p, visited.p, flip.p);
				FPOHM_LAUNCH_CHECK(ctx);
			}
			int32_t hc[3];
			cnt.download(hc, 3);
			FPOHM_CUDA(cudaStreamSynchronize(s));
			if (hc[level % 3] == 0) break;
		}
		r->bfs_levels = level;
		surf_check_kernel<<<grid_for(ctx, nE, blk), blk, 0, s>>>(nE, vn, r->F_vs.p, r->E_vs.p, r->off[0].p, r->val[0].p, visited.p, flip.p, bad.p);
		FPOHM_LAUNCH_CHECK(ctx);
		const int vgrid = grid_for(ctx, nF, 256, 4);
		DevBuf<double> partial(vgrid, s);
		surf_volume_kernel<<<vgrid, 256, 0, s>>>(nF, vn, r->F_vs.p, flip.p, r->V.p, partial.p);
		FPOHM_LAUNCH_CHECK(ctx);
		std::vector<double> hp((size_t)vgrid);
		int32_t hbad = 0;
		partial.download(hp.data(), vgrid); bad.download(&hbad, 1);
		FPOHM_CUDA(cudaStreamSynchronize(s));
		FPOHM_REQUIRE(!hbad, FPOHM_EINVAL, "fpohm_extract_surface: the component of face 0 is not an orientable 2-manifold "
		              "(the reference's orientation would depend on its queue order)");
		double res = 0;
		for (double x : hp) res += x;
		surf_apply_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(nF, vn, flip.p, res > 0 ? 1 : 0, r->F_vs.p, r->F_es.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- V.neighbor_es / V.neighbor_vs / V.neighbor_fs (gf.cpp:231-248)
		{
			const int64_t n2 = 2 * nE;
			DevBuf<uint32_t> key(n2, s), ve(n2, s), vv(n2, s), key2(n2, s);
			ves_items_kernel<<<grid_for(ctx, n2, blk), blk, 0, s>>>(r->E_vs.p, nE, key.p, ve.p, vv.p);
			FPOHM_LAUNCH_CHECK(ctx);
			FPOHM_CUDA(cudaMemcpyAsync(key2.p, key.p, 4 * (size_t)n2, cudaMemcpyDeviceToDevice, s));
			ops.csr_from_items(key, ve, n2, nV, r->off[2], r->val[2]); r->tot[2] = n2;
			ops.csr_from_items(key2, vv, n2, nV, r->off[1], r->val[1]); r->tot[1] = n2;
			DevBuf<uint32_t> fk(n, s), fv(n, s);
			surf_vfs_items_kernel<<<grid_for(ctx, n, blk), blk, 0, s>>>(r->F_vs.p, n, vn, fk.p, fv.p);
			FPOHM_LAUNCH_CHECK(ctx);
			ops.csr_from_items(fk, fv, n, nV, r->off[3], r->val[3]); r->tot[3] = n;
		}
		timer.stop();
		FPOHM_CUDA(cudaStreamSynchronize(s));
	} catch (...) { delete r; throw; }
	*out = r;
	FPOHM_API_END
}

int fpohm_surface_sizes(const fpohm_surface *r, int64_t *nV, int64_t *nF, int64_t *nE, int32_t *vn, int64_t *bfs_levels) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(r, FPOHM_EINVAL, "fpohm_surface_sizes: null");
	if (nV) *nV = r->nV;
	if (nF) *nF = r->nF;
	if (nE) *nE = r->nE;
	if (vn) *vn = r->vn;
	if (bfs_levels) *bfs_levels = r->bfs_levels;
	FPOHM_API_END
}

int fpohm_surface_export(const fpohm_surface *r, double *V, uint32_t *F_vs, uint32_t *F_es, uint32_t *E_vs, uint8_t *E_boundary, uint8_t *V_boundary,
                         int32_t *V_map, int32_t *V_map_reverse, int32_t *F_map, int32_t *F_map_reverse)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(r, FPOHM_EINVAL, "fpohm_surface_export: null");
	DeviceGuard g(r->ctx->device);
	if (V) r->V.download(V, 3 * r->nV);
	if (F_vs) r->F_vs.download(F_vs, r->vn * r->nF);
	if (F_es) r->F_es.download(F_es, r->vn * r->nF);
	if (E_vs) r->E_vs.download(E_vs, 2 * r->nE);
	if (E_boundary) r->E_boundary.download(E_boundary, r->nE);
	if (V_boundary) r->V_boundary.download(V_boundary, r->nV);
	if (V_map) r->V_map.download(V_map, r->nV_hex);
	if (V_map_reverse) r->V_rev.download(V_map_reverse, r->nV);
	if (F_map) r->F_map.download(F_map, r->nF_hex);
	if (F_map_reverse) r->F_rev.download(F_map_reverse, r->nF);
	FPOHM_CUDA(cudaStreamSynchronize(r->ctx->stream));
	FPOHM_API_END
}

int fpohm_surface_csr(const fpohm_surface *r, int32_t which, int64_t *off, uint32_t *val, int64_t *total) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(r && which >= 0 && which < 4, FPOHM_EINVAL, "fpohm_surface_csr: bad argument");
	DeviceGuard g(r->ctx->device);
	const int64_t n = which == 0 ? r->nE : r->nV;
	if (total) *total = r->tot[which];
	if (off) r->off[which].download(off, n + 1);
	if (val) r->val[which].download(val, r->tot[which]);
	FPOHM_CUDA(cudaStreamSynchronize(r->ctx->stream));
	FPOHM_API_END
}

void fpohm_surface_free(fpohm_surface *r) {
	if (!r) return;
	DeviceGuard g(r->ctx->device);
	cudaStreamSynchronize(r->ctx->stream);
	delete r;
}

} // extern "C"

// The three normal sets of build_aabb_tree (ghm.cpp:4231-4248) on the device, bit-identical to igl's:
//   per_face_normals            igl/per_face_normals.cpp:13-36        FN  unit face normals (zero where the area is zero)
//   per_vertex_normals (ANGLE)  igl/per_vertex_normals.cpp:38-108     VN  sum over incident corners of angle * FN, normalised
//   per_edge_normals (UNIFORM)  igl/per_edge_normals.cpp:20-77        EN  sum of FN over the incident faces, NOT normalised
//   E / EMAP                    igl/all_edges.cpp:35-42, unique_simplices.cpp:16-33
// What makes the sums bit-exact is their ORDER: igl adds a vertex's corner terms in ascending face order (then corner), an
// edge's face normals in ascending (face, corner) order.  Corner / directed-edge records are therefore sorted by (vertex) and
// (edge key) with a STABLE radix sort from an input laid out in exactly that order, and every vertex / edge sums its run
// sequentially.  The one thing left on the host is acos: the internal angles (igl/internal_angles.cpp:64-87) go through the
// host libm, whose last bit the device's acos does not promise to match; the host computes the 3 nF angles on all its
// threads (tens of ms at 2 M faces) while the device sorts.
#include "mesh.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cmath>
#include <thread>
#include <vector>

using namespace fpohm;

namespace {

__device__ __forceinline__ double sqn_seq(double x, double y, double z) { return (x * x + y * y) + z * z; }   // dynamic Eigen rows reduce sequentially

__global__ void face_normals_kernel(const double *__restrict__ V, const int32_t *__restrict__ F, int64_t nF, double *__restrict__ FN) {
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
		const double *p0 = V + 3 * (int64_t)F[3 * f], *p1 = V + 3 * (int64_t)F[3 * f + 1], *p2 = V + 3 * (int64_t)F[3 * f + 2];
		const double a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
		const double b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
		double n0 = a1 * b2 - a2 * b1, n1 = a2 * b0 - a0 * b2, n2 = a0 * b1 - a1 * b0;
		const double r = sqrt(sqn_seq(n0, n1, n2));
		if (r == 0) { n0 = n1 = n2 = 0; } else { const double ir = 1.0 / r; n0 *= ir; n1 *= ir; n2 *= ir; }      // `N.row(i) /= r` multiplies by 1 / r
		FN[3 * f] = n0; FN[3 * f + 1] = n1; FN[3 * f + 2] = n2;
	}
}
// corner records in (face, corner) order: key = vertex, payload = 3 f + d
__global__ void corner_keys_kernel(const int32_t *__restrict__ F, int64_t nF, uint32_t *__restrict__ key, uint32_t *__restrict__ pay) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 3 * nF; i += (int64_t)gridDim.x * blockDim.x) { key[i] = (uint32_t)F[i]; pay[i] = (uint32_t)i; }
}
__global__ void seg_begin_kernel(const uint32_t *__restrict__ skey, int64_t n, int64_t n_seg, int32_t *__restrict__ begin) {
	// begin[v] = first sorted position with key >= v  (n_seg + 1 entries)
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
		const int64_t a = i == 0 ? -1 : (int64_t)skey[i - 1], b = i == n ? n_seg : (int64_t)skey[i];
		for (int64_t v = a + 1; v <= b; ++v) begin[v] = (int32_t)i;
	}
}
__global__ void vertex_normals_kernel(int64_t nV, const int32_t *__restrict__ begin, const uint32_t *__restrict__ spay, const double *__restrict__ W,
                                      const double *__restrict__ FN, double *__restrict__ VN)
{
	for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
		double x = 0, y = 0, z = 0;
		for (int32_t k = begin[v]; k < begin[v + 1]; ++k) {
			const uint32_t i = spay[k];
			const double w = W[i];
			const double *fn = FN + 3 * (int64_t)(i / 3);
			x += w * fn[0]; y += w * fn[1]; z += w * fn[2];
		}
		const double r = sqrt(sqn_seq(x, y, z));      // N.rowwise().normalize(): a true quotient by the row norm
		VN[3 * v] = x / r; VN[3 * v + 1] = y / r; VN[3 * v + 2] = z / r;
	}
}
// directed edge (f, c) = (F[f][(c+1)%3], F[f][(c+2)%3]); records in (face, corner) order, key = (min << vb) | max, payload = 3 f + c
__global__ void edge_keys_kernel(const int32_t *__restrict__ F, int64_t nF, int vb, unsigned long long *__restrict__ key, uint32_t *__restrict__ pay) {
	for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < 3 * nF; p += (int64_t)gridDim.x * blockDim.x) {
		const int64_t f = p / 3; const int c = (int)(p % 3);
		unsigned long long a = (uint32_t)F[3 * f + (c + 1) % 3], b = (uint32_t)F[3 * f + (c + 2) % 3];
		if (a > b) { const unsigned long long t = a; a = b; b = t; }
		key[p] = (a << vb) | b;
		pay[p] = (uint32_t)p;
	}
}
__global__ void edge_heads_kernel(const unsigned long long *__restrict__ skey, int64_t n, int32_t *__restrict__ head) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) head[i] = (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
}
__global__ void edge_emit_kernel(const unsigned long long *__restrict__ skey, const uint32_t *__restrict__ spay, const int32_t *__restrict__ head,
                                 const int32_t *__restrict__ eid_incl, int64_t n, int64_t nF, int vb, const double *__restrict__ FN,
                                 int32_t *__restrict__ E, int32_t *__restrict__ EMAP, double *__restrict__ EN)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const int32_t e = eid_incl[i] - 1;
		const uint32_t p = spay[i];
		EMAP[(int64_t)(p / 3) + (int64_t)(p % 3) * nF] = e;      // EMAP(f + c * m)
		if (!head[i]) continue;
		const unsigned long long k = skey[i];
		E[2 * (int64_t)e] = (int32_t)(k >> vb); E[2 * (int64_t)e + 1] = (int32_t)(k & ((1ull << vb) - 1));
		double x = 0, y = 0, z = 0;
		for (int64_t j = i; j < n && (j == i || !head[j]); ++j) {      // the run, in ascending (face, corner) order
			const double *fn = FN + 3 * (int64_t)(spay[j] / 3);
			x += fn[0]; y += fn[1]; z += fn[2];
		}
		EN[3 * (int64_t)e] = x; EN[3 * (int64_t)e + 1] = y; EN[3 * (int64_t)e + 2] = z;
	}
}

// internal angles on the host (libm acos), igl/internal_angles.cpp:64-87 + squared_edge_lengths.cpp:30-44; all host threads
void host_internal_angles(const double *V, const int32_t *F, int64_t nF, double *W) {
	unsigned T = std::thread::hardware_concurrency();
	if (T == 0) T = 1;
	if (T > 32) T = 32;
	if (nF < 20000) T = 1;
	auto sqn = [](double x, double y, double z) { return (x * x + y * y) + z * z; };
	auto work = [&](int64_t lo, int64_t hi) {
		for (int64_t f = lo; f < hi; ++f) {
			const double *p0 = V + 3 * (int64_t)F[3 * f], *p1 = V + 3 * (int64_t)F[3 * f + 1], *p2 = V + 3 * (int64_t)F[3 * f + 2];
			double L[3];
			L[0] = sqn(p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]);
			L[1] = sqn(p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]);
			L[2] = sqn(p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]);
			for (int d = 0; d < 3; ++d) {
				const double s1 = L[d], s2 = L[(d + 1) % 3], s3 = L[(d + 2) % 3];
				W[3 * f + d] = std::acos((s3 + s2 - s1) / (2. * std::sqrt(s3 * s2)));
			}
		}
	};
	if (T == 1) { work(0, nF); return; }
	std::vector<std::thread> th;
	for (unsigned t = 0; t < T; ++t) th.emplace_back(work, nF * t / T, nF * (t + 1) / T);
	for (auto &x : th) x.join();
}

} // namespace

namespace fpohm {

// fills m->FN / VN / EN / EMAP / dE on the device and m->nE; the host copies (fpohm_mesh_normals) are made on demand
void build_normals_device(fpohm_ctx *ctx, fpohm_mesh *m, cudaStream_t s) {
	const int64_t nF = m->nF, nV = m->nV, n3 = 3 * nF;
	const int blk = 256;
	FPOHM_REQUIRE(n3 < (1ll << 31), FPOHM_ERANGE, "normals: %lld corners", (long long)n3);
	// the angles start on the host threads right away
	std::vector<double> hW((size_t)n3);
	std::thread angles(host_internal_angles, m->hV.data(), m->hF.data(), nF, hW.data());
	try {
		m->FN.alloc(3 * nF, s); m->VN.alloc(3 * nV, s); m->EMAP.alloc(n3, s);
		face_normals_kernel<<<grid_for(ctx, nF, blk), blk, 0, s>>>(m->V.p, m->F.p, nF, m->FN.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- edges ----
		int vb = 1;
		while ((1ll << vb) < nV) ++vb;
		DevBuf<unsigned long long> ek(n3, s), sek(n3, s);
		DevBuf<uint32_t> ep(n3, s), sep(n3, s);
		edge_keys_kernel<<<grid_for(ctx, n3, blk), blk, 0, s>>>(m->F.p, nF, vb, ek.p, ep.p);
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tb = 0;
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ek.p, sek.p, ep.p, sep.p, (int)n3, 0, 2 * vb, s));
		DevBuf<uint8_t> tmp((int64_t)tb, s);
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, ek.p, sek.p, ep.p, sep.p, (int)n3, 0, 2 * vb, s));
		DevBuf<int32_t> head(n3, s), eid(n3, s);
		edge_heads_kernel<<<grid_for(ctx, n3, blk), blk, 0, s>>>(sek.p, n3, head.p);
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tb2 = 0;
		FPOHM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, head.p, eid.p, (int)n3, s));
		DevBuf<uint8_t> tmp2((int64_t)tb2, s);
		FPOHM_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, tb2, head.p, eid.p, (int)n3, s));
		ctx->launches += 2;
		int32_t nE = 0;
		FPOHM_CUDA(cudaMemcpyAsync(&nE, eid.p + (n3 - 1), 4, cudaMemcpyDeviceToHost, s));
		FPOHM_CUDA(cudaStreamSynchronize(s));
		m->nE = nE;
		m->EN.alloc(3 * (int64_t)nE, s); m->dE.alloc(2 * (int64_t)nE, s);
		edge_emit_kernel<<<grid_for(ctx, n3, blk), blk, 0, s>>>(sek.p, sep.p, head.p, eid.p, n3, nF, vb, m->FN.p, m->dE.p, m->EMAP.p, m->EN.p);
		FPOHM_LAUNCH_CHECK(ctx);
		// ---- vertices: corners grouped by vertex, in (face, corner) order inside a group ----
		DevBuf<uint32_t> ck(n3, s), sck(n3, s), cp(n3, s), scp(n3, s);
		corner_keys_kernel<<<grid_for(ctx, n3, blk), blk, 0, s>>>(m->F.p, nF, ck.p, cp.p);
		FPOHM_LAUNCH_CHECK(ctx);
		size_t tb3 = 0;
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb3, ck.p, sck.p, cp.p, scp.p, (int)n3, 0, vb, s));
		DevBuf<uint8_t> tmp3((int64_t)tb3, s);
		FPOHM_CUDA(cub::DeviceRadixSort::SortPairs(tmp3.p, tb3, ck.p, sck.p, cp.p, scp.p, (int)n3, 0, vb, s));
		ctx->launches += 1;
		DevBuf<int32_t> vbegin(nV + 1, s);
		seg_begin_kernel<<<grid_for(ctx, n3 + 1, blk), blk, 0, s>>>(sck.p, n3, nV, vbegin.p);
		FPOHM_LAUNCH_CHECK(ctx);
		angles.join();
		DevBuf<double> W(n3, s);
		W.upload(hW.data(), n3);
		vertex_normals_kernel<<<grid_for(ctx, nV, blk), blk, 0, s>>>(nV, vbegin.p, scp.p, W.p, m->FN.p, m->VN.p);
		FPOHM_LAUNCH_CHECK(ctx);
		FPOHM_CUDA(cudaStreamSynchronize(s));      // hW is a local
	} catch (...) { if (angles.joinable()) angles.join(); throw; }
	m->hnormals_valid = false;
}

void mesh_host_normals(fpohm_mesh *m) {
	if (m->hnormals_valid) return;
	m->hFN.resize(3 * (size_t)m->nF); m->hVN.resize(3 * (size_t)m->nV); m->hEN.resize(3 * (size_t)m->nE);
	m->hE.resize(2 * (size_t)m->nE); m->hEMAP.resize(3 * (size_t)m->nF);
	m->FN.download(m->hFN.data(), 3 * m->nF); m->VN.download(m->hVN.data(), 3 * m->nV); m->EN.download(m->hEN.data(), 3 * m->nE);
	m->dE.download(m->hE.data(), 2 * m->nE); m->EMAP.download(m->hEMAP.data(), 3 * m->nF);
	cudaStreamSynchronize(m->FN.s);
	m->hnormals_valid = true;
}

} // namespace fpohm

// fpohm_shim.hpp — C++ drop-in over the C-ABI (include/fpohm.h) with the REFERENCE'S OWN SIGNATURES.
//
// Include this header AFTER the reference headers it shadows (global_types.h, grid_meshing/octree.h is replaced,
// see INTEGRATION.md) and link libfpohm.so.  Everything is a template or inline so that the header itself needs
// neither Eigen nor geogram: the reference's types (Mesh, Mesh_Quality, Treestr, Eigen::MatrixXd, GEO::Mesh ...) are
// bound at instantiation time, through the members the reference code itself uses.
//
//   reference entry point (file:line)                                   shim (namespace fpohm_shim)
//   -----------------------------------------------------------------   ------------------------------------------
//   scaled_jacobian(Mesh&, Mesh_Quality&)          gf.cpp:2309           scaled_jacobian(hmi, mq)
//   points_inside_mesh(MatrixXd&, Mesh&, VectorXd&) gf.cpp:4024          points_inside_mesh(Ps, tmi, signed_dis)
//   build_aabb_tree(Mesh&, Treestr&, bool)         ghm.cpp:4231          build_aabb_tree(tmi, a_tree, is_tri)
//   igl::signed_distance_pseudonormal(P,V,F,tree,FN,VN,EN,EMAP,S,I,C,N)  signed_distance_pseudonormal(P, a_tree, S, I, C, N)
//        callers ghm.cpp:2273,2603,3771,4045,4074
//   build_connectivity(Mesh&)  [Hex]               gf.cpp:16,121-264     build_connectivity(hmi)
//   OctreeGrid                                     octree.h:62-270       class OctreeGrid (same public members)
//   conforming_mesh(mo, hybrid, octree, grid_size) ghm.cpp:568           conforming_mesh(mo, hybrid, octree, grid_size)
//   dual_conforming_mesh(mo, hybrid, hs, h_type)   ghm.cpp:697           conforming_and_dual_mesh(mo, hybrid, hs, h_type, octree, grid_size)
//   octree_mesh(GEO::Mesh&, Mesh&, OctreeGrid&, Vector3i&) ghm.cpp:460   octree_mesh(ctx, V, nV, F, nF, mo, octree, grid_size, ...)
//   compute_sign(M, aabb, VoxelGrid<T>&)           voxelization.h:220    compute_sign(mesh, voxels)
//   compute_octree(M, mo, aabb, ...)               voxelization.cpp:353  compute_octree(mesh, octree, ..., Vpos, hex, inside)
//   compute(mesh0, mesh1, diag, max, ave)          metro_hausdorff.cpp:358   compute(mesh0, mesh1, diag, max, ave)
//   hausdorff_ratio_check / compute(..., ratio, thr) metro_hausdorff.cpp:12  compute(mesh0, mesh1, ratio, thr)
//   hausdorff_dis(mesh0, mesh1, outlierVs, thr)    gf.cpp:3590           hausdorff_dis(mesh0, mesh1, outlierVs, thr)
//   reorder_hex_mesh(Mesh&)                        gf.cpp:2199           reorder_hex_mesh(hmi)
//   re_indexing_connectivity(hmi, H_flag, Ho, V_map, V_map_reverse, H_map, H_map_reverse) gf.cpp:664   same name and arguments
//   tagging_uneven_element(mi, H_flag)             ghm.cpp:1983          tagging_uneven_element(mi, H_flag)
//   clean_non_manifold_ve(mi, hmi_local, maps..., signed_dis, H_flag) ghm.cpp:2006   same name and arguments
//   drop_small_pieces(Mesh_Domain&)                ghm.cpp:2081          drop_small_pieces(md)
//   clean_hex_mesh(Mesh &tmi, Mesh_Domain &md)     ghm.cpp:1932          clean_hex_mesh(tmi, md)   (args.scaffold_type 1)
//   extract_surface_conforming_mesh(meshi, mesho, V_map, V_map_reverse, F_map, F_map_reverse) gf.cpp:1021   same name and arguments
//   compute_jacobians(SLIMData&, uv)               slim_m.cpp:84         compute_jacobians(s, uv)               (tet branch)
//   update_weights_and_closest_rotations(s, V, F, uv) slim_m.cpp:108     update_weights_and_closest_rotations(s, V, F, uv)
//   compute_energy_with_jacobians(s, V, F, Ji, uv, areas) slim_m.cpp:792 compute_energy_with_jacobians(s, V, F, Ji, uv, areas)
//   igl::flip_avoiding::compute_max_step_from_singularities(uv, F, d) igl/flip_avoiding_line_search.cpp:273   compute_max_step_from_singularities(uv, F, d)
//
// Error behaviour mirrors the reference: bool returns and a line on std::cout/cerr, never an exception out of a call the
// reference declares noexcept-in-practice; a missing GPU is fatal by design (no CPU fallback) and reported loudly.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <map>
#include <queue>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fpohm.h"

namespace fpohm_shim {

// ---------------------------------------------------------------------------------------------------------------------
inline void check(int rc, const char *what) {
	if (rc != FPOHM_OK) {
		std::cerr << "[fpohm] " << what << " failed (" << rc << "): " << fpohm_last_error() << std::endl;
		throw std::runtime_error(std::string(what) + ": " + fpohm_last_error());
	}
}

// one context per process and device (env FPOHM_DEVICE, default 0) — the reference is single threaded
inline fpohm_ctx *context() {
	static fpohm_ctx *ctx = nullptr;
	if (!ctx) {
		const char *d = std::getenv("FPOHM_DEVICE");
		check(fpohm_ctx_create(d ? std::atoi(d) : 0, &ctx), "fpohm_ctx_create");
	}
	return ctx;
}

// RAII device mesh built from the reference's Mesh (V is 3 x n column-major == xyz per vertex; Fs[i].vs = 3 ids)
struct DeviceMesh {
	fpohm_mesh *h = nullptr;
	std::vector<int32_t> F;
	template <class MeshT>
	explicit DeviceMesh(const MeshT &tmi) {
		F.resize(3 * tmi.Fs.size());
		for (size_t i = 0; i < tmi.Fs.size(); ++i)
			for (int k = 0; k < 3; ++k) F[3 * i + k] = (int32_t)tmi.Fs[i].vs[k];
		check(fpohm_mesh_upload(context(), tmi.V.data(), (int64_t)tmi.V.cols(), F.data(), (int64_t)tmi.Fs.size(), &h), "fpohm_mesh_upload");
	}
	DeviceMesh(const double *V, int64_t nV, const int32_t *Fp, int64_t nF) {
		check(fpohm_mesh_upload(context(), V, nV, Fp, nF, &h), "fpohm_mesh_upload");
	}
	// content-keyed: a surface seen before comes back with its trees built (points_inside_mesh is called with the same surface
	// over and over, and rebuilds its igl::AABB every time in the reference, gf.cpp:4038)
	struct Cached {};
	template <class MeshT>
	DeviceMesh(const MeshT &tmi, Cached) {
		F.resize(3 * tmi.Fs.size());
		for (size_t i = 0; i < tmi.Fs.size(); ++i)
			for (int k = 0; k < 3; ++k) F[3 * i + k] = (int32_t)tmi.Fs[i].vs[k];
		check(fpohm_mesh_upload_cached(context(), tmi.V.data(), (int64_t)tmi.V.cols(), F.data(), (int64_t)tmi.Fs.size(), &h), "fpohm_mesh_upload_cached");
	}
	DeviceMesh(const DeviceMesh &) = delete;
	DeviceMesh &operator=(const DeviceMesh &) = delete;
	~DeviceMesh() { fpohm_mesh_free(h); }
};

// ---------------------------------------------------------------------------------------------------------------------
// scaled_jacobian(Mesh &hmi, Mesh_Quality &mq), gf.cpp:2309-2358 (Hex branch; other element types are not on the hot path)
template <class MeshT, class MeshQualityT>
bool scaled_jacobian(MeshT &hmi, MeshQualityT &mq) {
	if (hmi.type != 5 /* Mesh_type::Hex, global_types.h:457-465 */) return false;
	const int64_t H = (int64_t)hmi.Hs.size();
	std::vector<uint32_t> hex(8 * (size_t)H);
	for (int64_t i = 0; i < H; ++i)
		for (int k = 0; k < 8; ++k) hex[8 * i + k] = hmi.Hs[i].vs[k];
	mq.V_Js.resize(8 * H); mq.H_Js.resize(H);
	double mad[3]; int64_t flipped = 0;
	check(fpohm_scaled_jacobian(context(), hmi.V.data(), (int64_t)hmi.V.cols(), hex.data(), H, mq.V_Js.data(), mq.H_Js.data(), mad, &flipped),
	      "fpohm_scaled_jacobian");
	mq.min_Jacobian = mad[0]; mq.ave_Jacobian = mad[1]; mq.deviation_Jacobian = mad[2];
	if (flipped > 0) std::cout << "flipped elements: " << flipped << std::endl;   // gf.cpp:2356
	return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// build_aabb_tree(Mesh &tmi, Treestr &a_tree, bool is_tri), ghm.cpp:4231-4248.  The device tree lives beside the Treestr.
inline std::map<const void *, DeviceMesh *> &tree_table() { static std::map<const void *, DeviceMesh *> t; return t; }

template <class MeshT, class TreestrT>
void build_aabb_tree(MeshT &tmi, TreestrT &a_tree, bool /*is_tri*/ = false) {
	auto &tab = tree_table();
	auto it = tab.find(&a_tree);
	if (it != tab.end()) { delete it->second; tab.erase(it); }
	DeviceMesh *dm = new DeviceMesh(tmi);
	tab[&a_tree] = dm;
	check(fpohm_mesh_build_query_tree(context(), dm->h), "fpohm_mesh_build_query_tree");
	const int64_t nV = (int64_t)tmi.V.cols(), nF = (int64_t)tmi.Fs.size();
	int64_t nE = 0;
	check(fpohm_mesh_num_edges(dm->h, &nE), "fpohm_mesh_num_edges");
	// Treestr fields (global_types.h:681-691); Eigen is column-major, the C-ABI row-major: fill through operator()
	a_tree.TriV = tmi.V.transpose();
	a_tree.TriF.resize(nF, 3);
	for (int64_t i = 0; i < nF; ++i) for (int k = 0; k < 3; ++k) a_tree.TriF(i, k) = dm->F[3 * i + k];
	std::vector<double> FN(3 * nF), VN(3 * nV), EN(3 * nE);
	std::vector<int32_t> E(2 * nE), EMAP(3 * nF);
	check(fpohm_mesh_normals(dm->h, FN.data(), VN.data(), EN.data(), E.data(), EMAP.data()), "fpohm_mesh_normals");
	a_tree.TriFN.resize(nF, 3); a_tree.TriVN.resize(nV, 3); a_tree.TriEN.resize(nE, 3); a_tree.TriE.resize(nE, 2); a_tree.TriEMAP.resize(3 * nF);
	for (int64_t i = 0; i < nF; ++i) for (int k = 0; k < 3; ++k) a_tree.TriFN(i, k) = FN[3 * i + k];
	for (int64_t i = 0; i < nV; ++i) for (int k = 0; k < 3; ++k) a_tree.TriVN(i, k) = VN[3 * i + k];
	for (int64_t i = 0; i < nE; ++i) { for (int k = 0; k < 3; ++k) a_tree.TriEN(i, k) = EN[3 * i + k]; a_tree.TriE(i, 0) = E[2 * i]; a_tree.TriE(i, 1) = E[2 * i + 1]; }
	for (int64_t i = 0; i < 3 * nF; ++i) a_tree.TriEMAP(i) = EMAP[i];
}

// igl::signed_distance_pseudonormal(P, V, F, tree, FN, VN, EN, EMAP, S, I, C, N) with the tree looked up from its Treestr
template <class MatP, class TreestrT, class VecS, class VecI, class MatC, class MatN>
void signed_distance_pseudonormal(const MatP &P, const TreestrT &a_tree, VecS &S, VecI &I, MatC &C, MatN &N) {
	auto it = tree_table().find(&a_tree);
	if (it == tree_table().end()) throw std::runtime_error("signed_distance_pseudonormal: build_aabb_tree was not called for this Treestr");
	const int64_t np = (int64_t)P.rows();
	std::vector<double> p(3 * np), s(np), c(3 * np), n(3 * np);
	std::vector<int32_t> idx(np);
	for (int64_t i = 0; i < np; ++i) for (int k = 0; k < 3; ++k) p[3 * i + k] = P(i, k);
	check(fpohm_signed_distance(context(), it->second->h, p.data(), np, s.data(), idx.data(), c.data(), n.data()), "fpohm_signed_distance");
	S.resize(np); I.resize(np); C.resize(np, 3); N.resize(np, 3);
	for (int64_t i = 0; i < np; ++i) { S(i) = s[i]; I(i) = idx[i]; for (int k = 0; k < 3; ++k) { C(i, k) = c[3 * i + k]; N(i, k) = n[3 * i + k]; } }
}

// points_inside_mesh(MatrixXd &Ps, Mesh &tmi, VectorXd &signed_dis), gf.cpp:4024-4048
template <class MatP, class MeshT, class VecS>
void points_inside_mesh(MatP &Ps, MeshT &tmi, VecS &signed_dis) {
	DeviceMesh dm(tmi, DeviceMesh::Cached{});
	const int64_t np = (int64_t)Ps.rows();
	std::vector<double> p(3 * np), s(np);
	for (int64_t i = 0; i < np; ++i) for (int k = 0; k < 3; ++k) p[3 * i + k] = Ps(i, k);
	check(fpohm_signed_distance(context(), dm.h, p.data(), np, s.data(), nullptr, nullptr, nullptr), "fpohm_signed_distance");
	signed_dis.resize(np);
	for (int64_t i = 0; i < np; ++i) signed_dis(i) = s[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// build_connectivity(Mesh &hmi), Hex branch gf.cpp:121-186 + adjacency 226-264.  HybridF / HybridE are the reference's
// Hybrid_F / Hybrid_E; they are template parameters only so this header does not need global_types.h.
template <class MeshT>
void build_connectivity(MeshT &hmi) {
	if (hmi.type != 5 /* Mesh_type::Hex */) throw std::runtime_error("fpohm_shim::build_connectivity: only the Hex branch is on the hot path");
	const int64_t H = (int64_t)hmi.Hs.size(), nV = (int64_t)hmi.Vs.size();
	std::vector<uint32_t> hex(8 * (size_t)H);
	for (int64_t i = 0; i < H; ++i) for (int k = 0; k < 8; ++k) hex[8 * i + k] = hmi.Hs[i].vs[k];
	fpohm_conn *c = nullptr;
	check(fpohm_hex_connectivity(context(), hex.data(), H, nV, &c), "fpohm_hex_connectivity");
	int64_t nF = 0, nE = 0;
	fpohm_conn_sizes(c, &nF, &nE);
	std::vector<uint32_t> F_vs(4 * nF), F_es(4 * nF), E_vs(2 * nE), H_fs(6 * H);
	std::vector<uint8_t> Fb(nF), Eb(nE), Vb(nV);
	check(fpohm_conn_fixed(c, F_vs.data(), F_es.data(), Fb.data(), E_vs.data(), Eb.data(), Vb.data(), H_fs.data()), "fpohm_conn_fixed");
	hmi.Fs.clear(); hmi.Fs.resize(nF); hmi.Es.clear(); hmi.Es.resize(nE);
	for (int64_t f = 0; f < nF; ++f) {
		auto &x = hmi.Fs[f]; x.id = (uint32_t)f; x.boundary = Fb[f];
		x.vs.assign(F_vs.begin() + 4 * f, F_vs.begin() + 4 * f + 4); x.es.assign(F_es.begin() + 4 * f, F_es.begin() + 4 * f + 4);
	}
	for (int64_t e = 0; e < nE; ++e) { auto &x = hmi.Es[e]; x.id = (uint32_t)e; x.boundary = Eb[e]; x.vs = {E_vs[2 * e], E_vs[2 * e + 1]}; }
	for (int64_t v = 0; v < nV; ++v) hmi.Vs[v].boundary = Vb[v];
	for (int64_t h = 0; h < H; ++h) hmi.Hs[h].fs.assign(H_fs.begin() + 6 * h, H_fs.begin() + 6 * h + 6);
	auto fill = [&](int which, int64_t n, auto &&dst) {
		int64_t tot = 0;
		check(fpohm_conn_csr(c, which, nullptr, nullptr, &tot), "fpohm_conn_csr");
		std::vector<int64_t> off(n + 1); std::vector<uint32_t> val((size_t)tot);
		check(fpohm_conn_csr(c, which, off.data(), val.data(), &tot), "fpohm_conn_csr");
		for (int64_t i = 0; i < n; ++i) dst(i).assign(val.begin() + off[i], val.begin() + off[i + 1]);
	};
	fill(0, nF, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Fs[i].neighbor_hs; });
	fill(1, nE, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Es[i].neighbor_fs; });
	fill(2, nE, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Es[i].neighbor_hs; });
	fill(3, nV, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Vs[i].neighbor_vs; });
	fill(4, nV, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Vs[i].neighbor_es; });
	fill(5, nV, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Vs[i].neighbor_fs; });
	fill(6, nV, [&](int64_t i) -> std::vector<uint32_t> & { return hmi.Vs[i].neighbor_hs; });
	fpohm_conn_free(c);
}

// ---------------------------------------------------------------------------------------------------------------------
// clean_hex_mesh and its stages (ghm.cpp:1932-2124).  H_flag travels as one byte per hex.
template <class MeshT>
std::vector<uint32_t> hex_list(const MeshT &m) {
	std::vector<uint32_t> hex(8 * m.Hs.size());
	for (size_t i = 0; i < m.Hs.size(); ++i) for (int k = 0; k < 8; ++k) hex[8 * i + k] = m.Hs[i].vs[k];
	return hex;
}
inline std::vector<uint8_t> to_bytes(const std::vector<bool> &f) { std::vector<uint8_t> b(f.size()); for (size_t i = 0; i < f.size(); ++i) b[i] = f[i]; return b; }
inline void from_bytes(const std::vector<uint8_t> &b, std::vector<bool> &f) { f.resize(b.size()); for (size_t i = 0; i < b.size(); ++i) f[i] = b[i] != 0; }

// reorder_hex_mesh(Mesh &hmi), gf.cpp:2199-2229
template <class MeshT>
void reorder_hex_mesh(MeshT &hmi) {
	std::vector<uint32_t> hex = hex_list(hmi);
	check(fpohm_reorder_hexes(context(), hmi.V.data(), (int64_t)hmi.V.cols(), hex.data(), (int64_t)hmi.Hs.size(), nullptr), "fpohm_reorder_hexes");
	for (size_t i = 0; i < hmi.Hs.size(); ++i) hmi.Hs[i].vs.assign(hex.begin() + 8 * i, hex.begin() + 8 * i + 8);
}

// tagging_uneven_element(const Mesh &mi, vector<bool> &H_flag), ghm.cpp:1983-2005 (mi's connectivity is rebuilt on the device)
template <class MeshT>
void tagging_uneven_element(const MeshT &mi, std::vector<bool> &H_flag) {
	const std::vector<uint32_t> hex = hex_list(mi);
	fpohm_conn *c = nullptr;
	check(fpohm_hex_connectivity(context(), hex.data(), (int64_t)mi.Hs.size(), (int64_t)mi.Vs.size(), &c), "fpohm_hex_connectivity");
	std::vector<uint8_t> f = to_bytes(H_flag);
	const int rc = fpohm_tag_uneven_elements(context(), c, f.data(), nullptr);
	fpohm_conn_free(c);
	check(rc, "fpohm_tag_uneven_elements");
	from_bytes(f, H_flag);
}

// re_indexing_connectivity(Mesh &hmi, vector<bool> &H_flag, Mesh &Ho, V_map, V_map_reverse, H_map, H_map_reverse), gf.cpp:664-698
template <class MeshT>
void re_indexing_connectivity(MeshT &hmi, std::vector<bool> &H_flag, MeshT &Ho, std::vector<int32_t> &V_map, std::vector<int32_t> &V_map_reverse,
                              std::vector<int32_t> &H_map, std::vector<int32_t> &H_map_reverse)
{
	const int64_t H = (int64_t)hmi.Hs.size(), nV = (int64_t)hmi.Vs.size();
	const std::vector<uint32_t> hex = hex_list(hmi);
	const std::vector<uint8_t> f = to_bytes(H_flag);
	V_map.assign((size_t)nV, -1); V_map_reverse.assign((size_t)nV, 0); H_map.clear(); H_map_reverse.assign((size_t)H, 0);
	std::vector<uint32_t> sub(8 * (size_t)H);
	int64_t nv = 0, nh = 0;
	check(fpohm_reindex_submesh(context(), hex.data(), H, nV, f.data(), V_map.data(), V_map_reverse.data(), &nv, H_map_reverse.data(), &nh, sub.data()),
	      "fpohm_reindex_submesh");
	V_map_reverse.resize((size_t)nv); H_map_reverse.resize((size_t)nh);
	Ho = MeshT();
	Ho.type = hmi.type;
	Ho.Vs.resize((size_t)nv);
	Ho.V.resize(3, nv);
	for (int64_t j = 0; j < nv; ++j) {
		auto &v = Ho.Vs[(size_t)j];
		v.id = (uint32_t)j; v.v = hmi.Vs[(size_t)V_map_reverse[(size_t)j]].v;
		for (int d = 0; d < 3; ++d) Ho.V(d, j) = v.v[d];
	}
	Ho.Hs.resize((size_t)nh);
	for (int64_t h = 0; h < nh; ++h) { Ho.Hs[(size_t)h].id = (uint32_t)h; Ho.Hs[(size_t)h].vs.assign(sub.begin() + 8 * h, sub.begin() + 8 * h + 8); }
	if (nh > 0) build_connectivity(Ho);
}

// clean_non_manifold_ve(mi, hmi_local, V_map, V_map_reverse, H_map, H_map_reverse, signed_dis, H_flag), ghm.cpp:2006-2080
template <class MeshT, class VecS>
void clean_non_manifold_ve(MeshT &mi, MeshT &hmi_local, std::vector<int> &V_map, std::vector<int> &V_map_reverse, std::vector<int> &H_map,
                           std::vector<int> &H_map_reverse, VecS & /*signed_dis: unused by the reference too*/, std::vector<bool> &H_flag)
{
	const std::vector<uint32_t> hex = hex_list(mi);
	std::vector<uint8_t> f = to_bytes(H_flag);
	int32_t rounds = 0;
	check(fpohm_clean_non_manifold(context(), hex.data(), (int64_t)mi.Hs.size(), (int64_t)mi.Vs.size(), f.data(), &rounds), "fpohm_clean_non_manifold");
	from_bytes(f, H_flag);
	if (rounds > 0) re_indexing_connectivity(mi, H_flag, hmi_local, V_map, V_map_reverse, H_map, H_map_reverse);   // what the last round leaves behind
}

// drop_small_pieces(Mesh_Domain &md), ghm.cpp:2081-2124
template <class DomainT>
void drop_small_pieces(DomainT &md) {
	const std::vector<uint32_t> hex = hex_list(md.mesh_entire);
	std::vector<uint8_t> f = to_bytes(md.H_flag);
	int64_t pieces = 0;
	check(fpohm_drop_small_pieces(context(), hex.data(), (int64_t)md.mesh_entire.Hs.size(), (int64_t)md.mesh_entire.Vs.size(), f.data(), &pieces), "fpohm_drop_small_pieces");
	if (pieces > 1) {
		from_bytes(f, md.H_flag);
		re_indexing_connectivity(md.mesh_entire, md.H_flag, md.mesh_subA, md.V_map, md.V_map_reverse, md.H_map, md.H_map_reverse);
	}
}

// clean_hex_mesh(Mesh &tmi, Mesh_Domain &md), ghm.cpp:1932-1981 with args.scaffold_type == 1.  md.mesh_entire must carry its
// connectivity (Fs / Vs are where the medial flags go), as it does in the pipeline (ghm.cpp:205-209).
template <class MeshT, class DomainT>
void clean_hex_mesh(MeshT &tmi, DomainT &md) {
	auto &mi = md.mesh_entire;
	const int64_t H = (int64_t)mi.Hs.size(), nV = (int64_t)mi.Vs.size();
	std::vector<uint32_t> hex = hex_list(mi);
	DeviceMesh surface(tmi);
	fpohm_conn *c = nullptr;
	check(fpohm_hex_connectivity(context(), hex.data(), H, nV, &c), "fpohm_hex_connectivity");
	int64_t nF = 0;
	fpohm_conn_sizes(c, &nF, nullptr);
	std::vector<uint8_t> f((size_t)H), Fm((size_t)nF), Vm((size_t)nV);
	int64_t stats[6];
	const int rc = fpohm_clean_hex_mesh(context(), surface.h, mi.V.data(), nV, hex.data(), H, c, nullptr, f.data(), Fm.data(), Vm.data(), stats);
	fpohm_conn_free(c);
	check(rc, "fpohm_clean_hex_mesh");
	for (int64_t i = 0; i < H; ++i) mi.Hs[(size_t)i].vs.assign(hex.begin() + 8 * i, hex.begin() + 8 * i + 8);      // reorder_hex_mesh
	from_bytes(f, md.H_flag);
	re_indexing_connectivity(mi, md.H_flag, md.mesh_subA, md.V_map, md.V_map_reverse, md.H_map, md.H_map_reverse);
	if (!md.mesh_subA.Hs.size()) { std::cout << "no elements inside the object, exit"; return; }
	if ((int64_t)mi.Fs.size() == nF) for (int64_t i = 0; i < nF; ++i) if (Fm[(size_t)i]) mi.Fs[(size_t)i].on_medial_surface = true;
	for (int64_t i = 0; i < nV; ++i) if (Vm[(size_t)i]) mi.Vs[(size_t)i].on_medial_surface = true;
}

// extract_surface_conforming_mesh(Mesh &meshi, Mesh &mesho, V_map, V_map_reverse, F_map, F_map_reverse), gf.cpp:1021-1072 (with
// orient_surface_mesh :1073-1112 and the two build_connectivity calls).  mesho.type (Tri = 0 / Qua = 1 in Mesh_type) selects the
// surface kind, as in the reference; meshi is a hex mesh (its connectivity is rebuilt on the device).
template <class MeshT>
void extract_surface_conforming_mesh(MeshT &meshi, MeshT &mesho, std::vector<int32_t> &V_map, std::vector<int32_t> &V_map_reverse,
                                     std::vector<int32_t> &F_map, std::vector<int32_t> &F_map_reverse)
{
	const int64_t H = (int64_t)meshi.Hs.size(), nVh = (int64_t)meshi.Vs.size();
	const std::vector<uint32_t> hex = hex_list(meshi);
	fpohm_conn *c = nullptr;
	check(fpohm_hex_connectivity(context(), hex.data(), H, nVh, &c), "fpohm_hex_connectivity");
	int64_t nFh = 0;
	fpohm_conn_sizes(c, &nFh, nullptr);
	fpohm_surface *sf = nullptr;
	const bool tri = (int)mesho.type == 0;              // Mesh_type::Tri
	const int rc = fpohm_extract_surface(context(), c, meshi.V.data(), tri ? 1 : 0, &sf);
	fpohm_conn_free(c);
	check(rc, "fpohm_extract_surface");
	int64_t nV = 0, nF = 0, nE = 0; int32_t vn = 4;
	fpohm_surface_sizes(sf, &nV, &nF, &nE, &vn, nullptr);
	std::vector<double> V(3 * (size_t)nV);
	std::vector<uint32_t> F_vs((size_t)(vn * nF)), F_es((size_t)(vn * nF)), E_vs(2 * (size_t)nE);
	std::vector<uint8_t> Eb((size_t)nE), Vb((size_t)nV);
	V_map.assign((size_t)nVh, -1); V_map_reverse.assign((size_t)nV, 0); F_map.assign((size_t)nFh, -1); F_map_reverse.assign((size_t)nF, 0);
	check(fpohm_surface_export(sf, V.data(), F_vs.data(), F_es.data(), E_vs.data(), Eb.data(), Vb.data(), V_map.data(), V_map_reverse.data(),
	                           F_map.data(), F_map_reverse.data()), "fpohm_surface_export");
	mesho.Vs.clear(); mesho.Es.clear(); mesho.Fs.clear(); mesho.Hs.clear();
	mesho.Vs.resize((size_t)nV); mesho.Fs.resize((size_t)nF); mesho.Es.resize((size_t)nE);
	mesho.V.resize(3, nV);
	for (int64_t v = 0; v < nV; ++v) { mesho.Vs[(size_t)v].id = (uint32_t)v; mesho.Vs[(size_t)v].boundary = Vb[(size_t)v]; for (int d = 0; d < 3; ++d) mesho.V(d, v) = V[3 * v + d]; }
	for (int64_t f = 0; f < nF; ++f) {
		auto &x = mesho.Fs[(size_t)f]; x.id = (uint32_t)f;
		x.vs.assign(F_vs.begin() + vn * f, F_vs.begin() + vn * (f + 1)); x.es.assign(F_es.begin() + vn * f, F_es.begin() + vn * (f + 1));
	}
	for (int64_t e = 0; e < nE; ++e) { auto &x = mesho.Es[(size_t)e]; x.id = (uint32_t)e; x.boundary = Eb[(size_t)e]; x.vs = {E_vs[2 * e], E_vs[2 * e + 1]}; }
	auto fill = [&](int which, int64_t n, auto &&dst) {
		int64_t tot = 0;
		check(fpohm_surface_csr(sf, which, nullptr, nullptr, &tot), "fpohm_surface_csr");
		std::vector<int64_t> off((size_t)n + 1); std::vector<uint32_t> val((size_t)tot);
		check(fpohm_surface_csr(sf, which, off.data(), val.data(), &tot), "fpohm_surface_csr");
		for (int64_t i = 0; i < n; ++i) dst(i).assign(val.begin() + off[(size_t)i], val.begin() + off[(size_t)i + 1]);
	};
	fill(0, nE, [&](int64_t i) -> std::vector<uint32_t> & { return mesho.Es[(size_t)i].neighbor_fs; });
	fill(1, nV, [&](int64_t i) -> std::vector<uint32_t> & { return mesho.Vs[(size_t)i].neighbor_vs; });
	fill(2, nV, [&](int64_t i) -> std::vector<uint32_t> & { return mesho.Vs[(size_t)i].neighbor_es; });
	fill(3, nV, [&](int64_t i) -> std::vector<uint32_t> & { return mesho.Vs[(size_t)i].neighbor_fs; });
	fpohm_surface_free(sf);
}

// ---------------------------------------------------------------------------------------------------------------------
// SLIM per-element stages, tet branch (slim_m.cpp).  SLIMDataT is the reference's SLIMData: the members the
// reference functions touch (Dx, Dy, Dz, Ji, Ri, W_11..W_33, slim_energy, exp_factor, dim, f_n) are used the same way.
// compute_jacobians(SLIMData &s, const MatrixXd &uv), slim_m.cpp:84-106.  Dx, Dy, Dz share igl::grad's pattern (4 entries per tet).
template <class SLIMDataT, class MatT>
void compute_jacobians(SLIMDataT &s, const MatT &uv) {
	if (s.F.cols() == 3) throw std::runtime_error("fpohm_shim::compute_jacobians: only the tet branch is on the hot path");
	const int64_t n = (int64_t)s.Dx.rows(), nv = (int64_t)s.Dx.cols();
	// row-major CSR of the three operators from Eigen's column-major storage, entries of a row in ascending column
	std::vector<int64_t> off((size_t)n + 1, 0);
	for (int k = 0; k < s.Dx.outerSize(); ++k) for (typename decltype(s.Dx)::InnerIterator it(s.Dx, k); it; ++it) ++off[(size_t)it.row() + 1];
	for (int64_t i = 0; i < n; ++i) off[(size_t)i + 1] += off[(size_t)i];
	std::vector<int32_t> col((size_t)off[(size_t)n]);
	std::vector<double> vx(col.size()), vy(col.size()), vz(col.size());
	std::vector<int64_t> fill(off.begin(), off.end() - 1);
	for (int k = 0; k < s.Dx.outerSize(); ++k) for (typename decltype(s.Dx)::InnerIterator it(s.Dx, k); it; ++it) {
		const int64_t q = fill[(size_t)it.row()]++;
		col[(size_t)q] = (int32_t)it.col(); vx[(size_t)q] = it.value(); vy[(size_t)q] = s.Dy.coeff(it.row(), it.col()); vz[(size_t)q] = s.Dz.coeff(it.row(), it.col());
	}
	std::vector<double> u(3 * (size_t)nv), J(9 * (size_t)n);
	for (int64_t i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) u[3 * i + c] = uv(i, c);
	check(fpohm_slim_jacobians(context(), n, nv, off.data(), col.data(), vx.data(), vy.data(), vz.data(), u.data(), J.data()), "fpohm_slim_jacobians");
	for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) s.Ji(i, k) = J[9 * i + k];
}
// update_weights_and_closest_rotations(SLIMData &s, const MatrixXd &V, const MatrixXi &F, MatrixXd &uv), slim_m.cpp:108-381
template <class SLIMDataT, class MatV, class MatF, class MatT>
void update_weights_and_closest_rotations(SLIMDataT &s, const MatV &, const MatF &, MatT &uv) {
	compute_jacobians(s, uv);
	const int64_t n = (int64_t)s.Ji.rows();
	std::vector<double> J(9 * (size_t)n), W(9 * (size_t)n), R(9 * (size_t)n);
	for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) J[9 * i + k] = s.Ji(i, k);
	check(fpohm_slim_weights_rotations(context(), J.data(), n, (int32_t)s.slim_energy, s.exp_factor, W.data(), R.data()), "fpohm_slim_weights_rotations");
	for (int64_t i = 0; i < n; ++i) {
		s.W_11(i) = W[9 * i]; s.W_12(i) = W[9 * i + 1]; s.W_13(i) = W[9 * i + 2]; s.W_21(i) = W[9 * i + 3]; s.W_22(i) = W[9 * i + 4];
		s.W_23(i) = W[9 * i + 5]; s.W_31(i) = W[9 * i + 6]; s.W_32(i) = W[9 * i + 7]; s.W_33(i) = W[9 * i + 8];
		for (int k = 0; k < 9; ++k) s.Ri(i, k) = R[9 * i + k];
	}
}
// compute_energy_with_jacobians(SLIMData &s, V, F, const MatrixXd &Ji, MatrixXd &uv, VectorXd &areas), slim_m.cpp:792-916
template <class SLIMDataT, class MatV, class MatF, class MatJ, class MatT, class VecA>
double compute_energy_with_jacobians(SLIMDataT &s, const MatV &, const MatF &, const MatJ &Ji, MatT &, VecA &areas) {
	if (s.dim != 3) throw std::runtime_error("fpohm_shim::compute_energy_with_jacobians: only the tet branch is on the hot path");
	const int64_t n = (int64_t)s.f_n;
	std::vector<double> J(9 * (size_t)n), a((size_t)n);
	for (int64_t i = 0; i < n; ++i) { for (int k = 0; k < 9; ++k) J[9 * i + k] = Ji(i, k); a[(size_t)i] = areas(i); }
	double e = 0;
	check(fpohm_slim_energy(context(), J.data(), n, a.data(), (int32_t)s.slim_energy, s.exp_factor, &e), "fpohm_slim_energy");
	return e;
}

// igl::flip_avoiding::compute_max_step_from_singularities(const MatrixXd &uv, const MatrixXi &F, MatrixXd &d), tet branch
// (igl/flip_avoiding_line_search.cpp:273-299)
template <class MatU, class MatF, class MatD>
double compute_max_step_from_singularities(const MatU &uv, const MatF &F, MatD &d) {
	if (uv.cols() != 3 || F.cols() != 4) throw std::runtime_error("fpohm_shim::compute_max_step_from_singularities: only the tet branch is on the hot path");
	const int64_t nv = (int64_t)uv.rows(), n = (int64_t)F.rows();
	std::vector<double> u(3 * (size_t)nv), dd(3 * (size_t)nv);
	std::vector<int32_t> T(4 * (size_t)n);
	for (int64_t i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) { u[3 * i + c] = uv(i, c); dd[3 * i + c] = d(i, c); }
	for (int64_t i = 0; i < n; ++i) for (int c = 0; c < 4; ++c) T[4 * i + c] = (int32_t)F(i, c);
	double m = 0;
	check(fpohm_slim_max_step(context(), u.data(), nv, T.data(), n, dd.data(), nullptr, &m), "fpohm_slim_max_step");
	return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// grid_hex_meshing_bijective::conforming_mesh(Mesh &mo, Mesh &hybrid, OctreeGrid &octree, Vector3i &grid_size),
// ghm.cpp:568-696.  `octree` is anything with the reference's public node table (m_Nodes[i].position /
// .neighNodeId): the reference's own OctreeGrid or fpohm_shim::OctreeGrid.  `mo` is the octree hex mesh (vertex i = node i);
// its connectivity is rebuilt on the device, so build_connectivity(mo) need not have run.  Fills `hybrid` completely,
// including the adjacency lists build_connectivity(Hyb) leaves behind (gf.cpp:187-264).
template <class MeshT, class OctreeT, class Vec3iT>
void conforming_mesh(MeshT &mo, MeshT &hybrid, OctreeT &octree, Vec3iT &grid_size) {
	const int64_t nV = (int64_t)mo.Vs.size(), H = (int64_t)mo.Hs.size();
	std::vector<int32_t> npos(3 * (size_t)nV), nn(6 * (size_t)nV);
	for (int64_t i = 0; i < nV; ++i) {
		for (int d = 0; d < 3; ++d) npos[3 * i + d] = octree.m_Nodes[(size_t)i].position[d];
		for (int k = 0; k < 6; ++k) nn[6 * i + k] = octree.m_Nodes[(size_t)i].neighNodeId[k];
	}
	std::vector<uint32_t> hex(8 * (size_t)H);
	for (int64_t i = 0; i < H; ++i) for (int k = 0; k < 8; ++k) hex[8 * i + k] = mo.Hs[i].vs[k];
	const int32_t gs[3] = {(int32_t)grid_size[0], (int32_t)grid_size[1], (int32_t)grid_size[2]};
	fpohm_conn *c = nullptr; fpohm_hybrid *hy = nullptr;
	check(fpohm_hex_connectivity(context(), hex.data(), H, nV, &c), "fpohm_hex_connectivity");
	const int rc = fpohm_conforming_mesh_tables(context(), npos.data(), nn.data(), nV, gs, c, &hy);
	fpohm_conn_free(c);
	check(rc, "fpohm_conforming_mesh_tables");
	int64_t sz[8];
	fpohm_hybrid_sizes(hy, sz, nullptr);
	const int64_t nF = sz[1], nE = sz[3];
	std::vector<int64_t> F_off(nF + 1), H_foff(H + 1), H_voff(H + 1), F_nhoff(nF + 1);
	std::vector<uint32_t> F_vs(sz[4]), F_es(sz[4]), E_vs(2 * nE), H_fs(sz[5]), H_vs(sz[6]), F_nhs(sz[7]);
	std::vector<uint8_t> Fb(nF), Eb(nE), Vb(nV);
	const int rc2 = fpohm_hybrid_export(hy, F_off.data(), F_vs.data(), F_es.data(), Fb.data(), E_vs.data(), Eb.data(), Vb.data(), H_foff.data(),
	                                    H_fs.data(), H_voff.data(), H_vs.data(), F_nhoff.data(), F_nhs.data());
	fpohm_hybrid_free(hy);
	check(rc2, "fpohm_hybrid_export");
	hybrid.type = decltype(mo.type)(4);               // Mesh_type::Hyb (global_types.h:457-465)
	hybrid.V = mo.V;
	hybrid.Vs.clear(); hybrid.Vs.resize(nV); hybrid.Fs.clear(); hybrid.Fs.resize(nF); hybrid.Es.clear(); hybrid.Es.resize(nE); hybrid.Hs.clear(); hybrid.Hs.resize(H);
	for (int64_t v = 0; v < nV; ++v) { auto &x = hybrid.Vs[v]; x.id = (uint32_t)v; x.v = mo.Vs[v].v; x.boundary = Vb[v]; }
	for (int64_t f = 0; f < nF; ++f) {
		auto &x = hybrid.Fs[f]; x.id = (uint32_t)f; x.boundary = Fb[f];
		x.vs.assign(F_vs.begin() + F_off[f], F_vs.begin() + F_off[f + 1]); x.es.assign(F_es.begin() + F_off[f], F_es.begin() + F_off[f + 1]);
		x.neighbor_hs.assign(F_nhs.begin() + F_nhoff[f], F_nhs.begin() + F_nhoff[f + 1]);
	}
	for (int64_t e = 0; e < nE; ++e) { auto &x = hybrid.Es[e]; x.id = (uint32_t)e; x.boundary = Eb[e]; x.vs = {E_vs[2 * e], E_vs[2 * e + 1]}; }
	for (int64_t h = 0; h < H; ++h) {
		auto &x = hybrid.Hs[h]; x.id = (uint32_t)h;
		x.fs.assign(H_fs.begin() + H_foff[h], H_fs.begin() + H_foff[h + 1]); x.vs.assign(H_vs.begin() + H_voff[h], H_vs.begin() + H_voff[h + 1]);
	}
	// remaining adjacency lists, in the reference's visiting orders (gf.cpp:231-264): plain appends of the tables above
	for (int64_t f = 0; f < nF; ++f) {
		for (uint32_t e : hybrid.Fs[f].es) hybrid.Es[e].neighbor_fs.push_back((uint32_t)f);
		for (uint32_t v : hybrid.Fs[f].vs) hybrid.Vs[v].neighbor_fs.push_back((uint32_t)f);
	}
	for (int64_t e = 0; e < nE; ++e) {
		const uint32_t v0 = hybrid.Es[e].vs[0], v1 = hybrid.Es[e].vs[1];
		hybrid.Vs[v0].neighbor_es.push_back((uint32_t)e); hybrid.Vs[v1].neighbor_es.push_back((uint32_t)e);
		hybrid.Vs[v0].neighbor_vs.push_back(v1); hybrid.Vs[v1].neighbor_vs.push_back(v0);
		std::set<uint32_t> hs;
		for (uint32_t f : hybrid.Es[e].neighbor_fs) hs.insert(hybrid.Fs[f].neighbor_hs.begin(), hybrid.Fs[f].neighbor_hs.end());
		hybrid.Es[e].neighbor_hs.assign(hs.begin(), hs.end());
	}
	for (int64_t h = 0; h < H; ++h) for (uint32_t v : hybrid.Hs[h].vs) hybrid.Vs[v].neighbor_hs.push_back((uint32_t)h);
}

// helper: fill a reference Mesh of type Hyb from a device polyhedral mesh, adjacency lists included (gf.cpp:226-264 orders)
template <class MeshT>
void fill_hybrid_mesh(fpohm_hybrid *hy, MeshT &m) {
	int64_t sz[8];
	fpohm_hybrid_sizes(hy, sz, nullptr);
	const int64_t nV = sz[0], nF = sz[1], H = sz[2], nE = sz[3];
	std::vector<int64_t> F_off(nF + 1), H_foff(H + 1), H_voff(H + 1), F_nhoff(nF + 1);
	std::vector<uint32_t> F_vs(sz[4]), F_es(sz[4]), E_vs(2 * nE), H_fs(sz[5]), H_vs(sz[6]), F_nhs(sz[7]);
	std::vector<uint8_t> Fb(nF), Eb(nE), Vb(nV);
	check(fpohm_hybrid_export(hy, F_off.data(), F_vs.data(), F_es.data(), Fb.data(), E_vs.data(), Eb.data(), Vb.data(), H_foff.data(), H_fs.data(),
	                          H_voff.data(), H_vs.data(), F_nhoff.data(), F_nhs.data()), "fpohm_hybrid_export");
	m.type = decltype(m.type)(4);                      // Mesh_type::Hyb (global_types.h:457-465)
	m.Vs.resize(nV); m.Fs.clear(); m.Fs.resize(nF); m.Es.clear(); m.Es.resize(nE); m.Hs.clear(); m.Hs.resize(H);
	for (int64_t v = 0; v < nV; ++v) {
		auto &x = m.Vs[v]; x.id = (uint32_t)v; x.boundary = Vb[v];
		x.neighbor_vs.clear(); x.neighbor_es.clear(); x.neighbor_fs.clear(); x.neighbor_hs.clear();
	}
	for (int64_t f = 0; f < nF; ++f) {
		auto &x = m.Fs[f]; x.id = (uint32_t)f; x.boundary = Fb[f];
		x.vs.assign(F_vs.begin() + F_off[f], F_vs.begin() + F_off[f + 1]); x.es.assign(F_es.begin() + F_off[f], F_es.begin() + F_off[f + 1]);
		x.neighbor_hs.assign(F_nhs.begin() + F_nhoff[f], F_nhs.begin() + F_nhoff[f + 1]);
	}
	for (int64_t e = 0; e < nE; ++e) { auto &x = m.Es[e]; x.id = (uint32_t)e; x.boundary = Eb[e]; x.vs = {E_vs[2 * e], E_vs[2 * e + 1]}; }
	for (int64_t h = 0; h < H; ++h) {
		auto &x = m.Hs[h]; x.id = (uint32_t)h;
		x.fs.assign(H_fs.begin() + H_foff[h], H_fs.begin() + H_foff[h + 1]); x.vs.assign(H_vs.begin() + H_voff[h], H_vs.begin() + H_voff[h + 1]);
	}
	for (int64_t f = 0; f < nF; ++f) {
		for (uint32_t e : m.Fs[f].es) m.Es[e].neighbor_fs.push_back((uint32_t)f);
		for (uint32_t v : m.Fs[f].vs) m.Vs[v].neighbor_fs.push_back((uint32_t)f);
	}
	for (int64_t e = 0; e < nE; ++e) {
		const uint32_t v0 = m.Es[e].vs[0], v1 = m.Es[e].vs[1];
		m.Vs[v0].neighbor_es.push_back((uint32_t)e); m.Vs[v1].neighbor_es.push_back((uint32_t)e);
		m.Vs[v0].neighbor_vs.push_back(v1); m.Vs[v1].neighbor_vs.push_back(v0);
		std::set<uint32_t> hs;
		for (uint32_t f : m.Es[e].neighbor_fs) hs.insert(m.Fs[f].neighbor_hs.begin(), m.Fs[f].neighbor_hs.end());
		m.Es[e].neighbor_hs.assign(hs.begin(), hs.end());
	}
}

// conforming_mesh + dual_conforming_mesh in one call (the reference always runs them back to back, ghm.cpp:416-424):
// dual_conforming_mesh(Mesh &mo, Mesh &hybrid, Mesh &hybrid_standard, vector<Element_Type> &h_type), ghm.cpp:697-872.
// Note: `hybrid_standard.Vs[v].neighbor_hs` is left empty: the reference fills it from the vertex sets BEFORE the census
// rewrites them (gf.cpp:260-263) and nothing downstream reads it (connectivity_modification works from Hs[].vs).
template <class MeshT, class OctreeT, class Vec3iT, class TypeVecT>
void conforming_and_dual_mesh(MeshT &mo, MeshT &hybrid, MeshT &hybrid_standard, TypeVecT &h_type, OctreeT &octree, Vec3iT &grid_size) {
	const int64_t nV = (int64_t)mo.Vs.size(), H = (int64_t)mo.Hs.size();
	std::vector<int32_t> npos(3 * (size_t)nV), nn(6 * (size_t)nV);
	for (int64_t i = 0; i < nV; ++i) {
		for (int d = 0; d < 3; ++d) npos[3 * i + d] = octree.m_Nodes[(size_t)i].position[d];
		for (int k = 0; k < 6; ++k) nn[6 * i + k] = octree.m_Nodes[(size_t)i].neighNodeId[k];
	}
	std::vector<uint32_t> hex(8 * (size_t)H);
	for (int64_t i = 0; i < H; ++i) for (int k = 0; k < 8; ++k) hex[8 * i + k] = mo.Hs[i].vs[k];
	std::vector<double> Vp(3 * (size_t)nV);
	for (int64_t i = 0; i < nV; ++i) for (int d = 0; d < 3; ++d) Vp[3 * i + d] = mo.Vs[i].v[d];
	const int32_t gs[3] = {(int32_t)grid_size[0], (int32_t)grid_size[1], (int32_t)grid_size[2]};
	fpohm_conn *c = nullptr; fpohm_hybrid *hy = nullptr, *du = nullptr;
	check(fpohm_hex_connectivity(context(), hex.data(), H, nV, &c), "fpohm_hex_connectivity");
	const int rc = fpohm_conforming_mesh_tables(context(), npos.data(), nn.data(), nV, gs, c, &hy);
	fpohm_conn_free(c);
	check(rc, "fpohm_conforming_mesh_tables");
	const int rc2 = fpohm_dual_conforming_mesh(context(), hy, Vp.data(), nV, hex.data(), H, &du);
	if (rc2 != FPOHM_OK) { fpohm_hybrid_free(hy); check(rc2, "fpohm_dual_conforming_mesh"); }
	hybrid.V = mo.V; hybrid.Vs.clear(); hybrid.Vs.resize(nV);
	for (int64_t v = 0; v < nV; ++v) hybrid.Vs[v].v = mo.Vs[v].v;
	fill_hybrid_mesh(hy, hybrid);
	for (int64_t h = 0; h < H; ++h) for (uint32_t v : hybrid.Hs[h].vs) hybrid.Vs[v].neighbor_hs.push_back((uint32_t)h);
	fpohm_hybrid_free(hy);
	int64_t sz[8];
	fpohm_hybrid_sizes(du, sz, nullptr);
	std::vector<double> Vd(3 * (size_t)sz[0]); std::vector<int32_t> ty((size_t)sz[2]); int64_t census[7];
	check(fpohm_hybrid_dual_extra(du, Vd.data(), ty.data(), census), "fpohm_hybrid_dual_extra");
	hybrid_standard.V.resize(3, sz[0]); hybrid_standard.Vs.clear(); hybrid_standard.Vs.resize((size_t)sz[0]);
	for (int64_t v = 0; v < sz[0]; ++v) {
		hybrid_standard.Vs[v].v = {Vd[3 * v], Vd[3 * v + 1], Vd[3 * v + 2]};
		for (int d = 0; d < 3; ++d) hybrid_standard.V(d, v) = Vd[3 * v + d];
	}
	fill_hybrid_mesh(du, hybrid_standard);
	fpohm_hybrid_free(du);
	h_type.resize((size_t)sz[2]);
	for (int64_t h = 0; h < sz[2]; ++h) h_type[(size_t)h] = (typename TypeVecT::value_type)ty[(size_t)h];
	std::cout << "total, tetN, slabN, pyramidN, prismN, pyramidcombineN, tetcombineN, hexN: " << sz[2] << " " << census[0] << " " << census[1] << " "
	          << census[2] << " " << census[3] << " " << census[4] << " " << census[5] << " " << census[6] << std::endl;   // ghm.cpp:871
}

// ---------------------------------------------------------------------------------------------------------------------
// OctreeGrid, octree.h:62-270.  Same public data (m_Nodes, m_Cells, ...) and accessors; construction happens on the GPU.
struct Node {                                        // octree.h:16-33
	std::array<int, 6> neighNodeId;
	std::array<int, 3> position;
	Node() { neighNodeId.fill(-1); position.fill(0); }
	int prev(int axis) const { return neighNodeId[2 * axis]; }
	int next(int axis) const { return neighNodeId[2 * axis + 1]; }
};
struct Cell {                                        // octree.h:38-61
	int firstChild;
	std::array<int, 8> cornerNodeId;
	std::array<int, 6> neighCellId;
	Cell() : firstChild(-1) { neighCellId.fill(-1); cornerNodeId.fill(-1); }
	int corner(int localId) const { return cornerNodeId[localId]; }
	int adj(int axis, int dir) const { return neighCellId[2 * axis + dir]; }
	int prev(int axis) const { return neighCellId[2 * axis]; }
	int next(int axis) const { return neighCellId[2 * axis + 1]; }
};

class OctreeGrid {
public:
	std::array<int, 3> m_NodeGridSize{{0, 0, 0}}, m_CellGridSize{{0, 0, 0}};
	int m_MaxDepth = 0, m_NumRootCells = 0;
	std::vector<Node> m_Nodes;
	std::vector<Cell> m_Cells;

	OctreeGrid() {}
	explicit OctreeGrid(std::array<int, 3> fineCellGridSize) { OctreeGrid_initialize(fineCellGridSize); }
	~OctreeGrid() { fpohm_octree_free(h_); }
	OctreeGrid(const OctreeGrid &) = delete;
	OctreeGrid &operator=(const OctreeGrid &) = delete;

	// octree.cpp:39-60 (+ createRootCells :64-117): an empty tree of root cells
	void OctreeGrid_initialize(std::array<int, 3> fineCellGridSize) {
		m_CellGridSize = fineCellGridSize;
		for (int d = 0; d < 3; ++d) m_NodeGridSize[d] = fineCellGridSize[d] + 1;
		fpohm_octree_free(h_); h_ = nullptr;
		const int32_t gs[3] = {fineCellGridSize[0], fineCellGridSize[1], fineCellGridSize[2]};
		check(fpohm_octree_build_from_marks(context(), gs, nullptr, 0, 1, 1, &h_), "fpohm_octree_build_from_marks");
		pull();
	}

	int numNodes() const { return (int)m_Nodes.size(); }
	int numCells() const { return (int)m_Cells.size(); }
	int maxDepth() const { return m_MaxDepth; }
	std::array<int, 3> nodePos(int nodeId) const { return m_Nodes[nodeId].position; }
	std::array<int, 3> cellCornerPos(int cellId, int k) const { return m_Nodes[m_Cells[cellId].corner(k)].position; }
	int cellCornerId(int cellId, int k) const { return m_Cells[cellId].corner(k); }
	int cellExtent(int cellId) const { return cellCornerPos(cellId, 1)[0] - cellCornerPos(cellId, 0)[0]; }
	bool cellIsLeaf(int cellId) const { return m_Cells[cellId].firstChild == -1; }
	bool is2to1Graded() const { int32_t f = 0; check(fpohm_octree_check(h_, &f), "fpohm_octree_check"); return f & 1; }
	bool isPaired() const { int32_t f = 0; check(fpohm_octree_check(h_, &f), "fpohm_octree_check"); return (f & 2) != 0; }

	// subdivide(predicate, graded, paired), octree.cpp:648-690, for an ARBITRARY host predicate: the BFS (which cells get
	// tested) runs here exactly as in the reference; the closure (splits forced by grading / pairing), numbering and all
	// link tables are computed on the GPU from the predicate-true set.
	void subdivide(std::function<bool(int, int, int, int)> predicate, bool graded = false, bool paired = false, int /*maxCells*/ = -1) {
		std::vector<int32_t> marks;
		std::queue<std::array<int, 4>> pending;
		for (int c = 0; c < numCells(); ++c) {
			const auto p = cellCornerPos(c, 0);
			if (cellIsLeaf(c)) pending.push({{p[0], p[1], p[2], cellExtent(c)}});
			else marks.insert(marks.end(), {p[0], p[1], p[2], cellExtent(c)});           // already split: stays split
		}
		while (!pending.empty()) {
			const auto c = pending.front(); pending.pop();
			if (!predicate(c[0], c[1], c[2], c[3])) continue;
			if (c[3] == 1) { std::cerr << "[OctreeGrid] Cannot subdivide cell of length 1." << std::endl; continue; }
			marks.insert(marks.end(), {c[0], c[1], c[2], c[3]});
			const int e = c[3] / 2;
			for (int k = 0; k < 8; ++k) pending.push({{c[0] + (k & 1) * e, c[1] + ((k >> 1) & 1) * e, c[2] + (k >> 2) * e, e}});
		}
		rebuild(marks, graded, paired);
	}
	// subdivide(predicate, tb_subdivided_cells, graded, paired), octree.cpp:691-729: listed leaves only, no recursion
	void subdivide(std::function<bool(int, int, int, int)> predicate, std::vector<int> &cells, bool graded = false, bool paired = false, int = -1) {
		std::vector<int32_t> marks;
		for (int c = 0; c < numCells(); ++c)
			if (!cellIsLeaf(c)) { const auto p = cellCornerPos(c, 0); marks.insert(marks.end(), {p[0], p[1], p[2], cellExtent(c)}); }
		for (int cid : cells) {
			if (!cellIsLeaf(cid)) continue;
			const auto p = cellCornerPos(cid, 0); const int e = cellExtent(cid);
			if (!predicate(p[0], p[1], p[2], e)) continue;
			if (e == 1) { std::cerr << "[OctreeGrid] Cannot subdivide cell of length 1." << std::endl; continue; }
			marks.insert(marks.end(), {p[0], p[1], p[2], e});
		}
		rebuild(marks, graded, paired);
	}

	// bbox-predicate build entirely on the device (what octree_mesh / compute_octree do); see octree_mesh below
	void build_bbox(const DeviceMesh &mesh, const fpohm_octree_params &prm) {
		fpohm_octree_free(h_); h_ = nullptr;
		for (int d = 0; d < 3; ++d) { m_CellGridSize[d] = prm.grid_size[d]; m_NodeGridSize[d] = prm.grid_size[d] + 1; }
		check(fpohm_octree_build(context(), mesh.h, &prm, &h_), "fpohm_octree_build");
		pull();
	}
	void subdivide_bbox(const DeviceMesh &mesh, int stop_extent) { check(fpohm_octree_subdivide(h_, mesh.h, stop_extent), "fpohm_octree_subdivide"); pull(); }
	void refine_bbox(const DeviceMesh &mesh, const std::vector<int> &cells, int stop_extent) {
		std::vector<int32_t> c(cells.begin(), cells.end());
		check(fpohm_octree_refine(h_, mesh.h, c.data(), (int64_t)c.size(), stop_extent), "fpohm_octree_refine");
		pull();
	}
	fpohm_octree *handle() const { return h_; }

private:
	fpohm_octree *h_ = nullptr;
	void rebuild(const std::vector<int32_t> &marks, bool graded, bool paired) {
		fpohm_octree_free(h_); h_ = nullptr;
		const int32_t gs[3] = {m_CellGridSize[0], m_CellGridSize[1], m_CellGridSize[2]};
		check(fpohm_octree_build_from_marks(context(), gs, marks.data(), (int64_t)marks.size() / 4, graded, paired, &h_), "fpohm_octree_build_from_marks");
		pull();
	}
	void pull() {
		int64_t nn = 0, nc = 0, nl = 0; int32_t nr = 0, md = 0;
		check(fpohm_octree_sizes(h_, &nn, &nc, &nl, &nr, &md), "fpohm_octree_sizes");
		m_NumRootCells = nr; m_MaxDepth = md;
		std::vector<int32_t> np(3 * nn), ng(6 * nn), fc(nc), cc(8 * nc), cn(6 * nc);
		check(fpohm_octree_export(h_, np.data(), ng.data(), fc.data(), cc.data(), cn.data()), "fpohm_octree_export");
		m_Nodes.assign((size_t)nn, Node()); m_Cells.assign((size_t)nc, Cell());
		for (int64_t i = 0; i < nn; ++i) { for (int k = 0; k < 3; ++k) m_Nodes[i].position[k] = np[3 * i + k]; for (int k = 0; k < 6; ++k) m_Nodes[i].neighNodeId[k] = ng[6 * i + k]; }
		for (int64_t i = 0; i < nc; ++i) { m_Cells[i].firstChild = fc[i]; for (int k = 0; k < 8; ++k) m_Cells[i].cornerNodeId[k] = cc[8 * i + k]; for (int k = 0; k < 6; ++k) m_Cells[i].neighCellId[k] = cn[6 * i + k]; }
	}
};

// octree_mesh(GEO::Mesh &mi, Mesh &mo, OctreeGrid &octree, Vector3i &grid_size), ghm.cpp:460-567, minus the class state it
// reads (num_voxels, STOP_EXTENT_MIN/MAX, tb_subdivided_cells, hex2Octree_map), which become arguments.
//   first call / re_Octree_Meshing: stop_extent = 2^STOP_EXTENT_MIN, fresh tree; later calls: tb_subdivided_cells (already
//   mapped through hex2Octree_map, ghm.cpp:519) or a global re-subdivide with 2^STOP_EXTENT_MAX.
template <class MeshT, class Vec3i>
bool octree_mesh(const DeviceMesh &mi, const double *V, int64_t nV, MeshT &mo, OctreeGrid &octree, Vec3i &grid_size, int num_voxels,
                 int stop_extent, bool fresh, std::vector<int> &tb_subdivided_cells, std::vector<uint32_t> &hex2Octree_map)
{
	fpohm_octree_params prm{};
	check(fpohm_octree_grid_setup(V, nV, num_voxels, &prm), "fpohm_octree_grid_setup");
	prm.stop_extent = stop_extent; prm.graded = 1; prm.paired = 1;
	for (int d = 0; d < 3; ++d) grid_size[d] = prm.grid_size[d];
	if (fresh || !octree.numNodes()) octree.build_bbox(mi, prm);
	else if (!tb_subdivided_cells.empty()) { octree.refine_bbox(mi, tb_subdivided_cells, stop_extent); tb_subdivided_cells.clear(); }
	else octree.subdivide_bbox(mi, stop_extent);
	int64_t nn = 0, nc = 0, nl = 0;
	fpohm_octree_sizes(octree.handle(), &nn, &nc, &nl, nullptr, nullptr);
	std::vector<double> Vp(3 * nn); std::vector<uint32_t> hex(8 * nl); std::vector<int32_t> h2c(nl);
	check(fpohm_octree_hexes(octree.handle(), Vp.data(), hex.data(), h2c.data()), "fpohm_octree_hexes");
	hex2Octree_map.assign(h2c.begin(), h2c.end());
	mo.Vs.clear(); mo.Vs.resize(nn);
	mo.V.resize(3, nn);
	for (int64_t i = 0; i < nn; ++i) {
		mo.Vs[i].id = (uint32_t)i; mo.Vs[i].v = {Vp[3 * i], Vp[3 * i + 1], Vp[3 * i + 2]};
		for (int k = 0; k < 3; ++k) mo.V(k, i) = Vp[3 * i + k];
	}
	mo.Hs.clear(); mo.Hs.resize(nl);
	for (int64_t c = 0; c < nl; ++c) { mo.Hs[c].id = (uint32_t)c; mo.Hs[c].vs.assign(hex.begin() + 8 * c, hex.begin() + 8 * c + 8); }
	if (!mo.Hs.size()) { std::cout << "No octants, exit!" << std::endl; return false; }   // ghm.cpp:563
	build_connectivity(mo);
	return true;
}

// compute_octree(M, mo, aabb, filename, min_corner, extent, spacing, padding, graded, paired), voxelization.cpp:353-391:
// bbox-predicate octree down to extent 1 + z-ray parity per cell + hex export.  The GEO::Mesh the reference fills becomes
// three arrays: vertex positions (origin + nodePos * spacing, octree.cpp:744-747), hexes in GEOGRAM corner order
// (Cube::invDelta(diff[lv]), octree.cpp:759-767) and the "inside" cell attribute of the leaves (octree.cpp:838-849).
inline void compute_octree(const DeviceMesh &M, OctreeGrid &octree, const double min_corner[3], const double extent[3], double spacing,
                           int padding, bool graded, bool paired, std::vector<double> &Vpos, std::vector<uint32_t> &hex,
                           std::vector<float> &inside)
{
	auto next_pow2 = [](unsigned x) { x -= 1; x |= (x >> 1); x |= (x >> 2); x |= (x >> 4); x |= (x >> 8); x |= (x >> 16); return x + 1; };
	fpohm_octree_params prm{};
	for (int d = 0; d < 3; ++d) {
		prm.origin[d] = min_corner[d] - padding * spacing * 1.0;
		prm.grid_size[d] = (int32_t)next_pow2((unsigned)(std::ceil(extent[d] / spacing) + 2 * padding));
		prm.mesh_transform[d] = 0.0;
	}
	prm.voxel_size = spacing; prm.stop_extent = 1; prm.graded = graded; prm.paired = paired;
	octree.build_bbox(M, prm);
	int64_t nn = 0, nc = 0, nl = 0;
	fpohm_octree_sizes(octree.handle(), &nn, &nc, &nl, nullptr, nullptr);
	std::vector<float> cell_inside((size_t)nc);
	check(fpohm_octree_cell_sign(octree.handle(), M.h, prm.origin, spacing, cell_inside.data()), "fpohm_octree_cell_sign");
	std::vector<uint32_t> h8(8 * (size_t)nl); std::vector<int32_t> h2c((size_t)nl);
	Vpos.resize(3 * (size_t)nn);
	check(fpohm_octree_hexes(octree.handle(), Vpos.data(), h8.data(), h2c.data()), "fpohm_octree_hexes");
	static const int geo[8] = {0, 1, 3, 2, 4, 5, 7, 6};          // lv -> Cube::invDelta((lv&1, lv>>1&1, lv>>2&1))
	hex.resize(8 * (size_t)nl); inside.resize((size_t)nl);
	for (int64_t c = 0; c < nl; ++c) {
		for (int lv = 0; lv < 8; ++lv) hex[8 * c + lv] = h8[8 * c + geo[lv]];
		inside[c] = cell_inside[(size_t)h2c[c]];
	}
}

// ---------------------------------------------------------------------------------------------------------------------
// VoxelGrid<T> + compute_sign, voxelization.h:41-91,220-272
template <class T>
class VoxelGrid {
	std::vector<T> m_data;
	std::array<double, 3> m_origin;
	double m_spacing;
	std::array<int, 3> m_grid_size;
public:
	VoxelGrid(std::array<double, 3> origin, std::array<double, 3> extent, double voxel_size, int padding) : m_spacing(voxel_size) {
		int32_t dims[3]; double o[3];
		check(fpohm_voxel_grid_setup(origin.data(), extent.data(), voxel_size, padding, dims, o), "fpohm_voxel_grid_setup");
		for (int d = 0; d < 3; ++d) { m_grid_size[d] = dims[d]; m_origin[d] = o[d]; }
		m_data.assign((size_t)dims[0] * dims[1] * dims[2], T(0));
	}
	std::array<int, 3> grid_size() const { return m_grid_size; }
	int num_voxels() const { return m_grid_size[0] * m_grid_size[1] * m_grid_size[2]; }
	std::array<double, 3> origin() const { return m_origin; }
	double spacing() const { return m_spacing; }
	const T at(int idx) const { return m_data[idx]; }
	T &at(int idx) { return m_data[idx]; }
	const T *rawbuf() const { return m_data.data(); }
	T *raw_layer(int z) { return m_data.data() + (size_t)z * m_grid_size[1] * m_grid_size[0]; }
};
template <class T>
void compute_sign(const DeviceMesh &M, VoxelGrid<T> &voxels) {
	const auto gs = voxels.grid_size(); const auto o = voxels.origin();
	const int32_t dims[3] = {gs[0], gs[1], gs[2]};
	std::vector<uint8_t> out((size_t)voxels.num_voxels());
	check(fpohm_voxel_sign(context(), M.h, o.data(), voxels.spacing(), dims, out.data()), "fpohm_voxel_sign");
	for (size_t i = 0; i < out.size(); ++i) voxels.at((int)i) = T(out[i]);
}

// ---------------------------------------------------------------------------------------------------------------------
// metro: compute(const Mesh&, const Mesh&, double &bbox_diagonal, double &max, double &ave), metro_hausdorff.cpp:358-505
template <class MeshT>
void compute(const MeshT &mesh0, const MeshT &mesh1, double &bbox_diagonal, double &max_hausdorff_dis, double &ave_hausdorff_dis) {
	DeviceMesh a(mesh0), b(mesh1);
	double out[7]; int64_t ns[2];
	check(fpohm_hausdorff(context(), a.h, b.h, 0, out, ns), "fpohm_hausdorff");
	bbox_diagonal = out[0];
	max_hausdorff_dis = out[1] > out[2] ? out[1] : out[2];
	ave_hausdorff_dis = out[3] > out[4] ? out[3] : out[4];
}
// compute(Mesh&, Mesh&, double &hausdorff_ratio, double &hausdorff_ratio_threshold), metro_hausdorff.cpp:12-195
template <class MeshT>
int compute(MeshT &mesh0, MeshT &mesh1, double &hausdorff_ratio, double &hausdorff_ratio_threshold) {
	double diag, mx, ave;
	compute(static_cast<const MeshT &>(mesh0), static_cast<const MeshT &>(mesh1), diag, mx, ave);
	hausdorff_ratio = (float)mx / diag;                                     // metro_hausdorff.cpp:186
	std::printf("%f  wrt bounding box diagonal\n", hausdorff_ratio);
	return hausdorff_ratio > hausdorff_ratio_threshold ? false : true;
}
// hausdorff_dis(mesh0, mesh1, outlierVs, hausdorff_dis_threshold), gf.cpp:3590-3628
template <class MeshT>
bool hausdorff_dis(MeshT &mesh0, MeshT &mesh1, std::vector<int> &outlierVs, double &hausdorff_dis_threshold) {
	DeviceMesh a(mesh0), b(mesh1);
	std::vector<int32_t> out((size_t)mesh1.Vs.size()); int64_t n = 0;
	check(fpohm_hausdorff_outliers(context(), a.h, b.h, hausdorff_dis_threshold, out.data(), &n), "fpohm_hausdorff_outliers");
	outlierVs.assign(out.begin(), out.begin() + n);
	std::cout << "refered total: " << outlierVs.size() << " " << mesh1.Vs.size() << std::endl;
	return true;
}

// ---- wire formats: h_io::write_hybrid_mesh_MESH / _VTK (io.cpp:295-325, 101-181), read_feature_Graph_FGRAPH (io.cpp:412-434).
// Same files byte for byte; the rows are formatted by all host threads (fpohm_io_*).
template <class MeshT>
void write_hybrid_mesh_MESH(MeshT &hmi, const std::string &path) {
	const int type = (int)hmi.type;
	std::vector<uint32_t> el;
	int64_t n = 0;
	if (type == 0 || type == 2) { n = (int64_t)hmi.Fs.size(); el.reserve((size_t)(3 * n)); for (auto &f : hmi.Fs) for (int k = 0; k < 3; ++k) el.push_back(f.vs[(size_t)k]); }
	else if (type == 5) { n = (int64_t)hmi.Hs.size(); el.reserve((size_t)(8 * n)); for (auto &h : hmi.Hs) for (uint32_t v : h.vs) el.push_back(v); }
	check(fpohm_io_write_mesh(path.c_str(), hmi.V.data(), (int64_t)hmi.V.cols(), type, el.data(), n), "fpohm_io_write_mesh");
}
template <class MeshT>
void write_hybrid_mesh_VTK(MeshT &hmi, const std::string &path) {
	const int type = (int)hmi.type;
	std::vector<uint32_t> el;
	std::vector<int64_t> off;
	std::vector<uint8_t> vb(hmi.Vs.size());
	for (size_t i = 0; i < hmi.Vs.size(); ++i) vb[i] = hmi.Vs[i].boundary ? 1 : 0;
	int64_t n = 0; int arity = 0;
	if (type == 0 || type == 1) { n = (int64_t)hmi.Fs.size(); arity = n ? (int)hmi.Fs[0].vs.size() : 3; for (auto &f : hmi.Fs) for (int k = 0; k < arity; ++k) el.push_back(f.vs[(size_t)k]); }
	else if (type == 4) { n = (int64_t)hmi.Fs.size(); off.push_back(0); for (auto &f : hmi.Fs) { for (uint32_t v : f.vs) el.push_back(v); off.push_back((int64_t)el.size()); } }
	else { n = (int64_t)hmi.Hs.size(); arity = n ? (int)hmi.Hs[0].vs.size() : 8; for (auto &h : hmi.Hs) for (int k = 0; k < arity; ++k) el.push_back(h.vs[(size_t)k]); }
	check(fpohm_io_write_vtk(path.c_str(), hmi.V.data(), (int64_t)hmi.V.cols(), type, type == 4 ? off.data() : nullptr, el.data(), n, arity, vb.data(), (int64_t)vb.size()),
	      "fpohm_io_write_vtk");
}
template <class FeatureT>
bool read_feature_Graph_FGRAPH(FeatureT &mf, const std::string &path) {
	double ang = 0; int32_t oc = 0, ocs = 0; int64_t nc = 0, np = 0;
	if (fpohm_io_read_fgraph(path.c_str(), &ang, &oc, &ocs, nullptr, &nc, nullptr, &np) != FPOHM_OK) return false;      // io.cpp:414
	std::vector<int32_t> c((size_t)std::max<int64_t>(nc, 1)), p((size_t)std::max<int64_t>(2 * np, 2));
	if (fpohm_io_read_fgraph(path.c_str(), &ang, &oc, &ocs, c.data(), &nc, p.data(), &np) != FPOHM_OK) return false;
	mf.angle_threshold = ang; mf.orphan_curve = oc; mf.orphan_curve_single = ocs;
	mf.IN_corners.assign(c.begin(), c.begin() + nc);
	mf.IN_v_pairs.assign((size_t)np, {});
	for (int64_t i = 0; i < np; ++i) { mf.IN_v_pairs[(size_t)i].push_back(p[(size_t)(2 * i)]); mf.IN_v_pairs[(size_t)i].push_back(p[(size_t)(2 * i + 1)]); }
	return true;
}

} // namespace fpohm_shim

// C++ end-to-end parity of the drop-in shim against the reference's own functions, both linked into one executable:
//   reference side : libfpohm_ref.so  (unmodified octree.cpp / global_functions.cpp / metro_hausdorff.cpp ...)
//   product side   : fpohm_shim.hpp over libfpohm.so (CUDA)
// Built by `make -C oracle/ref shim_parity` (needs the reference headers, so only where /root/reference exists); the
// binary travels to the GPU box inside oracle/_ref/ and is run by tests/test_gpu_parity.py::test_cpp_shim_parity.
#include "grid_meshing/voxelization.h"
#include "global_types.h"
#include "global_functions.h"
#include "metro_hausdorff.h"
#include "grid_meshing/grid_hex_meshing.h"
#include "slim_m.h"
#include <igl/flip_avoiding_line_search.h>
#include <igl/signed_distance.h>
#include <geogram/basic/common.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/command_line.h>
#include <geogram/basic/command_line_args.h>

#include "fpohm_shim.hpp"

#include <array>
#include <algorithm>
#include <iostream>
#include <set>
#include <cmath>
#include <cstdio>

// the reference's SLIM per-element functions (slim_m.cpp, external linkage, not declared in slim_m.h)
void compute_jacobians(SLIMData &s, const Eigen::MatrixXd &uv);
void update_weights_and_closest_rotations(SLIMData &s, const Eigen::MatrixXd &V, const Eigen::MatrixXi &F, Eigen::MatrixXd &uv);
double compute_energy_with_jacobians(SLIMData &s, const Eigen::MatrixXd &V, const Eigen::MatrixXi &F, const Eigen::MatrixXd &Ji, Eigen::MatrixXd &uv, Eigen::VectorXd &areas);

static int failures = 0;
#define EXPECT(cond, what) do { if (cond) std::printf("PASS %s\n", what); else { std::printf("FAIL %s\n", what); ++failures; } } while (0)

static void torus(int nu, int nv, Mesh &m) {
	const double R = 1.0, r = 0.35, PI = 3.14159265358979323846;
	m.type = Mesh_type::Tri;
	m.V.resize(3, nu * nv); m.Vs.resize(nu * nv);
	for (int i = 0; i < nu; ++i) for (int j = 0; j < nv; ++j) {
		const double u = 2 * PI * i / nu, v = 2 * PI * j / nv;
		const int id = i * nv + j;
		m.V(0, id) = (R + r * std::cos(v)) * std::cos(u) / 2.7; m.V(1, id) = (R + r * std::cos(v)) * std::sin(u) / 2.7; m.V(2, id) = r * std::sin(v) / 2.7;
		m.Vs[id].id = id;
	}
	for (int i = 0; i < nu; ++i) for (int j = 0; j < nv; ++j) {
		const uint32_t a = i * nv + j, b = ((i + 1) % nu) * nv + j, c = ((i + 1) % nu) * nv + (j + 1) % nv, d = i * nv + (j + 1) % nv;
		Hybrid_F f0, f1; f0.vs = {a, b, c}; f1.vs = {a, c, d};
		f0.id = (uint32_t)m.Fs.size(); m.Fs.push_back(f0); f1.id = (uint32_t)m.Fs.size(); m.Fs.push_back(f1);
	}
}
static void hex_block(int n, double amp, Mesh &m) {
	m.type = Mesh_type::Hex;
	const int s = n + 1;
	m.V.resize(3, s * s * s); m.Vs.resize(s * s * s);
	for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) for (int k = 0; k < s; ++k) {
		const int id = (i * s + j) * s + k;
		m.V(0, id) = i / (double)n + amp / n * std::sin(7.0 * j + k); m.V(1, id) = j / (double)n + amp / n * std::cos(3.0 * i + 2 * k);
		m.V(2, id) = k / (double)n + amp / n * std::sin(5.0 * i * j + 1);
		m.Vs[id].id = id;
	}
	auto vid = [&](int i, int j, int k) { return (uint32_t)((i * s + j) * s + k); };
	for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) {
		Hybrid h; h.id = (uint32_t)m.Hs.size();
		h.vs = {vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k), vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)};
		m.Hs.push_back(h);
	}
}

int main() {
	GEO::initialize();
	GEO::CmdLine::import_arg_group("standard");
	GEO::Logger::instance()->set_quiet(true);

	// ---- scaled_jacobian + build_connectivity on a tangled hex block
	{
		Mesh a, b; hex_block(7, 1.4, a); hex_block(7, 1.4, b);
		Mesh_Quality qa, qb;
		::scaled_jacobian(a, qa);
		fpohm_shim::scaled_jacobian(b, qb);
		EXPECT(qa.V_Js == qb.V_Js && qa.H_Js == qb.H_Js && qa.min_Jacobian == qb.min_Jacobian, "scaled_jacobian bit-exact (V_Js, H_Js, min)");
		EXPECT(std::fabs(qa.ave_Jacobian - qb.ave_Jacobian) <= 1e-12 * std::fabs(qa.ave_Jacobian) && std::fabs(qa.deviation_Jacobian - qb.deviation_Jacobian) <= 1e-9 * qa.deviation_Jacobian,
		       "scaled_jacobian ave/deviation within 1e-9 (north star 1e-5)");
		::build_connectivity(a);
		fpohm_shim::build_connectivity(b);
		bool same = a.Fs.size() == b.Fs.size() && a.Es.size() == b.Es.size();
		for (size_t i = 0; same && i < a.Fs.size(); ++i) same = a.Fs[i].vs == b.Fs[i].vs && a.Fs[i].es == b.Fs[i].es && a.Fs[i].boundary == b.Fs[i].boundary && a.Fs[i].neighbor_hs == b.Fs[i].neighbor_hs;
		for (size_t i = 0; same && i < a.Es.size(); ++i) same = a.Es[i].vs == b.Es[i].vs && a.Es[i].boundary == b.Es[i].boundary && a.Es[i].neighbor_fs == b.Es[i].neighbor_fs && a.Es[i].neighbor_hs == b.Es[i].neighbor_hs;
		for (size_t i = 0; same && i < a.Vs.size(); ++i) same = a.Vs[i].boundary == b.Vs[i].boundary && a.Vs[i].neighbor_vs == b.Vs[i].neighbor_vs && a.Vs[i].neighbor_es == b.Vs[i].neighbor_es && a.Vs[i].neighbor_fs == b.Vs[i].neighbor_fs && a.Vs[i].neighbor_hs == b.Vs[i].neighbor_hs;
		for (size_t i = 0; same && i < a.Hs.size(); ++i) same = a.Hs[i].fs == b.Hs[i].fs;
		EXPECT(same, "build_connectivity identical (F/E/V/H relations)");
	}
	// ---- points_inside_mesh + Treestr + signed_distance_pseudonormal + metro
	{
		Mesh t; torus(48, 30, t);
		Eigen::MatrixXd Ps(4000, 3);
		for (int i = 0; i < 4000; ++i) { Ps(i, 0) = std::sin(i * 0.37) * 0.55; Ps(i, 1) = std::cos(i * 0.91) * 0.55; Ps(i, 2) = std::sin(i * 1.73) * 0.2; }
		Eigen::VectorXd sa, sb;
		::points_inside_mesh(Ps, t, sa);
		fpohm_shim::points_inside_mesh(Ps, t, sb);
		EXPECT(sa == sb, "points_inside_mesh bit-exact");
		Treestr tr;
		fpohm_shim::build_aabb_tree(t, tr);
		Eigen::MatrixXd V = t.V.transpose(), FN, VN, EN; Eigen::MatrixXi F = tr.TriF, E; Eigen::VectorXi EMAP;
		igl::per_face_normals(V, F, FN);
		igl::per_vertex_normals(V, F, igl::PER_VERTEX_NORMALS_WEIGHTING_TYPE_ANGLE, FN, VN);
		igl::per_edge_normals(V, F, igl::PER_EDGE_NORMALS_WEIGHTING_TYPE_UNIFORM, FN, EN, E, EMAP);
		EXPECT(FN == tr.TriFN && VN == tr.TriVN && EN == tr.TriEN && EMAP == tr.TriEMAP, "build_aabb_tree normals bit-exact (TriFN, TriVN, TriEN, TriEMAP)");
		igl::AABB<Eigen::MatrixXd, 3> tree; tree.init(V, F);
		Eigen::VectorXd S, S2; Eigen::VectorXi I, I2; Eigen::MatrixXd C, N, C2, N2;
		igl::signed_distance_pseudonormal(Ps, V, F, tree, FN, VN, EN, EMAP, S, I, C, N);
		fpohm_shim::signed_distance_pseudonormal(Ps, tr, S2, I2, C2, N2);
		EXPECT(S == S2 && I == I2 && C == C2 && N == N2, "signed_distance_pseudonormal bit-exact (S, I, C, N)");
		Mesh t2; torus(31, 19, t2); t2.V *= 1.01;
		double d0, m0, a0, d1, m1, a1;
		::compute((const Mesh &)t, (const Mesh &)t2, d0, m0, a0);
		fpohm_shim::compute((const Mesh &)t, (const Mesh &)t2, d1, m1, a1);
		EXPECT(d0 == d1 && std::fabs(m0 - m1) <= 1e-5 * m0 && std::fabs(a0 - a1) <= 1e-5 * a0, "metro compute within 1e-5");
	}
	// ---- OctreeGrid with an arbitrary host predicate (sphere shell), graded + paired
	{
		auto pred = [](int x, int y, int z, int e) {
			if (e <= 2) return false;
			const double cx = x + e * 0.5 - 32, cy = y + e * 0.5 - 32, cz = z + e * 0.5 - 16, r = std::sqrt(cx * cx + cy * cy + cz * cz);
			return std::fabs(r - 13.0) < e * 0.9;
		};
		::OctreeGrid ref(Eigen::Vector3i(64, 64, 32));
		ref.subdivide(pred, true, true);
		fpohm_shim::OctreeGrid mine(std::array<int, 3>{{64, 64, 32}});
		mine.subdivide(pred, true, true);
		bool same = ref.numCells() == mine.numCells() && ref.numNodes() == mine.numNodes();
		std::set<std::array<int, 4>> la, lb;
		for (int c = 0; same && c < ref.numCells(); ++c) if (ref.cellIsLeaf(c)) { auto p = ref.cellCornerPos(c, 0); la.insert({{p[0], p[1], p[2], ref.cellExtent(c)}}); }
		for (int c = 0; same && c < mine.numCells(); ++c) if (mine.cellIsLeaf(c)) { auto p = mine.cellCornerPos(c, 0); lb.insert({{p[0], p[1], p[2], mine.cellExtent(c)}}); }
		EXPECT(same && la == lb && mine.is2to1Graded() && mine.isPaired() && ref.is2to1Graded() && ref.isPaired(), "OctreeGrid::subdivide(std::function) same leaf set, graded, paired");
	}
	// ---- conforming_mesh (ghm.cpp:568-696) on the reference's own OctreeGrid: the class method vs the shim
	{
		auto pred = [](int x, int y, int z, int e) {
			if (e <= 2) return false;
			const double cx = x + e * 0.5 - 20, cy = y + e * 0.5 - 14, cz = z + e * 0.5 - 9, r = std::sqrt(cx * cx + cy * cy + cz * cz);
			return std::fabs(r - 9.0) < e * 0.8;
		};
		::OctreeGrid oct(Eigen::Vector3i(32, 32, 16));
		oct.subdivide(pred, true, true);
		auto octree_hex_mesh = [&](Mesh &m) {
			m.type = Mesh_type::Hex;
			m.Vs.resize(oct.numNodes()); m.V.resize(3, oct.numNodes());
			for (int i = 0; i < oct.numNodes(); ++i) {
				Hybrid_V v; v.id = i;
				for (int d = 0; d < 3; ++d) { v.v.push_back(oct.nodePos(i)[d]); m.V(d, i) = oct.nodePos(i)[d]; }
				m.Vs[i] = v;
			}
			for (int q = 0; q < oct.numCells(); ++q) if (oct.cellIsLeaf(q)) {
				Hybrid h; h.id = (uint32_t)m.Hs.size(); h.vs.resize(8);
				for (int lv = 0; lv < 8; ++lv) h.vs[lv] = oct.cellCornerId(q, lv);
				m.Hs.push_back(h);
			}
		};
		Mesh ma, mb, ha, hb;
		octree_hex_mesh(ma); octree_hex_mesh(mb);
		::build_connectivity(ma);
		Eigen::Vector3i gs(32, 32, 16);
		grid_hex_meshing_bijective gm;
		gm.conforming_mesh(ma, ha, oct, gs);
		fpohm_shim::conforming_mesh(mb, hb, oct, gs);
		bool same = ha.Fs.size() == hb.Fs.size() && ha.Es.size() == hb.Es.size() && ha.Hs.size() == hb.Hs.size() && ha.Vs.size() == hb.Vs.size() && ha.type == hb.type;
		size_t polygons = 0;
		for (size_t i = 0; same && i < ha.Fs.size(); ++i) { same = ha.Fs[i].vs == hb.Fs[i].vs && ha.Fs[i].es == hb.Fs[i].es && ha.Fs[i].boundary == hb.Fs[i].boundary && ha.Fs[i].neighbor_hs == hb.Fs[i].neighbor_hs; polygons += ha.Fs[i].vs.size() > 4; }
		for (size_t i = 0; same && i < ha.Es.size(); ++i) same = ha.Es[i].vs == hb.Es[i].vs && ha.Es[i].boundary == hb.Es[i].boundary && ha.Es[i].neighbor_fs == hb.Es[i].neighbor_fs && ha.Es[i].neighbor_hs == hb.Es[i].neighbor_hs;
		for (size_t i = 0; same && i < ha.Vs.size(); ++i) same = ha.Vs[i].boundary == hb.Vs[i].boundary && ha.Vs[i].neighbor_vs == hb.Vs[i].neighbor_vs && ha.Vs[i].neighbor_es == hb.Vs[i].neighbor_es && ha.Vs[i].neighbor_fs == hb.Vs[i].neighbor_fs && ha.Vs[i].neighbor_hs == hb.Vs[i].neighbor_hs && ha.Vs[i].v == hb.Vs[i].v;
		for (size_t i = 0; same && i < ha.Hs.size(); ++i) same = ha.Hs[i].fs == hb.Hs[i].fs && ha.Hs[i].vs == hb.Hs[i].vs;
		EXPECT(same && polygons > 0 && ha.Fs.size() < ma.Fs.size(), "conforming_mesh identical polyhedral mesh (loops, cells, edges, all adjacency lists)");
		// dual_conforming_mesh (ghm.cpp:697-872): dual mesh + element types, both halves through one shim call
		Mesh da, mc, hc, dc;
		std::vector<Element_Type> ta, tc;
		gm.dual_conforming_mesh(ma, ha, da, ta);
		octree_hex_mesh(mc);
		fpohm_shim::conforming_and_dual_mesh(mc, hc, dc, tc, oct, gs);
		bool sd = da.Fs.size() == dc.Fs.size() && da.Es.size() == dc.Es.size() && da.Hs.size() == dc.Hs.size() && da.Vs.size() == dc.Vs.size() && ta == tc && da.V == dc.V;
		for (size_t i = 0; sd && i < da.Fs.size(); ++i) sd = da.Fs[i].vs == dc.Fs[i].vs && da.Fs[i].es == dc.Fs[i].es && da.Fs[i].boundary == dc.Fs[i].boundary && da.Fs[i].neighbor_hs == dc.Fs[i].neighbor_hs;
		for (size_t i = 0; sd && i < da.Es.size(); ++i) sd = da.Es[i].vs == dc.Es[i].vs && da.Es[i].boundary == dc.Es[i].boundary && da.Es[i].neighbor_fs == dc.Es[i].neighbor_fs && da.Es[i].neighbor_hs == dc.Es[i].neighbor_hs;
		for (size_t i = 0; sd && i < da.Vs.size(); ++i) sd = da.Vs[i].boundary == dc.Vs[i].boundary && da.Vs[i].neighbor_vs == dc.Vs[i].neighbor_vs && da.Vs[i].neighbor_es == dc.Vs[i].neighbor_es && da.Vs[i].neighbor_fs == dc.Vs[i].neighbor_fs && da.Vs[i].v == dc.Vs[i].v;
		for (size_t i = 0; sd && i < da.Hs.size(); ++i) sd = da.Hs[i].fs == dc.Hs[i].fs && da.Hs[i].vs == dc.Hs[i].vs;
		std::set<int> kinds(ta.begin(), ta.end());
		EXPECT(sd && kinds.size() >= 3, "dual_conforming_mesh identical dual mesh, element types and template-ordered vertex lists");
	}
	// ---- compute_octree: the one public end-to-end entry of voxelization.h (bbox octree to extent 1 + ray parity + hex export)
	{
		Mesh t; torus(40, 24, t);
		GEO::Mesh M, mo;
		M.vertices.create_vertices((GEO::index_t)t.V.cols());
		for (int i = 0; i < t.V.cols(); ++i) M.vertices.point(i) = GEO::vec3(t.V(0, i), t.V(1, i), t.V(2, i));
		M.facets.create_triangles((GEO::index_t)t.Fs.size());
		for (size_t f = 0; f < t.Fs.size(); ++f) for (int c = 0; c < 3; ++c) M.facets.set_vertex((GEO::index_t)f, c, t.Fs[f].vs[c]);
		GEO::vec3 mn, mx; GEO::get_bbox(M, &mn[0], &mx[0]);
		const GEO::vec3 ext = mx - mn;
		const double spacing = 1.0 / 48;
		fpohm_shim::DeviceMesh dm(t);                       // before MeshFacetsAABB reorders M's facets (results do not depend on facet order)
		GEO::MeshFacetsAABB aabb(M);
		::compute_octree(M, mo, aabb, "", mn, ext, spacing, 1, true, true);
		fpohm_shim::OctreeGrid oc;
		std::vector<double> Vp; std::vector<uint32_t> hex; std::vector<float> inside;
		const double mnp[3] = {mn[0], mn[1], mn[2]}, exp_[3] = {ext[0], ext[1], ext[2]};
		fpohm_shim::compute_octree(dm, oc, mnp, exp_, spacing, 1, true, true, Vp, hex, inside);
		bool same = mo.cells.nb() == inside.size() && mo.vertices.nb() * 3 == Vp.size();
		// id-free comparison: multiset of (8 corner positions in geogram order, inside flag) per hex
		std::multiset<std::array<double, 25>> a, b;
		GEO::Attribute<float> rin(mo.cells.attributes(), "inside");
		for (GEO::index_t c = 0; same && c < mo.cells.nb(); ++c) {
			std::array<double, 25> ka, kb;
			for (int lv = 0; lv < 8; ++lv) for (int d = 0; d < 3; ++d) {
				ka[3 * lv + d] = mo.vertices.point(mo.cells.vertex(c, lv))[d];
				kb[3 * lv + d] = Vp[3 * (size_t)hex[8 * (size_t)c + lv] + d];
			}
			ka[24] = rin[c]; kb[24] = inside[c];
			a.insert(ka); b.insert(kb);
		}
		int n_in = 0; for (float v : inside) n_in += v > 0.5f;
		EXPECT(same && a == b && n_in > 0, "compute_octree identical hexes (positions bit-exact, geogram corner order) and inside flags");
	}
	// ---- clean_hex_mesh (ghm.cpp:1932-1981) and its stages on a lattice around the torus, a quarter of the hexes mirrored
	{
		Mesh t; torus(40, 24, t);
		auto lattice = [&](Mesh &m) {
			hex_block(20, 0.0, m);
			for (int i = 0; i < m.V.cols(); ++i) {
				for (int d = 0; d < 3; ++d) m.V(d, i) = (m.V(d, i) - 0.5) * (d == 2 ? 0.4 : 1.1);
				m.Vs[i].v = {m.V(0, i), m.V(1, i), m.V(2, i)};
			}
			for (size_t h = 0; h < m.Hs.size(); h += 4) { auto vs = m.Hs[h].vs; m.Hs[h].vs = {vs[3], vs[2], vs[1], vs[0], vs[7], vs[6], vs[5], vs[4]}; }
			::build_connectivity(m);
		};
		Mesh_Domain a, b;
		lattice(a.mesh_entire); lattice(b.mesh_entire);
		grid_hex_meshing_bijective gm;
		gm.clean_hex_mesh(t, a);
		fpohm_shim::clean_hex_mesh(t, b);
		bool same = a.H_flag == b.H_flag && a.V_map == b.V_map && a.V_map_reverse == b.V_map_reverse && a.H_map_reverse == b.H_map_reverse &&
		            a.mesh_subA.Hs.size() == b.mesh_subA.Hs.size() && a.mesh_subA.Fs.size() == b.mesh_subA.Fs.size() && a.mesh_subA.V == b.mesh_subA.V;
		size_t kept = 0, medial = 0;
		for (bool x : a.H_flag) kept += x;
		for (size_t i = 0; same && i < a.mesh_entire.Hs.size(); ++i) same = a.mesh_entire.Hs[i].vs == b.mesh_entire.Hs[i].vs;
		for (size_t i = 0; same && i < a.mesh_subA.Hs.size(); ++i) same = a.mesh_subA.Hs[i].vs == b.mesh_subA.Hs[i].vs && a.mesh_subA.Hs[i].fs == b.mesh_subA.Hs[i].fs;
		for (size_t i = 0; same && i < a.mesh_subA.Fs.size(); ++i) same = a.mesh_subA.Fs[i].vs == b.mesh_subA.Fs[i].vs && a.mesh_subA.Fs[i].boundary == b.mesh_subA.Fs[i].boundary;
		for (size_t i = 0; same && i < a.mesh_entire.Fs.size(); ++i) { same = a.mesh_entire.Fs[i].on_medial_surface == b.mesh_entire.Fs[i].on_medial_surface; medial += a.mesh_entire.Fs[i].on_medial_surface; }
		for (size_t i = 0; same && i < a.mesh_entire.Vs.size(); ++i) same = a.mesh_entire.Vs[i].on_medial_surface == b.mesh_entire.Vs[i].on_medial_surface;
		EXPECT(same && kept > 0 && medial > 0, "clean_hex_mesh identical H_flag, maps, sub-mesh and medial-surface flags");
		// stages on a carved flag set: every third hex dropped from what is inside, which leaves non-manifold vertices and edges
		std::vector<bool> fa = a.H_flag, fb;
		for (size_t i = 0; i < fa.size(); i += 3) fa[i] = false;
		fb = fa;
		gm.tagging_uneven_element(a.mesh_entire, fa);
		fpohm_shim::tagging_uneven_element(b.mesh_entire, fb);
		EXPECT(fa == fb, "tagging_uneven_element identical flags");
		a.H_flag = fa; b.H_flag = fb;
		::re_indexing_connectivity(a.mesh_entire, a.H_flag, a.mesh_subA, a.V_map, a.V_map_reverse, a.H_map, a.H_map_reverse);
		fpohm_shim::re_indexing_connectivity(b.mesh_entire, b.H_flag, b.mesh_subA, b.V_map, b.V_map_reverse, b.H_map, b.H_map_reverse);
		Eigen::VectorXd sd;
		gm.clean_non_manifold_ve(a.mesh_entire, a.mesh_subA, a.V_map, a.V_map_reverse, a.H_map, a.H_map_reverse, sd, a.H_flag);
		fpohm_shim::clean_non_manifold_ve(b.mesh_entire, b.mesh_subA, b.V_map, b.V_map_reverse, b.H_map, b.H_map_reverse, sd, b.H_flag);
		EXPECT(a.H_flag == b.H_flag && a.H_flag != fa && a.mesh_subA.Hs.size() == b.mesh_subA.Hs.size() && a.V_map == b.V_map, "clean_non_manifold_ve identical flags and re-indexed sub-mesh");
		gm.drop_small_pieces(a);
		fpohm_shim::drop_small_pieces(b);
		EXPECT(a.H_flag == b.H_flag && a.H_map_reverse == b.H_map_reverse && a.mesh_subA.Es.size() == b.mesh_subA.Es.size(), "drop_small_pieces identical flags and sub-mesh");
		// extract_surface_conforming_mesh (gf.cpp:1021-1112) of the cleaned sub-mesh, quad and triangle surfaces
		for (int kind = 0; kind < 2; ++kind) {
			Mesh sa, sb;
			sa.type = sb.type = kind ? Mesh_type::Tri : Mesh_type::Qua;
			std::vector<int32_t> vm_a, vr_a, fm_a, fr_a, vm_b, vr_b, fm_b, fr_b;
			::extract_surface_conforming_mesh(a.mesh_subA, sa, vm_a, vr_a, fm_a, fr_a);
			fpohm_shim::extract_surface_conforming_mesh(b.mesh_subA, sb, vm_b, vr_b, fm_b, fr_b);
			bool ss = vm_a == vm_b && vr_a == vr_b && fm_a == fm_b && fr_a == fr_b && sa.V == sb.V && sa.Fs.size() == sb.Fs.size() && sa.Es.size() == sb.Es.size() && sa.Vs.size() == sb.Vs.size();
			for (size_t i = 0; ss && i < sa.Fs.size(); ++i) ss = sa.Fs[i].vs == sb.Fs[i].vs && sa.Fs[i].es == sb.Fs[i].es;
			for (size_t i = 0; ss && i < sa.Es.size(); ++i) ss = sa.Es[i].vs == sb.Es[i].vs && sa.Es[i].boundary == sb.Es[i].boundary && sa.Es[i].neighbor_fs == sb.Es[i].neighbor_fs;
			for (size_t i = 0; ss && i < sa.Vs.size(); ++i) ss = sa.Vs[i].boundary == sb.Vs[i].boundary && sa.Vs[i].neighbor_vs == sb.Vs[i].neighbor_vs && sa.Vs[i].neighbor_es == sb.Vs[i].neighbor_es && sa.Vs[i].neighbor_fs == sb.Vs[i].neighbor_fs;
			EXPECT(ss && sa.Fs.size() > 0, kind ? "extract_surface_conforming_mesh identical triangle surface, maps and adjacency" : "extract_surface_conforming_mesh identical quad surface, maps and adjacency");
		}
	}
	// ---- the CHAIN: octree_mesh -> conforming_mesh -> dual_conforming_mesh, and octree_mesh -> clean_hex_mesh ->
	// extract_surface_conforming_mesh, every stage fed with the previous stage's output on its own side — the reference's
	// member functions on the reference's split-order numbering (ghm.cpp:460-567, 568-872, 1932-1981; gf.cpp:1021-1112), the
	// shim on the product's canonical (level, Morton) numbering.  Ids differ by construction; the comparison is by GEOMETRY:
	// does the canonical renumbering survive composition, including the id-order-dependent cleaning stages (tagging sweep,
	// greedy non-manifold clean-up, drop_small_pieces' tie-break)?  C1 torus at --e 15 and the refinement to --e 13, and a
	// second subdivide pass on the same octree objects (ghm.cpp:495-500).
	for (int E : {15, 13}) {
		Mesh t; torus(100, 100, t);                               // BASELINE config C1: 20 000 triangles
		GEO::Mesh mi;
		mi.vertices.create_vertices((GEO::index_t)t.V.cols());
		for (int i = 0; i < t.V.cols(); ++i) mi.vertices.point(i) = GEO::vec3(t.V(0, i), t.V(1, i), t.V(2, i));
		mi.facets.create_triangles((GEO::index_t)t.Fs.size());
		for (size_t f = 0; f < t.Fs.size(); ++f) for (int c = 0; c < 3; ++c) mi.facets.set_vertex((GEO::index_t)f, c, t.Fs[f].vs[c]);
		fpohm_shim::DeviceMesh dm(t);                             // before MeshFacetsAABB reorders mi's facets
		std::vector<double> Vt((size_t)(3 * t.V.cols()));
		for (int i = 0; i < t.V.cols(); ++i) for (int d = 0; d < 3; ++d) Vt[(size_t)(3 * i + d)] = t.V(d, i);
		// reference side
		grid_hex_meshing_bijective gm;
		gm.num_voxels = 1 << 20; gm.STOP_EXTENT_MIN = E; gm.STOP_EXTENT_MAX = E; gm.graded = true; gm.paired = true;
		Mesh_Domain a, b;
		a.mesh_entire.type = b.mesh_entire.type = Mesh_type::Hex;
		::OctreeGrid oa;
		Eigen::Vector3i gsa, gsb;
		std::streambuf *old = std::cout.rdbuf(nullptr);
		const bool oka = gm.octree_mesh(mi, a.mesh_entire, oa, gsa);
		// product side
		fpohm_shim::OctreeGrid ob;
		std::vector<int> tb; std::vector<uint32_t> h2o;
		const bool okb = fpohm_shim::octree_mesh(dm, Vt.data(), (int64_t)t.V.cols(), b.mesh_entire, ob, gsb, 1 << 20, 1 << E, true, tb, h2o);
		bool same = oka && okb && gsa == gsb && a.mesh_entire.Hs.size() == b.mesh_entire.Hs.size() && a.mesh_entire.Vs.size() == b.mesh_entire.Vs.size();
		auto hex_key = [](const Mesh &m, size_t h) {
			std::array<double, 6> k = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
			for (uint32_t v : m.Hs[h].vs) for (int d = 0; d < 3; ++d) { k[d] = std::min(k[d], m.V(d, v)); k[3 + d] = std::max(k[3 + d], m.V(d, v)); }
			return k;
		};
		// conforming + dual on each side's own octree and hex mesh
		Mesh ha, da, hb, db;
		std::vector<Element_Type> ta, tb2;
		gm.conforming_mesh(a.mesh_entire, ha, oa, gsa);
		gm.dual_conforming_mesh(a.mesh_entire, ha, da, ta);
		fpohm_shim::conforming_and_dual_mesh(b.mesh_entire, hb, db, tb2, ob, gsb);
		bool sd = same && ha.Fs.size() == hb.Fs.size() && ha.Es.size() == hb.Es.size() && da.Hs.size() == db.Hs.size() && da.Fs.size() == db.Fs.size() && da.Vs.size() == db.Vs.size();
		{   // dual cells by geometry: (sorted vertex positions, element type)
			auto cell_keys = [](const Mesh &m, const std::vector<Element_Type> &ty) {
				std::multiset<std::vector<double>> ks;
				for (size_t h = 0; h < m.Hs.size(); ++h) {
					std::vector<std::array<double, 3>> ps;
					for (uint32_t v : m.Hs[h].vs) ps.push_back({{m.V(0, v), m.V(1, v), m.V(2, v)}});
					std::sort(ps.begin(), ps.end());
					std::vector<double> k; k.push_back((double)(int)ty[h]);
					for (auto &q : ps) for (double x : q) k.push_back(x);
					ks.insert(k);
				}
				return ks;
			};
			if (sd) sd = cell_keys(da, ta) == cell_keys(db, tb2);
		}
		EXPECT(sd && da.Hs.size() > 0, E == 15 ? "chain --e 15: octree_mesh -> conforming_mesh -> dual_conforming_mesh, same dual cells (geometry + element type)"
		                                        : "chain --e 13: octree_mesh -> conforming_mesh -> dual_conforming_mesh, same dual cells (geometry + element type)");
		// clean_hex_mesh on each side's own octree hex mesh, then the surface of what is kept
		gm.clean_hex_mesh(t, a);
		fpohm_shim::clean_hex_mesh(t, b);
		std::set<std::array<double, 6>> ka, kb;
		for (size_t h = 0; h < a.mesh_entire.Hs.size(); ++h) if (a.H_flag[h]) ka.insert(hex_key(a.mesh_entire, h));
		for (size_t h = 0; h < b.mesh_entire.Hs.size(); ++h) if (b.H_flag[h]) kb.insert(hex_key(b.mesh_entire, h));
		bool sc = same && ka == kb && !ka.empty() && a.mesh_subA.Hs.size() == b.mesh_subA.Hs.size() && a.mesh_subA.Vs.size() == b.mesh_subA.Vs.size();
		Mesh sa, sb;
		sa.type = sb.type = Mesh_type::Qua;
		std::vector<int32_t> vm_a, vr_a, fm_a, fr_a, vm_b, vr_b, fm_b, fr_b;
		::extract_surface_conforming_mesh(a.mesh_subA, sa, vm_a, vr_a, fm_a, fr_a);
		fpohm_shim::extract_surface_conforming_mesh(b.mesh_subA, sb, vm_b, vr_b, fm_b, fr_b);
		auto quad_keys = [](const Mesh &m) {       // oriented quads by geometry: rotation to the smallest corner, direction kept
			std::multiset<std::array<double, 12>> ks;
			for (auto &f : m.Fs) {
				std::array<std::array<double, 3>, 4> ps;
				for (int j = 0; j < 4; ++j) ps[j] = {{m.V(0, f.vs[j]), m.V(1, f.vs[j]), m.V(2, f.vs[j])}};
				int s0 = 0; for (int j = 1; j < 4; ++j) if (ps[j] < ps[s0]) s0 = j;
				std::array<double, 12> k;
				for (int j = 0; j < 4; ++j) for (int d = 0; d < 3; ++d) k[3 * j + d] = ps[(s0 + j) % 4][d];
				ks.insert(k);
			}
			return ks;
		};
		bool ss = sc && sa.Fs.size() == sb.Fs.size() && sa.Vs.size() == sb.Vs.size() && sa.Es.size() == sb.Es.size() && quad_keys(sa) == quad_keys(sb);
		// The cleaning stages depend on the ORDER of the hex and vertex ids (tagging_uneven_element is an in-place sweep in hex
		// order, clean_non_manifold_ve is greedy in vertex / edge order, drop_small_pieces breaks ties by id: ghm.cpp:1983-2124), so
		// the reference's own answer changes when its octree is merely renumbered.  Where the two chains differ, that is what has
		// to be shown: the REFERENCE's clean_hex_mesh run on the product's numbering of the same octree must reproduce the
		// product's result id for id — then the difference is the reference's order dependence, not the product.
		bool by_order = false;
		size_t only_a = 0, only_b = 0;
		if (!ss) {
			for (auto &k : ka) only_a += !kb.count(k);
			for (auto &k : kb) only_b += !ka.count(k);
			Mesh_Domain c;
			c.mesh_entire.type = Mesh_type::Hex;
			c.mesh_entire.V = b.mesh_entire.V;
			c.mesh_entire.Vs.resize(b.mesh_entire.Vs.size());
			for (size_t i = 0; i < c.mesh_entire.Vs.size(); ++i) { c.mesh_entire.Vs[i].id = (uint32_t)i; c.mesh_entire.Vs[i].v = b.mesh_entire.Vs[i].v; }
			fpohm_shim::OctreeGrid ob2;
			std::vector<int> tb3; std::vector<uint32_t> h2o3;
			Mesh fresh; fresh.type = Mesh_type::Hex;
			Eigen::Vector3i gsc;
			fpohm_shim::octree_mesh(dm, Vt.data(), (int64_t)t.V.cols(), fresh, ob2, gsc, 1 << 20, 1 << E, true, tb3, h2o3);   // the hex list BEFORE reorder_hex_mesh
			c.mesh_entire.Hs.resize(fresh.Hs.size());
			for (size_t h = 0; h < fresh.Hs.size(); ++h) { c.mesh_entire.Hs[h].id = (uint32_t)h; c.mesh_entire.Hs[h].vs = fresh.Hs[h].vs; }
			::build_connectivity(c.mesh_entire);
			gm.clean_hex_mesh(t, c);
			by_order = c.H_flag == b.H_flag && c.V_map == b.V_map && c.H_map_reverse == b.H_map_reverse && c.mesh_subA.Hs.size() == b.mesh_subA.Hs.size();
			for (size_t i = 0; by_order && i < c.mesh_subA.Hs.size(); ++i) by_order = c.mesh_subA.Hs[i].vs == b.mesh_subA.Hs[i].vs;
			std::cout.rdbuf(old);
			std::printf("     chain --e %d: %zu hexes; kept by the reference on ITS numbering %zu, by the product %zu (only reference %zu, only product %zu); "
			            "reference clean_hex_mesh on the PRODUCT's numbering: %s the product id for id\n",
			            E, a.mesh_entire.Hs.size(), ka.size(), kb.size(), only_a, only_b, by_order ? "equals" : "DIFFERS FROM");
			old = std::cout.rdbuf(nullptr);
		}
		EXPECT((ss || by_order) && sa.Fs.size() > 0, E == 15 ? "chain --e 15: octree_mesh -> clean_hex_mesh -> extract_surface, same kept hexes and same ORIENTED surface quads (geometry)"
		                                        : "chain --e 13: octree_mesh -> clean_hex_mesh -> extract_surface: same geometry, or a difference the reference reproduces on the product's numbering (order-dependent cleaning)");
		// second pass of the outer loop on the same octree objects: smaller stop extent
		if (E == 15) {
			gm.STOP_EXTENT_MAX = 14;
			Mesh ma2, mb2;
			ma2.type = mb2.type = Mesh_type::Hex;
			const bool o2a = gm.octree_mesh(mi, ma2, oa, gsa);
			const bool o2b = fpohm_shim::octree_mesh(dm, Vt.data(), (int64_t)t.V.cols(), mb2, ob, gsb, 1 << 20, 1 << 14, false, tb, h2o);
			std::set<std::array<double, 6>> k2a, k2b;
			for (size_t h = 0; h < ma2.Hs.size(); ++h) k2a.insert(hex_key(ma2, h));
			for (size_t h = 0; h < mb2.Hs.size(); ++h) k2b.insert(hex_key(mb2, h));
			EXPECT(o2a && o2b && k2a == k2b && ma2.Hs.size() > a.mesh_entire.Hs.size() && ma2.Vs.size() == mb2.Vs.size(), "chain: second octree_mesh pass on the same octree (stop extent 2^14), same leaves");
		}
		std::cout.rdbuf(old);
	}
	// ---- SLIM per-element stages (slim_m.cpp:84-381, 792-916) on the 8-tets-per-hex split of a warped block: gradient operators
	// with 4 entries per tet, deformed positions uv, every energy
	{
		Mesh hb; hex_block(6, 0.8, hb);
		static const int T[8][4] = {{0, 1, 3, 4}, {1, 2, 0, 5}, {2, 3, 1, 6}, {3, 0, 2, 7}, {4, 7, 5, 0}, {5, 4, 6, 1}, {6, 5, 7, 2}, {7, 6, 4, 3}};
		const int nt = 8 * (int)hb.Hs.size(), nv = (int)hb.V.cols();
		auto fill = [&](SLIMData &s) {
			s.dim = 3; s.f_n = s.f_num = nt; s.v_num = nv;
			s.F.resize(nt, 4); s.Ji.resize(nt, 9); s.Ri.resize(nt, 9);
			for (Eigen::VectorXd *w : {&s.W_11, &s.W_12, &s.W_13, &s.W_21, &s.W_22, &s.W_23, &s.W_31, &s.W_32, &s.W_33}) w->resize(nt);
			std::vector<Eigen::Triplet<double>> tx, ty, tz;
			for (size_t h = 0; h < hb.Hs.size(); ++h) for (int k = 0; k < 8; ++k) {
				const int t = 8 * (int)h + k;
				int v[4]; for (int j = 0; j < 4; ++j) { v[j] = (int)hb.Hs[h].vs[T[k][j]]; s.F(t, j) = v[j]; }
				Eigen::Matrix3d E; for (int j = 0; j < 3; ++j) E.row(j) = (hb.V.col(v[j + 1]) - hb.V.col(v[0])).transpose();
				const Eigen::Matrix3d G = E.inverse();                          // column j = gradient of hat function j+1
				const Eigen::Vector3d g0 = -(G.col(0) + G.col(1) + G.col(2));
				for (int j = 0; j < 4; ++j) {
					const Eigen::Vector3d g = j == 0 ? g0 : Eigen::Vector3d(G.col(j - 1));
					tx.emplace_back(t, v[j], g[0]); ty.emplace_back(t, v[j], g[1]); tz.emplace_back(t, v[j], g[2]);
				}
			}
			s.Dx.resize(nt, nv); s.Dy.resize(nt, nv); s.Dz.resize(nt, nv);
			s.Dx.setFromTriplets(tx.begin(), tx.end()); s.Dy.setFromTriplets(ty.begin(), ty.end()); s.Dz.setFromTriplets(tz.begin(), tz.end());
			s.Dx.makeCompressed(); s.Dy.makeCompressed(); s.Dz.makeCompressed();
		};
		Eigen::MatrixXd uv(nv, 3);
		for (int i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) uv(i, c) = hb.V(c, i) * (1.0 + 0.2 * c) + 0.03 * std::sin(11.0 * i + c);
		Eigen::VectorXd areas(nt); for (int i = 0; i < nt; ++i) areas(i) = 0.5 + 0.001 * (i % 97);
		Eigen::MatrixXd Vd; Eigen::MatrixXi Fd;
		bool ok = true; double worst = 0;
		for (int en = 0; en < 6 && ok; ++en) {
			SLIMData a, b; fill(a); fill(b);
			a.slim_energy = b.slim_energy = (SLIM_ENERGY)en; a.exp_factor = b.exp_factor = 0.5;
			Eigen::MatrixXd ua = uv, ub = uv;
			::update_weights_and_closest_rotations(a, Vd, Fd, ua);
			fpohm_shim::update_weights_and_closest_rotations(b, Vd, Fd, ub);
			const Eigen::VectorXd *wa[9] = {&a.W_11, &a.W_12, &a.W_13, &a.W_21, &a.W_22, &a.W_23, &a.W_31, &a.W_32, &a.W_33};
			const Eigen::VectorXd *wb[9] = {&b.W_11, &b.W_12, &b.W_13, &b.W_21, &b.W_22, &b.W_23, &b.W_31, &b.W_32, &b.W_33};
			for (int i = 0; i < nt && ok; ++i) {
				double sw = 0, sr = 0, sj = 0;
				for (int k = 0; k < 9; ++k) { sw = std::max(sw, std::abs((*wa[k])(i))); sr = std::max(sr, std::abs(a.Ri(i, k))); sj = std::max(sj, std::abs(a.Ji(i, k))); }
				for (int k = 0; k < 9 && ok; ++k) {
					// overflowing exponential weights are inf / nan in both implementations: equal non-finite values agree
					auto rel = [&](double x, double y, double sc) { return (x == y || (std::isnan(x) && std::isnan(y))) ? 0.0 : (std::isfinite(x) && std::isfinite(y) ? std::abs(x - y) / sc : 1.0); };
					const double dw = rel((*wa[k])(i), (*wb[k])(i), sw), dr = rel(a.Ri(i, k), b.Ri(i, k), sr), dj = rel(a.Ji(i, k), b.Ji(i, k), sj);
					worst = std::max(worst, std::max(dw, std::max(dr, dj)));
					ok = dw <= 1e-9 && dr <= 1e-9 && dj <= 1e-10;     // Ji: Eigen sums a row of D over columns, the kernel over the CSR row
				}
			}
			const double ea = ::compute_energy_with_jacobians(a, Vd, Fd, a.Ji, ua, areas), eb = fpohm_shim::compute_energy_with_jacobians(b, Vd, Fd, b.Ji, ub, areas);
			if (!(ea == eb || std::abs(ea - eb) <= 1e-9 * std::abs(ea))) { std::printf("     energy %d: %.17g vs %.17g\n", en, ea, eb); ok = false; }
		}
		// line-search step bound along a direction that inverts some tets
		{
			SLIMData a; fill(a);
			Eigen::MatrixXd pos(nv, 3), dir(nv, 3);
			for (int i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) { pos(i, c) = hb.V(c, i); dir(i, c) = 0.4 * std::sin(3.0 * i + 1.7 * c); }
			const double ma = igl::flip_avoiding::compute_max_step_from_singularities(pos, a.F, dir), mb = fpohm_shim::compute_max_step_from_singularities(pos, a.F, dir);
			std::printf("     step bound %.17g vs %.17g\n", ma, mb);
			ok = ok && std::isfinite(ma) && std::abs(ma - mb) <= 1e-9 * ma;
		}
		std::printf("     SLIM stages: worst relative difference %.2e\n", worst);
		EXPECT(ok, "SLIM compute_jacobians / update_weights_and_closest_rotations / compute_energy_with_jacobians / max step within 1e-9 (north star: 1e-5), all six energies");
	}
	std::printf("%s (%d failures)\n", failures ? "SHIM PARITY FAILED" : "SHIM PARITY OK", failures);
	return failures ? 1 : 0;
}

// TEST INFRASTRUCTURE ONLY — extern "C" driver for the reference's own conforming_mesh
// (grid_meshing/grid_hex_meshing.cpp:568-696), compiled UNMODIFIED from /root/reference next to this file.
// The octree comes from the handles of ref_driver.cpp; the octree hex mesh `mo` is filled the way octree_mesh
// does (ghm.cpp:527-562: one vertex per node, one hex per leaf in cell order, then build_connectivity).
#include "grid_meshing/grid_hex_meshing.h"
#include <cstdint>

struct RefOctreeView;   // ref_driver.cpp
extern "C" {
// provided by ref_driver.cpp
const OctreeGrid *ref_octree_grid(void *hv);
void ref_octree_frame(void *hv, double origin[3], double mesh_transform[3], double *voxel_size, int32_t grid_size[3]);

struct RefHybrid { Mesh mo, hybrid, dual; std::vector<Element_Type> types; };

static void fill_sizes(const Mesh &hy, int64_t sizes[8]) {
	int64_t fv = 0, hf = 0, hv = 0, fn = 0;
	for (auto &f : hy.Fs) { fv += (int64_t)f.vs.size(); fn += (int64_t)f.neighbor_hs.size(); }
	for (auto &h : hy.Hs) { hf += (int64_t)h.fs.size(); hv += (int64_t)h.vs.size(); }
	sizes[0] = (int64_t)hy.Vs.size(); sizes[1] = (int64_t)hy.Fs.size(); sizes[2] = (int64_t)hy.Hs.size(); sizes[3] = (int64_t)hy.Es.size();
	sizes[4] = fv; sizes[5] = hf; sizes[6] = hv; sizes[7] = fn;
}

// conforming_mesh on an octree GIVEN AS TABLES (any numbering): node positions / node links fill OctreeGrid::m_Nodes — the
// only octree members the function reads (numNodes, nodePos, m_Nodes[i].neighNodeId, ghm.cpp:572-583) — and `mo` is the
// hex mesh with vertex i = node i.  This lets the product's canonical numbering be fed to the reference unchanged.
void *ref_conforming_mesh_tables(const int32_t *node_pos, const int32_t *node_neigh, int64_t n_nodes, const double *Vpos,
                                 const uint32_t *hex, int64_t n_hex, const int32_t grid_size_in[3], int64_t sizes[8])
{
	RefHybrid *r = new RefHybrid;
	OctreeGrid octree;
	octree.m_Nodes.resize((size_t)n_nodes);
	for (int64_t i = 0; i < n_nodes; ++i) {
		for (int d = 0; d < 3; ++d) octree.m_Nodes[(size_t)i].position[d] = node_pos[3 * i + d];
		for (int k = 0; k < 6; ++k) octree.m_Nodes[(size_t)i].neighNodeId[k] = node_neigh[6 * i + k];
	}
	Mesh &mo = r->mo;
	mo.type = Mesh_type::Hex;
	mo.Vs.resize((size_t)n_nodes);
	mo.V.resize(3, n_nodes);
	for (int64_t i = 0; i < n_nodes; ++i) {
		Hybrid_V v; v.id = (uint32_t)i;
		for (int d = 0; d < 3; ++d) { v.v.push_back(Vpos[3 * i + d]); mo.V(d, i) = Vpos[3 * i + d]; }
		mo.Vs[(size_t)i] = v;
	}
	mo.Hs.resize((size_t)n_hex);
	for (int64_t h = 0; h < n_hex; ++h) { mo.Hs[(size_t)h].id = (uint32_t)h; mo.Hs[(size_t)h].vs.assign(hex + 8 * h, hex + 8 * h + 8); }
	build_connectivity(mo);
	grid_hex_meshing_bijective gm;
	Eigen::Vector3i grid_size(grid_size_in[0], grid_size_in[1], grid_size_in[2]);
	gm.conforming_mesh(mo, r->hybrid, octree, grid_size);
	fill_sizes(r->hybrid, sizes);
	return r;
}

void *ref_conforming_mesh(void *octree_handle, int64_t sizes[8]) {
	const OctreeGrid &octree_c = *ref_octree_grid(octree_handle);
	OctreeGrid &octree = const_cast<OctreeGrid &>(octree_c);
	double org[3], mt[3], vs; int32_t gs[3];
	ref_octree_frame(octree_handle, org, mt, &vs, gs);
	RefHybrid *r = new RefHybrid;
	Mesh &mo = r->mo;
	mo.type = Mesh_type::Hex;
	const int nn = octree.numNodes();
	mo.Vs.resize(nn);
	mo.V.resize(3, nn);
	for (int i = 0; i < nn; ++i) {
		const Eigen::Vector3i p = octree.nodePos(i);
		Hybrid_V v; v.id = i;
		for (int d = 0; d < 3; ++d) { const double x = (mt[d] + org[d]) + (double)p[d] * vs; v.v.push_back(x); mo.V(d, i) = x; }
		mo.Vs[i] = v;
	}
	for (int q = 0; q < octree.numCells(); ++q) {
		if (!octree.cellIsLeaf(q)) continue;
		Hybrid h; h.id = (uint32_t)mo.Hs.size(); h.vs.resize(8);
		for (int lv = 0; lv < 8; ++lv) h.vs[lv] = octree.cellCornerId(q, lv);
		mo.Hs.push_back(h);
	}
	build_connectivity(mo);
	grid_hex_meshing_bijective gm;
	Eigen::Vector3i grid_size(gs[0], gs[1], gs[2]);
	gm.conforming_mesh(mo, r->hybrid, octree, grid_size);
	fill_sizes(r->hybrid, sizes);
	return r;
}
// F_off nF+1, F_vs/F_es sizes[4], F_boundary nF, E_vs 2 nE, E_boundary nE, V_boundary nV,
// H_foff nH+1, H_fs sizes[5], H_voff nH+1, H_vs sizes[6], F_nhoff nF+1, F_nhs sizes[7]
static void export_mesh(const Mesh &hy, int64_t *F_off, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs, uint8_t *E_boundary,
                        uint8_t *V_boundary, int64_t *H_foff, uint32_t *H_fs, int64_t *H_voff, uint32_t *H_vs, int64_t *F_nhoff, uint32_t *F_nhs)
{
	int64_t t = 0, u = 0;
	for (size_t f = 0; f < hy.Fs.size(); ++f) {
		F_off[f] = t; F_nhoff[f] = u;
		for (size_t k = 0; k < hy.Fs[f].vs.size(); ++k) { F_vs[t] = hy.Fs[f].vs[k]; F_es[t] = hy.Fs[f].es[k]; ++t; }
		for (uint32_t h : hy.Fs[f].neighbor_hs) F_nhs[u++] = h;
		F_boundary[f] = hy.Fs[f].boundary;
	}
	F_off[hy.Fs.size()] = t; F_nhoff[hy.Fs.size()] = u;
	for (size_t e = 0; e < hy.Es.size(); ++e) { E_vs[2 * e] = hy.Es[e].vs[0]; E_vs[2 * e + 1] = hy.Es[e].vs[1]; E_boundary[e] = hy.Es[e].boundary; }
	for (size_t v = 0; v < hy.Vs.size(); ++v) V_boundary[v] = hy.Vs[v].boundary;
	int64_t a = 0, b = 0;
	for (size_t h = 0; h < hy.Hs.size(); ++h) {
		H_foff[h] = a; H_voff[h] = b;
		for (uint32_t x : hy.Hs[h].fs) H_fs[a++] = x;
		for (uint32_t x : hy.Hs[h].vs) H_vs[b++] = x;
	}
	H_foff[hy.Hs.size()] = a; H_voff[hy.Hs.size()] = b;
}
void ref_hybrid_export(void *rv, int64_t *F_off, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs, uint8_t *E_boundary,
                       uint8_t *V_boundary, int64_t *H_foff, uint32_t *H_fs, int64_t *H_voff, uint32_t *H_vs, int64_t *F_nhoff, uint32_t *F_nhs)
{
	export_mesh(((RefHybrid *)rv)->hybrid, F_off, F_vs, F_es, F_boundary, E_vs, E_boundary, V_boundary, H_foff, H_fs, H_voff, H_vs, F_nhoff, F_nhs);
}

// dual_conforming_mesh (ghm.cpp:697-872) on the result of ref_conforming_mesh*: dual polyhedral mesh + element types
void ref_dual_conforming_mesh(void *rv, int64_t sizes[8]) {
	RefHybrid *r = (RefHybrid *)rv;
	grid_hex_meshing_bijective gm;
	r->dual = Mesh(); r->types.clear();
	gm.dual_conforming_mesh(r->mo, r->hybrid, r->dual, r->types);
	fill_sizes(r->dual, sizes);
}
void ref_dual_export(void *rv, double *V, int32_t *h_type, int64_t *F_off, uint32_t *F_vs, uint32_t *F_es, uint8_t *F_boundary, uint32_t *E_vs,
                     uint8_t *E_boundary, uint8_t *V_boundary, int64_t *H_foff, uint32_t *H_fs, int64_t *H_voff, uint32_t *H_vs, int64_t *F_nhoff, uint32_t *F_nhs)
{
	RefHybrid *r = (RefHybrid *)rv;
	const Mesh &d = r->dual;
	for (size_t v = 0; v < d.Vs.size(); ++v) for (int k = 0; k < 3; ++k) V[3 * v + k] = d.V(k, v);
	for (size_t h = 0; h < d.Hs.size(); ++h) h_type[h] = (int32_t)r->types[h];
	export_mesh(d, F_off, F_vs, F_es, F_boundary, E_vs, E_boundary, V_boundary, H_foff, H_fs, H_voff, H_vs, F_nhoff, F_nhs);
}
void ref_hybrid_free(void *rv) { delete (RefHybrid *)rv; }

// ---- §8(f)-2: clean_hex_mesh and its stages (ghm.cpp:1932-2126, gf.cpp:664-698,2199-2229), the compiled reference
// methods themselves on a hex mesh given as (Vpos, hex).  Every stage is callable on its own so that the product's
// stage entry points can be compared one by one.
struct RefClean { Mesh_Domain md; Eigen::VectorXd signed_dis; };

void *ref_clean_new(const double *Vpos, int64_t nV, const uint32_t *hex, int64_t H) {
	RefClean *r = new RefClean;
	Mesh &m = r->md.mesh_entire;
	m.type = Mesh_type::Hex;
	m.Vs.resize((size_t)nV);
	m.V.resize(3, nV);
	for (int64_t i = 0; i < nV; ++i) {
		Hybrid_V v; v.id = (uint32_t)i;
		for (int d = 0; d < 3; ++d) { v.v.push_back(Vpos[3 * i + d]); m.V(d, i) = Vpos[3 * i + d]; }
		m.Vs[(size_t)i] = v;
	}
	m.Hs.resize((size_t)H);
	for (int64_t h = 0; h < H; ++h) { m.Hs[(size_t)h].id = (uint32_t)h; m.Hs[(size_t)h].vs.assign(hex + 8 * h, hex + 8 * h + 8); }
	build_connectivity(m);
	r->md.H_flag.assign((size_t)H, false);
	r->signed_dis = Eigen::VectorXd::Zero(H);
	return r;
}
void ref_clean_free(void *rv) { delete (RefClean *)rv; }
void ref_clean_reorder(void *rv, uint32_t *hex_out) {
	RefClean *r = (RefClean *)rv;
	reorder_hex_mesh(r->md.mesh_entire);
	const Mesh &m = r->md.mesh_entire;
	for (size_t h = 0; h < m.Hs.size(); ++h) for (int k = 0; k < 8; ++k) hex_out[8 * h + k] = m.Hs[h].vs[k];
}
void ref_clean_set_flags(void *rv, const uint8_t *f) {
	RefClean *r = (RefClean *)rv;
	for (size_t i = 0; i < r->md.H_flag.size(); ++i) r->md.H_flag[i] = f[i] != 0;
}
void ref_clean_get_flags(void *rv, uint8_t *f) {
	RefClean *r = (RefClean *)rv;
	for (size_t i = 0; i < r->md.H_flag.size(); ++i) f[i] = r->md.H_flag[i] ? 1 : 0;
}
void ref_clean_tagging(void *rv) {
	RefClean *r = (RefClean *)rv;
	grid_hex_meshing_bijective gm;
	gm.tagging_uneven_element(r->md.mesh_entire, r->md.H_flag);
}
void ref_clean_reindex(void *rv, int64_t sizes[4]) {
	RefClean *r = (RefClean *)rv;
	Mesh_Domain &md = r->md;
	re_indexing_connectivity(md.mesh_entire, md.H_flag, md.mesh_subA, md.V_map, md.V_map_reverse, md.H_map, md.H_map_reverse);
	sizes[0] = (int64_t)md.mesh_subA.Vs.size(); sizes[1] = (int64_t)md.mesh_subA.Hs.size();
	sizes[2] = (int64_t)md.mesh_subA.Fs.size(); sizes[3] = (int64_t)md.mesh_subA.Es.size();
}
// V_map nV(entire), V_map_reverse nV(sub), H_map_reverse nH(sub), sub_hex 8 nH(sub), sub_V 3 nV(sub); any may be NULL
void ref_clean_sub_export(void *rv, int32_t *V_map, int32_t *V_map_reverse, int32_t *H_map_reverse, uint32_t *sub_hex, double *sub_V) {
	RefClean *r = (RefClean *)rv;
	const Mesh_Domain &md = r->md;
	if (V_map) for (size_t i = 0; i < md.V_map.size(); ++i) V_map[i] = md.V_map[i];
	if (V_map_reverse) for (size_t i = 0; i < md.V_map_reverse.size(); ++i) V_map_reverse[i] = md.V_map_reverse[i];
	if (H_map_reverse) for (size_t i = 0; i < md.H_map_reverse.size(); ++i) H_map_reverse[i] = md.H_map_reverse[i];
	if (sub_hex) for (size_t h = 0; h < md.mesh_subA.Hs.size(); ++h) for (int k = 0; k < 8; ++k) sub_hex[8 * h + k] = md.mesh_subA.Hs[h].vs[k];
	if (sub_V) for (size_t v = 0; v < md.mesh_subA.Vs.size(); ++v) for (int d = 0; d < 3; ++d) sub_V[3 * v + d] = md.mesh_subA.V(d, v);
}
// needs ref_clean_reindex first (mesh_subA + maps are inputs of both)
void ref_clean_non_manifold(void *rv) {
	RefClean *r = (RefClean *)rv;
	Mesh_Domain &md = r->md;
	grid_hex_meshing_bijective gm;
	gm.clean_non_manifold_ve(md.mesh_entire, md.mesh_subA, md.V_map, md.V_map_reverse, md.H_map, md.H_map_reverse, r->signed_dis, md.H_flag);
}
void ref_clean_drop_small(void *rv) {
	RefClean *r = (RefClean *)rv;
	grid_hex_meshing_bijective gm;
	gm.drop_small_pieces(r->md);
}
// the whole clean_hex_mesh (args.scaffold_type keeps its default 1: no scaffold layers)
void ref_clean_full(void *rv, const double *tV, int64_t ntV, const int32_t *tF, int64_t ntF) {
	RefClean *r = (RefClean *)rv;
	Mesh tmi;
	tmi.type = Mesh_type::Tri;
	tmi.V.resize(3, ntV);
	tmi.Vs.resize((size_t)ntV);
	for (int64_t i = 0; i < ntV; ++i) { for (int c = 0; c < 3; ++c) tmi.V(c, i) = tV[3 * i + c]; tmi.Vs[(size_t)i].id = (uint32_t)i; }
	tmi.Fs.resize((size_t)ntF);
	for (int64_t f = 0; f < ntF; ++f) {
		tmi.Fs[(size_t)f].id = (uint32_t)f;
		tmi.Fs[(size_t)f].vs = {(uint32_t)tF[3 * f], (uint32_t)tF[3 * f + 1], (uint32_t)tF[3 * f + 2]};
	}
	grid_hex_meshing_bijective gm;
	gm.clean_hex_mesh(tmi, r->md);
}
void ref_clean_entire_sizes(void *rv, int64_t sizes[4]) {
	const Mesh &m = ((RefClean *)rv)->md.mesh_entire;
	sizes[0] = (int64_t)m.Vs.size(); sizes[1] = (int64_t)m.Hs.size(); sizes[2] = (int64_t)m.Fs.size(); sizes[3] = (int64_t)m.Es.size();
}
void ref_clean_medial(void *rv, uint8_t *F_medial, uint8_t *V_medial) {
	const Mesh &m = ((RefClean *)rv)->md.mesh_entire;
	for (size_t f = 0; f < m.Fs.size(); ++f) F_medial[f] = m.Fs[f].on_medial_surface ? 1 : 0;
	for (size_t v = 0; v < m.Vs.size(); ++v) V_medial[v] = m.Vs[v].on_medial_surface ? 1 : 0;
}

// ---- extract_surface_conforming_mesh (global_functions.cpp:1021-1072: boundary faces -> quad / triangle surface,
// build_connectivity, orient_surface_mesh :1073-1112, build_connectivity) on a hex mesh given as (Vpos, hex)
struct RefSurface { Mesh hexm, sur; std::vector<int32_t> V_map, V_map_reverse, F_map, F_map_reverse; };
void *ref_extract_surface(const double *Vpos, int64_t nV, const uint32_t *hex, int64_t H, int as_triangles, int64_t sizes[6]) {
	RefSurface *r = new RefSurface;
	Mesh &m = r->hexm;
	m.type = Mesh_type::Hex;
	m.Vs.resize((size_t)nV);
	m.V.resize(3, nV);
	for (int64_t i = 0; i < nV; ++i) {
		Hybrid_V v; v.id = (uint32_t)i;
		for (int d = 0; d < 3; ++d) { v.v.push_back(Vpos[3 * i + d]); m.V(d, i) = Vpos[3 * i + d]; }
		m.Vs[(size_t)i] = v;
	}
	m.Hs.resize((size_t)H);
	for (int64_t h = 0; h < H; ++h) { m.Hs[(size_t)h].id = (uint32_t)h; m.Hs[(size_t)h].vs.assign(hex + 8 * h, hex + 8 * h + 8); }
	build_connectivity(m);
	r->sur.type = as_triangles ? Mesh_type::Tri : Mesh_type::Qua;
	extract_surface_conforming_mesh(m, r->sur, r->V_map, r->V_map_reverse, r->F_map, r->F_map_reverse);
	int64_t nfs = 0, nvf = 0;
	for (auto &e : r->sur.Es) nfs += (int64_t)e.neighbor_fs.size();
	for (auto &v : r->sur.Vs) nvf += (int64_t)v.neighbor_fs.size();
	sizes[0] = (int64_t)r->sur.Vs.size(); sizes[1] = (int64_t)r->sur.Fs.size(); sizes[2] = (int64_t)r->sur.Es.size();
	sizes[3] = nfs; sizes[4] = nvf; sizes[5] = (int64_t)m.Fs.size();
	return r;
}
// V 3 nV, F_vs / F_es vn nF, E_vs 2 nE, flags, V_map nV(hex mesh), V_map_reverse nV, F_map nF(hex mesh), F_map_reverse nF
void ref_surface_export(void *rv, double *V, uint32_t *F_vs, uint32_t *F_es, uint32_t *E_vs, uint8_t *E_boundary, uint8_t *V_boundary,
                        int32_t *V_map, int32_t *V_map_reverse, int32_t *F_map, int32_t *F_map_reverse)
{
	RefSurface *r = (RefSurface *)rv;
	const Mesh &s = r->sur;
	for (size_t v = 0; v < s.Vs.size(); ++v) { for (int d = 0; d < 3; ++d) V[3 * v + d] = s.V(d, v); V_boundary[v] = s.Vs[v].boundary; }
	size_t t = 0;
	for (size_t f = 0; f < s.Fs.size(); ++f) for (size_t k = 0; k < s.Fs[f].vs.size(); ++k, ++t) { F_vs[t] = s.Fs[f].vs[k]; F_es[t] = s.Fs[f].es[k]; }
	for (size_t e = 0; e < s.Es.size(); ++e) { E_vs[2 * e] = s.Es[e].vs[0]; E_vs[2 * e + 1] = s.Es[e].vs[1]; E_boundary[e] = s.Es[e].boundary; }
	for (size_t i = 0; i < r->V_map.size(); ++i) V_map[i] = r->V_map[i];
	for (size_t i = 0; i < r->V_map_reverse.size(); ++i) V_map_reverse[i] = r->V_map_reverse[i];
	for (size_t i = 0; i < r->F_map.size(); ++i) F_map[i] = r->F_map[i];
	for (size_t i = 0; i < r->F_map_reverse.size(); ++i) F_map_reverse[i] = r->F_map_reverse[i];
}
// which: 0 E.neighbor_fs 1 V.neighbor_vs 2 V.neighbor_es 3 V.neighbor_fs; off has n+1 entries
void ref_surface_csr(void *rv, int which, int64_t *off, uint32_t *val) {
	const Mesh &s = ((RefSurface *)rv)->sur;
	const size_t n = which == 0 ? s.Es.size() : s.Vs.size();
	int64_t t = 0;
	for (size_t i = 0; i < n; ++i) {
		const std::vector<uint32_t> &l = which == 0 ? s.Es[i].neighbor_fs : which == 1 ? s.Vs[i].neighbor_vs : which == 2 ? s.Vs[i].neighbor_es : s.Vs[i].neighbor_fs;
		off[i] = t;
		for (uint32_t x : l) val[t++] = x;
	}
	off[n] = t;
}
void ref_surface_free(void *rv) { delete (RefSurface *)rv; }
// grid_hex_meshing_bijective::voxel_meshing itself (ghm.cpp:215-296, the `--o 0` lattice): bbox of the input GEO::Mesh,
// dim = ceil(extent / (max_extent / num_voxels)), float grid_length, vertex lattice + unit hexes + build_connectivity.
// Two-phase: sizes first (dims[3], nV, nH), then the arrays.
struct RefLattice { Mesh hmi; };
void *ref_voxel_meshing(const double *V, int64_t nV, const int32_t *F, int64_t nF, int num_voxels, int64_t sizes[2]) {
	GEO::Mesh mi;
	mi.vertices.create_vertices((GEO::index_t)nV);
	for (int64_t i = 0; i < nV; ++i) mi.vertices.point((GEO::index_t)i) = GEO::vec3(V[3 * i], V[3 * i + 1], V[3 * i + 2]);
	mi.facets.create_triangles((GEO::index_t)nF);
	for (int64_t f = 0; f < nF; ++f) for (int c = 0; c < 3; ++c) mi.facets.set_vertex((GEO::index_t)f, c, (GEO::index_t)F[3 * f + c]);
	RefLattice *r = new RefLattice;
	r->hmi.type = Mesh_type::Hex;
	grid_hex_meshing_bijective gm;
	gm.num_voxels = num_voxels;
	gm.voxel_meshing(mi, r->hmi);
	sizes[0] = (int64_t)r->hmi.Vs.size(); sizes[1] = (int64_t)r->hmi.Hs.size();
	return r;
}
void ref_voxel_meshing_export(void *rv, double *Vpos, uint32_t *hex) {
	const Mesh &m = ((RefLattice *)rv)->hmi;
	for (size_t v = 0; v < m.Vs.size(); ++v) for (int d = 0; d < 3; ++d) Vpos[3 * v + d] = m.V(d, v);
	for (size_t h = 0; h < m.Hs.size(); ++h) for (int k = 0; k < 8; ++k) hex[8 * h + k] = m.Hs[h].vs[k];
}
void ref_voxel_meshing_free(void *rv) { delete (RefLattice *)rv; }
}

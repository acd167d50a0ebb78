// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// extern "C" driver over the UNMODIFIED reference sources compiled in place by
// oracle/ref/Makefile (grid_meshing/octree.cpp, grid_meshing/voxelization.cpp,
// global_functions.cpp, metro_hausdorff.cpp + vendored geogram / libigl / VCG).
// Each entry point marshals plain arrays into the reference's own containers
// (GEO::Mesh, Mesh, Eigen matrices), calls the reference function, and copies
// the result out.  The few lines that live *inside* grid_hex_meshing.cpp's
// class (the should_subdivide lambda, ghm.cpp:502-517; the grid set-up,
// ghm.cpp:463-493; the LINE branch of dirty_graph_projection, ghm.cpp:3967-3994)
// cannot be linked on their own, so they are restated here around the real
// OctreeGrid / MeshFacetsAABB / point_line_projection objects.
#include "grid_meshing/voxelization.h"
#include "global_types.h"
#include "global_functions.h"
#include "metro_hausdorff.h"

#include <geogram/basic/common.h>
#include <geogram/basic/command_line.h>
#include <geogram/basic/command_line_args.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/process.h>
#include <igl/AABB.h>
#include <igl/per_face_normals.h>
#include <igl/per_vertex_normals.h>
#include <igl/per_edge_normals.h>
#include <igl/signed_distance.h>
#include <igl/point_mesh_squared_distance.h>

#include <cstdint>
#include <cstring>
#include <mutex>

namespace {

std::once_flag g_once;
void ensure_init() {
	std::call_once(g_once, []() {
		GEO::initialize();
		GEO::CmdLine::import_arg_group("standard");
		GEO::CmdLine::import_arg_group("algo");
		GEO::Logger::instance()->set_quiet(true);
	});
}

void fill_geomesh(GEO::Mesh &M, const double *V, int64_t nV, const int32_t *F, int64_t nF) {
	M.clear(false, false);
	M.vertices.create_vertices((GEO::index_t)nV);
	for (int64_t i = 0; i < nV; ++i) {
		M.vertices.point((GEO::index_t)i) = GEO::vec3(V[3 * i], V[3 * i + 1], V[3 * i + 2]);
	}
	M.facets.create_triangles((GEO::index_t)nF);
	for (int64_t f = 0; f < nF; ++f) {
		for (int c = 0; c < 3; ++c) M.facets.set_vertex((GEO::index_t)f, c, (GEO::index_t)F[3 * f + c]);
	}
}

// Mesh (global_types.h:512) of type Tri with V (3 x n) and Fs[].vs
void fill_trimesh(Mesh &m, const double *V, int64_t nV, const int32_t *F, int64_t nF) {
	m.type = Mesh_type::Tri;
	m.V.resize(3, nV);
	m.Vs.resize(nV);
	for (int64_t i = 0; i < nV; ++i) {
		for (int c = 0; c < 3; ++c) m.V(c, i) = V[3 * i + c];
		m.Vs[i].id = (uint32_t)i;
	}
	m.Fs.resize(nF);
	for (int64_t f = 0; f < nF; ++f) {
		m.Fs[f].id = (uint32_t)f;
		m.Fs[f].vs = {(uint32_t)F[3 * f], (uint32_t)F[3 * f + 1], (uint32_t)F[3 * f + 2]};
	}
}

struct RefOctree {
	GEO::Mesh M;
	GEO::MeshFacetsAABB *aabb = nullptr;
	OctreeGrid octree;
	GEO::vec3 origin;
	Eigen::Vector3d mesh_transform;
	double voxel_size = 0;
	Eigen::Vector3i grid_size;
	bool graded = true, paired = true;
	~RefOctree() { delete aabb; }
};

struct RefTree {
	Eigen::MatrixXd V;
	Eigen::MatrixXi F;
	igl::AABB<Eigen::MatrixXd, 3> tree;
	Eigen::MatrixXd FN, VN, EN;
	Eigen::MatrixXi E;
	Eigen::VectorXi EMAP;
};

} // namespace

extern "C" {

int ref_num_cores() { ensure_init(); return (int)GEO::Process::number_of_cores(); }

// ---------------------------------------------------------------------------------------
// ghm.cpp:463-493 — grid set-up of octree_mesh (scaffold types 2/3 not exercised: args default 1)
// out: grid_size[3], origin[3], mesh_transform[3], voxel_size
void ref_octree_grid_setup(const double *V, int64_t nV, const int32_t *F, int64_t nF, int num_voxels,
	int32_t *grid_size, double *origin_out, double *mesh_transform_out, double *voxel_size_out)
{
	ensure_init();
	GEO::Mesh mi;
	fill_geomesh(mi, V, nV, F, nF);
	GEO::vec3 min_corner, max_corner;
	GEO::get_bbox(mi, &min_corner[0], &max_corner[0]);
	GEO::vec3 mesh_center = (min_corner + max_corner) / 2;
	GEO::vec3 extent = (max_corner - min_corner);
	double voxel_size = 0;
	if (num_voxels > 0) {
		double max_extent = std::max(extent[0], std::max(extent[1], extent[2]));
		voxel_size = max_extent / num_voxels;
	}
	int padding = 0;
	GEO::vec3 origin = min_corner - padding * voxel_size * GEO::vec3(1, 1, 1);
	Eigen::Vector3i gs(
		next_pow2(std::ceil(extent[0] / voxel_size) + 2 * padding),
		next_pow2(std::ceil(extent[1] / voxel_size) + 2 * padding),
		next_pow2(std::ceil(extent[2] / voxel_size) + 2 * padding));
	GEO::vec3 origin_max = origin + GEO::vec3(voxel_size * gs[0], voxel_size * gs[1], voxel_size * gs[2]);
	GEO::vec3 origin_center = (origin_max + origin) * 0.5;
	for (int d = 0; d < 3; ++d) {
		grid_size[d] = gs[d];
		origin_out[d] = origin[d];
		mesh_transform_out[d] = mesh_center[d] - origin_center[d];
	}
	*voxel_size_out = voxel_size;
}

// ---------------------------------------------------------------------------------------
// OctreeGrid + should_subdivide (ghm.cpp:495-524).  stop_extent is in finest-voxel units.
void *ref_octree_build(const double *V, int64_t nV, const int32_t *F, int64_t nF,
	const int32_t *grid_size, const double *origin, const double *mesh_transform, double voxel_size,
	int stop_extent, int graded, int paired)
{
	ensure_init();
	RefOctree *h = new RefOctree;
	fill_geomesh(h->M, V, nV, F, nF);
	h->aabb = new GEO::MeshFacetsAABB(h->M); // Morton-reorders h->M's facets, mesh_AABB.cpp:337
	h->origin = GEO::vec3(origin[0], origin[1], origin[2]);
	h->mesh_transform = Eigen::Vector3d(mesh_transform[0], mesh_transform[1], mesh_transform[2]);
	h->voxel_size = voxel_size;
	h->grid_size = Eigen::Vector3i(grid_size[0], grid_size[1], grid_size[2]);
	h->graded = graded; h->paired = paired;
	h->octree.OctreeGrid_initialize(h->grid_size);
	const GEO::vec3 &o = h->origin; const Eigen::Vector3d &mt = h->mesh_transform; const double vs = voxel_size;
	GEO::MeshFacetsAABB &aabb_tree = *h->aabb;
	auto should_subdivide = [&](int x, int y, int z, int extent) {
		if (extent <= stop_extent) return false;
		GEO::Box box;
		box.xyz_min[0] = mt[0] + o[0] + vs * x;
		box.xyz_min[1] = mt[1] + o[1] + vs * y;
		box.xyz_min[2] = mt[2] + o[2] + vs * z;
		box.xyz_max[0] = box.xyz_min[0] + vs * extent;
		box.xyz_max[1] = box.xyz_min[1] + vs * extent;
		box.xyz_max[2] = box.xyz_min[2] + vs * extent;
		bool has_triangles = false;
		auto action = [&has_triangles](int) { has_triangles = true; };
		aabb_tree.compute_bbox_facet_bbox_intersections(box, action);
		return has_triangles;
	};
	h->octree.subdivide(should_subdivide, (bool)graded, (bool)paired);
	return h;
}

// incremental refinement of listed cells, octree.cpp:691-729 via ghm.cpp:518-521
void ref_octree_refine(void *hv, const int32_t *cells, int64_t n, int stop_extent) {
	RefOctree *h = (RefOctree *)hv;
	std::vector<int> list(cells, cells + n);
	const GEO::vec3 &o = h->origin; const Eigen::Vector3d &mt = h->mesh_transform; const double vs = h->voxel_size;
	GEO::MeshFacetsAABB &aabb_tree = *h->aabb;
	auto should_subdivide = [&](int x, int y, int z, int extent) {
		if (extent <= stop_extent) return false;
		GEO::Box box;
		box.xyz_min[0] = mt[0] + o[0] + vs * x;
		box.xyz_min[1] = mt[1] + o[1] + vs * y;
		box.xyz_min[2] = mt[2] + o[2] + vs * z;
		box.xyz_max[0] = box.xyz_min[0] + vs * extent;
		box.xyz_max[1] = box.xyz_min[1] + vs * extent;
		box.xyz_max[2] = box.xyz_min[2] + vs * extent;
		bool has_triangles = false;
		auto action = [&has_triangles](int) { has_triangles = true; };
		aabb_tree.compute_bbox_facet_bbox_intersections(box, action);
		return has_triangles;
	};
	h->octree.subdivide(should_subdivide, list, h->graded, h->paired);
}

// random-split fuzz of the reference itself (octree.cpp:900-961); predicate-free topology oracle
void *ref_octree_random(const int32_t *grid_size, int graded, int paired) {
	ensure_init();
	RefOctree *h = new RefOctree;
	h->grid_size = Eigen::Vector3i(grid_size[0], grid_size[1], grid_size[2]);
	h->octree.OctreeGrid_initialize(h->grid_size);
	std::streambuf *old = std::cout.rdbuf(nullptr);
	h->octree.testSubdivideRandom((bool)graded, (bool)paired);
	std::cout.rdbuf(old);
	return h;
}

// explicit split list driven through subdivide(): split cell iff listed (x,y,z,extent) — used to feed the
// SAME predicate set to the reference and to the GPU without any floating point in between.
void *ref_octree_from_marks(const int32_t *grid_size, const int32_t *marks /*4 x n: x,y,z,extent*/, int64_t n,
	int graded, int paired)
{
	ensure_init();
	RefOctree *h = new RefOctree;
	h->grid_size = Eigen::Vector3i(grid_size[0], grid_size[1], grid_size[2]);
	h->graded = graded; h->paired = paired;
	h->octree.OctreeGrid_initialize(h->grid_size);
	std::set<std::array<int, 4>> S;
	for (int64_t i = 0; i < n; ++i) S.insert({marks[4 * i], marks[4 * i + 1], marks[4 * i + 2], marks[4 * i + 3]});
	auto pred = [&](int x, int y, int z, int e) { return S.count({x, y, z, e}) > 0; };
	h->octree.subdivide(pred, (bool)graded, (bool)paired);
	return h;
}

void ref_octree_sizes(void *hv, int64_t *nNodes, int64_t *nCells, int64_t *nLeaves, int32_t *nRoots, int32_t *maxDepth) {
	RefOctree *h = (RefOctree *)hv;
	*nNodes = h->octree.numNodes();
	*nCells = h->octree.numCells();
	int64_t l = 0;
	for (int c = 0; c < h->octree.numCells(); ++c) l += h->octree.cellIsLeaf(c);
	*nLeaves = l;
	*nRoots = h->octree.m_NumRootCells;
	*maxDepth = h->octree.m_MaxDepth;
}

void ref_octree_export(void *hv, int32_t *node_pos, int32_t *node_neigh, int32_t *cell_first_child,
	int32_t *cell_corner, int32_t *cell_neigh)
{
	RefOctree *h = (RefOctree *)hv;
	const OctreeGrid &o = h->octree;
	for (int i = 0; i < o.numNodes(); ++i) {
		for (int d = 0; d < 3; ++d) node_pos[3 * i + d] = o.m_Nodes[i].position[d];
		for (int d = 0; d < 6; ++d) node_neigh[6 * i + d] = o.m_Nodes[i].neighNodeId[d];
	}
	for (int c = 0; c < o.numCells(); ++c) {
		cell_first_child[c] = o.m_Cells[c].firstChild;
		for (int k = 0; k < 8; ++k) cell_corner[8 * c + k] = o.m_Cells[c].cornerNodeId[k];
		for (int k = 0; k < 6; ++k) cell_neigh[6 * c + k] = o.m_Cells[c].neighCellId[k];
	}
}

int ref_octree_flags(void *hv) {
	RefOctree *h = (RefOctree *)hv;
	return (h->octree.is2to1Graded() ? 1 : 0) | (h->octree.isPaired() ? 2 : 0);
}

// ray-parity inside flag per cell: compute_sign(OctreeGrid) voxelization.cpp:101-163
// (origin/spacing as compute_octree passes them, voxelization.cpp:384)
void ref_octree_cell_sign(void *hv, const double *origin, double spacing, float *inside) {
	RefOctree *h = (RefOctree *)hv;
	compute_sign(h->M, *h->aabb, h->octree, GEO::vec3(origin[0], origin[1], origin[2]), spacing);
	const Eigen::VectorXf &in = h->octree.cellAttributes.get<float>("inside");
	for (int c = 0; c < h->octree.numCells(); ++c) inside[c] = in(c);
}

// hex export of octree_mesh, ghm.cpp:527-562
void ref_octree_hexes(void *hv, double *Vpos, uint32_t *hex, int32_t *hex2cell) {
	RefOctree *h = (RefOctree *)hv;
	const OctreeGrid &octree = h->octree;
	Eigen::Vector3d o(h->origin[0], h->origin[1], h->origin[2]);
	Eigen::Vector3d s(h->voxel_size, h->voxel_size, h->voxel_size);
	for (int idx = 0; idx < octree.numNodes(); ++idx) {
		Eigen::Vector3d pos = h->mesh_transform + o + octree.nodePos(idx).cast<double>().cwiseProduct(s);
		for (int d = 0; d < 3; ++d) Vpos[3 * idx + d] = pos[d];
	}
	for (int q = 0, c = 0; q < octree.numCells(); ++q) {
		if (!octree.cellIsLeaf(q)) continue;
		hex2cell[c] = q;
		for (int lv = 0; lv < 8; ++lv) hex[8 * c + lv] = octree.cellCornerId(q, lv);
		++c;
	}
}

void ref_octree_free(void *hv) { delete (RefOctree *)hv; }
// views for ref_driver_ghm.cpp
const OctreeGrid *ref_octree_grid(void *hv) { return &((RefOctree *)hv)->octree; }
void ref_octree_frame(void *hv, double origin[3], double mesh_transform[3], double *voxel_size, int32_t grid_size[3]) {
	RefOctree *h = (RefOctree *)hv;
	for (int d = 0; d < 3; ++d) { origin[d] = h->origin[d]; mesh_transform[d] = h->mesh_transform[d]; grid_size[d] = h->grid_size[d]; }
	*voxel_size = h->voxel_size;
}

// ---------------------------------------------------------------------------------------
// compute_octree, voxelization.cpp:353-391 (the one public end-to-end entry of voxelization.h)
// returns geogram hex mesh: vertices + hexes (geogram corner order) + "inside" cell attribute
void *ref_compute_octree(const double *V, int64_t nV, const int32_t *F, int64_t nF,
	const double *min_corner, const double *extent, double spacing, int padding, int graded, int paired,
	int64_t *n_vertices, int64_t *n_hexes)
{
	ensure_init();
	GEO::Mesh *M = new GEO::Mesh, *mo = new GEO::Mesh;
	fill_geomesh(*M, V, nV, F, nF);
	GEO::MeshFacetsAABB aabb(*M);
	compute_octree(*M, *mo, aabb, "", GEO::vec3(min_corner[0], min_corner[1], min_corner[2]),
		GEO::vec3(extent[0], extent[1], extent[2]), spacing, padding, (bool)graded, (bool)paired);
	delete M;
	*n_vertices = mo->vertices.nb();
	*n_hexes = mo->cells.nb();
	return mo;
}
void ref_compute_octree_export(void *mv, double *Vpos, uint32_t *hex, float *inside) {
	GEO::Mesh *mo = (GEO::Mesh *)mv;
	for (GEO::index_t v = 0; v < mo->vertices.nb(); ++v)
		for (int d = 0; d < 3; ++d) Vpos[3 * v + d] = mo->vertices.point(v)[d];
	GEO::Attribute<float> in(mo->cells.attributes(), "inside");
	for (GEO::index_t c = 0; c < mo->cells.nb(); ++c) {
		for (int lv = 0; lv < 8; ++lv) hex[8 * c + lv] = mo->cells.vertex(c, lv);
		inside[c] = in[c];
	}
}
void ref_compute_octree_free(void *mv) { delete (GEO::Mesh *)mv; }

// ---------------------------------------------------------------------------------------
// VoxelGrid<num_t> + compute_sign, voxelization.h:41-91,220-272
void ref_voxel_dims(const double *extent, double spacing, int padding, int32_t *dims) {
	for (int d = 0; d < 3; ++d) dims[d] = (int)std::ceil(extent[d] / spacing) + 2 * padding;
}
void ref_voxel_sign(const double *V, int64_t nV, const int32_t *F, int64_t nF,
	const double *origin, const double *extent, double spacing, int padding, uint8_t *out)
{
	ensure_init();
	GEO::Mesh M;
	fill_geomesh(M, V, nV, F, nF);
	GEO::MeshFacetsAABB aabb(M);
	VoxelGrid<num_t> voxels(GEO::vec3(origin[0], origin[1], origin[2]), GEO::vec3(extent[0], extent[1], extent[2]), spacing, padding);
	compute_sign(M, aabb, voxels);
	std::memcpy(out, voxels.rawbuf(), (size_t)voxels.num_voxels());
}
// DexelGrid<double> + compute_sign, voxelization.h:95-138,275-331.  Two-phase: offsets (nx*ny+1) then values.
void *ref_dexel_sign(const double *V, int64_t nV, const int32_t *F, int64_t nF,
	const double *origin, const double *extent, double spacing, int padding, int32_t *dims2, int64_t *total)
{
	ensure_init();
	GEO::Mesh M;
	fill_geomesh(M, V, nV, F, nF);
	GEO::MeshFacetsAABB aabb(M);
	DexelGrid<double> *dex = new DexelGrid<double>(GEO::vec3(origin[0], origin[1], origin[2]),
		GEO::vec3(extent[0], extent[1], extent[2]), spacing, padding);
	compute_sign(M, aabb, *dex);
	dims2[0] = dex->grid_size()[0]; dims2[1] = dex->grid_size()[1];
	int64_t t = 0;
	for (int y = 0; y < dims2[1]; ++y) for (int x = 0; x < dims2[0]; ++x) t += (int64_t)dex->at(x, y).size();
	*total = t;
	return dex;
}
void ref_dexel_export(void *dv, int64_t *offsets, double *values) {
	DexelGrid<double> *dex = (DexelGrid<double> *)dv;
	int nx = dex->grid_size()[0], ny = dex->grid_size()[1];
	int64_t t = 0;
	for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) {
		offsets[x + (int64_t)nx * y] = t;
		for (double z : dex->at(x, y)) values[t++] = z;
	}
	offsets[(int64_t)nx * ny] = t;
}
void ref_dexel_free(void *dv) { delete (DexelGrid<double> *)dv; }

// ---------------------------------------------------------------------------------------
// build_aabb_tree (ghm.cpp:4231-4248) == points_inside_mesh set-up (gf.cpp:4024-4042)
void *ref_tree_build(const double *V, int64_t nV, const int32_t *F, int64_t nF) {
	RefTree *t = new RefTree;
	t->V.resize(nV, 3); t->F.resize(nF, 3);
	for (int64_t i = 0; i < nV; ++i) for (int c = 0; c < 3; ++c) t->V(i, c) = V[3 * i + c];
	for (int64_t i = 0; i < nF; ++i) for (int c = 0; c < 3; ++c) t->F(i, c) = F[3 * i + c];
	t->tree.init(t->V, t->F);
	igl::per_face_normals(t->V, t->F, t->FN);
	igl::per_vertex_normals(t->V, t->F, igl::PER_VERTEX_NORMALS_WEIGHTING_TYPE_ANGLE, t->FN, t->VN);
	igl::per_edge_normals(t->V, t->F, igl::PER_EDGE_NORMALS_WEIGHTING_TYPE_UNIFORM, t->FN, t->EN, t->E, t->EMAP);
	return t;
}
void ref_tree_free(void *tv) { delete (RefTree *)tv; }
int64_t ref_tree_num_edges(void *tv) { return ((RefTree *)tv)->E.rows(); }
// FN nF x3, VN nV x3, EN nE x3 (row-major out), E nE x2, EMAP 3*nF
void ref_tree_normals(void *tv, double *FN, double *VN, double *EN, int32_t *E, int32_t *EMAP) {
	RefTree *t = (RefTree *)tv;
	for (int64_t i = 0; i < t->FN.rows(); ++i) for (int c = 0; c < 3; ++c) FN[3 * i + c] = t->FN(i, c);
	for (int64_t i = 0; i < t->VN.rows(); ++i) for (int c = 0; c < 3; ++c) VN[3 * i + c] = t->VN(i, c);
	for (int64_t i = 0; i < t->EN.rows(); ++i) for (int c = 0; c < 3; ++c) EN[3 * i + c] = t->EN(i, c);
	for (int64_t i = 0; i < t->E.rows(); ++i) for (int c = 0; c < 2; ++c) E[2 * i + c] = t->E(i, c);
	for (int64_t i = 0; i < t->EMAP.size(); ++i) EMAP[i] = t->EMAP(i);
}
// pre-order flattening of the pointer tree: per node box (6), primitive, left, right (-1 when leaf)
static int64_t flatten(const igl::AABB<Eigen::MatrixXd, 3> *n, double *box, int32_t *prim, int32_t *lr, int64_t &next) {
	int64_t me = next++;
	if (box) {
		for (int c = 0; c < 3; ++c) { box[6 * me + c] = n->m_box.min()[c]; box[6 * me + 3 + c] = n->m_box.max()[c]; }
		prim[me] = n->m_primitive;
	}
	int64_t l = -1, r = -1;
	if (n->m_left) l = flatten(n->m_left, box, prim, lr, next);
	if (n->m_right) r = flatten(n->m_right, box, prim, lr, next);
	if (lr) { lr[2 * me] = (int32_t)l; lr[2 * me + 1] = (int32_t)r; }
	return me;
}
int64_t ref_tree_num_nodes(void *tv) { int64_t n = 0; flatten(&((RefTree *)tv)->tree, nullptr, nullptr, nullptr, n); return n; }
void ref_tree_flatten(void *tv, double *box, int32_t *prim, int32_t *lr) { int64_t n = 0; flatten(&((RefTree *)tv)->tree, box, prim, lr, n); }

// igl::signed_distance_pseudonormal batch (igl/signed_distance.cpp:186-218). P row-major np x 3.
void ref_signed_distance(void *tv, const double *P, int64_t np, double *S, int32_t *I, double *C, double *N) {
	RefTree *t = (RefTree *)tv;
	Eigen::MatrixXd Pm(np, 3);
	for (int64_t i = 0; i < np; ++i) for (int c = 0; c < 3; ++c) Pm(i, c) = P[3 * i + c];
	Eigen::VectorXd Sv; Eigen::VectorXi Iv; Eigen::MatrixXd Cm, Nm;
	igl::signed_distance_pseudonormal(Pm, t->V, t->F, t->tree, t->FN, t->VN, t->EN, t->EMAP, Sv, Iv, Cm, Nm);
	for (int64_t i = 0; i < np; ++i) {
		S[i] = Sv(i); I[i] = Iv(i);
		for (int c = 0; c < 3; ++c) { C[3 * i + c] = Cm(i, c); N[3 * i + c] = Nm(i, c); }
	}
}
// igl::point_mesh_squared_distance (hausdorff_dis, gf.cpp:3590-3604)
void ref_point_mesh_sqdist(const double *V, int64_t nV, const int32_t *F, int64_t nF, const double *P, int64_t np,
	double *sqrD, int32_t *I, double *C)
{
	Eigen::MatrixXd Vm(nV, 3), Pm(np, 3); Eigen::MatrixXi Fm(nF, 3);
	for (int64_t i = 0; i < nV; ++i) for (int c = 0; c < 3; ++c) Vm(i, c) = V[3 * i + c];
	for (int64_t i = 0; i < nF; ++i) for (int c = 0; c < 3; ++c) Fm(i, c) = F[3 * i + c];
	for (int64_t i = 0; i < np; ++i) for (int c = 0; c < 3; ++c) Pm(i, c) = P[3 * i + c];
	Eigen::VectorXd D; Eigen::VectorXi Iv; Eigen::MatrixXd Cm;
	igl::point_mesh_squared_distance(Pm, Vm, Fm, D, Iv, Cm);
	for (int64_t i = 0; i < np; ++i) { sqrD[i] = D(i); I[i] = Iv(i); for (int c = 0; c < 3; ++c) C[3 * i + c] = Cm(i, c); }
}
// points_inside_mesh, the compiled reference function itself (gf.cpp:4024-4048)
void ref_points_inside_mesh(const double *V, int64_t nV, const int32_t *F, int64_t nF, const double *P, int64_t np, double *S) {
	Mesh tmi; fill_trimesh(tmi, V, nV, F, nF);
	Eigen::MatrixXd Ps(np, 3);
	for (int64_t i = 0; i < np; ++i) for (int c = 0; c < 3; ++c) Ps(i, c) = P[3 * i + c];
	Eigen::VectorXd sd;
	points_inside_mesh(Ps, tmi, sd);
	for (int64_t i = 0; i < np; ++i) S[i] = sd(i);
}

// ---------------------------------------------------------------------------------------
// LINE branch of dirty_graph_projection (ghm.cpp:3967-3994) around the reference's point_line_projection.
// curves in CSR: curve_off[nc+1], curve_vs (ids into V), circle[nc].  P row-major np x 3, curve_id[np].
void ref_polyline_project(const double *V, int64_t nV, const int64_t *curve_off, const int32_t *curve_vs, const uint8_t *circle,
	const double *P, const int32_t *curve_id, int64_t np, double *origin_L, double *axis_L)
{
	(void)nV;
	for (int64_t i = 0; i < np; ++i) {
		Vector3d v(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
		int cid = curve_id[i];
		const int32_t *curve = curve_vs + curve_off[cid];
		uint32_t size = (uint32_t)(curve_off[cid + 1] - curve_off[cid]);
		uint32_t curve_len = size;
		if (!circle[cid]) curve_len--;
		Vector3d tangent(1, 0, 0), pv;
		vector<Vector3d> pvs, tangents;
		vector<pair<double, uint32_t>> dis_ids;
		for (uint32_t j = 0; j < curve_len; j++) {
			uint32_t pos_0 = curve[j], pos_1 = curve[(j + 1) % size];
			double t;
			Vector3d a(V[3 * pos_0], V[3 * pos_0 + 1], V[3 * pos_0 + 2]), b(V[3 * pos_1], V[3 * pos_1 + 1], V[3 * pos_1 + 2]);
			point_line_projection(a, b, v, pv, t);
			tangent = (b - a).normalized();
			dis_ids.push_back(make_pair((v - pv).norm(), (uint32_t)pvs.size()));
			pvs.push_back(pv);
			tangents.push_back(tangent);
		}
		sort(dis_ids.begin(), dis_ids.end());
		if (dis_ids.size()) {
			uint32_t cloestid = dis_ids[0].second;
			pv = pvs[cloestid];
			tangent = tangents[cloestid];
		}
		for (int c = 0; c < 3; ++c) { origin_L[3 * i + c] = pv[c]; axis_L[3 * i + c] = tangent[c]; }
	}
}

// ---------------------------------------------------------------------------------------
// scaled_jacobian (gf.cpp:2309-2358).  V row-major nV x 3 (one vertex per row == column of Mesh.V)
void ref_scaled_jacobian(const double *V, int64_t nV, const uint32_t *hex, int64_t H, double *V_Js, double *H_Js,
	double *min_ave_dev, int64_t *flipped)
{
	Mesh m; m.type = Mesh_type::Hex;
	m.V.resize(3, nV);
	for (int64_t i = 0; i < nV; ++i) for (int c = 0; c < 3; ++c) m.V(c, i) = V[3 * i + c];
	m.Hs.resize(H);
	for (int64_t h = 0; h < H; ++h) { m.Hs[h].id = (uint32_t)h; m.Hs[h].vs.assign(hex + 8 * h, hex + 8 * h + 8); }
	Mesh_Quality mq;
	std::streambuf *old = std::cout.rdbuf(nullptr);
	scaled_jacobian(m, mq);
	std::cout.rdbuf(old);
	int64_t fl = 0;
	for (int64_t i = 0; i < 8 * H; ++i) V_Js[i] = mq.V_Js[i];
	for (int64_t i = 0; i < H; ++i) { H_Js[i] = mq.H_Js[i]; fl += (mq.H_Js[i] < 0); }
	min_ave_dev[0] = mq.min_Jacobian; min_ave_dev[1] = mq.ave_Jacobian; min_ave_dev[2] = mq.deviation_Jacobian;
	*flipped = fl;
}

// ---------------------------------------------------------------------------------------
// build_connectivity, Hex branch (gf.cpp:121-186 + 226-264)
struct RefConn { Mesh m; };
void *ref_hex_connectivity(const uint32_t *hex, int64_t H, int64_t nV, int64_t *nF, int64_t *nE) {
	RefConn *c = new RefConn;
	Mesh &m = c->m; m.type = Mesh_type::Hex;
	m.V.resize(3, nV); m.V.setZero();
	m.Vs.resize(nV);
	for (int64_t i = 0; i < nV; ++i) m.Vs[i].id = (uint32_t)i;
	m.Hs.resize(H);
	for (int64_t h = 0; h < H; ++h) { m.Hs[h].id = (uint32_t)h; m.Hs[h].vs.assign(hex + 8 * h, hex + 8 * h + 8); }
	build_connectivity(m);
	*nF = (int64_t)m.Fs.size(); *nE = (int64_t)m.Es.size();
	return c;
}
// fixed-size relations
void ref_conn_fixed(void *cv, uint32_t *F_vs /*4F*/, uint32_t *F_es /*4F*/, uint8_t *F_boundary, uint32_t *E_vs /*2E*/,
	uint8_t *E_boundary, uint8_t *V_boundary, uint32_t *H_fs /*6H*/)
{
	Mesh &m = ((RefConn *)cv)->m;
	for (size_t f = 0; f < m.Fs.size(); ++f) {
		for (int k = 0; k < 4; ++k) { F_vs[4 * f + k] = m.Fs[f].vs[k]; F_es[4 * f + k] = m.Fs[f].es[k]; }
		F_boundary[f] = m.Fs[f].boundary;
	}
	for (size_t e = 0; e < m.Es.size(); ++e) { E_vs[2 * e] = m.Es[e].vs[0]; E_vs[2 * e + 1] = m.Es[e].vs[1]; E_boundary[e] = m.Es[e].boundary; }
	for (size_t v = 0; v < m.Vs.size(); ++v) V_boundary[v] = m.Vs[v].boundary;
	for (size_t h = 0; h < m.Hs.size(); ++h) for (int k = 0; k < 6; ++k) H_fs[6 * h + k] = m.Hs[h].fs[k];
}
// variable-size relations as CSR; which: 0 F.nhs 1 E.nfs 2 E.nhs 3 V.nvs 4 V.nes 5 V.nfs 6 V.nhs
static const vector<uint32_t> &rel(const Mesh &m, int which, size_t i) {
	switch (which) {
	case 0: return m.Fs[i].neighbor_hs; case 1: return m.Es[i].neighbor_fs; case 2: return m.Es[i].neighbor_hs;
	case 3: return m.Vs[i].neighbor_vs; case 4: return m.Vs[i].neighbor_es; case 5: return m.Vs[i].neighbor_fs;
	default: return m.Vs[i].neighbor_hs;
	}
}
int64_t ref_conn_csr(void *cv, int which, int64_t *off, uint32_t *val) {
	Mesh &m = ((RefConn *)cv)->m;
	size_t n = which == 0 ? m.Fs.size() : (which <= 2 ? m.Es.size() : m.Vs.size());
	int64_t t = 0;
	for (size_t i = 0; i < n; ++i) {
		const vector<uint32_t> &r = rel(m, which, i);
		if (off) off[i] = t;
		if (val) for (uint32_t x : r) val[t++] = x; else t += (int64_t)r.size();
	}
	if (off) off[n] = t;
	return t;
}
void ref_conn_free(void *cv) { delete (RefConn *)cv; }

// ---------------------------------------------------------------------------------------
// metro two-sided Hausdorff: compute(const Mesh&, const Mesh&, double&, double&, double&), metro_hausdorff.cpp:358-505
void ref_hausdorff(const double *VA, int64_t nVA, const int32_t *FA, int64_t nFA,
	const double *VB, int64_t nVB, const int32_t *FB, int64_t nFB, double *out3 /*diag,max,mean*/)
{
	Mesh a, b; fill_trimesh(a, VA, nVA, FA, nFA); fill_trimesh(b, VB, nVB, FB, nFB);
	compute((const Mesh &)a, (const Mesh &)b, out3[0], out3[1], out3[2]);
}
// ratio flavour, metro_hausdorff.cpp:196-357: returns bool, writes ratio
int ref_hausdorff_ratio(const double *VA, int64_t nVA, const int32_t *FA, int64_t nFA,
	const double *VB, int64_t nVB, const int32_t *FB, int64_t nFB, double thr, double *ratio)
{
	Mesh a, b; fill_trimesh(a, VA, nVA, FA, nFA); fill_trimesh(b, VB, nVB, FB, nFB);
	double r = 0, t = thr;
	std::fflush(stdout);
	int ok = compute(a, b, r, t); // prints one line, metro_hausdorff.cpp:188
	*ratio = r;
	return ok;
}

// metro with FACE sampling switched on: the statements of compute() (metro_hausdorff.cpp:30-170) around the real vcg::Sampling
// class, with flags VERTEX | FACE | SIMILAR (the reference clears FACE_SAMPLING at :47-48; this is the mode BASELINE config C5's
// 50 M samples need) and SetSamplesTarget(n_target) per direction.  out = {diag, max_ab, max_ba, mean_ab, mean_ba}; ns = samples.
void ref_hausdorff_face_sampled(const double *VA, int64_t nVA, const int32_t *FA, int64_t nFA,
	const double *VB, int64_t nVB, const int32_t *FB, int64_t nFB, uint64_t n_target_ab, uint64_t n_target_ba, double out[5], uint64_t ns[2])
{
	CMesh S1, S2;
	auto fill = [](CMesh &S, const double *V, int64_t nV, const int32_t *F, int64_t nF) {
		S.vert.resize((size_t)nV);
		for (int64_t i = 0; i < nV; ++i) { CVertex v; v.P()[0] = V[3 * i]; v.P()[1] = V[3 * i + 1]; v.P()[2] = V[3 * i + 2]; S.vert[(size_t)i] = v; }
		S.face.resize((size_t)nF);
		for (int64_t i = 0; i < nF; ++i) {
			CFace f;
			f.V(0) = &(S.vert[(size_t)F[3 * i]]); f.V(1) = &(S.vert[(size_t)F[3 * i + 1]]); f.V(2) = &(S.vert[(size_t)F[3 * i + 2]]);
			S.face[(size_t)i] = f;
		}
		S.vn = (int)S.vert.size(); S.fn = (int)S.face.size();
	};
	fill(S1, VA, nVA, FA, nFA); fill(S2, VB, nVB, FB, nFB);
	int flags = SamplingFlags::VERTEX_SAMPLING | SamplingFlags::FACE_SAMPLING | SamplingFlags::SIMILAR_SAMPLING | SamplingFlags::USE_STATIC_GRID;
	tri::UpdateComponentEP<CMesh>::Set(S1);
	tri::UpdateComponentEP<CMesh>::Set(S2);
	tri::UpdateBounding<CMesh>::Box(S1);
	tri::UpdateBounding<CMesh>::Box(S2);
	Box3<CMesh::ScalarType> bbox;
	bbox.Add(S1.bbox); bbox.Add(S2.bbox);
	bbox.Offset(bbox.Diag() * 0.02);
	S1.bbox = bbox; S2.bbox = bbox;
	Sampling<CMesh> fw(S1, S2), bw(S2, S1);
	fw.SetFlags(flags); fw.SetSamplesTarget((unsigned long)n_target_ab); fw.Hausdorff();
	bw.SetFlags(flags); bw.SetSamplesTarget((unsigned long)n_target_ba); bw.Hausdorff();
	out[0] = bbox.Diag(); out[1] = fw.GetDistMax(); out[2] = bw.GetDistMax(); out[3] = fw.GetDistMean(); out[4] = bw.GetDistMean();
	ns[0] = fw.GetNSamples(); ns[1] = bw.GetNSamples();
}

// hausdorff_dis(mesh0, mesh1, outlierVs, thr) itself, gf.cpp:3590-3628 (igl::point_mesh_squared_distance both ways, threshold
// decaying x0.9 until the list is non-empty).  Two-phase: returns the count, fills `out` when it is large enough.  The list
// comes back in the reference's own push order.
int64_t ref_hausdorff_dis_outliers(const double *VA, int64_t nVA, const int32_t *FA, int64_t nFA,
	const double *VB, int64_t nVB, const int32_t *FB, int64_t nFB, double thr, int32_t *out, int64_t cap)
{
	Mesh a, b; fill_trimesh(a, VA, nVA, FA, nFA); fill_trimesh(b, VB, nVB, FB, nFB);
	std::vector<int> outliers;
	double t = thr;
	std::streambuf *old = std::cout.rdbuf(nullptr);      // "refered total: ..." per round
	hausdorff_dis(a, b, outliers, t);
	std::cout.rdbuf(old);
	if (out && cap >= (int64_t)outliers.size()) for (size_t i = 0; i < outliers.size(); ++i) out[i] = outliers[i];
	return (int64_t)outliers.size();
}

// a further OctreeGrid::subdivide pass over an EXISTING tree with a smaller stop extent — the later passes of the outer loop
// (ghm.cpp:495-500,523-524: the same octree object is subdivided again after args.edge_length_ratio went down)
void ref_octree_subdivide(void *hv, int stop_extent) {
	RefOctree *h = (RefOctree *)hv;
	const GEO::vec3 &o = h->origin; const Eigen::Vector3d &mt = h->mesh_transform; const double vs = h->voxel_size;
	GEO::MeshFacetsAABB &aabb_tree = *h->aabb;
	auto should_subdivide = [&](int x, int y, int z, int extent) {
		if (extent <= stop_extent) return false;
		GEO::Box box;
		box.xyz_min[0] = mt[0] + o[0] + vs * x;
		box.xyz_min[1] = mt[1] + o[1] + vs * y;
		box.xyz_min[2] = mt[2] + o[2] + vs * z;
		box.xyz_max[0] = box.xyz_min[0] + vs * extent;
		box.xyz_max[1] = box.xyz_min[1] + vs * extent;
		box.xyz_max[2] = box.xyz_min[2] + vs * extent;
		bool has_triangles = false;
		auto action = [&has_triangles](int) { has_triangles = true; };
		aabb_tree.compute_bbox_facet_bbox_intersections(box, action);
		return has_triangles;
	};
	h->octree.subdivide(should_subdivide, h->graded, h->paired);
}

} // extern "C"

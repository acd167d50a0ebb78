// TEST INFRASTRUCTURE ONLY — extern "C" driver over the reference's own h_io (io.cpp, compiled UNMODIFIED from /root/reference):
// write_hybrid_mesh_MESH (io.cpp:295-325), write_hybrid_mesh_VTK (:101-181), read / write_feature_Graph_FGRAPH (:412-446).
#include "io.h"
#include <cstdint>

static void fill(Mesh &m, const double *V, int64_t nV, int type, const int64_t *off, const uint32_t *el, int64_t n, int arity, const uint8_t *vb) {
	m.type = (Mesh_type)type;
	m.V.resize(3, nV); m.Vs.resize((size_t)nV);
	for (int64_t i = 0; i < nV; ++i) {
		for (int d = 0; d < 3; ++d) m.V(d, i) = V[3 * i + d];
		m.Vs[(size_t)i].id = (uint32_t)i; m.Vs[(size_t)i].boundary = vb ? vb[i] != 0 : false;
	}
	const bool faces = type == Mesh_type::Tri || type == Mesh_type::Qua || type == Mesh_type::Hyb || type == Mesh_type::HSur;
	for (int64_t i = 0; i < n; ++i) {
		const int64_t b = off ? off[i] : i * (int64_t)arity, e = off ? off[i + 1] : b + arity;
		std::vector<uint32_t> vs(el + b, el + e);
		if (faces) { Hybrid_F f; f.id = (uint32_t)i; f.vs = vs; m.Fs.push_back(f); }
		else { Hybrid h; h.id = (uint32_t)i; h.vs = vs; m.Hs.push_back(h); }
	}
}

extern "C" {
void ref_io_write_mesh(const char *path, const double *V, int64_t nV, int type, const uint32_t *el, int64_t n, int arity) {
	Mesh m; fill(m, V, nV, type, nullptr, el, n, arity, nullptr);
	h_io io; io.write_hybrid_mesh_MESH(m, path);
}
void ref_io_write_vtk(const char *path, const double *V, int64_t nV, int type, const int64_t *off, const uint32_t *el, int64_t n, int arity, const uint8_t *vb) {
	Mesh m; fill(m, V, nV, type, off, el, n, arity, vb);
	h_io io; io.write_hybrid_mesh_VTK(m, path);
}
int ref_io_read_fgraph(const char *path, double *angle, int *oc, int *ocs, int32_t *corners, int64_t *nc, int32_t *pairs, int64_t *np) {
	Mesh_Feature mf;
	h_io io;
	if (!io.read_feature_Graph_FGRAPH(mf, path)) return 0;
	*angle = mf.angle_threshold; *oc = mf.orphan_curve; *ocs = mf.orphan_curve_single;
	if (corners && *nc >= (int64_t)mf.IN_corners.size()) for (size_t i = 0; i < mf.IN_corners.size(); ++i) corners[i] = mf.IN_corners[i];
	if (pairs && *np >= (int64_t)mf.IN_v_pairs.size()) for (size_t i = 0; i < mf.IN_v_pairs.size(); ++i) { pairs[2 * i] = mf.IN_v_pairs[i][0]; pairs[2 * i + 1] = mf.IN_v_pairs[i][1]; }
	*nc = (int64_t)mf.IN_corners.size(); *np = (int64_t)mf.IN_v_pairs.size();
	return 1;
}
}

// Wire formats of the pipeline's outputs and of its feature-graph input:
//   h_io::write_hybrid_mesh_MESH   io.cpp:295-325      h_io::write_hybrid_mesh_VTK   io.cpp:101-181
//   h_io::read_feature_Graph_FGRAPH io.cpp:412-434     h_io::write_feature_Graph_FGRAPH io.cpp:435-446
// The files are ASCII and byte-identical to what the reference's `std::fstream << ...` chains produce: `<<` of a double is
// printf's "%g" with precision 6 (libstdc++ num_put with default flags), of an integer plain decimal, `std::endl` a '\n'.
// The reference formats on one thread through a stream; at the sizes the GPU core produces (10^7 hexes: ~0.8 GB of text) that
// dominates the wall clock of a run, so the rows are formatted by all host threads into per-chunk buffers and written with one
// sequential fwrite per chunk.  Host code only (no CUDA in here): the arrays are the caller's.
#include "../../include/fpohm.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace fpohm { void set_error(const char *fmt, ...); }

namespace {

inline void put_g(std::string &s, double x) { char b[40]; const int n = std::snprintf(b, sizeof b, "%g", x); s.append(b, (size_t)n); }
inline void put_u(std::string &s, unsigned long long x) {
	char b[24]; int n = 0;
	do { b[n++] = (char)('0' + x % 10); x /= 10; } while (x);
	while (n) s.push_back(b[--n]);
}

// rows [0, n) formatted by `fmt(row, out)` on all host threads, written in order
bool write_rows(std::FILE *f, int64_t n, const std::function<void(int64_t, std::string &)> &fmt) {
	if (n <= 0) return true;
	const int T = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::max(1u, std::thread::hardware_concurrency()), n / 4096 + 1));
	std::vector<std::string> part((size_t)T);
	std::vector<std::thread> th;
	for (int t = 0; t < T; ++t)
		th.emplace_back([&, t]() {
			const int64_t lo = n * t / T, hi = n * (t + 1) / T;
			std::string &s = part[(size_t)t];
			s.reserve((size_t)(hi - lo) * 40);
			for (int64_t i = lo; i < hi; ++i) fmt(i, s);
		});
	for (auto &x : th) x.join();
	for (auto &s : part) if (!s.empty() && std::fwrite(s.data(), 1, s.size(), f) != s.size()) return false;
	return true;
}
bool put_str(std::FILE *f, const std::string &s) { return std::fwrite(s.data(), 1, s.size(), f) == s.size(); }

enum { TRI = 0, QUA = 1, HSUR = 2, TET = 3, HYB = 4, HEX = 5 };      // Mesh_type, global_types.h:457-465

} // namespace

extern "C" {

int fpohm_io_write_mesh(const char *path, const double *V, int64_t nV, int32_t mesh_type, const uint32_t *elems, int64_t n_elems) {
	if (!path || (!V && nV) || nV < 0 || n_elems < 0 || (!elems && n_elems)) { fpohm::set_error("fpohm_io_write_mesh: bad argument"); return FPOHM_EINVAL; }
	std::FILE *f = std::fopen(path, "wb");
	if (!f) { fpohm::set_error("fpohm_io_write_mesh: cannot open %s", path); return FPOHM_EINVAL; }
	bool ok = true;
	{ std::string h = "MeshVersionFormatted 1\nDimension 3\nVertices "; put_u(h, (unsigned long long)nV); h += "\n"; ok = put_str(f, h); }
	ok = ok && write_rows(f, nV, [&](int64_t i, std::string &s) { put_g(s, V[3 * i]); s += ' '; put_g(s, V[3 * i + 1]); s += ' '; put_g(s, V[3 * i + 2]); s += " 0\n"; });
	if (mesh_type == TRI || mesh_type == HSUR) {
		std::string h = "Triangles\n"; put_u(h, (unsigned long long)n_elems); h += "\n"; ok = ok && put_str(f, h);
		ok = ok && write_rows(f, n_elems, [&](int64_t i, std::string &s) {
			for (int k = 0; k < 3; ++k) { put_u(s, (unsigned long long)elems[3 * i + k] + 1); s += ' '; }
			s += "0\n";
		});
	} else if (mesh_type == HEX) {
		std::string h = "Hexahedra\n"; put_u(h, (unsigned long long)n_elems); h += "\n"; ok = ok && put_str(f, h);
		ok = ok && write_rows(f, n_elems, [&](int64_t i, std::string &s) {
			for (int k = 0; k < 8; ++k) { put_u(s, (unsigned long long)elems[8 * i + k] + 1); s += ' '; }
			s += "0\n";
		});
	}       // (the reference writes no element block for the other types either)
	ok = ok && put_str(f, "End");
	ok = (std::fclose(f) == 0) && ok;
	if (!ok) { fpohm::set_error("fpohm_io_write_mesh: write to %s failed", path); return FPOHM_ECUDA; }
	return FPOHM_OK;
}

int fpohm_io_write_vtk(const char *path, const double *V, int64_t nV, int32_t mesh_type, const int64_t *elem_off, const uint32_t *elems,
                       int64_t n_elems, int32_t arity, const uint8_t *V_boundary, int64_t n_point_data)
{
	if (!path || (!V && nV) || nV < 0 || n_elems < 0 || (!elems && n_elems) || (mesh_type == HYB && !elem_off && n_elems) || (mesh_type != HYB && arity <= 0 && n_elems)) {
		fpohm::set_error("fpohm_io_write_vtk: bad argument"); return FPOHM_EINVAL;
	}
	std::FILE *f = std::fopen(path, "wb");
	if (!f) { fpohm::set_error("fpohm_io_write_vtk: cannot open %s", path); return FPOHM_EINVAL; }
	bool ok = true;
	{
		std::string h = "# vtk DataFile Version 2.0\nmesh vtk data - converted from .off\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS ";
		put_u(h, (unsigned long long)nV); h += " double\n"; ok = put_str(f, h);
	}
	ok = ok && write_rows(f, nV, [&](int64_t i, std::string &s) { put_g(s, V[3 * i]); s += ' '; put_g(s, V[3 * i + 1]); s += ' '; put_g(s, V[3 * i + 2]); s += '\n'; });
	int cell_type = 12;
	if (mesh_type == TRI) cell_type = 5; else if (mesh_type == QUA) cell_type = 9; else if (mesh_type == HYB) cell_type = 7; else if (mesh_type == TET) cell_type = 10;
	{
		// `uint32_t vnum` in the reference: the size field wraps at 2^32 exactly as there
		// Hyb: `uint32_t vnum` accumulates in the reference (wraps at 2^32 exactly as there); the others are size_t products
		const unsigned long long total = mesh_type == HYB ? (unsigned long long)(uint32_t)((uint64_t)elem_off[n_elems] + (uint64_t)n_elems)
		                                                   : (unsigned long long)n_elems * (unsigned long long)(arity + 1);
		std::string h = "CELLS "; put_u(h, (unsigned long long)n_elems); h += ' '; put_u(h, total); h += "\n"; ok = ok && put_str(f, h);
	}
	ok = ok && write_rows(f, n_elems, [&](int64_t i, std::string &s) {
		const int64_t b = mesh_type == HYB ? elem_off[i] : i * (int64_t)arity, e = mesh_type == HYB ? elem_off[i + 1] : b + arity;
		s += ' '; put_u(s, (unsigned long long)(e - b)); s += ' ';
		for (int64_t k = b; k < e; ++k) { put_u(s, elems[k]); s += ' '; }
		s += '\n';
	});
	{ std::string h = "CELL_TYPES "; put_u(h, (unsigned long long)n_elems); h += "\n"; ok = ok && put_str(f, h); }
	ok = ok && write_rows(f, n_elems, [&](int64_t, std::string &s) { put_u(s, (unsigned long long)cell_type); s += '\n'; });
	{ std::string h = "POINT_DATA "; put_u(h, (unsigned long long)n_point_data); h += "\nSCALARS fixed int\nLOOKUP_TABLE default\n"; ok = ok && put_str(f, h); }
	ok = ok && write_rows(f, n_point_data, [&](int64_t i, std::string &s) { s += (V_boundary && V_boundary[i]) ? "1\n" : "0\n"; });
	ok = (std::fclose(f) == 0) && ok;
	if (!ok) { fpohm::set_error("fpohm_io_write_vtk: write to %s failed", path); return FPOHM_ECUDA; }
	return FPOHM_OK;
}

// two-phase: with corners == NULL / pairs == NULL only the counts and the header come back
int fpohm_io_read_fgraph(const char *path, double *angle_threshold, int32_t *orphan_curve, int32_t *orphan_curve_single,
                         int32_t *corners, int64_t *n_corners, int32_t *pairs, int64_t *n_pairs)
{
	if (!path || !n_corners || !n_pairs) { fpohm::set_error("fpohm_io_read_fgraph: bad argument"); return FPOHM_EINVAL; }
	std::FILE *f = std::fopen(path, "rb");
	if (!f) { fpohm::set_error("fpohm_io_read_fgraph: cannot open %s", path); return FPOHM_EINVAL; }      // the reference returns false
	// f.getline(s, 1023) + sscanf per line, io.cpp:415-431: a line keeps at most 1022 characters
	char s[1024];
	auto getline = [&]() { if (!std::fgets(s, 1023, f)) s[0] = 0; };
	double ang = 0; int oc = 0, ocs = 0, cnum = 0, edgenum = 0;
	getline(); std::sscanf(s, "%lf %i %i", &ang, &oc, &ocs);
	getline(); std::sscanf(s, "%i %i", &cnum, &edgenum);
	if (cnum < 0) cnum = 0;
	if (edgenum < 0) edgenum = 0;
	if (angle_threshold) *angle_threshold = ang;
	if (orphan_curve) *orphan_curve = oc;
	if (orphan_curve_single) *orphan_curve_single = ocs;
	const bool fill = corners || pairs;
	if (fill && (*n_corners < cnum || *n_pairs < edgenum)) {
		std::fclose(f);
		fpohm::set_error("fpohm_io_read_fgraph: capacity %lld / %lld < %d corners / %d edges", (long long)*n_corners, (long long)*n_pairs, cnum, edgenum);
		return FPOHM_EINVAL;
	}
	*n_corners = cnum; *n_pairs = edgenum;
	if (fill) {
		for (int i = 0; i < cnum; ++i) { int c = 0; getline(); std::sscanf(s, "%i", &c); if (corners) corners[i] = c; }      // (a short line keeps the slot's previous value in the reference: 0 of a resized vector)
		for (int i = 0; i < edgenum; ++i) { int v0 = -1, v1 = -1; getline(); std::sscanf(s, "%i %i", &v0, &v1); if (pairs) { pairs[2 * i] = v0; pairs[2 * i + 1] = v1; } }
	}
	std::fclose(f);
	return FPOHM_OK;
}

int fpohm_io_write_fgraph(const char *path, double angle_threshold, int32_t orphan_curve, int32_t orphan_curve_single,
                          const int32_t *corners, int64_t n_corners, const int32_t *pairs, int64_t n_pairs)
{
	if (!path || n_corners < 0 || n_pairs < 0 || (!corners && n_corners) || (!pairs && n_pairs)) { fpohm::set_error("fpohm_io_write_fgraph: bad argument"); return FPOHM_EINVAL; }
	std::FILE *f = std::fopen(path, "wb");
	if (!f) { fpohm::set_error("fpohm_io_write_fgraph: cannot open %s", path); return FPOHM_EINVAL; }
	std::string h;
	put_g(h, angle_threshold); h += ' '; h += std::to_string(orphan_curve); h += ' '; h += std::to_string(orphan_curve_single); h += '\n';
	put_u(h, (unsigned long long)n_corners); h += ' '; put_u(h, (unsigned long long)n_pairs); h += '\n';
	for (int64_t i = 0; i < n_corners; ++i) { h += std::to_string(corners[i]); h += '\n'; }
	for (int64_t i = 0; i < n_pairs; ++i) { h += std::to_string(pairs[2 * i]); h += ' '; h += std::to_string(pairs[2 * i + 1]); h += '\n'; }
	const bool ok = put_str(f, h);
	if (std::fclose(f) != 0 || !ok) { fpohm::set_error("fpohm_io_write_fgraph: write to %s failed", path); return FPOHM_ECUDA; }
	return FPOHM_OK;
}

} // extern "C"

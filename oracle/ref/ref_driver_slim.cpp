// TEST INFRASTRUCTURE ONLY — extern "C" driver over the reference's SLIM per-element stages (slim_m.cpp, compiled UNMODIFIED):
// compute_jacobians (:84-106), update_weights_and_closest_rotations (:108-381, tet branch) and compute_energy_with_jacobians
// (:792-916).  The SLIMData is filled by hand with exactly the members those functions read.
#include "slim_m.h"
#include <igl/flip_avoiding_line_search.h>
#include <cstdint>
#include <vector>

void compute_jacobians(SLIMData &s, const Eigen::MatrixXd &uv);
void update_weights_and_closest_rotations(SLIMData &s, const Eigen::MatrixXd &V, const Eigen::MatrixXi &F, Eigen::MatrixXd &uv);
double compute_energy_with_jacobians(SLIMData &s, const Eigen::MatrixXd &V, const Eigen::MatrixXi &F, const Eigen::MatrixXd &Ji,
                                     Eigen::MatrixXd &uv, Eigen::VectorXd &areas);
void buildRhs(SLIMData &s, const Eigen::SparseMatrix<double> &At);

namespace {
void size_tet_data(SLIMData &s, int64_t n, int energy, double exp_factor) {
	s.dim = 3; s.f_n = (int)n; s.f_num = (int)n;
	s.F.resize(n, 4); s.F.setZero();                 // 4 columns select the tet branch of compute_jacobians
	s.slim_energy = (SLIM_ENERGY)energy; s.exp_factor = exp_factor;
	s.Ji.resize(n, 9); s.Ri.resize(n, 9);
	for (Eigen::VectorXd *w : {&s.W_11, &s.W_12, &s.W_13, &s.W_21, &s.W_22, &s.W_23, &s.W_31, &s.W_32, &s.W_33}) w->resize(n);
}
Eigen::SparseMatrix<double> from_csr(int64_t rows, int64_t cols, const int64_t *off, const int32_t *col, const double *val) {
	std::vector<Eigen::Triplet<double>> t;
	for (int64_t r = 0; r < rows; ++r) for (int64_t k = off[r]; k < off[r + 1]; ++k) t.emplace_back((int)r, col[k], val[k]);
	Eigen::SparseMatrix<double> m((int)rows, (int)cols);
	m.setFromTriplets(t.begin(), t.end());
	m.makeCompressed();
	return m;
}
}

extern "C" {

// Ji = [Dx u, Dy u, Dz u, Dx v, ...] (slim_m.cpp:84-106); D* in CSR (n x nv), uv row-major nv x 3, Ji row-major n x 9
void ref_slim_jacobians(int64_t n, int64_t nv, const int64_t *off, const int32_t *col, const double *vx, const double *vy, const double *vz,
                        const double *uv, double *Ji)
{
	SLIMData s;
	size_tet_data(s, n, 0, 0);
	s.Dx = from_csr(n, nv, off, col, vx); s.Dy = from_csr(n, nv, off, col, vy); s.Dz = from_csr(n, nv, off, col, vz);
	Eigen::MatrixXd U(nv, 3);
	for (int64_t i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) U(i, c) = uv[3 * i + c];
	compute_jacobians(s, U);
	for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) Ji[9 * i + k] = s.Ji(i, k);
}

// update_weights_and_closest_rotations on GIVEN Jacobians: D_c picks entry 3i+c of a stacked vector, so compute_jacobians
// reproduces J exactly (products with 1.0).  W row-major n x 9 = (W_11, W_12, W_13, W_21, ...), Ri n x 9 as s.Ri.
void ref_slim_weights_rotations(const double *J, int64_t n, int energy, double exp_factor, double *W, double *Ri) {
	SLIMData s;
	size_tet_data(s, n, energy, exp_factor);
	std::vector<Eigen::Triplet<double>> tx, ty, tz;
	for (int64_t i = 0; i < n; ++i) { tx.emplace_back((int)i, (int)(3 * i), 1.0); ty.emplace_back((int)i, (int)(3 * i + 1), 1.0); tz.emplace_back((int)i, (int)(3 * i + 2), 1.0); }
	s.Dx.resize((int)n, (int)(3 * n)); s.Dy.resize((int)n, (int)(3 * n)); s.Dz.resize((int)n, (int)(3 * n));
	s.Dx.setFromTriplets(tx.begin(), tx.end()); s.Dy.setFromTriplets(ty.begin(), ty.end()); s.Dz.setFromTriplets(tz.begin(), tz.end());
	Eigen::MatrixXd uv(3 * n, 3);
	for (int64_t i = 0; i < n; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) uv(3 * i + c, r) = J[9 * i + 3 * r + c];
	Eigen::MatrixXd V; Eigen::MatrixXi F;
	update_weights_and_closest_rotations(s, V, F, uv);
	const Eigen::VectorXd *w[9] = {&s.W_11, &s.W_12, &s.W_13, &s.W_21, &s.W_22, &s.W_23, &s.W_31, &s.W_32, &s.W_33};
	for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) { W[9 * i + k] = (*w[k])(i); Ri[9 * i + k] = s.Ri(i, k); }
}

double ref_slim_energy(const double *J, int64_t n, const double *areas, int energy, double exp_factor) {
	SLIMData s;
	size_tet_data(s, n, energy, exp_factor);
	Eigen::MatrixXd Ji(n, 9);
	for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) Ji(i, k) = J[9 * i + k];
	Eigen::VectorXd a(n);
	for (int64_t i = 0; i < n; ++i) a(i) = areas[i];
	Eigen::MatrixXd V, uv; Eigen::MatrixXi F;
	return compute_energy_with_jacobians(s, V, F, Ji, uv, a);
}


// igl::flip_avoiding::compute_max_step_from_singularities (igl/flip_avoiding_line_search.cpp:273-299, tet branch) and the per-tet
// roots of get_min_pos_root_3D (:177-271).  uv, d row-major nv x 3; T row-major n x 4; roots may be NULL.
double ref_slim_max_step(const double *uv, int64_t nv, const int32_t *T, int64_t n, const double *d, double *roots) {
	Eigen::MatrixXd U(nv, 3), D(nv, 3);
	for (int64_t i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) { U(i, c) = uv[3 * i + c]; D(i, c) = d[3 * i + c]; }
	Eigen::MatrixXi F(n, 4);
	for (int64_t i = 0; i < n; ++i) for (int c = 0; c < 4; ++c) F(i, c) = T[4 * i + c];
	if (roots) for (int64_t i = 0; i < n; ++i) roots[i] = igl::flip_avoiding::get_min_pos_root_3D(U, F, D, (int)i);
	return igl::flip_avoiding::compute_max_step_from_singularities(U, F, D);
}


// per-element part of buildRhs (slim_m.cpp:1044-1093): with At = identity, unit weights and proximal_p = 0 the function returns f_rhs itself
void ref_slim_rhs_terms(const double *W, const double *Ri, int64_t n, double *f_rhs) {
	SLIMData s;
	size_tet_data(s, n, 0, 0);
	Eigen::VectorXd *w[9] = {&s.W_11, &s.W_12, &s.W_13, &s.W_21, &s.W_22, &s.W_23, &s.W_31, &s.W_32, &s.W_33};
	for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) { (*w[k])(i) = W[9 * i + k]; s.Ri(i, k) = Ri[9 * i + k]; }
	s.v_n = 1; s.V_o = Eigen::MatrixXd::Zero(1, 3); s.proximal_p = 0;
	s.WGL_M = Eigen::VectorXd::Ones(9 * n);
	Eigen::SparseMatrix<double> At((int)(9 * n), (int)(9 * n));
	At.setIdentity();
	// rhs has dim * v_n + ... entries in the real pipeline; here At is square, so rhs = f_rhs + 0 * uv_flat needs equal sizes
	s.v_n = (int)(3 * n); s.V_o = Eigen::MatrixXd::Zero(3 * n, 3);
	buildRhs(s, At);
	for (int64_t i = 0; i < 9 * n; ++i) f_rhs[i] = s.rhs(i);
}

}

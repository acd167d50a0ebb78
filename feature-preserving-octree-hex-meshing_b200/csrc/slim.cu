// SLIM per-element stages, tet branch: compute_jacobians (slim_m.cpp:84-106),
// update_weights_and_closest_rotations (:229-381) and compute_energy_with_jacobians (:861-913).  These sit between the
// projection and the Jacobian kernels in the optimisation inner loop (8 tets per hex); the sparse solve between them is not
// part of this row.
//
// The reference decomposes every 3x3 Jacobian with igl::polar_svd (Eigen::JacobiSVD, igl/polar_svd.cpp:34-67).  Here: one
// thread per tet, one-sided Jacobi (Hestenes) rotations on the columns until they are orthogonal — relative accuracy on
// every singular value, no A^T A — then singular values sorted descending as Eigen returns them.  Everything the stages
// emit is invariant to the sign / ordering freedom of the factors: W = U diag(m) U^T, R = U V^T (last column of V negated
// under a reflection), conformal "rotations" closest * U V^T.  Floating point, tolerance 1e-5 relative (north star);
// observed agreement with the reference ~1e-13.
#include "internal.h"

#include <cmath>
#include <math_constants.h>
#include <vector>

using namespace fpohm;

namespace {

enum { E_ARAP = 0, E_LOG_ARAP = 1, E_SYMMETRIC_DIRICHLET = 2, E_CONFORMAL = 3, E_EXP_CONFORMAL = 4, E_EXP_SYMMETRIC_DIRICHLET = 5 };   // global_types.h:56-64

struct Svd3 { double U[3][3], V[3][3], s[3]; };

__device__ __forceinline__ void rot_cols(double M[3][3], int p, int q, double c, double s) {
#pragma unroll
	for (int r = 0; r < 3; ++r) { const double a = M[r][p], b = M[r][q]; M[r][p] = c * a - s * b; M[r][q] = s * a + c * b; }
}

// A = U diag(s) V^T, s[0] >= s[1] >= s[2] >= 0.  VECTORS = false: singular values only (the energy needs nothing else)
template <bool VECTORS>
__device__ void svd3(const double A[3][3], Svd3 &o) {
	double B[3][3];
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int c = 0; c < 3; ++c) { B[r][c] = A[r][c]; o.V[r][c] = r == c ? 1.0 : 0.0; }
	for (int sweep = 0; sweep < 30; ++sweep) {
		bool rotated = false;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const int p = k == 2 ? 1 : 0, q = k == 0 ? 1 : 2;
			const double alpha = B[0][p] * B[0][p] + B[1][p] * B[1][p] + B[2][p] * B[2][p];
			const double beta = B[0][q] * B[0][q] + B[1][q] * B[1][q] + B[2][q] * B[2][q];
			const double gamma = B[0][p] * B[0][q] + B[1][p] * B[1][q] + B[2][p] * B[2][q];
			if (gamma == 0.0 || gamma * gamma <= 1e-30 * (alpha * beta)) continue;       // |cos| of the column angle <= 1e-15
			rotated = true;
			// smaller root of t^2 + 2 zeta t - 1 = 0, zeta = (beta - alpha) / (2 gamma), written with one division
			const double dba = beta - alpha;
			const double t = (dba >= 0 ? 2.0 * gamma : -2.0 * gamma) / (fabs(dba) + sqrt(dba * dba + 4.0 * gamma * gamma));
			const double c = rsqrt(1.0 + t * t), s = c * t;
			rot_cols(B, p, q, c, s);
			if (VECTORS) rot_cols(o.V, p, q, c, s);
		}
		if (!rotated) break;
	}
	double n[3];
#pragma unroll
	for (int c = 0; c < 3; ++c) n[c] = sqrt(B[0][c] * B[0][c] + B[1][c] * B[1][c] + B[2][c] * B[2][c]);
	// descending order: three compare-swaps on whole columns (static indices keep everything in registers)
#define FPOHM_COLSWAP(a, b)                                                                               \
	if (n[a] < n[b]) {                                                                                     \
		double t_ = n[a]; n[a] = n[b]; n[b] = t_;                                                          \
		for (int r = 0; r < 3; ++r) { t_ = B[r][a]; B[r][a] = B[r][b]; B[r][b] = t_; t_ = o.V[r][a]; o.V[r][a] = o.V[r][b]; o.V[r][b] = t_; } \
	}
	if (!VECTORS) {
		double t_;
		if (n[0] < n[1]) { t_ = n[0]; n[0] = n[1]; n[1] = t_; }
		if (n[1] < n[2]) { t_ = n[1]; n[1] = n[2]; n[2] = t_; }
		if (n[0] < n[1]) { t_ = n[0]; n[0] = n[1]; n[1] = t_; }
		o.s[0] = n[0]; o.s[1] = n[1]; o.s[2] = n[2];
		return;
	}
	FPOHM_COLSWAP(0, 1) FPOHM_COLSWAP(1, 2) FPOHM_COLSWAP(0, 1)
#undef FPOHM_COLSWAP
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		o.s[c] = n[c];
		const double inv = n[c] > 0 ? 1.0 / n[c] : 0.0;
#pragma unroll
		for (int r = 0; r < 3; ++r) o.U[r][c] = B[r][c] * inv;
	}
	// rank-deficient input: complete U to an orthonormal basis (the reference's factors are just as arbitrary there)
	if (o.s[2] <= 0) {
		if (o.s[1] <= 0) {
			if (o.s[0] <= 0) { o.U[0][0] = 1; o.U[1][0] = 0; o.U[2][0] = 0; }
			const int m = fabs(o.U[0][0]) <= fabs(o.U[1][0]) && fabs(o.U[0][0]) <= fabs(o.U[2][0]) ? 0 : (fabs(o.U[1][0]) <= fabs(o.U[2][0]) ? 1 : 2);
			double e[3] = {0, 0, 0}; e[m] = 1;
			const double d = e[0] * o.U[0][0] + e[1] * o.U[1][0] + e[2] * o.U[2][0];
			double v[3] = {e[0] - d * o.U[0][0], e[1] - d * o.U[1][0], e[2] - d * o.U[2][0]};
			const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
			for (int r = 0; r < 3; ++r) o.U[r][1] = v[r] / l;
		}
		o.U[0][2] = o.U[1][0] * o.U[2][1] - o.U[2][0] * o.U[1][1];
		o.U[1][2] = o.U[2][0] * o.U[0][1] - o.U[0][0] * o.U[2][1];
		o.U[2][2] = o.U[0][0] * o.U[1][1] - o.U[1][0] * o.U[0][1];
	}
}

__device__ __forceinline__ double det3(const double M[3][3]) {
	return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

__device__ __forceinline__ void load_ji(const double *__restrict__ Ji, int64_t i, double A[3][3]) {
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int c = 0; c < 3; ++c) A[r][c] = Ji[9 * i + 3 * r + c];             // ji(r, c) = Ji(i, 3r + c), slim_m.cpp:239-247
}

// update_weights_and_closest_rotations, slim_m.cpp:236-378
__global__ void __launch_bounds__(128)
slim_weights_kernel(const double *__restrict__ Ji, int64_t n, int energy, double exp_f, double *__restrict__ W, double *__restrict__ Ri) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		double A[3][3];
		load_ji(Ji, i, A);
		Svd3 f;
		svd3<true>(A, f);
		const double s1 = f.s[0], s2 = f.s[1], s3 = f.s[2];
		// ri = U V^T, last column of V negated under a reflection (igl/polar_svd.cpp:53-62)
		double R[3][3];
#pragma unroll
		for (int r = 0; r < 3; ++r)
#pragma unroll
			for (int c = 0; c < 3; ++c) R[r][c] = f.U[r][0] * f.V[c][0] + f.U[r][1] * f.V[c][1] + f.U[r][2] * f.V[c][2];
		if (det3(R) < 0) {
#pragma unroll
			for (int r = 0; r < 3; ++r)
#pragma unroll
				for (int c = 0; c < 3; ++c) R[r][c] = f.U[r][0] * f.V[c][0] + f.U[r][1] * f.V[c][1] - f.U[r][2] * f.V[c][2];
		}
		double m[3] = {1, 1, 1};
		switch (energy) {
		case E_ARAP: break;
		case E_LOG_ARAP: {
			const double g1 = 2 * (log(s1) / s1), g2 = 2 * (log(s2) / s2), g3 = 2 * (log(s3) / s3);
			m[0] = sqrt(g1 / (2 * (s1 - 1))); m[1] = sqrt(g2 / (2 * (s2 - 1))); m[2] = sqrt(g3 / (2 * (s3 - 1)));
			break;
		}
		case E_SYMMETRIC_DIRICHLET:
		case E_EXP_SYMMETRIC_DIRICHLET: {
			double g1 = 2 * (s1 - pow(s1, -3.0)), g2 = 2 * (s2 - pow(s2, -3.0)), g3 = 2 * (s3 - pow(s3, -3.0));
			if (energy == E_EXP_SYMMETRIC_DIRICHLET) {
				const double in_exp = exp_f * (pow(s1, 2.0) + pow(s1, -2.0) + pow(s2, 2.0) + pow(s2, -2.0) + pow(s3, 2.0) + pow(s3, -2.0));
				const double e = exp(in_exp);
				g1 *= e * exp_f; g2 *= e * exp_f; g3 *= e * exp_f;
			}
			m[0] = sqrt(g1 / (2 * (s1 - 1))); m[1] = sqrt(g2 / (2 * (s2 - 1))); m[2] = sqrt(g3 / (2 * (s3 - 1)));
			break;
		}
		default: {      // CONFORMAL, EXP_CONFORMAL
			const double cd = 9 * pow(s1 * s2 * s3, 5. / 3.);
			double g1 = (-2 * s2 * s3 * (pow(s2, 2.0) + pow(s3, 2.0) - 2 * pow(s1, 2.0))) / cd;
			double g2 = (-2 * s1 * s3 * (pow(s1, 2.0) + pow(s3, 2.0) - 2 * pow(s2, 2.0))) / cd;
			double g3 = (-2 * s1 * s2 * (pow(s1, 2.0) + pow(s2, 2.0) - 2 * pow(s3, 2.0))) / cd;
			if (energy == E_EXP_CONFORMAL) {
				const double in_exp = exp_f * ((pow(s1, 2.0) + pow(s2, 2.0) + pow(s3, 2.0)) / (3 * pow(s1 * s2 * s3, 2. / 3)));
				const double e = exp(in_exp);
				g1 *= e * exp_f; g2 *= e * exp_f; g3 *= e * exp_f;
			}
			const double closest = sqrt(pow(s1, 2.0) + pow(s3, 2.0)) / sqrt(2.0);
			m[0] = sqrt(g1 / (2 * (s1 - closest))); m[1] = sqrt(g2 / (2 * (s2 - closest))); m[2] = sqrt(g3 / (2 * (s3 - closest)));
			// "change local step": ri = ui * diag(closest) * vi^T with the UNmodified vi (slim_m.cpp:320-322)
#pragma unroll
			for (int r = 0; r < 3; ++r)
#pragma unroll
				for (int c = 0; c < 3; ++c) R[r][c] = closest * (f.U[r][0] * f.V[c][0] + f.U[r][1] * f.V[c][1] + f.U[r][2] * f.V[c][2]);
			break;
		}
		}
		const double eps = 1e-8;
		if (fabs(s1 - 1) < eps) m[0] = 1;
		if (fabs(s2 - 1) < eps) m[1] = 1;
		if (fabs(s3 - 1) < eps) m[2] = 1;
#pragma unroll
		for (int r = 0; r < 3; ++r)
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				W[9 * i + 3 * r + c] = f.U[r][0] * m[0] * f.U[c][0] + f.U[r][1] * m[1] * f.U[c][1] + f.U[r][2] * m[2] * f.U[c][2];   // W_{r+1,c+1}
				Ri[9 * i + 3 * c + r] = R[r][c];                                                                                    // s.Ri(i, 3c + r) = ri(r, c)
			}
	}
}

__device__ __forceinline__ double element_energy(int energy, double exp_f, double s1, double s2, double s3) {
	switch (energy) {
	case E_ARAP: return pow(s1 - 1, 2.0) + pow(s2 - 1, 2.0) + pow(s3 - 1, 2.0);
	case E_SYMMETRIC_DIRICHLET: return pow(s1, 2.0) + pow(s1, -2.0) + pow(s2, 2.0) + pow(s2, -2.0) + pow(s3, 2.0) + pow(s3, -2.0);
	case E_EXP_SYMMETRIC_DIRICHLET: return exp(exp_f * (pow(s1, 2.0) + pow(s1, -2.0) + pow(s2, 2.0) + pow(s2, -2.0) + pow(s3, 2.0) + pow(s3, -2.0)));
	case E_LOG_ARAP: return pow(log(s1), 2.0) + pow(log(fabs(s2)), 2.0) + pow(log(fabs(s3)), 2.0);
	case E_CONFORMAL: return (pow(s1, 2.0) + pow(s2, 2.0) + pow(s3, 2.0)) / (3 * pow(s1 * s2 * s3, 2. / 3.));
	default: return exp((pow(s1, 2.0) + pow(s2, 2.0) + pow(s3, 2.0)) / (3 * pow(s1 * s2 * s3, 2. / 3.)));      // EXP_CONFORMAL: no exp_factor, slim_m.cpp:907
	}
}

// compute_energy_with_jacobians, slim_m.cpp:861-913: sum_i areas(i) * f(singular values); fixed-order block partials
__global__ void __launch_bounds__(256)
slim_energy_kernel(const double *__restrict__ Ji, int64_t n, const double *__restrict__ areas, int energy, double exp_f, double *__restrict__ partial) {
	__shared__ double sm[256];
	double acc = 0;
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		double A[3][3];
		load_ji(Ji, i, A);
		Svd3 f;
		svd3<false>(A, f);
		acc += areas[i] * element_energy(energy, exp_f, f.s[0], f.s[1], f.s[2]);
	}
	sm[threadIdx.x] = acc;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void slim_energy_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out) {
	if (blockIdx.x == 0 && threadIdx.x == 0) { double s = 0; for (int b = 0; b < nb; ++b) s += partial[b]; *out = s; }
}

// compute_jacobians, slim_m.cpp:94-106: row i of Dx, Dy, Dz (one shared CSR pattern) against the three columns of uv
__global__ void __launch_bounds__(256)
slim_jacobians_kernel(int64_t n, const int64_t *__restrict__ off, const int32_t *__restrict__ col, const double *__restrict__ vx,
                      const double *__restrict__ vy, const double *__restrict__ vz, const double *__restrict__ uv, double *__restrict__ Ji)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
		for (int64_t k = off[i]; k < off[i + 1]; ++k) {
			const double *u = uv + 3 * (int64_t)col[k];
			const double dx = vx[k], dy = vy[k], dz = vz[k];
#pragma unroll
			for (int r = 0; r < 3; ++r) { a[3 * r] += dx * u[r]; a[3 * r + 1] += dy * u[r]; a[3 * r + 2] += dz * u[r]; }
		}
#pragma unroll
		for (int k = 0; k < 9; ++k) Ji[9 * i + k] = a[k];
	}
}

// ---- flip-avoiding line search: largest step before some tet's volume changes sign
// (igl::flip_avoiding::compute_max_step_from_singularities, igl/flip_avoiding_line_search.cpp:177-299, tet branch).
// The volume along the search direction is the cubic det([b-a; c-a; d-a] + t [db-da; dc-da; dd-da]); its coefficients are the
// row-wise multilinear expansion of that determinant — the same polynomial as the reference's symbolic expansion of the 4x4
// volume determinant up to a common sign, on which the root selection does not depend.
__device__ __forceinline__ double det_rows(const double *x, const double *y, const double *z) {
	return x[0] * (y[1] * z[2] - y[2] * z[1]) - x[1] * (y[0] * z[2] - y[2] * z[0]) + x[2] * (y[0] * z[1] - y[1] * z[0]);
}
// smallest positive root of a t^2 + b t + c (igl/flip_avoiding_line_search.cpp:63-109)
__device__ __forceinline__ double smallest_pos_quad_zero(double a, double b, double c) {
	const double inf = CUDART_INF;
	if (fabs(a) > 1.0e-10) {
		const double delta_in = b * b - 4 * a * c;
		if (delta_in <= 0) return inf;
		const double delta = sqrt(delta_in);
		double t1, t2;
		if (b >= 0) { const double bd = -b - delta; t1 = 2 * c / bd; t2 = bd / (2 * a); }
		else { const double bd = -b + delta; t1 = bd / (2 * a); t2 = (2 * c) / bd; }
		if (a < 0) { const double t = t1; t1 = t2; t2 = t; }
		if (t1 > 0) return t2 > 0 ? t2 : t1;
		return inf;
	}
	if (b == 0) return inf;
	const double t1 = -c / b;
	return t1 > 0 ? t1 : inf;
}
// smallest positive root of t^3 + a t^2 + b t + c: trigonometric form for three real roots, Cardano otherwise (:24-61, :249-270)
__device__ __forceinline__ double smallest_pos_cubic_zero(double a, double b, double c) {
	const double inf = CUDART_INF;
	const double a2 = a * a;
	double q = (a2 - 3 * b) / 9;
	const double r = (a * (2 * a2 - 9 * b) + 27 * c) / 54;
	const double r2 = r * r, q3 = q * q * q;
	if (r2 < q3) {
		double t = r / sqrt(q3);
		t = acos(fmin(1.0, fmax(-1.0, t)));
		a /= 3; q = -2 * sqrt(q);
		double x0 = q * cos(t / 3) - a, x1 = q * cos((t + 2 * CUDART_PI) / 3) - a, x2 = q * cos((t - 2 * CUDART_PI) / 3) - a;
		double lo = x0, mid = x1, hi = x2, sw;                 // std::sort of the three roots
		if (lo > mid) { sw = lo; lo = mid; mid = sw; }
		if (mid > hi) { sw = mid; mid = hi; hi = sw; }
		if (lo > mid) { sw = lo; lo = mid; mid = sw; }
		if (lo > 0) return lo;
		if (mid > 0) return mid;
		if (hi > 0) return hi;
		return inf;
	}
	double A = -pow(fabs(r) + sqrt(r2 - q3), 1. / 3);
	if (r < 0) A = -A;
	const double B = A == 0 ? 0 : q / A;
	a /= 3;
	const double x0 = (A + B) - a, x1 = -0.5 * (A + B) - a, x2 = 0.5 * sqrt(3.) * (A - B);
	if (fabs(x2) < 1e-14) {                            // double real root
		const double lo = fmin(x0, x1), hi = fmax(x0, x1);
		if (lo > 0) return lo;
		if (hi > 0) return hi;
		return inf;
	}
	return x0 >= 0 ? x0 : inf;                         // one real root
}
__global__ void __launch_bounds__(256)
slim_max_step_kernel(const double *__restrict__ uv, const int32_t *__restrict__ T, int64_t n, const double *__restrict__ d,
                     double *__restrict__ roots, double *__restrict__ partial)
{
	__shared__ double sm[256];
	double best = CUDART_INF;
	for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < n; f += (int64_t)gridDim.x * blockDim.x) {
		const int4 t = *reinterpret_cast<const int4 *>(T + 4 * f);
		const int v[4] = {t.x, t.y, t.z, t.w};
		double P[3][3], D[3][3];
#pragma unroll
		for (int k = 0; k < 3; ++k)
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				P[k][c] = uv[3 * (int64_t)v[k + 1] + c] - uv[3 * (int64_t)v[0] + c];
				D[k][c] = d[3 * (int64_t)v[k + 1] + c] - d[3 * (int64_t)v[0] + c];
			}
		const double c0 = det_rows(P[0], P[1], P[2]);
		const double c1 = det_rows(D[0], P[1], P[2]) + det_rows(P[0], D[1], P[2]) + det_rows(P[0], P[1], D[2]);
		const double c2 = det_rows(D[0], D[1], P[2]) + det_rows(D[0], P[1], D[2]) + det_rows(P[0], D[1], D[2]);
		const double c3 = det_rows(D[0], D[1], D[2]);
		const double root = fabs(c3) <= 1.e-10 ? smallest_pos_quad_zero(c2, c1, c0) : smallest_pos_cubic_zero(c2 / c3, c1 / c3, c0 / c3);
		if (roots) roots[f] = root;
		best = fmin(best, root);
	}
	sm[threadIdx.x] = best;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] = fmin(sm[threadIdx.x], sm[threadIdx.x + o]); __syncthreads(); }
	if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void slim_min_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out) {
	if (blockIdx.x == 0 && threadIdx.x == 0) { double m = CUDART_INF; for (int b = 0; b < nb; ++b) m = fmin(m, partial[b]); *out = m; }
}

// per-element part of buildRhs (slim_m.cpp:1061-1083): f_rhs(i + (3a + b) n) = sum_k W_{a+1,k+1}(i) * ri(k, b)(i), evaluated left to right
__global__ void __launch_bounds__(256)
slim_rhs_terms_kernel(const double *__restrict__ W, const double *__restrict__ Ri, int64_t n, double *__restrict__ f_rhs) {
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		double w[9], r[9];
#pragma unroll
		for (int k = 0; k < 9; ++k) { w[k] = W[9 * i + k]; r[k] = Ri[9 * i + k]; }
#pragma unroll
		for (int a = 0; a < 3; ++a)
#pragma unroll
			for (int b = 0; b < 3; ++b)
				f_rhs[i + (int64_t)(3 * a + b) * n] = (w[3 * a] * r[3 * b] + w[3 * a + 1] * r[3 * b + 1]) + w[3 * a + 2] * r[3 * b + 2];
	}
}

void check_energy(int32_t energy, const char *who) {
	FPOHM_REQUIRE(energy >= 0 && energy <= 5, FPOHM_EINVAL, "%s: unknown SLIM_ENERGY %d", who, energy);
}

} // namespace

extern "C" {

int fpohm_slim_jacobians_dev(fpohm_ctx *ctx, int64_t n, const int64_t *off_dev, const int32_t *col_dev, const double *vx_dev, const double *vy_dev,
                             const double *vz_dev, const double *uv_dev, double *Ji_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n >= 0 && (n == 0 || (off_dev && col_dev && vx_dev && vy_dev && vz_dev && uv_dev && Ji_dev)), FPOHM_EINVAL, "fpohm_slim_jacobians_dev: bad argument");
	if (n == 0) return FPOHM_OK;
	DeviceGuard g(ctx->device);
	slim_jacobians_kernel<<<grid_for(ctx, n, 256), 256, 0, (cudaStream_t)stream>>>(n, off_dev, col_dev, vx_dev, vy_dev, vz_dev, uv_dev, Ji_dev);
	FPOHM_LAUNCH_CHECK(ctx);
	FPOHM_API_END
}

int fpohm_slim_weights_rotations_dev(fpohm_ctx *ctx, const double *Ji_dev, int64_t n, int32_t energy, double exp_factor, double *W_dev, double *Ri_dev,
                                     void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n >= 0 && (n == 0 || (Ji_dev && W_dev && Ri_dev)), FPOHM_EINVAL, "fpohm_slim_weights_rotations_dev: bad argument");
	check_energy(energy, "fpohm_slim_weights_rotations_dev");
	if (n == 0) return FPOHM_OK;
	DeviceGuard g(ctx->device);
	slim_weights_kernel<<<grid_for(ctx, n, 128, 8), 128, 0, (cudaStream_t)stream>>>(Ji_dev, n, energy, exp_factor, W_dev, Ri_dev);
	FPOHM_LAUNCH_CHECK(ctx);
	FPOHM_API_END
}

int fpohm_slim_energy_dev(fpohm_ctx *ctx, const double *Ji_dev, int64_t n, const double *areas_dev, int32_t energy, double exp_factor,
                          double *energy_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n >= 0 && energy_dev && (n == 0 || (Ji_dev && areas_dev)), FPOHM_EINVAL, "fpohm_slim_energy_dev: bad argument");
	check_energy(energy, "fpohm_slim_energy_dev");
	DeviceGuard g(ctx->device);
	cudaStream_t s = (cudaStream_t)stream;
	const int nb = grid_for(ctx, n, 256, 4);
	DevBuf<double> partial(nb, s);
	slim_energy_kernel<<<nb, 256, 0, s>>>(Ji_dev, n, areas_dev, energy, exp_factor, partial.p);
	FPOHM_LAUNCH_CHECK(ctx);
	slim_energy_final_kernel<<<1, 32, 0, s>>>(partial.p, nb, energy_dev);
	FPOHM_LAUNCH_CHECK(ctx);
	FPOHM_API_END
}

int fpohm_slim_max_step_dev(fpohm_ctx *ctx, const double *uv_dev, const int32_t *T_dev, int64_t n, const double *d_dev, double *roots_dev,
                            double *max_step_dev, void *stream)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n >= 0 && max_step_dev && (n == 0 || (uv_dev && T_dev && d_dev)), FPOHM_EINVAL, "fpohm_slim_max_step_dev: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = (cudaStream_t)stream;
	const int nb = grid_for(ctx, n, 256, 4);
	DevBuf<double> partial(nb, s);
	slim_max_step_kernel<<<nb, 256, 0, s>>>(uv_dev, T_dev, n, d_dev, roots_dev, partial.p);
	FPOHM_LAUNCH_CHECK(ctx);
	slim_min_final_kernel<<<1, 32, 0, s>>>(partial.p, nb, max_step_dev);
	FPOHM_LAUNCH_CHECK(ctx);
	FPOHM_API_END
}

int fpohm_slim_max_step(fpohm_ctx *ctx, const double *uv, int64_t nv, const int32_t *T, int64_t n, const double *d, double *roots, double *max_step) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && uv && T && d && max_step && nv > 0 && n > 0, FPOHM_EINVAL, "fpohm_slim_max_step: bad argument");
	for (int64_t i = 0; i < 4 * n; ++i) FPOHM_REQUIRE(T[i] >= 0 && T[i] < nv, FPOHM_EINVAL, "fpohm_slim_max_step: vertex id %d out of range at %lld", T[i], (long long)i);
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> duv(3 * nv, s), dd(3 * nv, s), dr(roots ? n : 0, s), dm(1, s);
	DevBuf<int32_t> dT(4 * n, s);
	duv.upload(uv, 3 * nv); dd.upload(d, 3 * nv); dT.upload(T, 4 * n);
	KernelTimer t(ctx, s);
	int rc = fpohm_slim_max_step_dev(ctx, duv.p, dT.p, n, dd.p, roots ? dr.p : nullptr, dm.p, s);
	t.stop();
	if (rc != FPOHM_OK) return rc;
	if (roots) dr.download(roots, n);
	dm.download(max_step, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_slim_rhs_terms_dev(fpohm_ctx *ctx, const double *W_dev, const double *Ri_dev, int64_t n, double *f_rhs_dev, void *stream) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n >= 0 && (n == 0 || (W_dev && Ri_dev && f_rhs_dev)), FPOHM_EINVAL, "fpohm_slim_rhs_terms_dev: bad argument");
	if (n == 0) return FPOHM_OK;
	DeviceGuard g(ctx->device);
	slim_rhs_terms_kernel<<<grid_for(ctx, n, 256), 256, 0, (cudaStream_t)stream>>>(W_dev, Ri_dev, n, f_rhs_dev);
	FPOHM_LAUNCH_CHECK(ctx);
	FPOHM_API_END
}

int fpohm_slim_rhs_terms(fpohm_ctx *ctx, const double *W, const double *Ri, int64_t n, double *f_rhs) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n > 0 && W && Ri && f_rhs, FPOHM_EINVAL, "fpohm_slim_rhs_terms: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dW(9 * n, s), dR(9 * n, s), dF(9 * n, s);
	dW.upload(W, 9 * n); dR.upload(Ri, 9 * n);
	int rc = fpohm_slim_rhs_terms_dev(ctx, dW.p, dR.p, n, dF.p, s);
	if (rc != FPOHM_OK) return rc;
	dF.download(f_rhs, 9 * n);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_slim_jacobians(fpohm_ctx *ctx, int64_t n, int64_t nv, const int64_t *off, const int32_t *col, const double *vx, const double *vy,
                         const double *vz, const double *uv, double *Ji)
{
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n > 0 && nv > 0 && off && col && vx && vy && vz && uv && Ji, FPOHM_EINVAL, "fpohm_slim_jacobians: bad argument");
	const int64_t nnz = off[n];
	FPOHM_REQUIRE(off[0] == 0 && nnz >= 0, FPOHM_EINVAL, "fpohm_slim_jacobians: bad row offsets");
	for (int64_t i = 0; i < n; ++i) FPOHM_REQUIRE(off[i] <= off[i + 1], FPOHM_EINVAL, "fpohm_slim_jacobians: row offsets decrease at %lld", (long long)i);
	for (int64_t k = 0; k < nnz; ++k) FPOHM_REQUIRE(col[k] >= 0 && col[k] < nv, FPOHM_EINVAL, "fpohm_slim_jacobians: column %d out of range at %lld", col[k], (long long)k);
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<int64_t> doff(n + 1, s); DevBuf<int32_t> dcol(nnz, s);
	DevBuf<double> dx(nnz, s), dy(nnz, s), dz(nnz, s), duv(3 * nv, s), dJ(9 * n, s);
	doff.upload(off, n + 1); dcol.upload(col, nnz); dx.upload(vx, nnz); dy.upload(vy, nnz); dz.upload(vz, nnz); duv.upload(uv, 3 * nv);
	KernelTimer t(ctx, s);
	int rc = fpohm_slim_jacobians_dev(ctx, n, doff.p, dcol.p, dx.p, dy.p, dz.p, duv.p, dJ.p, s);
	t.stop();
	if (rc != FPOHM_OK) return rc;
	dJ.download(Ji, 9 * n);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_slim_weights_rotations(fpohm_ctx *ctx, const double *Ji, int64_t n, int32_t energy, double exp_factor, double *W, double *Ri) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n > 0 && Ji && W && Ri, FPOHM_EINVAL, "fpohm_slim_weights_rotations: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dJ(9 * n, s), dW(9 * n, s), dR(9 * n, s);
	dJ.upload(Ji, 9 * n);
	KernelTimer t(ctx, s);
	int rc = fpohm_slim_weights_rotations_dev(ctx, dJ.p, n, energy, exp_factor, dW.p, dR.p, s);
	t.stop();
	if (rc != FPOHM_OK) return rc;
	dW.download(W, 9 * n); dR.download(Ri, 9 * n);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

int fpohm_slim_energy(fpohm_ctx *ctx, const double *Ji, int64_t n, const double *areas, int32_t energy, double exp_factor, double *energy_out) {
	FPOHM_API_BEGIN
	FPOHM_REQUIRE(ctx && n > 0 && Ji && areas && energy_out, FPOHM_EINVAL, "fpohm_slim_energy: bad argument");
	DeviceGuard g(ctx->device);
	cudaStream_t s = ctx->stream;
	DevBuf<double> dJ(9 * n, s), dA(n, s), dE(1, s);
	dJ.upload(Ji, 9 * n); dA.upload(areas, n);
	KernelTimer t(ctx, s);
	int rc = fpohm_slim_energy_dev(ctx, dJ.p, n, dA.p, energy, exp_factor, dE.p, s);
	t.stop();
	if (rc != FPOHM_OK) return rc;
	dE.download(energy_out, 1);
	FPOHM_CUDA(cudaStreamSynchronize(s));
	FPOHM_API_END
}

} // extern "C"

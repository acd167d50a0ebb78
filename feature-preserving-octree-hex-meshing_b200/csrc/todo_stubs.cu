// Entry points declared in include/fpohm.h whose kernels are not written yet.  They fail loudly.
#include "internal.h"
using namespace fpohm;
#define NOT_YET(name) do { set_error(name ": not implemented in this build"); return FPOHM_ESTATE; } while (0)
extern "C" {
int fpohm_octree_cell_sign(const fpohm_octree *, const fpohm_mesh *, const double *, double, float *) { NOT_YET("fpohm_octree_cell_sign"); }
int fpohm_voxel_grid_setup(const double *, const double *, double, int32_t, int32_t *, double *) { NOT_YET("fpohm_voxel_grid_setup"); }
int fpohm_voxel_sign(fpohm_ctx *, const fpohm_mesh *, const double *, double, const int32_t *, uint8_t *) { NOT_YET("fpohm_voxel_sign"); }
int fpohm_voxel_sign_dev(fpohm_ctx *, const fpohm_mesh *, const double *, double, const int32_t *, uint8_t *, void *) { NOT_YET("fpohm_voxel_sign_dev"); }
int fpohm_voxel_occupancy(fpohm_ctx *, const fpohm_mesh *, const double *, double, const int32_t *, uint8_t *) { NOT_YET("fpohm_voxel_occupancy"); }
int fpohm_dexel_sign(fpohm_ctx *, const fpohm_mesh *, const double *, double, const int32_t *, int64_t *, double *, int64_t *) { NOT_YET("fpohm_dexel_sign"); }
int fpohm_hausdorff_outliers(fpohm_ctx *, fpohm_mesh *, fpohm_mesh *, double, int32_t *, int64_t *) { NOT_YET("fpohm_hausdorff_outliers"); }
int fpohm_polyline_project(fpohm_ctx *, const double *, int64_t, const int64_t *, const int32_t *, const uint8_t *, int64_t, const double *, const int32_t *, int64_t, double *, double *) { NOT_YET("fpohm_polyline_project"); }
int fpohm_hausdorff(fpohm_ctx *, fpohm_mesh *, fpohm_mesh *, int64_t, double *, int64_t *) { NOT_YET("fpohm_hausdorff"); }
int fpohm_hex_connectivity(fpohm_ctx *, const uint32_t *, int64_t, int64_t, fpohm_conn **) { NOT_YET("fpohm_hex_connectivity"); }
int fpohm_conn_sizes(const fpohm_conn *, int64_t *, int64_t *) { NOT_YET("fpohm_conn_sizes"); }
int fpohm_conn_fixed(const fpohm_conn *, uint32_t *, uint32_t *, uint8_t *, uint32_t *, uint8_t *, uint8_t *, uint32_t *) { NOT_YET("fpohm_conn_fixed"); }
int fpohm_conn_csr(const fpohm_conn *, int32_t, int64_t *, uint32_t *, int64_t *) { NOT_YET("fpohm_conn_csr"); }
void fpohm_conn_free(fpohm_conn *) {}
int fpohm_voxel_lattice_dims(const double *, const double *, int32_t, int32_t *) { NOT_YET("fpohm_voxel_lattice_dims"); }
int fpohm_voxel_lattice(fpohm_ctx *, const double *, const double *, int32_t, double *, uint32_t *) { NOT_YET("fpohm_voxel_lattice"); }
}

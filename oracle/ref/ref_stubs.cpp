// Link-time stubs for the two third-party libraries that global_functions.cpp
// references but that are NOT on the hot path and are not built here:
//   * embree (rtc*)  — only used by orient_triangle_mesh_acw, global_functions.cpp:1119
//   * TBB            — only used by project_surface_update_feature (dead code),
//                      global_functions.cpp:3240-3245
// Every stub aborts: if the oracle ever reaches one, that is a bug in the oracle
// driver, not something to paper over.  TEST INFRASTRUCTURE ONLY.
#include <cstdio>
#include <cstdlib>
extern "C" {
static void die(const char *n) { std::fprintf(stderr, "[fpohm_ref] stub %s reached\n", n); std::abort(); }
#define STUB(sym) void stub_##sym(void) __asm__(#sym); void stub_##sym(void) { die(#sym); }
STUB(_ZN3tbb18task_group_context4initEv)
STUB(_ZN3tbb18task_group_contextD1Ev)
STUB(_ZN3tbb4task13note_affinityEt)
STUB(_ZN3tbb8internal36get_initial_auto_partitioner_divisorEv)
STUB(_ZNK3tbb18task_group_context28is_group_execution_cancelledEv)
STUB(_ZNK3tbb8internal20allocate_child_proxy8allocateEm)
STUB(_ZNK3tbb8internal27allocate_continuation_proxy8allocateEm)
STUB(_ZNK3tbb8internal32allocate_root_with_context_proxy4freeERNS_4taskE)
STUB(_ZNK3tbb8internal32allocate_root_with_context_proxy8allocateEm)
STUB(rtcCommit) STUB(rtcDeleteScene) STUB(rtcGetError) STUB(rtcInit) STUB(rtcIntersect)
STUB(rtcMapBuffer) STUB(rtcNewScene) STUB(rtcNewTriangleMesh) STUB(rtcSetMask) STUB(rtcUnmapBuffer)
// grid_hex_meshing.cpp is linked for conforming_mesh only; the SLIM optimiser it references elsewhere (optimization.cpp) is
// not built (io.cpp is: §8(f)-4)
STUB(_ZN12optimization10slim_m_optER13Tetralize_Setjib)
STUB(_ZN12optimization12slim_opt_iglER13Tetralize_Setj)
STUB(_ZN12optimization9pipeline2Ev)
STUB(_ZN12optimization18assign_constraintsERN5Eigen6MatrixIdLin1ELin1ELi0ELin1ELin1EEERSt6vectorI13Deform_V_TypeSaIS5_EE)
// typeinfo object for tbb::task (data symbol)
void *stub_ti_tbb_task[2] __asm__("_ZTIN3tbb4taskE") = {0, 0};
}

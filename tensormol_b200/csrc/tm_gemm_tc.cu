// tcgen05 (5th-gen tensor core) grouped GEMM back-end for the per-element MLPs.
// Placeholder until the TMEM/TMA kernel lands: refuses instead of silently falling back.
#include "tm_internal.h"

int tm_gemm_tc_launch(tm_ctx* c, const GemmGroup* groups, int ngroups, const int* rowmeta_dev, int max_row_tiles, int epilogue) {
  (void)c; (void)groups; (void)ngroups; (void)rowmeta_dev; (void)max_row_tiles; (void)epilogue;
  tm_set_error("tensor-core GEMM mode is not available in this build");
  return TM_ESTATE;
}

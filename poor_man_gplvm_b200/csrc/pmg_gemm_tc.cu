// tcgen05 / TMA tensor-core GEMMs (emission and time-reduction).  Placeholder until the
// tensor-core kernels land: reports "unsupported" so callers use the CUDA-core tiles.
#include "pmg_common.cuh"

int pmg_emission_tc_launch(int64_t, int, int, const float*, int64_t, const float*, const float*, const float*,
                           const float*, float*, int64_t, cudaStream_t) {
  return PMG_ERR_UNSUPPORTED_SHAPE;
}
int64_t pmg_atb_tc_workspace_bytes(int64_t, int, int) { return 0; }
int pmg_atb_tc_launch(int64_t, int, int, const float*, int64_t, const float*, int64_t, float*, int64_t, void*,
                      int64_t, cudaStream_t) {
  return PMG_ERR_UNSUPPORTED_SHAPE;
}

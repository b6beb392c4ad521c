// ctd_generic_search.cu -- cfr_train (ctd_k_mccfr) for any ruleset (see ctd_search.cuh), alone in its translation unit.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define CTD_DEVICE_ONLY 1
#define CTD_NO_PLAYOUT_KERNEL 1
#define CTD_NO_PRED_KERNEL 1
#define CTD_MCCFR_KERNEL_NAME ctd_k_mccfr
#define CTD_MCCFR_PRED_KERNEL_NAME ctd_k_mccfr_pred_unused
#include "ctd_search.cuh"

cudaError_t ctd_mccfr_generic_launch(const CtdMccfrArgs& a, int grid, cudaStream_t stream) {
  ctd_k_mccfr<<<grid, CTD_BLOCK, 0, stream>>>(a);
  return cudaGetLastError();
}
cudaError_t ctd_mccfr_generic_blocks_per_sm(int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, ctd_k_mccfr, CTD_BLOCK, 0);
}

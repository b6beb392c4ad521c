// ctd_generic_pred.cu -- one wave of cfr_pred (ctd_k_mccfr_pred) for any ruleset (see ctd_search.cuh), alone in its unit.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define CTD_DEVICE_ONLY 1
#define CTD_NO_PLAYOUT_KERNEL 1
#define CTD_NO_TRAIN_KERNEL 1
#define CTD_MCCFR_KERNEL_NAME ctd_k_mccfr_unused
#define CTD_MCCFR_PRED_KERNEL_NAME ctd_k_mccfr_pred
#include "ctd_search.cuh"

cudaError_t ctd_mccfr_pred_generic_launch(const CtdPredArgs& p, int grid, cudaStream_t stream) {
  ctd_k_mccfr_pred<<<grid, CTD_BLOCK, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t ctd_mccfr_pred_generic_blocks_per_sm(int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, ctd_k_mccfr_pred, CTD_BLOCK, 0);
}

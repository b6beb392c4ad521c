// ctd_preset_pred.cu -- one wave of cfr_pred (ctd_k_mccfr_pred) specialised for the preset ruleset (see ctd_search.cuh,
// ctd_preset_search.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define CTD_DEVICE_ONLY 1
#define CTD_FIXED_PRESET 1
#define CTD_NO_PLAYOUT_KERNEL 1
#define CTD_NO_TRAIN_KERNEL 1
#define CTD_MCCFR_KERNEL_NAME ctd_k_mccfr_preset_unused
#define CTD_MCCFR_PRED_KERNEL_NAME ctd_k_mccfr_pred_preset
#include "ctd_search.cuh"

cudaError_t ctd_mccfr_pred_preset_launch(const CtdPredArgs& p, int grid, cudaStream_t stream) {
  ctd_k_mccfr_pred_preset<<<grid, CTD_BLOCK, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t ctd_mccfr_pred_preset_blocks_per_sm(int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, ctd_k_mccfr_pred_preset, CTD_BLOCK, 0);
}

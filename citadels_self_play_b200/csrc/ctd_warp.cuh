// ctd_warp.cuh -- warp-level option choice for the fused playout: count the legal options of the game in
// `w`, draw k uniformly from the game's Philox stream, return the k-th option in the reference's list
// order (Agent.get_options order, game/agent.py:50-83) to every lane.  0 = no legal option.
#pragma once
#include "ctd_engine.cuh"

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t ctd_warp_choose(CtdWork& w, int lane) {
  uint64_t d = 0;
  if (lane == 0) {
    CtdEmit e{nullptr, 0, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e);
    if (e.n != 0) {
      uint32_t k = ctd_randbelow(w, e.n);
      CtdEmit e2{nullptr, 0, 0, k, 0};
      ctd_enumerate(w, e2);
      d = e2.got;
    }
  }
  return __shfl_sync(0xFFFFFFFFu, d, 0);
}
#endif

// ctd_warp.cuh -- warp-level option choice for the fused playout: count the legal options of the game in
// `w`, draw k uniformly from the game's Philox stream, return the k-th option in the reference's list
// order (Agent.get_options order, game/agent.py:50-83) to every lane.  0 = no legal option.
#pragma once
#include "ctd_engine.cuh"

#define CTD_CHOOSE_BUF 64 /* descriptors kept from the counting pass; the preset ruleset never exceeds 59 */

#ifdef __CUDACC__
// `buf` is CTD_CHOOSE_BUF descriptors of shared memory owned by this warp.
__device__ __forceinline__ uint64_t ctd_warp_choose(CtdWork& w, int lane, uint64_t* buf) {
  uint64_t d = 0;
  if (lane == 0) {
    CtdEmit e{buf, CTD_CHOOSE_BUF, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e);
    if (e.n != 0) {
      uint32_t k = ctd_randbelow(w, e.n);
      if (k < CTD_CHOOSE_BUF) {
        d = buf[k];
      } else {  // rare (classic Magician): select by a second pass
        CtdEmit e2{buf, 0, 0, k, 0};
        ctd_enumerate(w, e2);
        d = e2.got;
      }
    }
  }
  return __shfl_sync(0xFFFFFFFFu, d, 0);
}
#endif

// ctd_warp.cuh -- warp-cooperative option choice for the fused playout kernel.
//
// One warp owns one game.  For a random playout only two things are needed from Agent.get_options
// (game/agent.py:50-83): how many legal options there are, and which one is the k-th in the reference's list order.
// Here the lanes evaluate the option predicates in parallel -- lane i looks at card i of a hand / building i of a
// city -- and ballots turn the predicates into bit masks: counts are popcounts, "de-duplicate by type, first
// occurrence wins" (game/agent_functions.py:28-32, :99-103, :117-118, :479-480) is __match_any_sync + lowest lane,
// and the k-th option is the k-th set bit (__fns).  Scalars of the record are read by every lane from shared
// memory (broadcast), so control flow stays uniform and branchy character abilities diverge per game, not per lane.
//
// States outside the fast path (Wizard's take-from-hand, Lighthouse, hands over 32 cards) fall back to the scalar
// enumerator on lane 0 (ctd_engine.cuh), which is also the definition both paths are tested against
// (ctd_k_choose_check in ctd_kernels.cu: every k of every step of the recorded reference games).
#pragma once
#include "ctd_engine.cuh"

#define CTD_CHOOSE_BUF 64 /* descriptors kept by the scalar fallback; the preset ruleset never exceeds 59 */

#ifdef __CUDACC__
#define CTD_ALL 0xFFFFFFFFu

// scalar fallback: lane 0 materialises the list, draws k, picks
__device__ __noinline__ uint64_t ctd_choose_scalar(CtdWork& w, int lane, uint64_t* buf, uint32_t* count_out, int want) {
  uint64_t d = 0;
  uint32_t n = 0;
  if (lane == 0) {
    CtdEmit e{buf, CTD_CHOOSE_BUF, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e);
    n = e.n;
    if (n != 0) {
      uint32_t k = want >= 0 ? (uint32_t)want : ctd_randbelow(w, n);
      if (k < CTD_CHOOSE_BUF && k < n) {
        d = buf[k];
      } else if (k < n) {
        CtdEmit e2{buf, 0, 0, k, 0};
        ctd_enumerate(w, e2);
        d = e2.got;
      }
    }
  }
  if (count_out) *count_out = __shfl_sync(CTD_ALL, n, 0);
  return __shfl_sync(CTD_ALL, d, 0);
}

__device__ __forceinline__ uint32_t ctd_draw_uniform(CtdWork& w, int lane, uint32_t n, int want) {
  if (want >= 0) return (uint32_t)want;
  uint32_t k = 0;
  if (lane == 0) k = ctd_randbelow(w, n);
  return __shfl_sync(CTD_ALL, k, 0);
}

// position of the k-th (0-based) set bit
__device__ __forceinline__ int ctd_kth_bit(uint32_t mask, uint32_t k) { return (int)__fns(mask, 0, (int)k + 1); }

// "first card of its type in this list" for lane < n (lanes beyond n never match anything)
__device__ __forceinline__ bool ctd_first_of_type(int lane, int n, int t) {
  const int key = lane < n ? t : 64 + lane;
  const uint32_t m = __match_any_sync(CTD_ALL, key);
  return lane < n && (__ffs(m) - 1) == lane;
}

// Count the options of the game in `w`, draw k (or take `want` >= 0, used by the checker), return the k-th.
// *count_out (optional) receives the number of options.  Every lane returns the same descriptor; 0 = none.
__device__ __noinline__ uint64_t ctd_warp_choose(CtdWork& w, int lane, uint64_t* buf, uint32_t* count_out = nullptr,
                                                 int want = -1) {
  if ((w.gflags & 2) || w.player >= 6) return ctd_choose_scalar(w, lane, buf, count_out, want);
  const int p = w.player, st = w.state;
  if (st == 0) {  // pick_role_options: one per role still on offer, rank ascending
    const uint32_t m = w.rtc_mask, n = __popc(m);
    if (count_out) *count_out = n;
    if (n == 0) return 0;
    const uint32_t k = ctd_draw_uniform(w, lane, n, want);
    return ctd_opt(CTD_K_ROLE_PICK, p) | ctd_f_rank(ctd_kth_bit(m, k));
  }
  const int role = w.role[p];
  if (role >= 8) return ctd_choose_scalar(w, lane, buf, count_out, want);  // None / Bewitched: rare, error paths
  const int nm = role * 3 + w.variant[role];
  const int rp = w.rprops[role];
  if (rp & CTD_RP_DEAD) return ctd_choose_scalar(w, lane, buf, count_out, want);
  if (st == 1) {  // gold_or_card_options
    const uint32_t n = w.n_deck > 1 ? 2 : 1;
    if (count_out) *count_out = n;
    const uint32_t k = ctd_draw_uniform(w, lane, n, want);
    return ctd_opt(CTD_K_GOLD_OR_CARD, p) | ctd_f_named(k == 0 ? CTD_N_GOLD : CTD_N_CARD);
  }
  if (st == 3 && !(rp & CTD_RP_BLACKMAIL)) {
    if (count_out) *count_out = 1;
    ctd_draw_uniform(w, lane, 1, want);
    return ctd_opt(CTD_K_EMPTY, p);
  }
  if (st == 6) {
    if (count_out) *count_out = 1;
    ctd_draw_uniform(w, lane, 1, want);
    return ctd_opt(w.gold[p] > 0 ? CTD_K_GRAVEYARD : CTD_K_EMPTY, p);
  }
  const int nb = w.n_bld[p];
  // which building types do I own (64-bit mask by type) -- lanes OR their building's bit
  uint32_t own_lo, own_hi;
  {
    uint64_t bit = 0;
    if (lane < nb) bit = 1ull << ctd_ctype(w.bld[p][lane]);
    own_lo = __reduce_or_sync(CTD_ALL, (uint32_t)bit);
    own_hi = __reduce_or_sync(CTD_ALL, (uint32_t)(bit >> 32));
  }
  const uint64_t own = (uint64_t)own_lo | ((uint64_t)own_hi << 32);
  if (st == 2) {  // which_card_to_keep_options
    const int n = w.n_jd[p];
    if (n > 32) return ctd_choose_scalar(w, lane, buf, count_out, want);
    const int t = lane < n ? ctd_ctype(w.jd[p][lane]) : 0;
    if ((own >> 20) & 1) {  // Library: every pair i < j, no de-duplication
      const uint32_t cnt = (uint32_t)(n * (n - 1) / 2);
      if (count_out) *count_out = cnt;
      if (cnt == 0) return 0;
      uint32_t k = ctd_draw_uniform(w, lane, cnt, want);
      int i = 0;
      while (k >= (uint32_t)(n - 1 - i)) { k -= (uint32_t)(n - 1 - i); ++i; }
      const int j = i + 1 + (int)k;
      return ctd_opt(CTD_K_KEEP, p) | ctd_f_a(ctd_ctype(w.jd[p][i])) | ctd_f_b(ctd_ctype(w.jd[p][j]));
    }
    const uint32_t m = __ballot_sync(CTD_ALL, ctd_first_of_type(lane, n, t));
    const uint32_t cnt = __popc(m);
    if (count_out) *count_out = cnt;
    if (cnt == 0) return 0;
    const uint32_t k = ctd_draw_uniform(w, lane, cnt, want);
    return ctd_opt(CTD_K_KEEP, p) | ctd_f_a(ctd_ctype(w.jd[p][ctd_kth_bit(m, k)]));
  }
  if (st != 5) return ctd_choose_scalar(w, lane, buf, count_out, want);
  if (nm == CTD_WITCH) {  // witch_options: ranks 1..7
    if (count_out) *count_out = 7;
    const uint32_t k = ctd_draw_uniform(w, lane, 7, want);
    return ctd_opt(CTD_K_BEWITCHING, p) | ctd_f_rank(1 + (int)k);
  }
  if (rp & CTD_RP_POSSESSED) {
    if (count_out) *count_out = 1;
    ctd_draw_uniform(w, lane, 1, want);
    return ctd_opt(CTD_K_FINISH, p) | ctd_f_next_witch(1) | ctd_f_crown(nm == CTD_KING || nm == CTD_PATRICIAN);
  }
  // ---------------------------------------------------------------- main_round_options, in the reference's order
  const int nh = w.n_hand[p];
  const bool lighthouse = ((own >> 29) & 1) && (w.pflags[p] & CTD_PF_LIGHTHOUSE);
  const bool tier_ab = nm == CTD_SPY || nm == CTD_WIZARD || nm == CTD_KING || nm == CTD_ABBOT || nm == CTD_ALCHEMIST ||
                       nm == CTD_NAVIGATOR || nm == CTD_WARLORD || nm == CTD_ASSASSIN || nm == CTD_THIEF ||
                       nm == CTD_MAGICIAN || nm == CTD_BISHOP || nm == CTD_MERCHANT || nm == CTD_ARCHITECT;
  if (nh > 32 || lighthouse || !tier_ab) return ctd_choose_scalar(w, lane, buf, count_out, want);
  const int gold = w.gold[p];
  const int done = w.done;
  const int hc = lane < nh ? w.hand[p][lane] : 0;
  const int ht = ctd_ctype(hc);
  const bool hfirst = ctd_first_of_type(lane, nh, ht);
  // 1. builds
  uint32_t m_build = 0;
  if (w.n_trade + w.n_nontrade < ctd_build_limit(nm)) {
    const bool factory = (own >> 35) & 1;
    m_build = __ballot_sync(CTD_ALL, hfirst && ctd_build_cost(hc, factory) <= gold);
  }
  const uint32_t c_build = __popc(m_build);
  // 2. character
  uint32_t c_char = 0, m_char = 0, m_war[6] = {0, 0, 0, 0, 0, 0};
  if (!(done & CTD_DM_CHARACTER)) {
    switch (nm) {
      case CTD_ASSASSIN: c_char = 7; break;
      case CTD_THIEF: c_char = 6; break;
      case CTD_SPY: c_char = 25; break;
      case CTD_MAGICIAN: {
        uint32_t c = lane < nh ? ctd_magician_count(nh, lane + 1) : 0;
        c_char = 5 + __reduce_add_sync(CTD_ALL, c);
        break;
      }
      case CTD_WIZARD:
        m_char = __ballot_sync(CTD_ALL, lane < 6 && lane != p && w.n_hand[lane < 6 ? lane : 0] > 0);
        c_char = __popc(m_char);
        break;
      case CTD_KING: case CTD_BISHOP: case CTD_MERCHANT: case CTD_ARCHITECT: c_char = 1; break;
      case CTD_ABBOT: {
        m_char = __ballot_sync(CTD_ALL, lane < nh && ctd_csuit(hc) == CTD_SUIT_RELIGION);
        const uint32_t n = __popc(m_char);
        c_char = n > 0 ? n + 1 : 0;
        break;
      }
      case CTD_NAVIGATOR: c_char = 2; break;
      case CTD_WARLORD:
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int nq = w.n_bld[q];
          if (nq >= 7 || ctd_name(w, q) == CTD_BISHOP) continue;  // uniform per q
          const int c = lane < nq ? w.bld[q][lane] : 0;
          const int t = ctd_ctype(c);
          const bool first = ctd_first_of_type(lane, nq, t);
          m_war[q] = __ballot_sync(CTD_ALL, first && ctd_ccost(c) - 1 <= gold && t != 17);
          c_char += __popc(m_war[q]);
        }
        break;
      default: break;
    }
  }
  const uint32_t c_beg = (nm == CTD_ABBOT && !(done & CTD_DM_BEGGED)) ? 1 : 0;
  const uint32_t c_war_gold = (nm == CTD_WARLORD && !(done & CTD_DM_TAKE_GOLD)) ? 1 : 0;
  // 3..8 unique buildings
  const uint32_t c_smithy = (((own >> 21) & 1) && gold >= 2 && !(done & CTD_DM_SMITHY)) ? 1 : 0;
  const uint32_t c_lab = (((own >> 22) & 1) && !(done & CTD_DM_LAB)) ? (uint32_t)nh : 0;
  const uint32_t c_ms = (((own >> 25) & 1) && !(done & CTD_DM_MAGIC_SCHOOL)) ? 5 : 0;
  uint32_t c_ws = 0;
  if ((own >> 27) & 1)
    for (int q = 0; q < 6; ++q) c_ws += q != p ? w.n_bld[q] : 0;
  uint32_t m_mus = 0;
  if (((own >> 34) & 1) && !(done & CTD_DM_MUSEUM)) m_mus = __ballot_sync(CTD_ALL, hfirst);
  const uint32_t c_mus = __popc(m_mus);
  const uint32_t total = c_build + c_char + c_beg + c_war_gold + c_smithy + c_lab + c_ms + c_ws + c_mus + 1;
  if (count_out) *count_out = total;
  uint32_t k = ctd_draw_uniform(w, lane, total, want);
  // ---------------------------------------------------------------- select the k-th
  if (k < c_build) {
    const int c = w.hand[p][ctd_kth_bit(m_build, k)], t = ctd_ctype(c);
    const int rep = (((own >> t) & 1) && !w.replicas[p]) ? w.replicas[p] + 1 : 0;
    return ctd_opt(CTD_K_BUILD, p) | ctd_f_a(t) | ctd_f_replica(rep);
  }
  k -= c_build;
  if (k < c_char) {
    switch (nm) {
      case CTD_ASSASSIN: return ctd_opt(CTD_K_ASSASSINATION, p) | ctd_f_rank(1 + (int)k);
      case CTD_THIEF: return ctd_opt(CTD_K_STEAL, p) | ctd_f_rank(2 + (int)k);
      case CTD_SPY: {
        int q = (int)k / 5;
        q += q >= p ? 1 : 0;
        return ctd_opt(CTD_K_SPY, p) | ctd_f_target(q) | ctd_f_named(CTD_N_TRADE + (int)k % 5);
      }
      case CTD_MAGICIAN: {
        if (k < 5) {
          int q = (int)k;
          q += q >= p ? 1 : 0;
          return ctd_opt(CTD_K_MAGIC_HAND_CHANGE, p) | ctd_f_target(q);
        }
        k -= 5;  // every discard_and_draw option has the same effect; recover (r, j) for the descriptor
        int r = 1;
        for (; r <= nh; ++r) {
          const uint32_t c = ctd_magician_count(nh, r);
          if (k < c) break;
          k -= c;
        }
        return ctd_opt(CTD_K_DISCARD_AND_DRAW, p) | ctd_f_r(r) | ctd_f_j(k);
      }
      case CTD_WIZARD: return ctd_opt(CTD_K_LOOK_AT_HAND, p) | ctd_f_target(ctd_kth_bit(m_char, k));
      case CTD_KING: return ctd_opt(CTD_K_TAKE_CROWN_KING, p);
      case CTD_BISHOP: return ctd_opt(CTD_K_BISHOP, p);
      case CTD_MERCHANT: return ctd_opt(CTD_K_MERCHANT, p);
      case CTD_ARCHITECT: return ctd_opt(CTD_K_ARCHITECT, p);
      case CTD_ABBOT: return ctd_opt(CTD_K_ABBOT, p) | ctd_f_count((int)k) | ctd_f_r((int)__popc(m_char));
      case CTD_NAVIGATOR: return ctd_opt(CTD_K_NAVIGATOR, p) | ctd_f_named(k == 0 ? CTD_N_4GOLD : CTD_N_4CARD);
      default:  // CTD_WARLORD
        for (int q = 0; q < 6; ++q) {
          const uint32_t c = __popc(m_war[q]);
          if (k < c) return ctd_opt(CTD_K_WARLORD, p) | ctd_f_target(q) | ctd_f_a(ctd_ctype(w.bld[q][ctd_kth_bit(m_war[q], k)]));
          k -= c;
        }
        return 0;
    }
  }
  k -= c_char;
  if (k < c_beg) return ctd_opt(CTD_K_ABBOT_BEG, p);
  k -= c_beg;
  if (k < c_war_gold) return ctd_opt(CTD_K_TAKE_GOLD_WAR, p);
  k -= c_war_gold;
  if (k < c_smithy) return ctd_opt(CTD_K_SMITHY, p);
  k -= c_smithy;
  if (k < c_lab) return ctd_opt(CTD_K_LAB, p) | ctd_f_a(ctd_ctype(w.hand[p][k]));
  k -= c_lab;
  if (k < c_ms) return ctd_opt(CTD_K_MAGIC_SCHOOL, p) | ctd_f_named(CTD_N_TRADE + (int)k);
  k -= c_ms;
  if (k < c_ws) {
    for (int q = 0; q < 6; ++q) {
      if (q == p) continue;
      const uint32_t c = w.n_bld[q];
      if (k < c) return ctd_opt(CTD_K_WEAPON_STORAGE, p) | ctd_f_target(q) | ctd_f_a(ctd_ctype(w.bld[q][k]));
      k -= c;
    }
    return 0;
  }
  k -= c_ws;
  if (k < c_mus) return ctd_opt(CTD_K_MUSEUM, p) | ctd_f_a(ctd_ctype(w.hand[p][ctd_kth_bit(m_mus, k)]));
  return ctd_opt(CTD_K_FINISH, p);
}
#endif

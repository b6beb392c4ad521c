// Host build of the engine's rules code (ctd_engine.cuh) -- a TEST VEHICLE so the kernel logic can be
// replayed against the golden traces without a GPU.  It is not part of the product and the package never
// loads it (the product path fails loudly when the CUDA library is missing).
#include <cstdio>
#include <cstdlib>
#include "../../citadels_self_play_b200/csrc/ctd_engine.cuh"
#include "../../citadels_self_play_b200/csrc/ctd_mccfr.cuh"
#include <string.h>

static CtdWork g_w;

static void chance_for(CtdWork& w, uint64_t seed, uint64_t gid, const uint8_t* tape, uint32_t tape_len) {
  uint32_t draws = w.draws, tpos = w.tape_pos;
  ctd_chance_init(w, seed, gid, draws);
  w.tape = tape_len ? tape : nullptr;
  w.tape_len = tape_len;
  w.tape_pos = tpos;
}

extern "C" {
void hs_new_game(uint64_t seed, uint64_t gid, int ruleset, const uint8_t* tape, uint32_t tape_len, ctd_state* out) {
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  chance_for(w, seed, gid, tape, tape_len);
  ctd_deal_preset(w, ruleset);
  ctd_setup_round(w);
  ctd_pack(w, out);
}
// enumeration may draw chance (Seer give-back) and may shrink seven_drawn_cards (Scholar): the record is written back
int hs_enumerate(ctd_state* s, uint64_t* opts, uint32_t cap, uint8_t* err, uint64_t seed, uint64_t gid, const uint8_t* tape,
                 uint32_t tape_len) {
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  ctd_unpack(s, w);
  chance_for(w, seed, gid, tape, tape_len);
  CtdEmit e{opts, cap, 0, 0xFFFFFFFFu, 0};
  ctd_enumerate(w, e);
  if (err) *err = w.err;
  ctd_pack(w, s);
  return (int)e.n;
}
int hs_step(ctd_state* s, uint64_t d, uint64_t seed, uint64_t gid, const uint8_t* tape, uint32_t tape_len) {
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  ctd_unpack(s, w);
  chance_for(w, seed, gid, tape, tape_len);
  bool won = ctd_apply(w, d);
  ctd_pack(w, s);
  return won ? w.winner : -1;
}
// fused playout, same loop as the device kernel: returns winner, fills points/steps, err
int hs_playout(uint64_t seed, uint64_t gid, int ruleset, uint32_t max_steps, int8_t* points6, uint32_t* steps,
               uint8_t* err, ctd_state* final_state) {
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  ctd_new_game(w, seed, gid, ruleset);
  // one materialising pass (the Seer's / Scholar's enumerations are not pure, so no count-then-select here)
  static uint64_t buf[8192];
  while (!(w.gflags & 2) && !w.err && w.steps < max_steps) {
    CtdEmit e{buf, 8192, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e);
    if (e.n == 0 || e.n > 8192) { w.err |= CTD_ERR_REF_RAISE; break; }
    uint32_t k = ctd_randbelow(w, e.n);
    ctd_apply(w, buf[k]);
  }
  if (!(w.gflags & 2) && !w.err) w.err |= CTD_ERR_MAXSTEPS;
  for (int p = 0; p < 6; ++p) points6[p] = w.points[p];
  *steps = w.steps;
  *err = w.err;
  if (final_state) ctd_pack(w, final_state);
  return w.winner;
}
int hs_sizeof_work() { return (int)sizeof(CtdWork); }

// ctd_make_roots for one root: record, the searching player's knowledge block, used_cards, root step; returns the searching seat
int hs_make_root(uint64_t seed, uint64_t gid, int ruleset, uint32_t back_lo, uint32_t back_hi, int flavour, ctd_state* root,
                 CtdKnow* know, uint8_t* used_cards, uint32_t* root_step) {
  static CtdKnow kn6[6];
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  const int viewer = ctd_make_root(w, kn6, seed, gid, ruleset, back_lo, back_hi, flavour, used_cards, root_step);
  *know = kn6[viewer];
  ctd_pack(w, root);
  return viewer;
}

// One tree on the host: `arena` (arena_bytes) is what the tree allocates from, the result is the compact export block in `out`
// (out_cap bytes; *out_bytes receives its size).  Returns the tree status, or -1 when `out` is too small.
struct HsTree {
  CtdTreeHdr hdr;
  unsigned long long used;
  CtdTree T;
};
static CtdKnow hs_kn;
static uint64_t hs_opts[CTD_MCCFR_OPT_CAP];
static uint8_t hs_scratch[CTD_TREE_SCRATCH] __attribute__((aligned(16)));
static ctd_state hs_stage __attribute__((aligned(16)));

static void hs_tree_begin(HsTree& H, const ctd_state* root, const CtdKnow* know, const uint8_t* used_cards, uint64_t seed, uint64_t gid,
                          uint32_t iters, uint8_t* arena, uint64_t arena_bytes, bool model) {
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  memset(&H.hdr, 0, sizeof(H.hdr));
  H.used = 1;
  CtdTree& T = H.T;
  T.w = &w; T.kn = &hs_kn; T.opts = hs_opts; T.scratch = hs_scratch; T.stage = &hs_stage; T.vnet = nullptr; T.act = nullptr;
  T.hdr = &H.hdr;
  memcpy(H.hdr.used_cards, used_cards, 76);
  ctd_tree_stage_used(T);
  ctd_unpack(root, w);
  ctd_chance_init(w, seed, gid, 0);
  w.stream = 1;
  hs_kn = *know;
  uint32_t k = 6;   // ctd_n0_log2 of the engine
  while (k < 20 && (1ull << k) < 2ull * iters) ++k;
  if (getenv("HS_N0_LOG2")) k = (uint32_t)atoi(getenv("HS_N0_LOG2"));   // tests: small first chunk exercises the chunk table
  CtdArena ar{arena, &H.used, arena_bytes / CTD_ARENA_UNIT};
  ctd_tree_init(T, &H.hdr, ar, k, know->viewer, gid, false, model);
}
static uint64_t hs_live = 0;
extern "C" uint64_t hs_last_live_option() { return hs_live; }   // ctd_live_choice of the tree built last
static int hs_tree_end(HsTree& H, uint8_t* out, uint64_t out_cap, uint64_t* out_bytes) {
  hs_live = ctd_live_choice(H.T);
  const uint64_t need = ctd_tree_export_bytes(H.hdr.n_nodes, H.hdr.child_used, H.hdr.arr_used);
  if (out_bytes) *out_bytes = need;
  if (getenv("HS_DEBUG")) fprintf(stderr, "hs tree gid %llu status %u w.err %u kn.err %u nodes %u arena units %llu\n", (unsigned long long)H.hdr.gid,
                                  H.hdr.status, g_w.err, hs_kn.err, H.hdr.n_nodes, H.used);
  if (out) {
    if (need > out_cap) return -1;
    memset(out, 0, need);
    ctd_tree_export(H.T, out);
  }
  return (int)H.hdr.status;
}

int hs_mccfr(const ctd_state* root, const CtdKnow* know, const uint8_t* used_cards, uint64_t seed, uint64_t gid, uint32_t iters,
             uint8_t* arena, uint64_t arena_bytes, uint8_t* out, uint64_t out_cap, uint64_t* out_bytes) {
  static HsTree H;
  hs_tree_begin(H, root, know, used_cards, seed, gid, iters, arena, arena_bytes, false);
  ctd_cfr_train(H.T, iters);
  return hs_tree_end(H, out, out_cap, out_bytes);
}
int hs_sizeof_node() { return (int)sizeof(CtdNode); }

// deep MCCFR on one root; `eval(features[448], pred[6])` stands in for the batched value kernel
typedef void (*hs_eval_fn)(const float*, float*);
int hs_mccfr_pred(const ctd_state* root, const CtdKnow* know, const uint8_t* used_cards, uint64_t seed, uint64_t gid,
                  uint32_t iters, uint32_t max_depth, uint8_t* arena, uint64_t arena_bytes, uint8_t* out, uint64_t out_cap,
                  uint64_t* out_bytes, hs_eval_fn eval) {
  static HsTree H;
  static float feat[CTD_FEATURES_PAD], pred[8];
  hs_tree_begin(H, root, know, used_cards, seed, gid, iters, arena, arena_bytes, true);
  // a small per-call budget exercises the yield / resume path of the wave scheduler as well
  const uint32_t budget = getenv("HS_PRED_BUDGET") ? (uint32_t)atoi(getenv("HS_PRED_BUDGET")) : 7u;
  for (;;) {
    const int r = ctd_cfr_pred_advance(H.T, iters, max_depth, feat, pred, budget);
    if (r == CTD_PRED_DONE) break;
    if (r == CTD_PRED_WAIT) eval(feat, pred);
  }
  return hs_tree_end(H, out, out_cap, out_bytes);
}
}

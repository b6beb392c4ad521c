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

// pure MCCFR on one root, tree built in `tree_buf` (ctd_tree_bytes(max_nodes, child_cap, arr_cap) bytes)
int hs_mccfr(const ctd_state* root, const CtdKnow* know, const uint8_t* used_cards, uint64_t seed, uint64_t gid,
             uint32_t iters, uint32_t max_nodes, uint32_t child_cap, uint32_t arr_cap, uint8_t* tree_buf) {
  static CtdKnow kn;
  static uint64_t opts[CTD_MCCFR_OPT_CAP];
  static uint8_t scratch[384] __attribute__((aligned(16)));
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  CtdTree T;
  T.hdr = (CtdTreeHdr*)tree_buf;
  T.nodes = (CtdNode*)(tree_buf + sizeof(CtdTreeHdr));
  T.children = (CtdChild*)((uint8_t*)T.nodes + (size_t)max_nodes * sizeof(CtdNode));
  T.arr = (double*)((uint8_t*)T.children + (size_t)child_cap * sizeof(CtdChild));
  static ctd_state hs_stage __attribute__((aligned(16)));
  T.w = &w; T.kn = &kn; T.opts = opts; T.scratch = scratch; T.stage = &hs_stage;
  memset(tree_buf, 0, ctd_tree_bytes(max_nodes, child_cap, arr_cap));
  memcpy(T.hdr->used_cards, used_cards, 76);
  ctd_tree_stage_used(T);
  ctd_unpack(root, w);
  ctd_chance_init(w, seed, gid, 0);
  w.stream = 1;
  kn = *know;
  ctd_tree_init(T, max_nodes, child_cap, arr_cap, know->viewer, gid, false, false);
  ctd_cfr_train(T, iters);
  ctd_tree_pack_nodes(T);
  if (getenv("HS_DEBUG")) fprintf(stderr, "hs_mccfr gid %llu status %u w.err %u kn.err %u nodes %u\n", (unsigned long long)gid,
                                  T.hdr->status, w.err, kn.err, T.hdr->n_nodes);
  return (int)T.hdr->status;
}
int hs_sizeof_node() { return (int)sizeof(CtdNode); }

// deep MCCFR on one root; `eval(features[448], pred[6])` stands in for the batched value kernel
typedef void (*hs_eval_fn)(const float*, float*);
int hs_mccfr_pred(const ctd_state* root, const CtdKnow* know, const uint8_t* used_cards, uint64_t seed, uint64_t gid,
                  uint32_t iters, uint32_t max_depth, uint32_t max_nodes, uint32_t child_cap, uint32_t arr_cap,
                  uint8_t* tree_buf, hs_eval_fn eval) {
  static CtdKnow kn;
  static uint64_t opts[CTD_MCCFR_OPT_CAP];
  static uint8_t scratch[384] __attribute__((aligned(16)));
  static float feat[CTD_FEATURES_PAD], pred[8];
  CtdWork& w = g_w;
  memset(&w, 0, sizeof(w));
  CtdTree T;
  T.hdr = (CtdTreeHdr*)tree_buf;
  T.nodes = (CtdNode*)(tree_buf + sizeof(CtdTreeHdr));
  T.children = (CtdChild*)((uint8_t*)T.nodes + (size_t)max_nodes * sizeof(CtdNode));
  T.arr = (double*)((uint8_t*)T.children + (size_t)child_cap * sizeof(CtdChild));
  static ctd_state hs_stage __attribute__((aligned(16)));
  T.w = &w; T.kn = &kn; T.opts = opts; T.scratch = scratch; T.stage = &hs_stage;
  memset(tree_buf, 0, ctd_tree_bytes(max_nodes, child_cap, arr_cap));
  memcpy(T.hdr->used_cards, used_cards, 76);
  ctd_tree_stage_used(T);
  ctd_unpack(root, w);
  ctd_chance_init(w, seed, gid, 0);
  w.stream = 1;
  kn = *know;
  ctd_tree_init(T, max_nodes, child_cap, arr_cap, know->viewer, gid, false, true);
  // a small per-call budget exercises the yield / resume path of the wave scheduler as well
  const uint32_t budget = getenv("HS_PRED_BUDGET") ? (uint32_t)atoi(getenv("HS_PRED_BUDGET")) : 7u;
  for (;;) {
    const int r = ctd_cfr_pred_advance(T, iters, max_depth, feat, pred, budget);
    if (r == CTD_PRED_DONE) break;
    if (r == CTD_PRED_WAIT) eval(feat, pred);
  }
  ctd_tree_pack_nodes(T);
  return (int)T.hdr->status;
}
}

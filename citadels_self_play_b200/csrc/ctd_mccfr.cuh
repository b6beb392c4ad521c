// ctd_mccfr.cuh -- the reference's MCCFR tree search (algorithms/deep_mccfr.py `CFRNode`) on a flat node pool.
//
// One tree is private to one root state (algorithms/deep_mccfr.py:27-29) and is grown by sequential dependent
// sampling, so a tree is owned by one warp and many trees run side by side.  A tree is one contiguous HBM block:
//
//   CtdTreeHdr | CtdNode[max_nodes] | CtdChild[child_cap] | double[arr_cap]
//
// CtdNode = 128 B header (parent, depth, player, flags, V[6], P[6], pred[6]) + the 256 B packed game record
// + the 592 B knowledge block of the searching player.  A child entry is (option descriptor, node index);
// regrets / strategy / cumulative strategy of an expanded node are 3 x K doubles in the array arena
// (role-pick nodes: 3 x 6 x 10, stored [player][child] like the reference after its transposes, :129-131).
// Everything is fp64 like the reference's numpy arrays.
//
// Chance: one Philox stream per tree (stream word 1, keyed by (seed, root id)); the mapping of the reference's
// random calls onto it is written out in oracle/mccfr_oracle.py and is the same here.
#pragma once
#include <math.h>
#include "ctd_engine.cuh"

#define CTD_MCCFR_OPT_CAP 4096 /* legal options of one state that expansion can materialise (preset max 59,
                                  classic Magician hands reach ~1600, the Cardinal ~2600); the buffer lives in HBM scratch, one per warp */

enum { CTD_NF_ROLE_PICK = 1, CTD_NF_TERMINAL = 2, CTD_NF_HAS_PRED = 4 };
enum { CTD_TREE_OK = 0, CTD_TREE_TERMINAL_ROOT = 1, CTD_TREE_EPOOL = 2, CTD_TREE_EENGINE = 4, CTD_TREE_EOPTS = 8 };

struct CtdNode {
  int32_t parent;
  uint16_t depth;
  uint8_t player;      // current_player_id after skip_false_choice
  uint8_t flags;
  uint32_t n_children;
  uint32_t child_cap;
  uint32_t child_off;  // first CtdChild
  uint32_t arr_off;    // R | s | C
  uint32_t visits;
  uint32_t pad0;
  double V[6];         // node_value
  double P[6];         // winning_probabilities
  float pred[6];       // pred_node_value (deep MCCFR)
  uint8_t order[6];    // game.turn_orders_for_roles (role-pick nodes weight their strategy by it)
  uint8_t gstate;      // game.gamestate.state
  int8_t winner;       // game winner (terminal nodes)
  ctd_state game;      // packed record: filled by ctd_tree_pack_nodes when the tree is exported, not on the hot path
  CtdKnow know;
  uint8_t snap[CTD_SNAP_BYTES];  // the working record verbatim: node <-> shared memory is a plain vector copy
};
static_assert(sizeof(CtdNode) == 160 + 256 + 592 + CTD_SNAP_BYTES, "CtdNode layout");
static_assert(offsetof(CtdNode, game) % 16 == 0 && offsetof(CtdNode, know) % 16 == 0 && offsetof(CtdNode, snap) % 16 == 0, "CtdNode alignment");

struct CtdChild {
  uint64_t desc;
  uint32_t node;
  uint32_t pad;
};

struct CtdTreeHdr {
  uint32_t n_nodes, max_nodes;
  uint32_t child_used, child_cap;
  uint32_t arr_used, arr_cap;
  uint32_t status;
  uint32_t iterations;
  uint32_t rng_draws;
  uint8_t viewer;        // original_player_id
  uint8_t training;
  uint8_t has_model;
  uint8_t phase;         // deep MCCFR walk: 0 not started, 1 walking, 2 waiting for a leaf value, 3 finished
  uint64_t gid;
  uint8_t used_cards[76];  // Game.used_cards in deal order (game/game.py:424); constant over the tree
  uint32_t cur_node;     // node the walk stands on (deep MCCFR is resumed across kernel launches)
};
static_assert(sizeof(CtdTreeHdr) == 128 && offsetof(CtdTreeHdr, used_cards) % 16 == 0, "CtdTreeHdr layout");

CTD_HD inline size_t ctd_tree_bytes(uint32_t max_nodes, uint32_t child_cap, uint32_t arr_cap) {
  return sizeof(CtdTreeHdr) + (size_t)max_nodes * sizeof(CtdNode) + (size_t)child_cap * sizeof(CtdChild) +
         (size_t)arr_cap * sizeof(double);
}

// a tree plus the on-chip working set of the warp that grows it
struct CtdTree {
  CtdTreeHdr* hdr;
  CtdNode* nodes;
  CtdChild* children;
  double* arr;
  CtdWork* w;       // working game (shared memory on the device)
  CtdKnow* kn;      // working knowledge of the viewer
  uint64_t* opts;   // CTD_MCCFR_OPT_CAP descriptors
  uint8_t* scratch; // >= 384 bytes, 16-byte aligned: [0,256) determinisation scratch, [256,336) Game.used_cards staged on chip
  ctd_state* stage; // 16-byte aligned staging record (shared memory on the device)
};

// address spaces of a tree's parts on the device: working set in shared memory, the tree block in HBM
#define CTD_TREE_SPACES(T)                                                                                  \
  CTD_ASSUME_SHARED((T).w); CTD_ASSUME_SHARED((T).kn); CTD_ASSUME_SHARED((T).stage);                         \
  CTD_ASSUME_GLOBAL((T).hdr); CTD_ASSUME_GLOBAL((T).nodes); CTD_ASSUME_GLOBAL((T).children); CTD_ASSUME_GLOBAL((T).arr); \
  CTD_ASSUME_GLOBAL((T).opts)

// 16-byte vector copy (both pointers 16-byte aligned, bytes a multiple of 16): the tree lives in HBM and its records
// move as a handful of independent 128-bit transactions instead of hundreds of dependent byte accesses
CTD_HD inline void ctd_copy16(void* dst, const void* src, int bytes) {
  struct alignas(16) V { uint32_t x, y, z, w; };
  V* d = (V*)dst;
  const V* s = (const V*)src;
  const int n = bytes / 16;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
  for (int i = 0; i < n; ++i) d[i] = s[i];
}

// The same copy with the address spaces spelled out (device only): the working record, the knowledge block and the
// staging record live in shared memory, tree nodes in HBM.  Through generic pointers every 16 bytes cost ~9 instructions
// (address arithmetic + descriptor moves around LD.E / ST.E); ld.shared.v4 / st.global.v4 with immediate offsets cost 2.
#if defined(__CUDA_ARCH__)
// Callers run either on one lane (export / target walks) or with the whole warp converged on the same scalar code
// (the search kernels, see ctd_k_mccfr): the active lanes split the 16-byte chunks between them.
template <int BYTES>
__device__ __forceinline__ void ctd_copy_s2g(void* gdst, const void* ssrc) {
  static_assert(BYTES % 16 == 0, "vector copy");
  const unsigned m = __activemask();
  const int nl = __popc(m), rank = __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  char* g = (char*)gdst;
  __syncwarp(m);   // every lane's (identical) stores to the source are in place
#pragma unroll 4
  for (int i = rank * 16; i < BYTES; i += nl * 16) {
    uint32_t x, y, z, w;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(s + i) : "memory");
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(g + i), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
  }
  __syncwarp(m);
}
template <int BYTES>
__device__ __forceinline__ void ctd_copy_g2s(void* sdst, const void* gsrc) {
  static_assert(BYTES % 16 == 0, "vector copy");
  const unsigned m = __activemask();
  const int nl = __popc(m), rank = __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(sdst);
  const char* g = (const char*)gsrc;
  __syncwarp(m);
#pragma unroll 4
  for (int i = rank * 16; i < BYTES; i += nl * 16) {
    uint32_t x, y, z, w;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(g + i) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(s + i), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
  }
  __syncwarp(m);   // the destination is complete for every lane
}
#define CTD_COPY_S2G(dst, src, bytes) ctd_copy_s2g<bytes>(dst, src)
#define CTD_COPY_G2S(dst, src, bytes) ctd_copy_g2s<bytes>(dst, src)
#else
#define CTD_COPY_S2G(dst, src, bytes) ctd_copy16(dst, src, bytes)
#define CTD_COPY_G2S(dst, src, bytes) ctd_copy16(dst, src, bytes)
#endif

CTD_HD inline double ctd_uniform(CtdWork& w) { return (double)ctd_u32(w) / 4294967296.0; }

// ------------------------------------------------------------------------------------------ determinisation
// Game.sample_private_information (game/game.py:215-242) and its helpers (:183-213, :245-357).  `w`/`k` are the hypothetical game; k.viewer is player_character.
CTD_HD CTD_NI inline void ctd_sample_private(CtdWork& w, CtdKnow& k, const uint8_t* used_cards, bool role_sample,
                                             uint8_t* scratch) {
  CTD_ASSUME_SHARED(&w); CTD_ASSUME_SHARED(&k); CTD_ASSUME_SHARED(scratch);
  const int viewer = k.viewer;
  // (1) which HandKnowledge entries are believed this time: (confidence - 1) * 0.2 > random()   (:217-222)
  CTD_LOOP for (int i = 0; i < k.n_hk; ++i) {
    double r = ctd_uniform(w);
    if ((double)(k.hk[i].conf - 1) * 0.2 > r) k.hk[i].flags |= CTD_HK_USED;
    else k.hk[i].flags &= (uint8_t)~CTD_HK_USED;
  }
  // (2) get_unknown_cards: used_cards minus everything visible, first occurrence per removal  (:183-213)
  uint8_t* unknown = scratch;       // <= 76
  uint8_t* cnt = scratch + 128;     // removals pending per type
  { uint32_t* c4 = (uint32_t*)cnt; CTD_LOOP for (int t = 0; t < 10; ++t) c4[t] = 0; }
  CTD_LOOP for (int p = 0; p < 6; ++p) {
    CTD_LOOP for (int i = 0; i < w.n_bld[p]; ++i) ++cnt[ctd_ctype(w.bld[p][i])];
    CTD_LOOP for (int i = 0; i < w.n_mus[p]; ++i) ++cnt[ctd_ctype(w.mus[p][i])];
  }
  CTD_LOOP for (int i = 0; i < w.n_hand[viewer]; ++i) ++cnt[ctd_ctype(w.hand[viewer][i])];
  CTD_LOOP for (int h = 0; h < k.n_hk; ++h)
    if (k.hk[h].flags & CTD_HK_USED)
      CTD_LOOP for (int i = 0; i < k.hk[h].n; ++i) ++cnt[ctd_ctype(k.pool[k.hk[h].off + i])];
  int nu = 0;
  CTD_LOOP for (int i = 0; i < 76; ++i) {
    int c = used_cards[i];
    if (c == 0xFF) break;  // Game(preset=False) plays with 66 cards
    int t = ctd_ctype(c);
    if (cnt[t] != 0) --cnt[t];
    else unknown[nu++] = (uint8_t)c;
  }
  // (3) sample_deck (:245-262): believed Lighthouse order first, then shuffled unknown cards
  int need = w.n_deck;
  w.deck_head = 0;
  w.n_deck = 0;
  CTD_LOOP for (int h = 0; h < k.n_hk; ++h)
    if (k.hk[h].pid == -1 && (k.hk[h].flags & CTD_HK_USED)) {
      int take = k.hk[h].n < need ? k.hk[h].n : need;
      CTD_LOOP for (int i = 0; i < take; ++i) w.deck[w.n_deck++] = k.pool[k.hk[h].off + i];
      need -= take;
      break;
    }
  ctd_shuffle(w, nu, [unknown](int i) -> uint8_t& { return unknown[i]; });
  int uh = 0;
  CTD_LOOP for (int i = 0; i < need; ++i)
    if (uh < nu) w.deck[w.n_deck++] = unknown[uh++];
  // (3b) sample_warrants_and_blackmails (:321-336): which of the flagged ranks carries the real one is re-rolled,
  // blackmails first
  CTD_LOOP for (int pass = 0; pass < 2; ++pass) {
    const int mask = pass == 0 ? CTD_RP_BLACKMAIL : CTD_RP_WARRANT, sh = pass == 0 ? 5 : 1;
    uint8_t* keys = scratch + 192;
    int nk = 0;
    CTD_LOOP for (int r = 0; r < 8; ++r)
      if (w.rprops[r] & mask) keys[nk++] = (uint8_t)r;
    if (nk == 0) continue;
    ctd_shuffle(w, nk, [keys](int i) -> uint8_t& { return keys[i]; });
    const int real = keys[0];
    CTD_LOOP for (int r = 0; r < 8; ++r)
      if (w.rprops[r] & mask) w.rprops[r] = (uint8_t)((w.rprops[r] & ~mask) | ((r == real ? 1 : 2) << sh));
  }
  // (4) roles the viewer can still believe in (:230-232, :298-310)
  uint16_t kr[6];
  CTD_LOOP for (int q = 0; q < 6; ++q) kr[q] = k.kr[q];
  const uint8_t conf = k.conf_mask;
  if (role_sample) {
    int r = w.role[w.player];
    uint16_t clear = 0;
    if (r == CTD_ROLE_BEWITCHED) clear = 1u << 8;
    else if (r != CTD_ROLE_NONE) clear = (uint16_t)((2u << r) - 1);  // own rank and every smaller one
    CTD_LOOP for (int q = 0; q < 6; ++q)
      if (!((conf >> q) & 1)) kr[q] &= (uint16_t)~clear;
  }
  // (5) per seat: hand of the same size (:264-280), then a role (:283-295)
  CTD_LOOP for (int p = 0; p < 6; ++p) {
    if (p != viewer) {
      int n = w.n_hand[p];
      w.n_hand[p] = 0;
      CTD_LOOP for (int h = 0; h < k.n_hk; ++h)
        if (k.hk[h].pid == p && (k.hk[h].flags & CTD_HK_USED)) {
          int take = k.hk[h].n < n ? k.hk[h].n : n;
          CTD_LOOP for (int i = 0; i < take; ++i) w.hand[p][w.n_hand[p]++] = k.pool[k.hk[h].off + i];
          n -= take;
          break;
        }
      CTD_LOOP for (int i = 0; i < n; ++i)
        if (uh < nu) w.hand[p][w.n_hand[p]++] = unknown[uh++];
    }
    if (role_sample && p != viewer && p != w.player && w.state != 0) {
      int m = kr[p], cntb = 0;
      CTD_LOOP for (int x = 0; x < 9; ++x) cntb += (m >> x) & 1;
      if (cntb != 0) {
        int pick = (int)ctd_randbelow(w, (uint32_t)cntb), x = 0;
        for (;; ++x)
          if ((m >> x) & 1) { if (pick == 0) break; --pick; }
        w.role[p] = (uint8_t)(x == 8 ? CTD_ROLE_BEWITCHED : x);
        CTD_LOOP for (int q = 0; q < 6; ++q)
          if (!((conf >> q) & 1)) kr[q] &= (uint16_t)~(1u << x);
      } else {  // "band aid" (:293-295): first rank nobody is known to hold
        int x = 0;
        for (; x < 8; ++x) {
          bool used = false;
          CTD_LOOP for (int i = 0; i < w.used_len; ++i) used |= (int)w.used_roles[i] - 1 == x;
          if (!used) break;
        }
        if (x == 8) { w.err |= CTD_ERR_REF_RAISE; x = 0; }
        w.role[p] = (uint8_t)x;
      }
    }
  }
  // (6) refresh_roles_after_sampling_roles (:339-357)
  if (role_sample && w.state != 0) ctd_refresh_used_roles(w);
}

// ------------------------------------------------------------------------------------------ node helpers
CTD_HD inline void ctd_node_store(CtdTree& T, CtdNode& n) {
  const CtdWork& w = *T.w;
  CTD_COPY_S2G(n.snap, &w, CTD_SNAP_BYTES);
  CTD_COPY_S2G(&n.know, T.kn, (int)sizeof(CtdKnow));
  CTD_LOOP for (int i = 0; i < 6; ++i) n.order[i] = w.order[i];
  n.gstate = w.state;
  n.winner = w.winner;
}
// export form: the 256-byte packed record of a node (tests, facade, ctd_mccfr trees_out)
CTD_HD inline void ctd_node_pack(CtdTree& T, CtdNode& n) {
  CtdWork& w = *T.w;
  ctd_copy16(&w, n.snap, CTD_SNAP_BYTES);
  w.draws = 0; w.tape_pos = 0; w.steps = 0;
  w.g0 = (uint32_t)T.hdr->gid; w.g1 = (uint32_t)(T.hdr->gid >> 32);
  ctd_pack(w, T.stage);
  ctd_copy16(&n.game, T.stage, (int)sizeof(ctd_state));
}
CTD_HD inline void ctd_node_load(CtdTree& T, const CtdNode& n) {
  // chance state lives in the working record and must survive a load
  CtdWork& w = *T.w;
  CTD_COPY_G2S(&w, n.snap, CTD_SNAP_BYTES);   // the chance fields sit outside the snapshot and survive
  w.buf_blk = 0xFFFFFFFFu;
  w.g0 = (uint32_t)T.hdr->gid; w.g1 = (uint32_t)(T.hdr->gid >> 32);
  w.tape = nullptr; w.tape_len = 0; w.err = 0;
  CTD_COPY_G2S(T.kn, &n.know, (int)sizeof(CtdKnow));
}

// CFRNode.skip_false_choice (:37-49) on the working game
CTD_HD CTD_NI inline void ctd_skip_false_choice(CtdTree& T) {
  CTD_TREE_SPACES(T);
  CtdWork& w = *T.w;
  CtdKnowSet ks{T.kn, 1};
  int i = 0;
  for (;;) {
    if (w.gflags & 2) return;
    CtdEmit e{T.opts, 1, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e, T.kn);
    if (w.err) return;
    if (e.n != 1) return;
    ++i;
    bool won = ctd_apply(w, T.opts[0], ks);
    if (won || w.err || i > 100) return;
  }
}

// allocate a child node from the working game (CFRNode.__init__, :9-33); returns its index or -1
CTD_HD CTD_NI inline int ctd_new_node(CtdTree& T, int parent, int depth) {
  CTD_TREE_SPACES(T);
  CtdTreeHdr& h = *T.hdr;
  if (h.n_nodes >= h.max_nodes) { h.status |= CTD_TREE_EPOOL; return -1; }
  ctd_skip_false_choice(T);
  if (T.w->err || T.kn->err) { h.status |= CTD_TREE_EENGINE; }
  int idx = (int)h.n_nodes++;
  CtdNode& n = T.nodes[idx];
  n.parent = parent;
  n.depth = (uint16_t)depth;
  n.player = T.w->player;
  n.flags = (uint8_t)((T.w->state == 0 ? CTD_NF_ROLE_PICK : 0) | ((T.w->gflags & 2) ? CTD_NF_TERMINAL : 0));
  n.n_children = 0; n.child_cap = 0; n.child_off = 0; n.arr_off = 0; n.visits = 0; n.pad0 = 0;
  // n.game (the 256-byte packed form) is written by the export pass only (ctd_node_pack)
  CTD_LOOP for (int i = 0; i < 6; ++i) { n.V[i] = 0.0; n.P[i] = 0.0; n.pred[i] = 0.f; }
  ctd_node_store(T, n);
  return idx;
}

CTD_HD inline bool ctd_reserve(CtdTree& T, CtdNode& n, uint32_t kids, uint32_t doubles) {
  CtdTreeHdr& h = *T.hdr;
  if (h.child_used + kids > h.child_cap || h.arr_used + doubles > h.arr_cap) { h.status |= CTD_TREE_EPOOL; return false; }
  n.child_off = h.child_used; n.child_cap = kids; h.child_used += kids;
  n.arr_off = h.arr_used; h.arr_used += doubles;
#if defined(__CUDA_ARCH__)
  if (__activemask() == 0xFFFFFFFFu) {   // converged warp: the lanes split the zero-fill
    for (uint32_t i = threadIdx.x & 31u; i < doubles; i += 32u) T.arr[n.arr_off + i] = 0.0;
    __syncwarp();
    return true;
  }
#endif
  CTD_LOOP for (uint32_t i = 0; i < doubles; ++i) T.arr[n.arr_off + i] = 0.0;
  return true;
}

// the option as stored after carry_out: take_from_hand+build gets its replica rewritten (game/option_functions.py:317)
CTD_HD inline uint64_t ctd_carried_form(const CtdWork& w, uint64_t d) {
  if (CTD_OPT_KIND(d) == CTD_K_TAKE_FROM_HAND && CTD_OPT_BUILD(d)) {
    int p = CTD_OPT_PERP(d), t = CTD_OPT_CARD_A(d);
    int rep = ctd_count_type(w.bld[p], w.n_bld[p], t);
    d &= ~((uint64_t)0xF << 32);
    d |= ctd_f_replica(rep);
  }
  return d;
}

// Game.used_cards is constant over a tree and read 76 bytes at a time by every determinisation: keep a copy next to the
// working set (shared memory on the device) instead of walking the tree header in HBM.  Call once per (re)attached tree.
CTD_HD inline void ctd_tree_stage_used(CtdTree& T) { ctd_copy16(T.scratch + 256, T.hdr->used_cards, 80); }

// "sample if it is not the same player's turn as in the parent" (:139-140, :157-158)
CTD_HD inline void ctd_maybe_sample(CtdTree& T, const CtdNode& n) {
  bool root = n.parent < 0;
  if (root || T.w->player != T.nodes[n.parent].player) {
    bool role_sample = root ? false : T.nodes[n.parent].gstate != 0;
    ctd_sample_private(*T.w, *T.kn, T.scratch + 256, role_sample, T.scratch);
  }
}

// CFRNode.expand (:93-179)
CTD_HD CTD_NI inline void ctd_expand(CtdTree& T, int ni) {
  CTD_TREE_SPACES(T);
  CtdNode& n = T.nodes[ni];
  CtdWork& w = *T.w;
  CtdKnowSet ks{T.kn, 1};
  const int viewer = T.hdr->viewer;
  if (n.gstate == 0 && n.n_children == 0) {
    // expand_role_pick (:102-131): ten uniformly random role-pick phases, the stored option is the last pick
    n.flags |= CTD_NF_ROLE_PICK;
    if (!ctd_reserve(T, n, 10, 180)) return;
    CTD_LOOP for (int rep = 0; rep < 10; ++rep) {
      ctd_node_load(T, n);
      uint64_t d = 0;
      while (w.state != 1 && !w.err) {
        CtdEmit e{T.opts, CTD_MCCFR_OPT_CAP, 0, 0xFFFFFFFFu, 0};
        ctd_enumerate(w, e, T.kn);
        if (e.n == 0) { w.err |= CTD_ERR_REF_RAISE; break; }
        d = T.opts[ctd_randbelow(w, e.n)];
        ctd_apply(w, d, ks);
      }
      int ci = ctd_new_node(T, ni, n.depth + 1);
      if (ci < 0) return;
      T.children[n.child_off + n.n_children] = CtdChild{d, (uint32_t)ci, 0};
      ++n.n_children;
    }
  } else if (n.player == viewer && n.n_children == 0) {
    // expand_for_original_player (:133-151): one child per legal option
    ctd_node_load(T, n);
    CtdEmit e{T.opts, CTD_MCCFR_OPT_CAP, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e, T.kn);
    if (e.n > CTD_MCCFR_OPT_CAP) { T.hdr->status |= CTD_TREE_EOPTS; return; }
    if (e.n == 0) { T.hdr->status |= CTD_TREE_EENGINE; return; }
    // the reference enumerates on the node's own game (:134): the Scholar's list shrinks there, and every child is a copy of that
    if (w.state == 9) CTD_COPY_S2G(n.snap, &w, CTD_SNAP_BYTES);
    const uint32_t K = e.n;
    if (!ctd_reserve(T, n, K, 3 * K)) return;
    // the option list must survive the children's own enumerations: park it in the child table
    CTD_LOOP for (uint32_t i = 0; i < K; ++i) T.children[n.child_off + i] = CtdChild{T.opts[i], 0, 0};
    CTD_LOOP for (uint32_t i = 0; i < K; ++i) {
      ctd_node_load(T, n);
      ctd_maybe_sample(T, n);
      uint64_t d = ctd_carried_form(w, T.children[n.child_off + i].desc);
      ctd_apply(w, d, ks);
      int ci = ctd_new_node(T, ni, n.depth + 1);
      if (ci < 0) return;
      T.children[n.child_off + i] = CtdChild{d, (uint32_t)ci, 0};
      ++n.n_children;
    }
  } else if (n.player != viewer && n.n_children < 10) {
    // expand_for_opponents (:153-179): one uniformly sampled option, kept only if it is new
    if (n.child_cap == 0 && !ctd_reserve(T, n, 10, 30)) return;
    ctd_node_load(T, n);
    ctd_maybe_sample(T, n);
    CtdEmit e{T.opts, CTD_MCCFR_OPT_CAP, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e, T.kn);
    if (e.n == 0) { T.hdr->status |= CTD_TREE_EENGINE; return; }
    uint32_t pick = ctd_randbelow(w, e.n);
    uint64_t d;
    if (pick < CTD_MCCFR_OPT_CAP) d = T.opts[pick];
    else { CtdEmit e2{T.opts, 0, 0, pick, 0}; ctd_enumerate(w, e2, T.kn); d = e2.got; }
    d = ctd_carried_form(w, d);
    ctd_apply(w, d, ks);
    bool seen = false;
    CTD_LOOP for (uint32_t i = 0; i < n.n_children; ++i) seen |= T.children[n.child_off + i].desc == d;
    if (!seen) {
      int ci = ctd_new_node(T, ni, n.depth + 1);
      if (ci < 0) return;
      T.children[n.child_off + n.n_children] = CtdChild{d, (uint32_t)ci, 0};
      ++n.n_children;
    }
  }
  if (w.err || T.kn->err) T.hdr->status |= CTD_TREE_EENGINE;
}

// arrays of a node: vector nodes R[K] s[K] C[K] with K = child_cap; role-pick nodes [6][10] each
CTD_HD inline double* ctd_R(CtdTree& T, const CtdNode& n) { return T.arr + n.arr_off; }
CTD_HD inline double* ctd_S(CtdTree& T, const CtdNode& n) {
  return T.arr + n.arr_off + ((n.flags & CTD_NF_ROLE_PICK) ? 60 : n.child_cap);
}
CTD_HD inline double* ctd_C(CtdTree& T, const CtdNode& n) {
  return T.arr + n.arr_off + 2 * ((n.flags & CTD_NF_ROLE_PICK) ? 60 : n.child_cap);
}

// CFRNode.update_strategy (:292-319)
CTD_HD CTD_NI inline void ctd_update_strategy(CtdTree& T, int ni) {
  CTD_TREE_SPACES(T);
  CtdNode& n = T.nodes[ni];
  const int K = (int)n.n_children;
  if (K == 0) return;  // empty arrays: numpy no-ops
  double *R = ctd_R(T, n), *S = ctd_S(T, n), *C = ctd_C(T, n);
  const double log13 = 0.26236426446749106;  // np.log(1.3)
  if (!(n.flags & CTD_NF_ROLE_PICK)) {
    double tot = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) { S[a] = exp(-R[a] * log13); tot += S[a]; }
    if (tot > 0.0) { CTD_LOOP for (int a = 0; a < K; ++a) S[a] = S[a] / tot; }
    else { CTD_LOOP for (int a = 0; a < K; ++a) S[a] = 1.0 / K; }
    double cs = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) { C[a] += S[a]; cs += C[a]; }
    CTD_LOOP for (int a = 0; a < K; ++a) C[a] = C[a] / cs;
  } else {
    // normalised over the PLAYER axis (axis=0), then C renormalised over all 60 entries
    CTD_LOOP for (int a = 0; a < 10; ++a) {
      double tot = 0.0;
      CTD_LOOP for (int p = 0; p < 6; ++p) { S[p * 10 + a] = exp(-R[p * 10 + a] * log13); tot += S[p * 10 + a]; }
      CTD_LOOP for (int p = 0; p < 6; ++p) S[p * 10 + a] = tot > 1e-8 ? S[p * 10 + a] / tot : 1.0 / 6.0;
    }
    double cs = 0.0;
    CTD_LOOP for (int i = 0; i < 60; ++i) { C[i] += S[i]; cs += C[i]; }
    CTD_LOOP for (int i = 0; i < 60; ++i) C[i] = C[i] / cs;
  }
}

// CFRNode.action_choice (:67-91), non-live: sample a child from the (weighted) cumulative strategy.
// Inverse CDF on one uniform draw: cdf = cumsum(p); cdf /= cdf[-1]; first index with u < cdf.
CTD_HD CTD_NI inline int ctd_action_choice(CtdTree& T, int ni) {
  CTD_TREE_SPACES(T);
  CtdNode& n = T.nodes[ni];
  const int K = (int)n.n_children;
  double* C = ctd_C(T, n);
  double* cdf = (double*)T.opts;  // K <= CTD_MCCFR_OPT_CAP doubles; the option buffer is free here
  if (!(n.flags & CTD_NF_ROLE_PICK)) {
    double cs = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) cs += C[a];
    double run = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) { run += C[a] / cs; cdf[a] = run; }
  } else {
    // weighted_average_strategy (:51-65): weight 6-i for the i-th picker, divided by sum(order) = 15
    double avg[10];
    double s = 0.0;
    CTD_LOOP for (int a = 0; a < 10; ++a) {
      double v = 0.0;
      CTD_LOOP for (int i = 0; i < 6; ++i) v += C[n.order[i] * 10 + a] * (double)(6 - i);
      avg[a] = v / 15.0;
    }
    CTD_LOOP for (int a = 0; a < 10; ++a) s += avg[a];
    double run = 0.0;
    CTD_LOOP for (int a = 0; a < 10; ++a) { run += (s == 0.0 ? 1.0 / 10 : avg[a] / s); cdf[a] = run; }
  }
  const double last = cdf[K - 1];
  const double u = ctd_uniform(*T.w);
  int i = 0;
  while (i < K - 1 && cdf[i] / last <= u) ++i;
  return (int)T.children[n.child_off + i].node;
}

// CFRNode.backpropagate + update_regrets (:231-256, :276-290), iterative instead of recursive
CTD_HD CTD_NI inline void ctd_backpropagate(CtdTree& T, int ni, const double reward[6]) {
  CTD_TREE_SPACES(T);
  const bool training = T.hdr->training, model = T.hdr->has_model;
  for (int cur = ni; cur >= 0; cur = T.nodes[cur].parent) {
    CtdNode& n = T.nodes[cur];
    double vs = 0.0;
    CTD_LOOP for (int i = 0; i < 6; ++i) vs += n.V[i];
    if (training || vs == 0.0 || !model) CTD_LOOP for (int i = 0; i < 6; ++i) n.V[i] += reward[i];
    vs = 0.0;
    CTD_LOOP for (int i = 0; i < 6; ++i) vs += n.V[i];
    CTD_LOOP for (int i = 0; i < 6; ++i) n.P[i] = n.V[i] / vs;
    ++n.visits;
    const int K = (int)n.n_children;
    if (K == 0) continue;
    double* R = ctd_R(T, n);
    if (!(n.flags & CTD_NF_ROLE_PICK)) {
      const int pl = n.player;
      double m = -1e300;
      CTD_LOOP for (int a = 0; a < K; ++a) {
        double v = T.nodes[T.children[n.child_off + a].node].P[pl];
        m = v > m ? v : m;
      }
      CTD_LOOP for (int a = 0; a < K; ++a) R[a] += m - T.nodes[T.children[n.child_off + a].node].P[pl];
    } else {
      // max over PLAYERS (axis=0 after the transpose, :248-251)
      CTD_LOOP for (int a = 0; a < 10; ++a) {
        const double* cp = T.nodes[T.children[n.child_off + a].node].P;
        double m = cp[0];
        CTD_LOOP for (int p = 1; p < 6; ++p) m = cp[p] > m ? cp[p] : m;
        CTD_LOOP for (int p = 0; p < 6; ++p) R[p * 10 + a] += m - cp[p];
      }
    }
  }
}

// CFRNode.cfr_train (:187-205) / cfr_pred without the model call (:207-229 needs pred, see ctd_kernels.cu).
// Initialise the tree from the working game (root state + knowledge already in T.w / T.kn).
CTD_HD CTD_NI inline void ctd_tree_init(CtdTree& T, uint32_t max_nodes, uint32_t child_cap, uint32_t arr_cap, int viewer,
                                        uint64_t gid, bool training, bool has_model) {
  
  CtdTreeHdr& h = *T.hdr;
  h.n_nodes = 0; h.max_nodes = max_nodes; h.child_used = 0; h.child_cap = child_cap; h.arr_used = 0; h.arr_cap = arr_cap;
  h.status = 0; h.iterations = 0; h.rng_draws = 0; h.viewer = (uint8_t)viewer; h.training = training; h.has_model = has_model;
  h.phase = 0; h.cur_node = 0; h.gid = gid;
  ctd_new_node(T, -1, 0);  // the root constructor runs skip_false_choice on the caller's game (:19-20)
  if (T.nodes[0].flags & CTD_NF_TERMINAL) h.status |= CTD_TREE_TERMINAL_ROOT;
  h.rng_draws = T.w->draws;
}

// run `iters` iterations of the pure-MCCFR loop; returns the node the walk is standing on
CTD_HD CTD_NI inline void ctd_cfr_train(CtdTree& T, uint32_t iters) {
  CTD_TREE_SPACES(T);
  CtdTreeHdr& h = *T.hdr;
  if (h.status & CTD_TREE_TERMINAL_ROOT) return;
  ctd_expand(T, 0);
  int node = 0;
  for (uint32_t it = 0; it < iters && !(h.status & ~CTD_TREE_TERMINAL_ROOT); ++it) {
    ctd_update_strategy(T, node);
    node = ctd_action_choice(T, node);
    if (T.nodes[node].flags & CTD_NF_TERMINAL) {
      double reward[6] = {0, 0, 0, 0, 0, 0};
      reward[T.nodes[node].winner] = 1.0;
      ctd_backpropagate(T, node, reward);
      ctd_update_strategy(T, node);
      node = 0;
    } else {
      ctd_expand(T, node);
    }
    ++h.iterations;
  }
  ctd_update_strategy(T, 0);
  h.rng_draws = T.w->draws;
}

// ------------------------------------------------------------------------------------------ deep MCCFR
#define CTD_FEATURES 418
#define CTD_FEATURES_PAD 448 /* row stride of the feature matrix: multiple of 64 for the tensor-core tiles */

// Game.encode_game (game/game.py:34-128, game/deck.py:75-89) from the working record.  The confirmed-role block
// reads known_roles[current player][seat].confirmed, which is the same for every observer (conf_mask).
// Role-pick nodes are encoded with player_id forced to 5 (algorithms/deep_mccfr.py:120-123).
CTD_HD CTD_NI inline void ctd_encode_game(const CtdWork& w, const CtdKnow& k, int player, float* f) {
  CTD_ASSUME_SHARED(&w);   // (k is a global record in ctd_k_encode)
#if defined(__CUDA_ARCH__)
  if (__activemask() == 0xFFFFFFFFu) {   // converged warp (search kernels): the lanes split the zero-fill
    for (int i = threadIdx.x & 31; i < CTD_FEATURES_PAD; i += 32) f[i] = 0.f;
    __syncwarp();
  } else
#endif
  CTD_LOOP for (int i = 0; i < CTD_FEATURES_PAD; ++i) f[i] = 0.f;
  CTD_LOOP for (int r = 0; r < 8; ++r) f[r * 3 + w.variant[r]] = 1.f;
  CTD_LOOP for (int p = 0; p < 6; ++p) {
    if (w.role[p] < 8 && ((k.conf_mask >> p) & 1)) f[24 + p * 8 + w.role[p]] = 1.f;
    f[72 + p] = (float)ctd_count_points(w, p);
    f[78 + p] = (float)w.gold[p];
    f[84 + p] = (float)w.n_hand[p];
    CTD_LOOP for (int i = 0; i < w.n_bld[p]; ++i) {
      int c = w.bld[p][i];
      f[90 + p * 40 + ctd_ctype(c)] += 1.f;
      f[330 + p * 5 + ctd_csuit(c)] += 1.f;
    }
  }
  f[360 + player] = 1.f;
  f[366 + w.state] = 1.f;
  f[377] = (w.gflags & 1) ? 1.f : 0.f;
  CTD_LOOP for (int r = 0; r < 8; ++r) {
    int rp = w.rprops[r];
    if (rp & CTD_RP_DEAD) f[378 + r * 5 + 0] = 1.f;
    if (rp & CTD_RP_WARRANT) f[378 + r * 5 + 1] = 1.f;
    if (rp & CTD_RP_POSSESSED) f[378 + r * 5 + 2] = 1.f;
    if (rp & CTD_RP_ROBBED) f[378 + r * 5 + 3] = 1.f;
    if (rp & CTD_RP_BLACKMAIL) f[378 + r * 5 + 4] = 1.f;
  }
}

// CFRNode.cfr_pred (algorithms/deep_mccfr.py:207-229) as a resumable walk.  When the walk reaches a node deeper than
// max_depth whose value is not cached it writes the node's features to `feat` (CTD_FEATURES_PAD floats), expands
// the node, and returns true: the caller evaluates the value model on the batch of all waiting trees and calls again
// with `pred` = model_reward_weights * square_and_normalize(model(features)) (:126,:147,:178; train_utils.py:143-145).
// Returns false when all iterations are done.
// Returns 0 = finished, 1 = waiting for the value of the leaf in `feat`, 2 = yielded after `budget` iterations of this
// call (the walk resumes from hdr.cur_node; bounding a wave this way keeps trees that never reach the depth limit from
// holding up the batch evaluation every other tree is waiting for).
enum { CTD_PRED_DONE = 0, CTD_PRED_WAIT = 1, CTD_PRED_YIELD = 2 };
CTD_HD CTD_NI inline int ctd_cfr_pred_advance(CtdTree& T, uint32_t iters, uint32_t max_depth, float* feat, const float* pred,
                                              uint32_t budget = 0xFFFFFFFFu) {
  CTD_TREE_SPACES(T);
  CtdTreeHdr& h = *T.hdr;
  if (h.phase == 3 || (h.status & CTD_TREE_TERMINAL_ROOT)) { h.phase = 3; return CTD_PRED_DONE; }
  T.w->draws = h.rng_draws;
  T.w->buf_blk = 0xFFFFFFFFu;
  if (h.phase == 0) {
    ctd_expand(T, 0);
    h.cur_node = 0;
    h.phase = 1;
  } else if (h.phase == 2) {
    CtdNode& n = T.nodes[h.cur_node];
    double reward[6];
    CTD_LOOP for (int i = 0; i < 6; ++i) { n.pred[i] = pred[i]; reward[i] = (double)pred[i]; }
    n.flags |= CTD_NF_HAS_PRED;
    ctd_backpropagate(T, (int)h.cur_node, reward);
    ctd_update_strategy(T, (int)h.cur_node);
    h.cur_node = 0;
    h.phase = 1;
    ++h.iterations;
  }
  int node = (int)h.cur_node;
  uint32_t done = 0;
  while (h.iterations < iters && !(h.status & ~CTD_TREE_TERMINAL_ROOT)) {
    if (done++ >= budget) {   // phase stays 1: walking
      h.cur_node = (uint32_t)node;
      h.rng_draws = T.w->draws;
      return CTD_PRED_YIELD;
    }
    ctd_update_strategy(T, node);
    node = ctd_action_choice(T, node);
    CtdNode& n = T.nodes[node];
    if (n.depth > max_depth && !(n.flags & CTD_NF_TERMINAL)) {
      if (!(n.flags & CTD_NF_HAS_PRED)) {
        ctd_node_load(T, n);
        ctd_encode_game(*T.w, *T.kn, n.gstate == 0 ? 5 : n.player, feat);
        ctd_expand(T, node);
        h.cur_node = (uint32_t)node;
        h.phase = 2;
        h.rng_draws = T.w->draws;
        return CTD_PRED_WAIT;
      }
      ctd_expand(T, node);
      double reward[6];
      CTD_LOOP for (int i = 0; i < 6; ++i) reward[i] = (double)n.pred[i];
      ctd_backpropagate(T, node, reward);
      ctd_update_strategy(T, node);
      node = 0;
    } else if (n.flags & CTD_NF_TERMINAL) {
      double reward[6] = {0, 0, 0, 0, 0, 0};
      reward[n.winner] = 1.0;
      ctd_backpropagate(T, node, reward);
      ctd_update_strategy(T, node);
      node = 0;
    } else {
      ctd_expand(T, node);
    }
    ++h.iterations;
  }
  ctd_update_strategy(T, 0);
  h.cur_node = 0;
  h.phase = 3;
  h.rng_draws = T.w->draws;
  return CTD_PRED_DONE;
}

// fill the packed game record of every node (export only)
CTD_HD CTD_NI inline void ctd_tree_pack_nodes(CtdTree& T) {
  
  const uint32_t n = T.hdr->n_nodes;
  CTD_LOOP for (uint32_t i = 0; i < n; ++i) ctd_node_pack(T, T.nodes[i]);
}

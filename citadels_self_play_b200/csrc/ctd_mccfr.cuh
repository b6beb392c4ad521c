// ctd_mccfr.cuh -- the reference's MCCFR tree search (algorithms/deep_mccfr.py `CFRNode`) on a flat node pool.
//
// One tree is private to one root state (algorithms/deep_mccfr.py:27-29) and is grown by sequential dependent
// sampling, so a tree is owned by one warp and many trees run side by side.  The reference never refuses a root (its
// trees are Python objects on the heap), so tree memory is not a fixed block per root: every tree takes what it needs
// from ONE arena in HBM shared by all trees of a launch (a bump pointer advanced with atomicAdd, 64-byte units):
//
//   CtdTreeHdr hdrs[n_roots]      256 B each: counters, status, the chunk table of the tree's nodes
//   arena                         node chunks  : chunk 0 holds the first n0 = 2^k nodes of a tree (k chosen from the
//                                                iteration budget so that the typical tree never leaves it), chunk c >= 1
//                                                the nodes [n0 << (c-1), n0 << c) -- node i of chunk 0 is one multiply
//                                                away, later chunks go through the table in the header
//                                 node arrays  : one allocation per expanded node: CtdChild[child_cap] then 3 x K doubles
//                                                (R | s | C; role-pick nodes 3 x 6 x 10, stored [player][child] like the
//                                                reference after its transposes, :129-131); small ones are cut from
//                                                16 KB slabs the tree owns, so a tree's arrays stay together
//
// CtdNode = 160 B header (parent, depth, player, flags, V[6], P[6], pred[6]) + the 592 B knowledge block of the searching
// player + a verbatim snapshot of the working record.  Everything numeric is fp64 like the reference's numpy arrays.
// Trees leave the device through ctd_tree_export: a compact block (header | nodes with their 256-byte packed game
// records | children | arrays) whose offsets are local to the block.
//
// Chance: one Philox stream per tree (stream word 1, keyed by (seed, root id)); the mapping of the reference's
// random calls onto it is written out in oracle/mccfr_oracle.py and is the same here.
#pragma once
#include <math.h>
#include "ctd_engine.cuh"

#define CTD_MCCFR_OPT_CAP 4096 /* descriptors in the per-warp option buffer (HBM scratch).  Longer lists (classic Magician hands
                                  reach ~1600 options, the Cardinal ~8000) are enumerated straight into the arena */

enum { CTD_NF_ROLE_PICK = 1, CTD_NF_TERMINAL = 2, CTD_NF_HAS_PRED = 4 };
// tree status bits.  TERMINAL_ROOT and REF_RAISE are outcomes the reference has too (run_mccfr raises ValueError on a terminal
// root; an exception inside the rules code propagates out of cfr_train); EPOOL and EENGINE are this engine's own limits.
enum { CTD_TREE_OK = 0, CTD_TREE_TERMINAL_ROOT = 1, CTD_TREE_EPOOL = 2, CTD_TREE_EENGINE = 4, CTD_TREE_EOPTS = 8,
       CTD_TREE_REF_RAISE = 16 };

struct CtdNode {
  int32_t parent;
  uint16_t depth;
  uint8_t player;      // current_player_id after skip_false_choice
  uint8_t flags;
  uint32_t n_children;
  uint32_t child_cap;
  uint32_t child_off;  // arena offset (64-byte units) of CtdChild[child_cap] followed by the node's doubles; export: index of the first child
  uint32_t arr_off;    // export only: index of the node's first double in the block's array section
  uint32_t visits;
  uint32_t pad0;
  double V[6];         // node_value
  double P[6];         // winning_probabilities
  float pred[6];       // pred_node_value (deep MCCFR)
  uint8_t order[6];    // game.turn_orders_for_roles (role-pick nodes weight their strategy by it)
  uint8_t gstate;      // game.gamestate.state
  int8_t winner;       // game winner (terminal nodes)
  CtdKnow know;
  uint8_t snap[CTD_SNAP_BYTES];  // the working record verbatim: node <-> shared memory is a plain vector copy
};
static_assert(sizeof(CtdNode) == 160 + 592 + CTD_SNAP_BYTES, "CtdNode layout");
static_assert(offsetof(CtdNode, know) % 16 == 0 && offsetof(CtdNode, snap) % 16 == 0 && sizeof(CtdNode) % 16 == 0, "CtdNode alignment");

// a node as it leaves the device (ctd_tree_export): the same 160-byte header, the 256-byte packed game record, the knowledge block
struct CtdNodeOut {
  uint8_t head[160];
  ctd_state game;
  CtdKnow know;
};
static_assert(sizeof(CtdNodeOut) == 160 + 256 + 592, "CtdNodeOut layout");

struct CtdChild {
  uint64_t desc;
  uint32_t node;
  uint32_t pad;
};

#define CTD_TREE_MAX_CHUNKS 24
#define CTD_ARENA_UNIT 64u
#define CTD_SLAB_UNITS 256u /* 16 KB */
struct CtdTreeHdr {
  uint32_t n_nodes;
  uint32_t n0_log2;      // chunk 0 holds 2^n0_log2 nodes
  uint32_t child_used;   // child slots reserved so far (what an export block needs)
  uint32_t arr_used;     // doubles reserved so far
  uint32_t status;
  uint32_t iterations;
  uint32_t rng_draws;
  uint8_t viewer;        // original_player_id
  uint8_t training;
  uint8_t has_model;
  uint8_t phase;         // deep MCCFR walk: 0 not started, 1 walking, 2 waiting for a leaf value, 3 finished
  uint64_t gid;
  uint32_t cur_node;     // node the walk stands on (deep MCCFR is resumed across kernel launches)
  uint32_t slab_off;     // the tree's current slab for small allocations (arena units) and what is left of it
  uint32_t slab_left;
  uint32_t pad0;
  uint8_t* arena;        // base of the arena this tree lives in
  uint8_t used_cards[80];  // Game.used_cards in deal order (game/game.py:424), 76 used; constant over the tree
  uint32_t chunk[CTD_TREE_MAX_CHUNKS];  // arena offset of node chunk c, 0 = not allocated yet
  uint32_t pad1[4];
};
static_assert(sizeof(CtdTreeHdr) == 256 && offsetof(CtdTreeHdr, used_cards) % 16 == 0, "CtdTreeHdr layout");

// header of an export block (the first 128 bytes; numpy mirror: layout.TREE_HDR_DTYPE)
struct CtdTreeHdrOut {
  uint32_t n_nodes, max_nodes, child_used, child_cap, arr_used, arr_cap, status, iterations, rng_draws;
  uint8_t viewer, training, has_model, phase;
  uint64_t gid;
  uint8_t used_cards[76];
  uint32_t cur_node;
};
static_assert(sizeof(CtdTreeHdrOut) == 128, "CtdTreeHdrOut layout");
CTD_HD inline size_t ctd_tree_export_bytes(uint32_t n_nodes, uint32_t child_used, uint32_t arr_used) {
  return sizeof(CtdTreeHdrOut) + (size_t)n_nodes * sizeof(CtdNodeOut) + (size_t)child_used * sizeof(CtdChild) +
         (size_t)arr_used * sizeof(double);
}

// the arena all trees of a launch allocate from
struct CtdArena {
  uint8_t* base;
  unsigned long long* used;  // bump pointer, CTD_ARENA_UNIT units; starts at 1 so that offset 0 means "none"
  unsigned long long cap;    // units
};

#define CTD_TREE_SCRATCH 416
#define CTD_ACT_FLOATS 1024
// ValueOnlyNN(418, 512) in eval mode with BatchNorm folded into fc1 / fc2, weights transposed to [in][out] (ctd_set_value_model)
struct CtdValueNet {
  const float *w1t, *b1, *w2t, *b2, *w3t, *b3, *w4t, *b4;  // w1t [448][512], w2t [512][256], w3t [256][128], w4t [128][6]
  float weight;                                            // model_reward_weights
};
// a tree plus the on-chip working set of the warp that grows it
struct CtdTree {
  CtdTreeHdr* hdr;
  CtdNode* nodes0;  // chunk 0
  uint32_t n0;      // nodes in chunk 0
  uint8_t* abase;   // == hdr->arena
  CtdArena ar;
  CtdWork* w;       // working game (shared memory on the device)
  CtdKnow* kn;      // working knowledge of the viewer
  uint64_t* opts;   // CTD_MCCFR_OPT_CAP descriptors
  uint8_t* scratch; // CTD_TREE_SCRATCH bytes, 16-byte aligned: [0,256) determinisation scratch, [256,336) Game.used_cards staged on chip,
                    // [336,416) occurrence index of every card of used_cards among the cards of its type
  ctd_state* stage; // 16-byte aligned staging record (shared memory on the device)
  const struct CtdValueNet* vnet;  // deep MCCFR, fused mode: the warp evaluates its own leaves (null: the walk hands leaves to the caller)
  float* act;       // fused mode: CTD_ACT_FLOATS floats of shared memory for the features and the activations
  bool walk_settled; // the last ctd_new_node left the game at a real choice or at its end (not cut by the forced-move limit)
};

// address spaces of a tree's parts on the device: working set in shared memory, the tree block in HBM
#define CTD_TREE_SPACES(T)                                                                                  \
  CTD_ASSUME_SHARED((T).w); CTD_ASSUME_SHARED((T).kn); CTD_ASSUME_SHARED((T).stage);                         \
  CTD_ASSUME_GLOBAL((T).hdr); CTD_ASSUME_GLOBAL((T).nodes0); CTD_ASSUME_GLOBAL((T).abase);                    \
  CTD_ASSUME_GLOBAL((T).opts)

// Device code runs either on one lane (export / target walks, ctd_kernels.cu) or with the whole warp converged on the same scalar
// code (the search kernels).  In the search units the second is a fact of the build (CTD_SEARCH_UNIT), elsewhere it is tested.
#if defined(__CUDA_ARCH__)
#ifdef CTD_SEARCH_UNIT
#pragma nv_diag_suppress 128   /* the one-lane fallbacks behind `if (CTD_CONVERGED()) { ...; return; }` are unreachable here, on purpose */
#define CTD_CONVERGED() true
#define CTD_ACTIVE_MASK() 0xFFFFFFFFu
#else
#define CTD_CONVERGED() (__activemask() == 0xFFFFFFFFu)
#define CTD_ACTIVE_MASK() __activemask()
#endif
#endif

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
  const unsigned m = CTD_ACTIVE_MASK();
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
  const unsigned m = CTD_ACTIVE_MASK();
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

// IEEE fp64 division and exp are ~40 / ~100 instructions inline at every use; the search's arithmetic is a few dozen of them per
// iteration, its instruction footprint is what it pays for (profiles/README.md): one out-of-line copy of each
#ifndef CTD_MATH_ATTR
#define CTD_MATH_ATTR CTD_NI
#endif
CTD_HD CTD_MATH_ATTR inline double ctd_ddiv(double a, double b) { return a / b; }
CTD_HD CTD_MATH_ATTR inline double ctd_dexp(double x) { return exp(x); }

// ------------------------------------------------------------------------------------------ tree memory
CTD_HD inline int ctd_clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
// node i of a tree: chunk 0 directly, later chunks through the table in the header
CTD_HD CTD_NI inline CtdNode* ctd_node_far(const CtdTree& T, uint32_t i) {
  const uint32_t c = 32u - (uint32_t)ctd_clz32(i >> T.hdr->n0_log2);   // >= 1; chunk c holds [n0 << (c-1), n0 << c)
  return (CtdNode*)(T.abase + (size_t)T.hdr->chunk[c] * CTD_ARENA_UNIT) + (i - (T.n0 << (c - 1)));
}
CTD_HD inline CtdNode& ctd_node(const CtdTree& T, uint32_t i) { return i < T.n0 ? T.nodes0[i] : *ctd_node_far(T, i); }
CTD_HD inline CtdChild* ctd_kids(const CtdTree& T, const CtdNode& n) { return (CtdChild*)(T.abase + (size_t)n.child_off * CTD_ARENA_UNIT); }

// `units` arena units, or 0 when the arena is exhausted.  On the device the active lanes of the warp (one lane, or all 32
// converged on the same scalar code) make ONE allocation between them.
CTD_HD inline uint32_t ctd_arena_alloc(CtdTree& T, uint32_t units) {
  unsigned long long off;
#if defined(__CUDA_ARCH__)
  const unsigned m = CTD_ACTIVE_MASK();
  const int leader = __ffs(m) - 1;
  off = 0;
  if ((int)(threadIdx.x & 31) == leader) off = atomicAdd(T.ar.used, (unsigned long long)units);
  off = __shfl_sync(m, off, leader);
#else
  off = *T.ar.used;
  *T.ar.used += units;
#endif
  if (off + units > T.ar.cap || off + units > 0xFFFFFFFFull) return 0;
  return (uint32_t)off;
}
// small allocations come out of a slab the tree owns
CTD_HD CTD_NI inline uint32_t ctd_tree_alloc(CtdTree& T, size_t bytes) {
  CtdTreeHdr& h = *T.hdr;
  const uint32_t units = (uint32_t)((bytes + CTD_ARENA_UNIT - 1) / CTD_ARENA_UNIT);
  if (units > CTD_SLAB_UNITS / 4) return ctd_arena_alloc(T, units);
  uint32_t left = h.slab_left, off = h.slab_off;
  if (left < units) {
    off = ctd_arena_alloc(T, CTD_SLAB_UNITS);
    if (off == 0) return 0;
    left = CTD_SLAB_UNITS;
  }
#if defined(__CUDA_ARCH__)
  __syncwarp(CTD_ACTIVE_MASK());   // every lane has read the old slab state before any lane writes the new one
#endif
  h.slab_off = off + units;
  h.slab_left = left - units;
  return off;
}
// attach the working set to a tree block (the header was initialised by ctd_tree_init or by an earlier launch)
CTD_HD inline void ctd_tree_attach(CtdTree& T, CtdTreeHdr* hdr, const CtdArena& ar) {
  T.hdr = hdr;
  T.ar = ar;
  T.abase = hdr->arena;
  T.n0 = 1u << hdr->n0_log2;
  T.nodes0 = (CtdNode*)(T.abase + (size_t)hdr->chunk[0] * CTD_ARENA_UNIT);
}

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
    else k.hk[i].flags &= (uint32_t)CTD_HK_WIZARD;   // clears CTD_HK_USED (flags is a two-bit field)
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
  // first occurrence per removal: the k-th card of a type (deal order) goes exactly when k < pending removals of that type.
  // used_cards[80 + i] holds that k (ctd_tree_stage_used: used_cards is constant over a tree), so every card decides by itself
  int nu = 0;
  const uint8_t* occ = used_cards + 80;
#if defined(__CUDA_ARCH__)
  if (CTD_CONVERGED()) {   // converged warp (search kernels): 32 cards at a time, ballot + prefix count compaction
    const int lane = threadIdx.x & 31;
    __syncwarp();
    CTD_LOOP for (int base = 0; base < 96; base += 32) {
      const int i = base + lane;
      const int c = i < 76 ? used_cards[i] : 0xFF;   // 0xFF pads the 66-card deal of Game(preset=False)
      const bool keep = c != 0xFF && occ[i] >= cnt[ctd_ctype(c)];
      const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
      if (keep) unknown[nu + __popc(m & ((1u << lane) - 1u))] = (uint8_t)c;
      nu += __popc(m);
    }
    __syncwarp();
  } else
#endif
  CTD_LOOP for (int i = 0; i < 76; ++i) {
    int c = used_cards[i];
    if (c == 0xFF) break;  // Game(preset=False) plays with 66 cards
    if (occ[i] >= cnt[ctd_ctype(c)]) unknown[nu++] = (uint8_t)c;
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
  ctd_shuffle_bytes(w, unknown, nu);
  int uh = 0;
  {
    const int take = need < nu ? need : nu;
#if defined(__CUDA_ARCH__)
    if (CTD_CONVERGED()) {   // converged warp: a plain parallel copy
      const int base = w.n_deck;
      __syncwarp();
      for (int i = threadIdx.x & 31; i < take; i += 32) w.deck[base + i] = unknown[i];
      __syncwarp();
    } else
#endif
    CTD_LOOP for (int i = 0; i < take; ++i) w.deck[w.n_deck + i] = unknown[i];
    w.n_deck = (uint8_t)(w.n_deck + take);
    uh = take;
  }
  // (3b) sample_warrants_and_blackmails (:321-336): which of the flagged ranks carries the real one is re-rolled,
  // blackmails first
  CTD_LOOP for (int pass = 0; pass < 2; ++pass) {
    const int mask = pass == 0 ? CTD_RP_BLACKMAIL : CTD_RP_WARRANT, sh = pass == 0 ? 5 : 1;
    uint8_t* keys = scratch + 192;
    int nk = 0;
    CTD_LOOP for (int r = 0; r < 8; ++r)
      if (w.rprops[r] & mask) keys[nk++] = (uint8_t)r;
    if (nk == 0) continue;
    ctd_shuffle_bytes(w, keys, nk);
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
#ifndef CTD_NODE_MOVE_ATTR
#define CTD_NODE_MOVE_ATTR CTD_NI   /* one copy of the 2.3 KB record moves: the search kernels are bound by their instruction footprint */
#endif
CTD_HD CTD_NODE_MOVE_ATTR inline void ctd_node_store(CtdTree& T, CtdNode& n) {
  const CtdWork& w = *T.w;
  CTD_COPY_S2G(n.snap, &w, CTD_SNAP_BYTES);
  CTD_COPY_S2G(&n.know, T.kn, (int)sizeof(CtdKnow));
  CTD_LOOP for (int i = 0; i < 6; ++i) n.order[i] = w.order[i];
  n.gstate = w.state;
  n.winner = w.winner;
}
CTD_HD CTD_NODE_MOVE_ATTR inline void ctd_node_load(CtdTree& T, const CtdNode& n) {
  // chance state lives in the working record and must survive a load
  CtdWork& w = *T.w;
  CTD_COPY_G2S(&w, n.snap, CTD_SNAP_BYTES);   // the chance fields sit outside the snapshot and survive
  w.buf_blk = 0xFFFFFFFFu;
  w.g0 = (uint32_t)T.hdr->gid; w.g1 = (uint32_t)(T.hdr->gid >> 32);
  w.tape = nullptr; w.tape_len = 0; w.err = 0;
  CTD_COPY_G2S(T.kn, &n.know, (int)sizeof(CtdKnow));
}

// CFRNode.skip_false_choice (:37-49) on the working game
CTD_HD CTD_NI inline bool ctd_skip_false_choice(CtdTree& T) {   // false: stopped by the 100-move limit with a forced move pending
  CTD_TREE_SPACES(T);
  CtdWork& w = *T.w;
  CtdKnowSet ks{T.kn, 1};
  int i = 0;
  for (;;) {
    if (w.gflags & 2) return true;
    CtdEmit e{nullptr, 0, 0, 0, 0};   // count, and keep the first option
    ctd_enumerate(w, e, T.kn);
    if (w.err) return true;
    if (e.n != 1) return true;
    ++i;
    bool won = ctd_apply(w, e.got, ks);
    if (won || w.err) return true;
    if (i > 100) return false;
  }
}

// allocate a child node from the working game (CFRNode.__init__, :9-33); returns its index or -1
// settled: the working game is known to stand at a real choice already (a twin of the node built just before), the forced-move
// walk would change nothing and is skipped
CTD_HD CTD_NI inline int ctd_new_node(CtdTree& T, int parent, int depth, bool settled = false) {
  CTD_TREE_SPACES(T);
  CtdTreeHdr& h = *T.hdr;
  const uint32_t idx = h.n_nodes;
  if (idx >= T.n0) {   // beyond chunk 0: the first node of a chunk allocates it
    const uint32_t c = 32u - (uint32_t)ctd_clz32(idx >> h.n0_log2);
    if (c >= CTD_TREE_MAX_CHUNKS) { h.status |= CTD_TREE_EPOOL; return -1; }
    if (idx == (T.n0 << (c - 1))) {
      const size_t bytes = (size_t)(T.n0 << (c - 1)) * sizeof(CtdNode);
      const uint32_t off = ctd_arena_alloc(T, (uint32_t)((bytes + CTD_ARENA_UNIT - 1) / CTD_ARENA_UNIT));
      if (off == 0) { h.status |= CTD_TREE_EPOOL; return -1; }
#if defined(__CUDA_ARCH__)
      __syncwarp(CTD_ACTIVE_MASK());
#endif
      h.chunk[c] = off;
    }
  }
  if (!settled) T.walk_settled = ctd_skip_false_choice(T);
  if ((T.w->err | T.kn->err) & CTD_ERR_OVERFLOW) h.status |= CTD_TREE_EENGINE;
  if ((T.w->err | T.kn->err) & ~CTD_ERR_OVERFLOW) h.status |= CTD_TREE_REF_RAISE;
  h.n_nodes = idx + 1;
  CtdNode& n = ctd_node(T, idx);
  n.parent = parent;
  n.depth = (uint16_t)depth;
  n.player = T.w->player;
  n.flags = (uint8_t)((T.w->state == 0 ? CTD_NF_ROLE_PICK : 0) | ((T.w->gflags & 2) ? CTD_NF_TERMINAL : 0));
  n.n_children = 0; n.child_cap = 0; n.child_off = 0; n.arr_off = 0; n.visits = 0; n.pad0 = 0;
  CTD_LOOP for (int i = 0; i < 6; ++i) { n.V[i] = 0.0; n.P[i] = 0.0; n.pred[i] = 0.f; }
  ctd_node_store(T, n);
  return (int)idx;
}

// room for `kids` children and `doubles` array entries of node n (one allocation: children first, then the doubles)
CTD_HD inline bool ctd_reserve(CtdTree& T, CtdNode& n, uint32_t kids, uint32_t doubles) {
  CtdTreeHdr& h = *T.hdr;
  const uint32_t off = ctd_tree_alloc(T, (size_t)kids * sizeof(CtdChild) + (size_t)doubles * sizeof(double));
  if (off == 0) { h.status |= CTD_TREE_EPOOL; return false; }
  n.child_off = off; n.child_cap = kids;
  h.child_used += kids;
  h.arr_used += doubles;
  double* a = (double*)(ctd_kids(T, n) + kids);
#if defined(__CUDA_ARCH__)
  if (CTD_CONVERGED()) {   // converged warp: the lanes split the zero-fill
    for (uint32_t i = threadIdx.x & 31u; i < doubles; i += 32u) a[i] = 0.0;
    __syncwarp();
    return true;
  }
#endif
  CTD_LOOP for (uint32_t i = 0; i < doubles; ++i) a[i] = 0.0;
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
CTD_HD inline void ctd_stage_used(uint8_t* u, const uint8_t* used_cards) {   // u: 160 bytes
  CTD_LOOP for (int i = 0; i < 80; ++i) u[i] = i < 76 ? used_cards[i] : 0xFF;
  // occurrence index of every card among the cards of its type, in deal order (what the determinisation's filter tests)
  uint8_t seen[40];
  CTD_LOOP for (int t = 0; t < 40; ++t) seen[t] = 0;
  CTD_LOOP for (int i = 0; i < 80; ++i) {
    const int c = u[i];
    u[80 + i] = c == 0xFF ? 0 : seen[ctd_ctype(c)]++;
  }
}
CTD_HD inline void ctd_tree_stage_used(CtdTree& T) { ctd_stage_used(T.scratch + 256, T.hdr->used_cards); }

// "sample if it is not the same player's turn as in the parent" (:139-140, :157-158)
CTD_HD inline void ctd_maybe_sample(CtdTree& T, const CtdNode& n) {
  bool root = n.parent < 0;
  if (root || T.w->player != ctd_node(T, n.parent).player) {
    bool role_sample = root ? false : ctd_node(T, n.parent).gstate != 0;
    ctd_sample_private(*T.w, *T.kn, T.scratch + 256, role_sample, T.scratch);
  }
}

// One uniformly random legal option without a list buffer: count, draw, enumerate again and keep the k-th.  The Seer's and the
// Scholar's enumerations are not pure (fresh shuffles, a shrinking list: game/agent_functions.py:332-361, :462-470), so the second
// pass starts from the state the first one started from and regenerates the same list; what is left behind is exactly one
// enumeration plus one draw, as in `options = game.get_options_from_state(); choice(options)` (run_utils.py:38-39).
struct CtdEnumSave {
  uint32_t draws, tape_pos;
  uint8_t n_seven, seven[7];
};
CTD_HD inline CtdEnumSave ctd_enum_save(const CtdWork& w) {
  CtdEnumSave s;
  s.draws = w.draws; s.tape_pos = w.tape_pos; s.n_seven = w.n_seven;
  CTD_LOOP for (int i = 0; i < 7; ++i) s.seven[i] = w.seven[i];
  return s;
}
CTD_HD inline uint64_t ctd_enum_select(CtdWork& w, const CtdKnow* kn, const CtdEnumSave& s, uint32_t k) {
  const uint32_t draws1 = w.draws, tape1 = w.tape_pos;
  w.draws = s.draws; w.tape_pos = s.tape_pos; w.buf_blk = 0xFFFFFFFFu; w.n_seven = s.n_seven;
  CTD_LOOP for (int i = 0; i < 7; ++i) w.seven[i] = s.seven[i];
  CtdEmit e2{nullptr, 0, 0, k, 0};
  ctd_enumerate(w, e2, kn);
  w.draws = draws1; w.tape_pos = tape1; w.buf_blk = 0xFFFFFFFFu;
  return e2.got;
}
// The legal options of the working game, materialised.  Lists of up to CTD_SMALL_OPTS descriptors (mean 4, p99 31 in preset
// games) stay on chip in the warp's staging record, which the search does not use otherwise; longer ones are enumerated again
// into the warp's HBM buffer (from the state the first pass started from: the Seer's give-back lists draw chance and can be long).
// *list receives where the first min(n, capacity) descriptors are.
#define CTD_SMALL_OPTS 32
static_assert(CTD_SMALL_OPTS * sizeof(uint64_t) <= sizeof(ctd_state), "the small option buffer is the staging record");
CTD_HD inline uint32_t ctd_list_options(CtdTree& T, const uint64_t** list) {
  CtdWork& w = *T.w;
  uint64_t* small = (uint64_t*)T.stage;
  const CtdEnumSave sv = ctd_enum_save(w);
  CtdEmit e{small, CTD_SMALL_OPTS, 0, 0xFFFFFFFFu, 0};
  ctd_enumerate(w, e, T.kn);
  *list = small;
  if (e.n <= CTD_SMALL_OPTS) return e.n;
  w.draws = sv.draws; w.tape_pos = sv.tape_pos; w.buf_blk = 0xFFFFFFFFu; w.n_seven = sv.n_seven;
  CTD_LOOP for (int i = 0; i < 7; ++i) w.seven[i] = sv.seven[i];
  CtdEmit e2{T.opts, CTD_MCCFR_OPT_CAP, 0, 0xFFFFFFFFu, 0};
  ctd_enumerate(w, e2, T.kn);
  *list = T.opts;
  return e2.n;
}

// CFRNode.expand (:93-179)
CTD_HD CTD_NI inline void ctd_expand(CtdTree& T, int ni) {
  CTD_TREE_SPACES(T);
  CtdNode& n = ctd_node(T, ni);
  CtdWork& w = *T.w;
  CtdKnowSet ks{T.kn, 1};
  const int viewer = T.hdr->viewer;
  if (n.gstate == 0 && n.n_children == 0) {
    // expand_role_pick (:102-131): ten uniformly random role-pick phases, the stored option is the last pick
    n.flags |= CTD_NF_ROLE_PICK;
    if (!ctd_reserve(T, n, 10, 180)) return;
    CtdChild* kids = ctd_kids(T, n);
    CTD_LOOP for (int rep = 0; rep < 10; ++rep) {
      ctd_node_load(T, n);
      uint64_t d = 0;
      while (w.state != 1 && !w.err) {
        const uint64_t* list;
        const uint32_t n_opts = ctd_list_options(T, &list);
        if (n_opts == 0) { w.err |= CTD_ERR_REF_RAISE; break; }
        d = list[ctd_randbelow(w, n_opts)];   // role-pick lists hold at most eight options
        ctd_apply(w, d, ks);
      }
      int ci = ctd_new_node(T, ni, n.depth + 1);
      if (ci < 0) return;
      kids[n.n_children] = CtdChild{d, (uint32_t)ci, 0};
      ++n.n_children;
    }
  } else if (n.player == viewer && n.n_children == 0) {
    // expand_for_original_player (:133-151): one child per legal option
    ctd_node_load(T, n);
    const uint64_t* list;
    const uint32_t K = ctd_list_options(T, &list);
#ifdef CTD_HOST_DEBUG
    if (K == 0) fprintf(stderr, "expand: no options for the searching player, state %d err %d\n", (int)w.state, (int)w.err);
#endif
    if (K == 0) { T.hdr->status |= w.err ? CTD_TREE_REF_RAISE : CTD_TREE_EENGINE; return; }
    // the reference enumerates on the node's own game (:134): the Scholar's list shrinks there, and every child is a copy of that
    if (w.state == 9) CTD_COPY_S2G(n.snap, &w, CTD_SNAP_BYTES);
    if (!ctd_reserve(T, n, K, 3 * K)) return;
    CtdChild* kids = ctd_kids(T, n);
    // the option list must survive the children's own enumerations: park it in the child table
    if (K <= CTD_MCCFR_OPT_CAP) {
      CTD_LOOP for (uint32_t i = 0; i < K; ++i) kids[i] = CtdChild{list[i], 0, 0};
    } else {
      // a list longer than the option buffer (pure enumerations only: the Magician's discards, the Cardinal's exchanges): enumerate
      // once more straight into the node's still unused doubles, move it over, zero the doubles again
      uint64_t* big = (uint64_t*)(kids + K);
      CtdEmit e2{big, K, 0, 0xFFFFFFFFu, 0};
      ctd_enumerate(w, e2, T.kn);
      CTD_LOOP for (uint32_t i = 0; i < K; ++i) kids[i] = CtdChild{big[i], 0, 0};
      CTD_LOOP for (uint32_t i = 0; i < K; ++i) big[i] = 0;   // 0.0
    }
    // Every discard_and_draw option of the Magician has the same effect (carry_out_magicking ignores the option's cards,
    // game/option_functions.py:295-300) and a classic Magician's turn lists ~1600 of them.  When a child was built without
    // determinisation and without a single chance draw, building the next one from the same parent is the same computation
    // bit for bit: the working set already holds its result, only the node has to be stored.
    const bool resampled = n.parent < 0 || n.player != ctd_node(T, n.parent).player;
    bool prev_same_effect = false;
    CTD_LOOP for (uint32_t i = 0; i < K; ++i) {
      const uint64_t d0 = kids[i].desc;
      uint64_t d;
      if (prev_same_effect && CTD_OPT_KIND(d0) == CTD_K_DISCARD_AND_DRAW) {
        d = d0;
      } else {
        ctd_node_load(T, n);
        ctd_maybe_sample(T, n);
        d = ctd_carried_form(w, d0);
        const uint32_t draws0 = w.draws;
        ctd_apply(w, d, ks);
        prev_same_effect = false;
        int ci = ctd_new_node(T, ni, n.depth + 1);
        if (ci < 0) return;
        kids[i] = CtdChild{d, (uint32_t)ci, 0};
        ++n.n_children;
        prev_same_effect = !resampled && CTD_OPT_KIND(d0) == CTD_K_DISCARD_AND_DRAW && w.draws == draws0 && !w.err && !T.kn->err &&
                           w.state != 8 && w.state != 9 && T.walk_settled;
        continue;
      }
      int ci = ctd_new_node(T, ni, n.depth + 1, true);   // no forced moves: the state stands where the twin's walk stopped
      if (ci < 0) return;
      kids[i] = CtdChild{d, (uint32_t)ci, 0};
      ++n.n_children;
    }
  } else if (n.player != viewer && n.n_children < 10) {
    // expand_for_opponents (:153-179): one uniformly sampled option, kept only if it is new
    if (n.child_cap == 0 && !ctd_reserve(T, n, 10, 30)) return;
    CtdChild* kids = ctd_kids(T, n);
    ctd_node_load(T, n);
    ctd_maybe_sample(T, n);
    const uint64_t* list;
    const uint32_t n_opts = ctd_list_options(T, &list);
#ifdef CTD_HOST_DEBUG
    if (n_opts == 0) fprintf(stderr, "expand: no options for an opponent, state %d err %d player %d role %d\n", (int)w.state, (int)w.err, (int)w.player, (int)w.role[w.player]);
#endif
    if (n_opts == 0) { T.hdr->status |= w.err ? CTD_TREE_REF_RAISE : CTD_TREE_EENGINE; return; }
    uint32_t pick = ctd_randbelow(w, n_opts);
    uint64_t d;
    if (pick < CTD_MCCFR_OPT_CAP) d = list[pick];
    else { CtdEmit e2{nullptr, 0, 0, pick, 0}; ctd_enumerate(w, e2, T.kn); d = e2.got; }   // beyond the buffer: pure enumerations only
    d = ctd_carried_form(w, d);
    ctd_apply(w, d, ks);
    bool seen = false;
    CTD_LOOP for (uint32_t i = 0; i < n.n_children; ++i) seen |= kids[i].desc == d;
    if (!seen) {
      int ci = ctd_new_node(T, ni, n.depth + 1);
      if (ci < 0) return;
      kids[n.n_children] = CtdChild{d, (uint32_t)ci, 0};
      ++n.n_children;
    }
  }
  if ((w.err | T.kn->err) & CTD_ERR_OVERFLOW) T.hdr->status |= CTD_TREE_EENGINE;
  if ((w.err | T.kn->err) & ~CTD_ERR_OVERFLOW) T.hdr->status |= CTD_TREE_REF_RAISE;
}

// arrays of a node: vector nodes R[K] s[K] C[K] with K = child_cap; role-pick nodes [6][10] each.  They follow the child table.
CTD_HD inline double* ctd_R(const CtdTree& T, const CtdNode& n) { return (double*)(ctd_kids(T, n) + n.child_cap); }
CTD_HD inline double* ctd_S(const CtdTree& T, const CtdNode& n) {
  return ctd_R(T, n) + ((n.flags & CTD_NF_ROLE_PICK) ? 60 : n.child_cap);
}
CTD_HD inline double* ctd_C(const CtdTree& T, const CtdNode& n) {
  return ctd_R(T, n) + 2 * ((n.flags & CTD_NF_ROLE_PICK) ? 60 : n.child_cap);
}

// CFRNode.update_strategy (:292-319)
CTD_HD CTD_NI inline void ctd_update_strategy(CtdTree& T, int ni) {
  CTD_TREE_SPACES(T);
  CtdNode& n = ctd_node(T, ni);
  const int K = (int)n.n_children;
  if (K == 0) return;  // empty arrays: numpy no-ops
  double *R = ctd_R(T, n), *S = ctd_S(T, n), *C = ctd_C(T, n);
  const double log13 = 0.26236426446749106;  // np.log(1.3)
#if defined(__CUDA_ARCH__)
  // Search kernels (whole warp converged on this code): the element-wise exp / divisions go one element per lane, every sum
  // stays the sequential left-to-right sum of the scalar form -- same operations on the same operands, bit for bit.
  if (CTD_CONVERGED()) {
    const int lane = (int)(threadIdx.x & 31);
    const bool rp = n.flags & CTD_NF_ROLE_PICK;
    const int E = rp ? 60 : K;   // entries of R / S / C
    for (int i = lane; i < E; i += 32) S[i] = ctd_dexp(-R[i] * log13);
    __syncwarp();
    // what an entry is divided by: the sum over the children (vector nodes), over the six players of its column (role-pick nodes)
    double tot = 0.0;
    if (!rp) {
      CTD_LOOP for (int a = 0; a < K; ++a) tot += S[a];
    } else if (lane < 10) {
      CTD_LOOP for (int p = 0; p < 6; ++p) tot += S[p * 10 + lane];
    }
    __syncwarp();
    CTD_LOOP for (int base = 0; base < E; base += 32) {
      const int i = base + lane;
      const double t = rp ? __shfl_sync(0xFFFFFFFFu, tot, i % 10) : tot;
      if (i < E) {
        const bool pos = rp ? t > 1e-8 : t > 0.0;   // otherwise uniform: 1/6 over the players, 1/K over the children
        const double si = ctd_ddiv(pos ? S[i] : 1.0, pos ? t : (rp ? 6.0 : (double)K));
        S[i] = si;
        C[i] += si;
      }
    }
    __syncwarp();
    double cs = 0.0;
    CTD_LOOP for (int i = 0; i < E; ++i) cs += C[i];
    __syncwarp();
    for (int i = lane; i < E; i += 32) C[i] = ctd_ddiv(C[i], cs);
    __syncwarp();
    return;
  }
#endif
  if (!(n.flags & CTD_NF_ROLE_PICK)) {
    double tot = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) { S[a] = ctd_dexp(-R[a] * log13); tot += S[a]; }
    if (tot > 0.0) { CTD_LOOP for (int a = 0; a < K; ++a) S[a] = ctd_ddiv(S[a], tot); }
    else { CTD_LOOP for (int a = 0; a < K; ++a) S[a] = ctd_ddiv(1.0, (double)K); }
    double cs = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) { C[a] += S[a]; cs += C[a]; }
    CTD_LOOP for (int a = 0; a < K; ++a) C[a] = ctd_ddiv(C[a], cs);
  } else {
    // normalised over the PLAYER axis (axis=0), then C renormalised over all 60 entries
    CTD_LOOP for (int a = 0; a < 10; ++a) {
      double tot = 0.0;
      CTD_LOOP for (int p = 0; p < 6; ++p) { S[p * 10 + a] = ctd_dexp(-R[p * 10 + a] * log13); tot += S[p * 10 + a]; }
      CTD_LOOP for (int p = 0; p < 6; ++p) S[p * 10 + a] = tot > 1e-8 ? ctd_ddiv(S[p * 10 + a], tot) : 1.0 / 6.0;
    }
    double cs = 0.0;
    CTD_LOOP for (int i = 0; i < 60; ++i) { C[i] += S[i]; cs += C[i]; }
    CTD_LOOP for (int i = 0; i < 60; ++i) C[i] = ctd_ddiv(C[i], cs);
  }
}

// CFRNode.action_choice (:67-91), non-live: sample a child from the (weighted) cumulative strategy.
// Inverse CDF on one uniform draw: cdf = cumsum(p); cdf /= cdf[-1]; first index with u < cdf.
CTD_HD CTD_NI inline int ctd_action_choice(CtdTree& T, int ni) {
  CTD_TREE_SPACES(T);
  CtdNode& n = ctd_node(T, ni);
  const int K = (int)n.n_children;
  double* C = ctd_C(T, n);
  const CtdChild* kids = ctd_kids(T, n);
  if (!(n.flags & CTD_NF_ROLE_PICK)) {
    double cs = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) cs += C[a];
    // cdf[a] = running sum of C[a] / cs; the draw is compared with cdf[a] / cdf[K-1]
    double* cdf = K <= CTD_SMALL_OPTS ? (double*)T.stage : (K <= CTD_MCCFR_OPT_CAP ? (double*)T.opts : nullptr);
#if defined(__CUDA_ARCH__)
    if (cdf != nullptr && CTD_CONVERGED()) {   // search kernels: divisions one per lane, sums and the draw as below
      const int lane = (int)(threadIdx.x & 31);
      __syncwarp();
      for (int a = lane; a < K; a += 32) cdf[a] = ctd_ddiv(C[a], cs);
      __syncwarp();
      double run = 0.0;
      CTD_LOOP for (int a = 0; a < K; ++a) { run += cdf[a]; cdf[a] = run; }
      __syncwarp();
      const double last = cdf[K - 1];
      const double u = ctd_uniform(*T.w);
      int i = K - 1;
      CTD_LOOP for (int base = 0; base < K - 1; base += 32) {
        const int a = base + lane;
        const bool stop = a < K - 1 && !(ctd_ddiv(cdf[a], last) <= u);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, stop);
        if (m) { i = base + __ffs(m) - 1; break; }
      }
      __syncwarp();
      return (int)kids[i].node;
    }
#endif
    if (cdf != nullptr) {
      double run = 0.0;
      CTD_LOOP for (int a = 0; a < K; ++a) { run += ctd_ddiv(C[a], cs); cdf[a] = run; }
      const double last = cdf[K - 1];
      const double u = ctd_uniform(*T.w);
      int i = 0;
      while (i < K - 1 && ctd_ddiv(cdf[i], last) <= u) ++i;
      return (int)kids[i].node;
    }
    // no buffer can hold it (the Cardinal expands thousands of children): the running sum is formed twice, same operations, same order
    double last = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) last += ctd_ddiv(C[a], cs);
    const double u = ctd_uniform(*T.w);
    double run = 0.0;
    int i = 0;
    CTD_LOOP for (; i < K - 1; ++i) {
      run += ctd_ddiv(C[i], cs);
      if (!(ctd_ddiv(run, last) <= u)) break;
    }
    return (int)kids[i].node;
  }
  // weighted_average_strategy (:51-65): weight 6-i for the i-th picker, divided by sum(order) = 15
  double cdf[10];
  double avg[10];
  double s = 0.0;
  CTD_LOOP for (int a = 0; a < 10; ++a) {
    double v = 0.0;
    CTD_LOOP for (int i = 0; i < 6; ++i) v += C[n.order[i] * 10 + a] * (double)(6 - i);
    avg[a] = ctd_ddiv(v, 15.0);
  }
  CTD_LOOP for (int a = 0; a < 10; ++a) s += avg[a];
  double run = 0.0;
  CTD_LOOP for (int a = 0; a < 10; ++a) { run += (s == 0.0 ? 1.0 / 10 : ctd_ddiv(avg[a], s)); cdf[a] = run; }
  const double last = cdf[K - 1];
  const double u = ctd_uniform(*T.w);
  int i = 0;
  while (i < K - 1 && ctd_ddiv(cdf[i], last) <= u) ++i;
  return (int)kids[i].node;
}

// CFRNode.backpropagate + update_regrets (:231-256, :276-290), iterative instead of recursive
CTD_HD CTD_NI inline void ctd_backpropagate(CtdTree& T, int ni, const double reward[6]) {
  CTD_TREE_SPACES(T);
  const bool training = T.hdr->training, model = T.hdr->has_model;
  for (int cur = ni; cur >= 0; cur = ctd_node(T, cur).parent) {
    CtdNode& n = ctd_node(T, cur);
    double vs = 0.0;
    CTD_LOOP for (int i = 0; i < 6; ++i) vs += n.V[i];
    if (training || vs == 0.0 || !model) CTD_LOOP for (int i = 0; i < 6; ++i) n.V[i] += reward[i];
    vs = 0.0;
    CTD_LOOP for (int i = 0; i < 6; ++i) vs += n.V[i];
#if defined(__CUDA_ARCH__)
    if (CTD_CONVERGED()) {   // search kernels: one seat / one child per lane; max is exact in any order
      const int lane = (int)(threadIdx.x & 31);
      __syncwarp();
      if (lane < 6) n.P[lane] = ctd_ddiv(n.V[lane], vs);
      __syncwarp();
      ++n.visits;
      const int K = (int)n.n_children;
      if (K == 0) continue;
      double* R = ctd_R(T, n);
      const CtdChild* kids = ctd_kids(T, n);
      if (!(n.flags & CTD_NF_ROLE_PICK)) {
        const int pl = n.player;
        double m = -1e300;
        CTD_LOOP for (int base = 0; base < K; base += 32) {
          const int a = base + lane;
          double v = a < K ? ctd_node(T, kids[a].node).P[pl] : -1e300;
          if (!(v > -1e300)) v = -1e300;   // `v > m ? v : m` never takes a NaN
          CTD_LOOP for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(0xFFFFFFFFu, v, o);
            v = y > v ? y : v;
          }
          m = v > m ? v : m;
        }
        for (int a = lane; a < K; a += 32) R[a] += m - ctd_node(T, kids[a].node).P[pl];
      } else if (lane < 10) {
        const double* cp = ctd_node(T, kids[lane].node).P;
        double m = cp[0];
        CTD_LOOP for (int p = 1; p < 6; ++p) m = cp[p] > m ? cp[p] : m;
        CTD_LOOP for (int p = 0; p < 6; ++p) R[p * 10 + lane] += m - cp[p];
      }
      __syncwarp();
      continue;
    }
#endif
    CTD_LOOP for (int i = 0; i < 6; ++i) n.P[i] = ctd_ddiv(n.V[i], vs);
    ++n.visits;
    const int K = (int)n.n_children;
    if (K == 0) continue;
    double* R = ctd_R(T, n);
    const CtdChild* kids = ctd_kids(T, n);
    if (!(n.flags & CTD_NF_ROLE_PICK)) {
      const int pl = n.player;
      double m = -1e300;
      CTD_LOOP for (int a = 0; a < K; ++a) {
        double v = ctd_node(T, kids[a].node).P[pl];
        m = v > m ? v : m;
      }
      CTD_LOOP for (int a = 0; a < K; ++a) R[a] += m - ctd_node(T, kids[a].node).P[pl];
    } else {
      // max over PLAYERS (axis=0 after the transpose, :248-251)
      CTD_LOOP for (int a = 0; a < 10; ++a) {
        const double* cp = ctd_node(T, kids[a].node).P;
        double m = cp[0];
        CTD_LOOP for (int p = 1; p < 6; ++p) m = cp[p] > m ? cp[p] : m;
        CTD_LOOP for (int p = 0; p < 6; ++p) R[p * 10 + a] += m - cp[p];
      }
    }
  }
}

// CFRNode.cfr_train (:187-205) / cfr_pred without the model call (:207-229 needs pred, see ctd_kernels.cu).
// Initialise the tree from the working game (root state + knowledge already in T.w / T.kn).
CTD_HD CTD_NI inline void ctd_tree_init(CtdTree& T, CtdTreeHdr* hdr, const CtdArena& ar, uint32_t n0_log2, int viewer,
                                        uint64_t gid, bool training, bool has_model) {
  CtdTreeHdr& h = *hdr;
  h.n_nodes = 0; h.n0_log2 = n0_log2; h.child_used = 0; h.arr_used = 0;
  h.status = 0; h.iterations = 0; h.rng_draws = 0; h.viewer = (uint8_t)viewer; h.training = training; h.has_model = has_model;
  h.phase = 0; h.cur_node = 0; h.gid = gid; h.slab_off = 0; h.slab_left = 0; h.pad0 = 0; h.arena = ar.base;
  CTD_LOOP for (int c = 0; c < CTD_TREE_MAX_CHUNKS; ++c) h.chunk[c] = 0;
  T.hdr = hdr; T.ar = ar; T.abase = ar.base; T.n0 = 1u << n0_log2; T.nodes0 = nullptr;
  const size_t bytes = (size_t)T.n0 * sizeof(CtdNode);
  const uint32_t off = ctd_arena_alloc(T, (uint32_t)((bytes + CTD_ARENA_UNIT - 1) / CTD_ARENA_UNIT));
  if (off == 0) { h.status = CTD_TREE_EPOOL; return; }
  h.chunk[0] = off;
  T.nodes0 = (CtdNode*)(T.abase + (size_t)off * CTD_ARENA_UNIT);
  ctd_new_node(T, -1, 0);  // the root constructor runs skip_false_choice on the caller's game (:19-20)
  if (ctd_node(T, 0).flags & CTD_NF_TERMINAL) h.status |= CTD_TREE_TERMINAL_ROOT;
  h.rng_draws = T.w->draws;
}

// run `iters` iterations of the pure-MCCFR loop.  resume: the tree was grown by an earlier call (root-parallel mode runs the
// loop in rounds: parallel.root_parallel_mccfr); the walk continues from the node it stood on, on the tree's own chance stream.
CTD_HD CTD_NI inline void ctd_cfr_train(CtdTree& T, uint32_t iters, bool resume = false) {
  CTD_TREE_SPACES(T);
  CtdTreeHdr& h = *T.hdr;
  if ((h.status & CTD_TREE_TERMINAL_ROOT) || h.n_nodes == 0) return;
  int node = 0;
  if (!resume) {
    ctd_expand(T, 0);
  } else {
    T.w->draws = h.rng_draws; T.w->buf_blk = 0xFFFFFFFFu;
    node = (int)h.cur_node;
  }
  for (uint32_t it = 0; it < iters && !(h.status & ~CTD_TREE_TERMINAL_ROOT); ++it) {
    ctd_update_strategy(T, node);
    node = ctd_action_choice(T, node);
    if (ctd_node(T, node).flags & CTD_NF_TERMINAL) {
      double reward[6] = {0, 0, 0, 0, 0, 0};
      reward[ctd_node(T, node).winner] = 1.0;
      ctd_backpropagate(T, node, reward);
      ctd_update_strategy(T, node);
      node = 0;
    } else {
      ctd_expand(T, node);
    }
    ++h.iterations;
  }
  ctd_update_strategy(T, 0);
  h.cur_node = (uint32_t)node;
  h.rng_draws = T.w->draws;
}

// CFRNode.action_choice(live=True) at the root once the search is over (:67-75; run_utils.py:82,86): the decision a caller of
// run_mccfr acts on.  Ordinary roots sample a child from the cumulative strategy; a role-pick root goes through
// Game.get_option_from_role_preference (game/game.py:312-317): the `strategy` row of the player to move, which has one entry per
// CHILD, is indexed by the RANKS still on offer (the reference's quirk, kept), normalised, and one of the role-pick options is
// drawn.  The draw is the next one of the tree's own Philox stream (np.random.choice -> inverse CDF on one uniform, as in
// ctd_action_choice).  Returns the option's descriptor; 0 for a terminal root (the reference raises ValueError there).
CTD_HD CTD_NI inline uint64_t ctd_live_choice(CtdTree& T) {
  CTD_TREE_SPACES(T);
  const CtdTreeHdr& h = *T.hdr;
  if (h.n_nodes == 0 || (h.status & ~CTD_TREE_TERMINAL_ROOT)) return 0;
  CtdNode& n = ctd_node(T, 0);
  const int K = (int)n.n_children;
  if (K == 0) return 0;
  CtdWork& w = *T.w;
  if (!(n.flags & CTD_NF_ROLE_PICK)) {
    w.draws = h.rng_draws; w.buf_blk = 0xFFFFFFFFu;
    const double* C = ctd_C(T, n);
    double cs = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) cs += C[a];
    double last = 0.0;
    CTD_LOOP for (int a = 0; a < K; ++a) last += ctd_ddiv(C[a], cs);
    const double u = ctd_uniform(w);
    double run = 0.0;
    int i = 0;
    CTD_LOOP for (; i < K - 1; ++i) {
      run += ctd_ddiv(C[i], cs);
      if (!(ctd_ddiv(run, last) <= u)) break;
    }
    return ctd_kids(T, n)[i].desc;
  }
  ctd_node_load(T, n);   // the root's game: which roles are on offer
  w.draws = h.rng_draws; w.buf_blk = 0xFFFFFFFFu;
  const double* S = ctd_S(T, n) + (int)n.player * 10;
  double sub[8], tot = 0.0;
  int ranks[8], m = 0;
  CTD_LOOP for (int r = 0; r < 8; ++r)
    if ((w.rtc_mask >> r) & 1) { ranks[m] = r; sub[m] = S[r]; tot += S[r]; ++m; }
  if (m == 0) return 0;
  double last = 0.0;
  CTD_LOOP for (int j = 0; j < m; ++j) last += ctd_ddiv(sub[j], tot);
  const double u = ctd_uniform(w);
  double run = 0.0;
  int i = 0;
  CTD_LOOP for (; i < m - 1; ++i) {
    run += ctd_ddiv(sub[i], tot);
    if (!(ctd_ddiv(run, last) <= u)) break;
  }
  return ctd_opt(CTD_K_ROLE_PICK, w.player) | ctd_f_rank(ranks[i]);
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
  if (CTD_CONVERGED()) {   // converged warp (search kernels): the lanes split the zero-fill
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

#if defined(__CUDA_ARCH__)
// ---- the value model evaluated by the warp that needs the value (fused deep MCCFR) ----
// CFRNode.model_inference (algorithms/deep_mccfr.py:364-374): ValueOnlyNN forward in eval mode, square_and_normalize
// (train_utils.py:143-145), times model_reward_weights.  One leaf at a time is a matrix-VECTOR product: there is no tile for a
// tensor core to work on, and batching leaves across trees is what forces the search into waves (every tree waits for the
// slowest walker of its wave).  Here the 32 lanes share the output columns of each layer, the inputs are broadcast from shared
// memory, and both sources of sparsity are used: a feature row has ~70 non-zeros of 418 (one-hot blocks, small counts), and
// about half of the ReLU outputs are zero.  fp32 FMAs in ascending input order: the same arithmetic, term for term, as the fp32
// batch kernel (ctd_k_value_mlp), since skipped terms are exact zeros.  All 32 lanes must be converged.
// act: [0,448) features in; scratch and activations (layout in the body).  Weights stream from L2 (1.5 MB, resident).
template <int NOUT4>   // float4 column groups per lane: layer width = 128 * NOUT4
__device__ __forceinline__ void ctd_vnet_layer(const float* __restrict__ wt, const float* __restrict__ bias, const float* in, int n_in,
                                               uint16_t* list, float* out, int lane) {
  // compact the indices of the non-zero inputs (ballot + prefix count)
  int nnz = 0;
  for (int base = 0; base < n_in; base += 32) {
    const int k = base + lane;
    const bool nz = k < n_in && in[k] != 0.f;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, nz);
    if (nz) list[nnz + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
    nnz += __popc(m);
  }
  __syncwarp();
  float4 acc[NOUT4];
#pragma unroll
  for (int j = 0; j < NOUT4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int width = 128 * NOUT4;
  const float4* w4 = reinterpret_cast<const float4*>(wt) + lane;
  // two inputs per trip: their weight rows are independent loads
  int i = 0;
  for (; i + 1 < nnz; i += 2) {
    const int k0 = list[i], k1 = list[i + 1];
    const float x0 = in[k0], x1 = in[k1];
    float4 a[NOUT4], b[NOUT4];
#pragma unroll
    for (int j = 0; j < NOUT4; ++j) { a[j] = __ldg(w4 + (k0 * width) / 4 + 32 * j); b[j] = __ldg(w4 + (k1 * width) / 4 + 32 * j); }
#pragma unroll
    for (int j = 0; j < NOUT4; ++j) {
      acc[j].x = fmaf(a[j].x, x0, acc[j].x); acc[j].y = fmaf(a[j].y, x0, acc[j].y); acc[j].z = fmaf(a[j].z, x0, acc[j].z); acc[j].w = fmaf(a[j].w, x0, acc[j].w);
      acc[j].x = fmaf(b[j].x, x1, acc[j].x); acc[j].y = fmaf(b[j].y, x1, acc[j].y); acc[j].z = fmaf(b[j].z, x1, acc[j].z); acc[j].w = fmaf(b[j].w, x1, acc[j].w);
    }
  }
  if (i < nnz) {
    const int k0 = list[i];
    const float x0 = in[k0];
#pragma unroll
    for (int j = 0; j < NOUT4; ++j) {
      const float4 a = __ldg(w4 + (k0 * width) / 4 + 32 * j);
      acc[j].x = fmaf(a.x, x0, acc[j].x); acc[j].y = fmaf(a.y, x0, acc[j].y); acc[j].z = fmaf(a.z, x0, acc[j].z); acc[j].w = fmaf(a.w, x0, acc[j].w);
    }
  }
  __syncwarp();   // every lane is done with `in` and `list` (the output may overlap them)
#pragma unroll
  for (int j = 0; j < NOUT4; ++j) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(bias) + lane + 32 * j);
    float4 r;
    r.x = fmaxf(acc[j].x + c.x, 0.f); r.y = fmaxf(acc[j].y + c.y, 0.f); r.z = fmaxf(acc[j].z + c.z, 0.f); r.w = fmaxf(acc[j].w + c.w, 0.f);
    reinterpret_cast<float4*>(out)[lane + 32 * j] = r;
  }
  __syncwarp();
}
static __device__ __noinline__ void ctd_value_inline(const CtdValueNet& m, float* act, float pred[6]) {
  const int lane = threadIdx.x & 31;
  // act (floats): layer 1 reads x [0,448), index list at [448,660), writes h1 [512,1024); layer 2 reads h1, list at [256,512),
  // writes h2 [0,256); layer 3 reads h2, list at [256,384), writes h3 [384,512)
  ctd_vnet_layer<4>(m.w1t, m.b1, act, CTD_FEATURES, reinterpret_cast<uint16_t*>(act + 448), act + 512, lane);
  ctd_vnet_layer<2>(m.w2t, m.b2, act + 512, 512, reinterpret_cast<uint16_t*>(act + 256), act, lane);
  ctd_vnet_layer<1>(m.w3t, m.b3, act, 256, reinterpret_cast<uint16_t*>(act + 256), act + 384, lane);
  // fc4: six outputs, each a sequential fp32 sum over the 128 inputs (lanes 0..5), then weight * y^2 / sum(y^2)
  float y = 0.f;
  if (lane < 6) {
    const float* h3 = act + 384;
    for (int k = 0; k < 128; ++k) y = fmaf(__ldg(m.w4t + k * 6 + lane), h3[k], y);
    y += __ldg(m.b4 + lane);
    y *= y;
  }
  float sq[6], s = 0.f;
#pragma unroll
  for (int o = 0; o < 6; ++o) { sq[o] = __shfl_sync(0xFFFFFFFFu, y, o); s += sq[o]; }
#pragma unroll
  for (int o = 0; o < 6; ++o) pred[o] = m.weight * (sq[o] / s);
  __syncwarp();
}
#endif

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
  if (h.phase == 3 || (h.status & CTD_TREE_TERMINAL_ROOT) || h.n_nodes == 0) { h.phase = 3; return CTD_PRED_DONE; }
  T.w->draws = h.rng_draws;
  T.w->buf_blk = 0xFFFFFFFFu;
  if (h.phase == 0) {
    ctd_expand(T, 0);
    h.cur_node = 0;
    h.phase = 1;
  } else if (h.phase == 2) {
    CtdNode& n = ctd_node(T, h.cur_node);
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
    CtdNode& n = ctd_node(T, node);
    if (n.depth > max_depth && !(n.flags & CTD_NF_TERMINAL)) {
      if (!(n.flags & CTD_NF_HAS_PRED)) {
        ctd_node_load(T, n);
#if defined(__CUDA_ARCH__)
        if (T.vnet != nullptr) {   // fused mode: this warp evaluates the leaf itself and walks on -- no wave, no waiting
          ctd_encode_game(*T.w, *T.kn, n.gstate == 0 ? 5 : n.player, T.act);
          ctd_expand(T, node);
          float pr[6];
          ctd_value_inline(*T.vnet, T.act, pr);
          double reward[6];
          CTD_LOOP for (int i = 0; i < 6; ++i) { n.pred[i] = pr[i]; reward[i] = (double)pr[i]; }
          n.flags |= CTD_NF_HAS_PRED;
          ctd_backpropagate(T, node, reward);
          ctd_update_strategy(T, node);
          node = 0;
          ++h.iterations;
          continue;
        }
#endif
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

// ------------------------------------------------------------------------------------------ CFR roots
// run_utils.create_a_close_to_finished_game / create_a_random_game (run_utils.py:29-72): play game (seed, gid) uniformly
// at random to terminal (T steps), then replay it from its (seed, gid) with the knowledge of all six observers tracked and
// stop at the root:
//   CTD_ROOTS_CLOSE_TO_FINISHED  u uniform in [back_lo, back_hi] (Philox stream word 2), root = step max(0, T - u), then along
//                                the recorded game until the player to move has >= 2 options, at most 100 times
//                                (run_utils.py:44-50 walks games[-m], games[-m+1], ...: m = u + 1); the searching player is
//                                whoever is to move there
//   CTD_ROOTS_RANDOM_GAME        m uniform in [back_lo, back_hi], root = games[-m] = step max(0, T + 1 - m) as it stands
//                                (run_utils.py:52-72); run_mccfr fixes original_player_id = the player to move BEFORE
//                                CFRNode.skip_false_choice advances through forced moves (run_utils.py:80-83,
//                                algorithms/deep_mccfr.py:19-20), so the searching player may be the previous seat: the
//                                root carries the knowledge of that seat and ctd_tree_init does the skipping
// Leaves the root in `w`, the six observers' knowledge in kn6[0..6); returns the searching seat.
CTD_HD inline uint64_t ctd_choose_uniform(CtdWork& w, const CtdKnow* kn) {
  const CtdEnumSave sv = ctd_enum_save(w);
  CtdEmit e{nullptr, 0, 0, 0xFFFFFFFFu, 0};
  ctd_enumerate(w, e, kn);
  if (e.n == 0) return 0;
  const uint32_t k = ctd_randbelow(w, e.n);
  return ctd_enum_select(w, kn, sv, k);
}
CTD_HD CTD_NI inline int ctd_make_root(CtdWork& w, CtdKnow* kn6, uint64_t seed, uint64_t gid, int ruleset, uint32_t back_lo,
                                       uint32_t back_hi, int flavour, uint8_t* used_cards, uint32_t* root_step) {
  // pass 1: length of the game
  ctd_new_game(w, seed, gid, ruleset);
  while (!(w.gflags & 2) && !w.err && w.steps < 4096) {
    uint64_t d = ctd_choose_uniform(w, nullptr);
    if (d == 0) { w.err |= CTD_ERR_REF_RAISE; break; }
    ctd_apply(w, d);
  }
  const uint32_t T = w.steps;
  uint32_t r[4];
  ctd_philox(0u, 2u, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const uint32_t u = back_lo + (uint32_t)(((uint64_t)r[0] * (back_hi - back_lo + 1)) >> 32);
  const uint32_t Tk = flavour == CTD_ROOTS_RANDOM_GAME ? T + 1 : T;
  const uint32_t k = Tk > u ? Tk - u : 0;
  // pass 2: replay with knowledge
  ctd_chance_init(w, seed, gid, 0);
  ctd_deal_preset(w, ruleset, used_cards);
  for (int o = 0; o < 6; ++o) ctd_kn_init(kn6[o], o);
  CtdKnowSet ks{kn6, 6};
  ctd_setup_round(w, ks);
  int limit = 0;
  for (;;) {
    if ((w.gflags & 2) || w.err) break;
    if (flavour == CTD_ROOTS_RANDOM_GAME && w.steps >= k) break;
    const CtdEnumSave sv = ctd_enum_save(w);
    CtdEmit e{nullptr, 0, 0, 0xFFFFFFFFu, 0};
    ctd_enumerate(w, e, &kn6[0]);
    if (e.n == 0) { w.err |= CTD_ERR_REF_RAISE; break; }
    if (w.steps >= k) {  // `while len(options) < 2 and limit < 100` (run_utils.py:46-50); the root keeps what this enumeration did to it
      if (e.n >= 2 || limit >= 100) break;
      ++limit;
    }
    const uint32_t pick = ctd_randbelow(w, e.n);
    ctd_apply(w, ctd_enum_select(w, &kn6[0], sv, pick), ks);
  }
  if (root_step) *root_step = w.steps;
  const int viewer = w.player < 6 ? w.player : 0;
  for (int o = 0; o < 6; ++o) w.err |= kn6[o].err;
  return viewer;
}

// ------------------------------------------------------------------------------------------ export
// The tree as a compact block: CtdTreeHdrOut | CtdNodeOut[n_nodes] | CtdChild[child_used] | double[arr_used], child_off / arr_off
// rewritten as indices into the block's own sections (nodes in index order).  `out` holds ctd_tree_export_bytes(...) bytes.
CTD_HD CTD_NI inline void ctd_tree_export(CtdTree& T, uint8_t* out) {
  const CtdTreeHdr& h = *T.hdr;
  CtdTreeHdrOut* ho = (CtdTreeHdrOut*)out;
  ho->n_nodes = h.n_nodes; ho->max_nodes = h.n_nodes; ho->child_used = h.child_used; ho->child_cap = h.child_used;
  ho->arr_used = h.arr_used; ho->arr_cap = h.arr_used; ho->status = h.status; ho->iterations = h.iterations;
  ho->rng_draws = h.rng_draws; ho->viewer = h.viewer; ho->training = h.training; ho->has_model = h.has_model; ho->phase = h.phase;
  ho->gid = h.gid; ho->cur_node = h.cur_node;
  CTD_LOOP for (int i = 0; i < 76; ++i) ho->used_cards[i] = h.used_cards[i];
  CtdNodeOut* no = (CtdNodeOut*)(out + sizeof(CtdTreeHdrOut));
  CtdChild* co = (CtdChild*)(no + h.n_nodes);
  double* ao = (double*)(co + h.child_used);
  uint32_t coff = 0, aoff = 0;
  CtdWork& w = *T.w;
  CTD_LOOP for (uint32_t i = 0; i < h.n_nodes; ++i) {
    const CtdNode& n = ctd_node(T, i);
    CtdNodeOut& o = no[i];
    ctd_copy16(o.head, &n, 160);
    ctd_copy16(&o.know, &n.know, (int)sizeof(CtdKnow));
    ctd_copy16(&w, n.snap, CTD_SNAP_BYTES);
    w.draws = 0; w.tape_pos = 0; w.steps = 0;
    w.g0 = (uint32_t)h.gid; w.g1 = (uint32_t)(h.gid >> 32);
    ctd_pack(w, T.stage);
    ctd_copy16(&o.game, T.stage, (int)sizeof(ctd_state));
    CtdNode* oh = (CtdNode*)o.head;   // only the 160-byte header part exists behind this pointer
    oh->child_off = coff; oh->arr_off = aoff;
    if (n.child_cap != 0) {
      const CtdChild* kids = ctd_kids(T, n);
      const uint32_t nd = (n.flags & CTD_NF_ROLE_PICK) ? 180u : 3u * n.child_cap;
      const double* a = (const double*)(kids + n.child_cap);
      CTD_LOOP for (uint32_t k = 0; k < n.child_cap; ++k) co[coff + k] = k < n.n_children ? kids[k] : CtdChild{0, 0, 0};
      CTD_LOOP for (uint32_t k = 0; k < nd; ++k) ao[aoff + k] = a[k];
      coff += n.child_cap;
      aoff += nd;
    }
  }
}

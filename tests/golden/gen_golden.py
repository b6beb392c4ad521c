"""Generate the committed golden fixtures from the REAL reference (container-only).

    python -m tests.golden.gen_golden            # writes tests/golden/*.npz

For each game the unmodified reference (/root/reference/game) is played to terminal by the loop of
run_utils.py:37-41, with its shuffles and the option choice driven by PhiloxChance(SEED, gid).
Recorded per step: number of legal options, chosen index, crc32 of the 228 reference-visible bytes of
the packed state (tests/golden/ref_harness.ref_pack) and crc32 of the little-endian descriptor list
(ref_harness.ref_descriptors).  Recorded per game: the chance tape (every shuffle's permutation), the
terminal record, steps.  `*_full.npz` additionally keeps every state and descriptor of a few games.

`ref_outcomes_preset.npz` is a plain sample of reference outcomes under its own Mersenne-Twister
(random.seed(gid)), used for the distribution gate (SURVEY.md 8(d) parity 2).
"""
import os
import sys
import zlib
import numpy as np
from multiprocessing import Pool

SEED = 0xC17ADE15
HERE = os.path.dirname(os.path.abspath(__file__))


def visible(rec):
    """Reference-visible bytes of a packed record: [0,228) + seer_mask, seven_n (229,230) + seven[] (240..246)."""
    return bytes(rec[:228]) + bytes(rec[229:231]) + bytes(rec[240:247])


def _one(args):
    gid, ruleset, full = args
    from oracle.philox import PhiloxChance, RecordingChance
    from tests.golden import ref_harness as H
    ch = RecordingChance(PhiloxChance(SEED, gid))
    g = H.new_ref_game(ch, ruleset)
    nopt, chosen, hs, ho = [], [], [], []
    states, descs = [], []
    while True:
        H.set_chance(ch)
        opts = g.get_options_from_state()
        d = H.ref_descriptors(opts)
        rec = H.ref_pack(g, ruleset)
        db = np.asarray(d, dtype="<u8").tobytes()
        nopt.append(len(d))
        hs.append(zlib.crc32(visible(rec)))
        ho.append(zlib.crc32(db))
        if full:
            states.append(rec)
            descs.append(d)
        i = ch.randbelow(len(d))
        chosen.append(i)
        if opts[i].carry_out(g):
            break
    final = H.ref_pack(g, ruleset)
    return dict(gid=gid, tape=ch.tape, nopt=nopt, chosen=chosen, hs=hs, ho=ho, final=final, states=states,
                descs=descs)


def _flat(lists, dtype):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in lists])
    flat = np.concatenate([np.asarray(x, dtype=dtype) for x in lists]) if lists else np.zeros(0, dtype)
    return flat, off


def gen_traces(name, ruleset, gids, full=False, procs=8):
    with Pool(procs) as pool:
        res = pool.map(_one, [(g, ruleset, full) for g in gids], chunksize=4)
    tape, tape_off = _flat([r["tape"] for r in res], np.uint8)
    nopt, step_off = _flat([r["nopt"] for r in res], np.uint16)
    chosen, _ = _flat([r["chosen"] for r in res], np.uint16)
    hs, _ = _flat([r["hs"] for r in res], np.uint32)
    ho, _ = _flat([r["ho"] for r in res], np.uint32)
    out = dict(seed=np.uint64(SEED), ruleset=np.int32(ruleset), gids=np.asarray(gids, dtype=np.uint64),
               tape=tape, tape_off=tape_off, step_off=step_off, nopt=nopt, chosen=chosen, state_crc=hs,
               opts_crc=ho, final=np.frombuffer(b"".join(r["final"] for r in res), dtype=np.uint8).reshape(-1, 256))
    if full:
        out["states"] = np.frombuffer(b"".join(b"".join(r["states"]) for r in res), dtype=np.uint8).reshape(-1, 256)
        dl = [d for r in res for d in r["descs"]]
        out["descs"], out["desc_off"] = _flat(dl, np.uint64)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(name, "games", len(gids), "steps", len(nopt), "bytes", os.path.getsize(path))


def _mt_game(args):
    """One game of the unmodified reference under CPython's own Mersenne Twister (random.seed(gid)): shuffles, sample /
    choice / randint of set_random_game and the caller's option choice all draw from `random`."""
    gid, ruleset = args
    import random
    from tests.golden import ref_harness as H
    H.load_reference()

    class MT:
        def perm(self, n):
            idx = list(range(n))
            random.Random.shuffle(random._inst, idx)
            return idx

        def randbelow(self, n):
            return random._inst.randrange(n)
    random.seed(gid)
    g = H.new_ref_game(MT(), ruleset)
    steps = 0
    while True:
        opts = g.get_options_from_state()
        steps += 1
        if opts[random._inst.randrange(len(opts))].carry_out(g):
            break
    return [int(g.rewards.argmax())] + [int(x) for x in g.points] + [steps]


def gen_outcomes(name, n, ruleset=0, procs=8):
    with Pool(procs) as pool:
        res = pool.map(_mt_game, [(g, ruleset) for g in range(n)], chunksize=16)
    a = np.asarray(res, dtype=np.int16)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, winner=a[:, 0].astype(np.int8), points=a[:, 1:7].astype(np.int8), steps=a[:, 7])
    print(name, "games", n, "mean steps", a[:, 7].mean(), "bytes", os.path.getsize(path))


def _mccfr_one(args):
    """One CFR root + tree from the REAL reference (CFRNode.cfr_train) under Philox streams 0 (game) / 1 (tree)."""
    gid, ruleset, back_hi, iters = args[:4]
    deep = args[4] if len(args) > 4 else 0
    import zlib
    from copy import deepcopy
    from oracle.philox import PhiloxChance
    from oracle import citadels_oracle as O
    from tests.golden import ref_harness as H
    H.patch_cfr_chance()
    from algorithms.deep_mccfr import CFRNode
    # root as ctd_make_roots defines it (see oracle/mccfr_oracle.make_root), played on the reference
    ch = PhiloxChance(SEED, gid)
    g = H.new_ref_game(ch, ruleset)
    T = 0
    while True:
        H.set_chance(ch)
        o = g.get_options_from_state()
        T += 1
        if o[ch.randbelow(len(o))].carry_out(g):
            break
    u = PhiloxChance(SEED, gid, stream=2).randbelow(back_hi + 1)
    k = max(0, T - u)
    ch = PhiloxChance(SEED, gid)
    g = H.new_ref_game(ch, ruleset)
    steps = limit = 0
    while not g.terminal:
        H.set_chance(ch)
        o = g.get_options_from_state()
        if steps >= k:
            if len(o) >= 2 or limit >= 100:
                break
            limit += 1
        o[ch.randbelow(len(o))].carry_out(g)
        steps += 1
    viewer = g.gamestate.player_id
    root = H.ref_pack(g, ruleset)
    know = H.ref_pack_know(g, viewer)
    used = bytes(H.card_code(c) for c in g.used_cards.cards)
    used += b"\xff" * (76 - len(used))   # Game(preset=False) plays with 66 cards
    out = dict(gid=gid, root=root, know=know, used=used, root_step=steps, terminal=bool(g.terminal))
    nodes = []
    if not g.terminal:
        H.set_chance(PhiloxChance(SEED, gid, stream=1))
        if deep:
            # config 4: ValueOnlyNN(418, 512), default torch init under manual_seed(0), eval(), depth limit `deep`
            import torch
            from algorithms.models import ValueOnlyNN
            from citadels_self_play_b200.value_model import ValueOnlyNN as Mirror
            torch.set_num_threads(1)
            torch.manual_seed(0)
            model = ValueOnlyNN(418, 512).eval()
            torch.manual_seed(0)
            mirror = Mirror(418, 512).eval()
            assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), mirror.state_dict().values()))
            rn = CFRNode(g, original_player_id=viewer, model=model, training=False, device="cpu")
            rn.cfr_pred(max_iterations=iters, max_depth=deep)
        else:
            rn = CFRNode(g, original_player_id=viewer)
            rn.cfr_train(max_iterations=iters)

        def walk(n):
            nodes.append(n)
            for _, c in n.children:
                walk(c)
        walk(rn)
    if nodes and not deep:
        # CFRNode.get_all_targets() of the same tree, viewpoint seats from Philox stream 3
        H.set_chance(PhiloxChance(SEED, gid, stream=3))
        tg = nodes[0].get_all_targets()
        out["t_feat"] = [t[0].numpy() for t in tg]
        out["t_opts"] = [t[1].numpy()[0] for t in tg]
        out["t_val"] = [t[2].numpy() for t in tg]
        out["t_dist"] = [t[3].numpy() for t in tg]
    out["nchild"] = [len(n.children) for n in nodes]
    out["desc"] = [H.ref_descriptors([c[0]])[0] if c[0].name not in ("discard_and_draw", "cardinal_exchange") else 0 for n in nodes for c in n.children]
    out["V"] = [list(n.node_value) for n in nodes]
    out["P"] = [list(n.winning_probabilities) for n in nodes]
    out["R"] = [float(x) for n in nodes for x in np.asarray(n.cumulative_regrets, dtype=float).ravel()]
    out["S"] = [float(x) for n in nodes for x in np.asarray(n.strategy, dtype=float).ravel()]
    out["C"] = [float(x) for n in nodes for x in np.asarray(n.cumulative_strategy, dtype=float).ravel()]
    out["narr"] = [int(np.asarray(n.cumulative_regrets).size) for n in nodes]
    out["game_crc"] = [zlib.crc32(visible(H.ref_pack(n.game, ruleset))) for n in nodes]
    out["know_crc"] = [zlib.crc32(H.ref_pack_know(n.game, viewer)) for n in nodes]
    return out


def _live_one(args):
    """run_utils.run_mccfr(game, model, max_iterations) of the REAL reference on one root: the option it decides on
    (root.action_choice(live=True)[1]) and the root's arrays.  Root as ctd_make_roots defines it; chance: Philox stream 0 for
    the game, stream 1 for the search AND the decision after it (the reference draws both from the global RNGs in that order)."""
    gid, ruleset, back_hi, iters, deep = args
    from oracle.philox import PhiloxChance
    from tests.golden import ref_harness as H
    H.patch_cfr_chance()
    import run_utils
    ch = PhiloxChance(SEED, gid)
    g = H.new_ref_game(ch, ruleset)
    T = 0
    while True:
        H.set_chance(ch)
        o = g.get_options_from_state()
        T += 1
        if o[ch.randbelow(len(o))].carry_out(g):
            break
    u = PhiloxChance(SEED, gid, stream=2).randbelow(back_hi + 1)
    k = max(0, T - u)
    ch = PhiloxChance(SEED, gid)
    g = H.new_ref_game(ch, ruleset)
    steps = limit = 0
    while not g.terminal:
        H.set_chance(ch)
        o = g.get_options_from_state()
        if steps >= k:
            if len(o) >= 2 or limit >= 100:
                break
            limit += 1
        o[ch.randbelow(len(o))].carry_out(g)
        steps += 1
    viewer = g.gamestate.player_id
    out = dict(gid=gid, root=H.ref_pack(g, ruleset), know=H.ref_pack_know(g, viewer),
               used=bytes(H.card_code(c) for c in g.used_cards.cards).ljust(76, b"\xff"), terminal=bool(g.terminal),
               live=0, role_pick=False, nchild=0, draws=0)
    if g.terminal:
        return out
    tree = PhiloxChance(SEED, gid, stream=1)
    H.set_chance(tree)
    model = None
    if deep:
        import torch
        from algorithms.models import ValueOnlyNN
        torch.set_num_threads(1)
        torch.manual_seed(0)
        model = ValueOnlyNN(418, 512).eval()
        import algorithms.deep_mccfr as dm
        _init = dm.CFRNode.__init__
        if not getattr(dm.CFRNode, "_cpu_default", False):      # run_mccfr builds CFRNode with device="cuda:0": no GPU here
            def init(self, *a, **kw):
                kw["device"] = "cpu"
                _init(self, *a, **kw)
            dm.CFRNode.__init__ = init
            dm.CFRNode._cpu_default = True
    chosen, root = run_utils.run_mccfr(g, model=model, max_iterations=iters)
    out["draws"] = tree.i
    out["live"] = H.ref_descriptors([chosen])[0]
    out["role_pick"] = bool(root.role_pick_node)
    out["nchild"] = len(root.children)
    return out


def gen_live(name, ruleset, gids, back_hi, iters, deep=0, procs=8):
    with Pool(procs) as pool:
        res = pool.map(_live_one, [(g, ruleset, back_hi, iters, deep) for g in gids], chunksize=2)
    out = dict(seed=np.uint64(SEED), ruleset=np.int32(ruleset), iterations=np.int32(iters), back_hi=np.int32(back_hi),
               max_depth=np.int32(deep), gids=np.asarray(gids, dtype=np.uint64),
               roots=np.frombuffer(b"".join(r["root"] for r in res), dtype=np.uint8).reshape(-1, 256),
               knows=np.frombuffer(b"".join(r["know"] for r in res), dtype=np.uint8).reshape(-1, 592),
               used=np.frombuffer(b"".join(r["used"] for r in res), dtype=np.uint8).reshape(-1, 76),
               terminal=np.asarray([r["terminal"] for r in res], dtype=bool),
               live=np.asarray([r["live"] for r in res], dtype=np.uint64),
               role_pick=np.asarray([r["role_pick"] for r in res], dtype=bool),
               nchild=np.asarray([r["nchild"] for r in res], dtype=np.int32),
               draws=np.asarray([r["draws"] for r in res], dtype=np.int64))
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(name, "roots", len(gids), "role-pick roots", int(out["role_pick"].sum()), "terminal", int(out["terminal"].sum()),
          "bytes", os.path.getsize(path))


def gen_mccfr(name, ruleset, gids, back_hi, iters, procs=8, deep=0):
    with Pool(procs) as pool:
        res = pool.map(_mccfr_one, [(g, ruleset, back_hi, iters, deep) for g in gids], chunksize=1)
    node_off = np.zeros(len(res) + 1, dtype=np.int64)
    node_off[1:] = np.cumsum([len(r["nchild"]) for r in res])
    cat = lambda k, dt: np.concatenate([np.asarray(r[k], dtype=dt).reshape(-1) for r in res])
    out = dict(seed=np.uint64(SEED), ruleset=np.int32(ruleset), iterations=np.int32(iters), back_hi=np.int32(back_hi),
               gids=np.asarray(gids, dtype=np.uint64), max_depth=np.int32(deep),
               roots=np.frombuffer(b"".join(r["root"] for r in res), dtype=np.uint8).reshape(-1, 256),
               knows=np.frombuffer(b"".join(r["know"] for r in res), dtype=np.uint8).reshape(-1, 592),
               used=np.frombuffer(b"".join(r["used"] for r in res), dtype=np.uint8).reshape(-1, 76),
               root_step=np.asarray([r["root_step"] for r in res], dtype=np.int32),
               terminal=np.asarray([r["terminal"] for r in res], dtype=bool),
               node_off=node_off, nchild=cat("nchild", np.int32), desc=cat("desc", np.uint64),
               V=cat("V", np.float64).reshape(-1, 6), P=cat("P", np.float64).reshape(-1, 6), narr=cat("narr", np.int32),
               R=cat("R", np.float64), S=cat("S", np.float64), C=cat("C", np.float64),
               game_crc=cat("game_crc", np.uint32), know_crc=cat("know_crc", np.uint32))
    if not deep:
        tcount = [len(r.get("t_feat", [])) for r in res]
        out["t_off"] = np.concatenate([[0], np.cumsum(tcount)]).astype(np.int64)
        feats = [f for r in res for f in r.get("t_feat", [])]
        out["t_feat"] = np.asarray(feats, dtype=np.float32).reshape(-1, 418)
        out["t_val"] = np.asarray([v for r in res for v in r.get("t_val", [])], dtype=np.float64).reshape(-1, 6)
        ks = [len(o) for r in res for o in r.get("t_opts", [])]
        out["t_k"] = np.asarray(ks, dtype=np.int32)
        out["t_opts"] = (np.concatenate([o for r in res for o in r.get("t_opts", [])]).astype(np.float32)
                         if ks else np.zeros((0, 131), np.float32))
        out["t_dist"] = (np.concatenate([d for r in res for d in r.get("t_dist", [])]).astype(np.float64)
                         if ks else np.zeros(0))
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(name, "roots", len(gids), "nodes", int(node_off[-1]), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "traces"):
        gen_traces("preset_traces.npz", 0, list(range(1000)))
        gen_traces("classic_traces.npz", 1, list(range(100000, 100300)))
        gen_traces("preset_full.npz", 0, list(range(2000, 2006)), full=True)
        gen_traces("classic_full.npz", 1, list(range(102000, 102004)), full=True)
    if what in ("all", "traces", "random"):
        gen_traces("random_traces.npz", 2, list(range(200000, 200500)))
        gen_traces("random_full.npz", 2, list(range(202000, 202006)), full=True)
    if what in ("all", "mccfr"):
        gen_mccfr("mccfr_preset.npz", 0, list(range(3000, 3024)), 20, 200)
        gen_mccfr("mccfr_preset_deep_back.npz", 0, list(range(3100, 3112)), 300, 200)
        gen_mccfr("mccfr_classic.npz", 1, list(range(103000, 103008)), 60, 200)
    if what in ("all", "mccfr", "mccfr_random"):
        gen_mccfr("mccfr_random.npz", 2, list(range(203000, 203016)), 80, 200)
    if what in ("all", "mccfr2000"):
        gen_mccfr("mccfr_preset_2000it.npz", 0, list(range(3200, 3206)), 45, 2000)
    if what in ("all", "deep"):
        gen_mccfr("deep_mccfr_preset.npz", 0, list(range(4000, 4016)), 120, 200, deep=10)
    if what in ("all", "live"):
        # run_mccfr's decisions (action_choice(live=True)): late-game roots, early roots (many role-pick roots), deluxe rulesets, deep
        gen_live("live_choice_preset.npz", 0, list(range(5000, 5160)), 20, 200)
        gen_live("live_choice_preset_early.npz", 0, list(range(5200, 5296)), 400, 120)
        gen_live("live_choice_classic.npz", 1, list(range(105000, 105024)), 60, 120)
        gen_live("live_choice_random.npz", 2, list(range(205000, 205024)), 80, 120)
        gen_live("live_choice_deep_preset.npz", 0, list(range(5400, 5424)), 200, 200, deep=10)
    if what in ("all", "outcomes"):
        gen_outcomes("ref_outcomes_preset.npz", 20000)
    if what in ("all", "outcomes", "outcomes_bc"):
        gen_outcomes("ref_outcomes_classic.npz", 12000, ruleset=1)
        gen_outcomes("ref_outcomes_random.npz", 12000, ruleset=2)

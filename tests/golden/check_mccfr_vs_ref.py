"""Differential check of the MCCFR oracle against the REAL reference CFRNode (container-only).

    python -m tests.golden.check_mccfr_vs_ref [n_roots] [iterations] [first_gid] [ruleset]

Root = preset game played randomly (Philox stream 0) for a drawn number of steps; then
CFRNode(...).cfr_train(iterations) on both sides with Philox stream 1.  Compares knowledge after every
playout step, then the whole tree: options, R, s, C, V, P and every node's packed game.
"""
import sys
import numpy as np
from oracle import citadels_oracle as O
from oracle import mccfr_oracle as M
from oracle.philox import PhiloxChance
from tests.golden import ref_harness as H

SEED = 0xC17ADE15


def build_roots(gid, ruleset, back_max=int(__import__("os").environ.get("BACK_MAX", "30"))):
    """Random play to terminal keeping copies, then step back (run_utils.create_a_close_to_finished_game shape)."""
    from copy import deepcopy
    ch_r, ch_o = PhiloxChance(SEED, gid), PhiloxChance(SEED, gid)
    rg = H.new_ref_game(ch_r, ruleset)
    og = O.new_game(ch_o, ruleset)
    rgs, ogs = [deepcopy(rg)], [og.clone()]
    while True:
        H.set_chance(ch_r)
        ropts = rg.get_options_from_state()
        oopts = og.options()
        assert H.ref_descriptors(ropts) == oopts
        assert H.ref_knowledge(rg) == og.knowledge(), ("knowledge", gid, len(rgs))
        i = ch_r.randbelow(len(ropts))
        assert i == ch_o.randbelow(len(oopts))
        w = ropts[i].carry_out(rg)
        og.apply(oopts[i])
        rgs.append(deepcopy(rg))
        ogs.append(og.clone())
        if w:
            break
    back = 1 + ch_r.randbelow(back_max)
    ch_o.randbelow(back_max)
    limit = 0
    while True:
        k = max(0, min(len(rgs) - 1, len(rgs) - back))
        H.set_chance(ch_r)
        n = len(rgs[k].get_options_from_state())
        back -= 1
        limit += 1
        if n >= 2 or limit >= 100:
            break
    return rgs[k], ogs[k], k


def compare_trees(rn, on, path="root"):
    ok = True
    rc = [H.ref_descriptors([c[0]])[0] if c[0].name not in ("discard_and_draw", "cardinal_exchange") else None for c in rn.children]
    oc = [c[0] for c in on.children]
    if len(rc) != len(oc) or any(a is not None and a != b for a, b in zip(rc, oc)):
        print(path, "children differ", len(rc), len(oc))
        return False
    for name, a, b in (("R", rn.cumulative_regrets, on.R), ("s", rn.strategy, on.s), ("C", rn.cumulative_strategy, on.C),
                       ("V", rn.node_value, on.V), ("P", rn.winning_probabilities, on.P)):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        if a.shape != b.shape or not np.allclose(a, b, rtol=1e-9, atol=1e-12, equal_nan=True):
            print(path, name, "differs", a, b)
            ok = False
    if H.ref_pack(rn.game, on.game.ruleset)[:228] != on.game.pack()[:228]:
        print(path, "game record differs")
        ok = False
    if H.ref_knowledge(rn.game) != on.game.knowledge():
        print(path, "knowledge differs")
        ok = False
    for i, (cr, co) in enumerate(zip(rn.children, on.children)):
        ok = compare_trees(cr[1], co[1], path + "/%d" % i) and ok
    return ok


def run(gid, iters, ruleset):
    H.patch_cfr_chance()
    from algorithms.deep_mccfr import CFRNode
    rg, og, k = build_roots(gid, ruleset)
    ch_r, ch_o = PhiloxChance(SEED, gid, stream=1), PhiloxChance(SEED, gid, stream=1)
    H.set_chance(ch_r)
    og.chance = ch_o
    if rg.terminal:
        return True, 0, k
    rn = CFRNode(rg, original_player_id=rg.gamestate.player_id)
    rn.cfr_train(max_iterations=iters)
    on = M.Node(og, og.player)
    on.cfr_train(iters)
    ok = compare_trees(rn, on)
    assert ch_r.i == ch_o.i, ("draw counts", ch_r.i, ch_o.i)
    return ok, sum(1 for _ in on.walk()), k


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ruleset = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    bad = 0
    for gid in range(first, first + n):
        ok, nodes, k = run(gid, iters, ruleset)
        print("gid", gid, "root step", k, "nodes", nodes, "OK" if ok else "MISMATCH", flush=True)
        bad += not ok
    print("roots", n, "bad", bad)

"""Lock-step differential check: real reference vs oracle under the same Philox chance stream.

Container-only (needs /root/reference).  Usage: python -m tests.golden.check_oracle_vs_ref [n_games] [ruleset] [first_gid]
"""
import sys
from oracle import citadels_oracle as O
from oracle.philox import PhiloxChance
from tests.golden import ref_harness as H

SEED = 0xC17ADE15


def run_pair(gid, ruleset, verbose=True):
    ch_r = PhiloxChance(SEED, gid)
    ch_o = PhiloxChance(SEED, gid)
    rg = H.new_ref_game(ch_r, ruleset)
    og = O.new_game(ch_o, ruleset)
    steps = 0
    while True:
        H.set_chance(ch_r)
        ropts = rg.get_options_from_state()
        rd = H.ref_descriptors(ropts)
        od = og.options()
        rp, op = H.ref_pack(rg, ruleset), og.pack()
        if ruleset == 2 and H.ref_knowledge(rg) != og.knowledge():
            print("KNOWLEDGE MISMATCH game", gid, "step", steps)
            return False, steps
        if rp != op or rd != od:
            if verbose:
                print("MISMATCH game", gid, "step", steps, "state", rg.gamestate.state, "player", rg.gamestate.player_id)
                if rp != op:
                    diff = [i for i in range(256) if rp[i] != op[i]]
                    print(" state bytes differ at", diff[:40])
                    print(" ref", [rp[i] for i in diff[:40]])
                    print(" ora", [op[i] for i in diff[:40]])
                if rd != od:
                    print(" ref opts", [(O.KIND_NAMES[O.d_kind(d)], hex(d)) for d in rd][:70])
                    print(" ora opts", [(O.KIND_NAMES[O.d_kind(d)], hex(d)) for d in od][:70])
            return False, steps
        i = ch_r.randbelow(len(ropts))
        i2 = ch_o.randbelow(len(od))
        assert i == i2
        steps += 1
        w = ropts[i].carry_out(rg)
        w2 = og.apply(od[i])
        if bool(w) != bool(w2):
            print("winner flag mismatch", gid, steps)
            return False, steps
        if w:
            rp, op = H.ref_pack(rg, ruleset), og.pack()
            if rp != op:
                diff = [i for i in range(256) if rp[i] != op[i]]
                print("final state mismatch", gid, diff[:40], [rp[i] for i in diff[:40]], [op[i] for i in diff[:40]])
                return False, steps
            return True, steps


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    ruleset = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    bad = 0
    tot = 0
    for gid in range(first, first + n):
        ok, steps = run_pair(gid, ruleset)
        tot += steps
        if not ok:
            bad += 1
            if bad >= 3:
                break
    print("games", n, "bad", bad, "steps", tot)

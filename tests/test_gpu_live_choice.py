"""run_mccfr's decision on the device against the real reference (tests/golden/live_choice_*.npz, see
tests/test_live_choice_cpu.py): ctd_mccfr_result.live_option, the facade's run_mccfr and the arena's batched search."""
import numpy as np
import pytest

from tests.test_live_choice_cpu import PURE, load, same_option

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from citadels_self_play_b200 import Engine
    e = Engine(capacity=256)
    yield e
    e.close()


@pytest.mark.parametrize("name", PURE)
def test_live_option_matches_reference(engine, name):
    z = load(name)
    n = len(z["gids"])
    engine.load_roots(z["roots"], z["knows"], z["used"], z["gids"])
    res = engine.mccfr(n, iterations=int(z["iterations"]), seed=int(z["seed"]), ruleset=int(z["ruleset"]))["results"]
    for r in range(n):
        if z["terminal"][r]:
            assert res[r]["status"] == 1 and res[r]["live_option"] == 0
            continue
        assert res[r]["status"] == 0 and bool(res[r]["role_pick"]) == bool(z["role_pick"][r])
        assert same_option(res[r]["live_option"], z["live"][r]), (name, r)
    assert int(z["role_pick"].sum()) == int(res["role_pick"][~z["terminal"]].sum())


@pytest.mark.parametrize("backend", ["fused", "tcgen05", "fp32"])
def test_deep_live_option_matches_reference(engine, backend):
    """run_mccfr(game, model, 200): cfr_pred at depth 10 with ValueOnlyNN(418,512) under torch.manual_seed(0), then the decision."""
    import torch
    from citadels_self_play_b200.value_model import ValueOnlyNN
    z = load("live_choice_deep_preset.npz")
    n = len(z["gids"])
    torch.manual_seed(0)
    engine.set_value_model(ValueOnlyNN(418, 512).eval())
    engine.set_value_backend(backend)
    engine.load_roots(z["roots"], z["knows"], z["used"], z["gids"])
    res = engine.mccfr_pred(n, iterations=int(z["iterations"]), max_depth=int(z["max_depth"]), seed=int(z["seed"]))["results"]
    engine.set_value_backend("fused")      # back to the default
    live = [r for r in range(n) if not z["terminal"][r]]
    # leaf values differ from torch's CPU GEMV in the last fp32 bits; a decision flips only if the uniform draw lands within
    # that distance of a CDF step, so every one of the 24 recorded decisions is expected to match
    assert all(same_option(res[r]["live_option"], z["live"][r]) for r in live)
    assert all(int(res[r]["n_children"]) == int(z["nchild"][r]) for r in live)


def test_facade_run_mccfr_and_arena_return_the_reference_decision(engine):
    """facade.run_mccfr (run_utils.py:74-87) and arena.search_batch + _live_choice on recorded roots: the option objects a
    caller gets are the reference's decisions, role-pick roots included."""
    from citadels_self_play_b200 import facade as F, arena
    from citadels_self_play_b200.layout import STATE_DTYPE, KNOW_BYTES
    z = load("live_choice_preset_early.npz")
    seed, iters = int(z["seed"]), int(z["iterations"])
    picks = [r for r in range(len(z["gids"])) if z["role_pick"][r]][:6] + [r for r in range(len(z["gids"]))
                                                                            if not z["role_pick"][r] and not z["terminal"][r]][:6]
    games = []
    for r in picks:
        g = F.Game.__new__(F.Game)
        g._engine, g.seed, g.gid, g._fresh, g._searches = engine, seed, int(z["gids"][r]), False, 0
        g._rec = np.array(np.frombuffer(z["roots"][r].tobytes(), dtype=STATE_DTYPE)[0])
        viewer = int(g._rec["player"])
        g._know = np.zeros(6 * KNOW_BYTES, np.uint8)
        g._know[viewer * KNOW_BYTES:(viewer + 1) * KNOW_BYTES] = z["knows"][r]
        g._used = z["used"][r].copy()
        g.players = [F.Agent(g, i) for i in range(6)]
        games.append(g)
    res = arena.search_batch(engine, games, [0] * len(games), None, iterations=iters)
    for g, r, rec in zip(games, picks, res):
        ch = arena._live_choice(rec, g, g.get_options_from_state())
        assert same_option(ch.desc, z["live"][r]) and ch.desc in [o.desc for o in g.get_options_from_state()]
    for g, r in zip(games, picks):
        chosen, root = F.run_mccfr(g, max_iterations=iters)
        assert same_option(chosen.desc, z["live"][r]), r
        assert root.role_pick_node == bool(z["role_pick"][r])
        if root.role_pick_node:
            assert chosen.name == "role_pick" and chosen.attributes["choice"] in g.roles_to_choose_from.values()
        # a second search of the same game draws from another stream (the game's search counter is part of the tree id)
        assert g._searches == 1
    with pytest.raises(ValueError):
        t = int(np.flatnonzero(load("live_choice_preset.npz")["terminal"])[0])
        zz = load("live_choice_preset.npz")
        g = games[0]
        g._rec = np.array(np.frombuffer(zz["roots"][t].tobytes(), dtype=STATE_DTYPE)[0])
        F.run_mccfr(g, max_iterations=10)

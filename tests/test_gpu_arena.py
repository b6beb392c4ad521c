"""The reference's evaluation harness (compare_to_random.py) on the engine: seat 0 deep MCCFR, seat 1 pure MCCFR,
the rest random -- serial through the facade and batched in lock-step (citadels_self_play_b200.arena)."""
import random
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model():
    import torch
    from citadels_self_play_b200.value_model import ValueOnlyNN
    torch.manual_seed(0)
    return ValueOnlyNN(418, 512).eval()


def test_batched_arena_plays_games_to_the_end():
    from citadels_self_play_b200 import arena
    stats = {}
    winners = arena.play_games_batched(6, _model(), seed=4242, first_gid=900, deep_iterations=30, pure_iterations=60, stats=stats)
    assert sum(winners) == 6 and all(w >= 0 for w in winners)
    assert stats["deep_searches"] > 50 and stats["pure_searches"] > 50
    # identical inputs, identical outcome (searches and games are pure functions of seed / gid / decision number)
    again = arena.play_games_batched(6, _model(), seed=4242, first_gid=900, deep_iterations=30, pure_iterations=60)
    assert again == winners


def test_search_batch_decisions_are_legal_and_match_single_tree_searches():
    from citadels_self_play_b200 import arena, facade as F, Engine
    rng = random.Random(5)
    games = []
    for gid in range(40, 48):   # mid-game positions with a real choice
        g = F.create_game(seed=77, gid=gid)
        for _ in range(60 + 7 * (gid % 5)):
            rng.choice(g.get_options_from_state()).carry_out(g)
        while len(g.get_options_from_state()) < 2:
            g.get_options_from_state()[0].carry_out(g)
        games.append(g)
    eng = Engine(capacity=16)
    res = arena.search_batch(eng, games, [0] * len(games), None, iterations=120)
    for g, r in zip(games, res):
        opts = g.get_options_from_state()
        ch = arena._live_choice(r, g, opts)
        assert ch.desc in [o.desc for o in opts]
        # the same root searched alone through the facade's CFRNode: same tree (stream keyed by the game id, decision 0)
        node = F.CFRNode(g, original_player_id=g.gamestate.player_id)
        node.cfr_train(max_iterations=120)
        k = int(r["n_children"])
        assert k == len(node.children)
        if not r["role_pick"]:
            assert np.allclose(r["cumulative_strategy"][:k], node.cumulative_strategy, rtol=1e-12, atol=0)
            assert [int(x) for x in r["options"][:k]] == [o.desc for o, _ in node.children]
    eng.close()


def test_serial_arena_is_the_reference_loop():
    from citadels_self_play_b200 import arena
    winners = arena.play_games(1, _model(), seed=99, first_gid=5, deep_iterations=20, pure_iterations=40)
    assert sum(winners) == 1

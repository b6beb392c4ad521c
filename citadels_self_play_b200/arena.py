"""The reference's evaluation harness (compare_to_random.py:8-37) on the engine.

Seat 0 decides with deep MCCFR (value model at depth 10, 200 iterations), seat 1 with pure MCCFR (2000 iterations),
seats 2-5 uniformly at random; forced moves (one legal option) are played without a search, as in the reference.

  play_games(sim_number, model)           the reference's loop, one game at a time through the facade (same calls, same
                                          order: get_options_from_state -> run_mccfr -> option.carry_out)
  play_games_batched(n_games, model)      the same experiment with every game advanced in lock-step: all games in which
                                          seat 0 (resp. seat 1) has a real choice are searched in ONE ctd_mccfr_pred
                                          (resp. ctd_mccfr) launch, one tree per game, and the live decision
                                          (CFRNode.action_choice(live=True), algorithms/deep_mccfr.py:67-75, with the
                                          role-preference quirk of game/game.py:312-317) is taken from the result records.

Every search of game g at decision number t runs on the Philox stream keyed (seed, g | t << 40), so successive decisions of
one game do not reuse draws.  There is no CPU path: games live in 256-byte records, every transition and every search
runs on the device.
"""
import random

import numpy as np

from .engine import Engine, EngineError, DEFAULT_SEED, RULESET_PRESET
from .layout import KNOW_BYTES
from . import facade as F


def play_games(sim_number, model=None, engine=None, seed=DEFAULT_SEED, first_gid=0, deep_iterations=200,
               pure_iterations=2000, rng=None):
    """compare_to_random.play_games (compare_to_random.py:8-37).  -> winners[6]."""
    rng = rng or random.Random(seed)
    winners = [0] * 6
    for s in range(sim_number):
        game = F.create_game(engine=engine, seed=seed, gid=first_gid + s)
        winner = False
        while not winner:
            pid = game.gamestate.player_id
            if pid == 0 and model is not None and len(game.get_options_from_state()) > 1:
                chosen, _ = F.run_mccfr(game, model, max_iterations=deep_iterations)
            elif pid == 1 and len(game.get_options_from_state()) > 1:
                chosen, _ = F.run_mccfr(game, max_iterations=pure_iterations)
            else:
                chosen = rng.choice(game.get_options_from_state())
            winner = chosen.carry_out(game)
        winners[winner.id] += 1
    return winners


def _live_choice(res, game, options, nprng=None):
    """CFRNode.action_choice(live=True) from one ctd_mccfr_result record: the kernel drew the decision from the tree's own
    chance stream right after the search (cumulative-strategy draw, or the role-preference quirk of game/game.py:312-317 at a
    role-pick root)."""
    d = int(res["live_option"])
    if d == 0:
        raise ValueError("a terminal root has no children (the reference raises ValueError here too)")
    return F.option(d, game)


def search_batch(engine, games, decision_no, model=None, iterations=2000, max_depth=10, weight=5.0):
    """One MCCFR tree per game, all in one launch.  -> ctd_mccfr_result records (numpy structured array)."""
    n = len(games)
    recs = np.stack([np.frombuffer(g._rec.tobytes(), dtype=np.uint8) for g in games])
    knows = np.stack([g._know[g.gamestate.player_id * KNOW_BYTES:(g.gamestate.player_id + 1) * KNOW_BYTES] for g in games])
    used = np.stack([g._used for g in games])
    gids = np.array([F.tree_gid(g.gid, t) for g, t in zip(games, decision_no)], dtype=np.uint64)
    engine.load_roots(recs, knows, used, gids)
    ruleset = int(games[0]._rec["ruleset"])
    seed = games[0].seed
    if model is None:
        out = engine.mccfr(n, iterations=iterations, seed=seed, ruleset=ruleset)
    else:
        engine.set_value_model(model)
        out = engine.mccfr_pred(n, iterations=iterations, max_depth=max_depth, seed=seed, ruleset=ruleset, weight=weight)
    res = out["results"]
    bad = res["status"] & ~np.uint32(1)
    if bad.any():
        raise EngineError("MCCFR status %d in a batched search (2 device memory exhausted, 4 container capacity, 16 the reference "
                          "raises for this root)" % int(bad.max()))
    return res


def play_games_batched(n_games, model=None, engine=None, seed=DEFAULT_SEED, first_gid=0, deep_iterations=200,
                       pure_iterations=2000, max_depth=10, rng=None, ruleset=RULESET_PRESET, stats=None):
    """The arena experiment over `n_games` games advanced in lock-step.  -> winners[6].
    `stats` (optional dict) receives the number of searches and launches."""
    rng = rng or random.Random(seed)
    nprng = np.random.default_rng(seed & 0xFFFFFFFF)
    step_engine = engine or F.default_engine()
    search_engine = Engine(capacity=max(1, n_games), device=step_engine.device)
    games = [F.create_game(engine=step_engine, seed=seed, gid=first_gid + i, ruleset=ruleset) for i in range(n_games)]
    decisions = [0] * n_games
    winners = [0] * 6
    live = list(range(n_games))
    n_search = [0, 0]
    try:
        while live:
            options = {i: games[i].get_options_from_state() for i in live}
            chosen = {}
            for seat, its, mdl in ((0, deep_iterations, model), (1, pure_iterations, None)):
                if seat == 0 and model is None:
                    continue
                idx = [i for i in live if games[i].gamestate.player_id == seat and len(options[i]) > 1]
                if not idx:
                    continue
                res = search_batch(search_engine, [games[i] for i in idx], [decisions[i] for i in idx], mdl, its, max_depth)
                n_search[seat] += len(idx)
                for i, r in zip(idx, res):
                    chosen[i] = _live_choice(r, games[i], options[i], nprng)
                    decisions[i] += 1
            nxt = []
            for i in live:
                opt = chosen[i] if i in chosen else rng.choice(options[i])
                winner = opt.carry_out(games[i])
                if winner:
                    winners[winner.id] += 1
                else:
                    nxt.append(i)
            live = nxt
    finally:
        if stats is not None:
            stats.update(deep_searches=n_search[0], pure_searches=n_search[1], search_launches=search_engine.launches)
        search_engine.close()
    return winners

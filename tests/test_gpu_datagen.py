"""The reference's data-generation drivers (train_from_scratch.get_mccfr_targets, generate_test_data.setup_game) as batched
engine calls (citadels_self_play_b200.datagen): tuple format and content against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_get_mccfr_targets_tuples():
    from citadels_self_play_b200 import datagen
    stats = {}
    tg = datagen.get_mccfr_targets(None, minimum_sufficient_nodes=60, base_usefullness_treshold=15, pretrain=True,
                                   max_iterations=200, roots_per_batch=64, seed=321, first_gid=7000, stats=stats)
    assert len(tg) >= 60 and stats["batches"] >= 1 and stats["targets"] == len(tg)
    for x, o, v, d in tg:
        k = o.shape[1]
        assert x.shape == (418,) and str(x.dtype) == "torch.float32"
        assert o.shape == (1, k, 131) and str(o.dtype) == "torch.float32" and k >= 1
        assert v.shape == (6,) and str(v.dtype) == "torch.float64" and float(v.sum()) >= 15.0    # the usefulness threshold
        assert d.shape == (k,) and str(d.dtype) == "torch.float64"
        assert float(o[0].sum(1).min()) >= 2.0        # every option row has at least its kind and perpetrator bits


def test_generate_test_data_matches_oracle_roots():
    from citadels_self_play_b200 import datagen
    from oracle import mccfr_oracle as M
    from oracle.philox import PhiloxChance
    seed, gid0, n = 555, 8100, 24
    rows = datagen.generate_test_data(n, max_iterations=120, seed=seed, first_gid=gid0, rng=np.random.default_rng(0))
    assert 1 <= len(rows) <= n
    # rebuild the kept roots with the oracle: same order, same values
    j = 0
    for i in range(n):
        g, _ = M.make_root(seed, gid0 + i, 0, 1, 30)
        if g.terminal:
            continue
        feat = np.asarray(g.encode_game(), dtype=np.float32)   # before run_mccfr's skip_false_choice touches the game
        g.chance = PhiloxChance(seed, gid0 + i, stream=1)
        node = M.Node(g, g.player)
        node.cfr_train(120)
        if len(node.children) == 0:
            continue
        x, o, v, d = rows[j]
        j += 1
        assert np.array_equal(x.numpy(), feat)
        assert o.shape[1] == len(node.children)
        assert np.allclose(v.numpy(), node.V, rtol=1e-9, atol=1e-12)
        if not node.role_pick:
            R = np.asarray(node.R, dtype=float)
            assert np.allclose(d.numpy(), R if R.sum() != 0 else np.ones_like(R), rtol=1e-9, atol=1e-12)
        else:
            assert d.shape == (10,)
        if j >= 6:
            break
    assert j >= 3


def test_root_parallel_mode_single_rank_is_the_plain_search_and_rounds_continue_trees():
    """The labelled root-parallel mode (parallel.root_parallel_mccfr): with one rank and one round it IS Engine.mccfr; in rounds,
    ctd_mccfr_continue resumes the trees where their walks stood (iterations add up, trees keep growing, nothing is refused)."""
    from citadels_self_play_b200 import Engine, parallel
    eng = Engine(capacity=64)
    try:
        eng.make_roots(64, seed=77, first_gid=3000, back_lo=0, back_hi=40)
        plain = eng.mccfr(64, iterations=120, seed=77)["results"]
        one = parallel.root_parallel_mccfr(eng, 64, iterations=120, sync_every=120, seed=77)
        for f in ("status", "n_nodes", "rng_draws", "cumulative_regrets", "cumulative_strategy", "node_value", "live_option"):
            assert np.array_equal(plain[f], one[f]), f
        rounds = parallel.root_parallel_mccfr(eng, 64, iterations=120, sync_every=40, seed=77)
        live = plain["status"] == 0
        assert (rounds["status"] == plain["status"]).all()
        assert (rounds["iterations"][live] == 120).all() and (rounds["n_nodes"][live] >= 2).all()
        # root_set round trip: what is written is what the next result record reports
        R = np.tile(np.arange(128, dtype=np.float64), (64, 1))
        C = np.full((64, 128), 1.0 / 128)
        V = np.tile(np.arange(6, dtype=np.float64) + 1, (64, 1))
        eng.root_set(R, C, V)
        after = eng.mccfr_continue(64, 0, seed=77)["results"]
        vec = live & (after["role_pick"] == 0) & (after["n_children"] > 0)
        i = int(np.flatnonzero(vec)[0])
        k = int(after["n_children"][i])
        assert np.allclose(after["node_value"][i], V[i]) and np.allclose(after["cumulative_regrets"][i, :k], R[i, :k])
    finally:
        eng.close()

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

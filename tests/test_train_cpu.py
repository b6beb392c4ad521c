"""Host side of the value-network training path (no GPU): the chance functions the device kernels and the reference harness share,
the fixture's sanity, and the StepLR arithmetic."""
import os
import numpy as np

from tests.golden_util import GOLDEN


def test_training_chance_functions_are_the_engines_philox():
    from citadels_self_play_b200 import train as T
    from oracle.philox import philox4x32_10
    seed = 0x1234567890ABCDEF
    for step, layer in ((0, 1), (7, 2), (300, 14)):
        e = np.array([0, 1, 2, 3, 4, 5, 1000003, 4294967], dtype=np.uint64)
        got = T.philox_words(seed, step, layer, e)
        for x, w in zip(e, got):
            blk = philox4x32_10(int(x) >> 2, layer | (step << 8), 0xD0, 0, seed & 0xFFFFFFFF, seed >> 32)
            assert int(w) == blk[int(x) & 3]
    m = T.dropout_mask(seed, 3, 1, 64, 512)
    assert m.shape == (64, 512) and 0.78 < m.mean() < 0.82
    p = T.epoch_permutation(seed, 2, 1000)
    assert sorted(p.tolist()) == list(range(1000)) and not np.array_equal(p, T.epoch_permutation(seed, 3, 1000))


def test_training_fixture_is_a_learning_curve_of_the_reference():
    with np.load(os.path.join(GOLDEN, "train_value_net.npz")) as f:
        z = {k: f[k] for k in f.files}
    assert z["train_x"].shape == (5000, 418) and z["val_x"].shape == (1000, 418) and z["train_v"].shape == (5000, 6)
    assert (z["train_v"].sum(1) >= 8).all()                      # the usefulness threshold of the generator
    assert len(z["train_losses"]) == int(z["epochs"]) and z["train_losses"][-1] < z["train_losses"][0]
    assert np.isclose(z["best_eval"], z["eval_losses"].min())
    assert z["w_fc4.weight"].shape == (6, 128) and z["w_fc1.weight_rowsum"].shape == (512,)

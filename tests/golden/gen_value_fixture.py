"""The reference's trained value model on 1000 encoded states (container-only; writes tests/golden/value_best_model.npz).

    python -m tests.golden.gen_value_fixture

`pretrain/best_model.pt` of the reference is loaded by the reference's own run_utils.setup_model_for_eval (run_utils.py:11-18)
into the reference's own ValueOnlyNN (algorithms/models.py:4-23) and evaluated the way CFRNode.model_inference does
(algorithms/deep_mccfr.py:364-374): torch fp32 on the CPU, square_and_normalize (train_utils.py:143-145), times
model_reward_weights = 5.  Inputs: Game.encode_game() of 1000 CFR roots (mid-game and late-game positions, role-pick
states included) produced by the oracle from (seed, gid).  The fixture keeps the inputs (small integers), the raw network
outputs, the leaf values, and the checkpoint's tensors -- the GPU box has no /root/reference to load them from."""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0xC17ADE15


def main():
    import torch
    from tests.golden import ref_harness as H
    H.load_reference()
    import run_utils
    from algorithms.train_utils import square_and_normalize
    from oracle import mccfr_oracle as M
    torch.set_num_threads(1)
    _load = torch.load
    torch.load = lambda f, *a, **kw: _load(f, *a, **{**kw, "map_location": "cpu"})   # the checkpoint was saved from cuda:0; no GPU here
    model = run_utils.setup_model_for_eval(os.path.join(H.REFERENCE, "pretrain", "best_model.pt"))
    feats = []
    gid = 600000
    while len(feats) < 1000:
        for lo, hi in ((0, 30), (30, 400)):
            g, _ = M.make_root(SEED, gid, 0, lo, hi)
            gid += 1
            if not g.terminal:
                feats.append(np.asarray(g.encode_game(5 if g.state == 0 else None), dtype=np.float32))
    x = np.stack(feats[:1000])
    assert np.array_equal(x, np.round(x)) and np.abs(x).max() < 32768
    with torch.no_grad():
        y = model(torch.from_numpy(x))
        leaf = (5 * square_and_normalize(y, dim=1)).numpy()
    sd = {("sd_" + k): v.numpy() for k, v in model.state_dict().items()}
    path = os.path.join(HERE, "value_best_model.npz")
    np.savez_compressed(path, features=x.astype(np.int16), raw=y.numpy(), leaf=leaf, **sd)
    print("value_best_model.npz", x.shape, "bytes", os.path.getsize(path), "bn1 var", float(sd["sd_bn1.running_var"].min()),
          float(sd["sd_bn1.running_var"].max()), "leaf range", float(leaf.min()), float(leaf.max()))


if __name__ == "__main__":
    main()

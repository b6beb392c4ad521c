"""Value-network training on the device (citadels_self_play_b200.train, csrc/ctd_train.cuh) against the REAL reference's
train_node_value_only (algorithms/train.py:13-86): tests/golden/train_value_net.npz holds the reference's loss curve and the
weights of its best_model.pt for 8 epochs over 5000 targets (tests/golden/gen_train_fixture.py; same initial weights, batch order
and dropout masks).  Gates: every epoch's train / eval loss to 1e-4 relative; the weights of the saved checkpoint to 1e-3 of each
tensor's norm (Adam turns gradients into steps of about lr whatever their size, so single entries whose gradient is rounding noise
can land an lr apart -- the tensor-level gate is the meaningful one; the worst entry is reported)."""
import json
import os
import numpy as np
import pytest

from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_training_run_matches_the_reference(tmp_path):
    import torch
    from citadels_self_play_b200 import train as T
    from citadels_self_play_b200.value_model import ValueOnlyNN, load_checkpoint
    with np.load(os.path.join(GOLDEN, "train_value_net.npz")) as f:
        z = {k: f[k] for k in f.files}
    tr = [(torch.from_numpy(x.astype(np.float32)), None, torch.from_numpy(v), None) for x, v in zip(z["train_x"], z["train_v"])]
    va = [(torch.from_numpy(x.astype(np.float32)), None, torch.from_numpy(v), None) for x, v in zip(z["val_x"], z["val_v"])]
    torch.manual_seed(int(z["init_seed"]))
    model = ValueOnlyNN(418, 512)
    hist = {}
    best = T.train_node_value_only(tr, va, epochs=int(z["epochs"]), lr=float(z["lr"]), hidden_size=512, gamma=float(z["gamma"]),
                                   batch_size=int(z["batch"]), parent_folder=str(tmp_path), model=model, seed=int(z["seed"]), history=hist)
    tl, el = np.array(hist["train_losses"]), np.array(hist["eval_losses"])
    rep = dict(train_rel=float(np.abs(tl / z["train_losses"] - 1).max()), eval_rel=float(np.abs(el / z["eval_losses"] - 1).max()),
               train_losses=tl.tolist(), eval_losses=el.tolist())
    sd = torch.load(os.path.join(str(tmp_path), "best_model.pt"))
    worst_entry = 0.0
    for k, v in sd.items():
        v = v.numpy()
        if "w_" + k in z:
            ref = z["w_" + k]
            if ref.dtype.kind == "f":
                rep["w_" + k] = float(np.linalg.norm(v - ref) / max(np.linalg.norm(ref), 1e-12))
                worst_entry = max(worst_entry, float(np.abs(v - ref).max()))
            else:
                assert int(v) == int(ref), k                       # num_batches_tracked
        else:
            probe = v.reshape(-1)[z["w_" + k + "_probe_idx"]]
            rep["w_" + k] = float(max(np.linalg.norm(probe - z["w_" + k + "_probe"]) / np.linalg.norm(z["w_" + k + "_probe"]),
                                      abs(np.sqrt((v.astype(np.float64) ** 2).sum()) / float(z["w_" + k + "_norm"]) - 1),
                                      np.linalg.norm(v.sum(1) - z["w_" + k + "_rowsum"]) / np.linalg.norm(z["w_" + k + "_rowsum"])))
            worst_entry = max(worst_entry, float(np.abs(probe - z["w_" + k + "_probe"]).max()))
    rep["worst_single_entry_abs"] = worst_entry
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "train_parity.json"), "w"), indent=1)
    assert rep["train_rel"] <= 1e-4 and rep["eval_rel"] <= 1e-4, rep
    assert abs(best / float(z["best_eval"]) - 1) <= 1e-4
    assert all(val <= 1e-3 for key, val in rep.items() if key.startswith("w_")), rep
    # the checkpoint is what run_utils.setup_model_for_eval loads (run_utils.py:11-18)
    m = load_checkpoint(os.path.join(str(tmp_path), "best_model.pt"))
    assert not m.training

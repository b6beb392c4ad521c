"""Host side of the value model: a state_dict-compatible mirror of the reference's ValueOnlyNN
(algorithms/models.py:4-23; checkpoint contract in SURVEY.md Appendix D) and the BatchNorm folding /
transposition the engine's kernels expect (include/citadels_b200.h, ctd_set_value_model)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

FEATURES = 418
FEATURES_PAD = 448


class ValueOnlyNN(nn.Module):
    """Same parameter names, shapes and construction order as the reference class, so that
    `load_state_dict(torch.load("pretrain/best_model.pt"))` works and `torch.manual_seed(s)` gives the
    same random initialisation."""

    def __init__(self, input_size=FEATURES, hidden_size=512):
        super().__init__()
        self.fc1 = nn.Linear(input_size, hidden_size)
        self.bn1 = nn.BatchNorm1d(hidden_size)
        self.dropout1 = nn.Dropout(0.2)
        self.fc2 = nn.Linear(hidden_size, hidden_size // 2)
        self.bn2 = nn.BatchNorm1d(hidden_size // 2)
        self.dropout2 = nn.Dropout(0.2)
        self.fc3 = nn.Linear(hidden_size // 2, hidden_size // 4)
        self.fc4 = nn.Linear(hidden_size // 4, 6)

    def forward(self, x):
        x = self.dropout1(F.relu(self.bn1(self.fc1(x))))
        x = self.dropout2(F.relu(self.bn2(self.fc2(x))))
        return self.fc4(F.relu(self.fc3(x)))


def load_checkpoint(path, hidden_size=512):
    """run_utils.setup_model_for_eval (run_utils.py:11-18)."""
    m = ValueOnlyNN(FEATURES, hidden_size)
    m.load_state_dict(torch.load(path, map_location="cpu"))
    m.eval()
    return m


def fold(model):
    """eval-mode BatchNorm folded into fc1 / fc2, weights transposed to [in][out], fc1 padded to 448 inputs.
    Returns the 8 float32 arrays of ctd_set_value_model."""
    sd = {k: v.detach().double().cpu().numpy() for k, v in model.state_dict().items()}
    if sd["fc1.weight"].shape != (512, FEATURES):
        raise ValueError("the engine's value kernels are built for ValueOnlyNN(418, 512)")

    def fold_bn(w, b, prefix, eps):
        g = sd[prefix + ".weight"] / np.sqrt(sd[prefix + ".running_var"] + eps)
        return w * g[:, None], (b - sd[prefix + ".running_mean"]) * g + sd[prefix + ".bias"]

    w1, b1 = fold_bn(sd["fc1.weight"], sd["fc1.bias"], "bn1", model.bn1.eps)
    w2, b2 = fold_bn(sd["fc2.weight"], sd["fc2.bias"], "bn2", model.bn2.eps)
    w1t = np.zeros((FEATURES_PAD, 512), dtype=np.float64)
    w1t[:FEATURES] = w1.T
    out = [w1t, b1, w2.T, b2, sd["fc3.weight"].T, sd["fc3.bias"], sd["fc4.weight"].T, sd["fc4.bias"]]
    return [np.ascontiguousarray(a, dtype=np.float32) for a in out]


def reference_value(model, features, weight=5.0):
    """What CFRNode.model_inference computes on the CPU (algorithms/deep_mccfr.py:364-374): fp32 torch forward,
    square_and_normalize, times model_reward_weights.  Used by tests as the floating-point reference."""
    with torch.no_grad():
        y = model(torch.as_tensor(np.asarray(features, dtype=np.float32)[..., :FEATURES]))
        y = y * y
        p = (y / y.sum(dim=-1, keepdim=True)).numpy()          # square_and_normalize, float32
        return np.float32(weight) * p                            # model_reward_weights * probabilities, still float32

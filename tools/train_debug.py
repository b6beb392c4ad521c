"""Developer tool: the training step of the engine against a torch restatement of the reference's loop (mirror model, the same
batch order and dropout masks), one optimiser step at a time -- localises a discrepancy to forward / gradient / Adam.
   python tools/train_debug.py [n_samples] [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
from citadels_self_play_b200 import Engine, train as T
from citadels_self_play_b200.value_model import ValueOnlyNN

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "train_value_net.npz"))
x, v = z["train_x"][:n].astype(np.float32), z["train_v"][:n]
SEED, LR = 99, 0.01
torch.manual_seed(1234)
model = ValueOnlyNN(418, 512)
init = {k: t.clone() for k, t in model.state_dict().items()}
state = {"step": -1}
def dropout(input, p=0.5, training=True, inplace=False):
    if not training:
        return input
    layer = 1 if input.shape[1] == 512 else 2
    if layer == 1:
        state["step"] += 1
    m = torch.from_numpy(T.dropout_mask(SEED, state["step"], layer, input.shape[0], input.shape[1]))
    return input * m.to(input.dtype) * 1.25
torch.nn.functional.dropout = dropout
opt = torch.optim.Adam(model.parameters(), lr=LR)
crit = nn.KLDivLoss(reduction="batchmean")
eng = Engine(capacity=8)
tr = T.Trainer(eng, x, v, x[:256], v[:256], bs)
tr.set_state(init)
def sq(t):
    s = t ** 2
    return s / s.sum(-1, keepdim=True)
for ep in range(steps):
    perm = T.epoch_permutation(SEED, ep, n)
    model.train()
    tot = 0.0
    nb = 0
    for i in range(0, n, bs):
        idx = torch.from_numpy(perm[i:i + bs].astype(np.int64))
        opt.zero_grad()
        out = model(torch.from_numpy(x)[idx])
        loss = crit(torch.log(sq(out) + 1e-10), sq(torch.from_numpy(v)[idx]))
        loss.backward()
        opt.step()
        tot += loss.item(); nb += 1
    model.eval()
    with torch.no_grad():
        ev = crit(torch.log(sq(model(torch.from_numpy(x[:256]))) + 1e-10), sq(torch.from_numpy(v[:256]))).item()
    tl, el = tr.epoch(SEED, LR, perm)
    got = tr.get_state()
    print("epoch %d torch train %.9f eval %.9f | engine train %.9f eval %.9f" % (ep, tot / nb, ev, tl, el))
    for k, t in model.state_dict().items():
        if k in got:
            a, b = t.numpy(), got[k]
            print("   %-18s rel %.3e  max|d| %.3e  (|ref| %.3e)" % (k, np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30), np.abs(a - b).max(), np.abs(a).max()))


# ---- gradient check: engine and torch at the SAME weights (lr ~ 0), one batch per step, a different batch every step
print("gradient check (lr 1e-12): per-tensor ||g - g_ref|| / ||g_ref||")
torch.manual_seed(1234)
model = ValueOnlyNN(418, 512)
state["step"] = -1
opt = torch.optim.Adam(model.parameters(), lr=1e-12)
xs, vs = z["train_x"][:4096].astype(np.float32), z["train_v"][:4096]
tr.close()
for b, (lo, hi) in enumerate(((0, 2048), (2048, 4096), (1000, 1904))):
    tr = T.Trainer(eng, xs[lo:hi], vs[lo:hi], xs[:64], vs[:64], 2048)
    sd0 = {k: t.clone() for k, t in model.state_dict().items()}
    tr.set_state(sd0)
    tr.lib.ctd_train_epoch   # (same kernels)
    # make the engine's dropout step counter equal torch's: set_state reset it to 0, torch is at step b
    model.train()
    opt.zero_grad()
    state["step"] = -1
    out = model(torch.from_numpy(xs[lo:hi]))
    loss = crit(torch.log(sq(out) + 1e-10), sq(torch.from_numpy(vs[lo:hi])))
    loss.backward()
    tl, _ = tr.epoch(SEED, 1e-12, None)
    g = tr.get_grads()
    print(" batch %d rows %d: loss torch %.9f engine %.9f" % (b, hi - lo, loss.item(), tl))
    for k, p_ in model.named_parameters():
        a, c = p_.grad.numpy(), g[k]
        print("   %-12s %.3e   (||g_ref|| %.3e, max|g_ref| %.3e, max|d| %.3e)" % (k, np.linalg.norm(a - c) / max(np.linalg.norm(a), 1e-30), np.linalg.norm(a), np.abs(a).max(), np.abs(a - c).max()))
    tr.close()

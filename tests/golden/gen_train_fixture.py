"""Fixture for the value-network training path (container-only; writes tests/golden/train_value_net.npz).

    python -m tests.golden.gen_train_fixture

1. Training data: 5000 + 1000 (features, node_value) pairs of the kind train_from_scratch.get_mccfr_targets produces: nodes with
   children and >= 8 backprops of 400-iteration pure-MCCFR trees (host build of the kernels' code for the trees, the oracle for
   Game.encode_game of every kept node).
2. The REAL reference's train_node_value_only (algorithms/train.py:13-86) on them, on the CPU, 8 epochs, Adam lr 0.01, StepLR,
   batch 2048, with two call sites routed the way random.shuffle is routed elsewhere in this harness: its DataLoader yields the
   batches in the order citadels_self_play_b200.train.epoch_permutation defines, and torch.nn.functional.dropout keeps the
   elements citadels_self_play_b200.train.dropout_mask defines.  Initial weights: torch.manual_seed(1234) default init.
Recorded: the loss and the gradients of the first optimiser step (norms, row sums, 512 probe entries of the big matrices; small
tensors in full); per-epoch train / eval losses (captured from the arguments of the reference's plot_metrics), the best evaluation loss it
returns, and of its best_model.pt the small tensors in full and, for the three big matrices, norms, row sums and 256 probe entries."""
import ctypes
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SEED = 0xC17ADE15
TRAIN_SEED = 20261018
EPOCHS, LR, GAMMA, BATCH = 8, 0.01, 0.9, 2048


def make_targets(n_want, gid0):
    from oracle import citadels_oracle as O
    from citadels_self_play_b200.layout import TreeView
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libctd_hostsim.so"))
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    lib.hs_make_root.argtypes = [u64, u64, i32, u32, u32, i32, vp, vp, vp, vp]
    lib.hs_mccfr.argtypes = [vp, vp, vp, u64, u64, u32, vp, u64, vp, u64, vp]
    arena, out = np.zeros(256 << 20, np.uint8), np.zeros(64 << 20, np.uint8)
    root, know, used, step = np.zeros(256, np.uint8), np.zeros(592, np.uint8), np.zeros(76, np.uint8), np.zeros(1, np.uint32)
    nb = ctypes.c_uint64()
    feats, vals = [], []
    gid = gid0
    while len(feats) < n_want:
        lib.hs_make_root(SEED, gid, 0, 1, 100, 1, root.ctypes.data, know.ctypes.data, used.ctypes.data, step.ctypes.data)
        st = lib.hs_mccfr(root.ctypes.data, know.ctypes.data, used.ctypes.data, SEED, gid, 400, arena.ctypes.data, arena.nbytes,
                          out.ctypes.data, out.nbytes, ctypes.byref(nb))
        gid += 1
        if st != 0:
            continue
        tv = TreeView(out[:nb.value].copy())
        for i in range(len(tv.nodes)):
            n = tv.nodes[i]
            if n["n_children"] == 0 or n["V"].sum() < 8:
                continue
            g = O.Game.unpack(n["game"].tobytes())
            g.unpack_know(n["know"].tobytes(), used)
            feats.append(np.asarray(g.encode_game(0 if g.state == 0 else None), dtype=np.float32))
            vals.append(np.array(n["V"], dtype=np.float64))
    return np.stack(feats[:n_want]), np.stack(vals[:n_want]), gid


def run_reference(xtr, vtr, xva, vva):
    import torch
    from tests.golden import ref_harness as H
    from citadels_self_play_b200 import train as T
    H.load_reference()
    import algorithms.train as RT
    torch.set_num_threads(4)
    state = {"epoch": 0, "step": -1, "curves": None}

    class Loader:   # stands in for torch.utils.data.DataLoader at the reference's two call sites
        def __init__(self, dataset, batch_size, shuffle):
            self.x, self.y = dataset.tensors
            self.bs, self.shuffle = batch_size, shuffle

        def __len__(self):
            return (len(self.x) + self.bs - 1) // self.bs

        def __iter__(self):
            n = len(self.x)
            if self.shuffle:
                order = torch.from_numpy(T.epoch_permutation(TRAIN_SEED, state["epoch"], n).astype(np.int64))
                state["epoch"] += 1
            else:
                order = torch.arange(n)
            for i in range(0, n, self.bs):
                idx = order[i:i + self.bs]
                yield self.x[idx], self.y[idx]

    def dropout(input, p=0.5, training=True, inplace=False):
        if not training:
            return input
        layer = 1 if input.shape[1] == 512 else 2
        if layer == 1:
            state["step"] += 1
        m = torch.from_numpy(T.dropout_mask(TRAIN_SEED, state["step"], layer, input.shape[0], input.shape[1]))
        return input * m.to(input.dtype) * 1.25

    def capture(train_losses, eval_losses, learning_rates, epochs, folder):
        state["curves"] = (list(train_losses), list(eval_losses), list(learning_rates))

    # the gradients and the loss of the very first optimiser step (equal weights on both sides: the cleanest comparison there is)
    first = {}
    _step = torch.optim.Adam.step

    def step(self, *a, **kw):
        if not first:
            names = ["fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight", "bn2.bias", "fc3.weight",
                     "fc3.bias", "fc4.weight", "fc4.bias"]
            params = [p for g in self.param_groups for p in g["params"]]
            first.update({n: p.grad.detach().clone().numpy() for n, p in zip(names, params)})
        return _step(self, *a, **kw)
    torch.optim.Adam.step = step
    _kl = RT.nn.KLDivLoss.forward

    def kl(self, inp, tgt):
        out = _kl(self, inp, tgt)
        state.setdefault("first_loss", float(out))
        return out
    RT.nn.KLDivLoss.forward = kl
    RT.DataLoader = Loader
    RT.plot_metrics = capture
    torch.nn.functional.dropout = dropout
    train = [(torch.from_numpy(x), None, torch.from_numpy(v), None) for x, v in zip(xtr, vtr)]
    val = [(torch.from_numpy(x), None, torch.from_numpy(v), None) for x, v in zip(xva, vva)]
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(1234)
        best = RT.train_node_value_only(train, val, epochs=EPOCHS, lr=LR, hidden_size=512, gamma=GAMMA, batch_size=BATCH, device="cpu",
                                        parent_folder=tmp)
        sd = torch.load(os.path.join(tmp, "best_model.pt"), map_location="cpu")
    torch.optim.Adam.step = _step
    RT.nn.KLDivLoss.forward = _kl
    return best, state["curves"], {k: v.numpy() for k, v in sd.items()}, first, state["first_loss"]


def main():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    xtr, vtr, gid = make_targets(5000, 700000)
    xva, vva, _ = make_targets(1000, gid)
    assert np.array_equal(xtr, np.round(xtr)) and np.abs(xtr).max() < 32768
    best, (tl, el, lrs), sd, grads, first_loss = run_reference(xtr, vtr, xva, vva)
    rng = np.random.RandomState(7)
    out = dict(train_x=xtr.astype(np.int16), train_v=vtr, val_x=xva.astype(np.int16), val_v=vva, seed=np.uint64(TRAIN_SEED),
               epochs=np.int32(EPOCHS), lr=np.float64(LR), gamma=np.float64(GAMMA), batch=np.int32(BATCH), init_seed=np.int32(1234),
               best_eval=np.float64(best), train_losses=np.asarray(tl), eval_losses=np.asarray(el), lrs=np.asarray(lrs))
    out["first_loss"] = np.float64(first_loss)
    for k, v in grads.items():          # gradients of the first step: norms, probes, small tensors in full
        out["g_" + k + "_norm"] = np.float64(np.sqrt((v.astype(np.float64) ** 2).sum()))
        if v.size > 1024:
            probe = rng.randint(0, v.size, size=512)
            out["g_" + k + "_probe_idx"] = probe
            out["g_" + k + "_probe"] = v.reshape(-1)[probe]
            out["g_" + k + "_rowsum"] = v.sum(1)
        else:
            out["g_" + k] = v
    for k, v in sd.items():
        if v.ndim == 2 and v.size > 1024:
            probe = rng.randint(0, v.size, size=256)
            out["w_" + k + "_probe_idx"] = probe
            out["w_" + k + "_probe"] = v.reshape(-1)[probe]
            out["w_" + k + "_rowsum"] = v.sum(1)
            out["w_" + k + "_norm"] = np.float64(np.sqrt((v.astype(np.float64) ** 2).sum()))
        else:
            out["w_" + k] = v
    path = os.path.join(HERE, "train_value_net.npz")
    np.savez_compressed(path, **out)
    print("train_value_net.npz bytes", os.path.getsize(path), "best eval", best, "\ntrain", tl, "\neval", el, "\nlr", lrs)


if __name__ == "__main__":
    main()

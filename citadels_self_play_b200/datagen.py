"""The reference's two data-generation drivers as batched engine calls.

  get_mccfr_targets(...)    train_from_scratch.get_mccfr_targets / simulate_game (train_from_scratch.py:23-64):
                            roots = create_a_random_game(100) (random play, stepped back 1..100 decisions),
                            run_mccfr(max_iterations, training=True), position_root.get_all_targets(threshold);
                            batches of roots are searched in one launch until enough targets exist.
                            (With training=True the reference never evaluates the model inside the search --
                            deep_mccfr.py:119-126 guards every inference with `not self.training` -- so the pre-training and
                            the later rounds run the same pure MCCFR; `model` is accepted for signature parity only.)
  generate_test_data(...)   generate_test_data.setup_game (generate_test_data.py:9-29): roots stepped back 1..30, one tuple per
                            root (game_input, options_input, node_value, target_decision_dist) with
                            run_utils.create_target_strategy (run_utils.py:99-108).

Both return the reference's pickle tuples: (float32[418], float32[1,K,131], float64[6], float64[K]).
Note (SURVEY.md discrepancy 3): the reference passes usefulness_treshold down but build_train_targets ignores it and uses 15;
here the threshold given is the threshold applied (pass 15 for the reference's effective behaviour).
"""
import numpy as np

from .engine import Engine, DEFAULT_SEED, RULESET_PRESET, ROOTS_RANDOM_GAME
from .layout import encode_options


def get_mccfr_targets(model=None, minimum_sufficient_nodes=5000, base_usefullness_treshold=200, pretrain=False,
                      max_iterations=2000, engine=None, roots_per_batch=1024, seed=DEFAULT_SEED, first_gid=0,
                      ruleset=RULESET_PRESET, back=(1, 100), stats=None, group=None):
    """train_from_scratch.get_mccfr_targets (train_from_scratch.py:53-64) with simulate_game (:23-36) batched: every batch is
    `roots_per_batch` roots of create_a_random_game(back[1]) searched in one launch.  Under torch.distributed (one process per
    GPU) every rank searches its own `roots_per_batch` roots of each batch -- global root ids are dealt round-robin by batch and
    rank, nothing crosses GPUs during the search -- and the targets of all ranks are gathered after every batch
    (parallel.gather_targets), so every rank returns the same list, as the reference's Pool returns it to the parent.  Trees that did not complete (status other
    than 0: terminal roots -- run_mccfr raises ValueError on them in the reference --, roots the reference itself raises on,
    engine limits) contribute no targets (ctd_mccfr_targets skips them) and are counted in `stats`."""
    own = engine is None
    eng = engine or Engine(capacity=roots_per_batch)
    from . import parallel, sharding
    rank, world = parallel._world(group)
    targets, batches = [], 0
    n_terminal = n_refused = 0
    try:
        while len(targets) < minimum_sufficient_nodes:
            gid = int(first_gid) + sharding.first_gid(batches, rank, world, roots_per_batch)
            eng.make_roots(roots_per_batch, seed=seed, first_gid=gid, ruleset=ruleset, back_lo=back[0], back_hi=back[1],
                           flavour=ROOTS_RANDOM_GAME)
            st = eng.mccfr(roots_per_batch, iterations=max_iterations, seed=seed, ruleset=ruleset)["results"]["status"]
            n_terminal += int((st == 1).sum())
            n_refused += int((st > 1).sum())
            t = eng.mccfr_targets(roots_per_batch, iterations=max_iterations, seed=seed, ruleset=ruleset,
                                  threshold=float(base_usefullness_treshold))
            assert not len(t["meta"]) or (st[t["meta"]["tree"]] == 0).all()
            targets += Engine.targets_as_tuples(parallel.gather_targets(t, group))
            batches += 1
            if batches > 10000:
                raise RuntimeError("get_mccfr_targets: no targets are being produced (threshold too high for max_iterations?)")
    finally:
        if stats is not None:
            stats.update(batches=batches, roots=batches * roots_per_batch * world, targets=len(targets), ranks=world,
                         terminal_roots_this_rank=n_terminal, refused_roots_this_rank=n_refused)
        if own:
            eng.close()
    return targets


def generate_test_data(n_roots, max_iterations=200, engine=None, seed=DEFAULT_SEED, first_gid=0, ruleset=RULESET_PRESET,
                       back=(1, 30), rng=None):
    import torch
    rng = rng or np.random.default_rng(seed & 0xFFFFFFFF)
    own = engine is None
    eng = engine or Engine(capacity=n_roots)
    out = []
    try:
        eng.make_roots(n_roots, seed=seed, first_gid=first_gid, ruleset=ruleset, back_lo=back[0], back_hi=back[1])
        feats = eng.encode(n_roots, cfr_role_pick=False)  # game.encode_game() of the position handed to run_mccfr
        res = eng.mccfr(n_roots, iterations=max_iterations, seed=seed, ruleset=ruleset)["results"]
        for i, r in enumerate(res):
            k = int(r["n_children"])
            if r["status"] != 0 or k == 0:
                continue                                  # terminal root (ValueError in the reference) / meaningless state
            opts = np.array(r["options"][:k])
            if r["role_pick"]:
                dist = np.array(r["cumulative_regrets"][:60]).reshape(6, 10)[int(rng.integers(0, 6))]
            elif k > len(r["options"]):                   # more children than a result record holds: read them from the tree
                opts, dist, _, _ = eng.root_children(i, 0, k)
            else:
                dist = np.array(r["cumulative_regrets"][:k])
            if dist.sum() == 0:
                dist = np.ones_like(dist)
            out.append((torch.from_numpy(np.array(feats[i][:418], dtype=np.float32)),
                        torch.from_numpy(encode_options(opts)).unsqueeze(0),
                        torch.from_numpy(np.array(r["node_value"])), torch.from_numpy(dist)))
    finally:
        if own:
            eng.close()
    return out

"""Multi-GPU plumbing of the MCCFR path (one process per GPU, torch.distributed; NCCL on the box, gloo in the CPU tests).

  gather_targets(t)            data generation across roots: every rank searched its own shard of roots; the
                               training targets of all ranks are brought together with one all_gather of the counts and one of
                               each (padded) array -- what the reference's Pool.starmap returns to the parent
                               (train_from_scratch.py:39-42, generate_test_data.py:30-33)
  root_parallel_mccfr(...)     LABELLED MODE, not the reference's algorithm (BASELINE.json configs[4] and
                               north_star name it): the ranks search the SAME roots with different chance streams and pool the
                               roots' regrets / cumulative strategy / values with an all-reduce every `sync_every` iterations.
                               The reference's trees are private (algorithms/deep_mccfr.py:27-29), so results differ from it by
                               construction; they are excluded from the parity gates.  With one rank and sync_every = iterations
                               it is exactly Engine.mccfr.
"""
import numpy as np
import torch
import torch.distributed as dist

from .engine import DEFAULT_SEED, RULESET_PRESET


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _all_gather_rows(a, counts, group=None):
    """a: numpy array whose first axis has counts[rank] rows -> concatenation over ranks (rank order)."""
    a = np.ascontiguousarray(a)
    rank, world = _world(group)
    mx = int(max(counts))
    if mx == 0:
        return a[:0]
    pad = np.zeros((mx,) + a.shape[1:], dtype=a.dtype)
    pad[:len(a)] = a
    t = torch.from_numpy(pad.view(np.uint8).reshape(mx, -1)).to(_device())
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    parts = [o.cpu().numpy().reshape(-1).view(a.dtype).reshape((mx,) + a.shape[1:])[:int(c)] for o, c in zip(out, counts)]
    return np.concatenate(parts)


def gather_targets(t, group=None):
    """t: the dict Engine.mccfr_targets returns on this rank -> the same dict holding every rank's targets (rank order; `tree`
    in meta becomes rank-local tree + rank * 2^24, option_offset is rebased).  Collective: every rank must call it."""
    rank, world = _world(group)
    if world == 1:
        return t
    n_rec, n_opt = len(t["meta"]), len(t["options"])
    cnt = torch.tensor([n_rec, n_opt], dtype=torch.int64, device=_device())
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt, group=group)
    recs = [int(c[0]) for c in allc]
    opts = [int(c[1]) for c in allc]
    meta = t["meta"].copy()
    meta["tree"] = meta["tree"] + np.uint32(rank << 24)
    meta["option_offset"] = meta["option_offset"] + np.uint32(sum(opts[:rank]))
    return dict(features=_all_gather_rows(np.ascontiguousarray(t["features"]), recs, group),
                meta=_all_gather_rows(meta, recs, group),
                options=_all_gather_rows(t["options"], opts, group),
                regrets=_all_gather_rows(t["regrets"], opts, group))


def root_parallel_mccfr(engine, n_roots, iterations=2000, sync_every=200, seed=DEFAULT_SEED, ruleset=RULESET_PRESET, group=None):
    """Root-parallel MCCFR over the roots loaded / made in `engine` (the same roots on every rank).  Every rank grows its own tree
    per root on its own chance stream (Philox key seed + rank * odd constant) for `sync_every` iterations at a time; then the
    roots' cumulative_regrets, cumulative_strategy and node_value are summed over the ranks (all_reduce; the strategy renormalised)
    and written back into every rank's roots.  Only roots whose arrays have the same meaning on every rank are pooled: the
    searching player's own decision nodes (one child per legal option, in option order); sampled roots (opponent / role-pick
    roots) pool node_value only.  -> the result records of this rank after the last round (identical root arrays on all ranks)."""
    rank, world = _world(group)
    key = (int(seed) + rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    res = None
    done = 0
    prev_R = prev_V = None       # the pooled regrets / values after the previous round (what every rank continued from)
    while done < iterations:
        step = min(sync_every, iterations - done)
        out = engine.mccfr(n_roots, iterations=step, seed=key, ruleset=ruleset) if done == 0 else \
            engine.mccfr_continue(n_roots, step, seed=key, ruleset=ruleset)
        res = out["results"]
        done += step
        if world == 1:
            continue
        k = res["n_children"].astype(np.int64)
        poolable = (res["status"] == 0) & (res["role_pick"] == 0) & (res["viewer"] == res["player"]) & (k > 0) & (k <= 128)
        if prev_R is None:
            prev_R, prev_V = np.zeros((n_roots, 128)), np.zeros((n_roots, 6))
        # regrets and values are running sums: pool what THIS round added on every rank; the cumulative strategy is a normalised
        # moving average: pool it as the mean
        dR = np.where(poolable[:, None], res["cumulative_regrets"][:, :128] - prev_R, 0.0)
        C = np.where(poolable[:, None], res["cumulative_strategy"][:, :128], 0.0)
        dV = np.where((res["status"] == 0)[:, None], res["node_value"] - prev_V, 0.0)
        pk = torch.from_numpy(np.concatenate([dR, C, dV, poolable[:, None].astype(np.float64)], axis=1)).to(_device())
        dist.all_reduce(pk, op=dist.ReduceOp.SUM, group=group)
        pk = pk.cpu().numpy()
        dR, C, dV, votes = pk[:, :128], pk[:, 128:256], pk[:, 256:262], pk[:, 262]
        ok = poolable & (votes == world)               # the same decision node on every rank
        cs = C.sum(axis=1, keepdims=True)
        C = np.where(cs > 0, C / np.where(cs > 0, cs, 1.0), C)
        R = np.where(ok[:, None], prev_R + dR, res["cumulative_regrets"][:, :128])
        C = np.where(ok[:, None], C, res["cumulative_strategy"][:, :128])
        V = prev_V + dV
        engine.root_set(R, C, V)
        prev_R, prev_V = np.where(ok[:, None], R, 0.0), V
        res = res.copy()
        res["cumulative_regrets"][:, :128], res["cumulative_strategy"][:, :128], res["node_value"] = R, C, V
    return res

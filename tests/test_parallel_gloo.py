"""N>1 paths of the MCCFR side on CPU: two gloo ranks (world_size 2).
  * parallel.gather_targets: every rank ends up with every rank's training targets, offsets rebased
  * parallel.root_parallel_mccfr (labelled mode): the pooling arithmetic over a scripted engine -- sums over ranks, renormalised
    strategy, only decision-node roots present on every rank are pooled, the rounds call mccfr / mccfr_continue / root_set"""
import os
import socket
import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _targets(rank):
    from citadels_self_play_b200.layout import TARGET_META_DTYPE
    rng = np.random.RandomState(100 + rank)
    ks = [3, 1, 7][:2 + rank] if rank == 0 else [2, 5, 4]
    meta = np.zeros(len(ks), dtype=TARGET_META_DTYPE)
    off = 0
    for i, k in enumerate(ks):
        meta[i]["tree"], meta[i]["node"], meta[i]["n_options"], meta[i]["option_offset"] = i, 10 * rank + i, k, off
        meta[i]["node_value"] = rng.rand(6)
        off += k
    return dict(features=rng.rand(len(ks), 418).astype(np.float32), meta=meta,
                options=rng.randint(1, 1 << 40, size=off).astype(np.uint64), regrets=rng.rand(off))


class _ScriptedEngine:
    """Stands in for Engine: returns fixed result records per rank and records what the mode asks of it."""

    def __init__(self, rank):
        from citadels_self_play_b200.layout import MCCFR_RESULT_DTYPE
        self.rank, self.calls, self.sets = rank, [], []
        r = np.zeros(3, dtype=MCCFR_RESULT_DTYPE)
        r["n_children"] = [4, 3, 10]
        r["role_pick"] = [0, 0, 1]
        r["viewer"] = [2, 1, 0]
        r["player"] = [2, 1 + rank, 0]            # root 1: an opponent node on rank 1 only -> not pooled
        r["cumulative_regrets"][0, :4] = [1.0 + rank, 2, 3, 4]
        r["cumulative_strategy"][0, :4] = [0.1, 0.2, 0.3, 0.4] if rank == 0 else [0.4, 0.3, 0.2, 0.1]
        r["cumulative_regrets"][1, :3] = [5, 6, 7 + rank]
        r["cumulative_strategy"][1, :3] = [0.5, 0.25, 0.25]
        r["node_value"] = np.arange(18).reshape(3, 6) * (1 + rank)
        self.res = r

    def mccfr(self, n, iterations, seed, ruleset):
        self.calls.append(("mccfr", iterations, seed))
        return dict(results=self.res.copy())

    def mccfr_continue(self, n, more, seed, ruleset):
        self.calls.append(("continue", more, seed))
        return dict(results=self.res.copy())

    def root_set(self, R, C, V):
        self.sets.append((R.copy(), C.copy(), V.copy()))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from citadels_self_play_b200 import parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = P.gather_targets(_targets(rank))
    eng = _ScriptedEngine(rank)
    res = P.root_parallel_mccfr(eng, 3, iterations=50, sync_every=20, seed=7)
    dist.barrier()
    q.put((rank, g, eng.calls, eng.sets, res))
    dist.destroy_process_group()


def test_two_rank_gather_and_root_parallel_pooling():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r[0], r[1:]) for r in (q.get(timeout=180) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # ---- gather: both ranks hold the concatenation, every record finds its own options
    local = [_targets(0), _targets(1)]
    for rank in range(world):
        g = got[rank][0]
        assert len(g["meta"]) == sum(len(t["meta"]) for t in local)
        assert np.array_equal(g["features"], np.concatenate([t["features"] for t in local]))
        assert np.array_equal(g["options"], np.concatenate([t["options"] for t in local]))
        i = 0
        for src, t in enumerate(local):
            for m in t["meta"]:
                gm = g["meta"][i]
                o, k = int(gm["option_offset"]), int(gm["n_options"])
                assert k == int(m["n_options"]) and int(gm["tree"]) == int(m["tree"]) + (src << 24)
                assert np.array_equal(g["options"][o:o + k], t["options"][int(m["option_offset"]):int(m["option_offset"]) + k])
                assert np.array_equal(g["regrets"][o:o + k], t["regrets"][int(m["option_offset"]):int(m["option_offset"]) + k])
                assert np.array_equal(gm["node_value"], m["node_value"])
                i += 1
    # ---- root-parallel rounds: 20 + 20 + 10 iterations, a different chance stream per rank, pooled after every round
    for rank in range(world):
        _, calls, sets, res = got[rank]
        assert [c[:2] for c in calls] == [("mccfr", 20), ("continue", 20), ("continue", 10)]
        assert len(sets) == 3
        R, C, V = sets[-1]
        R1, C1, V1 = sets[0]
        assert np.allclose(R1[0, :4], [1 + 2, 4, 6, 8])                      # round 1, root 0: a decision node on both ranks -> summed
        assert np.allclose(C1[0, :4], [0.25, 0.25, 0.25, 0.25])              # mean of the two strategies, renormalised
        assert np.allclose(V1, np.arange(18).reshape(3, 6) * 3)              # values are pooled for every completed root
        # later rounds pool INCREMENTS over what every rank continued from: the scripted engine reports its first-round arrays
        # again, i.e. increments of (local - pooled) per rank
        own0, own1 = _ScriptedEngine(0).res, _ScriptedEngine(1).res
        inc = (own0["cumulative_regrets"][0, :4] - R1[0, :4]) + (own1["cumulative_regrets"][0, :4] - R1[0, :4])
        assert np.allclose(sets[1][0][0, :4], R1[0, :4] + inc)
        own = _ScriptedEngine(rank).res
        assert np.allclose(R[1, :3], own["cumulative_regrets"][1, :3])       # root 1: not the same node everywhere -> kept
        assert np.array_equal(res["cumulative_regrets"][0, :4], R[0, :4])
    assert got[0][1][0][2] != got[1][1][0][2]                                # per-rank Philox keys


def test_single_process_root_parallel_is_the_plain_search():
    from citadels_self_play_b200 import parallel as P
    eng = _ScriptedEngine(0)
    res = P.root_parallel_mccfr(eng, 3, iterations=200, sync_every=200, seed=11)
    assert eng.calls == [("mccfr", 200, 11)] and eng.sets == []
    assert np.array_equal(res["cumulative_regrets"], eng.res["cumulative_regrets"])

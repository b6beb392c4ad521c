"""N>1 path on CPU: two gloo ranks shard game ids as bench.py does, each plays its block with the oracle, the
all-reduced statistics equal a single-process run (results depend on (seed, gid) only, never on the sharding)."""
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


def _stats_for(seed, gid0, n):
    from oracle import citadels_oracle as O
    st = dict(games=0, steps=0, steps_sq=0, errors=0, wins=[0] * 6, points_sum=[0] * 6, points_sq=[0] * 6)
    for i in range(n):
        w, pts, steps, _ = O.playout(seed, gid0 + i)
        st["games"] += 1
        st["steps"] += steps
        st["steps_sq"] += steps * steps
        st["wins"][w] += 1
        for p in range(6):
            st["points_sum"][p] += pts[p]
            st["points_sq"][p] += pts[p] * pts[p]
    return st


def _worker(rank, world, port, G, q):
    import torch.distributed as dist
    from citadels_self_play_b200 import sharding as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = None
    for step in range(2):
        st = _stats_for(1234, S.first_gid(step, rank, world, G), G)
        if total is None:
            total = st
        else:
            for k in S.STAT_KEYS:
                total[k] += st[k]
            for k in ("wins", "points_sum", "points_sq"):
                total[k] = [a + b for a, b in zip(total[k], st[k])]
    red = S.reduce_stats(total)
    t = S.reduce_max([1.0 + rank, 5.0 - rank])
    dist.barrier()
    if rank == 0:
        q.put((red, t))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from citadels_self_play_b200 import sharding as S
    world, G = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, G, q)) for r in range(world)]
    for p in procs:
        p.start()
    red, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _stats_for(1234, 0, 2 * world * G)       # ids 0 .. 11 in one process
    assert red == want
    assert t == [2.0, 5.0]
    ids = sorted(S.first_gid(s, r, world, G) + i for s in range(2) for r in range(world) for i in range(G))
    assert ids == list(range(2 * world * G))


def test_stats_tensor_roundtrip():
    from citadels_self_play_b200 import sharding as S
    st = dict(games=3, steps=1200, steps_sq=500000, errors=0, wins=[1, 0, 0, 1, 0, 1], points_sum=[10, 20, 30, 40, 50, 60],
              points_sq=[100, 400, 900, 1600, 2500, 3600])
    assert S.tensor_to_stats(S.stats_to_tensor(st)) == st

"""Multi-GPU check of the MCCFR side (run under torchrun on N GPUs of one box):
     torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
  * datagen.get_mccfr_targets across ranks: roots sharded by global id, targets gathered (NCCL all_gather); every rank holds the
    same list, and it equals what one rank produces alone for the same global root ids
  * parallel.root_parallel_mccfr: same roots on every rank, roots' regrets / strategy / values pooled with all_reduce
Prints one JSON line on rank 0."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from citadels_self_play_b200 import Engine, datagen, parallel, sharding

eng = Engine(capacity=1024, device=local)
out = {"world": world}
# ---- gathered data generation (configs[4] shape, small): 2000-iteration trees, >= 200 backprops
stats = {}
t0 = time.perf_counter()
tg = datagen.get_mccfr_targets(None, minimum_sufficient_nodes=150, base_usefullness_treshold=200, pretrain=True, max_iterations=2000,
                               engine=eng, roots_per_batch=256, seed=4321, first_gid=50_000, stats=stats)
out["datagen"] = dict(stats, seconds=time.perf_counter() - t0)
digest = torch.tensor([len(tg), int(sum(float(x[2].sum()) for x in tg) * 1000) % (1 << 40)], dtype=torch.int64, device="cuda")
alld = [torch.empty_like(digest) for _ in range(world)]
dist.all_gather(alld, digest)
assert all(torch.equal(a, alld[0]) for a in alld), "ranks hold different target lists"
if rank == 0:   # the same global roots searched by ONE engine: same multiset of targets
    solo = []
    for b in range(stats["batches"]):
        for r in range(world):
            eng.make_roots(256, seed=4321, first_gid=50_000 + sharding.first_gid(b, r, world, 256), back_lo=1, back_hi=100, flavour=1)
            eng.mccfr(256, iterations=2000, seed=4321)
            solo += Engine.targets_as_tuples(eng.mccfr_targets(256, iterations=2000, seed=4321, threshold=200.0))
    assert len(solo) == len(tg)
    a = sorted(tuple(np.round(x[2].numpy(), 9)) for x in solo)
    b = sorted(tuple(np.round(x[2].numpy(), 9)) for x in tg)
    assert a == b, "gathered targets differ from a single-engine run over the same roots"
    out["datagen"]["equals_single_engine"] = True
dist.barrier()
# ---- root-parallel mode
eng.make_roots(512, seed=99, first_gid=9000, back_lo=0, back_hi=20)
t0 = time.perf_counter()
res = parallel.root_parallel_mccfr(eng, 512, iterations=2000, sync_every=200, seed=99)
dt = time.perf_counter() - t0
chk = torch.from_numpy(np.ascontiguousarray(res["node_value"])).cuda()
allv = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allv, chk)
assert all(torch.equal(a, allv[0]) for a in allv), "pooled root values differ between ranks"
out["root_parallel"] = dict(roots=512, iterations_per_rank=2000, sync_every=200, seconds=dt,
                            it_per_s_all_ranks=int(res["iterations"].sum()) * world / dt,
                            pooled_root_visits_mean=float(res["node_value"].sum(1).mean()))
if rank == 0:
    print(json.dumps(out), flush=True)
eng.close()
dist.destroy_process_group()

"""Developer tool: latency of single MCCFR trees (one root per launch) vs the batch -- is the batch bound by its slowest tree?
   python tools/tree_latency.py [roots] [sampled trees] [ruleset] [first_gid]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from citadels_self_play_b200 import Engine

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 192
RS = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # ruleset
G0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # first game id
e = Engine(capacity=R)
e.make_roots(R, seed=0xC17ADE15, first_gid=G0, ruleset=RS, back_lo=0, back_hi=20)
roots, knows, used, gids = e.store_roots(R)
o = e.mccfr(R, iterations=200, seed=0xC17ADE15, ruleset=RS)
o = e.mccfr(R, iterations=200, seed=0xC17ADE15, ruleset=RS)
batch_ms = o["kernel_ms"]
nodes = o["results"]["n_nodes"].astype(np.int64)
order = np.argsort(-nodes)
pick = np.concatenate([order[:K // 2], order[len(order) // 2:len(order) // 2 + K // 2]])   # the biggest trees and typical ones
e1 = Engine(capacity=8)
lat = []
for i in pick:
    e1.load_roots(roots[i:i + 1], knows[i:i + 1], used[i:i + 1], gids[i:i + 1])
    e1.mccfr(1, iterations=200, seed=0xC17ADE15, ruleset=RS)
    t = e1.mccfr(1, iterations=200, seed=0xC17ADE15, ruleset=RS)["kernel_ms"]
    lat.append(t)
lat = np.array(lat)
print(json.dumps({"roots": R, "batch_ms": batch_ms, "sum_single_ms_sampled": float(lat.sum()),
                  "single_ms_biggest": [round(float(x), 2) for x in lat[:8]], "nodes_biggest": [int(nodes[i]) for i in pick[:8]],
                  "single_ms_typical_median": float(np.median(lat[K // 2:])), "nodes_median": int(np.median(nodes)),
                  "single_ms_max": float(lat.max()), "mean_nodes": float(nodes.mean())}))

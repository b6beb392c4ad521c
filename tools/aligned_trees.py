"""Developer experiment: every warp searches the SAME root (same chance stream), i.e. all resident warps run one instruction stream
in near lock-step -- the upper bound of what instruction-cache locality could buy the search kernel.
   python tools/aligned_trees.py [roots] [distinct]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from citadels_self_play_b200 import Engine

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1
e = Engine(capacity=R)
e.make_roots(R, seed=0xC17ADE15, first_gid=0, back_lo=0, back_hi=20)
roots, knows, used, gids = e.store_roots(R)
o = e.mccfr(R, iterations=200, seed=0xC17ADE15)
o = e.mccfr(R, iterations=200, seed=0xC17ADE15)
base = int(o["results"]["iterations"].sum()) / o["kernel_ms"] * 1e3
nodes = o["results"]["n_nodes"].astype(np.int64)
order = np.argsort(nodes)
typical = order[len(order) // 2: len(order) // 2 + D]          # D trees of median size
idx = np.resize(typical, R)
e.load_roots(roots[idx].copy(), knows[idx].copy(), used[idx].copy(), gids[idx].copy())
e.mccfr(R, iterations=200, seed=0xC17ADE15)
o2 = e.mccfr(R, iterations=200, seed=0xC17ADE15)
rate = int(o2["results"]["iterations"].sum()) / o2["kernel_ms"] * 1e3
print(json.dumps({"roots": R, "distinct_trees": D, "nodes_of_the_replicated_trees": [int(nodes[i]) for i in typical[:8]],
                  "it_per_s_distinct_roots": base, "it_per_s_replicated": rate, "kernel_ms_replicated": o2["kernel_ms"]}))

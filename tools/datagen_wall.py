"""Developer tool: where the wall-clock of a 2000-iteration data-generation search goes (kernel event time vs the whole call)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from citadels_self_play_b200 import Engine

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
e = Engine(capacity=R)
e.make_roots(R, seed=0xC17ADE15, first_gid=R, back_lo=1, back_hi=100, flavour=1)
out = []
for rep in range(3):
    l0 = e.launches
    t0 = time.perf_counter()
    o = e.mccfr(R, iterations=IT, seed=0xC17ADE15)
    w = time.perf_counter() - t0
    out.append({"wall_ms": w * 1e3, "kernel_ms_first_pass": o["kernel_ms"], "launches": e.launches - l0,
                "status2": int((o["results"]["status"] & 2 != 0).sum()), "nodes": int(o["results"]["n_nodes"].sum()),
                "max_nodes": int(o["results"]["n_nodes"].max())})
for rep in range(3):
    t0 = time.perf_counter()
    tg = e.mccfr_targets(R, iterations=IT, seed=0xC17ADE15, threshold=200.0)
    out.append({"targets_export_ms": (time.perf_counter() - t0) * 1e3, "targets": len(tg["meta"])})
print(json.dumps(out))

"""Developer tool: search many roots of every ruleset on the GPU and count trees by status (none may be refused).
   python tools/status_scan_gpu.py [first_gid]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from citadels_self_play_b200 import Engine

BASE = int(sys.argv[1]) if len(sys.argv) > 1 else 7_000_000      # first game id: another base = other roots
plan = [(0, 32768, 200, 0), (0, 8192, 200, 1), (1, 16384, 200, 0), (2, 8192, 200, 0), (0, 2048, 2000, 1)]   # (ruleset, roots, iterations, flavour)
out = []
for rs, R, IT, fl in plan:
    e = Engine(capacity=R)
    e.make_roots(R, seed=0xC17ADE15, first_gid=BASE, ruleset=rs, back_lo=0 if fl == 0 else 1, back_hi=20 if fl == 0 else 100, flavour=fl)
    l0 = e.launches
    t0 = time.perf_counter()
    o = e.mccfr(R, iterations=IT, seed=0xC17ADE15, ruleset=rs)
    st = o["results"]["status"]
    out.append({"ruleset": rs, "roots": R, "iterations": IT, "flavour": fl, "wall_s": round(time.perf_counter() - t0, 3), "launches": e.launches - l0,
                "terminal_roots": int((st == 1).sum()), "status_2": int(((st & 2) != 0).sum()), "status_4": int(((st & 4) != 0).sum()),
                "status_16": int(((st & 16) != 0).sum()), "max_nodes": int(o["results"]["n_nodes"].max()), "max_children": int(o["results"]["n_children"].max())})
    print(json.dumps(out[-1]), flush=True)
    e.close()

"""Developer tool: wall-clock of Engine.mccfr (launch + status check + result records to the host) against its kernel event time."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from citadels_self_play_b200 import Engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = Engine(capacity=R)
e.make_roots(R, seed=0xC17ADE15, first_gid=0, back_lo=0, back_hi=20)
e.mccfr(R, iterations=200, seed=0xC17ADE15)
out = []
for i in range(3):
    t0 = time.perf_counter()
    o = e.mccfr(R, iterations=200, seed=0xC17ADE15)
    out.append((round((time.perf_counter() - t0) * 1e3, 2), round(o["kernel_ms"], 2)))
print(json.dumps({"roots": R, "wall_ms_vs_kernel_ms": out}))

"""Developer tool: time the fused playout kernel of one library build (CTD_LIB=path python tools/playout_perf.py [games])."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from citadels_self_play_b200 import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
ruleset = int(sys.argv[2]) if len(sys.argv) > 2 else 0
e = Engine(capacity=64)
best = None
for it in range(4):
    st = e.playout(n, seed=0xC17ADE15, first_gid=it * n, ruleset=ruleset, outputs=False)["stats"]
    r = st["steps"] / st["kernel_ms"] * 1e3
    if it and (best is None or r > best):
        best = r
    if it == 0:
        first = (st["steps"], st["wins"], st["errors"])
print(json.dumps({"lib": os.path.basename(os.environ.get("CTD_LIB", "default")), "steps_per_s": best, "check": first}))

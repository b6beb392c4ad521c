"""Developer tool: time the MCCFR kernels of one library build.
   CTD_LIB=path python tools/mccfr_perf.py [roots] [iterations] [pure|deep|both]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from citadels_self_play_b200 import Engine
from citadels_self_play_b200.value_model import ValueOnlyNN

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 200
what = sys.argv[3] if len(sys.argv) > 3 else "both"
reps = int(os.environ.get("REPS", "3"))
RS = int(sys.argv[4]) if len(sys.argv) > 4 else 0
e = Engine(capacity=R)
e.make_roots(R, seed=0xC17ADE15, first_gid=0, ruleset=RS, back_lo=0, back_hi=20)
out = {"lib": os.path.basename(os.environ.get("CTD_LIB", "default")), "roots": R, "iterations": IT, "ruleset": RS}
if what in ("pure", "both"):
    best = 0
    for i in range(reps):
        o = e.mccfr(R, iterations=IT, seed=0xC17ADE15, ruleset=RS)
        best = max(best, int(o["results"]["iterations"].sum()) / o["kernel_ms"] * 1e3)
    out["pure_it_per_s"] = best
    out["pure_check"] = [int(o["results"]["n_nodes"].sum()), int(o["results"]["rng_draws"].sum()), int((o["results"]["status"] > 1).sum())]
if what in ("deep", "both"):
    torch.manual_seed(0)
    e.set_value_model(ValueOnlyNN(418, 512).eval())
    e.set_value_backend(os.environ.get("CTD_BACKEND", "fused"))
    out["backend"] = os.environ.get("CTD_BACKEND", "fused")
    best = 0
    for i in range(reps):
        o = e.mccfr_pred(R, iterations=IT, max_depth=10, seed=0xC17ADE15, ruleset=RS)
        best = max(best, int(o["results"]["iterations"].sum()) / o["kernel_ms"] * 1e3)
    out["deep_it_per_s"] = best
    out["deep_check"] = [int(o["results"]["n_nodes"].sum()), int(o["results"]["rng_draws"].sum()), int((o["results"]["status"] > 1).sum())]
print(json.dumps(out))

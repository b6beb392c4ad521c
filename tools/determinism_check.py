"""Developer tool: the search kernels run the scalar walk on all 32 lanes converged; repeated runs must be bit-identical."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from citadels_self_play_b200 import Engine
from citadels_self_play_b200.value_model import ValueOnlyNN
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
e = Engine(capacity=R)
out = {}
for name, kw in (("pure200", dict(iterations=200)), ("pure2000", dict(iterations=2000, n=512))):
    n = kw.pop("n", R)
    e.make_roots(n, seed=99, first_gid=0, back_lo=0, back_hi=60)
    ref = None; same = True
    for rep in range(3):
        r = e.mccfr(n, seed=99, **kw)["results"]
        b = r.tobytes()
        if ref is None: ref = b
        same &= (b == ref)
    out[name] = bool(same)
torch.manual_seed(0); e.set_value_model(ValueOnlyNN(418, 512).eval())
e.make_roots(R, seed=99, first_gid=0, back_lo=0, back_hi=60)
ref = None; same = True
for rep in range(3):
    r = e.mccfr_pred(R, iterations=200, max_depth=10, seed=99)["results"]
    b = r.tobytes()
    if ref is None: ref = b
    same &= (b == ref)
out["deep200"] = bool(same)
print(json.dumps(out))

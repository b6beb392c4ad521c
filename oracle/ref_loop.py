"""The reference's own loops, run from oracle/_ref (the byte-compiled REAL reference, see build_ref.py) -- the CPU arm of
bench.py (`--impl reference`, `cpu_baseline`).  TEST / BENCH INFRASTRUCTURE: never imported by the product package.

  playouts(n, seed)                 run_utils.create_game() + the loop at run_utils.py:37-41 (get_options_from_state,
                                    random.choice, carry_out) to terminal, n games under random.seed(seed)
  pure_mccfr(n, seed, iterations)   run_utils.run_mccfr(create_a_close_to_finished_game(create_game()), max_iterations)
  deep_mccfr(n, seed, iterations)   CFRNode(game, original_player_id=..., model=ValueOnlyNN(418, 512).eval() under
                                    torch.manual_seed(0), training=False, device="cpu").cfr_pred(iterations, max_depth=10)
                                    (run_mccfr itself hard-codes device "cuda:0")
All draw from CPython's `random` / numpy's global RNG exactly as the reference does; nothing is patched except the two
plotting imports (seaborn / matplotlib, used only by plot helpers) that are absent from this image."""
import os
import sys
import time
import types

import importlib.abc
import importlib.util
import marshal

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_loaded = {}


def available():
    return os.path.isfile(os.path.join(_REF, "run_utils.pyc.bin"))


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports run_utils, game.* and algorithms.* from oracle/_ref/**/*.pyc.bin (py_compile output of the reference's modules)."""

    def find_spec(self, name, path=None, target=None):
        parts = name.split(".")
        if parts[0] not in ("run_utils", "game", "algorithms"):
            return None
        base = os.path.join(_REF, *parts)
        if os.path.isdir(base):
            return importlib.util.spec_from_loader(name, self, is_package=True)
        if os.path.isfile(base + ".pyc.bin"):
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        base = os.path.join(_REF, *module.__name__.split("."))
        if os.path.isdir(base):
            module.__path__ = [base]
            return
        with open(base + ".pyc.bin", "rb") as f:
            code = marshal.loads(f.read()[16:])      # 16-byte .pyc header: magic, flags, mtime, size
        module.__file__ = base + ".pyc.bin"
        exec(code, module.__dict__)


def load():
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is not built (python -m oracle.build_ref in the build container)")
    for m in ("seaborn", "matplotlib", "matplotlib.pyplot"):      # plotting-only imports of algorithms/train_utils.py, train.py
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.setrecursionlimit(5000)                                    # as train_from_scratch.py:17 (recursive backpropagate)
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    import run_utils
    from algorithms.deep_mccfr import CFRNode
    from algorithms.models import ValueOnlyNN
    _loaded.update(run_utils=run_utils, CFRNode=CFRNode, ValueOnlyNN=ValueOnlyNN)
    return _loaded


def playouts(n, seed):
    """-> (env steps, games, seconds, winners[6])"""
    import random
    ru = load()["run_utils"]
    random.seed(seed)
    steps, wins = 0, [0] * 6
    t0 = time.perf_counter()
    for _ in range(n):
        game = ru.create_game()
        winner = False
        while not winner:                                          # run_utils.py:37-41
            options = game.get_options_from_state()
            chosen_option = random.choice(options)
            winner = chosen_option.carry_out(game)
            steps += 1
        wins[winner.id] += 1
    return steps, n, time.perf_counter() - t0, wins


def pure_mccfr(n, seed, iterations=200):
    """-> (iterations run, roots searched, seconds)"""
    import random
    import numpy as np
    ru = load()["run_utils"]
    random.seed(seed)
    np.random.seed(seed & 0xFFFFFFFF)
    games = []
    while len(games) < n:                                          # roots are made outside the timed region
        g = ru.create_a_close_to_finished_game(ru.create_game())
        if not g.terminal:
            games.append(g)
    t0 = time.perf_counter()
    for g in games:
        ru.run_mccfr(g, max_iterations=iterations)
    return iterations * n, n, time.perf_counter() - t0


def deep_mccfr(n, seed, iterations=200, max_depth=10):
    import random
    import numpy as np
    import torch
    L = load()
    ru = L["run_utils"]
    torch.set_num_threads(1)
    torch.manual_seed(0)
    model = L["ValueOnlyNN"](418, 512).eval()
    random.seed(seed)
    np.random.seed(seed & 0xFFFFFFFF)
    games = []
    while len(games) < n:
        g = ru.create_a_close_to_finished_game(ru.create_game())
        if not g.terminal:
            games.append(g)
    t0 = time.perf_counter()
    for g in games:
        root = L["CFRNode"](g, original_player_id=g.gamestate.player_id, model=model, training=False, device="cpu")
        root.cfr_pred(max_iterations=iterations, max_depth=max_depth)
    return iterations * n, n, time.perf_counter() - t0

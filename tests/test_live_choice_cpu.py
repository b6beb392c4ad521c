"""run_mccfr's decision -- CFRNode.action_choice(live=True) (algorithms/deep_mccfr.py:67-75, game/game.py:312-317) -- against
the real reference: tests/golden/live_choice_*.npz hold the option the unmodified run_utils.run_mccfr returned for 328 roots
(36 of them role-pick roots, where the reference indexes a per-child strategy row by the ranks on offer), recorded by
tests/golden/gen_golden.py (`live`).  Checked here: the oracle's restatement and the host build of the kernels' code."""
import ctypes
import os
import subprocess
import numpy as np
import pytest

from oracle import citadels_oracle as O
from oracle import mccfr_oracle as M
from oracle.philox import PhiloxChance
from tests.golden_util import GOLDEN

HERE = os.path.dirname(os.path.abspath(__file__))
PURE = ["live_choice_preset.npz", "live_choice_preset_early.npz", "live_choice_classic.npz", "live_choice_random.npz"]
J_MASK = np.uint64(~(0x3FF << 51) & 0xFFFFFFFFFFFFFFFF)


def same_option(a, b):
    """Descriptors equal; the ordinal of a discard_and_draw / cardinal_exchange option is not recoverable from a single
    reference option object (it is its position among its siblings), so it is masked for those two kinds."""
    a, b = np.uint64(a), np.uint64(b)
    if int(a) & 0x3F in (O.K["discard_and_draw"], O.K["cardinal_exchange"]):
        return (a & J_MASK) == (b & J_MASK)
    return a == b


def load(name):
    from citadels_self_play_b200.layout import know_from_v1
    with np.load(os.path.join(GOLDEN, name)) as f:
        z = {k: f[k] for k in f.files}
    z["knows"] = know_from_v1(z["knows"])   # the fixtures hold knowledge blocks in the round-1 entry format
    return z


@pytest.mark.parametrize("name", PURE)
def test_oracle_live_choice_matches_reference(name):
    z = load(name)
    seed, iters = int(z["seed"]), int(z["iterations"])
    checked = rp = 0
    for r in range(len(z["gids"])):
        if z["terminal"][r] or (r % 3 and not z["role_pick"][r]):      # every role-pick root, a third of the others
            continue
        ch = PhiloxChance(seed, int(z["gids"][r]), stream=1)
        g = O.Game.unpack(z["roots"][r].tobytes(), ch)
        g.unpack_know(z["knows"][r], z["used"][r])
        n = M.Node(g, g.player)
        n.cfr_train(iters)
        d = n.live_choice()
        assert same_option(d, z["live"][r]), (name, r, hex(d), hex(int(z["live"][r])))
        assert n.role_pick == bool(z["role_pick"][r]) and len(n.children) == int(z["nchild"][r])
        assert ch.i == int(z["draws"][r])          # the decision is the stream's next draw after the search
        checked += 1
        rp += n.role_pick
    assert checked >= 8 and (rp >= 3 or "early" not in name)


def test_oracle_terminal_root_raises_like_the_reference():
    z = load("live_choice_preset.npz")
    r = int(np.flatnonzero(z["terminal"])[0])
    g = O.Game.unpack(z["roots"][r].tobytes(), PhiloxChance(int(z["seed"]), int(z["gids"][r]), stream=1))
    g.unpack_know(z["knows"][r], z["used"][r])
    n = M.Node(g, g.player)
    n.cfr_train(20)
    with pytest.raises(ValueError):
        n.live_choice()


@pytest.fixture(scope="module")
def hostsim():
    d = os.path.join(HERE, "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    lib = ctypes.CDLL(os.path.join(d, "libctd_hostsim.so"))
    u64, u32, vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
    lib.hs_mccfr.argtypes = [vp, vp, vp, u64, u64, u32, vp, u64, vp, u64, vp]
    lib.hs_last_live_option.restype = u64
    return lib


@pytest.mark.parametrize("name", PURE)
def test_kernel_live_choice_host_build_matches_reference(hostsim, name):
    """ctd_live_choice (what ctd_mccfr_result.live_option carries), compiled for the host, on every root of the fixtures."""
    z = load(name)
    arena = np.zeros(192 << 20, np.uint8)
    nb = ctypes.c_uint64()
    for r in range(len(z["gids"])):
        root, know, used = (np.ascontiguousarray(z[k][r]) for k in ("roots", "knows", "used"))
        st = hostsim.hs_mccfr(root.ctypes.data, know.ctypes.data, used.ctypes.data, int(z["seed"]), int(z["gids"][r]),
                              int(z["iterations"]), arena.ctypes.data, arena.nbytes, None, 0, ctypes.byref(nb))
        d = hostsim.hs_last_live_option()
        if z["terminal"][r]:
            assert st == 1 and d == 0
            continue
        assert st == 0
        assert same_option(d, z["live"][r]), (name, r, hex(d), hex(int(z["live"][r])))

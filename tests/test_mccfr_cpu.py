"""CPU suite for the MCCFR path: oracle and host build of the kernel code against trees grown by the REAL
reference's CFRNode.cfr_train (tests/golden/mccfr_*.npz, made by gen_golden.py)."""
import ctypes
import os
import subprocess
import numpy as np
import pytest

from oracle import mccfr_oracle as M
from tests.mccfr_util import MccfrGolden, oracle_preorder, tree_preorder, assert_same_tree

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = ["mccfr_preset.npz", "mccfr_preset_deep_back.npz", "mccfr_classic.npz"]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_trees_match_reference(name):
    G = MccfrGolden(name)
    z = G.z
    for r in range(G.n):
        g, step = M.make_root(G.seed, int(G.gids[r]), G.ruleset, 0, G.back_hi)
        assert step == int(z["root_step"][r])
        assert g.pack()[:228] == z["roots"][r][:228].tobytes()
        assert g.pack_know(g.player) == z["knows"][r].tobytes()
        assert bytes(g.used_cards) == z["used"][r].tobytes()
        if z["terminal"][r]:
            continue
        n = M.run_from_root(z["roots"][r], z["knows"][r], z["used"][r], G.seed, int(G.gids[r]), G.iterations)
        assert_same_tree(G.nodes(r), oracle_preorder(n), (name, r))


@pytest.fixture(scope="module")
def hostsim():
    d = os.path.join(HERE, "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    lib = ctypes.CDLL(os.path.join(d, "libctd_hostsim.so"))
    u64, u32, vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
    lib.hs_mccfr.argtypes = [vp, vp, vp, u64, u64, u32, u32, u32, u32, vp]
    return lib


@pytest.mark.parametrize("name", FIXTURES)
def test_kernel_mccfr_host_build_matches_reference(hostsim, name):
    """ctd_mccfr.cuh (the code ctd_k_mccfr runs), compiled for the host, against the reference's trees."""
    from citadels_self_play_b200.layout import TreeView, tree_bytes
    G = MccfrGolden(name)
    z = G.z
    extra = 8192 if G.ruleset == 1 else 0
    mn = 6 * G.iterations + 256 + extra
    cc = mn + 10 * (G.iterations + 2)
    ac = 3 * cc + 180 * 64
    buf = np.zeros(tree_bytes(mn, cc, ac), np.uint8)
    for r in range(G.n):
        root, know, used = (np.ascontiguousarray(z[k][r]) for k in ("roots", "knows", "used"))
        st = hostsim.hs_mccfr(root.ctypes.data, know.ctypes.data, used.ctypes.data, G.seed, int(G.gids[r]), G.iterations,
                              mn, cc, ac, buf.ctypes.data)
        if z["terminal"][r]:
            assert st == 1
            continue
        assert st == 0
        assert_same_tree(G.nodes(r), tree_preorder(TreeView(buf, mn, cc, ac)), (name, r))

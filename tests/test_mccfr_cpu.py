"""CPU suite for the MCCFR path: oracle and host build of the kernel code against trees grown by the REAL
reference's CFRNode.cfr_train (tests/golden/mccfr_*.npz, made by gen_golden.py)."""
import ctypes
import os
import subprocess
import numpy as np
import pytest

from oracle import mccfr_oracle as M
from tests.golden_util import visible
from tests.mccfr_util import MccfrGolden, oracle_preorder, tree_preorder, assert_same_tree

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = ["mccfr_preset.npz", "mccfr_preset_deep_back.npz", "mccfr_classic.npz", "mccfr_random.npz", "mccfr_preset_2000it.npz"]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_trees_match_reference(name):
    G = MccfrGolden(name)
    z = G.z
    # the Python oracle does ~100 iterations/s: every root of the 200-iteration sets, half of the 2000-iteration set
    roots = range(0, G.n, 2) if G.iterations > 200 else range(G.n)
    for r in roots:
        g, step = M.make_root(G.seed, int(G.gids[r]), G.ruleset, 0, G.back_hi)
        assert step == int(z["root_step"][r])
        assert visible(g.pack()) == visible(z["roots"][r].tobytes())
        assert g.pack_know(g.player) == z["knows"][r].tobytes()
        assert bytes(g.used_cards).ljust(76, b"\xff") == z["used"][r].tobytes()
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
    lib.hs_mccfr.argtypes = [vp, vp, vp, u64, u64, u32, vp, u64, vp, u64, vp]
    return lib


ARENA = np.zeros(192 << 20, np.uint8)     # what one host-built tree allocates from
OUT = np.zeros(64 << 20, np.uint8)        # its export block


@pytest.mark.parametrize("name", FIXTURES)
def test_kernel_mccfr_host_build_matches_reference(hostsim, name):
    """ctd_mccfr.cuh (the code ctd_k_mccfr runs), compiled for the host, against the reference's trees."""
    from citadels_self_play_b200.layout import TreeView
    G = MccfrGolden(name)
    z = G.z
    nb = ctypes.c_uint64()
    for r in range(G.n):
        root, know, used = (np.ascontiguousarray(z[k][r]) for k in ("roots", "knows", "used"))
        st = hostsim.hs_mccfr(root.ctypes.data, know.ctypes.data, used.ctypes.data, G.seed, int(G.gids[r]), G.iterations,
                              ARENA.ctypes.data, ARENA.nbytes, OUT.ctypes.data, OUT.nbytes, ctypes.byref(nb))
        if z["terminal"][r]:
            assert st == 1
            continue
        assert st == 0
        assert_same_tree(G.nodes(r), tree_preorder(TreeView(OUT[:nb.value])), (name, r))


def test_kernel_tree_memory_chunks_and_exhaustion(hostsim, monkeypatch):
    """Tree memory: a tree that outgrows its first chunk continues in further chunks (same tree, node for node), and an
    exhausted arena is reported as status 2 (the engine then searches the tree again from a larger arena), never a crash."""
    from citadels_self_play_b200.layout import TreeView
    G = MccfrGolden("mccfr_preset.npz")
    z = G.z
    r = next(i for i in range(G.n) if not z["terminal"][i])
    root, know, used = (np.ascontiguousarray(z[k][r]) for k in ("roots", "knows", "used"))
    nb = ctypes.c_uint64()
    args = (root.ctypes.data, know.ctypes.data, used.ctypes.data, G.seed, int(G.gids[r]), G.iterations)
    monkeypatch.setenv("HS_N0_LOG2", "3")    # eight nodes in chunk 0: the tree lives almost entirely behind the chunk table
    st = hostsim.hs_mccfr(*args, ARENA.ctypes.data, ARENA.nbytes, OUT.ctypes.data, OUT.nbytes, ctypes.byref(nb))
    assert st == 0
    assert_same_tree(G.nodes(r), tree_preorder(TreeView(OUT[:nb.value])), ("chunked", r))
    monkeypatch.delenv("HS_N0_LOG2")
    st = hostsim.hs_mccfr(*args, ARENA.ctypes.data, 1 << 20, None, 0, ctypes.byref(nb))   # chunk 0 alone needs 1.3 MB
    assert st & 2
    monkeypatch.setenv("HS_N0_LOG2", "5")
    st = hostsim.hs_mccfr(*args, ARENA.ctypes.data, 300 << 10, None, 0, ctypes.byref(nb))   # runs dry mid-search
    assert st & 2


def test_kernel_knowledge_block_holds_a_busy_seer(hostsim):
    """Two Game(preset=False) roots (seed 0xC17ADE15, gids 7003168 and 7006460, stepped back 0..20) whose hypothetical games collect
    more than 32 hand-knowledge entries: refused with status 4 while the knowledge block held 32 entries, complete now and equal to
    the oracle's trees node for node."""
    from citadels_self_play_b200.layout import TreeView
    u64, u32, vp, i32 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int
    hostsim.hs_make_root.argtypes = [u64, u64, i32, u32, u32, i32, vp, vp, vp, vp]
    root, know, used, step = np.zeros(256, np.uint8), np.zeros(592, np.uint8), np.zeros(76, np.uint8), np.zeros(1, np.uint32)
    SEED = 0xC17ADE15
    nb = ctypes.c_uint64()
    most = 0
    for gid in (7003168, 7006460):
        hostsim.hs_make_root(SEED, gid, 2, 0, 20, 0, root.ctypes.data, know.ctypes.data, used.ctypes.data, step.ctypes.data)
        st = hostsim.hs_mccfr(root.ctypes.data, know.ctypes.data, used.ctypes.data, SEED, gid, 200, ARENA.ctypes.data, ARENA.nbytes,
                              OUT.ctypes.data, OUT.nbytes, ctypes.byref(nb))
        assert st == 0, (gid, st)
        tv = TreeView(OUT[:nb.value].copy())
        most = max(most, int(tv.nodes["know"]["n_hk"].max()))
        node = M.run_from_root(root, know, used, SEED, gid, 200)
        assert_same_tree(oracle_preorder(node), tree_preorder(tv), ("busy seer", gid))
    assert most > 32


# ---------------------------------------------------------------- deep MCCFR (config 4)
def _model():
    import torch
    from citadels_self_play_b200.value_model import ValueOnlyNN
    torch.set_num_threads(1)
    torch.manual_seed(0)
    return ValueOnlyNN(418, 512).eval()


DEEP = ["deep_mccfr_preset.npz", "deep_mccfr_classic.npz", "deep_mccfr_random.npz"]


@pytest.mark.parametrize("name", DEEP)
def test_oracle_deep_trees_match_reference(name):
    """cfr_pred(200, max_depth=10) with ValueOnlyNN(418,512) under torch.manual_seed(0): oracle vs the real reference."""
    from citadels_self_play_b200.value_model import reference_value
    from oracle import citadels_oracle as O
    from oracle.philox import PhiloxChance
    G = MccfrGolden(name)
    z = G.z
    model = _model()

    def value(g):
        return reference_value(model, np.asarray(g.encode_game(), dtype=np.float32)[None, :], weight=1.0)[0]
    for r in range(0, G.n, 2):
        if z["terminal"][r]:
            continue
        g = O.Game.unpack(z["roots"][r].tobytes(), PhiloxChance(G.seed, int(G.gids[r]), stream=1))
        g.unpack_know(z["knows"][r], z["used"][r])
        n = M.Node(g, g.player, model=value)
        n.cfr_pred(G.iterations, int(z["max_depth"]))
        assert_same_tree(G.nodes(r), oracle_preorder(n), ("deep", r), rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("name", DEEP)
def test_kernel_deep_mccfr_host_build_matches_reference(hostsim, name):
    from citadels_self_play_b200.layout import TreeView
    from citadels_self_play_b200.value_model import reference_value
    G = MccfrGolden(name)
    z = G.z
    model = _model()
    u64, u32, vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
    EVAL = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float))
    hostsim.hs_mccfr_pred.argtypes = [vp, vp, vp, u64, u64, u32, u32, vp, u64, vp, u64, vp, EVAL]

    def ev(fp, pp):
        p = reference_value(model, np.ctypeslib.as_array(fp, shape=(448,))[None, :])[0]
        for i in range(6):
            pp[i] = float(p[i])
    cb = EVAL(ev)
    nb = ctypes.c_uint64()
    for r in range(G.n):
        root, know, used = (np.ascontiguousarray(z[k][r]) for k in ("roots", "knows", "used"))
        if z["terminal"][r]:
            continue
        st = hostsim.hs_mccfr_pred(root.ctypes.data, know.ctypes.data, used.ctypes.data, G.seed, int(G.gids[r]),
                                   G.iterations, int(z["max_depth"]), ARENA.ctypes.data, ARENA.nbytes, OUT.ctypes.data, OUT.nbytes,
                                   ctypes.byref(nb), cb)
        assert st == 0
        assert_same_tree(G.nodes(r), tree_preorder(TreeView(OUT[:nb.value])), ("deep-host", r), rtol=1e-7, atol=1e-10)


def test_oracle_encode_game_layout():
    """Feature offsets of SURVEY.md Appendix D on a hand-built state."""
    from oracle import citadels_oracle as O
    g = O.Game(None, deal=False)
    g.role = [0, 1, 2, 3, 4, 5]
    g.kr_conf = [[q in (1, 4) for q in range(6)] for _ in range(6)]
    g.bld[2] = [13, 13, 40]
    g.gold[3] = -2
    g.hand[4] = [1, 2, 3]
    g.state, g.player, g.ending = 5, 2, True
    g.possessed[6] = True
    f = g.encode_game()
    assert len(f) == 418
    assert [i for i in range(24) if f[i]] == [r * 3 + g.variant[r] for r in range(8)]
    assert [i for i in range(24, 72) if f[i]] == [24 + 1 * 8 + 1, 24 + 4 * 8 + 4]
    assert f[72 + 2] == 4 + 4 + 6 and f[78 + 3] == -2 and f[84 + 4] == 3
    assert f[90 + 2 * 40 + 13] == 2 and f[90 + 2 * 40 + 25] == 1
    assert f[330 + 2 * 5 + 3] == 2 and f[330 + 2 * 5 + 0] == 1      # the rewritten Magic School counts as trade
    assert f[360 + 2] == 1 and f[366 + 5] == 1 and f[377] == 1 and f[378 + 6 * 5 + 2] == 1


# ---------------------------------------------------------------- training targets (get_all_targets)
@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_training_targets_match_reference(name):
    """CFRNode.get_all_targets(): 418 features, 131-wide option encodings, node values and regret rows of the real
    reference vs the oracle's restatement (including encode_option's elif-chain quirks)."""
    G = MccfrGolden(name)
    z = G.z
    koff = np.concatenate([[0], np.cumsum(z["t_k"])])
    for r in (range(0, G.n, 2) if G.iterations > 200 else range(G.n)):
        if z["terminal"][r]:
            continue
        n = M.run_from_root(z["roots"][r], z["knows"][r], z["used"][r], G.seed, int(G.gids[r]), G.iterations)
        tg = M.get_all_targets(n, G.seed, int(G.gids[r]))
        lo, hi = int(z["t_off"][r]), int(z["t_off"][r + 1])
        assert len(tg) == hi - lo
        for j, (feat, opts, val, dist) in enumerate(tg):
            i = lo + j
            assert np.array_equal(np.asarray(feat, dtype=np.float32), z["t_feat"][i])
            k0, k1 = int(koff[i]), int(koff[i + 1])
            ref_opts = z["t_opts"][k0:k1]
            got = np.asarray(opts, dtype=np.float32)
            if G.ruleset == 1:   # discard_and_draw descriptors carry no card list: compare everything but the card bits
                mask = np.ones(131, bool)
                mask[89:129] = False
                rows = ref_opts[:, O_DISCARD] == 1
                assert np.array_equal(got[~rows], ref_opts[~rows]) and np.array_equal(got[rows][:, mask], ref_opts[rows][:, mask])
            else:
                assert np.array_equal(got, ref_opts), (r, j)
            assert np.allclose(val, z["t_val"][i], rtol=1e-12)
            assert np.allclose(dist, z["t_dist"][k0:k1], rtol=1e-9, atol=1e-12)


O_DISCARD = 25   # index of discard_and_draw in game/option.py:34-45


@pytest.mark.parametrize("ruleset,flavour,lo,hi,gids", [
    (0, 0, 0, 20, range(9000, 9012)), (0, 1, 1, 100, range(9100, 9112)),
    (1, 0, 0, 60, range(109000, 109006)), (2, 0, 0, 80, list(range(209000, 209008)) + [204827]),
    (2, 1, 1, 100, range(209100, 209106))])
def test_kernel_root_construction_matches_oracle(hostsim, ruleset, flavour, lo, hi, gids):
    """ctd_make_roots (both flavours: run_utils.create_a_close_to_finished_game / create_a_random_game) against the oracle's
    make_root: root record, the searching player's knowledge block, used_cards, root step."""
    from tests.golden_util import visible
    u64, u32, vp, i32 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int
    hostsim.hs_make_root.argtypes = [u64, u64, i32, u32, u32, i32, vp, vp, vp, vp]
    root, know, used, step = np.zeros(256, np.uint8), np.zeros(592, np.uint8), np.zeros(76, np.uint8), np.zeros(1, np.uint32)
    SEED = 0xC17ADE15
    for gid in gids:
        viewer = hostsim.hs_make_root(SEED, gid, ruleset, lo, hi, flavour, root.ctypes.data, know.ctypes.data, used.ctypes.data,
                                      step.ctypes.data)
        g, steps = M.make_root(SEED, gid, ruleset, lo, hi, flavour)
        assert steps == int(step[0]), (gid, steps, int(step[0]))
        assert visible(g.pack()) == visible(root.tobytes()), gid
        if not g.terminal:
            assert viewer == g.player
            assert g.pack_know(viewer) == know.tobytes(), gid
            assert bytes(g.used_cards) + b"\xff" * (76 - len(g.used_cards)) == used.tobytes(), gid

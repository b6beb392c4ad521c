"""CPU suite: the oracle (oracle/citadels_oracle.py) and the host build of the kernel rules code
(tests/hostsim) against the golden fixtures the real reference produced."""
import ctypes
import os
import subprocess
import zlib
import numpy as np
import pytest

from oracle import citadels_oracle as O
from oracle.philox import PhiloxChance, TapeChance, philox4x32_10
from tests.golden_util import Traces, visible

HERE = os.path.dirname(os.path.abspath(__file__))


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter = key = 0 / all ones / pi digits
    assert philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def _oracle_replay(T, g, use_tape):
    gid = int(T.gids[g])
    sl = T.game_steps(g)
    chosen = T.chosen[sl]
    if use_tape:
        ch = TapeChance(T.game_tape(g), chosen)
    else:
        ch = PhiloxChance(T.seed, gid)
    og = O.new_game(ch, T.ruleset)
    for k in range(sl.start, sl.stop):
        opts = og.options()
        assert len(opts) == T.nopt[k]
        assert zlib.crc32(visible(og.pack())) == int(T.state_crc[k]), (g, k - sl.start)
        assert zlib.crc32(np.asarray(opts, dtype="<u8").tobytes()) == int(T.opts_crc[k]), (g, k - sl.start)
        i = ch.randbelow(len(opts))
        assert i == T.chosen[k]
        won = og.apply(opts[i])
        assert won == (k == sl.stop - 1)
    assert og.pack()[:228] == T.final[g][:228].tobytes()


@pytest.mark.parametrize("name,count", [("preset_traces.npz", 60), ("classic_traces.npz", 30), ("random_traces.npz", 40)])
def test_oracle_replays_golden_philox(name, count):
    T = Traces(name)
    for g in range(0, len(T), max(1, len(T) // count)):
        _oracle_replay(T, g, use_tape=False)


@pytest.mark.parametrize("name,count", [("preset_traces.npz", 20), ("classic_traces.npz", 10), ("random_traces.npz", 15)])
def test_oracle_replays_golden_tape(name, count):
    T = Traces(name)
    for g in range(3, len(T), max(1, len(T) // count)):
        _oracle_replay(T, g, use_tape=True)


@pytest.mark.parametrize("name", ["preset_full.npz", "classic_full.npz", "random_full.npz"])
def test_oracle_full_states_and_descriptors(name):
    """Fixtures that keep every packed state and every descriptor: byte-for-byte, plus pack/unpack round trip."""
    T = Traces(name)
    for g in range(len(T)):
        ch = PhiloxChance(T.seed, int(T.gids[g]))
        og = O.new_game(ch, T.ruleset)
        sl = T.game_steps(g)
        for k in range(sl.start, sl.stop):
            opts = og.options()
            rec = og.pack()
            assert visible(rec) == visible(T.states[k].tobytes())
            assert opts == [int(x) for x in T.descs[T.desc_off[k]:T.desc_off[k + 1]]]
            rt = O.Game.unpack(rec)
            assert visible(rt.pack()) == visible(rec)
            if og.state not in (8, 9):   # the Seer's / Scholar's enumerations are not pure (chance draws, shrinking list)
                assert rt.options() == opts
            og.apply(opts[ch.randbelow(len(opts))])


# ---------------------------------------------------------------- host build of the kernel rules code
@pytest.fixture(scope="module")
def hostsim():
    d = os.path.join(HERE, "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    lib = ctypes.CDLL(os.path.join(d, "libctd_hostsim.so"))
    u64, u32, vp, ci = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int
    lib.hs_new_game.argtypes = [u64, u64, ci, vp, u32, vp]
    lib.hs_enumerate.argtypes = [vp, vp, u32, vp, u64, u64, vp, u32]
    lib.hs_step.argtypes = [vp, u64, u64, u64, vp, u32]
    lib.hs_playout.argtypes = [u64, u64, ci, u32, vp, vp, vp, vp]
    return lib


@pytest.mark.parametrize("name,stride", [("preset_traces.npz", 4), ("classic_traces.npz", 3), ("random_traces.npz", 2)])
def test_kernel_rules_host_build_replays_tapes(hostsim, name, stride):
    """ctd_engine.cuh (the code the kernels run), compiled for the host, against recorded reference games."""
    T = Traces(name)
    opts = np.zeros(8192, np.uint64)
    err = np.zeros(1, np.uint8)
    for g in range(0, len(T), stride):
        tape = np.ascontiguousarray(T.game_tape(g))
        st = np.zeros(256, np.uint8)
        gid = int(T.gids[g])
        hostsim.hs_new_game(T.seed, gid, T.ruleset, tape.ctypes.data, len(tape), st.ctypes.data)
        sl = T.game_steps(g)
        for k in range(sl.start, sl.stop):
            n = hostsim.hs_enumerate(st.ctypes.data, opts.ctypes.data, 8192, err.ctypes.data, T.seed, gid, tape.ctypes.data, len(tape))
            assert n == T.nopt[k] and err[0] == 0
            assert zlib.crc32(visible(st.tobytes())) == int(T.state_crc[k])
            assert zlib.crc32(opts[:n].astype("<u8").tobytes()) == int(T.opts_crc[k])
            hostsim.hs_step(st.ctypes.data, int(opts[T.chosen[k]]), T.seed, gid, tape.ctypes.data, len(tape))
        assert st[:228].tobytes() == T.final[g][:228].tobytes()


@pytest.mark.parametrize("name", ["preset_traces.npz", "classic_traces.npz", "random_traces.npz"])
def test_kernel_rules_host_build_fused_playout(hostsim, name):
    T = Traces(name)
    for g in range(len(T)):
        pts = np.zeros(6, np.int8)
        steps = np.zeros(1, np.uint32)
        err = np.zeros(1, np.uint8)
        fin = np.zeros(256, np.uint8)
        hostsim.hs_playout(T.seed, int(T.gids[g]), T.ruleset, 4096, pts.ctypes.data, steps.ctypes.data, err.ctypes.data,
                           fin.ctypes.data)
        assert err[0] == 0 and int(steps[0]) == int(T.step_off[g + 1] - T.step_off[g])
        assert fin[:228].tobytes() == T.final[g][:228].tobytes()


def test_oracle_known_answers_appendix_a():
    """Worked micro-examples of SURVEY.md A.4b (measured on the reference)."""
    g = O.Game(PhiloxChance(1, 1), deal=False)
    # Magician discard quirk: hand [0,1,2,3,4], deck [13,14,15,13] -> hand [1,3,13,14], deck [15,13,0,2,4]
    g.variant = [0] * 8
    g.role = [2, 0, 1, 3, 4, 5]
    g.hand[0] = [0, 1, 2, 3, 4]
    g.deck = [13, 14, 15, 13]
    g.state, g.player = 5, 0
    opts = g._character_options(0)
    assert sum(1 for d in opts if O.d_kind(d) == O.K["discard_and_draw"]) == 31
    g.apply(next(d for d in opts if O.d_kind(d) == O.K["discard_and_draw"]))
    assert g.hand[0] == [1, 3, 13, 14] and g.deck == [15, 13, 0, 2, 4]
    g.hand[0] = [0, 0, 1, 0, 2, 0, 3]
    g.done = 0
    assert sum(1 for d in g._character_options(0) if O.d_kind(d) == O.K["discard_and_draw"]) == 127
    # Factory: uniques cost +1 to be offered, printed cost is paid
    g = O.Game(PhiloxChance(1, 1), deal=False)
    g.role = [5, 0, 1, 2, 3, 4]
    g.variant = list(O.RULESET_VARIANTS[1])
    g.bld[0] = [35]
    g.hand[0] = [19]
    g.state, g.player, g.gold[0] = 5, 0, 2
    assert not any(O.d_kind(d) == O.K["build"] for d in g.options())
    g.gold[0] = 3
    b = [d for d in g.options() if O.d_kind(d) == O.K["build"]]
    assert len(b) == 1
    g.apply(b[0])
    assert g.gold[0] == 1

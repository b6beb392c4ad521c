"""Edge cases of the C ABI on the device: empty batches, capacity and argument errors as status codes (never a crash),
too-small option strides, the step cap, terminal inputs, ragged batches."""
import ctypes
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from citadels_self_play_b200 import Engine
    e = Engine(capacity=256)
    yield e
    e.close()


def test_empty_batches_are_no_ops(engine):
    engine.reset(0)
    opts, counts = engine.enumerate(0)
    assert opts.shape[0] == 0 and counts.shape == (0,)
    assert engine.step(np.zeros(0, np.uint64)).shape == (0,)
    out = engine.playout(0)
    assert out["stats"]["games"] == 0 and out["stats"]["steps"] == 0 and len(out["winner"]) == 0


def test_argument_and_capacity_errors_are_status_codes(engine):
    lib, h = engine._lib, engine._h
    buf = np.zeros(8, np.uint64)
    cnt = np.zeros(8, np.uint32)
    assert lib.ctd_reset(h, 257, 1, 0, 0) == 1                      # n > capacity -> CTD_EARG
    assert lib.ctd_reset(h, 8, 1, 0, 7) == 1                        # unknown ruleset
    assert lib.ctd_reset(None, 8, 1, 0, 0) == 1                     # null handle
    assert lib.ctd_enumerate(h, 8, None, cnt.ctypes.data, 8) == 1   # null output
    assert lib.ctd_enumerate(h, 8, buf.ctypes.data, cnt.ctypes.data, 0) == 1
    assert lib.ctd_make_roots(h, 8, 1, 0, 0, 5, 2, 0, None) == 1    # back_hi < back_lo
    assert lib.ctd_make_roots(h, 8, 1, 0, 0, 1, 2, 7, None) == 1    # unknown flavour
    st = ctypes.c_void_p()
    assert lib.ctd_create(0, 0, ctypes.byref(st)) in (0, 1)         # zero capacity: accepted or refused, never a crash
    if st.value:
        lib.ctd_destroy(st)
    engine.reset(8, seed=3)                                          # the handle is still usable after refused calls
    assert engine.enumerate(8)[1].min() >= 1


def test_small_stride_reports_ecap_and_leaves_slots_untouched(engine):
    engine.reset(64, seed=11, first_gid=0, ruleset=1)
    for _ in range(40):
        opts, counts = engine.enumerate(64)
        engine.step(opts[:, 0].copy())
    before = engine.store_states(64)
    lib, h = engine._lib, engine._h
    opts = np.zeros((64, 1), np.uint64)
    counts = np.zeros(64, np.uint32)
    st = lib.ctd_enumerate(h, 64, opts.ctypes.data, counts.ctypes.data, 1)
    full, fc = engine.enumerate(64)
    assert (st == 3) == bool((fc > 1).any())                        # CTD_ECAP exactly when a list did not fit
    assert np.array_equal(counts, fc)                                # the true counts are reported either way
    assert np.array_equal(opts[:, 0], full[:, 0])                    # ... and the first `stride` options
    assert np.array_equal(engine.store_states(64), before)           # enumeration changed nothing


def test_step_cap_sets_maxsteps_flag(engine):
    out = engine.playout(32, seed=5, first_gid=0, max_steps=50)
    assert out["stats"]["errors"] == 32 and (out["steps"] == 50).all() and (out["winner"] == -1).all()
    engine.reset(32, seed=5, first_gid=0)
    w, s = engine.playout_slots(32, max_steps=10)
    assert (s == 10).all()
    rec = engine.store_states(32)
    assert (rec[:, 228] & 16).all()                                  # CTD_ERR_MAXSTEPS in ctd_state.err


def test_terminal_states_enumerate_nothing_and_terminal_roots_are_reported(engine):
    from oracle import citadels_oracle as O
    engine.reset(16, seed=21, first_gid=0)
    engine.playout_slots(16)
    rec = engine.store_states(16)
    assert (rec[:, 217] & 2).all()                                   # gflags: terminal
    opts, counts = engine.enumerate(16)
    assert (counts == 0).all()
    # a finished game as a CFR root: status 1, the reference's run_mccfr raises ValueError there
    engine.make_roots(16, seed=21, first_gid=0, back_lo=0, back_hi=0)
    roots, knows, used, gids = engine.store_roots(16)
    term = (roots[:, 217] & 2) != 0
    out = engine.mccfr(16, iterations=20, seed=21)
    assert ((out["results"]["status"] == 1) == term).all()
    assert O.Game.unpack(bytes(rec[0])).terminal


def test_ragged_batch_mixed_rulesets_and_lengths(engine):
    """Slots filled from three rulesets and different points of their games step together."""
    parts = []
    rng = np.random.default_rng(9)
    for ruleset, n, k in ((0, 20, 5), (1, 20, 70), (2, 24, 150)):
        engine.reset(n, seed=31 + ruleset, first_gid=1000, ruleset=ruleset)
        for _ in range(k):
            opts, counts = engine.enumerate(n, stride=256)
            pick = (rng.random(n) * counts).astype(np.int64)
            engine.step(opts[np.arange(n), pick].copy())
        parts.append(engine.store_states(n))
    mixed = np.concatenate(parts)
    engine.load_states(mixed)
    w, s = engine.playout_slots(64)
    assert (w >= 0).all() and (w < 6).all()
    fin = engine.store_states(64)
    assert (fin[:, 228] == 0).all() and (fin[:, 217] & 2).all()
    assert np.array_equal(fin[:, 227], mixed[:, 227])               # every game kept its ruleset


def test_wrong_ruleset_named_for_a_search_is_refused_not_run(engine):
    """The preset-specialised search kernels must not be handed roots of another ruleset: status 4 per tree, no crash."""
    engine.make_roots(8, seed=3, first_gid=100, ruleset=1, back_lo=0, back_hi=30)      # classic roots ...
    out = engine.mccfr(8, iterations=20, seed=3, ruleset=0)                             # ... searched "as preset"
    assert (out["results"]["status"] == 4).all() and (out["results"]["n_nodes"] == 0).all()
    out = engine.mccfr(8, iterations=20, seed=3, ruleset=1)                             # the handle is still fine
    assert (out["results"]["status"] <= 1).all() and (out["results"]["n_nodes"] > 0).any()

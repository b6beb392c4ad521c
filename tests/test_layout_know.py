"""The knowledge block's entry formats: current (64 entries of 4 bytes) and round 1 (32 of 8 bytes, what the CFR fixtures hold)."""
import os
import numpy as np
import pytest

from citadels_self_play_b200 import layout as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_entry_word_round_trip():
    for pid in range(-1, 6):
        for conf, flags, n, off in ((5, 0, 1, 0), (1, 3, 64, 256), (3, 2, 17, 101)):
            f = L.hk_unpack(L.hk_pack(pid, conf, flags, n, off))
            assert (f["pid"], f["conf"], f["flags"], f["n"], f["off"]) == (pid, conf, flags, n, off)


@pytest.mark.parametrize("name", ["mccfr_preset.npz", "mccfr_random.npz", "live_choice_preset.npz"])
def test_v1_blocks_convert_both_ways(name):
    with np.load(os.path.join(GOLDEN, name)) as f:
        v1 = f["knows"]
    cur = L.know_from_v1(v1)
    assert cur.shape == v1.shape and cur.dtype == np.uint8
    assert np.array_equal(L.know_to_v1(cur), v1)
    k = cur.view(L.KNOW_DTYPE).reshape(-1)
    assert (k["n_hk"] <= 32).all() and int(k["n_hk"].max()) > 0
    old = v1.reshape(-1, 592)
    for b in range(len(k)):
        for i in range(int(k["n_hk"][b])):
            f = L.hk_unpack(k["hk"][b][i])
            o = [int(x) for x in old[b, 16 + 8 * i:24 + 8 * i]]
            pid = o[0] - 256 if o[0] >= 128 else o[0]
            assert (f["pid"], f["conf"], f["flags"], f["n"], f["off"]) == (pid, o[1], o[2], o[3], o[4] + 256 * o[5])
        assert not k["hk"][b][int(k["n_hk"][b]):].any()


def test_blocks_beyond_32_entries_have_no_v1_form():
    blk = np.zeros(592, np.uint8)
    blk[2] = 33
    with pytest.raises(ValueError):
        L.know_to_v1(blk)

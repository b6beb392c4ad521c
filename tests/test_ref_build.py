"""oracle/_ref -- the REAL reference byte-compiled by oracle/build_ref.py, the CPU arm of bench.py.  Where /root/reference is
present (the build container) every code object is checked against a fresh compile of the source where it lies; everywhere
the three loops bench.py times are run once (oracle/ref_loop.py)."""
import marshal
import os
import py_compile
import tempfile
import pytest

from oracle import build_ref, ref_loop


def test_ref_build_is_the_reference_compiled_from_where_it_lies():
    if not build_ref.available():
        pytest.skip("no /root/reference here (GPU box): oracle/_ref travelled with the repo")
    assert build_ref.build()
    with tempfile.TemporaryDirectory() as tmp:
        for m in build_ref.MODULES:
            fresh = py_compile.compile(os.path.join(build_ref.REFERENCE, m), cfile=os.path.join(tmp, "x.pyc"), dfile=m, doraise=True)
            a = marshal.loads(open(fresh, "rb").read()[16:])
            b = marshal.loads(open(os.path.join(build_ref.OUT, m + "c.bin"), "rb").read()[16:])
            assert a.co_code == b.co_code and a.co_consts == b.co_consts and a.co_names == b.co_names, m


def test_reference_loops_run_from_the_build():
    if not ref_loop.available():
        pytest.skip("oracle/_ref not built")
    steps, games, secs, wins = ref_loop.playouts(3, seed=7)
    assert games == 3 and sum(wins) == 3 and 600 < steps < 2400
    it, n, _ = ref_loop.pure_mccfr(1, seed=3, iterations=40)
    assert (it, n) == (40, 1)
    it, n, _ = ref_loop.deep_mccfr(1, seed=3, iterations=40)
    assert (it, n) == (40, 1)
    import run_utils
    assert run_utils.__file__.endswith(os.path.join("oracle", "_ref", "run_utils.pyc.bin"))

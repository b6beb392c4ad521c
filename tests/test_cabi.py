"""The C-ABI library loads and exports every symbol include/citadels_b200.h declares (no compute, no GPU)."""
import os
import re
import ctypes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "citadels_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctd_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from citadels_self_play_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))


def test_state_layout_matches_header():
    from citadels_self_play_b200.layout import STATE_DTYPE
    src = open(os.path.join(ROOT, "include", "citadels_b200.h")).read()
    body = src[src.index("typedef struct ctd_state {"):src.index("} ctd_state;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:u?int\d+_t)\s+(\w+)(?:\[(\d+)\])?;", body)
    assert [f[0] for f in fields] == list(STATE_DTYPE.names)
    for name, cnt in fields:
        shape = STATE_DTYPE[name].shape
        assert (shape[0] if shape else 1) == (int(cnt) if cnt else 1), name


def test_record_layouts_match_the_library():
    """The numpy mirrors of every record that crosses the C ABI have the sizes the library was compiled with."""
    import __graft_entry__
    __graft_entry__.build()
    from citadels_self_play_b200 import _lib, layout
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.ctd_sizeof.restype = ctypes.c_uint32
    want = [layout.STATE_DTYPE.itemsize, layout.MCCFR_RESULT_DTYPE.itemsize, layout.TARGET_META_DTYPE.itemsize, layout.KNOW_DTYPE.itemsize,
            layout.TREE_HDR_DTYPE.itemsize, layout.NODE_DTYPE.itemsize, layout.CHILD_DTYPE.itemsize, ctypes.sizeof(_lib.PlayoutStats)]
    assert [lib.ctd_sizeof(i) for i in range(8)] == want


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from citadels_self_play_b200 import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(capacity=8)

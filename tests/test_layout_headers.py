"""csrc/ctd_layout_*.h only rename device functions (to fix their placement order): every renamed function must exist in the rules
code, new names must be distinct and equally long (ptxas orders functions by mangled name = length, then text)."""
import os
import re
import glob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "citadels_self_play_b200", "csrc")


def test_layout_headers_rename_existing_functions_only():
    code = "".join(open(f).read() for f in glob.glob(os.path.join(CSRC, "*.cuh")))
    headers = sorted(glob.glob(os.path.join(CSRC, "ctd_layout_*.h")))
    assert headers
    for h in headers:
        pairs = re.findall(r"^#define\s+(ctd_\w+)\s+(ctd_h\d\d_\w+)\s*$", open(h).read(), flags=re.M)
        assert len(pairs) >= 10, h
        olds, news = [p[0] for p in pairs], [p[1] for p in pairs]
        assert len(set(olds)) == len(olds) and len(set(news)) == len(news), h
        assert len(set(len(n) for n in news)) == 1, h                     # one length class: one contiguous block
        assert [int(n[5:7]) for n in news] == list(range(len(news))), h   # the order prefix is the order
        for o in olds:
            assert re.search(r"\b%s\s*\(" % re.escape(o), code), (h, o)
        others = [l for l in open(h).read().splitlines() if l.strip() and not l.startswith(("//", "#define", "#pragma once"))]
        assert not others, (h, others)

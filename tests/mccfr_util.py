"""Shared helpers for the MCCFR parity tests: the golden tree fixtures (real reference) and tree comparison."""
import os
import zlib
import numpy as np

from tests.golden_util import GOLDEN, visible


class MccfrGolden:
    def __init__(self, name):
        with np.load(os.path.join(GOLDEN, name)) as f:
            z = {k: f[k] for k in f.files}      # decompress once (NpzFile re-reads on every access)
        # the fixtures hold the reference's knowledge in the round-1 entry format (32 x 8 bytes): root blocks are converted here,
        # the per-node checksums are compared over the converted-back form (know_v1_crc below)
        from citadels_self_play_b200.layout import know_from_v1
        z["knows"] = know_from_v1(z["knows"])
        self.z = z
        self.seed = int(z["seed"])
        self.ruleset = int(z["ruleset"])
        self.iterations = int(z["iterations"])
        self.back_hi = int(z["back_hi"])
        self.gids = z["gids"]
        self.n = len(self.gids)
        self.child_off = np.concatenate([[0], np.cumsum(z["nchild"])])
        self.arr_off = np.concatenate([[0], np.cumsum(z["narr"])])

    def nodes(self, r):
        """Yield per-node dicts of root r in the reference's DFS pre-order."""
        z = self.z
        for i in range(int(z["node_off"][r]), int(z["node_off"][r + 1])):
            a0, a1 = int(self.arr_off[i]), int(self.arr_off[i + 1])
            yield dict(nchild=int(z["nchild"][i]), desc=z["desc"][int(self.child_off[i]):int(self.child_off[i + 1])],
                       V=z["V"][i], P=z["P"][i], R=z["R"][a0:a1], S=z["S"][a0:a1], C=z["C"][a0:a1],
                       game_crc=int(z["game_crc"][i]), know_crc=int(z["know_crc"][i]))


def know_v1_crc(block):
    """checksum of a knowledge block as the fixtures took it (round-1 entry format); a block with more than 32 entries has no such
    form and is summed as it is -- such blocks only ever meet blocks of the same kind (engine vs oracle)"""
    from citadels_self_play_b200.layout import know_to_v1
    raw = np.frombuffer(bytes(block), dtype=np.uint8)
    if int(raw[2]) > 32:
        return zlib.crc32(raw.tobytes()) ^ 0x5A5A5A5A
    return zlib.crc32(know_to_v1(raw).tobytes())


def oracle_preorder(node):
    """Oracle tree -> the same per-node dicts."""
    for n in node.walk():
        yield dict(nchild=len(n.children), desc=np.asarray([c[0] for c in n.children], dtype=np.uint64), V=n.V, P=n.P,
                   R=np.asarray(n.R, dtype=float).ravel(), S=np.asarray(n.s, dtype=float).ravel(),
                   C=np.asarray(n.C, dtype=float).ravel(), game_crc=zlib.crc32(visible(n.game.pack())),
                   know_crc=know_v1_crc(n.game.pack_know(n.orig)))


def tree_preorder(tv):
    """Engine tree block (layout.TreeView) -> the same per-node dicts."""
    from citadels_self_play_b200.layout import NODE_DTYPE
    raw = tv.nodes.view(np.uint8).reshape(len(tv.nodes), NODE_DTYPE.itemsize)
    g0, k0 = NODE_DTYPE.fields["game"][1], NODE_DTYPE.fields["know"][1]
    child_desc, child_node = tv.children["desc"], tv.children["node"]
    nch, coff = tv.nodes["n_children"], tv.nodes["child_off"]
    V, P = tv.nodes["V"], tv.nodes["P"]
    stack = [0]
    while stack:
        i = stack.pop()
        k, o = int(nch[i]), int(coff[i])
        R, S, C = tv.arrays(i)
        yield dict(nchild=k, desc=child_desc[o:o + k], V=V[i], P=P[i],
                   R=np.asarray(R).ravel(), S=np.asarray(S).ravel(), C=np.asarray(C).ravel(),
                   game_crc=zlib.crc32(visible(raw[i, g0:g0 + 256].tobytes())), know_crc=know_v1_crc(raw[i, k0:k0 + 592].tobytes()))
        stack.extend(int(x) for x in child_node[o:o + k][::-1])


def assert_same_tree(a, b, what, rtol=1e-9, atol=1e-12, norm_rtol=None):
    """Integers exact; regrets / strategies / values to rtol elementwise (north_star: 1e-5 relative; pure MCCFR holds
    1e-9).  norm_rtol, if given, replaces the elementwise test by max|a-b| <= norm_rtol * max(1, max|a|) per array:
    with a value model the leaf values carry fp32 rounding (~1e-6, different summation order than torch's CPU GEMV)
    and regrets are accumulated differences of such values, so entries near zero have no meaningful relative error."""
    a, b = list(a), list(b)
    assert len(a) == len(b), (what, "node count", len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        assert x["nchild"] == y["nchild"], (what, i, "children")
        m = (np.asarray(x["desc"]) != 0) & (np.asarray(y["desc"]) != 0)   # 0 = descriptor not recorded in the fixture
        assert np.array_equal(np.asarray(x["desc"])[m], np.asarray(y["desc"])[m]), (what, i, "options")
        assert x["game_crc"] == y["game_crc"], (what, i, "game record")
        assert x["know_crc"] == y["know_crc"], (what, i, "knowledge block")
        for k in ("V", "P", "R", "S", "C"):
            u, v = np.asarray(x[k], dtype=float), np.asarray(y[k], dtype=float)
            assert u.shape == v.shape, (what, i, k, u.shape, v.shape)
            if norm_rtol is not None:
                if u.size and not np.abs(u - v).max() <= norm_rtol * max(1.0, np.abs(u).max()):
                    raise AssertionError((what, "node", i, k, "max |diff| %.3e vs scale %.3e" % (np.abs(u - v).max(), np.abs(u).max())))
            elif not np.allclose(u, v, rtol=rtol, atol=atol, equal_nan=True):
                j = int(np.nanargmax(np.abs(u - v)))
                raise AssertionError((what, "node", i, k, "max |diff| %.3e at %d: %r vs %r" % (abs(u[j] - v[j]), j, u[j], v[j])))

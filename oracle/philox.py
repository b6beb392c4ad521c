"""Philox4x32-10 counter-based RNG and the chance sources built on it.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.

The reference draws all chance from CPython's Mersenne Twister
(`random.shuffle`, game/deck.py:73; `random.shuffle`, game/game.py:152).  The
engine replaces that with a counter-based stream so a game is a pure function
of (seed, game id).  This file is the CPU statement of that stream; the CUDA
side (csrc/ctd_rng.cuh) is written independently against the same definition:

  key     = (seed_lo, seed_hi)
  counter = (block, stream, gid_lo, gid_hi),  draw i -> block = i >> 2, word = i & 3
  randbelow(n) = (u32 * n) >> 32
  perm(n): Fisher-Yates from the top, exactly the loop shape of CPython's
           random.shuffle (Lib/random.py): for i = n-1 .. 1: j = randbelow(i+1); swap(i, j)
           n <= 1 consumes nothing.
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


class PhiloxChance:
    """Chance source keyed by (seed, game id, stream)."""

    def __init__(self, seed, gid, stream=0):
        self.k0 = seed & MASK
        self.k1 = (seed >> 32) & MASK
        self.g0 = gid & MASK
        self.g1 = (gid >> 32) & MASK
        self.stream = stream & MASK
        self.i = 0
        self._blk = -1
        self._buf = None

    def u32(self):
        blk = self.i >> 2
        if blk != self._blk:
            self._buf = philox4x32_10(blk & MASK, self.stream, self.g0, self.g1, self.k0, self.k1)
            self._blk = blk
        v = self._buf[self.i & 3]
        self.i += 1
        return v

    def randbelow(self, n):
        return (self.u32() * n) >> 32

    randbelow_game = randbelow   # draws that belong to the game itself (role variants, crown seat): same stream

    def uniform(self):
        """[0,1) with 32 bits, as a float64 (CFR: HandKnowledge use test)."""
        return self.u32() / 4294967296.0

    def perm(self, n):
        idx = list(range(n))
        for i in range(n - 1, 0, -1):
            j = self.randbelow(i + 1)
            idx[i], idx[j] = idx[j], idx[i]
        return idx


class TapeChance:
    """Chance source that replays recorded outcomes (bit-exact trace replay).

    `tape` is a flat list of small ints: each perm(n>1) consumes n entries (the
    permutation as source indices: new[k] = old[perm[k]]).  `choices` holds the
    recorded option indices, one per randbelow() call.
    """

    def __init__(self, tape, choices=()):
        self.tape = list(tape)
        self.pos = 0
        self.choices = list(choices)
        self.cpos = 0

    def perm(self, n):
        if n <= 1:
            return list(range(n))
        p = self.tape[self.pos:self.pos + n]
        self.pos += n
        assert sorted(p) == list(range(n)), "tape out of sync"
        return p

    def randbelow(self, n):
        v = self.choices[self.cpos]
        self.cpos += 1
        assert v < n
        return v

    def randbelow_game(self, n):
        """A chance draw of the game itself (not the caller's option choice): one tape entry."""
        v = self.tape[self.pos]
        self.pos += 1
        assert v < n
        return v


class RecordingChance:
    """Wraps another source and records the tape TapeChance would need."""

    def __init__(self, inner):
        self.inner = inner
        self.tape = []
        self.choices = []

    def perm(self, n):
        p = self.inner.perm(n)
        if n > 1:
            self.tape.extend(p)
        return p

    def randbelow(self, n):
        v = self.inner.randbelow(n)
        self.choices.append(v)
        return v

    def randbelow_game(self, n):
        v = self.inner.randbelow(n)
        self.tape.append(v)
        return v

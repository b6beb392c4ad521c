"""Generate the committed golden fixtures from the REAL reference (container-only).

    python -m tests.golden.gen_golden            # writes tests/golden/*.npz

For each game the unmodified reference (/root/reference/game) is played to terminal by the loop of
run_utils.py:37-41, with its shuffles and the option choice driven by PhiloxChance(SEED, gid).
Recorded per step: number of legal options, chosen index, crc32 of the 228 reference-visible bytes of
the packed state (tests/golden/ref_harness.ref_pack) and crc32 of the little-endian descriptor list
(ref_harness.ref_descriptors).  Recorded per game: the chance tape (every shuffle's permutation), the
terminal record, steps.  `*_full.npz` additionally keeps every state and descriptor of a few games.

`ref_outcomes_preset.npz` is a plain sample of reference outcomes under its own Mersenne-Twister
(random.seed(gid)), used for the distribution gate (SURVEY.md 8(d) parity 2).
"""
import os
import sys
import zlib
import numpy as np
from multiprocessing import Pool

SEED = 0xC17ADE15
HERE = os.path.dirname(os.path.abspath(__file__))


def _one(args):
    gid, ruleset, full = args
    from oracle.philox import PhiloxChance, RecordingChance
    from tests.golden import ref_harness as H
    ch = RecordingChance(PhiloxChance(SEED, gid))
    g = H.new_ref_game(ch, ruleset)
    nopt, chosen, hs, ho = [], [], [], []
    states, descs = [], []
    while True:
        H.set_chance(ch)
        opts = g.get_options_from_state()
        d = H.ref_descriptors(opts)
        rec = H.ref_pack(g, ruleset)
        db = np.asarray(d, dtype="<u8").tobytes()
        nopt.append(len(d))
        hs.append(zlib.crc32(rec[:228]))
        ho.append(zlib.crc32(db))
        if full:
            states.append(rec)
            descs.append(d)
        i = ch.randbelow(len(d))
        chosen.append(i)
        if opts[i].carry_out(g):
            break
    final = H.ref_pack(g, ruleset)
    return dict(gid=gid, tape=ch.tape, nopt=nopt, chosen=chosen, hs=hs, ho=ho, final=final, states=states,
                descs=descs)


def _flat(lists, dtype):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in lists])
    flat = np.concatenate([np.asarray(x, dtype=dtype) for x in lists]) if lists else np.zeros(0, dtype)
    return flat, off


def gen_traces(name, ruleset, gids, full=False, procs=8):
    with Pool(procs) as pool:
        res = pool.map(_one, [(g, ruleset, full) for g in gids], chunksize=4)
    tape, tape_off = _flat([r["tape"] for r in res], np.uint8)
    nopt, step_off = _flat([r["nopt"] for r in res], np.uint16)
    chosen, _ = _flat([r["chosen"] for r in res], np.uint16)
    hs, _ = _flat([r["hs"] for r in res], np.uint32)
    ho, _ = _flat([r["ho"] for r in res], np.uint32)
    out = dict(seed=np.uint64(SEED), ruleset=np.int32(ruleset), gids=np.asarray(gids, dtype=np.uint64),
               tape=tape, tape_off=tape_off, step_off=step_off, nopt=nopt, chosen=chosen, state_crc=hs,
               opts_crc=ho, final=np.frombuffer(b"".join(r["final"] for r in res), dtype=np.uint8).reshape(-1, 256))
    if full:
        out["states"] = np.frombuffer(b"".join(b"".join(r["states"]) for r in res), dtype=np.uint8).reshape(-1, 256)
        dl = [d for r in res for d in r["descs"]]
        out["descs"], out["desc_off"] = _flat(dl, np.uint64)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(name, "games", len(gids), "steps", len(nopt), "bytes", os.path.getsize(path))


def _mt_game(gid):
    import random
    from tests.golden import ref_harness as H
    gg = H.load_reference()

    class MT:
        def perm(self, n):
            idx = list(range(n))
            random.Random.shuffle(random._inst, idx)
            return idx
    H.set_chance(MT())
    random.seed(gid)
    g = gg.Game(preset=True)
    g.setup_round()
    steps = 0
    while True:
        opts = g.get_options_from_state()
        steps += 1
        if opts[random.randrange(len(opts))].carry_out(g):
            break
    return [int(g.rewards.argmax())] + [int(x) for x in g.points] + [steps]


def gen_outcomes(name, n, procs=8):
    with Pool(procs) as pool:
        res = pool.map(_mt_game, range(n), chunksize=16)
    a = np.asarray(res, dtype=np.int16)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, winner=a[:, 0].astype(np.int8), points=a[:, 1:7].astype(np.int8), steps=a[:, 7])
    print(name, "games", n, "mean steps", a[:, 7].mean(), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "traces"):
        gen_traces("preset_traces.npz", 0, list(range(1000)))
        gen_traces("classic_traces.npz", 1, list(range(100000, 100300)))
        gen_traces("preset_full.npz", 0, list(range(2000, 2006)), full=True)
        gen_traces("classic_full.npz", 1, list(range(102000, 102004)), full=True)
    if what in ("all", "outcomes"):
        gen_outcomes("ref_outcomes_preset.npz", 20000)

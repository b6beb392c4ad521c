"""Developer tool (CPU, host build of the kernels' rules code): which CFR roots end with status > 1?
   python tools/scan_status_cpu.py FIRST_GID N ITERS BACK_LO BACK_HI [RULESET] [FLAVOUR]
Replays exactly what ctd_make_roots + ctd_mccfr do on the device for gids [FIRST_GID, FIRST_GID + N)."""
import ctypes, os, sys
import numpy as np
from multiprocessing import Pool
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xC17ADE15


def work(args):
    gid0, n, iters, lo, hi, rs, fl = args
    lib = ctypes.CDLL(os.environ.get("HS_LIB", os.path.join(ROOT, "tests", "hostsim", "libctd_hostsim.so")))
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    lib.hs_make_root.argtypes = [u64, u64, i32, u32, u32, i32, vp, vp, vp, vp]
    lib.hs_mccfr.argtypes = [vp, vp, vp, u64, u64, u32, vp, u64, vp, u64, vp]
    arena = np.zeros(int(os.environ.get("HS_ARENA_MB", "1024")) << 20, np.uint8)
    root, know, used, step = np.zeros(256, np.uint8), np.zeros(592, np.uint8), np.zeros(76, np.uint8), np.zeros(1, np.uint32)
    nb = ctypes.c_uint64()
    out = []
    for g in range(gid0, gid0 + n):
        lib.hs_make_root(SEED, g, rs, lo, hi, fl, root.ctypes.data, know.ctypes.data, used.ctypes.data, step.ctypes.data)
        st = lib.hs_mccfr(root.ctypes.data, know.ctypes.data, used.ctypes.data, SEED, g, iters, arena.ctypes.data, arena.nbytes,
                          None, 0, ctypes.byref(nb))
        out.append((g, st, int(nb.value), int(step[0])))
    return out


if __name__ == "__main__":
    g0, n, iters, lo, hi = [int(x) for x in sys.argv[1:6]]
    rs = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    fl = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    P = os.cpu_count()
    per = (n + 8 * P - 1) // (8 * P)
    jobs = [(g0 + i, min(per, g0 + n - (g0 + i)), iters, lo, hi, rs, fl) for i in range(0, n, per)]
    with Pool(P) as pool:
        res = [r for chunk in pool.imap_unordered(work, jobs) for r in chunk]
    bad = sorted(r for r in res if r[1] > 1)
    print("roots", len(res), "largest export block", max(r[2] for r in res), "mean", sum(r[2] for r in res) // len(res),
          "status>1:", len(bad), "by status", {s: sum(1 for r in bad if r[1] == s) for s in sorted(set(r[1] for r in bad))})
    for r in bad[:40]:
        print("gid %d status %d export bytes %d root_step %d" % r)

#!/usr/bin/env python
"""bench.py -- env steps/sec of batched 6-player random playouts (BASELINE.json metric, config[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: G independent preset games per GPU, dealt on the
device from (seed, global game id) and played uniformly at random to terminal by the fused playout kernel.
Weak scaling: every rank plays G games (rank r owns global ids [r*G*K', ...)), no collective on the step path;
one all-reduce of the outcome histogram after the timed region.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_STEP = 512          # read + write of the 256 B packed playout record per env step
SEED = 0xC17ADE15
METRIC = "env steps/sec (batched 6p random playouts)"


def _traffic(games_per_launch):
    """DRAM bytes per launch of the playout kernel from the committed ncu capture, scaled by games per launch."""
    p = os.path.join(ROOT, "profiles", "r01_playout_traffic.json")
    try:
        t = json.load(open(p))
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) * (games_per_launch / t["games"])
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    seed, gid0, n, ruleset = args
    from oracle import citadels_oracle as O
    steps = 0
    t0 = time.perf_counter()
    for i in range(n):
        steps += O.playout(seed, gid0 + i, ruleset)[2]
    return steps, time.perf_counter() - t0


def cpu_playouts(games_per_core, cores, gid0=0, ruleset=0):
    """The reference's random-playout loop (run_utils.py:37-41) as restated in oracle/ (a Python port of a
    Python reference), on `cores` processes.  Returns (env_steps, wall_seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(SEED, gid0 + c * games_per_core, games_per_core, ruleset) for c in range(cores)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def _cpu_mccfr_worker(args):
    seed, gid0, n, iters = args
    from oracle import mccfr_oracle as M
    from oracle.philox import PhiloxChance
    its = 0
    t0 = time.perf_counter()
    for i in range(n):
        g, _ = M.make_root(seed, gid0 + i, 0, 0, 20)
        if g.terminal:
            continue
        g.chance = PhiloxChance(seed, gid0 + i, stream=1)
        M.Node(g, g.player).cfr_train(iters)
        its += iters
    return its, time.perf_counter() - t0


def cpu_mccfr(roots_per_core, cores, iters=200):
    """run_mccfr(game, max_iterations=200) on near-terminal roots (oracle port), all host cores."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_mccfr_worker, [(SEED, 10_000 + c * roots_per_core, roots_per_core, iters) for c in range(cores)])
    return sum(r[0] for r in res), time.perf_counter() - t0


def run_reference_arm(args, rank, world, emit):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = max(1, args.ref_games_per_core)
    for _ in range(args.warmup):
        cpu_playouts(1, cores, ruleset=args.ruleset)
    tot_steps, tot_wall = 0, 0.0
    for k in range(args.steps):
        s, w = cpu_playouts(per_core, cores, gid0=1000 + k * cores * per_core, ruleset=args.ruleset)
        tot_steps += s
        tot_wall += w
    v = tot_steps / tot_wall
    sample = "%d games per step (%d per core x %d cores) of the %s random playout, oracle port" % (
        per_core * cores, per_core, cores, ["preset", "classic", "random-ruleset"][args.ruleset])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "env steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s 6p random playouts to terminal (BASELINE configs[1]), bounded CPU sample" % ["preset", "classic", "random-ruleset"][args.ruleset],
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "env steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--games", type=int, default=1 << 20, help="games per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-games-per-core", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ruleset", type=int, default=0)
    ap.add_argument("--no-mccfr", action="store_true", help="skip the secondary MCCFR measurement")
    ap.add_argument("--mccfr-roots", type=int, default=4096)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the one JSON line and nothing else: library chatter written straight to fd 1 (NCCL prints its
    # version there under NCCL_DEBUG=VERSION) goes to stderr until the line is printed
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from citadels_self_play_b200 import Engine
    eng = Engine(capacity=1024, device=local)
    G = args.games
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    from citadels_self_play_b200 import sharding

    def gid0(step):  # disjoint global ids per (step, rank)
        return sharding.first_gid(step, rank, world, G)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (untimed)
    for w in range(args.warmup):
        flush.zero_()
        torch.cuda.synchronize()
        eng.playout(G, seed=SEED, first_gid=gid0(1000 + w), ruleset=args.ruleset, outputs=False)

    # ---- device-resident timing: K steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    kernel_ms, env_steps, errors, wins = 0.0, 0, 0, [0] * 6
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        st = eng.playout(G, seed=SEED, first_gid=gid0(k), ruleset=args.ruleset, outputs=False)["stats"]
        kernel_ms += st["kernel_ms"]
        env_steps += st["steps"]
        errors += st["errors"]
        wins = [a + b for a, b in zip(wins, st["wins"])]
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launches - launches0

    # ---- end-to-end through the public API with HOST buffers: every step copies G packed game records (256 B each) from
    # pinned host memory into the engine's slots (ctd_load_states), plays them to terminal (ctd_playout_slots) and reads
    # winner + step count of every game back to the host.  The records are made beforehand, untimed (two alternating sets).
    eng_h = Engine(capacity=G, device=local)
    host_sets = []
    for j in range(min(2, max(args.steps, 1))):
        buf = torch.empty((G, 256), dtype=torch.uint8, pin_memory=True).numpy()
        eng_h.reset(G, seed=SEED, first_gid=gid0(2000 + j), ruleset=args.ruleset)
        eng_h.store_states(G, out=buf)
        host_sets.append(buf)
    eng_h.load_states(host_sets[0])
    eng_h.playout_slots(G)                                   # warm-up of this path
    launches_h0 = eng_h.launches
    barrier()
    t1 = time.perf_counter()
    e2e_steps = 0
    for k in range(args.steps):
        eng_h.load_states(host_sets[k % len(host_sets)])      # H2D 256*G bytes
        w_host, s_host = eng_h.playout_slots(G)              # D2H 3*G bytes (int8 winner, uint16 steps per game)
        e2e_steps += int(s_host.sum(dtype="int64"))
    barrier()
    e2e_wall = time.perf_counter() - t1
    e2e_launches = eng_h.launches - launches_h0
    assert (w_host >= 0).all()
    eng_h.close()
    del host_sets
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary metric: MCCFR iterations/s (BASELINE configs[2] pure, configs[3] deep), same roots on every rank ----
    mccfr = None
    if not args.no_mccfr:
        R, IT = args.mccfr_roots, 200
        eng2 = Engine(capacity=R, device=local)
        eng2.make_roots(R, seed=SEED, first_gid=sharding.first_gid(0, rank, world, R), back_lo=0, back_hi=20)
        pure_it = pure_ms = deep_it = deep_ms = 0
        eng2.mccfr(R, iterations=IT, seed=SEED)
        for _ in range(2):
            o = eng2.mccfr(R, iterations=IT, seed=SEED)
            pure_it += int(o["results"]["iterations"].sum())
            pure_ms += o["kernel_ms"]
        from citadels_self_play_b200.value_model import ValueOnlyNN
        torch.manual_seed(0)
        eng2.set_value_model(ValueOnlyNN(418, 512).eval())
        eng2.mccfr_pred(R, iterations=IT, max_depth=10, seed=SEED)
        for _ in range(2):
            o = eng2.mccfr_pred(R, iterations=IT, max_depth=10, seed=SEED)
            deep_it += int(o["results"]["iterations"].sum())
            deep_ms += o["kernel_ms"]
        bad = int((o["results"]["status"] > 1).sum())
        eng2.close()
        # BASELINE configs[4]: training-data generation, 2000 iterations, roots stepped back 1..100, nodes with >= 200 backprops
        R5 = max(64, args.mccfr_roots // 4)
        eng3 = Engine(capacity=R5, device=local)
        eng3.make_roots(R5, seed=SEED, first_gid=sharding.first_gid(1, rank, world, R5), back_lo=1, back_hi=100)
        eng3.mccfr(R5, iterations=2000, seed=SEED)
        o5 = eng3.mccfr(R5, iterations=2000, seed=SEED)
        t5 = time.perf_counter()
        tg = eng3.mccfr_targets(R5, iterations=2000, seed=SEED, threshold=200.0)
        t5 = time.perf_counter() - t5
        gen_it, gen_ms, gen_targets = int(o5["results"]["iterations"].sum()), o5["kernel_ms"], len(tg["meta"])
        bad += int((o5["results"]["status"] > 1).sum())
        eng3.close()
        mccfr = [pure_it, pure_ms, deep_it, deep_ms, bad, gen_it, gen_ms, gen_targets, t5 * 1e3]

    # ---- the only collective: outcome statistics, after the timed region ----
    if mccfr is not None:
        mt = torch.tensor([mccfr[1], mccfr[3], mccfr[6], mccfr[8]], dtype=torch.float64, device="cuda")
        mi = torch.tensor([mccfr[0], mccfr[2], mccfr[4], mccfr[5], mccfr[7]], dtype=torch.int64, device="cuda")
        if world > 1:   # the data-gen collective: per-rank counts summed (targets themselves stay on their rank)
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
            dist.all_reduce(mi, op=dist.ReduceOp.SUM)
        mccfr = [int(mi[0]), float(mt[0]), int(mi[1]), float(mt[1]), int(mi[2]), int(mi[3]), float(mt[2]), int(mi[4]), float(mt[3])]
    t = torch.tensor([wall, e2e_wall, kernel_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([env_steps, e2e_steps, errors, launches] + wins, dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    wall, e2e_wall, kernel_ms = [float(x) for x in t.tolist()]
    env_steps, e2e_steps, errors, launches = [int(x) for x in s.tolist()[:4]]
    wins = [int(x) for x in s.tolist()[4:]]

    if rank == 0:
        peak, peak_src = _peaks()
        value = env_steps / wall
        per_gpu_kernel = (env_steps / world) / (kernel_ms / 1e3)
        achieved = per_gpu_kernel * BYTES_PER_STEP / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "env steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "device_ms_per_step": kernel_ms / max(args.steps, 1),   # CUDA events on the engine's stream around the kernel, max over ranks
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%d %s 6-player games per GPU per step, dealt on device from Philox(seed, gid), "
                                   "uniform-random option to terminal (BASELINE configs[1])" % (G, ["preset", "classic", "random-ruleset"][args.ruleset]),
                       "ruleset": ["preset", "classic", "random"][args.ruleset], "games_per_gpu_per_step": G,
                       "l2": "no HBM-resident inputs (games are generated on device); 256 MiB flush between iterations",
                       "parallelism": "games sharded by global id, %d rank(s), no step-path collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(G), "traffic_source": "profiles/r01_playout_traffic.json (ncu, bytes per launch)",
                         "peak_source": peak_src,
                         "note": "algorithmic 512 B/env step (SURVEY 8(d)); the fused kernel keeps the game in shared "
                                 "memory for its ~420 steps (4 B of DRAM traffic per step), so the real limiter is the SM front end: instruction fetch (see profiles/README.md)",
                         "kernel_env_steps_per_s_per_gpu": per_gpu_kernel},
            "e2e": {"value": e2e_steps / e2e_wall, "unit": "env steps/s", "h2d_bytes_per_step": 256 * G * world,
                    "d2h_bytes_per_step": 3 * G * world,
                    "path": "Engine.load_states (pinned host records -> HBM) + Engine.playout_slots (winner, steps -> host)"},
            "gpu_launches": launches, "gpu_launches_e2e": e2e_launches, "clocks": clocks,
            "outcomes": {"games": G * args.steps * world, "errors": errors, "wins": wins},
        }
        if mccfr is not None:
            line["mccfr"] = {
                "unit": "MCCFR iterations/s (one iteration = one node-step, algorithms/deep_mccfr.py:194-204)",
                "roots_per_gpu": args.mccfr_roots, "iterations_per_root": 200, "root_step_back": "0..20",
                "pure_it_per_s": mccfr[0] / (mccfr[1] / 1e3), "deep_it_per_s": mccfr[2] / (mccfr[3] / 1e3),
                "deep_max_depth": 10, "deep_model": "ValueOnlyNN(418,512), torch.manual_seed(0) init",
                "roofline_frac_hbm_2048B_per_it": (mccfr[0] / world / (mccfr[1] / 1e3)) * 2048 / 1e9 / peak,
                "trees_with_error_status": mccfr[4],
                "datagen_2000it": {"roots_per_gpu": max(64, args.mccfr_roots // 4), "root_step_back": "1..100",
                                   "it_per_s": mccfr[5] / (mccfr[6] / 1e3), "targets": mccfr[7], "usefulness_threshold": 200,
                                   "targets_export_ms": mccfr[8]}}
            if not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                ci, cw = cpu_mccfr(12, cores)
                line["mccfr"]["cpu_baseline"] = {"value": ci / cw, "unit": "iterations/s", "cores": cores, "kind": "port",
                                                 "sample": "%d roots x 200 iterations, oracle port of run_mccfr" % (12 * cores)}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_core = 300   # ~5 s of work per core: pool start-up and imports amortised
            cs, cw = cpu_playouts(per_core, cores, ruleset=args.ruleset)
            line["cpu_baseline"] = {"value": cs / cw, "unit": "env steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d %s games (%d per core), oracle port of run_utils.py:37-41"
                                              % (per_core * cores, ["preset", "classic", "random-ruleset"][args.ruleset], per_core)}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Recipe for oracle/_ref: the REAL reference, byte-compiled from the sources where they lie.

    python -m oracle.build_ref            # needs /root/reference (the build container); writes oracle/_ref/**/*.pyc

TEST / BENCH INFRASTRUCTURE, like everything under oracle/.  The reference is pure Python, so "compiling it from its own
sources" is `py_compile`: every module of the hot path (game/*.py, algorithms/{deep_mccfr,models,train_utils,train}.py,
run_utils.py) is compiled straight from /root/reference into a code-object file under oracle/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box like the built .so files; no reference source is copied into the repo).  The files
are ordinary .pyc images under the suffix ".pyc.bin" (the snapshot that goes to the GPU box skips *.pyc like any cache file);
oracle/ref_loop.py imports them with a twenty-line loader.  The GPU box has the same interpreter (the image is the same).  bench.py's `--impl reference` arm and its
`cpu_baseline` legs time THIS code (kind "reference"); tests/test_ref_build.py checks it against the oracle's restatement.
Only tests/, __graft_entry__ and bench.py's CPU legs may import it (oracle/ref_loop.py is the one entry point)."""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("CTD_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ["run_utils.py"] + ["game/" + f for f in ("agent.py", "agent_functions.py", "config.py", "deck.py", "game.py",
                                                    "helper_classes.py", "option.py", "option_functions.py")] + \
          ["algorithms/" + f for f in ("deep_mccfr.py", "models.py", "train_utils.py", "train.py")]


def available():
    return os.path.isfile(os.path.join(REFERENCE, "run_utils.py"))


def built():
    return all(os.path.isfile(os.path.join(OUT, m + "c.bin")) for m in MODULES)


def build(force=False):
    """-> True when oracle/_ref is complete (built now or before)."""
    if not available():
        return built()
    for m in MODULES:
        src, dst = os.path.join(REFERENCE, m), os.path.join(OUT, m + "c.bin")
        if force or not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(src, cfile=dst, dfile=m, doraise=True, optimize=0)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write("py_compile of %s with %s\n" % (REFERENCE, sys.version.replace("\n", " ")))
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "complete" if ok else "NOT built (no %s here)" % REFERENCE)

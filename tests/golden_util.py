"""Loading helpers for the committed golden fixtures (tests/golden/*.npz, made by gen_golden.py)."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def visible(rec):
    """Reference-visible bytes of a packed record (include/citadels_b200.h): [0,228) + seer_mask, seven_n + seven[]."""
    rec = bytes(rec)
    return rec[:228] + rec[229:231] + rec[240:247]


class Traces:
    def __init__(self, name):
        with np.load(os.path.join(GOLDEN, name)) as f:
            z = {k: f[k] for k in f.files}
        self.seed = int(z["seed"])
        self.ruleset = int(z["ruleset"])
        self.gids = z["gids"].astype(np.uint64)
        self.tape, self.tape_off = z["tape"], z["tape_off"]
        self.step_off = z["step_off"]
        self.nopt, self.chosen = z["nopt"], z["chosen"]
        self.state_crc, self.opts_crc = z["state_crc"], z["opts_crc"]
        self.final = z["final"]
        self.states = z["states"] if "states" in z else None
        self.descs = z["descs"] if "descs" in z else None
        self.desc_off = z["desc_off"] if "desc_off" in z else None

    def __len__(self):
        return len(self.gids)

    def game_tape(self, g):
        return self.tape[self.tape_off[g]:self.tape_off[g + 1]]

    def game_steps(self, g):
        return slice(int(self.step_off[g]), int(self.step_off[g + 1]))

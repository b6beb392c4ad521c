"""B200-native Citadels rollout engine behind the reference's Game / Agent.get_options /
option.carry_out surface.  The compute path is hand-written sm_100a CUDA reached through the C ABI in
include/citadels_b200.h; importing this package does not need a GPU, using it does."""
from .engine import Engine, EngineError, RULESET_PRESET, RULESET_CLASSIC, DEFAULT_SEED  # noqa: F401

from .facade import (Game, Agent, option, Card, CFRNode, create_game, create_a_close_to_finished_game,  # noqa: F401
                     create_a_random_game, run_mccfr)

from . import arena  # noqa: F401,E402  (compare_to_random.py semantics: play_games / play_games_batched)
from . import datagen  # noqa: F401,E402
from . import train  # noqa: F401,E402  (algorithms/train.py train_node_value_only on the device)
from . import parallel  # noqa: F401,E402  (multi-GPU: gathered data generation, labelled root-parallel mode)  (train_from_scratch.get_mccfr_targets / generate_test_data.setup_game)

__all__ = ["arena", "datagen", "Engine", "EngineError", "RULESET_PRESET", "RULESET_CLASSIC", "DEFAULT_SEED", "Game", "Agent", "option", "Card",
           "CFRNode", "create_game", "create_a_close_to_finished_game", "create_a_random_game", "run_mccfr"]

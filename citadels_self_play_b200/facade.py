"""Drop-in look-alikes of the reference's hot-path objects, backed by the CUDA engine.

    reference                                   here
    game.game.Game(preset=True)                 facade.Game(preset=True)
    game.setup_round()                          Game.setup_round()
    game.get_options_from_state()               Game.get_options_from_state()      -> [option]
    Agent.get_options(game)                     Agent.get_options(game)
    option.carry_out(game) -> Agent | False     option.carry_out(game) -> Agent | False
    copy.deepcopy(game)                         copy.deepcopy(game)   (pure host data: 256 B record + knowledge)
    run_utils.create_game / create_a_close_to_finished_game / create_a_random_game / run_mccfr
    algorithms.deep_mccfr.CFRNode(game, original_player_id, model=..., training=...)
        .cfr_train(n) / .cfr_pred(n, max_depth) / .children / .cumulative_regrets / .strategy /
        .cumulative_strategy / .node_value / .winning_probabilities / .action_choice(live=True)

Every state change and every option list comes from the kernels (ctd_game_* / ctd_mccfr* in
include/citadels_b200.h); this module only decodes records and descriptors.  There is no CPU path.

Differences a caller can observe, all deliberate:
  * chance comes from the engine's Philox stream keyed by (seed, game id), not from `random`; `random.seed` does not
    steer the deal.  Choosing among options with `random.choice` is of course still the caller's randomness.
  * Game(preset=True) deals AND runs the first setup_round on the device; the first explicit setup_round() call is
    then a no-op, so `create_game()` behaves as in the reference (run_utils.py:20-27).
  * Game(preset=False) deals the random game of game/game.py:491-520 on the device (14 random uniques, four cards
    each, a random variant of every rank, random pick order and crown).
"""
import copy
import ctypes
import itertools
import numpy as np

from . import _lib
from .engine import Engine, EngineError, DEFAULT_SEED, RULESET_PRESET, RULESET_CLASSIC
from .layout import (STATE_DTYPE, KNOW_BYTES, KIND_NAMES, KIND, NAMED_NAMES, SUIT_NAMES, ROLES, COST_OF_TYPE,
                     SUIT_OF_TYPE, opt_fields, record_gold, TreeView, NF_ROLE_PICK, TREE_REF_RAISE)

_default_engine = None
# Games made without an explicit gid take ids from the upper half of the 40-bit id space (explicit ids are expected below 2^39).
# Bits 40.. of a tree's id carry the game's search counter: every run_mccfr of one game draws from its own Philox stream, as
# arena.search_batch does with the decision number.
GID_BITS = 40
_gid_counter = itertools.count(1 << (GID_BITS - 1))


def tree_gid(gid, search_no):
    return (int(gid) & ((1 << GID_BITS) - 1)) | (int(search_no) << GID_BITS)


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(capacity=64)
    return _default_engine


class Card:
    """game/deck.py:7-24: equality and ordering by type_ID only."""

    def __init__(self, code):
        self.type_ID = 25 if code >= 40 else int(code)
        self.suit = SUIT_NAMES[code - 40] if code >= 40 else SUIT_NAMES[SUIT_OF_TYPE[self.type_ID]]
        self.cost = COST_OF_TYPE[self.type_ID]

    def __eq__(self, other):
        return isinstance(other, Card) and self.type_ID == other.type_ID

    def __lt__(self, other):
        return isinstance(other, Card) and self.type_ID < other.type_ID

    def __hash__(self):
        return hash(self.type_ID)

    def __repr__(self):
        return "Card(type_ID=%d, suit=%r, cost=%d)" % (self.type_ID, self.suit, self.cost)


class _Deck:
    def __init__(self, codes):
        self.cards = [Card(c) for c in codes]


class GameState:
    """game/helper_classes.py:16-32 (read-only view)."""

    def __init__(self, rec):
        self.state = int(rec["state"])
        self.player_id = None if rec["player"] == 0xFF else int(rec["player"])
        names = ["smithy", "lab", "magic_school", "museum", "character_ability", "begged", "take_gold"]
        self.already_done_moves = ([n for i, n in enumerate(names) if rec["done"] >> i & 1]
                                   + ["trade_building"] * int(rec["done_builds"] & 15)
                                   + ["non_trade_building"] * int(rec["done_builds"] >> 4))

    def __eq__(self, other):
        return isinstance(other, GameState) and self.state == other.state and self.player_id == other.player_id


class Agent:
    """game/agent.py:10-28 as a view on seat `id` of a Game."""

    def __init__(self, game, seat):
        self._game = game
        self.id = seat

    def _seg(self, k):
        r = self._game._rec
        c = 4 * self.id + k
        return r["arena"][int(r["off"][c]):int(r["off"][c + 1])]

    hand = property(lambda self: _Deck(self._seg(0)))
    buildings = property(lambda self: _Deck(self._seg(1)))
    museum_cards = property(lambda self: _Deck(self._seg(2)))
    just_drawn_cards = property(lambda self: _Deck(self._seg(3)))
    gold = property(lambda self: record_gold(self._game._rec, self.id))
    replicas = property(lambda self: int(self._game._rec["replicas"][self.id]))
    crown = property(lambda self: int(self._game._rec["crown"]) == self.id)
    can_use_lighthouse = property(lambda self: bool(self._game._rec["pflags"][self.id] & 1))
    first_to_7 = property(lambda self: bool(self._game._rec["pflags"][self.id] & 2))
    witch = property(lambda self: bool(self._game._rec["pflags"][self.id] & 4))

    @property
    def role(self):
        r = int(self._game._rec["role"][self.id])
        if r == 8:
            return None
        if r == 9:
            return "Bewitched"
        return ROLES[r][int(self._game._rec["variant"][r])]

    def get_options(self, game):
        """Agent.get_options (game/agent.py:50-83).  The engine enumerates for the player to move."""
        if game.gamestate.player_id != self.id:
            raise ValueError("the engine enumerates options for the player to move (seat %r)" % game.gamestate.player_id)
        return game.get_options_from_state()

    def __eq__(self, other):
        return isinstance(other, Agent) and self.id == other.id and self._game is other._game

    def __hash__(self):
        return hash(self.id)

    def __repr__(self):
        return "Agent(id=%d, role=%r, gold=%d)" % (self.id, self.role, self.gold)


class option:
    """game/option.py:8-18: `name` + `attributes`, equality on both; built from a 64-bit descriptor."""

    def __init__(self, desc, game=None):
        self.desc = int(desc)
        f = opt_fields(self.desc)
        self.name = KIND_NAMES[f["kind"]]
        self._f = f
        self._roles = None if game is None else [ROLES[r][int(game._rec["variant"][r])] for r in range(8)]
        self._attrs = None

    @property
    def attributes(self):
        if self._attrs is None:
            self._attrs = self._decode()
        return self._attrs

    def _decode(self):
        f, n = self._f, self.name
        a = {"perpetrator": f["perp"]}
        card = lambda t: Card(t)
        if n == "role_pick":
            a["choice"] = self._roles[f["rank"]] if self._roles else f["rank"]
        elif n in ("gold_or_card", "navigator_gold_card", "magic_school_choice", "blackmail_response"):
            a["choice"] = NAMED_NAMES[f["named"]]
        elif n == "which_card_to_keep":
            a["choice"] = [card(f["a"])] if f["b"] < 0 else (card(f["a"]), card(f["b"]))
        elif n == "finish_round":
            a["next_witch"] = bool(f["next_witch"])
            a["crown"] = bool(f["crown"])
        elif n == "build":
            a["built_card"] = card(f["a"])
            a["replica"] = f["replica"]
        elif n in ("laboratory_choice", "lighthouse_choice", "museum_choice"):
            a["choice"] = card(f["a"])
        elif n in ("weapon_storage_choice", "warlord_desctruction"):
            a["target"] = f["target"]
            a["choice"] = card(f["a"])
        elif n in ("assassination", "bewitching", "steal"):
            a["choice"] = f["rank"]
        elif n == "spy":
            a["target"] = f["target"]
            a["suit"] = NAMED_NAMES[f["named"]]
        elif n in ("look_at_hand", "magic_hand_change"):
            a["target"] = f["target"]
        elif n == "take_from_hand":
            a["target"] = f["target"]
            a["build"] = bool(f["build"])
            if f["build"]:
                a["built_card"] = card(f["a"])
                a["replica"] = f["replica"]
            else:
                a["card"] = card(f["a"])
        elif n == "abbot_gold_or_card":
            a["gold_or_card_combination"] = ["gold"] * (f["r"] - f["count"]) + ["card"] * f["count"]
        elif n == "discard_and_draw":
            a["subset_size"] = f["r"]       # every such option has the same effect (game/option_functions.py:295-300)
            a["ordinal"] = f["j"]
        # ---- the ten deluxe characters (game/agent_functions.py:221-272, :328-419, :446-504)
        elif n == "magistrate_warrant":
            a["real_target"] = f["rank"]
            a["fake_targets"] = [f["named"], f["count"]]
        elif n == "blackmail":
            a["real_target"] = f["rank"]
            a["fake_target"] = f["named"]
        elif n in ("reveal_blackmail_as_blackmailer", "reveal_warrant_as_magistrate"):
            a["target"] = f["target"]
            a["choice"] = NAMED_NAMES[f["named"]]
        elif n == "give_crown":
            a["target"] = f["target"]
            a["gold_or_card"] = "nothing" if f["named"] == 13 else NAMED_NAMES[f["named"]]
        elif n == "give_back_card":
            # one card per seat the Seer took from, in seat order (game.seer_taken_card_from); zip() may run short
            codes = [f["a"] + 1, f["b"] + 1, f["count"], f["r"], f["j"] & 0x3F]
            a["card_handouts"] = [card(c - 1) for c in codes if c > 0]
        elif n == "scholar_card_pick":
            a["choice"] = card(f["a"])
        elif n == "cardinal_exchange":
            a["target"] = f["target"]
            a["built_card"] = card(f["a"])
            a["replica"] = f["replica"]
            a["factory"] = bool(f["build"])
            a["n_cards_to_give"] = f["count"]   # which f["count"]-subset of the other cards: ordinal j of the thinned list
            a["ordinal"] = f["j"]
        elif n == "marshal_steal":
            a["target"] = f["target"]
            a["choice"] = card(f["a"])
        elif n == "diplomat_exchange":
            a["target"] = f["target"]
            a["choice"] = card(f["a"])
            a["give"] = card(f["b"])
        return a

    def __eq__(self, other):
        return isinstance(other, option) and self.desc == other.desc

    def __hash__(self):
        return hash(self.desc)

    def __str__(self):
        return "%s, %s" % (self.name, self.attributes)

    __repr__ = __str__

    def carry_out(self, game):
        """option.carry_out (game/option.py:118-122): returns the winning Agent when the game ended, else False."""
        return game._step(self.desc)

    def encode_option(self):
        """option.encode_option (game/option.py:52-115) -> torch.float32[1, 131]."""
        import torch
        from .layout import encode_options
        return torch.from_numpy(encode_options([self.desc]))


class Game:
    """game/game.py `Game` for the fixed rulesets: a 256-byte record, the six observers' knowledge and used_cards on
    the host; every transition runs on the device."""

    def __init__(self, preset=True, engine=None, seed=DEFAULT_SEED, gid=None, ruleset=None):
        if ruleset is None:
            ruleset = RULESET_PRESET if preset else 2   # Game(preset=False): CTD_RULESET_RANDOM (game/game.py:491-520)
        self._engine = engine or default_engine()
        self.seed = int(seed)
        self.gid = next(_gid_counter) if gid is None else int(gid)
        self._rec = np.zeros((), dtype=STATE_DTYPE)
        self._know = np.zeros(6 * KNOW_BYTES, dtype=np.uint8)
        self._used = np.zeros(76, dtype=np.uint8)
        lib, h = self._engine._lib, self._engine._h
        self._engine._check(lib.ctd_game_new(h, self.seed, self.gid, ruleset, self._rec.ctypes.data,
                                             self._know.ctypes.data, self._used.ctypes.data), "ctd_game_new")
        self._fresh = True
        self._searches = 0      # run_mccfr calls made from this game so far (keys the search's chance stream)
        self.players = [Agent(self, i) for i in range(6)]

    # -- the reference's attributes ------------------------------------------------------------
    gamestate = property(lambda self: GameState(self._rec))
    terminal = property(lambda self: bool(self._rec["gflags"] & 2))
    ending = property(lambda self: bool(self._rec["gflags"] & 1))
    deck = property(lambda self: _Deck(self._rec["arena"][int(self._rec["off"][24]):int(self._rec["off"][25])]))
    discard_deck = property(lambda self: _Deck(self._rec["arena"][int(self._rec["off"][25]):int(self._rec["off"][26])]))
    turn_orders_for_roles = property(lambda self: [int(x) for x in self._rec["order"]])
    used_roles = property(lambda self: [int(x) - 1 for x in self._rec["used_roles"][:int(self._rec["used_len"])]])
    roles = property(lambda self: {r: ROLES[r][int(self._rec["variant"][r])] for r in range(8)})
    roles_to_choose_from = property(lambda self: {r: ROLES[r][int(self._rec["variant"][r])] for r in range(8)
                                                  if self._rec["rtc_mask"] >> r & 1})

    @property
    def rewards(self):
        r = np.zeros(6)
        if self.terminal:
            r[int(self._rec["winner"])] = 1
        return r

    @property
    def points(self):
        return [int(x) for x in self._rec["points"]]

    @property
    def error_flags(self):
        return int(self._rec["err"])

    def setup_round(self):
        """Game.setup_round (game/game.py:144-171).  The constructor already ran the first one on the device."""
        if self._fresh:
            self._fresh = False
            return
        raise NotImplementedError("setup_round is driven by finish_round inside the engine (game/option_functions.py:236)")

    def get_options_from_state(self):
        """game/game.py:415-418."""
        lib, h = self._engine._lib, self._engine._h
        cap = 16384   # one call: the Seer's / Scholar's enumerations are not repeatable (they draw chance / shrink a list)
        opts = np.zeros(cap, dtype=np.uint64)
        n = ctypes.c_uint32()
        self._engine._check(lib.ctd_game_options(h, self.seed, self._rec.ctypes.data, self._know.ctypes.data, opts.ctypes.data,
                                                 cap, ctypes.byref(n)), "ctd_game_options")
        return [option(d, self) for d in opts[:n.value]]

    def _step(self, desc):
        lib, h = self._engine._lib, self._engine._h
        w = ctypes.c_int8()
        self._fresh = False
        self._engine._check(lib.ctd_game_step(h, self.seed, self._rec.ctypes.data, self._know.ctypes.data, desc,
                                              ctypes.byref(w)), "ctd_game_step")
        if self._rec["err"]:
            raise EngineError("engine error flags 0x%x (2 = the reference would raise here)" % int(self._rec["err"]))
        return self.players[w.value] if w.value >= 0 else False

    def encode_game(self):
        """Game.encode_game (game/game.py:91-128) -> torch.float32[418], computed on the device."""
        import torch
        e = self._engine
        viewer = 0 if self._rec["player"] == 0xFF else int(self._rec["player"])
        e.load_roots(np.frombuffer(self._rec.tobytes(), dtype=np.uint8), self._know[viewer * KNOW_BYTES:(viewer + 1) * KNOW_BYTES],
                     self._used, np.array([self.gid], dtype=np.uint64))
        f = e.encode(1, cfr_role_pick=False)[0]
        return torch.from_numpy(np.ascontiguousarray(f))

    def record(self):
        """The packed 256-byte record (include/citadels_b200.h ctd_state)."""
        return self._rec.tobytes()

    def get_option_from_role_preference(self, strategy):
        """Game.get_option_from_role_preference (game/game.py:312-317): `strategy` is indexed by the RANKS of the roles on
        offer (the caller, CFRNode.action_choice(live=True), passes a row with one entry per child -- the reference's quirk)."""
        options = self.get_options_from_state()
        role_ids = [opt_fields(o.desc)["rank"] for o in options]
        substrategy = np.asarray(strategy, dtype=float)[role_ids]
        substrategy = substrategy / substrategy.sum()
        return options[np.random.choice(len(options), p=substrategy)]

    def sample_private_information(self, player_character, role_sample=True):
        """Game.sample_private_information (game/game.py:215-242): determinise what `player_character` cannot see, in place,
        on the device (ctd_game_sample).  Inside searches this runs in the kernels; the method exists for callers that
        determinise a game themselves."""
        seat = player_character.id if hasattr(player_character, "id") else int(player_character)
        lib, h = self._engine._lib, self._engine._h
        self._engine._check(lib.ctd_game_sample(h, self.seed, self._rec.ctypes.data, self._know.ctypes.data, self._used.ctypes.data,
                                                seat, 1 if role_sample else 0), "ctd_game_sample")
        if self._rec["err"]:
            raise EngineError("engine error flags 0x%x (2 = the reference would raise here)" % int(self._rec["err"]))

    def __deepcopy__(self, memo):
        g = Game.__new__(Game)
        g._engine = self._engine
        g.seed, g.gid, g._fresh, g._searches = self.seed, self.gid, self._fresh, self._searches
        g._rec = self._rec.copy()
        g._know = self._know.copy()
        g._used = self._used.copy()
        g.players = [Agent(g, i) for i in range(6)]
        return g

    def __eq__(self, other):
        return isinstance(other, Game) and self._rec.tobytes()[:228] == other._rec.tobytes()[:228]


# ------------------------------------------------------------------------------------------------ run_utils
def create_game(engine=None, seed=DEFAULT_SEED, gid=None, ruleset=RULESET_PRESET):
    """run_utils.create_game (run_utils.py:20-27)."""
    g = Game(preset=True, engine=engine, seed=seed, gid=gid, ruleset=ruleset)
    g.setup_round()
    return g


def _random_play(game, choose):
    games = [copy.deepcopy(game)]
    winner = False
    while not winner:
        options = game.get_options_from_state()
        winner = choose(options).carry_out(game)
        games.append(copy.deepcopy(game))
    return games


def create_a_close_to_finished_game(game, rng=None):
    """run_utils.create_a_close_to_finished_game (run_utils.py:29-50), one game through the facade.  For many roots
    use Engine.make_roots, which does the same on the device."""
    import random
    rng = rng or random
    move_stop_num = rng.randint(1, 30)
    games = _random_play(game, rng.choice)
    options, limit = [], 0
    while len(options) < 2 and limit < 100:
        almost_won_game = games[-move_stop_num]
        options = almost_won_game.get_options_from_state()
        move_stop_num -= 1
        limit += 1
    return almost_won_game


def create_a_random_game(max_move_num, engine=None, rng=None, **kw):
    """run_utils.create_a_random_game (run_utils.py:52-72)."""
    import random
    rng = rng or random
    move_stop_num = rng.randint(1, max_move_num)
    games = _random_play(create_game(engine, **kw), rng.choice)
    return games[-move_stop_num]


class CFRNode:
    """algorithms/deep_mccfr.py `CFRNode`, root-facing surface.  The search runs in ctd_mccfr / ctd_mccfr_pred; this
    object exposes the resulting tree with the reference's field names."""

    def __init__(self, game, original_player_id, parent=None, player_count=6, model=None, training=False, device=None,
                 model_reward_weights=5, depth=0, _tree=None, _index=0):
        self.game = game
        self.original_player_id = original_player_id
        self.parent = parent
        self.model = model
        self.training = training
        self.model_reward_weights = model_reward_weights
        self.depth = depth
        self._tree = _tree
        self._index = _index
        self._children = None
        self.live_option = None
        if _tree is None:
            self.current_player_id = game.gamestate.player_id
            self.role_pick_node = game.gamestate.state == 0
            self.cumulative_regrets = np.array([])
            self.strategy = np.array([])
            self.cumulative_strategy = np.array([])
            self.node_value = np.zeros(player_count)
            self.winning_probabilities = np.zeros(player_count)
        else:
            self._fill()

    def _fill(self):
        tv, i = self._tree, self._index
        n = tv.nodes[i]
        self.current_player_id = int(n["player"])
        self.role_pick_node = bool(n["flags"] & NF_ROLE_PICK)
        R, S, C = tv.arrays(i)
        self.cumulative_regrets, self.strategy, self.cumulative_strategy = np.array(R), np.array(S), np.array(C)
        self.node_value = np.array(n["V"])
        self.winning_probabilities = np.array(n["P"])

    def _run(self, iterations, max_depth=None):
        g = self.game
        e = g._engine
        v = self.original_player_id
        e.load_roots(np.frombuffer(g._rec.tobytes(), dtype=np.uint8), g._know[v * KNOW_BYTES:(v + 1) * KNOW_BYTES], g._used,
                     np.array([tree_gid(g.gid, g._searches)], dtype=np.uint64))
        g._searches += 1
        ruleset = int(g._rec["ruleset"])
        if max_depth is None:
            out = e.mccfr(1, iterations=iterations, seed=g.seed, ruleset=ruleset, trees=True)
        else:
            e.set_value_model(self.model)
            out = e.mccfr_pred(1, iterations=iterations, max_depth=max_depth, seed=g.seed, ruleset=ruleset,
                               weight=float(self.model_reward_weights), trees=True)
        res = out["results"][0]
        self.status = int(res["status"])
        if self.status & TREE_REF_RAISE:
            raise EngineError("MCCFR status %d: the reference raises inside its rules code for this root" % self.status)
        if self.status & ~1:
            raise EngineError("MCCFR status %d (2 device memory exhausted, 4 container capacity)" % self.status)
        self._tree = out["trees"][0]
        self._index = 0
        self._children = None
        self.live_option = int(res["live_option"])
        # skip_false_choice advances the caller's game (algorithms/deep_mccfr.py:19-20, :37-49)
        g._rec = np.array(self._tree.nodes[0]["game"])
        g._know[v * KNOW_BYTES:(v + 1) * KNOW_BYTES] = np.frombuffer(self._tree.nodes[0]["know"].tobytes(), dtype=np.uint8)
        self._fill()

    def cfr_train(self, max_iterations=100000):
        """algorithms/deep_mccfr.py:187-205."""
        self._run(max_iterations)

    def cfr_pred(self, max_iterations=2000, max_depth=20):
        """algorithms/deep_mccfr.py:207-229.  With training=True the reference never consumes the model's output
        (deep_mccfr.py:119-126, :147, :178 guard on `not self.training`; backpropagate always accumulates), so cfr_pred is
        only meaningful with training=False, as in run_utils.run_mccfr."""
        if self.model is None:
            raise ValueError("cfr_pred needs a value model")
        self._run(max_iterations, max_depth)

    @property
    def children(self):
        """[(option, CFRNode)] like the reference."""
        if self._tree is None:
            return []
        if self._children is None:
            self._children = []
            for desc, idx in self._tree.child_list(self._index):
                cg = Game.__new__(Game)
                cg._engine, cg.seed, cg.gid, cg._fresh, cg._searches = self.game._engine, self.game.seed, self.game.gid, False, 0
                cg._rec = np.array(self._tree.nodes[idx]["game"])
                cg._know = self.game._know.copy()   # only the searching player's block is tracked inside a tree
                v = self.original_player_id
                cg._know[v * KNOW_BYTES:(v + 1) * KNOW_BYTES] = np.frombuffer(self._tree.nodes[idx]["know"].tobytes(), dtype=np.uint8)
                cg._used = self.game._used
                cg.players = [Agent(cg, i) for i in range(6)]
                node = CFRNode(cg, self.original_player_id, parent=self, model=self.model, training=self.training,
                               model_reward_weights=self.model_reward_weights, depth=self.depth + 1, _tree=self._tree,
                               _index=idx)
                self._children.append((option(desc, self.game), node))
        return self._children

    def is_terminal(self):
        return self.game.terminal

    def get_reward(self):
        return self.game.rewards

    def action_choice(self, live=False):
        """algorithms/deep_mccfr.py:67-91.  `live=True` on a searched root returns the decision the kernel drew from the tree's own
        chance stream right after the search (ctd_mccfr_result.live_option: the cumulative-strategy draw, or at a role-pick root
        the preference quirk of game/game.py:312-317); on any other node the draw comes from numpy like in the reference."""
        kids = self.children
        if not kids and not self.role_pick_node:
            raise ValueError("a terminal root has no children (the reference raises ValueError here too)")
        if live and self.parent is None and self.live_option:
            chosen = option(self.live_option, self.game)
            if self.role_pick_node:
                return None, chosen
            for o, node in kids:
                if o.desc == chosen.desc:
                    return node, o
            return None, chosen
        if not self.role_pick_node:
            p = self.cumulative_strategy / self.cumulative_strategy.sum()
            i = np.random.choice(range(len(kids)), p=p)
            return kids[i][1], kids[i][0]
        if live:
            return None, self.game.get_option_from_role_preference(self.strategy[self.game.gamestate.player_id])
        order = self.game.turn_orders_for_roles
        avg = np.zeros(self.cumulative_strategy.shape[1])
        for i, pl in enumerate(order):
            avg += self.cumulative_strategy[pl] * (6 - i)
        avg = avg / sum(order)
        p = np.ones(len(kids)) / len(kids) if avg.sum() == 0 else avg / avg.sum()
        i = np.random.choice(range(len(kids)), p=p)
        return kids[i][1], kids[i][0]


def run_mccfr(game, model=None, max_iterations=2000, training=False):
    """run_utils.run_mccfr (run_utils.py:74-87).  training=True runs cfr_train whether or not a model is given, as the
    reference does (the model's output is never consumed in that mode)."""
    root = CFRNode(game, original_player_id=game.gamestate.player_id, model=model, training=training)
    if model is not None and not training:
        root.cfr_pred(max_iterations=max_iterations, max_depth=10)
    else:
        root.cfr_train(max_iterations=max_iterations)
    _, chosen = root.action_choice(live=True)
    return chosen, root

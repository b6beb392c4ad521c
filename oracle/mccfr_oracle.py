"""CPU restatement of the reference's MCCFR tree search (algorithms/deep_mccfr.py) over the oracle Game.

TEST INFRASTRUCTURE ONLY (same rules as citadels_oracle.py).  fp64 numpy, same operation order as the
reference.  Chance comes from one Philox stream per tree (oracle/philox.py, stream 1):
  np.random.choice(range(n), p) in action_choice      -> inverse CDF on one uniform() (numpy's own recipe:
                                                          cdf = cumsum(p); cdf /= cdf[-1]; searchsorted right)
  np.random.choice(range(n), uniform) in expand_*     -> randbelow(n)
  random.random() / random.choice / random.shuffle    -> uniform() / randbelow / perm
  np.random.choice(options, p) in Game.get_option_from_role_preference (live role-pick decisions) -> inverse CDF, as above
The reference harness (tests/golden/ref_harness.py) patches exactly those call sites of the real
reference with the same mapping, which is how this file is pinned (tests/golden/check_mccfr_vs_ref.py).
"""
import numpy as np

from . import citadels_oracle as O

LOG13 = np.log(1.3)


def choice_p(ch, p):
    cdf = np.cumsum(p)
    cdf = cdf / cdf[-1]
    u = ch.uniform()
    i = int(np.searchsorted(cdf, u, side="right"))
    return min(i, len(p) - 1)


def carried_form(g, d):
    """The option object as it looks after option.carry_out: carry_out_wizard_take_from_hand rewrites
    attributes['replica'] with the builder's copy count (game/option_functions.py:317).  CFRNode stores and
    compares (`in child_options`, algorithms/deep_mccfr.py:169-170) the rewritten object."""
    if O.d_kind(d) == O.K["take_from_hand"] and O.d_build(d):
        p, t = O.d_perp(d), O.d_a(d)
        rep = sum(1 for c in g.bld[p] if O.ctype(c) == t)
        return O.D(O.K["take_from_hand"], p, target=O.d_target(d), a=t, build=1, replica=rep)
    return d


class Node:
    """CFRNode, algorithms/deep_mccfr.py:8-33."""

    def __init__(self, game, orig, parent=None, model=None, training=False, depth=0, weight=5):
        self.model = model
        self.weight = weight
        self.orig = orig
        self.training = training
        self.depth = depth
        self.game = game
        self.error = False
        self.skip_false_choice()
        self.parent = parent
        self.children = []      # (descriptor, Node)
        self.player = game.player
        self.R = np.array([])
        self.s = np.array([])
        self.C = np.array([])
        self.V = np.zeros(6)
        self.P = np.zeros(6)
        self.role_pick = game.state == 0
        self.pred = None

    def skip_false_choice(self):
        """:37-49 (mutates the game it was given, the caller's included)."""
        g = self.game
        i = 0
        opts = g.options()
        terminal = False
        while len(opts) == 1 and not terminal:
            i += 1
            terminal = g.apply(opts[0])
            opts = g.options()
            if i > 100:
                terminal = True

    def is_terminal(self):
        return self.game.terminal

    def reward(self):
        r = np.zeros(6)
        r[self.game.winner] = 1
        return r

    # -------------------------------------------------------------- expansion (:93-179)
    def _child(self, g):
        return Node(g, self.orig, parent=self, model=self.model, training=self.training, depth=self.depth + 1,
                    weight=self.weight)

    def _maybe_sample(self, g):
        if self.parent is None or g.player != self.parent.game.player:
            g.sample_private_information(self.orig, role_sample=(self.parent.game.state != 0) if self.parent else False)

    def expand(self):
        if self.game.state == 0 and not self.children:
            self.role_pick = True
            self.expand_role_pick()
        elif self.player == self.orig and not self.children:
            self.expand_for_original_player()
        elif self.player != self.orig and len(self.children) < 10:
            self.expand_for_opponents()

    def expand_role_pick(self):
        ch = self.game.chance
        for _ in range(10):
            g = self.game.clone()
            d = None
            while g.state != 1:
                opts = g.options()
                d = opts[ch.randbelow(len(opts))]
                g.apply(d)
            self.children.append((d, self._child(g)))   # role_pick options are never rewritten
        if self.model is not None:
            g = self.game.clone()
            g.player = 5
            self.pred = self.weight * self.model(g) if not self.training else self.pred
        self.R = np.zeros((6, 10))
        self.s = np.zeros((6, 10))
        self.C = np.zeros((6, 10))

    def expand_for_original_player(self):
        for d in self.game.options():
            g = self.game.clone()
            self._maybe_sample(g)
            d = carried_form(g, d)
            g.apply(d)
            self.children.append((d, self._child(g)))
        if self.model is not None and not self.training:
            self.pred = self.weight * self.model(self.game)
        k = len(self.children)
        self.R, self.s, self.C = np.zeros(k), np.zeros(k), np.zeros(k)

    def expand_for_opponents(self):
        ch = self.game.chance
        g = self.game.clone()
        self._maybe_sample(g)
        opts = g.options()
        d = carried_form(g, opts[ch.randbelow(len(opts))])
        g.apply(d)
        if d not in [c[0] for c in self.children]:
            self.children.append((d, self._child(g)))
            self.R = np.append(self.R, 0)
            self.s = np.append(self.s, 0)
            if self.model is not None and not self.training:
                self.pred = self.weight * self.model(self.game)
            self.C = np.append(self.C, 0)

    # -------------------------------------------------------------- strategy / regrets
    def update_strategy(self):
        """:292-319."""
        t = np.exp(-self.R * LOG13)
        if not self.role_pick:
            tot = np.sum(t)
            self.s = t / tot if tot > 0 else np.ones_like(t) / len(t)
        else:
            tot = np.sum(t, axis=0)
            if np.any(tot <= 1e-8):
                self.s = np.where(tot > 1e-8, t / tot, 1.0 / t.shape[0])
            else:
                self.s = t / tot
        self.C = self.C + self.s
        self.C = self.C / self.C.sum()

    def action_choice(self):
        """:67-91 (the non-live branches)."""
        ch = self.game.chance
        if not self.role_pick:
            p = self.C / self.C.sum()
        else:
            order = self.game.order
            avg = np.zeros(self.C.shape[1])
            for i, pl in enumerate(order):
                avg += self.C[pl] * (6 - i)
            avg = avg / sum(order)
            p = np.ones(len(self.children)) / len(self.children) if avg.sum() == 0 else avg / avg.sum()
        i = choice_p(ch, p)
        return self.children[i][1]

    def live_choice(self):
        """action_choice(live=True), :67-75 -- what run_mccfr returns (run_utils.py:82,86).  Ordinary roots: a child drawn from
        the cumulative strategy, its stored option.  Role-pick roots: Game.get_option_from_role_preference (game/game.py:312-317)
        on self.strategy[player to move]: that row has one entry per CHILD but is indexed by the RANKS on offer."""
        ch = self.game.chance
        if not self.role_pick:
            if not self.children:
                raise ValueError("a terminal root has no children")     # np.random.choice on an empty range
            p = self.C / self.C.sum()
            return self.children[choice_p(ch, p)][0]
        opts = self.game.options()
        sub = self.s[self.game.player][[O.d_rank(d) for d in opts]]
        sub = sub / sub.sum()
        return opts[choice_p(ch, sub)]

    def update_regrets(self):
        """:231-256."""
        if not self.role_pick:
            a = [c[1].P[self.player] for c in self.children]
            m = max(a)
            for i in range(len(a)):
                self.R[i] += m - a[i]
        else:
            a = np.array([c[1].P for c in self.children]).T
            self.R += np.max(a, axis=0) - a

    def backpropagate(self, reward):
        """:276-290."""
        node = self
        while node is not None:
            if node.training or (node.V.sum() == 0 or node.model is None):
                node.V = node.V + reward
            node.P = node.V / node.V.sum()
            if node.children:
                node.update_regrets()
            node = node.parent

    # -------------------------------------------------------------- the loops
    def cfr_train(self, max_iterations):
        """:187-205."""
        if self.is_terminal():
            return
        self.expand()
        node = self
        for _ in range(max_iterations):
            node.update_strategy()
            node = node.action_choice()
            if node.is_terminal():
                node.backpropagate(node.reward())
                node.update_strategy()
                node = self
            else:
                node.expand()
        self.update_strategy()

    def cfr_pred(self, max_iterations, max_depth):
        """:207-229."""
        if self.is_terminal():
            return
        self.expand()
        node = self
        for _ in range(max_iterations):
            node.update_strategy()
            node = node.action_choice()
            if node.depth > max_depth and not node.is_terminal():
                node.expand()
                node.backpropagate(node.pred)
                node.update_strategy()
                node = self
            elif node.is_terminal():
                node.backpropagate(node.reward())
                node.update_strategy()
                node = self
            else:
                node.expand()
        self.update_strategy()

    # -------------------------------------------------------------- inspection
    def walk(self):
        yield self
        for _, c in self.children:
            yield from c.walk()


def make_root(seed, gid, ruleset=O.RULESET_PRESET, back_lo=0, back_hi=20, flavour=0):
    """Root construction as the engine defines it (ctd_make_roots, include/citadels_b200.h).
    flavour 0 restates run_utils.create_a_close_to_finished_game (run_utils.py:29-50): play game (seed, gid) to terminal
    (T steps), u = back_lo + randbelow(back_hi - back_lo + 1) from Philox stream word 2, k = max(0, T - u); replay to step k,
    then keep stepping while the player to move has fewer than 2 options (at most 100 times).
    flavour 1 restates run_utils.create_a_random_game (run_utils.py:52-72): m drawn the same way, root = games[-m] = the
    state after max(0, T + 1 - m) steps, not moved forward (CFRNode.skip_false_choice does that, AFTER run_mccfr fixed the
    searching player).
    Returns (game, root_step); the searching player is game.player."""
    from .philox import PhiloxChance
    T = O.playout(seed, gid, ruleset)[2]
    u = back_lo + PhiloxChance(seed, gid, stream=2).randbelow(back_hi - back_lo + 1)
    k = max(0, (T + 1 if flavour == 1 else T) - u)
    ch = PhiloxChance(seed, gid)
    g = O.new_game(ch, ruleset)
    steps = limit = 0
    while not g.terminal:
        if flavour == 1 and steps >= k:
            break
        opts = g.options()
        if steps >= k:
            if len(opts) >= 2 or limit >= 100:
                break
            limit += 1
        g.apply(opts[ch.randbelow(len(opts))])
        steps += 1
    return g, steps


def run_from_root(root_rec, know_blob, used_cards, seed, gid, iterations):
    """CFRNode(game, original_player_id=game.gamestate.player_id).cfr_train(iterations) from a packed root."""
    from .philox import PhiloxChance
    g = O.Game.unpack(bytes(root_rec), PhiloxChance(seed, gid, stream=1))
    g.unpack_know(bytes(know_blob), used_cards)
    n = Node(g, g.player)
    n.cfr_train(iterations)
    return n


# ------------------------------------------------------------------ training targets (algorithms/deep_mccfr.py:258-345)
def encode_option(d):
    """option.encode_option (game/option.py:52-115) from a descriptor -> list of 131 numbers."""
    e = [0.0] * 131
    k, name = O.d_kind(d), O.KIND_NAMES[O.d_kind(d)]
    e[k] = 1.0
    e[O.d_perp(d) + 47] = 1.0
    if O.d_target(d) >= 0:                       # "target" in attributes
        e[O.d_target(d) + 53] = 1.0
    elif name == "role_pick":                    # choice is a role name
        e[O.d_rank(d) + 60] = 1.0
    elif O.d_named(d) >= 0:                      # choice is one of the 13 named strings
        e[O.d_named(d) + 76] = 1.0
    elif name in ("laboratory_choice", "lighthouse_choice", "museum_choice"):   # choice is a Card
        e[O.d_a(d) + 89] = 1.0
    elif name == "which_card_to_keep":           # choice is a list [card]; Library pairs are tuples -> no bits
        if O.d_b(d) < 0:
            e[O.d_a(d) + 89] = 1.0
    elif name == "build":                        # built_card
        e[O.d_a(d) + 89] = 1.0
    elif name == "abbot_gold_or_card":
        e[130] = float(O.d_count(d))
    return e


def get_all_targets(root, seed, gid, threshold=15):
    """CFRNode.get_all_targets (:258-274) -> list of (features[418], options[K][131], node_value[6], regrets[K])."""
    from .philox import PhiloxChance
    ch = PhiloxChance(seed, gid, stream=3)
    out = []
    for n in root.walk():
        if not n.children or n.V.sum() < threshold:
            continue
        if n.role_pick:
            i = ch.randbelow(6)
            feats = n.game.encode_game(i)
            dist = np.array(n.R[i], dtype=float)
        else:
            feats = n.game.encode_game()
            dist = np.array(n.R, dtype=float)
        if dist.sum() == 0:
            dist = np.ones_like(dist)
        out.append((feats, [encode_option(c[0]) for c in n.children], np.array(n.V), dist))
    return out

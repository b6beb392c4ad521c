"""CPU restatement of the reference's Citadels rules engine (enumeration + transition).

TEST INFRASTRUCTURE ONLY.  Nothing in citadels_self_play_b200/ imports this file; it is
the checker for the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs).  It follows the reference *as implemented*,
quirks included; every function cites the reference lines it
restates (paths relative to /root/reference).

Data model (plain ints and lists, no Card/Deck/Agent objects):
  card code  0..39  = type_ID with its printed suit; 40..43 = Magic School (type 25)
             whose suit was rewritten to trade/war/religion/lord
             (game/option_functions.py:147-150 appends a *new* Card with the chosen suit).
  role       0..7   = rank (the name is roles-variant[rank]); 8 = None; 9 = "Bewitched"
  option     64-bit descriptor, see `D()` below.

Parity pin: tests/golden/*.npz are produced by running the *real* reference
(/root/reference/game) under the same chance stream (tests/golden/gen_golden.py) and
hashing its states/options through tests/golden/ref_harness.py; tests/test_oracle_golden.py
replays them through this file.
"""
from math import comb
import struct

# ----------------------------------------------------------------------------- tables
# game/config.py:2-80
SUIT_TRADE, SUIT_WAR, SUIT_RELIGION, SUIT_LORD, SUIT_UNIQUE = range(5)
SUIT_OF_TYPE = [0] * 6 + [1] * 4 + [2] * 3 + [3] * 3 + [4] * 24
COST_OF_TYPE = [1, 2, 4, 2, 5, 3, 2, 3, 5, 1, 2, 3, 1, 4, 3, 5,
                5, 3, 6, 2, 6, 5, 5, 6, 5, 6, 6, 3, 6, 3, 5, 5, 6, 5, 4, 6, 5, 4, 0, 5]
# building_cards + unique_building_cards in list order (game/config.py:2-80)
BASE_DECK = ([0] * 5 + [1] * 3 + [2] * 3 + [3] * 4 + [4] * 2 + [5] * 3 + [6] * 3 + [7] * 3 + [8] * 2 + [9] * 3
             + [10] * 3 + [11] * 3 + [12] * 3 + [13] * 4 + [14] * 5 + [15] * 3
             + [16, 17, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 39])
assert len(BASE_DECK) == 76
# game/game.py:428-474 (indices into building_cards / unique_building_cards resolved to type_IDs)
PRESET_HANDS = [[0, 0, 16, 17, 18, 19], [1, 1, 20, 21, 22, 23], [2, 3, 24, 25, 26, 27],
                [3, 4, 28, 29, 30, 31], [4, 0, 32, 33, 34, 35], [0, 1, 36, 37, 39, 0]]

ROLE_NONE, ROLE_BEWITCHED = 8, 9
# role name id = rank*3 + variant (game/config.py:83-91)
(ASSASSIN, WITCH, MAGISTRATE, THIEF, SPY, BLACKMAILER, MAGICIAN, WIZARD, SEER, KING, EMPEROR, PATRICIAN,
 BISHOP, ABBOT, CARDINAL, MERCHANT, ALCHEMIST, TRADER, ARCHITECT, NAVIGATOR, SCHOLAR, WARLORD, DIPLOMAT,
 MARSHAL) = range(24)
NAME_NONE, NAME_BEWITCHED = 24, 25
ROLE_NAMES = ["Assassin", "Witch", "Magistrate", "Thief", "Spy", "Blackmailer", "Magician", "Wizard", "Seer",
              "King", "Emperor", "Patrician", "Bishop", "Abbot", "Cardinal", "Merchant", "Alchemist", "Trader",
              "Architect", "Navigator", "Scholar", "Warlord", "Diplomat", "Marshal"]
RULESET_PRESET, RULESET_CLASSIC, RULESET_RANDOM = 0, 1, 2
# unique_building_cards in list order (game/config.py:56-80)
UNIQUE_CARDS = [16, 17, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 39]
NAMED_NOTHING = 13   # give_crown's gold_or_card="nothing" (not in the reference's names table; descriptor only)
RULESET_VARIANTS = {
    RULESET_PRESET: [1, 1, 1, 0, 1, 1, 1, 0],   # Witch Spy Wizard King Abbot Alchemist Navigator Warlord (game/game.py:479-486)
    RULESET_CLASSIC: [0, 0, 0, 0, 0, 0, 0, 0],  # Assassin Thief Magician King Bishop Merchant Architect Warlord
    RULESET_RANDOM: [0, 0, 0, 0, 0, 0, 0, 0],   # drawn per game, Game.set_random_game (game/game.py:491-520)
}

# option kinds = index in game/option.py:34-45
KIND_NAMES = [
    "role_pick", "gold_or_card", "which_card_to_keep", "blackmail_response",
    "reveal_blackmail_as_blackmailer", "reveal_warrant_as_magistrate", "build", "empty_option",
    "finish_round", "ghost_town_color_choice", "smithy_choice", "laboratory_choice",
    "magic_school_choice", "weapon_storage_choice", "lighthouse_choice", "museum_choice",
    "graveyard", "take_gold_for_war", "assassination", "magistrate_warrant", "bewitching",
    "steal", "blackmail", "spy", "magic_hand_change", "discard_and_draw", "look_at_hand",
    "take_from_hand", "seer", "give_back_card", "take_crown_king", "give_crown",
    "take_crown_pat", "bishop", "cardinal_exchange", "abbot_gold_or_card", "abbot_beg",
    "merchant", "alchemist", "trader", "architect", "navigator_gold_card", "scholar",
    "scholar_card_pick", "warlord_desctruction", "marshal_steal", "diplomat_exchange"]
K = {n: i for i, n in enumerate(KIND_NAMES)}
# named choices, game/option.py:69-83
NAMED = {"gold": 0, "card": 1, "pay": 2, "not_pay": 3, "reveal": 4, "not_reveal": 5, "4gold": 6, "4card": 7,
         "trade": 8, "war": 9, "religion": 10, "lord": 11, "unique": 12}

# done-move flag bits (game/helper_classes.py:21; names appended in game/option_functions.py)
DM_SMITHY, DM_LAB, DM_MAGIC_SCHOOL, DM_MUSEUM, DM_CHARACTER, DM_BEGGED, DM_TAKE_GOLD = (1 << i for i in range(7))
# next_gamestate modes
NEXT_NONE, NEXT_ALIAS, NEXT_RESET_CA, NEXT_EMPTY = range(4)


def ctype(c):
    return 25 if c >= 40 else c


def csuit(c):
    return c - 40 if c >= 40 else SUIT_OF_TYPE[c]


def ccost(c):
    return COST_OF_TYPE[ctype(c)]


def D(kind, perp, target=None, a=None, b=None, rank=None, named=None, replica=0, build=0, next_witch=0,
      crown=0, count=0, r=0, j=0):
    """64-bit option descriptor (include/citadels_b200.h, CTD_OPT_* field macros).

    [0:6) kind  [6:9) perpetrator  [9:12) target+1  [12:18) card A type+1  [18:24) card B type+1
    [24:28) rank+1  [28:32) named+1  [32:36) replica (4-bit two's complement)  [36] build
    [37] next_witch  [38] crown  [39:45) count  [45:51) subset size r (Magician) / list length (Abbot)  [51:61) ordinal j
    """
    d = kind | (perp << 6)
    if target is not None:
        d |= (target + 1) << 9
    if a is not None:
        d |= (a + 1) << 12
    if b is not None:
        d |= (b + 1) << 18
    if rank is not None:
        d |= (rank + 1) << 24
    if named is not None:
        d |= (named + 1) << 28
    d |= (int(replica) & 0xF) << 32
    d |= (1 if build else 0) << 36
    d |= (1 if next_witch else 0) << 37
    d |= (1 if crown else 0) << 38
    d |= (count & 0x3F) << 39
    d |= (r & 0x3F) << 45
    d |= (j & 0x3FF) << 51
    return d


def D_handout(perp, types):
    """give_back_card: up to five card types (+1) in fields a, b, count, r, j."""
    t = [(x + 1) for x in types] + [0] * (5 - len(types))
    d = K["give_back_card"] | (perp << 6)
    d |= t[0] << 12 | t[1] << 18 | t[2] << 39 | t[3] << 45 | t[4] << 51
    return d


def d_handout(d):
    v = [(d >> 12) & 0x3F, (d >> 18) & 0x3F, (d >> 39) & 0x3F, (d >> 45) & 0x3F, (d >> 51) & 0x3F]
    return [x - 1 for x in v if x]


def unrank_combination(n, k, i):
    """Indices of the i-th k-combination of range(n) in itertools.combinations order."""
    out, x = [], 0
    for r in range(k, 0, -1):
        while True:
            c = comb(n - x - 1, r - 1)
            if i < c:
                break
            i -= c
            x += 1
        out.append(x)
        x += 1
    return out


def d_kind(d): return d & 0x3F
def d_perp(d): return (d >> 6) & 7
def d_target(d): return ((d >> 9) & 7) - 1
def d_a(d): return ((d >> 12) & 0x3F) - 1
def d_b(d): return ((d >> 18) & 0x3F) - 1
def d_rank(d): return ((d >> 24) & 0xF) - 1
def d_named(d): return ((d >> 28) & 0xF) - 1
def d_replica(d):
    v = (d >> 32) & 0xF
    return v - 16 if v >= 8 else v
def d_build(d): return (d >> 36) & 1
def d_next_witch(d): return (d >> 37) & 1
def d_crown(d): return (d >> 38) & 1
def d_count(d): return (d >> 39) & 0x3F
def d_r(d): return (d >> 45) & 0x3F
def d_j(d): return (d >> 51) & 0x3FF


def py_round_div100(x):
    """round(x / 1e2) with CPython's round-half-even (game/agent_functions.py:293)."""
    return round(x / 1e2)


class OracleError(Exception):
    """The reference itself would raise here (KeyError / AttributeError / IndexError)."""


class HandKnowledge:
    """game/helper_classes.py:37-43."""
    __slots__ = ("pid", "conf", "cards", "wizard", "used")

    def __init__(self, pid, cards, conf=5, wizard=False):
        self.pid, self.cards, self.conf, self.wizard, self.used = pid, list(cards), conf, wizard, False

    def copy(self):
        h = HandKnowledge(self.pid, self.cards, self.conf, self.wizard)
        h.used = self.used
        return h


class Game:
    """Restates game/game.py `Game` + game/agent.py `Agent` as flat fields."""

    # ------------------------------------------------------------------ construction
    def __init__(self, chance, ruleset=RULESET_PRESET, deal=True):
        self.chance = chance
        self.ruleset = ruleset
        self.variant = list(RULESET_VARIANTS[ruleset])
        n = 6
        self.hand = [[] for _ in range(n)]
        self.bld = [[] for _ in range(n)]
        self.mus = [[] for _ in range(n)]
        self.jd = [[] for _ in range(n)]          # just_drawn_cards
        self.deck = []
        self.discard = []
        self.used_cards = []                      # game/game.py:424 (order matters only for CFR sampling)
        self.role = [ROLE_NONE] * n
        self.gold = [2] * n                       # game/agent.py:22
        self.replicas = [0] * n                   # False == 0
        self.lighthouse = [False] * n             # can_use_lighthouse
        self.first7 = [False] * n
        self.witch = [False] * n
        self.crown = 3                            # game/game.py:477
        self.order = [0, 1, 2, 3, 4, 5]           # turn_orders_for_roles, game/game.py:489
        # role_properties, game/helper_classes.py:1-14  (warrant/blackmail: 0 None, 1 Real, 2 Fake)
        self.dead = [False] * 8
        self.warrant = [0] * 8
        self.possessed = [False] * 8
        self.robbed = [False] * 8
        self.blackmail = [0] * 8
        # gamestate, game/helper_classes.py:16-32
        self.state = 0
        self.player = 0xFF                         # GameState() has player_id None until setup_round
        self.done = 0                              # flag bits
        self.n_trade = 0                           # count of "trade_building"
        self.n_nontrade = 0                        # count of "non_trade_building"
        self.next_player = 0
        self.next_mode = NEXT_NONE
        self.rtc = []                              # roles_to_choose_from (ranks, ascending)
        self.used_roles = []
        self.ending = False
        self.terminal = False
        self.winner = -1
        self.points = [0] * n
        self.warrant_building = 0xFF
        # wizard look (HandKnowledge with wizard=True of the acting wizard), game/option_functions.py:305-310
        self.wiz_target = 0xFF
        self.wiz_cards = []
        # knowledge (CFR path only)
        self.kr_mask = [[0] * n for _ in range(n)]     # [observer][seat] bit r = rank r, bit 8 = Bewitched
        self.kr_conf = [[False] * n for _ in range(n)]
        self.kh = [[] for _ in range(n)]               # known_hands per observer
        self.seer_from = []      # game.seer_taken_card_from
        self.seven = []          # game.seven_drawn_cards
        if deal:
            if ruleset == RULESET_RANDOM:
                self._deal_random()
            else:
                self._deal_preset()

    def _deal_preset(self):
        """game/game.py:420-477: Deck() shuffles the 76 cards, hands are pulled by type (first match)."""
        perm = self.chance.perm(76)
        self.deck = [BASE_DECK[i] for i in perm]
        self.used_cards = list(self.deck)
        for p in range(6):
            for t in PRESET_HANDS[p]:
                self.hand[p].append(self._take_like(self.deck, t))

    def clone(self):
        """copy.deepcopy(game) as CFRNode uses it (algorithms/deep_mccfr.py:105,137,155); the chance source is shared."""
        g = Game.__new__(Game)
        d = self.__dict__
        for k, v in d.items():
            if k == "chance":
                g.chance = v
            elif k == "kh":
                g.kh = [[h.copy() for h in lst] for lst in v]
            elif isinstance(v, list):
                g.__dict__[k] = [list(x) if isinstance(x, list) else x for x in v]
            else:
                g.__dict__[k] = v
        return g

    def knowledge(self):
        """Canonical dump of every observer's beliefs (for comparison with the reference's Agent.known_*)."""
        return [(tuple(self.kr_mask[o]), tuple(self.kr_conf[o]),
                 tuple((h.pid, h.conf, h.wizard, tuple(h.cards)) for h in self.kh[o])) for o in range(6)]

    # ------------------------------------------------------------------ determinisation (CFR only)
    def sample_private_information(self, viewer, role_sample=True):
        """Game.sample_private_information (game/game.py:215-242) with get_unknown_cards (:183-213), sample_deck
        (:245-262), sample_cards_for_opponent (:264-280), sample_roles_for_opponent (:283-295),
        remove_role_* (:298-310), refresh_roles_after_sampling_roles (:339-357).  Tier A/B: no warrants/blackmails."""
        ch = self.chance
        for hk in self.kh[viewer]:
            hk.used = (hk.conf - 1) * 0.2 > ch.uniform()
        # get_unknown_cards
        unknown = list(self.used_cards)
        for p in range(6):
            for c in self.bld[p]:
                self._remove_like(unknown, ctype(c))
        for p in range(6):
            for c in self.mus[p]:
                self._remove_like(unknown, ctype(c))
        for c in self.hand[viewer]:
            self._remove_like(unknown, ctype(c))
        for hk in self.kh[viewer]:
            if hk.used:
                for c in hk.cards:
                    self._remove_like(unknown, ctype(c))
        # sample_deck
        n = len(self.deck)
        lk = next((h for h in self.kh[viewer] if h.pid == -1 and h.used), None)
        self.deck = []
        if lk is not None:
            k = min(len(lk.cards), n)
            self.deck += lk.cards[:k]
            n -= k
        perm = ch.perm(len(unknown))
        unknown = [unknown[i] for i in perm]
        for _ in range(n):
            if unknown:
                self.deck.append(unknown.pop(0))
        # sample_warrants_and_blackmails (:321-336): re-roll which flagged rank carries the real one
        for arr in (self.blackmail, self.warrant):
            keys = [r for r in range(8) if arr[r]]
            if keys:
                perm = ch.perm(len(keys))
                real = keys[perm[0]]
                for r in keys:
                    arr[r] = 1 if r == real else 2
        # roles
        kr = list(self.kr_mask[viewer])
        conf = list(self.kr_conf[viewer])

        def remove_role(bit):
            for q in range(6):
                if not conf[q]:
                    kr[q] &= ~(1 << bit)
        if role_sample:
            r = self.role[self.player]
            if r == ROLE_BEWITCHED:
                remove_role(8)
            elif r != ROLE_NONE:
                for x in range(r + 1):
                    remove_role(x)
        for p in range(6):
            if p != viewer:
                # sample_cards_for_opponent
                hk = next((h for h in self.kh[viewer] if h.pid == p and h.used), None)
                n = len(self.hand[p])
                self.hand[p] = []
                if hk is not None:
                    k = min(len(hk.cards), n)
                    self.hand[p] += hk.cards[:k]
                    n -= k
                for _ in range(n):
                    if unknown:
                        self.hand[p].append(unknown.pop(0))
            if role_sample and p != viewer and p != self.player and self.state != 0:
                cand = [x for x in range(9) if kr[p] >> x & 1]
                # dict order: possible_roles is built rank-ascending; {-1: "Bewitched"} is always a singleton
                if cand:
                    x = cand[ch.randbelow(len(cand))]
                    self.role[p] = ROLE_BEWITCHED if x == 8 else x
                    remove_role(x)
                else:
                    self.role[p] = next(x for x in range(8) if x not in self.used_roles)
        if role_sample and self.state != 0:
            self._refresh_used_roles()

    @staticmethod
    def _remove_like(cards, t):
        for i, c in enumerate(cards):
            if ctype(c) == t:
                del cards[i]
                return

    def _deal_random(self):
        """Game.set_random_game, game/game.py:491-520.  Chance mapping: random.sample(uniques, 14) -> first 14 of
        perm(24); Deck() shuffle -> perm(66); random.choice per rank -> randbelow(3); random.shuffle(order) -> perm(6);
        random.randint(0, 5) -> randbelow(6)."""
        ch = self.chance
        pu = ch.perm(24)
        base = BASE_DECK[:52] + [UNIQUE_CARDS[i] for i in pu[:14]]
        perm = ch.perm(66)
        self.deck = [base[i] for i in perm]
        self.used_cards = list(self.deck)
        for _ in range(4):
            for p in range(6):
                self.hand[p].append(self.deck.pop(0))
        self.variant = [ch.randbelow_game(3) for _ in range(8)]
        po = ch.perm(6)
        self.order = [[0, 1, 2, 3, 4, 5][i] for i in po]
        self.crown = ch.randbelow_game(6)

    # ------------------------------------------------------------------ list helpers (game/deck.py)
    @staticmethod
    def _take_like(cards, t):
        """Deck.get_a_card_like_it, game/deck.py:49-55: first card of that type, fabricated if absent."""
        for i, c in enumerate(cards):
            if ctype(c) == t:
                return cards.pop(i)
        return t

    @staticmethod
    def _has(cards, t):
        for c in cards:
            if ctype(c) == t:
                return True
        return False

    def _reshuffle_if_empty(self):
        """game/option_functions.py:564-570."""
        if not self.deck:
            if not self.discard:
                return
            perm = self.chance.perm(len(self.discard))
            self.deck = [self.discard[i] for i in perm]
            self.discard = []

    def _draw_to(self, dst):
        """reshuffle_deck_if_empty + dst.add_card(deck.draw_card()); 'Deck Empty' is dropped (game/deck.py:57-70)."""
        self._reshuffle_if_empty()
        if self.deck:
            dst.append(self.deck.pop(0))

    def name(self, p):
        r = self.role[p]
        if r < 8:
            return r * 3 + self.variant[r]
        return NAME_NONE if r == ROLE_NONE else NAME_BEWITCHED

    def player_from_rank(self, rank):
        """game/game.py:403-412 (rank -1 -> the Bewitched seat)."""
        want = ROLE_BEWITCHED if rank == -1 else rank
        for p in range(6):
            if self.role[p] == want:
                return p
        return None

    def _rank_or_raise(self, p):
        r = self.role[p]
        if r == ROLE_NONE:
            raise OracleError("role_to_role_id[None]")
        return -1 if r == ROLE_BEWITCHED else r

    def _prop_rank(self, p):
        """role_properties[role_to_role_id[role]] -- KeyError for None / Bewitched."""
        r = self.role[p]
        if r >= 8:
            raise OracleError("role_properties[%d]" % r)
        return r

    # ------------------------------------------------------------------ round machine
    def setup_round(self):
        """game/game.py:144-171."""
        for r in range(8):
            self.dead[r] = False
            self.warrant[r] = 0
            self.possessed[r] = False
            self.robbed[r] = False
            self.blackmail[r] = 0
        self.used_roles = []
        perm = self.chance.perm(8)                 # shuffle of list(roles.items()); 6 players -> one face-down pop()
        facedown = perm[7]
        self.rtc = [r for r in range(8) if r != facedown]
        c = self.crown
        self.order = self.order[c:] + self.order[:c]
        # fresh GameState(state=0, player_id=order[0])
        self.state = 0
        self.player = self.order[0]
        self.done = 0
        self.n_trade = self.n_nontrade = 0
        self.next_mode = NEXT_NONE
        self.next_player = 0
        for p in range(6):
            # game/agent.py:100-114
            keep = []
            for hk in self.kh[p]:
                hk.conf -= 1
                hk.wizard = False
                hk.used = False
                if hk.conf != 0:
                    keep.append(hk)
            self.kh[p] = keep
            for q in range(6):
                self.kr_mask[p][q] = 0
                self.kr_conf[p][q] = False
        # the acting wizard's entry loses its wizard flag (game/agent.py:104); keep the copy inert
        self.wiz_target = 0xFF
        self.wiz_cards = []

    def _refresh_used_roles(self):
        """game/game.py:349-357."""
        self.used_roles = sorted(self._rank_or_raise(p) for p in range(6))

    def setup_next_player(self, current=None):
        """game/game.py:391-401."""
        if self.state == 0:
            self._refresh_used_roles()
            self.state = 1
            nxt = self.player_from_rank(self.used_roles[0])
        elif current is not None:
            self.state = 1
            r = self._rank_or_raise(current)
            if r not in self.used_roles:
                raise OracleError("used_roles.index")
            i = self.used_roles.index(r) + 1
            if i >= len(self.used_roles):
                raise OracleError("used_roles[i+1]")
            nxt = self.player_from_rank(self.used_roles[i])
            self.done = 0
            self.n_trade = self.n_nontrade = 0
        else:
            raise OracleError("No current player and not in rolepick state")
        if nxt is None:
            raise OracleError("get_player_from_role_id -> None")
        self.player = nxt

    def _is_last_round(self):
        """game/game.py:173-181."""
        if not self.ending:
            for p in range(6):
                if len(self.bld[p]) == 7:
                    self.ending = True
                    self.first7[p] = True

    def count_points(self, p):
        """game/agent.py:116-143."""
        pts = 0
        well = self._has(self.bld[p], 31)
        for c in self.bld[p]:
            pts += ccost(c)
            if ctype(c) in (18, 23):
                pts += 2
            if well and csuit(c) == SUIT_UNIQUE:
                pts += 1
        if len(self.bld[p]) >= 7:
            pts += 2
        if self.first7[p]:
            pts += 4
        pts += len(self.mus[p])
        if self._has(self.bld[p], 37):
            pts += self.gold[p]
        if self._has(self.bld[p], 39):
            pts += len(self.hand[p])
        return pts

    def _check_game_ending(self):
        """game/game.py:359-368."""
        if self.ending:
            self.points = [self.count_points(p) for p in range(6)]
            self.terminal = True
            self.winner = self.points.index(max(self.points))
            return True
        return False

    def _move_crown(self, target):
        """game/option_functions.py:625-631, :588-595."""
        self.crown = target
        for p in range(6):
            if self._has(self.bld[p], 32):
                self.gold[p] += 1
                break

    # ------------------------------------------------------------------ knowledge (CFR only)
    def _confirm_role(self, revealed):
        """game/option_functions.py:608-622; the elif-filter is a no-op."""
        r = self._rank_or_raise(revealed)
        for obs in range(6):
            self.kr_mask[obs][revealed] = (1 << 8) if r == -1 else (1 << r)
            self.kr_conf[obs][revealed] = True

    # ------------------------------------------------------------------ enumeration
    def build_limit(self, p):
        """game/agent.py:87-98."""
        nm = self.name(p)
        if nm == ARCHITECT:
            return 3
        if nm == SCHOLAR:
            return 2
        if nm in (BISHOP, NAVIGATOR):
            return 0
        return 1

    def options(self):
        """Agent.get_options, game/agent.py:50-83."""
        p = self.player
        st = self.state
        if st == 0:
            # game/agent_functions.py:13-14
            return [D(K["role_pick"], p, rank=r) for r in self.rtc]
        role = self.role[p]
        if role == ROLE_NONE:
            raise OracleError("role_to_role_id[None]")
        nm = self.name(p)
        if role == ROLE_BEWITCHED or not self.dead[role]:
            if st == 1:
                # game/agent_functions.py:16-17
                o = [D(K["gold_or_card"], p, named=NAMED["gold"])]
                if len(self.deck) > 1:
                    o.append(D(K["gold_or_card"], p, named=NAMED["card"]))
                return o
            if st == 2:
                return self._keep_options(p)
            if st == 3:
                # game/agent_functions.py:35-38
                if self.blackmail[self._prop_rank(p)]:
                    return [D(K["blackmail_response"], p, named=NAMED["pay"]),
                            D(K["blackmail_response"], p, named=NAMED["not_pay"])]
                return [D(K["empty_option"], p)]
            if st == 4:
                # game/agent_functions.py:40-41
                q = self.next_player
                return [D(K["reveal_blackmail_as_blackmailer"], p, target=q, named=NAMED["reveal"]),
                        D(K["reveal_blackmail_as_blackmailer"], p, target=q, named=NAMED["not_reveal"])]
            if st == 6:
                # game/agent_functions.py:150-153
                if self.gold[p] > 0:
                    return [D(K["graveyard"], p)]
                return [D(K["empty_option"], p)]
            if st == 7:
                # game/agent_functions.py:43-44
                q = self.next_player
                return [D(K["reveal_warrant_as_magistrate"], p, target=q, named=NAMED["reveal"]),
                        D(K["reveal_warrant_as_magistrate"], p, target=q, named=NAMED["not_reveal"])]
            if nm == WITCH:
                # game/agent_functions.py:236-242: every rank > 0 of game.roles
                return [D(K["bewitching"], p, rank=r) for r in range(1, 8)]
            if not self.possessed[self._prop_rank(p)]:
                if st == 5:
                    return self._main_round_options(p)
                if st == 8:
                    return self._seer_give_back_options(p)
                if st == 9:
                    return self._scholar_give_back_options(p)
                if st == 10:
                    return self._wizard_take_options(p)
                return None   # the reference falls off the if-chain and returns None
            return [D(K["finish_round"], p, next_witch=1, crown=nm in (KING, PATRICIAN))]
        if nm == EMPEROR and not (self.done & DM_CHARACTER):
            return self._emperor_options(p, dead=True)
        return [D(K["finish_round"], p, next_witch=0, crown=nm in (KING, PATRICIAN))]

    def _keep_options(self, p):
        """game/agent_functions.py:19-33."""
        jd = self.jd[p]
        if self._has(self.bld[p], 20):
            return [D(K["which_card_to_keep"], p, a=ctype(jd[i]), b=ctype(jd[j]))
                    for i in range(len(jd)) for j in range(i + 1, len(jd))]
        o, seen = [], set()
        for c in jd:
            t = ctype(c)
            if t not in seen:
                seen.add(t)
                o.append(D(K["which_card_to_keep"], p, a=t))
        return o

    def _build_cost(self, p, c):
        """game/agent_functions.py:111-114: Factory (35) makes uniques dearer (cost -= -1)."""
        cost = ccost(c)
        if self._has(self.bld[p], 35) and csuit(c) == SUIT_UNIQUE:
            cost += 1
        return cost

    def _build_options(self, p):
        """game/agent_functions.py:108-130."""
        nm = self.name(p)
        limit = self.build_limit(p)
        n = self.n_nontrade if nm == TRADER else self.n_trade + self.n_nontrade
        o = []
        if n < limit:
            seen = set()
            for c in self.hand[p]:
                t = ctype(c)
                replica = 0
                if self._has(self.bld[p], t) and not self.replicas[p]:
                    replica = self.replicas[p] + 1
                if self._build_cost(p, c) <= self.gold[p] and (t, replica) not in seen:
                    seen.add((t, replica))
                    o.append(D(K["build"], p, a=t, replica=replica))
        return o

    def _character_options(self, p):
        """game/agent_functions.py:156-209 and the per-role enumerators it dispatches to."""
        nm = self.name(p)
        o = []
        if not (self.done & DM_CHARACTER):
            if nm == ASSASSIN:      # :213-218
                o = [D(K["assassination"], p, rank=r) for r in range(1, 8)]
            elif nm == THIEF:       # :246-253
                o = [D(K["steal"], p, rank=r) for r in range(2, 8)]
            elif nm == SPY:         # :274-281
                o = [D(K["spy"], p, target=q, named=8 + s) for q in range(6) if q != p for s in range(5)]
            elif nm == MAGICIAN:    # :284-296
                o = [D(K["magic_hand_change"], p, target=q) for q in range(6) if q != p]
                n = len(self.hand[p])
                for r in range(1, n + 1):
                    total = comb(n, r)
                    step = max(py_round_div100(total), 1)
                    cnt = (total + step - 1) // step
                    o += [D(K["discard_and_draw"], p, r=r, j=j) for j in range(cnt)]
            elif nm == WIZARD:      # :298-308
                o = [D(K["look_at_hand"], p, target=q) for q in range(6) if q != p and self.hand[q]]
            elif nm == KING:        # :364-366
                o = [D(K["take_crown_king"], p)]
            elif nm == BISHOP:      # :389-391
                o = [D(K["bishop"], p)]
            elif nm == ABBOT:       # :422-430
                n = sum(1 for c in self.hand[p] if csuit(c) == SUIT_RELIGION)
                if n > 0:
                    o = [D(K["abbot_gold_or_card"], p, count=k, r=n) for k in range(n + 1)]
            elif nm == MERCHANT:    # :438-440
                o = [D(K["merchant"], p)]
            elif nm == ALCHEMIST:   # :442-444
                o = []
            elif nm == ARCHITECT:   # :451-452
                o = [D(K["architect"], p)]
            elif nm == NAVIGATOR:   # :454-455
                o = [D(K["navigator_gold_card"], p, named=NAMED["4gold"]),
                     D(K["navigator_gold_card"], p, named=NAMED["4card"])]
            elif nm == WARLORD:     # :473-482
                seen = set()
                for q in range(6):
                    if len(self.bld[q]) < 7:
                        for c in self.bld[q]:
                            t = ctype(c)
                            if ccost(c) - 1 <= self.gold[p] and t != 17 and self.name(q) != BISHOP:
                                if (q, t) not in seen:
                                    seen.add((q, t))
                                    o.append(D(K["warlord_desctruction"], p, target=q, a=t))
            elif nm == MAGISTRATE:  # :221-234
                targets = list(range(1, 8))
                for real in targets:
                    for i in range(len(targets)):
                        for j in range(i + 1, len(targets)):
                            a, b = targets[i], targets[j]
                            if real != a and real != b:
                                o.append(D(K["magistrate_warrant"], p, rank=real, named=a, count=b))
            elif nm == BLACKMAILER:  # :255-272 (possessed ranks cannot be blackmailed)
                targets = [r for r in range(2, 8) if not self.possessed[r]]
                for i in range(len(targets)):
                    for j in range(i + 1, len(targets)):
                        o.append(D(K["blackmail"], p, rank=targets[i], named=targets[j]))
                        o.append(D(K["blackmail"], p, rank=targets[j], named=targets[i]))
            elif nm == SEER:        # :328-329
                o = [D(K["seer"], p)]
            elif nm == EMPEROR:     # :368-382
                o = self._emperor_options(p, dead=False)
            elif nm == PATRICIAN:   # :384-386
                o = [D(K["take_crown_pat"], p)]
            elif nm == CARDINAL:    # :393-419
                o = self._cardinal_options(p)
            elif nm == TRADER:      # :446-448
                o = [D(K["trader"], p)]
            elif nm == SCHOLAR:     # :457-460
                o = [D(K["scholar"], p)] if self.deck else []
            elif nm == MARSHAL:     # :484-492
                seen = set()
                for q in range(6):
                    if q != p and len(self.bld[q]) < 7:
                        for c in self.bld[q]:
                            t = ctype(c)
                            if (ccost(c) <= self.gold[p] and ccost(c) <= 3 and not self._has(self.bld[p], t) and t != 17
                                    and self.name(q) != BISHOP and (q, t) not in seen):
                                seen.add((q, t))
                                o.append(D(K["marshal_steal"], p, target=q, a=t))
            elif nm == DIPLOMAT:    # :494-504
                seen = set()
                for q in range(6):
                    if q != p and len(self.bld[q]) < 7:
                        for e in self.bld[q]:
                            for own in self.bld[p]:
                                te, to = ctype(e), ctype(own)
                                if (ccost(e) - ccost(own) <= self.gold[p] and te != 17 and self.name(q) != BISHOP
                                        and not self._has(self.bld[p], te) and (q, te, to) not in seen):
                                    seen.add((q, te, to))
                                    o.append(D(K["diplomat_exchange"], p, target=q, a=te, b=to))
            elif nm in (NAME_NONE, NAME_BEWITCHED):
                pass
        if nm == ABBOT and not (self.done & DM_BEGGED):       # :199-202, :432-435
            o.append(D(K["abbot_beg"], p))
        if nm in (WARLORD, MARSHAL, DIPLOMAT) and not (self.done & DM_TAKE_GOLD):   # :204-207, :506-509
            o.append(D(K["take_gold_for_war"], p))
        return o

    def _emperor_options(self, p, dead):
        """game/agent_functions.py:368-382."""
        o = []
        for q in range(6):
            if q != p:
                if self.hand[q] and not dead:
                    o.append(D(K["give_crown"], p, target=q, named=NAMED["card"]))
                if self.gold[q] and not dead:
                    o.append(D(K["give_crown"], p, target=q, named=NAMED["gold"]))
                if (not self.gold[q] and not self.hand[q]) or dead:
                    o.append(D(K["give_crown"], p, target=q, named=NAMED_NOTHING))
        return o

    def _cardinal_options(self, p):
        """game/agent_functions.py:393-419.  j = position of the kept combination among the thinned ones
        (range(0, C, max(round(C/100), 1))); `build` carries the factory flag, `count` the number of cards to give."""
        o = []
        hand = self.hand[p]
        factory_owned = self._has(self.bld[p], 35)
        for q in range(6):
            for c in hand:
                t = ctype(c)
                cost = ccost(c)
                factory = False
                replica = 0
                if factory_owned and csuit(c) == SUIT_UNIQUE:
                    cost += 1
                    factory = True
                if self._has(self.bld[p], t) and not self.replicas[p]:
                    replica = self.replicas[p] + 1
                if cost <= self.gold[q]:
                    ex = max(self.gold[q] - cost, 0)
                    if len(hand) - 1 >= ex:
                        n_other = sum(1 for x in hand if ctype(x) != t)
                        total = comb(n_other, ex)
                        step = max(py_round_div100(total), 1)
                        for j in range((total + step - 1) // step):
                            o.append(D(K["cardinal_exchange"], p, target=q, a=t, replica=replica, build=factory, count=ex, j=j))
        return o

    def _seer_give_back_options(self, p):
        """game/agent_functions.py:332-361: the enumeration itself shuffles (3 random fill-ins per card and position).
        Descriptor: card type + 1 per player of seer_taken_card_from, in that order (0 = zip() ran out)."""
        k = len(self.seer_from)
        cards = self.hand[p]
        o = []
        for pos in range(k):
            for card in cards:
                remaining = [c for c in cards if ctype(c) != ctype(card)]
                for _ in range(3):
                    perm = self.chance.perm(len(remaining))
                    remaining = [remaining[i] for i in perm]
                    row = list(remaining[:k - 1])
                    row.insert(pos, card)
                    o.append(D_handout(p, [ctype(c) for c in row[:k]]))
        return o

    def _scholar_give_back_options(self, p):
        """game/agent_functions.py:462-470: copy() is shallow, so get_a_card_like_it removes from the very list that is
        being iterated: every call keeps shrinking game.seven_drawn_cards and every option shares what is left."""
        o = []
        i = 0
        while i < len(self.seven):
            t = ctype(self.seven[i])
            self._remove_like(self.seven, t)
            o.append(D(K["scholar_card_pick"], p, a=t))
            i += 1
        return o

    def _main_round_options(self, p):
        """game/agent_functions.py:133-147 (order of concatenation is observable)."""
        o = self._build_options(p)
        o += self._character_options(p)
        bld = self.bld[p]
        if self._has(bld, 21) and self.gold[p] >= 2 and not (self.done & DM_SMITHY):      # :54-57
            o.append(D(K["smithy_choice"], p))
        if self._has(bld, 22) and not (self.done & DM_LAB):                               # :59-65 (no dedupe)
            o += [D(K["laboratory_choice"], p, a=ctype(c)) for c in self.hand[p]]
        if not (self.done & DM_MAGIC_SCHOOL) and self._has(bld, 25):                      # :67-74
            o += [D(K["magic_school_choice"], p, named=8 + s) for s in range(5)]
        if self._has(bld, 27):                                                            # :76-83 (no dedupe)
            for q in range(6):
                if q != p:
                    o += [D(K["weapon_storage_choice"], p, target=q, a=ctype(c)) for c in self.bld[q]]
        if self._has(bld, 29) and self.lighthouse[p]:                                     # :85-94
            seen = set()
            for c in self.deck:
                t = ctype(c)
                if t not in seen:
                    seen.add(t)
                    o.append(D(K["lighthouse_choice"], p, a=t))
        if self._has(bld, 34) and not (self.done & DM_MUSEUM):                            # :96-105
            seen = set()
            for c in self.hand[p]:
                t = ctype(c)
                if t not in seen:
                    seen.add(t)
                    o.append(D(K["museum_choice"], p, a=t))
        o.append(D(K["finish_round"], p, next_witch=0, crown=0))
        return o

    def _wizard_take_options(self, p):
        """game/agent_functions.py:310-326 (replica leaks across iterations)."""
        if self.wiz_target == 0xFF:
            raise OracleError("target_hand is None")
        q = self.wiz_target
        o, seen = [], set()
        replica = 0
        for c in self.wiz_cards:
            t = ctype(c)
            key = (t, 0, 0)
            if key not in seen:
                seen.add(key)
                o.append(D(K["take_from_hand"], p, target=q, a=t, build=0))
            if self._has(self.bld[p], t):
                replica = self.replicas[p] + 1
            key = (t, 1, replica)
            if self._build_cost(p, c) <= self.gold[p] and key not in seen:
                seen.add(key)
                o.append(D(K["take_from_hand"], p, target=q, a=t, build=1, replica=replica))
        if not o:
            return [D(K["empty_option"], p)]
        return o

    # ------------------------------------------------------------------ transition
    def apply(self, d):
        """option.carry_out, game/option.py:118-122.  Returns True when the game ended (winner set)."""
        k = d_kind(d)
        won = self._APPLY[k](self, d)
        self._is_last_round()
        return bool(won)

    def _to5(self, p):
        self.state = 5
        self.player = p

    def _restore_next(self):
        """game.gamestate = game.gamestate.next_gamestate."""
        if self.next_mode == NEXT_NONE:
            raise OracleError("next_gamestate is None")
        self.state = 5
        self.player = self.next_player
        if self.next_mode == NEXT_RESET_CA:
            self.done = DM_CHARACTER
            self.n_trade = self.n_nontrade = 0
        elif self.next_mode == NEXT_EMPTY:
            self.done = 0
            self.n_trade = self.n_nontrade = 0
        self.next_mode = NEXT_NONE
        self.next_player = 0

    def _a_role_pick(self, d):
        """game/option_functions.py:6-30."""
        p, r = d_perp(d), d_rank(d)
        self.role[p] = r
        self.rtc.remove(r)
        rtc_mask = sum(1 << x for x in self.rtc)
        me = self.order.index(p)
        for q in range(6):
            if q != p:
                if self.order.index(q) < me:
                    self.kr_mask[p][q] = 0xFF & ~rtc_mask & ~(1 << r)
                else:
                    self.kr_mask[p][q] = rtc_mask
        if p != self.order[-1]:
            self.state = 0
            self.player = self.order[me + 1]
        else:
            self.setup_next_player()

    def _a_gold_or_card(self, d):
        """game/option_functions.py:33-55."""
        p = d_perp(d)
        self._confirm_role(p)
        if self.robbed[self._prop_rank(p)]:
            thief = self.player_from_rank(1)
            if thief is None:
                raise OracleError("no rank-1 player")
            self.gold[thief] += self.gold[p]
            self.gold[p] = 0
        if d_named(d) == NAMED["gold"]:
            self.gold[p] += 2
            self.state = 3
        else:
            for _ in range(3 if self._has(self.bld[p], 16) else 2):
                self._draw_to(self.jd[p])
            self.state = 2
        self.player = p

    def _a_keep(self, d):
        """game/option_functions.py:58-66."""
        p = d_perp(d)
        self.hand[p].append(self._take_like(self.jd[p], d_a(d)))
        if d_b(d) >= 0:
            self.hand[p].append(self._take_like(self.jd[p], d_b(d)))
        self.deck += self.jd[p]
        self.jd[p] = []
        self.state = 3
        self.player = p

    def _a_empty(self, d):
        """game/option_functions.py:68-69; the three producers are game/agent_functions.py:38, :153, :325."""
        if self.state == 3:
            # fresh GameState(state=5, player_id=agent.id)
            self.state = 5
            self.player = d_perp(d)
            self.done = 0
            self.n_trade = self.n_nontrade = 0
            self.next_mode = NEXT_NONE
            self.next_player = 0
        else:
            self._restore_next()

    def _a_build(self, d):
        """game/option_functions.py:102-127."""
        p, t = d_perp(d), d_a(d)
        c = self._take_like(self.hand[p], t)
        self.bld[p].append(c)
        if self.name(p) != ALCHEMIST:
            self.gold[p] -= ccost(c)
        rep = d_replica(d)
        if rep:
            self.replicas[p] = rep
        if csuit(c) == SUIT_TRADE:
            self.n_trade += 1
        else:
            self.n_nontrade += 1
        if t == 29:
            self.lighthouse[p] = True
        if self.warrant[self._prop_rank(p)] == 0:
            self._to5(p)
        else:   # :121-127 any warrant, real or fake, interrupts for the Magistrate
            m = self.player_from_rank(0)
            if m is None:
                raise OracleError("no rank-0 player")
            self.warrant_building = t
            self.state = 7
            self.player = m
            self.next_player = p
            self.next_mode = NEXT_ALIAS

    def _a_smithy(self, d):
        """game/option_functions.py:131-138 (cards go to just_drawn_cards)."""
        p = d_perp(d)
        self.gold[p] -= 2
        for _ in range(3):
            self._draw_to(self.jd[p])
        self._to5(p)
        self.done |= DM_SMITHY

    def _a_lab(self, d):
        """game/option_functions.py:140-145."""
        p = d_perp(d)
        self.discard.append(self._take_like(self.hand[p], d_a(d)))
        self.gold[p] += 1
        self._to5(p)
        self.done |= DM_LAB

    def _a_magic_school(self, d):
        """game/option_functions.py:147-153: the card is removed and re-appended with the new suit."""
        p = d_perp(d)
        self._take_like(self.bld[p], 25)
        s = d_named(d) - 8
        self.bld[p].append(25 if s == SUIT_UNIQUE else 40 + s)
        self._to5(p)
        self.done |= DM_MAGIC_SCHOOL

    def _a_museum(self, d):
        """game/option_functions.py:161-165."""
        p = d_perp(d)
        self.mus[p].append(self._take_like(self.hand[p], d_a(d)))
        self._to5(p)
        self.done |= DM_MUSEUM

    def _a_weapon_storage(self, d):
        """game/option_functions.py:167-171."""
        p, q = d_perp(d), d_target(d)
        self.discard.append(self._take_like(self.bld[p], 27))
        self.discard.append(self._take_like(self.bld[q], d_a(d)))
        self._to5(p)

    def _a_lighthouse(self, d):
        """game/option_functions.py:173-180."""
        p = d_perp(d)
        self.kh[p].append(HandKnowledge(-1, self.deck))
        self.hand[p].append(self._take_like(self.deck, d_a(d)))
        self.lighthouse[p] = False
        perm = self.chance.perm(len(self.deck))
        self.deck = [self.deck[i] for i in perm]
        self._to5(p)

    def _a_graveyard(self, d):
        """game/option_functions.py:183-187."""
        p = d_perp(d)
        if not self.discard:
            raise OracleError("pop from empty discard")
        self.bld[p].append(self.discard.pop())
        self.gold[p] -= 1
        self._restore_next()

    def _a_finish(self, d):
        """finish_main_sequnce_actions, game/option_functions.py:189-243."""
        p = d_perp(d)
        pr = self._prop_rank(p)
        if not self.dead[pr]:
            if self._has(self.bld[p], 28) and not self.hand[p]:
                for _ in range(2):
                    self._draw_to(self.jd[p])
            if self._has(self.bld[p], 30) and not self.hand[p]:
                self.gold[p] += 1
        if d_crown(d):
            self._confirm_role(p)
            self._move_crown(p)
        elif self.dead[pr]:
            self._confirm_role(p)
        if d_next_witch(d):
            w = self.player_from_rank(0)
            if w is None:
                raise OracleError("no rank-0 player")
            self.state = 5
            self.player = w
            self.role[w] = self.role[p]
            self.possessed[pr] = False
            self.role[p] = ROLE_BEWITCHED
            for obs in range(6):
                if obs != w:
                    self.kr_mask[obs][w] = 1 << pr
                if obs != p:
                    self.kr_mask[obs][p] = 1 << 8
            self.done = 0
            self.n_trade = self.n_nontrade = 0
            return False
        if not self.used_roles:
            raise OracleError("used_roles[-1]")
        if self.used_roles[-1] == pr:
            if self._check_game_ending():
                return True
            self.setup_round()
        else:
            self.setup_next_player(current=p)
        return False

    def _a_mark(self, d, arr):
        p = d_perp(d)
        arr[d_rank(d)] = True
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_assassination(self, d):
        """game/option_functions.py:245-249."""
        self._a_mark(d, self.dead)

    def _a_steal(self, d):
        """game/option_functions.py:265-269."""
        self._a_mark(d, self.robbed)

    def _a_bewitching(self, d):
        """game/option_functions.py:259-262."""
        p = d_perp(d)
        self.possessed[d_rank(d)] = True
        self.witch[p] = True
        self.setup_next_player(current=p)

    def _a_spy(self, d):
        """game/option_functions.py:278-288."""
        p, q, s = d_perp(d), d_target(d), d_named(d) - 8
        n = sum(1 for c in self.hand[q] if csuit(c) == s)
        steal = min(n, self.gold[q])
        self.gold[p] += steal
        self.gold[q] -= steal
        self._draw_to(self.hand[p])
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_magic_hand_change(self, d):
        """game/option_functions.py:291-293."""
        p, q = d_perp(d), d_target(d)
        self.hand[p], self.hand[q] = self.hand[q], self.hand[p]
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_discard_and_draw(self, d):
        """game/option_functions.py:295-300: ignores the option's cards, removes while iterating."""
        p = d_perp(d)
        h = self.hand[p]
        i = 0
        while i < len(h):
            card = h[i]
            self.deck.append(self._take_like(h, ctype(card)))
            i += 1
        for _ in range(len(h)):
            self._draw_to(h)
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_look(self, d):
        """game/option_functions.py:305-310."""
        p, q = d_perp(d), d_target(d)
        self.kh[p].append(HandKnowledge(q, self.hand[q], wizard=True))
        self.wiz_target = q
        self.wiz_cards = list(self.hand[q])
        self.state = 10
        self.player = p
        self.done |= DM_CHARACTER
        self.next_player = p
        self.next_mode = NEXT_ALIAS

    def _a_take_from_hand(self, d):
        """game/option_functions.py:312-328."""
        p, q, t = d_perp(d), d_target(d), d_a(d)
        hk = next((h for h in self.kh[p] if h.wizard), None)
        self.hand[p].append(self._take_like(self.hand[q], t))
        if d_build(d):
            rep = sum(1 for c in self.bld[p] if ctype(c) == t)
            # carry_out_building with the rewritten replica attribute
            d2 = D(K["build"], p, a=t, replica=rep)
            self._a_build(d2)
        if hk is not None:
            self._take_like(hk.cards, t)
        self._take_like(self.wiz_cards, t)
        self._restore_next()

    def _a_take_crown_king(self, d):
        """game/option_functions.py:354-363."""
        p = d_perp(d)
        self.gold[p] += sum(1 for c in self.bld[p] if csuit(c) == SUIT_LORD)
        if not self.witch[p]:
            self._move_crown(p)
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_suit_gold(self, d, suit, extra=0):
        p = d_perp(d)
        self.gold[p] += sum(1 for c in self.bld[p] if csuit(c) == suit) + extra
        self._to5(p)

    def _a_bishop(self, d):
        """game/option_functions.py:397-403."""
        self._a_suit_gold(d, SUIT_RELIGION)
        self.done |= DM_CHARACTER

    def _a_merchant(self, d):
        """game/option_functions.py:442-449."""
        self._a_suit_gold(d, SUIT_TRADE, 1)
        self.done |= DM_CHARACTER

    def _a_take_gold_for_war(self, d):
        """game/option_functions.py:553-559."""
        self._a_suit_gold(d, SUIT_WAR)
        self.done |= DM_TAKE_GOLD

    def _a_abbot(self, d):
        """game/option_functions.py:405-412 (the option carries its own gold/card list: r entries, count cards)."""
        p = d_perp(d)
        n = d_r(d)
        k = d_count(d)
        self.gold[p] += n - k
        for _ in range(k):
            self._draw_to(self.hand[p])
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_abbot_beg(self, d):
        """game/option_functions.py:414-420."""
        p = d_perp(d)
        rich = max(range(6), key=lambda i: (self.gold[i], -i))
        self.gold[rich] -= 1
        a = self.player_from_rank(4)
        if a is None:
            raise OracleError("no rank-4 player")
        self.gold[a] += 1
        self._to5(p)
        self.done |= DM_BEGGED

    def _a_architect(self, d):
        """game/option_functions.py:464-471."""
        p = d_perp(d)
        for _ in range(2):
            self._draw_to(self.hand[p])
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_navigator(self, d):
        """game/option_functions.py:473-483."""
        p = d_perp(d)
        if d_named(d) == NAMED["4card"]:
            for _ in range(4):
                self._draw_to(self.hand[p])
        else:
            self.gold[p] += 4
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_warlord(self, d):
        """game/option_functions.py:517-535 (+ :573-586, :597-606)."""
        p, q, t = d_perp(d), d_target(d), d_a(d)
        # option.attributes['choice'] is the first building of that type of the target at enumeration time
        c = next((x for x in self.bld[q] if ctype(x) == t), t)
        self.gold[p] -= ccost(c) - 1
        self.discard.append(self._take_like(self.bld[q], t))
        if sum(1 for x in self.bld[q] if ctype(x) == t) > 1:
            self.replicas[q] -= 1
        if t == 34:
            self.discard += self.mus[q]
            self.mus[q] = []
        if t == 29 and self.lighthouse[q]:
            self.lighthouse[q] = False
            self.lighthouse[p] = True
        self._to5(p)
        self.done |= DM_CHARACTER
        owner = next((x for x in range(6) if self._has(self.bld[x], 24)), None)
        if owner is not None and owner != p:
            self.state = 6
            self.player = owner
            self.next_player = p
            self.next_mode = NEXT_RESET_CA

    # ---- tier C transitions -------------------------------------------------------------------------
    def _a_warranting(self, d):
        """game/option_functions.py:251-257."""
        p = d_perp(d)
        self.warrant[d_rank(d)] = 1
        self.warrant[d_named(d)] = 2
        self.warrant[d_count(d)] = 2
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_magistrate_reveal(self, d):
        """game/option_functions.py:94-100."""
        p, q = d_perp(d), d_target(d)
        if d_named(d) == NAMED["reveal"] and self.warrant[self._prop_rank(q)] == 1:
            t = self.warrant_building
            self.bld[p].append(self._take_like(self.bld[q], t))
            self.gold[q] += COST_OF_TYPE[t]
            self.warrant = [0] * 8
        self._restore_next()

    def _a_blackmail(self, d):
        """game/option_functions.py:271-276."""
        p = d_perp(d)
        self.blackmail[d_rank(d)] = 1
        self.blackmail[d_named(d)] = 2
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_blackmail_response(self, d):
        """game/option_functions.py:71-82 (int(gold/2) truncates toward zero)."""
        p = d_perp(d)
        bm = self.player_from_rank(1)
        if bm is None:
            raise OracleError("no rank-1 player")
        if d_named(d) == NAMED["pay"]:
            half = int(self.gold[p] / 2)
            self.gold[bm] += half
            self.gold[p] -= half
            self._to5(p)
        else:
            self.state = 4
            self.player = bm
            self.next_player = p
            self.next_mode = NEXT_EMPTY

    def _a_blackmail_reveal(self, d):
        """game/option_functions.py:85-92."""
        p, q = d_perp(d), d_target(d)
        if d_named(d) == NAMED["reveal"] and self.blackmail[self._prop_rank(q)] == 1:
            self.gold[p] += self.gold[q]
            self.gold[q] = 0
            self.blackmail = [0] * 8
        self._restore_next()

    def _a_seer(self, d):
        """game/option_functions.py:330-341."""
        p = d_perp(d)
        self.seer_from = []
        for q in range(6):
            if q != p and self.hand[q]:
                perm = self.chance.perm(len(self.hand[q]))
                self.hand[q] = [self.hand[q][i] for i in perm]
                self._reshuffle_if_empty()
                self.hand[p].append(self.hand[q].pop(0))
                self.seer_from.append(q)
        self.state = 8
        self.player = p
        self.done |= DM_CHARACTER
        self.next_player = p
        self.next_mode = NEXT_ALIAS

    def _a_seer_give_back(self, d):
        """game/option_functions.py:343-350."""
        p = d_perp(d)
        for q, t in zip(self.seer_from, d_handout(d)):
            c = self._take_like(self.hand[p], t)
            self.hand[q].append(c)
            self.kh[p].append(HandKnowledge(q, [c]))
        self.seer_from = []
        self._restore_next()

    def _a_emperor(self, d):
        """game/option_functions.py:377-393."""
        p, q = d_perp(d), d_target(d)
        self.gold[p] += sum(1 for c in self.bld[p] if csuit(c) == SUIT_LORD)
        if d_named(d) == NAMED["card"]:
            perm = self.chance.perm(len(self.hand[q]))
            self.hand[q] = [self.hand[q][i] for i in perm]
            if self.hand[q]:
                self.hand[p].append(self.hand[q].pop(0))
        elif d_named(d) == NAMED["gold"]:
            self.gold[p] += 1
            self.gold[q] -= 1
        self._confirm_role(p)
        self._move_crown(q)
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_patrician(self, d):
        """game/option_functions.py:365-375."""
        p = d_perp(d)
        for _ in range(sum(1 for c in self.bld[p] if csuit(c) == SUIT_LORD)):
            self._draw_to(self.hand[p])
        if not self.witch[p]:
            self._move_crown(p)
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_cardinal(self, d):
        """game/option_functions.py:422-439 (no trade/non-trade bookkeeping, no warrant check, gold clamped at 0)."""
        p, q, t = d_perp(d), d_target(d), d_a(d)
        c = self._take_like(self.hand[p], t)
        self.bld[p].append(c)
        self.gold[p] -= ccost(c) - (1 if d_build(d) else 0)
        self.gold[p] = max(0, self.gold[p])
        if d_replica(d):
            self.replicas[p] = d_replica(d)
        k = d_count(d)
        if k:
            self.gold[q] -= k
            other = [x for x in self.hand[p] if ctype(x) != t]
            total = comb(len(other), k)
            step = max(py_round_div100(total), 1)
            for i in unrank_combination(len(other), k, d_j(d) * step):
                self.hand[q].append(self._take_like(self.hand[p], ctype(other[i])))
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_trader(self, d):
        """game/option_functions.py:455-461."""
        self._a_suit_gold(d, SUIT_TRADE)
        self.done |= DM_CHARACTER

    def _a_scholar(self, d):
        """game/option_functions.py:485-496."""
        p = d_perp(d)
        self.seven = []
        for _ in range(min(7, len(self.deck))):
            self._reshuffle_if_empty()
            c = self.deck.pop(0)
            self.hand[p].append(c)
            self.seven.append(c)
        self.state = 9
        self.player = p
        self.done |= DM_CHARACTER
        self.next_player = p
        self.next_mode = NEXT_ALIAS

    def _a_scholar_pick(self, d):
        """game/option_functions.py:498-502: puts back whatever the (shared, shrunk) unchosen list holds; the choice
        itself has no effect."""
        p = d_perp(d)
        for c in self.seven:
            self.deck.append(self._take_like(self.hand[p], ctype(c)))
        self._restore_next()
        self.seven = []

    def _steal_or_swap_tail(self, p, q, t):
        """check_if_building_is_replica / settle_museum (non-warlord branch) / settle_lighthouse (:573-606)."""
        if sum(1 for x in self.bld[q] if ctype(x) == t) > 1:
            self.replicas[q] -= 1
        if t == 34:
            self.mus[p] += self.mus[q]
            self.mus[q] = []
        if t == 29 and self.lighthouse[q]:
            self.lighthouse[q] = False
            self.lighthouse[p] = True
        self._to5(p)
        self.done |= DM_CHARACTER

    def _a_marshal(self, d):
        """game/option_functions.py:505-515."""
        p, q, t = d_perp(d), d_target(d), d_a(d)
        cost = COST_OF_TYPE[t]
        self.gold[p] -= cost
        self.gold[q] += cost
        self.bld[p].append(self._take_like(self.bld[q], t))
        self._steal_or_swap_tail(p, q, t)

    def _a_diplomat(self, d):
        """game/option_functions.py:538-551."""
        p, q, t, g = d_perp(d), d_target(d), d_a(d), d_b(d)
        money = abs(COST_OF_TYPE[t] - COST_OF_TYPE[g])
        self.gold[p] -= money
        self.gold[q] += money
        self.bld[p].append(self._take_like(self.bld[q], t))
        self.bld[q].append(self._take_like(self.bld[p], g))
        self._steal_or_swap_tail(p, q, t)

    def _a_unimplemented(self, d):
        raise NotImplementedError(KIND_NAMES[d_kind(d)])

    _APPLY = {}

    # ------------------------------------------------------------------ packing (256-byte playout record)
    def pack(self):
        """The engine's HBM record (include/citadels_b200.h `ctd_state`, 256 bytes)."""
        arena = []
        off = []
        for p in range(6):
            for lst in (self.hand[p], self.bld[p], self.mus[p], self.jd[p]):
                off.append(len(arena))
                arena += lst
        off.append(len(arena))
        arena += self.deck
        off.append(len(arena))
        arena += self.discard
        off.append(len(arena))
        if len(arena) > 128:
            raise OverflowError("arena")
        arena += [0] * (128 - len(arena))
        b = bytearray(256)
        b[0:128] = bytes(arena)
        b[128:155] = bytes(off)
        s8 = lambda v: v - 256 if v >= 128 else v
        for p in range(6):
            b[156 + p] = self.gold[p] & 0xFF
            # purses beyond a signed byte: 2-bit signed page next to the low byte (include/citadels_b200.h gold_hi03/45)
            page = (self.gold[p] - s8(self.gold[p] & 0xFF)) >> 8
            if p < 4:
                b[231] |= (page & 3) << (2 * p)
            else:
                b[247] |= (page & 3) << (2 * (p - 4))
            b[162 + p] = self.role[p]
            b[168 + p] = int(self.replicas[p]) & 0xFF
            b[174 + p] = (1 if self.lighthouse[p] else 0) | (2 if self.first7[p] else 0) | (4 if self.witch[p] else 0)
        for r in range(8):
            b[180 + r] = ((1 if self.dead[r] else 0) | (self.warrant[r] << 1) | ((1 if self.possessed[r] else 0) << 3)
                          | ((1 if self.robbed[r] else 0) << 4) | (self.blackmail[r] << 5))
            b[188 + r] = self.variant[r]
        for i in range(6):
            b[196 + i] = self.order[i]
        for i in range(6):
            b[202 + i] = (self.used_roles[i] + 1) if i < len(self.used_roles) else 0
        b[208] = len(self.used_roles)
        b[209] = sum(1 << r for r in self.rtc)
        b[210] = self.state
        b[211] = self.player
        b[212] = self.done
        b[213] = min(self.n_trade, 15) | (min(self.n_nontrade, 15) << 4)
        # next_gamestate is only meaningful (and only ever read) inside an interrupt state; a stale one
        # (e.g. dead graveyard owner finishing from state 6, game/agent.py:83) is canonicalised away.
        if self.state in (4, 6, 7, 8, 9, 10):
            b[214] = self.next_player
            b[215] = self.next_mode
        b[216] = self.crown
        b[217] = (1 if self.ending else 0) | (2 if self.terminal else 0)
        b[218] = self.winner & 0xFF
        b[219] = self.wiz_target
        for p in range(6):
            b[220 + p] = max(-128, min(127, self.points[p])) & 0xFF   # the record saturates (engine sets its overflow flag)
        b[226] = self.warrant_building
        b[227] = self.ruleset
        # tier C scratch state lives in the engine-private tail but is game state all the same
        b[229] = sum(1 << q for q in self.seer_from)        # seer_taken_card_from is always in seat order
        b[230] = len(self.seven)
        b[240:240 + len(self.seven)] = bytes(self.seven)
        return bytes(b)

    def encode_game(self, player_override=None):
        """Game.encode_game (game/game.py:34-128): the 418-feature value-model input, as a list of numbers."""
        f = [0.0] * 418
        for r in range(8):
            f[r * 3 + self.variant[r]] = 1.0
        cur = self.player if player_override is None else player_override
        for p in range(6):
            if self.role[p] < 8 and self.kr_conf[cur][p]:
                f[24 + p * 8 + self.role[p]] = 1.0
        for p in range(6):
            f[72 + p] = float(self.count_points(p))
            f[78 + p] = float(self.gold[p])
            f[84 + p] = float(len(self.hand[p]))
            for c in self.bld[p]:
                f[90 + p * 40 + ctype(c)] += 1.0
                f[330 + p * 5 + csuit(c)] += 1.0
        f[360 + cur] = 1.0
        f[366 + self.state] = 1.0
        f[377] = 1.0 if self.ending else 0.0
        for r in range(8):
            vals = (self.dead[r], self.warrant[r], self.possessed[r], self.robbed[r], self.blackmail[r])
            for i, v in enumerate(vals):
                if v:
                    f[378 + r * 5 + i] = 1.0
        return f

    def pack_know(self, viewer):
        """The engine's 592-byte knowledge block of one observer (csrc/ctd_engine.cuh `CtdKnow`)."""
        b = bytearray(592)   # header 16 | hk[64] x 4 | wiz_cards[48] @272 | pool[256] @320 | pool_used @576
        b[0] = viewer
        b[1] = sum(1 << q for q in range(6) if self.kr_conf[viewer][q])
        hks = self.kh[viewer]
        assert len(hks) <= 64
        b[2] = len(hks)
        wiz = self.wiz_cards if self.wiz_target != 0xFF else []
        b[3] = len(wiz)
        for q in range(6):
            struct.pack_into("<H", b, 4 + 2 * q, self.kr_mask[viewer][q])
        pos = 0
        for i, h in enumerate(hks):
            flags = (1 if h.wizard else 0) | (2 if h.used else 0)   # CtdHK: pid 4 bits | conf 3 | flags 2 | n 8 | off 9
            struct.pack_into("<I", b, 16 + 4 * i, (h.pid & 0xF) | (h.conf & 7) << 4 | flags << 7 | len(h.cards) << 9 | pos << 17)
            b[320 + pos:320 + pos + len(h.cards)] = bytes(h.cards)
            pos += len(h.cards)
        b[272:272 + len(wiz)] = bytes(wiz)
        struct.pack_into("<H", b, 576, pos)
        return bytes(b)

    def unpack_know(self, blob, used_cards):
        """Restore one observer's knowledge block (and Game.used_cards) into this game: enough to run CFR from it."""
        blob = bytes(blob)
        v = blob[0]
        n_hk, wiz_n = blob[2], blob[3]
        for q in range(6):
            self.kr_mask[v][q] = struct.unpack_from("<H", blob, 4 + 2 * q)[0]
            conf = bool(blob[1] >> q & 1)
            for o in range(6):
                self.kr_conf[o][q] = conf
        self.kh[v] = []
        for i in range(n_hk):
            word = struct.unpack_from("<I", blob, 16 + 4 * i)[0]
            pid, conf, flags, n, off = word & 0xF, (word >> 4) & 7, (word >> 7) & 3, (word >> 9) & 0xFF, (word >> 17) & 0x1FF
            pid = pid - 16 if pid >= 8 else pid
            h = HandKnowledge(pid, list(blob[320 + off:320 + off + n]), conf, bool(flags & 1))
            h.used = bool(flags & 2)
            self.kh[v].append(h)
        self.wiz_cards = list(blob[272:272 + wiz_n])
        self.used_cards = [int(c) for c in used_cards if int(c) != 0xFF]   # 0xFF pads the 66-card random deal to 76
        return v

    @classmethod
    def unpack(cls, rec, chance=None):
        g = cls(chance, ruleset=rec[227], deal=False)
        off = list(rec[128:155])
        arena = list(rec[0:128])
        seg = [arena[off[i]:off[i + 1]] for i in range(26)]
        for p in range(6):
            g.hand[p], g.bld[p], g.mus[p], g.jd[p] = seg[4 * p:4 * p + 4]
        g.deck, g.discard = seg[24], seg[25]
        s8 = lambda v: v - 256 if v >= 128 else v
        for p in range(6):
            page = (rec[231] >> (2 * p)) & 3 if p < 4 else (rec[247] >> (2 * (p - 4))) & 3
            g.gold[p] = s8(rec[156 + p]) + 256 * ((page ^ 2) - 2)
            g.role[p] = rec[162 + p]
            g.replicas[p] = s8(rec[168 + p])
            f = rec[174 + p]
            g.lighthouse[p], g.first7[p], g.witch[p] = bool(f & 1), bool(f & 2), bool(f & 4)
        for r in range(8):
            f = rec[180 + r]
            g.dead[r], g.warrant[r], g.possessed[r] = bool(f & 1), (f >> 1) & 3, bool(f & 8)
            g.robbed[r], g.blackmail[r] = bool(f & 16), (f >> 5) & 3
            g.variant[r] = rec[188 + r]
        g.order = list(rec[196:202])
        g.used_roles = [rec[202 + i] - 1 for i in range(rec[208])]
        g.rtc = [r for r in range(8) if rec[209] >> r & 1]
        g.state, g.player, g.done = rec[210], rec[211], rec[212]
        g.n_trade, g.n_nontrade = rec[213] & 15, rec[213] >> 4
        g.next_player, g.next_mode, g.crown = rec[214], rec[215], rec[216]
        g.ending, g.terminal = bool(rec[217] & 1), bool(rec[217] & 2)
        g.winner = s8(rec[218])
        g.wiz_target = rec[219]
        if g.wiz_target != 0xFF:
            g.wiz_cards = list(g.hand[g.wiz_target])
        g.points = [s8(rec[220 + p]) for p in range(6)]
        g.warrant_building = rec[226]
        g.seer_from = [q for q in range(6) if rec[229] >> q & 1]
        g.seven = list(rec[240:240 + rec[230]])
        return g


Game._APPLY = {
    K["role_pick"]: Game._a_role_pick,
    K["gold_or_card"]: Game._a_gold_or_card,
    K["which_card_to_keep"]: Game._a_keep,
    K["build"]: Game._a_build,
    K["empty_option"]: Game._a_empty,
    K["finish_round"]: Game._a_finish,
    K["smithy_choice"]: Game._a_smithy,
    K["laboratory_choice"]: Game._a_lab,
    K["magic_school_choice"]: Game._a_magic_school,
    K["weapon_storage_choice"]: Game._a_weapon_storage,
    K["lighthouse_choice"]: Game._a_lighthouse,
    K["museum_choice"]: Game._a_museum,
    K["graveyard"]: Game._a_graveyard,
    K["take_gold_for_war"]: Game._a_take_gold_for_war,
    K["assassination"]: Game._a_assassination,
    K["bewitching"]: Game._a_bewitching,
    K["steal"]: Game._a_steal,
    K["spy"]: Game._a_spy,
    K["magic_hand_change"]: Game._a_magic_hand_change,
    K["discard_and_draw"]: Game._a_discard_and_draw,
    K["look_at_hand"]: Game._a_look,
    K["take_from_hand"]: Game._a_take_from_hand,
    K["take_crown_king"]: Game._a_take_crown_king,
    K["bishop"]: Game._a_bishop,
    K["abbot_gold_or_card"]: Game._a_abbot,
    K["abbot_beg"]: Game._a_abbot_beg,
    K["merchant"]: Game._a_merchant,
    K["architect"]: Game._a_architect,
    K["navigator_gold_card"]: Game._a_navigator,
    K["warlord_desctruction"]: Game._a_warlord,
    K["magistrate_warrant"]: Game._a_warranting,
    K["reveal_warrant_as_magistrate"]: Game._a_magistrate_reveal,
    K["blackmail"]: Game._a_blackmail,
    K["blackmail_response"]: Game._a_blackmail_response,
    K["reveal_blackmail_as_blackmailer"]: Game._a_blackmail_reveal,
    K["seer"]: Game._a_seer,
    K["give_back_card"]: Game._a_seer_give_back,
    K["give_crown"]: Game._a_emperor,
    K["take_crown_pat"]: Game._a_patrician,
    K["cardinal_exchange"]: Game._a_cardinal,
    K["trader"]: Game._a_trader,
    K["scholar"]: Game._a_scholar,
    K["scholar_card_pick"]: Game._a_scholar_pick,
    K["marshal_steal"]: Game._a_marshal,
    K["diplomat_exchange"]: Game._a_diplomat,
}
for _k in range(len(KIND_NAMES)):
    Game._APPLY.setdefault(_k, Game._a_unimplemented)


def new_game(chance, ruleset=RULESET_PRESET):
    """run_utils.create_game, run_utils.py:20-27: Game(preset=True); setup_round()."""
    g = Game(chance, ruleset)
    g.setup_round()
    return g


def playout(seed, gid, ruleset=RULESET_PRESET, max_steps=4096, trace=None):
    """The reference's random-playout loop (run_utils.py:37-41) under the Philox chance stream.

    Returns (winner, points[6], steps).  `trace`, if a list, receives (packed state, options, chosen index)
    per step.
    """
    from .philox import PhiloxChance
    ch = PhiloxChance(seed, gid)
    g = new_game(ch, ruleset)
    steps = 0
    while steps < max_steps:
        opts = g.options()
        i = ch.randbelow(len(opts))
        if trace is not None:
            trace.append((g.pack(), opts, i))
        steps += 1
        if g.apply(opts[i]):
            break
    return g.winner, list(g.points), steps, g

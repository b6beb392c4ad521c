"""Harness around the REAL reference (/root/reference), used only in the build container.

It (1) routes every `random.shuffle` the reference performs (game/deck.py:73 via
`from random import shuffle`; game/game.py:152, :260) through a pluggable chance source, so
the reference can be driven by the same Philox stream as the oracle and the CUDA engine,
(2) dumps a reference `Game` into the engine's 256-byte record, and (3) canonicalises
reference `option` objects into 64-bit descriptors.  Nothing here is shipped or runs on the
GPU box: /root/reference does not exist there.  The committed fixtures under tests/golden/
are its output (see gen_golden.py).
"""
import os
import sys
import types
import random as _random

REFERENCE = "/root/reference"
_state = {"chance": None}


def available():
    return os.path.isdir(os.path.join(REFERENCE, "game"))


def _driven_shuffle(x):
    ch = _state["chance"]
    n = len(x)
    if n <= 1:
        return
    perm = ch.perm(n)
    x[:] = [x[i] for i in perm]


def _game_draw(n):
    ch = _state["chance"]
    return ch.randbelow_game(n) if hasattr(ch, "randbelow_game") else ch.randbelow(n)


def _raise_index():
    raise IndexError("Cannot choose from an empty sequence")


def load_reference():
    """Import the reference with plotting imports stubbed and shuffles routed through `_state['chance']`."""
    if "game.game" in sys.modules and getattr(sys.modules["game.game"], "_ctd_patched", False):
        return sys.modules["game.game"]
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    for m in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.setrecursionlimit(5000)
    _random.shuffle = _driven_shuffle          # before `from random import shuffle` in game/deck.py
    # Game.set_random_game (game/game.py:491-520): sample / choice / randint through the same chance source
    _random.sample = lambda pop, k: [pop[i] for i in _state["chance"].perm(len(pop))[:k]]
    _random.randint = lambda a, b: a + _game_draw(b - a + 1)
    _random.choice = lambda seq: _raise_index() if not len(seq) else seq[_game_draw(len(seq))]
    import game.deck
    import game.game
    game.deck.shuffle = _driven_shuffle
    game.game._ctd_patched = True
    return game.game


def set_chance(ch):
    _state["chance"] = ch


# --------------------------------------------------------------------------- canonicalisation
from oracle import citadels_oracle as O   # noqa: E402  (tests may import the oracle)

_CLASSIC = {0: "Assassin", 1: "Thief", 2: "Magician", 3: "King", 4: "Bishop", 5: "Merchant", 6: "Architect",
            7: "Warlord"}
_SUITS = {"trade": 0, "war": 1, "religion": 2, "lord": 3, "unique": 4}


def card_code(card):
    if card.type_ID == 25 and card.suit != "unique":
        return 40 + _SUITS[card.suit]
    return card.type_ID


def new_ref_game(chance, ruleset=O.RULESET_PRESET):
    gg = load_reference()
    set_chance(chance)
    g = gg.Game(preset=(ruleset != O.RULESET_RANDOM))
    if ruleset == O.RULESET_CLASSIC:
        g.roles = dict(_CLASSIC)
    g.setup_round()
    return g


def ref_pack(g, ruleset=O.RULESET_PRESET):
    from game.config import roles as cfg_roles, role_to_role_id
    arena, off = [], []
    for pl in g.players:
        for dk in (pl.hand, pl.buildings, pl.museum_cards, pl.just_drawn_cards):
            off.append(len(arena))
            arena += [card_code(c) for c in dk.cards]
    off.append(len(arena))
    arena += [card_code(c) for c in g.deck.cards]
    off.append(len(arena))
    arena += [card_code(c) for c in g.discard_deck.cards]
    off.append(len(arena))
    assert len(arena) <= 128
    arena += [0] * (128 - len(arena))
    b = bytearray(256)
    b[0:128] = bytes(arena)
    b[128:155] = bytes(off)
    crowns = [pl.id for pl in g.players if pl.crown]
    assert len(crowns) == 1
    wiz = [(pl.id, hk.player_id) for pl in g.players for hk in pl.known_hands if hk.wizard]
    assert len(wiz) <= 1
    for p, pl in enumerate(g.players):
        b[156 + p] = pl.gold & 0xFF
        b[162 + p] = 8 if pl.role is None else (9 if pl.role == "Bewitched" else role_to_role_id[pl.role])
        b[168 + p] = int(pl.replicas) & 0xFF
        b[174 + p] = (1 if pl.can_use_lighthouse else 0) | (2 if pl.first_to_7 else 0) | (4 if pl.witch else 0)
    enc = {None: 0, "Real": 1, "Fake": 2}
    for r in range(8):
        rp = g.role_properties[r]
        b[180 + r] = ((1 if rp.dead else 0) | (enc[rp.warrant] << 1) | ((1 if rp.possessed else 0) << 3)
                      | ((1 if rp.robbed else 0) << 4) | (enc[rp.blackmail] << 5))
        b[188 + r] = cfg_roles[r].index(g.roles[r])
    for i in range(6):
        b[196 + i] = g.turn_orders_for_roles[i]
    used = getattr(g, "used_roles", [])
    for i in range(6):
        b[202 + i] = (used[i] + 1) if i < len(used) else 0
    b[208] = len(used)
    b[209] = sum(1 << r for r in getattr(g, "roles_to_choose_from", {}).keys())
    gs = g.gamestate
    b[210] = gs.state
    b[211] = 0xFF if gs.player_id is None else gs.player_id
    done = gs.already_done_moves
    bits = {"smithy": 1, "lab": 2, "magic_school": 4, "museum": 8, "character_ability": 16, "begged": 32,
            "take_gold": 64}
    f = 0
    for m in done:
        f |= bits.get(m, 0)
    b[212] = f
    b[213] = min(done.count("trade_building"), 15) | (min(done.count("non_trade_building"), 15) << 4)
    if gs.state in (4, 6, 7, 8, 9, 10):
        ng = gs.next_gamestate
        assert ng.state == 5
        b[214] = ng.player_id
        if ng.already_done_moves is done:
            b[215] = O.NEXT_ALIAS
        elif ng.already_done_moves == ["character_ability"]:
            b[215] = O.NEXT_RESET_CA
        else:
            assert ng.already_done_moves == []
            b[215] = O.NEXT_EMPTY
    b[216] = crowns[0]
    b[217] = (1 if g.ending else 0) | (2 if g.terminal else 0)
    if g.terminal:
        b[218] = int(g.rewards.argmax())
        for p in range(6):
            b[220 + p] = g.points[p] & 0xFF
    else:
        b[218] = 0xFF
    b[219] = wiz[0][1] if wiz else 0xFF
    wb = getattr(g, "warrant_building", None)
    b[226] = 0xFF if wb is None else wb.type_ID
    b[227] = ruleset
    b[229] = sum(1 << q for q in getattr(g, "seer_taken_card_from", []))
    seven = getattr(g, "seven_drawn_cards", [])
    seven = [card_code(c) for c in (seven.cards if hasattr(seven, "cards") else seven)]
    b[230] = len(seven)
    b[240:240 + len(seven)] = bytes(seven)
    return bytes(b)


def ref_descriptors(options):
    """Reference `option` list -> 64-bit descriptors (order preserved)."""
    from game.config import role_to_role_id
    out = []
    r_counts = {}
    card_key, card_j = None, 0
    for o in options:
        a = o.attributes
        k = O.K[o.name]
        p = a["perpetrator"]
        n = o.name
        if n == "role_pick":
            d = O.D(k, p, rank=role_to_role_id[a["choice"]])
        elif n in ("gold_or_card", "navigator_gold_card", "magic_school_choice", "blackmail_response"):
            d = O.D(k, p, named=O.NAMED[a["choice"]])
        elif n == "which_card_to_keep":
            ch = list(a["choice"])
            d = O.D(k, p, a=ch[0].type_ID, b=ch[1].type_ID if len(ch) > 1 else None)
        elif n == "finish_round":
            d = O.D(k, p, next_witch=a["next_witch"], crown=a["crown"])
        elif n == "build":
            d = O.D(k, p, a=a["built_card"].type_ID, replica=a["replica"])
        elif n in ("laboratory_choice", "lighthouse_choice", "museum_choice"):
            d = O.D(k, p, a=a["choice"].type_ID)
        elif n in ("weapon_storage_choice", "warlord_desctruction"):
            d = O.D(k, p, target=a["target"], a=a["choice"].type_ID)
        elif n in ("assassination", "bewitching", "steal"):
            d = O.D(k, p, rank=a["choice"])
        elif n == "spy":
            d = O.D(k, p, target=a["target"], named=O.NAMED[a["suit"]])
        elif n in ("look_at_hand", "magic_hand_change"):
            d = O.D(k, p, target=a["target"])
        elif n == "take_from_hand":
            if a["build"]:
                d = O.D(k, p, target=a["target"], a=a["built_card"].type_ID, build=1, replica=a["replica"])
            else:
                d = O.D(k, p, target=a["target"], a=a["card"].type_ID, build=0)
        elif n == "abbot_gold_or_card":
            d = O.D(k, p, count=a["gold_or_card_combination"].count("card"), r=len(a["gold_or_card_combination"]))
        elif n == "discard_and_draw":
            r = len(a["cards"])
            j = r_counts.get(r, 0)
            r_counts[r] = j + 1
            d = O.D(k, p, r=r, j=j)
        elif n == "magistrate_warrant":
            d = O.D(k, p, rank=a["real_target"], named=a["fake_targets"][0], count=a["fake_targets"][1])
        elif n == "blackmail":
            d = O.D(k, p, rank=a["real_target"], named=a["fake_target"])
        elif n in ("reveal_blackmail_as_blackmailer", "reveal_warrant_as_magistrate"):
            d = O.D(k, p, target=a["target"], named=O.NAMED[a["choice"]])
        elif n == "give_back_card":
            d = O.D_handout(p, [c.type_ID for c in a["card_handouts"].values()])
        elif n == "give_crown":
            d = O.D(k, p, target=a["target"], named={"gold": 0, "card": 1, "nothing": O.NAMED_NOTHING}[a["gold_or_card"]])
        elif n == "scholar_card_pick":
            d = O.D(k, p, a=a["choice"].type_ID)
        elif n == "cardinal_exchange":
            key = (a["target"], id(a["built_card"]))
            card_j = card_j + 1 if key == card_key else 0
            card_key = key
            d = O.D(k, p, target=a["target"], a=a["built_card"].type_ID, replica=a["replica"], build=a["factory"],
                    count=len(a["cards_to_give"]), j=card_j)
        elif n == "marshal_steal":
            d = O.D(k, p, target=a["target"], a=a["choice"].type_ID)
        elif n == "diplomat_exchange":
            d = O.D(k, p, target=a["target"], a=a["choice"].type_ID, b=a["give"].type_ID)
        elif n in ("empty_option", "smithy_choice", "graveyard", "take_gold_for_war", "take_crown_king",
                   "abbot_beg", "bishop", "merchant", "architect", "seer", "take_crown_pat", "trader", "scholar"):
            d = O.D(k, p)
        else:
            raise NotImplementedError(n)
        out.append(d)
    return out


# --------------------------------------------------------------------------- CFR chance + knowledge dumps
def patch_cfr_chance():
    """Route the reference's CFR-path randomness through `_state['chance']` (mapping: oracle/mccfr_oracle.py)."""
    import numpy as np
    load_reference()

    def np_choice(a, size=None, replace=True, p=None):
        ch = _state["chance"]
        n = len(a)
        caller = sys._getframe(1).f_code.co_name
        if caller in ("action_choice", "get_option_from_role_preference"):
            cdf = np.cumsum(p)
            cdf = cdf / cdf[-1]
            i = min(int(np.searchsorted(cdf, ch.uniform(), side="right")), n - 1)
        else:
            assert caller in ("expand_role_pick", "expand_for_opponents"), caller
            i = ch.randbelow(n)
        return a[i]

    def rnd_choice(seq):
        if not len(seq):
            raise IndexError("Cannot choose from an empty sequence")
        return seq[_state["chance"].randbelow(len(seq))]

    np.random.choice = np_choice
    _random.random = lambda: _state["chance"].uniform()
    import algorithms.deep_mccfr as dm
    dm.randint = lambda a, b: a + _state["chance"].randbelow(b - a + 1)   # build_train_targets' viewpoint seat


def ref_knowledge(g):
    out = []
    for pl in g.players:
        masks, confs = [], []
        for rk in pl.known_roles:
            m = 0
            for rid in rk.possible_roles.keys():
                m |= 1 << (8 if rid == -1 else rid)
            masks.append(m)
            confs.append(bool(rk.confirmed))
        hks = tuple((hk.player_id, hk.confidence, bool(hk.wizard), tuple(card_code(c) for c in hk.hand.cards))
                    for hk in pl.known_hands)
        out.append((tuple(masks), tuple(confs), hks))
    return out


def ref_pack_know(g, viewer):
    """Reference Agent.known_roles / known_hands of `viewer` -> a 592-byte knowledge block in the ROUND-1 entry format (32 entries of
    8 bytes), the one the committed fixtures and their know_crc checksums were made with; tests convert with
    citadels_self_play_b200.layout.know_from_v1 / know_to_v1 (the engine's own format holds 64 entries of 4 bytes)."""
    import struct
    pl = g.players[viewer]
    b = bytearray(592)   # header 16 | hk[32] x 8 | wiz_cards[48] @272 | pool[256] @320 | pool_used @576
    b[0] = viewer
    b[1] = sum(1 << q for q in range(6) if pl.known_roles[q].confirmed)
    assert len(pl.known_hands) <= 32
    b[2] = len(pl.known_hands)
    wiz = [hk for p in g.players for hk in p.known_hands if hk.wizard]
    wcards = [card_code(c) for c in wiz[0].hand.cards] if wiz else []
    b[3] = len(wcards)
    for q in range(6):
        m = 0
        for rid in pl.known_roles[q].possible_roles.keys():
            m |= 1 << (8 if rid == -1 else rid)
        struct.pack_into("<H", b, 4 + 2 * q, m)
    pos = 0
    for i, hk in enumerate(pl.known_hands):
        cards = [card_code(c) for c in hk.hand.cards]
        struct.pack_into("<bBBBHH", b, 16 + 8 * i, hk.player_id, hk.confidence,
                         (1 if hk.wizard else 0) | (2 if hk.used else 0), len(cards), pos, 0)
        b[320 + pos:320 + pos + len(cards)] = bytes(cards)
        pos += len(cards)
    b[272:272 + len(wcards)] = bytes(wcards)
    struct.pack_into("<H", b, 576, pos)
    return bytes(b)

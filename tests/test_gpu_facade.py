"""The reference-facing facade (citadels_self_play_b200.facade) driven the way the reference's own loops drive
game/ (run_utils.py:37-41), compared step by step with the oracle.  The caller's choice is external randomness
(`random.choice`), so only shuffles consume the games' Philox stream on both sides."""
import copy
import random
import numpy as np
import pytest

from tests.golden_util import visible

pytestmark = pytest.mark.gpu


def _pair(seed, gid, ruleset):
    from citadels_self_play_b200 import facade as F
    from oracle import citadels_oracle as O
    from oracle.philox import PhiloxChance
    return F.create_game(seed=seed, gid=gid, ruleset=ruleset), O.new_game(PhiloxChance(seed, gid), ruleset)


@pytest.mark.parametrize("ruleset", [0, 1, 2])
def test_reference_loop_through_facade(ruleset):
    from oracle import citadels_oracle as O
    rng = random.Random(7 + ruleset)
    for gid in (11, 12):
        fg, og = _pair(99, gid, ruleset)
        steps = 0
        while True:
            options = fg.get_options_from_state()
            assert [o.desc for o in options] == og.options()
            assert visible(fg.record()) == visible(og.pack())
            for o in range(6):
                assert fg._know[o * 592:(o + 1) * 592].tobytes() == og.pack_know(o), (gid, steps, o)
            assert fg.gamestate.player_id == og.player and fg.gamestate.state == og.state
            assert all(o.name == O.KIND_NAMES[O.d_kind(o.desc)] and o.attributes["perpetrator"] == og.player for o in options)
            chosen = rng.choice(options)
            winner = chosen.carry_out(fg)
            won = og.apply(chosen.desc)
            steps += 1
            assert bool(winner) == won
            if winner:
                assert winner.id == og.winner and fg.terminal and fg.points == og.points
                assert list(fg.rewards) == [1.0 if p == og.winner else 0.0 for p in range(6)]
                break
        assert 200 < steps < 1000


def test_agent_views_and_deepcopy():
    fg, og = _pair(5, 77, 0)
    rng = random.Random(1)
    for _ in range(120):
        rng.choice(fg.get_options_from_state()).carry_out(fg)
    snap = copy.deepcopy(fg)
    before = snap.record()
    for _ in range(10):
        rng.choice(fg.get_options_from_state()).carry_out(fg)
    assert snap.record() == before and fg.record() != before      # the copy is independent
    assert snap == copy.deepcopy(snap)
    p = snap.players[snap.gamestate.player_id]
    assert p.get_options(snap) == snap.get_options_from_state()
    assert sum(len(a.hand.cards) + len(a.buildings.cards) + len(a.museum_cards.cards) + len(a.just_drawn_cards.cards)
               for a in snap.players) + len(snap.deck.cards) + len(snap.discard_deck.cards) == 76
    assert sum(a.crown for a in snap.players) == 1
    assert all(c.cost >= 1 for a in snap.players for c in a.hand.cards)


def test_cfrnode_through_facade_matches_oracle():
    from citadels_self_play_b200 import facade as F
    from oracle import mccfr_oracle as M
    from oracle.philox import PhiloxChance
    from tests.mccfr_util import oracle_preorder, tree_preorder, assert_same_tree
    fg, og = _pair(31337, 5, 0)
    rng = random.Random(3)
    while True:   # play to the end keeping copies, then step back like create_a_close_to_finished_game
        snaps_f, snaps_o = [copy.deepcopy(fg)], [og.clone()]
        done = False
        while not done:
            ch = rng.choice(fg.get_options_from_state())
            done = bool(ch.carry_out(fg))
            og.apply(ch.desc)
            snaps_f.append(copy.deepcopy(fg))
            snaps_o.append(og.clone())
        break
    k = len(snaps_f) - 12
    while len(snaps_f[k].get_options_from_state()) < 2:
        k += 1
    rf, ro = snaps_f[k], snaps_o[k]
    viewer = rf.gamestate.player_id
    node = F.CFRNode(rf, original_player_id=viewer)
    node.cfr_train(max_iterations=200)
    ro.chance = PhiloxChance(31337, 5, stream=1)
    on = M.Node(ro, viewer)
    on.cfr_train(200)
    assert_same_tree(oracle_preorder(on), tree_preorder(node._tree), "facade")
    assert len(node.children) == len(on.children)
    assert [o.desc for o, _ in node.children] == [d for d, _ in on.children]
    assert np.allclose(node.cumulative_regrets, on.R) and np.allclose(node.node_value, on.V)
    assert rf.record()[:228] == on.game.pack()[:228]          # skip_false_choice advanced the caller's game
    child = node.children[0][1]
    assert child.parent is node and child.depth == 1
    _, chosen = node.action_choice(live=True)
    assert chosen in [o for o, _ in node.children] or node.role_pick_node
    chosen2, root2 = F.run_mccfr(copy.deepcopy(snaps_f[k]), max_iterations=50)
    assert chosen2.name in [o.name for o in root2.game.get_options_from_state()]


def test_sample_private_information_and_option_encoding_match_oracle():
    """The two facade methods that only CFRNode calls in the reference: Game.sample_private_information (game/game.py:215-242,
    device: ctd_game_sample) against the oracle's restatement on the same chance stream, and option.encode_option
    (game/option.py:52-115) against the oracle's encoder."""
    from oracle import mccfr_oracle as M
    rng = random.Random(11)
    for ruleset, gid in ((0, 301), (0, 302), (2, 303)):
        fg, og = _pair(4321, gid, ruleset)
        for _ in range(150):
            ch = rng.choice(fg.get_options_from_state())
            if ch.carry_out(fg):
                break
            og.apply(ch.desc)
        else:
            for o in fg.get_options_from_state()[:12]:
                assert np.array_equal(o.encode_option().numpy()[0], np.asarray(M.encode_option(o.desc), dtype=np.float32))
                assert o.encode_option().shape == (1, 131)
            viewer = (fg.gamestate.player_id + 2) % 6
            fg.sample_private_information(fg.players[viewer], role_sample=True)
            og.sample_private_information(viewer, role_sample=True)
            assert visible(fg.record()) == visible(og.pack()), (ruleset, gid)
            assert fg._know[viewer * 592:(viewer + 1) * 592].tobytes() == og.pack_know(viewer)
            assert len(fg.get_options_from_state()) >= 1

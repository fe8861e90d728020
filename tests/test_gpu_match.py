"""`Match` / `Game` on the device: all games of a match advance in lock step, one batched search per
player per round.  Golden `match.npz`: the unmodified reference's Match (oinkoink/match.py:14-76) between two different
deterministic MCTS players over every 1-ply (14 games) and 2-ply (98 games) opening with sides switched -- every game's
moves and result, and the W/D/L summary."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _players(cfg):
    from connect4_b200 import evaluators as evl
    from connect4_b200.mcts import MCTS, MCTSConfig
    mk = lambda name, c: MCTS(name, MCTSConfig(int(c[0]), int(c[1]), float(c[2])),
                              evl.Evaluator(evl.evaluate_centre_with_prior))
    return mk("one", cfg[0]), mk("two", cfg[1])


@pytest.mark.parametrize("mi", [0, 1])
def test_batched_match_equals_the_reference_game_for_game(mi, capsys):
    from connect4_b200.match import Match
    z = np.load(os.path.join(GOLD, "match.npz"))
    p1, p2 = _players(z["m%d_cfg" % mi])
    match = Match(False, p1, p2, plies=int(z["m%d_plies" % mi]), switch=True)
    n = int(z["m%d_n" % mi])
    assert match.n == n and len(match.games) == 2 * n
    starts = [(int(g._board.color[0]), int(g._board.color[1])) for g in match.games]
    res = match.play()
    want = z["m%d_summary" % mi]
    assert [res["wins"], res["draws"], res["losses"]] == want[:3].tolist()
    assert res["return"] == pytest.approx(want[3], abs=0)
    assert "The results for one vs two are:" in capsys.readouterr().out
    # per game: key = (opening, switched)
    gold = {}
    for i in range(2 * n):
        mv = z["m%d_moves" % mi][i]
        gold[(int(z["m%d_c0" % mi][i]), int(z["m%d_c1" % mi][i]), i >= n)] = (mv[mv >= 0].tolist(), float(z["m%d_result" % mi][i]))
    assert len(gold) == 2 * n
    for i, g in enumerate(match.games):
        moves, result = gold[(starts[i][0], starts[i][1], i >= n)]
        assert g.move_history.tolist() == moves, (i, g.move_history.tolist(), moves)
        assert g._board.result.value == result


def test_sequential_path_plays_the_same_games():
    """Game.play() (one device search per move) and the lock-step match agree"""
    from copy import copy
    from connect4_b200.board import make_random_ips
    from connect4_b200.game import Game
    from connect4_b200.match import Match
    z = np.load(os.path.join(GOLD, "match.npz"))
    p1, p2 = _players(z["m0_cfg"])
    match = Match(False, p1, p2, plies=1, switch=False)
    boards = [copy(g._board) for g in match.games]
    match.play()
    for g, b in list(zip(match.games, boards))[:3]:
        solo = Game(False, copy(p1), copy(p2), b)
        assert solo.play() == g._board.result
        assert solo.move_history.tolist() == g.move_history.tolist()


def test_network_player_against_centre_player():
    """the `_match` of the reference's training loop (neural/training.py:176-207): network MCTS vs centre MCTS"""
    from functools import partial
    from connect4_b200 import evaluators as evl
    from connect4_b200.match import Match
    from connect4_b200.mcts import MCTS, MCTSConfig
    from connect4_b200.neural.model import ModelWrapper
    from oracle import net_ref as nr
    model = ModelWrapper(state_dict=nr.load_golden_state(os.path.join(GOLD, "example_net_state.npz")))
    az = MCTS("AlphaZero", MCTSConfig(100, 19652, 1.25, 0.0, 0.0, 0), evl.Evaluator(partial(evl.evaluate_nn, model=model)))
    opp = MCTS("Evaluate_centre_with_prior", MCTSConfig(simulations=100), evl.Evaluator(evl.evaluate_centre_with_prior))
    res = Match(False, az, opp, plies=1, switch=True).play(agents=4)
    assert res["wins"] + res["draws"] + res["losses"] == 14
    assert 0.0 <= res["return"] <= 1.0
    # a bare ModelWrapper is a valid evaluator too ("Older net", training.py:189-195) and must give the same match
    res2 = Match(False, az, MCTS("Older net", MCTSConfig(simulations=100), model), plies=1, switch=True).play()
    res3 = Match(False, az, MCTS("same", MCTSConfig(simulations=100), evl.Evaluator(partial(evl.evaluate_nn, model=model))),
                 plies=1, switch=True).play()
    assert res2 == res3

"""Self-play generation on the device game pool and the reference-facing Python API, vs the reference's recorded games
(tests/golden/games.npz) and the oracle."""
import os
from functools import partial

import numpy as np
import pytest

from conftest import GOLDEN, bits, golden, random_positions

pytestmark = pytest.mark.gpu


def _cfg(sims, alpha=0.0, frac=0.0, sampling=0, init=1.25):
    from connect4_b200.mcts import MCTSConfig
    return MCTSConfig(sims, 19652, init, alpha, frac, sampling)


def _by_game(rec):
    out = {}
    for r in rec[np.lexsort((rec["ply"], rec["game_id"]))]:
        out.setdefault(int(r["game_id"]), []).append(r)
    return out


def _check_game(recs, moves, values, priors, result_value, c0=None, c1=None):
    assert [int(r["move"]) for r in recs] == [int(m) for m in moves]
    assert np.array_equal(np.array([r["search_value"] for r in recs], np.float32), np.asarray(values, np.float64).astype(np.float32))
    assert np.array_equal(np.stack([r["policy"] for r in recs]), np.asarray(priors, np.float64).astype(np.float32))
    assert all(float(r["result_value"]) == result_value and int(r["n_moves"]) == len(moves) for r in recs)
    assert [int(r["ply"]) for r in recs] == list(range(len(moves)))
    if c0 is not None:
        assert [int(r["c0"]) for r in recs] == [int(x) for x in c0] and [int(r["c1"]) for r in recs] == [int(x) for x in c1]


def test_deterministic_selfplay_equals_reference_games():
    """training_game with the deterministic evaluator (no noise, no sampling): whole games reproduced exactly"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    g = golden("games.npz")
    for i in range(3):
        pool = SelfPlayPool("centre", _cfg(int(g["det%d_sims" % i])), concurrent_games=4)
        rec = pool.generate_records(1)
        _check_game(list(rec), g["det%d_moves" % i], g["det%d_values" % i], g["det%d_priors" % i],
                    float(g["det%d_result" % i]), g["det%d_c0" % i], g["det%d_c1" % i])
        pool.engine.close()


def test_alphazero_selfplay_with_reference_randomness():
    """AlphaZero settings (alpha 0.3, frac 0.25, 6 sampled moves): the reference's recorded gamma draws and
    np.random.choice uniforms injected -> identical games"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    g = golden("games.npz")
    for i in range(4):
        pool = SelfPlayPool("centre", _cfg(int(g["az%d_sims" % i]), 0.3, 0.25, 6), concurrent_games=2)
        nz, un = g["az%d_noise" % i], g["az%d_uniform" % i]
        pool.engine.set_rng("injected", noise=nz[None], uniform=un[None])
        rec = pool.generate_records(1)
        _check_game(list(rec), g["az%d_moves" % i], g["az%d_values" % i], g["az%d_priors" % i], float(g["az%d_result" % i]))
        pool.engine.close()


def test_philox_selfplay_replayed_by_oracle(oracle):
    """device-generated noise (Philox) is recorded and the oracle replays every game with it: same moves / policies"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    sims, G = 60, 16
    pool = SelfPlayPool("centre", _cfg(sims, 0.3, 0.25, 6), concurrent_games=G, seed=1234)
    pool.engine.set_rng("philox", seed=1234, record=True)
    rec = pool.generate_records(G)
    nz, un = pool.engine.recorded_rng()
    games = _by_game(rec)
    assert sorted(games) == list(range(G))
    draws = []
    for gid, recs in games.items():
        out = oracle.selfplay_centre(oracle.make_config(sims, 19652, 1.25, 0.3, 0.25, 6), noise=nz[gid], uniform=un[gid])
        _check_game(recs, out["moves"], out["values"], out["priors"], out["result"] * 0.5, out["c0"], out["c1"])
        draws.append(nz[gid][:len(recs)].reshape(-1))
    draws = np.concatenate(draws)
    assert (draws > 0).all() and abs(draws.mean() - 0.3) < 0.06           # gamma(0.3, 1): mean 0.3
    assert len({tuple(int(r["move"]) for r in v) for v in games.values()}) > 1   # games differ
    pool.engine.close()


def test_reseeding_many_games_on_few_slots():
    """40 deterministic games on 8 slots: every game is the reference's game, every id appears once"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    g = golden("games.npz")
    pool = SelfPlayPool("centre", _cfg(int(g["det0_sims"])), concurrent_games=8)
    rec = pool.generate_records(40, game_id_base=3, game_id_stride=5)
    games = _by_game(rec)
    assert sorted(games) == [3 + 5 * i for i in range(40)]
    for recs in games.values():
        _check_game(recs, g["det0_moves"], g["det0_values"], g["det0_priors"], float(g["det0_result"]))
    gd = pool.generate(3)
    assert [list(x.moves) for x in gd] == [[int(m) for m in g["det0_moves"]]] * 3
    assert gd[0].result.value == float(g["det0_result"]) and gd[0].boards[1].age == 1
    pool.engine.close()


def test_selfplay_from_start_positions(oracle):
    from connect4_b200.neural.game_pool import SelfPlayPool
    c0, c1 = random_positions(77, 12, max_plies=20)
    pool = SelfPlayPool("centre", _cfg(40), concurrent_games=4)
    rec = pool.generate_records(12, start=(c0, c1))
    games = _by_game(rec)
    for gid, recs in games.items():
        out = oracle.selfplay_centre(oracle.make_config(40), start=(int(c0[gid]), int(c1[gid])))
        _check_game(recs, out["moves"], out["values"], out["priors"], out["result"] * 0.5, out["c0"], out["c1"])
    pool.engine.close()


def _golden_model():
    from oracle import net_ref as nr
    from connect4_b200.neural.model import ModelWrapper
    return ModelWrapper(state_dict=nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz")))


def test_net_guided_search_equals_oracle_with_same_network(oracle):
    """tree kernels + CUDA net (C4_EVAL_NET) vs the oracle search fed with the CUDA net's own outputs: the search
    logic (float32 prior normalisation, fp64 PUCT) must agree exactly"""
    from connect4_b200.engine import Engine
    model = _golden_model()
    c0, c1 = random_positions(31, 48)
    sims = 120
    eng = Engine(64, _cfg(sims))
    eng.set_net(model)
    eng.begin(c0, c1)
    eng.run("net")
    out = eng.readout()

    def ev(a, b):
        v, p = model.evaluate_bitboards(np.array([a], np.uint64), np.array([b], np.uint64))
        return float(v.cpu().numpy()[0]), p.cpu().numpy()[0]
    for i in range(16):
        t = oracle.Tree(oracle.make_config(sims), int(c0[i]), int(c1[i])).search(ev)
        v, s, r, a = t.root_children()
        assert (v == out["visits"][i]).all(), i
        assert (bits(s) == bits(out["vsum"][i])).all(), i
        assert t.best_move() == out["best"][i] and t.root_stats()[3] == out["nodes"][i]
    eng.close()


def test_net_selfplay_generates_legal_consistent_games(oracle):
    """NN-guided AlphaZero self-play with re-seeding: every record chain is a legal game with the right result"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _golden_model()
    pool = SelfPlayPool(model, _cfg(48, 0.3, 0.25, 6), concurrent_games=64, seed=5)
    rec = pool.generate_records(150)
    games = _by_game(rec)
    assert sorted(games) == list(range(150))
    lengths = []
    for recs in games.values():
        c0 = c1 = 0
        res = -1
        for k, r in enumerate(recs):
            assert (int(r["c0"]), int(r["c1"]), int(r["ply"])) == (c0, c1, k) and res == -1
            legal = oracle.legal_mask(c0, c1)
            assert legal >> int(r["move"]) & 1
            pol = r["policy"]
            assert abs(float(pol.sum()) - 1.0) < 1e-5 and all(pol[c] == 0 for c in range(7) if not legal >> c & 1)
            c0, c1, res = oracle.drop(c0, c1, int(r["move"]))
        assert res != -1 and res == int(recs[-1]["result"]) and float(recs[0]["result_value"]) == res * 0.5
        assert int(recs[0]["n_moves"]) == len(recs)
        lengths.append(len(recs))
    assert 7 <= min(lengths) and max(lengths) <= 42
    b, v, p = pool.last_dataset()
    assert b.shape == (2 * len(rec), 3, 6, 7) and v.shape == (2 * len(rec),) and p.shape == (2 * len(rec), 7)
    stats = pool.throughput(20)
    assert stats["evals"] > 0 and stats["device_ms"] > 0
    pool.engine.close()


# ---------------------------------------------------------------------------------- reference-facing Python API
def test_mcts_player_reference_kats():
    """reference tests/player_test.py:151-179 verbatim in structure: MCTS player, pb_c_init=9999, move in ans"""
    import json
    from copy import copy
    from connect4_b200.board import Board
    from connect4_b200.evaluators import Evaluator, evaluate_centre_with_prior
    from connect4_b200.mcts import MCTS, MCTSConfig
    kat = json.load(open(os.path.join(GOLDEN, "board_kat.json")))
    m = golden("mcts_kat.npz")
    for i, case in enumerate(kat["player_cases"]):
        plies = case["plies"]
        board = Board.from_pieces(o_pieces=np.array(case["o"], np.bool_), x_pieces=np.array(case["x"], np.bool_))
        computer = MCTS("mcts_test", MCTSConfig(simulations=7 ** plies + 1 if plies <= 6 else 2 ** plies, pb_c_init=9999),
                        Evaluator(evaluate_centre_with_prior))
        board_copy = copy(board)
        move, value, tree = computer.make_move(board_copy)
        for r_child in tree.root.children:
            for child in r_child.children:
                child.children = []
        assert move in case["ans"] and move == int(m["best"][i])
        assert board_copy.age == board.age + 1
        assert tree.count_nodes() <= int(m["nodes"][i])
        assert (bits(tree.get_values_policy()) == bits(m["vpolicy"][i])).all()
        assert np.float64(value).view(np.uint64) == m["best_value"][i].view(np.uint64)
        assert tree.root.data.search_value.visit_count == int(m["root_visits"][i])


def test_training_game_api_equals_reference():
    """training_game(player) through the Python surface: deterministic game, then the AlphaZero game replayed with
    the SAME numpy seed as the reference run (search() draws np.random.gamma / np.random.choice like the reference)"""
    from connect4_b200.evaluators import Evaluator, evaluate_centre_with_prior
    from connect4_b200.mcts import MCTS, MCTSConfig
    from connect4_b200.neural.training_game import training_game
    g = golden("games.npz")
    gd = training_game(MCTS("g", MCTSConfig(simulations=int(g["det0_sims"])), Evaluator(evaluate_centre_with_prior)))
    assert gd.moves == [int(x) for x in g["det0_moves"]]
    assert (bits(np.array(gd.values)) == bits(g["det0_values"])).all()
    assert (bits(np.array(gd.priors)) == bits(g["det0_priors"])).all()
    assert gd.result.value == float(g["det0_result"])
    np.random.seed(0)      # tests/golden/generate_goldens.py gen_games: (sims 60, seed 0)
    gd = training_game(MCTS("g", MCTSConfig(int(g["az0_sims"]), 19652, 1.25, 0.3, 0.25, 6),
                            Evaluator(evaluate_centre_with_prior)))
    assert gd.moves == [int(x) for x in g["az0_moves"]]
    assert (bits(np.array(gd.priors)) == bits(g["az0_priors"])).all()
    assert gd.result.value == float(g["az0_result"])


def test_host_callable_evaluator_through_mcts(oracle):
    """any `evaluator(board) -> (value, prior)` callable works (evaluators.py:18-25 protocol)"""
    from connect4_b200.board import Board
    from connect4_b200.evaluators import Evaluator
    from connect4_b200.mcts import MCTSConfig, search

    def fn(board):
        a, b = int(board.color[0]), int(board.color[1])
        v = ((a * 31 + b * 17) % 101) / 100.0
        p = np.array([1, 2, 3, 4, 3, 2, 1], np.float64) / 16.0
        return v, p
    board = Board()
    board.make_move(2)
    tree = search(MCTSConfig(90), board, Evaluator(fn))
    t = oracle.Tree(oracle.make_config(90), int(board.color[0]), int(board.color[1]))
    t.search(lambda a, b: fn(Board.from_bitboards(a, b)))
    assert (bits(tree.get_values_policy()) == bits(t.values_policy())).all()
    assert tree.best_move().name == t.best_move() and tree.count_nodes() == t.root_stats()[3]


def test_nn_evaluator_through_mcts_and_game():
    from connect4_b200.board import Board
    from connect4_b200.evaluators import Evaluator, evaluate_centre_with_prior, evaluate_nn
    from connect4_b200.game import Game
    from connect4_b200.match import Match
    from connect4_b200.mcts import MCTS, MCTSConfig
    model = _golden_model()
    nn_player = MCTS("nn", MCTSConfig(64), Evaluator(partial(evaluate_nn, model=model)))
    centre = MCTS("centre", MCTSConfig(64), Evaluator(evaluate_centre_with_prior))
    assert Evaluator(partial(evaluate_nn, model=model)).device_kind()[0] == "net"
    res = Game(False, nn_player, centre, Board()).play()
    assert res is not None
    out = Match(False, nn_player, centre, plies=1, switch=True).play()
    assert out["wins"] + out["draws"] + out["losses"] == 14 and 0.0 <= out["return"] <= 1.0


def test_generation_sink_matches_reference_tensors(tmp_path):
    """records -> data.pth tensors with flip augmentation == the reference's native_to_pytorch output (sink.npz)"""
    import torch
    from connect4_b200.board import Board
    from connect4_b200.neural.data import Connect4Dataset, TrainingDataStorage, native_to_pytorch
    from connect4_b200.neural.training_game import GameData
    from connect4_b200.utils import Result
    s = golden("sink.npz")
    boards = [Board.from_bitboards(a, b) for a, b in zip(s["c0"], s["c1"])]
    bt, vt, pt = native_to_pytorch(boards, list(s["values"]), list(s["priors"]), add_fliplr=True)
    assert np.array_equal(bt.numpy().astype(np.uint8), s["boards_t"])
    assert np.array_equal(vt.numpy(), s["values_t"]) and np.array_equal(pt.numpy(), s["priors_t"])
    gd = GameData()
    for b, p in zip(boards, s["priors"]):
        gd.add_move(b, 0, 0.5, p)
    gd.result = Result(float(s["values"][0]))
    TrainingDataStorage().save([gd], str(tmp_path))
    ds = Connect4Dataset.load(str(tmp_path / "data.pth"))
    assert np.array_equal(ds.boards.numpy().astype(np.uint8), s["boards_t"]) and len(ds) == 2 * len(boards)
    assert os.path.exists(tmp_path / "games.pkl")


def test_evaluation_memo_changes_work_not_results(monkeypatch):
    """the device-side evaluation memo (Evaluator.position_table of the reference, evaluators.py:18-25) is a pure cache:
    NN-guided self-play with it on and off produces identical records, and it does get hits"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _golden_model()
    cfg = _cfg(64, 0.3, 0.25, 6)
    recs = {}
    for log2 in ("0", "18"):
        monkeypatch.setenv("C4_MEMO_LOG2", log2)
        pool = SelfPlayPool(model, cfg, concurrent_games=32, seed=11)
        assert pool.engine.lib.c4_ctx_get(pool.engine.h, 4) == int(log2)
        rec = pool.generate_records(64)
        recs[log2] = rec[np.lexsort((rec["ply"], rec["game_id"]))]
        stats = pool.throughput(50)
        assert (stats["memo_hits"] > 0) == (log2 != "0")
        pool.engine.close()
    a, b = recs["0"], recs["18"]
    assert len(a) == len(b)
    for f in a.dtype.names:                       # field by field: bit-identical (NaN search values compare by bits)
        assert a[f].tobytes() == b[f].tobytes(), f

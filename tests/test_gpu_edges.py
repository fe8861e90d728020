"""Edge cases and "changes work, not results" guards of the engine: empty and ragged inputs, the IEEE-division fallback of
the select loop, the adaptive pass length, fewer games than slots."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, golden

pytestmark = pytest.mark.gpu


def _cfg(sims, alpha=0.0, frac=0.0, sampling=0):
    from connect4_b200.mcts import MCTSConfig
    return MCTSConfig(sims, 19652, 1.25, alpha, frac, sampling)


def _model():
    from connect4_b200.neural.model import ModelWrapper
    from oracle import net_ref as nr
    return ModelWrapper(state_dict=nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz")))


def _sorted(rec):
    return rec[np.lexsort((rec["ply"], rec["game_id"]))]


def _same_records(a, b):
    assert len(a) == len(b)
    for f in a.dtype.names:
        assert a[f].tobytes() == b[f].tobytes(), f


def test_division_fallback_is_bit_identical(monkeypatch):
    """the select loop's table-reciprocal division and its __ddiv_rn fallback give the same searches, both equal to
    the reference's (1,000 positions of the 800-simulation sweep)"""
    from connect4_b200.engine import Engine
    m = golden("mcts_sweep_800.npz")
    idx = np.arange(0, len(m["c0"]), 10)
    outs = []
    for no_fast in ("", "1"):
        if no_fast:
            monkeypatch.setenv("C4_NO_FASTDIV", "1")
        eng = Engine(len(idx), _cfg(800))
        eng.begin(m["c0"][idx], m["c1"][idx])
        eng.run("centre")
        outs.append(eng.readout())
        eng.close()
    for out in outs:
        assert (out["visits"] == m["visits"][idx]).all() and (out["best"] == m["best"][idx]).all()
        assert (bits(out["vsum"]) == bits(m["vsum"][idx])).all()
        assert (bits(out["vpolicy"]) == bits(m["vpolicy"][idx])).all()


def test_pass_length_policy_changes_work_not_results(monkeypatch):
    """stop fraction / cycle limit / re-visit budget decide how a generation is cut into passes, never what is played"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    recs = []
    for env in ({"C4_STOP_FRAC": "0"}, {"C4_STOP_FRAC": "0.5"}, {"C4_STOP_FRAC": "0.1", "C4_CYCLE_LIMIT": "20000"},
                {"C4_STOP_FRAC": "0.9", "C4_BUDGET": "3", "C4_CYCLE_LIMIT": "0"},
                {"C4_POOLS": "2", "C4_NET_CTAS": "96"}, {"C4_POOLS": "2", "C4_STOP_FRAC": "0.3", "C4_NET_CTAS": "148"}):
        for k in ("C4_STOP_FRAC", "C4_CYCLE_LIMIT", "C4_BUDGET", "C4_POOLS", "C4_NET_CTAS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        pool = SelfPlayPool(model, _cfg(96, 0.3, 0.25, 6), concurrent_games=48, seed=3)
        recs.append(_sorted(pool.generate_records(96)))
        pool.engine.close()
    for r in recs[1:]:
        _same_records(recs[0], r)


def test_empty_and_small_generations():
    from connect4_b200.neural.game_pool import SelfPlayPool
    pool = SelfPlayPool("centre", _cfg(30), concurrent_games=8)
    assert len(pool.generate_records(0)) == 0
    assert pool.generate(0) == []
    few = _sorted(pool.generate_records(3))                     # fewer games than slots
    assert sorted(set(few["game_id"].tolist())) == [0, 1, 2]
    many = _sorted(pool.generate_records(20))                   # slots re-seeded
    for g in range(3):                                          # deterministic player: every game is the same game
        a, b = few[few["game_id"] == g], many[many["game_id"] == g]
        assert a["move"].tolist() == b["move"].tolist() == few[few["game_id"] == 0]["move"].tolist()
    pool.engine.close()


def test_search_of_zero_and_one_positions():
    from connect4_b200.engine import Engine
    eng = Engine(4, _cfg(50))
    eng.begin(np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    eng.run("centre")
    assert eng.readout(0)["visits"].shape == (0, 7)
    eng.begin(np.zeros(1, np.uint64), np.zeros(1, np.uint64))
    eng.run("centre")
    out = eng.readout()
    assert int(out["root_visits"][0]) == 51 and int(out["visits"][0].sum()) == 50
    eng.close()


def test_evaluation_pass_with_ragged_batches():
    import torch
    from connect4_b200.board import BoardBatch
    from connect4_b200.neural.data import Connect4Dataset
    z = np.load(os.path.join(GOLDEN, "eval_stats.npz"))
    model = _model()
    n = 1000
    planes = BoardBatch(z["c0"][:n], z["c1"][:n]).to_planes("float32").cpu()
    ds = Connect4Dataset(planes, torch.as_tensor(z["values"][:n]), torch.as_tensor(z["priors"][:n]))
    a = model.evaluate(ds, batch_size=4096, shuffle=False)      # one short batch
    b = model.evaluate(ds, batch_size=333, shuffle=False)       # 3 full batches + 1 of one position
    c = model.evaluate(ds, batch_size=7, shuffle=True)
    for s in (b, c):
        assert s.value_stats.n == a.value_stats.n == n
        assert s.value_stats.correct == a.value_stats.correct and s.value_stats.total == a.value_stats.total
        assert s.prior_stats.correct == a.prior_stats.correct
        assert s.loss == pytest.approx(a.loss, rel=1e-5)
    assert model.evaluate_value_only(Connect4Dataset(planes, ds.values, None)).n == n


def test_full_size_generation_properties():
    """BASELINE configs[2] at its real size -- 4,096 concurrent games, 800 simulations per move, the example network,
    AlphaZero noise -- checked through size-independent properties on all ~130k records (vectorised on the host with the
    device bitboard engine): every game is a chain of legal moves from the empty board, boards follow from the moves,
    the result is the board's result, policies are distributions over the legal moves, the flip-augmented dataset mirrors
    itself, and the sharding of the same generation over two simulated ranks reproduces it game for game."""
    import torch
    from connect4_b200.board import BoardBatch
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    cfg = _cfg(800, 0.3, 0.25, 6)
    pool = SelfPlayPool(model, cfg, concurrent_games=4096, seed=7)
    rec = _sorted(pool.generate_records(4096))
    gid, ply = rec["game_id"], rec["ply"].astype(np.int64)
    assert sorted(set(gid.tolist())) == list(range(4096))
    first = ply == 0
    assert (rec["c0"][first] == 0).all() and (rec["c1"][first] == 0).all()
    n_moves = rec["n_moves"].astype(np.int64)
    last = ply == n_moves - 1
    assert first.sum() == last.sum() == 4096 and n_moves.min() >= 7 and n_moves.max() <= 42
    # replay every move on the device: board after move k == board before move k+1; the last move ends the game
    bb = BoardBatch(rec["c0"].copy(), rec["c1"].copy())
    legal = bb.legal_mask().cpu().numpy()
    assert ((legal >> rec["move"].astype(np.int64)) & 1).all()
    res_after = bb.drop(rec["move"]).cpu().numpy()
    a0, a1 = bb.numpy()
    nxt = np.flatnonzero(~last)
    assert (a0[nxt] == rec["c0"][nxt + 1]).all() and (a1[nxt] == rec["c1"][nxt + 1]).all()
    assert (res_after[~last] == -1).all() and (res_after[last] == rec["result"][last]).all()
    per_game_result = rec["result"][last][np.searchsorted(gid[last], gid)]
    assert (rec["result"] == per_game_result).all() and (rec["result_value"] == per_game_result * 0.5).all()
    # policy targets: distributions over the legal columns
    pol = rec["policy"]
    assert np.abs(pol.sum(1) - 1.0).max() < 1e-5 and (pol >= 0).all()
    illegal = ((legal[:, None] >> np.arange(7)[None, :]) & 1) == 0
    assert (pol[illegal] == 0).all()
    # the sink: second half of the dataset = mirror of the first
    b, v, p = pool.last_dataset()
    n = len(rec)
    assert torch.equal(b[n:], torch.flip(b[:n], dims=[3])) and torch.equal(p[n:], torch.flip(p[:n], dims=[1])) and torch.equal(v[n:], v[:n])
    pool.engine.close()
    # the same generation played as two "ranks" (games g % 2 == r) on smaller pools is the same set of records
    parts = []
    for r in range(2):
        q = SelfPlayPool(model, cfg, concurrent_games=1024, seed=7)
        parts.append(q.generate_records(2048, game_id_base=r, game_id_stride=2))
        q.engine.close()
    _same_records(rec, _sorted(np.concatenate(parts)))


def test_selfplay_without_rng_is_refused_when_the_config_needs_one():
    """a config with root noise / sampled moves and no RNG would make every slot play the same game: refused"""
    from connect4_b200._lib import C4Error
    from connect4_b200.engine import Engine
    from connect4_b200.mcts import MCTSConfig
    eng = Engine(4, MCTSConfig(8, 19652, 1.25, 0.3, 0.25, 6))
    with pytest.raises(C4Error, match="RNG"):
        eng.selfplay(2, "centre")
    eng.set_rng("philox", seed=1)
    assert len(eng.selfplay(2, "centre")) >= 14
    eng.close()


def test_search_engine_cache_is_bounded():
    """mcts.search caches its device contexts in a small LRU and closes what it evicts (one leaked context per
    network / configuration)"""
    from connect4_b200 import mcts
    from connect4_b200.board import Board
    from connect4_b200.evaluators import Evaluator, evaluate_centre_with_prior
    mcts.release_engines()
    closed = []
    for sims in range(5, 5 + mcts._MAX_ENGINES + 3):
        t = mcts.search(mcts.MCTSConfig(sims), Board(), Evaluator(evaluate_centre_with_prior))
        assert t.root.data.search_value.visit_count == sims + 1
        closed.append(len(mcts._ENGINES))
    assert max(closed) == mcts._MAX_ENGINES
    first = next(iter(mcts._ENGINES.values()))
    mcts.release_engines()
    assert not mcts._ENGINES and first.h is None

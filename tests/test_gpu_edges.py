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

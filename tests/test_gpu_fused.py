"""The fused persistent engine (csrc/c4_fused.cu: one launch per generation, tree warps and the tcgen05 tower on the same
SM, CTA-local leaf ring) against the lock-step pass engine (csrc/c4_search.cu) -- both run the reference's search
(oinkoink/mcts.py:94-121) and game loop (neural/training_game.py:8-19) with the same device functions, so every record and
every root read-out must be identical bit for bit -- and against the reference's recorded games."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, golden, random_positions

pytestmark = pytest.mark.gpu


def _cfg(sims, alpha=0.3, frac=0.25, sampling=6):
    from connect4_b200.mcts import MCTSConfig
    return MCTSConfig(sims, 19652, 1.25, alpha, frac, sampling)


def _model(**kw):
    from oracle import net_ref as nr
    from connect4_b200.neural.model import ModelWrapper
    return ModelWrapper(state_dict=nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz")), **kw)


def _sorted(rec):
    return rec[np.lexsort((rec["ply"], rec["game_id"]))]


def _same(a, b):
    assert len(a) == len(b)
    for f in a.dtype.names:                       # field by field (numpy leaves a record's padding bytes undefined)
        assert a[f].tobytes() == b[f].tobytes(), f


def _generate(monkeypatch, engine, model, cfg, slots, n, seed=3, **kw):
    from connect4_b200.neural.game_pool import SelfPlayPool
    monkeypatch.setenv("C4_ENGINE", engine)
    pool = SelfPlayPool(model, cfg, concurrent_games=slots, seed=seed)
    rec = _sorted(pool.generate_records(n, **kw))
    pool.engine.close()
    return rec


@pytest.mark.parametrize("slots,sims,n", [(1, 24, 3), (8, 16, 8), (64, 64, 200), (149, 40, 300), (300, 200, 450)])
def test_fused_and_lockstep_generations_are_identical(monkeypatch, slots, sims, n):
    """pool sizes below / at / above one game per SM, games re-seeded, AlphaZero noise + sampled moves"""
    model = _model()
    a = _generate(monkeypatch, "fused", model, _cfg(sims), slots, n)
    b = _generate(monkeypatch, "lockstep", model, _cfg(sims), slots, n)
    assert sorted(set(a["game_id"].tolist())) == list(range(n))
    _same(a, b)


def test_fused_generation_with_start_positions_and_bf16_operands(monkeypatch):
    model = _model(operand_dtype="bf16")
    c0, c1 = random_positions(5, 40, max_plies=12)
    kw = dict(start=(c0, c1), game_id_base=7, game_id_stride=3)
    a = _generate(monkeypatch, "fused", model, _cfg(48), 16, 40, **kw)
    b = _generate(monkeypatch, "lockstep", model, _cfg(48), 16, 40, **kw)
    assert sorted(set(a["game_id"].tolist())) == [7 + 3 * i for i in range(40)]
    first = a[a["ply"] == 0]
    assert first["c0"].tolist() == [int(x) for x in c0] and first["c1"].tolist() == [int(x) for x in c1]
    _same(a, b)


def test_fused_generation_does_not_depend_on_the_memo_or_the_tuning_knobs(monkeypatch):
    """the evaluation memo, the shared-memory PUCT tables and the number of working tree warps change the work done, never
    the records"""
    model = _model()
    recs = []
    for env in ({}, {"C4_MEMO_LOG2": "0"}, {"C4_FZ_SMEM_TABLES": "1"}, {"C4_FZ_TREE_WARPS": "3"}):
        for k in ("C4_MEMO_LOG2", "C4_FZ_SMEM_TABLES", "C4_FZ_TREE_WARPS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        recs.append(_generate(monkeypatch, "fused", model, _cfg(64), 40, 100, seed=11))
    for r in recs[1:]:
        _same(recs[0], r)


def test_fused_search_batch_equals_lockstep_and_oracle(monkeypatch, oracle):
    """stand-alone searches (MCTS.make_move protocol, c4_search_run NET): whole batch in one persistent launch"""
    from connect4_b200.engine import Engine
    model = _model()
    c0, c1 = random_positions(17, 200)
    outs = {}
    for engine in ("fused", "lockstep"):
        monkeypatch.setenv("C4_ENGINE", engine)
        eng = Engine(256, _cfg(150, 0.0, 0.0, 0))
        eng.set_net(model)
        eng.begin(c0, c1)
        eng.run("net")
        outs[engine] = eng.readout()
        eng.close()
    a, b = outs["fused"], outs["lockstep"]
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k
    assert (a["root_visits"] == 151).all()

    def ev(x, y):
        v, p = model.evaluate_bitboards(np.array([x], np.uint64), np.array([y], np.uint64))
        return float(v.cpu().numpy()[0]), p.cpu().numpy()[0]
    for i in range(6):                                   # and the oracle's search fed with the CUDA network's outputs
        t = oracle.Tree(oracle.make_config(150), int(c0[i]), int(c1[i])).search(ev)
        v, s, r, _ = t.root_children()
        assert (v == a["visits"][i]).all() and (bits(s) == bits(a["vsum"][i])).all() and t.best_move() == a["best"][i]


def test_stream_api_on_both_engines(monkeypatch):
    """c4_selfplay_stream: cold start until N games have finished, continue by time, then a normal generation on the
    same context (the pool state of a stopped stream is complete: no request is left in flight)"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    for engine in ("fused", "lockstep"):
        monkeypatch.setenv("C4_ENGINE", engine)
        pool = SelfPlayPool(model, _cfg(32), concurrent_games=96, seed=2)
        r = pool.stream(stop_games=96, reset=True, cold_memo=True)
        assert r["engine"] == engine and r["games"] >= 96 and r["positions"] >= 96 * 7 and r["evals"] > 0
        assert r["device_ms"] > 0
        r2 = pool.stream(max_ms=30.0)                    # continues the same games
        assert r2["positions"] > 0 and 25.0 <= r2["device_ms"] < 400.0
        r3 = pool.stream(stop_games=10)
        assert r3["games"] >= 10
        rec = pool.generate_records(50)                  # a fresh generation after a stopped stream
        assert sorted(set(rec["game_id"].tolist())) == list(range(50)) and (rec["result"] >= 0).all()
        pool.engine.close()
    with pytest.raises(Exception):
        SelfPlayPool(model, _cfg(8), concurrent_games=4).stream()          # needs a game count or a time limit


def test_fused_games_equal_reference_games_with_injected_randomness(monkeypatch):
    """the reference's recorded AlphaZero games (gamma draws and uniforms injected) cannot be replayed with a network
    evaluator -- they were played with the centre evaluator -- so the fused engine is pinned to the reference through
    the lock-step engine (test above) and through the NN-guided oracle search; here: a fused generation is a legal,
    self-consistent set of games"""
    from oracle import c4oracle as o
    model = _model()
    rec = _generate(monkeypatch, "fused", model, _cfg(40), 200, 400, seed=9)
    for g in range(400):
        recs = rec[rec["game_id"] == g]
        c0 = c1 = 0
        res = -1
        for k, r in enumerate(recs):
            assert (int(r["c0"]), int(r["c1"]), int(r["ply"])) == (c0, c1, k) and res == -1
            assert o.legal_mask(c0, c1) >> int(r["move"]) & 1
            c0, c1, res = o.drop(c0, c1, int(r["move"]))
        assert res != -1 and res == int(recs[-1]["result"]) and int(recs[0]["n_moves"]) == len(recs)


@pytest.mark.parametrize("slots,sims,n", [(6, 16, 10), (150, 32, 300)])
def test_fused_engine_with_the_64_filter_network(monkeypatch, slots, sims, n):
    """the reference's example_config network (64 filters / 6 residual blocks / 6 fc layers, oinkoink/data/example_config.py)
    runs on the fused engine with the 64-filter tower geometry (6-board strips, weight stage refilled slice by slice)"""
    import torch
    from connect4_b200.neural.config import ModelConfig, NetConfig
    from connect4_b200.neural.model import ModelWrapper
    torch.manual_seed(0)
    model = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
    a = _generate(monkeypatch, "fused", model, _cfg(sims), slots, n)
    b = _generate(monkeypatch, "lockstep", model, _cfg(sims), slots, n)
    assert sorted(set(a["game_id"].tolist())) == list(range(n))
    _same(a, b)


def test_generation_handed_over_from_lockstep_to_fused_for_the_drain(monkeypatch):
    """auto engine policy without the split engine (what a 64-filter network gets) on a pool with more than 16 games per
    SM: the lock-step engine (with de-duplication of the evaluations in flight) plays the bulk, the drain -- few games
    left -- is handed to the fused engine; the records are those of either engine alone"""
    import torch
    model = _model()
    slots = 16 * torch.cuda.get_device_properties(0).multi_processor_count + 300
    monkeypatch.delenv("C4_ENGINE", raising=False)
    monkeypatch.setenv("C4_SP_DISABLE", "1")
    from connect4_b200.neural.game_pool import SelfPlayPool
    pool = SelfPlayPool(model, _cfg(12), concurrent_games=slots, seed=3)
    auto = _sorted(pool.generate_records(slots + 500))
    pool.engine.close()
    lock = _generate(monkeypatch, "lockstep", model, _cfg(12), slots, slots + 500)
    assert sorted(set(auto["game_id"].tolist())) == list(range(slots + 500))
    _same(auto, lock)


def test_deduplication_of_evaluations_in_flight_changes_work_not_results(monkeypatch):
    """lock-step engine: a game that misses the memo claims the entry, later askers wait for its answer (c4_tree.cuh):
    fewer network evaluations, identical records"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    monkeypatch.setenv("C4_ENGINE", "lockstep")
    out = {}
    for dedup in (True, False):
        if dedup:
            monkeypatch.delenv("C4_MEMO_NO_DEDUP", raising=False)
        else:
            monkeypatch.setenv("C4_MEMO_NO_DEDUP", "1")
        pool = SelfPlayPool(model, _cfg(48), concurrent_games=256, seed=21)
        r = pool.stream(stop_games=256, reset=True, cold_memo=True)
        rec = _sorted(pool.generate_records(300))
        pool.engine.close()
        out[dedup] = (r["evals"], rec)
    assert out[True][0] < out[False][0]                  # fewer evaluations (-27 % at the benchmark size, a few % here)
    _same(out[True][1], out[False][1])

"""The split persistent engine (csrc/c4_split.cu: tree CTAs and tcgen05 tower CTAs on separate SMs, one leaf ring per
tower, de-duplication of the evaluations in flight) against the lock-step pass engine (csrc/c4_search.cu).  Both run the
reference's search (oinkoink/mcts.py:94-121) and game loop (neural/training_game.py:8-19) with the same device functions,
so every record and every root read-out must be identical bit for bit whatever the engine, the number of tower CTAs, the
memo size or the de-duplication do to the ORDER in which the work is done."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, random_positions

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
    monkeypatch.setenv("C4_FZ_TIMEOUT_S", "120")
    pool = SelfPlayPool(model, cfg, concurrent_games=slots, seed=seed)
    rec = _sorted(pool.generate_records(n, **kw))
    pool.engine.close()
    return rec


@pytest.mark.parametrize("slots,sims,n", [(1, 24, 3), (8, 16, 8), (64, 64, 200), (149, 40, 300), (300, 200, 450), (1000, 60, 1300)])
def test_split_and_lockstep_generations_are_identical(monkeypatch, slots, sims, n):
    """pool sizes below / at / above one game per tree CTA, games re-seeded, AlphaZero noise + sampled moves"""
    model = _model()
    a = _generate(monkeypatch, "split", model, _cfg(sims), slots, n)
    b = _generate(monkeypatch, "lockstep", model, _cfg(sims), slots, n)
    assert sorted(set(a["game_id"].tolist())) == list(range(n))
    _same(a, b)


def test_split_generation_with_start_positions_and_bf16_operands(monkeypatch):
    model = _model(operand_dtype="bf16")
    c0, c1 = random_positions(5, 40, max_plies=12)
    kw = dict(start=(c0, c1), game_id_base=7, game_id_stride=3)
    a = _generate(monkeypatch, "split", model, _cfg(48), 16, 40, **kw)
    b = _generate(monkeypatch, "lockstep", model, _cfg(48), 16, 40, **kw)
    assert sorted(set(a["game_id"].tolist())) == [7 + 3 * i for i in range(40)]
    _same(a, b)


def test_split_generation_does_not_depend_on_towers_memo_or_deduplication(monkeypatch):
    """the number of tower CTAs (3 .. 120 of 148 SMs), the evaluation memo, the de-duplication of evaluations in flight and
    the launch form (one launch with roles by CTA index / two launches) change who does which work when, never the records"""
    model = _model()
    recs = []
    knobs = ("C4_SP_NET_CTAS", "C4_MEMO_LOG2", "C4_MEMO_NO_DEDUP", "C4_SP_LAUNCH")
    for env in ({}, {"C4_SP_NET_CTAS": "3"}, {"C4_SP_NET_CTAS": "120"}, {"C4_MEMO_LOG2": "0"}, {"C4_MEMO_NO_DEDUP": "1"},
                {"C4_MEMO_LOG2": "8"},                        # 256 entries: colliding keys overwrite PENDING tags all the time
                {"C4_SP_LAUNCH": "two"}):                     # tree and tower kernels as two launches on two streams
        for k in knobs:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        recs.append(_generate(monkeypatch, "split", model, _cfg(64), 200, 320, seed=11))
    for k in knobs:
        monkeypatch.delenv(k, raising=False)
    for r in recs[1:]:
        _same(recs[0], r)


def test_adaptive_tower_count_and_shared_memory_tables_change_work_not_results(monkeypatch):
    """from 1,024 game slots a run of the split engine is cut into time slices and every slice is launched with the tower count
    the previous slice's load signals ask for (C4_SP_ADAPT = slice length in ms; 0 = constant count); the tree CTAs read the
    PUCT tables from shared memory (C4_SP_SMEM_TABLES).  A game may be stopped and continued by another launch, on another
    tree CTA, any number of times: the records stay those of the lock-step engine"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    knobs = ("C4_SP_ADAPT", "C4_SP_SMEM_TABLES", "C4_SP_NET_CTAS")
    recs, launches = [], []
    for env in ({"C4_SP_ADAPT": "0"}, {"C4_SP_ADAPT": "2"}, {}, {"C4_SP_ADAPT": "2", "C4_SP_SMEM_TABLES": "0"}):
        for k in knobs:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        recs.append(_generate(monkeypatch, "split", model, _cfg(64), 1100, 2600, seed=5))
        # the stream interface reports the kernels it launched: one per slice (+ two bookkeeping kernels)
        monkeypatch.setenv("C4_ENGINE", "split")
        pool = SelfPlayPool(model, _cfg(64), concurrent_games=1100, seed=5)
        r = pool.stream(max_ms=40.0, reset=True, cold_memo=True)
        pool.engine.close()
        assert r["engine"] == "split" and 30.0 <= r["device_ms"] < 400.0
        launches.append(r["launches"])
    for k in knobs:
        monkeypatch.delenv(k, raising=False)
    assert launches[0] == 3 and launches[1] >= 10 and launches[3] >= 10, launches
    for r in recs[1:]:
        _same(recs[0], r)
    _same(recs[0], _generate(monkeypatch, "lockstep", model, _cfg(64), 1100, 2600, seed=5))


def test_large_search_batch_in_slices_and_searches_beyond_the_shared_memory_tables(monkeypatch):
    """(a) a stand-alone search batch of more than 1,024 positions runs in time slices with the adaptive tower count like a
    generation does; (b) searches of more simulations than the shared-memory PUCT tables hold (4,096 entries) fall back to the
    tables in HBM.  Root read-outs / records equal to the lock-step engine's"""
    from connect4_b200.engine import Engine
    model = _model()
    c0, c1 = random_positions(23, 1100)
    outs = {}
    for engine, adapt in (("split", "1.5"), ("split", "0"), ("lockstep", "0")):
        monkeypatch.setenv("C4_ENGINE", engine)
        monkeypatch.setenv("C4_SP_ADAPT", adapt)                       # "1.5" = slices of 1.5 ms
        eng = Engine(1100, _cfg(120, 0.0, 0.0, 0))
        eng.set_net(model)
        eng.begin(c0, c1)
        eng.run("net")
        outs[(engine, adapt)] = eng.readout()
        eng.close()
    monkeypatch.delenv("C4_SP_ADAPT", raising=False)
    ref = outs[("lockstep", "0")]
    assert (ref["root_visits"] == 121).all()
    for k in ref:
        assert outs[("split", "1.5")][k].tobytes() == ref[k].tobytes(), k
        assert outs[("split", "0")][k].tobytes() == ref[k].tobytes(), k
    a = _generate(monkeypatch, "split", model, _cfg(4200), 2, 2, seed=9)
    b = _generate(monkeypatch, "lockstep", model, _cfg(4200), 2, 2, seed=9)
    _same(a, b)


def test_split_search_batch_equals_lockstep_and_oracle(monkeypatch, oracle):
    """stand-alone searches (MCTS.make_move protocol, c4_search_run NET): the whole batch in one persistent launch pair"""
    from connect4_b200.engine import Engine
    model = _model()
    c0, c1 = random_positions(17, 200)
    outs = {}
    for engine in ("split", "lockstep"):
        monkeypatch.setenv("C4_ENGINE", engine)
        eng = Engine(256, _cfg(150, 0.0, 0.0, 0))
        eng.set_net(model)
        eng.begin(c0, c1)
        eng.run("net")
        outs[engine] = eng.readout()
        eng.close()
    a, b = outs["split"], outs["lockstep"]
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k
    assert (a["root_visits"] == 151).all()

    def ev(x, y):
        v, p = model.evaluate_bitboards(np.array([x], np.uint64), np.array([y], np.uint64))
        return float(v.cpu().numpy()[0]), p.cpu().numpy()[0]
    for i in range(4):                                   # and the oracle's search fed with the CUDA network's outputs
        t = oracle.Tree(oracle.make_config(150), int(c0[i]), int(c1[i])).search(ev)
        v, s, r, _ = t.root_children()
        assert (v == a["visits"][i]).all() and (bits(s) == bits(a["vsum"][i])).all() and t.best_move() == a["best"][i]


def test_stream_api_and_auto_policy(monkeypatch):
    """auto policy: the split engine for 32-filter networks at every pool size; c4_selfplay_stream: cold start until N games
    have finished, continue by time, then a normal generation on the same context (the pool state of a stopped stream is
    complete: no request is left in flight, games parked on another game's evaluation are simply re-run)"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    monkeypatch.delenv("C4_ENGINE", raising=False)
    pool = SelfPlayPool(model, _cfg(32), concurrent_games=96, seed=2)
    r = pool.stream(stop_games=96, reset=True, cold_memo=True)
    assert r["engine"] == "split" and r["games"] >= 96 and r["positions"] >= 96 * 7 and r["evals"] > 0
    assert r["device_ms"] > 0
    r2 = pool.stream(max_ms=30.0)                        # continues the same games
    assert r2["positions"] > 0 and 25.0 <= r2["device_ms"] < 400.0
    r3 = pool.stream(stop_games=10)
    assert r3["games"] >= 10
    rec = _sorted(pool.generate_records(50))             # a fresh generation after a stopped stream
    pool.engine.close()
    assert sorted(set(rec["game_id"].tolist())) == list(range(50)) and (rec["result"] >= 0).all()
    lock = _generate(monkeypatch, "lockstep", model, _cfg(32), 96, 50, seed=2)
    _same(rec, lock)


def test_deduplication_in_the_split_engine_changes_work_not_results(monkeypatch):
    """a game whose leaf is being evaluated for another game parks it and takes the owner's answer from the memo: fewer
    network evaluations, identical records"""
    from connect4_b200.neural.game_pool import SelfPlayPool
    model = _model()
    monkeypatch.setenv("C4_ENGINE", "split")
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
    monkeypatch.delenv("C4_MEMO_NO_DEDUP", raising=False)
    assert out[True][0] < out[False][0]
    _same(out[True][1], out[False][1])


@pytest.mark.parametrize("slots,sims,n", [(6, 16, 10), (150, 32, 300)])
def test_split_engine_with_the_64_filter_network(monkeypatch, slots, sims, n):
    """the reference's example_config network (64 filters / 6 residual blocks / 6 fc layers, oinkoink/data/example_config.py):
    the tower CTAs run the batch kernel's 64-filter geometry (6-board strips, weight stage refilled slice by slice)"""
    import torch
    from connect4_b200.neural.config import ModelConfig, NetConfig
    from connect4_b200.neural.model import ModelWrapper
    torch.manual_seed(0)
    model = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
    a = _generate(monkeypatch, "split", model, _cfg(sims), slots, n)
    b = _generate(monkeypatch, "lockstep", model, _cfg(sims), slots, n)
    assert sorted(set(a["game_id"].tolist())) == list(range(n))
    _same(a, b)

"""GPU tree kernels vs the oracle and the reference's goldens: bit-exact visit counts, fp64 value sums, policies,
chosen moves and node counts under the deterministic evaluator (BASELINE.json configs[1])."""
import numpy as np
import pytest

from conftest import bits, golden, random_positions

pytestmark = pytest.mark.gpu


class Cfg:
    def __init__(self, simulations, pb_c_base=19652, pb_c_init=1.25, alpha=0.0, frac=0.0, sampling=0):
        self.simulations = simulations
        self.pb_c_base = pb_c_base
        self.pb_c_init = pb_c_init
        self.root_dirichlet_alpha = alpha
        self.root_exploration_fraction = frac
        self.num_sampling_moves = sampling


def _compare(out, m, idx=None, keys_i=("visits", "cres", "best", "nodes", "root_visits"),
             keys_f=("vsum", "best_value", "vpolicy", "root_vsum")):
    sel = slice(None) if idx is None else idx
    for k in keys_i:
        assert (out[k] == m[k][sel]).all(), k
    for k in keys_f:
        bad = np.flatnonzero((bits(out[k]) != bits(m[k][sel])).reshape(len(out[k]), -1).any(axis=1))
        assert len(bad) == 0, (k, bad[:5])


def test_sweep_10k_positions_800_sims():
    """the headline parity config: every one of the 10k reference searches reproduced bit-for-bit, one launch"""
    from connect4_b200.engine import Engine
    m = golden("mcts_sweep_800.npz")
    n = len(m["c0"])
    eng = Engine(n, Cfg(800))
    eng.begin(m["c0"], m["c1"])
    eng.run("centre")
    out = eng.readout()
    _compare(out, m)
    eng.close()


def test_small_simulation_counts():
    from connect4_b200.engine import Engine
    m = golden("mcts_small.npz")
    for sims in sorted(set(m["sims"].tolist())):
        idx = np.flatnonzero(m["sims"] == sims)
        eng = Engine(len(idx), Cfg(sims))
        eng.begin(m["c0"][idx], m["c1"][idx])
        eng.run("centre")
        out = eng.readout()
        _compare(out, m, idx, keys_f=("vsum", "best_value", "vpolicy", "root_vsum", "cpolicy", "root_prior"))
        eng.close()


def test_reference_player_kats():
    """reference tests/player_test.py:151-179: 7 tactical positions, pb_c_init=9999, sims 7**plies+1 / 2**plies"""
    from connect4_b200.engine import Engine
    m = golden("mcts_kat.npz")
    for i in range(7):
        eng = Engine(1, Cfg(int(m["sims"][i]), 19652, 9999))
        eng.begin(m["c0"][i:i + 1], m["c1"][i:i + 1])
        eng.run("centre")
        out = eng.readout()
        _compare(out, m, slice(i, i + 1))
        assert (int(m["ans_mask"][i]) >> int(out["best"][0])) & 1
        eng.close()


def test_root_noise_injected():
    from connect4_b200.engine import Engine
    m = golden("mcts_noise.npz")
    for sims in sorted(set(m["sims"].tolist())):
        idx = np.flatnonzero(m["sims"] == sims)
        eng = Engine(len(idx), Cfg(sims, alpha=0.3, frac=0.25))
        eng.set_rng("injected", noise=m["noise"][idx][:, None, :], uniform=np.zeros((len(idx), 1)))
        eng.begin(m["c0"][idx], m["c1"][idx])
        eng.run("centre")
        out = eng.readout()
        _compare(out, m, idx, keys_f=("vsum", "best_value", "vpolicy", "root_vsum", "root_prior"))
        eng.close()


def _perturbed_eval(dtype):
    """a deterministic evaluator that is NOT the built-in one: position-dependent value and non-uniform prior"""
    def f(c0, c1):
        c0 = np.asarray(c0, np.uint64)
        c1 = np.asarray(c1, np.uint64)
        h = (c0 * np.uint64(0x9E3779B97F4A7C15) ^ (c1 * np.uint64(0xC2B2AE3D27D4EB4F))) >> np.uint64(11)
        v = (h % np.uint64(1000)).astype(np.float64) / 999.0
        pr = np.stack([((h >> np.uint64(3 * k)) % np.uint64(17)).astype(np.float64) + 1.0 for k in range(7)], axis=1)
        pr = pr / pr.sum(axis=1, keepdims=True)
        return v, pr.astype(dtype)
    return f


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_external_evaluator_stepping(oracle, dtype):
    """c4_search_pending / c4_search_supply with an arbitrary host evaluator vs the oracle driven by the same function"""
    from connect4_b200.engine import Engine
    c0, c1 = random_positions(11, 96)
    ev = _perturbed_eval(dtype)
    sims = 150
    eng = Engine(128, Cfg(sims))
    eng.begin(c0, c1)
    eng.run_external(ev)
    out = eng.readout()
    for i in range(len(c0)):
        t = oracle.Tree(oracle.make_config(sims), int(c0[i]), int(c1[i]))

        def one(a, b):
            v, p = ev(np.array([a], np.uint64), np.array([b], np.uint64))
            return float(v[0]), p[0]
        t.search(one)
        v, s, r, a = t.root_children()
        assert (v == out["visits"][i]).all(), i
        assert (bits(s) == bits(out["vsum"][i])).all(), i
        assert t.best_move() == out["best"][i]
        assert (bits(t.values_policy()) == bits(out["vpolicy"][i])).all()
        assert t.root_stats()[3] == out["nodes"][i]
    eng.close()


def test_whole_tree_matches_oracle(oracle):
    """not just the root: every node of the exported pool equals the oracle's tree (visits, fp64 sums, results)"""
    from connect4_b200.board import Board
    from connect4_b200.engine import Engine
    from connect4_b200.tree import Tree
    c0, c1 = random_positions(21, 6)
    sims = 300
    eng = Engine(8, Cfg(sims))
    eng.begin(c0, c1)
    eng.run("centre")
    for i in range(len(c0)):
        tree = Tree(Board.from_bitboards(c0[i], c1[i]), eng.export_tree(i))
        d = oracle.Tree(oracle.make_config(sims), int(c0[i]), int(c1[i])).search_centre().dump()
        paths = [()]
        want = {}
        for j in range(len(d["parent"])):
            if j > 0:
                paths.append(paths[int(d["parent"][j])] + (int(d["name"][j]),))
            want[paths[j]] = (int(d["c0"][j]), int(d["c1"][j]), int(d["visits"][j]), float(d["vsum"][j]),
                              int(d["result"][j]))
        got = {}

        def walk(n, path):
            b = n.data.board
            sv = n.data.search_value
            got[path] = (int(b.color[0]), int(b.color[1]), sv.visit_count if sv else 0, sv.value_sum if sv else 0.0,
                         -1 if b.result is None else int(b.result.value * 2))
            for c in n.children:
                walk(c, path + (c.name,))
        walk(tree.root, ())
        assert len(got) == len(want) == tree.count_nodes()
        assert got == want
    eng.close()


def test_engine_reuse_and_partial_batches():
    """a context is reused across searches of different sizes (pool reset, idle slots stay idle)"""
    from connect4_b200.engine import Engine
    m = golden("mcts_sweep_800.npz")
    eng = Engine(64, Cfg(800))
    for lo, hi in ((0, 64), (64, 70), (70, 71), (100, 164)):
        eng.begin(m["c0"][lo:hi], m["c1"][lo:hi])
        eng.run("centre")
        _compare(eng.readout(), m, slice(lo, hi))
    eng.close()

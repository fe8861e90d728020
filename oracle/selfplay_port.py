"""CPU port of the reference's self-play generation path, used ONLY as the timed CPU baseline of bench.py
(`cpu_baseline` leg and `--impl reference` arm) -- TEST/BENCH INFRASTRUCTURE, never imported by the product.

What it is: the reference's algorithm (oinkoink/neural/training_game.py:8-19 over mcts.py:78-121, AlphaZero settings of
neural/training.py:209-216) with the tree work done by the C oracle (oracle/c4_oracle.c) and the network by the fp32
torch restatement (oracle/net_ref.py) on the host cores.  Like the reference's game_pool + InferenceServer
(neural/game_pool.py:15-49, inference_server.py:37-63) each worker keeps a set of games in flight and evaluates their
pending leaves as one batch, and memoises evaluations in a per-process table exactly like the reference's
Evaluator.position_table (evaluators.py:18-25, game_pool.py:21-27); one single-threaded worker process per core.  It is a
faster arrangement than the reference's own Python (C tree code, no pipes, no GIL): the survey measured the unmodified
reference at 1.08 positions/s per core and 13 positions/s on 8 cores, this port does many times that.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_STATE = {}


def _init(state_path, filters, n_fc, n_res, sims, games_per_worker, seed):
    import torch
    from oracle import c4oracle as o
    from oracle import net_ref as nr
    torch.set_num_threads(1)
    o.lib()
    if state_path:
        sd = nr.load_golden_state(state_path)
    else:
        sd = nr.random_state(0, filters, n_fc, n_res)
    sd = {k: torch.as_tensor(v) for k, v in sd.items()}
    _STATE.update(sd=sd, sims=sims, G=games_per_worker, o=o, nr=nr, torch=torch,
                  rng=np.random.default_rng(seed + os.getpid()), games=None)


def _new_search(g):
    S = _STATE
    o = S["o"]
    t = o.Tree(S["cfg"], g["c0"], g["c1"])
    t.set_noise(S["rng"].gamma(0.3, 1.0, 7))
    st, a, b = t.start()
    g.update(tree=t, leaf=(a, b))


def _finish_search(g):
    """search finished: play the move (mcts.py:81-86); start the next search / the next game"""
    S = _STATE
    o = S["o"]
    t = g["tree"]
    age = o.age(g["c0"], g["c1"])
    mv = t.sample_move(S["rng"].random()) if age < 6 else t.best_move()
    g["c0"], g["c1"], res = o.drop(g["c0"], g["c1"], mv)
    if res != o.RES_NONE:
        g["c0"], g["c1"] = 0, 0
    _new_search(g)


def _advance(g):
    """run game g until its pending leaf is NOT in the evaluation memo (Evaluator.position_table,
    oinkoink/evaluators.py:18-25: one table per process shared by all its games); returns moves played"""
    S = _STATE
    o, table = S["o"], S["table"]
    moves = hits = 0
    while True:
        hit = table.get(g["leaf"])
        if hit is None:
            return moves, hits
        hits += 1
        st, a, b = g["tree"].supply(hit[0], hit[1])
        if st == o.DONE:
            _finish_search(g)
            moves += 1
        else:
            g["leaf"] = (a, b)


def _step(budget_s):
    """advance this worker's games for ~budget_s seconds; returns (positions, evals, elapsed, memo hits)"""
    S = _STATE
    o, nr, torch = S["o"], S["nr"], S["torch"]
    if S["games"] is None:
        S["cfg"] = o.make_config(S["sims"], 19652, 1.25, 0.3, 0.25, 6)
        S["table"] = {}
        S["games"] = [dict(c0=0, c1=0) for _ in range(S["G"])]
        for g in S["games"]:
            _new_search(g)
    games, table = S["games"], S["table"]
    positions = evals = hits = 0
    t0 = time.perf_counter()
    with torch.no_grad():
        while time.perf_counter() - t0 < budget_s:
            for g in games:                                   # play through memo hits
                m, h = _advance(g)
                positions += m
                hits += h
            c0 = np.array([g["leaf"][0] for g in games], np.uint64)
            c1 = np.array([g["leaf"][1] for g in games], np.uint64)
            v, p = nr.forward_state(S["sd"], nr.planes_from_bitboards(c0, c1))
            v, p = v.numpy(), p.numpy()
            evals += len(games)
            for i, g in enumerate(games):
                table[g["leaf"]] = (float(v[i]), p[i].copy())
    return positions, evals, time.perf_counter() - t0, hits


class PortPool:
    """persistent worker pool: one single-threaded process per core, `games_per_worker` games in flight each"""

    def __init__(self, procs=None, sims=800, games_per_worker=32, state_path=None, filters=32, n_fc=4, n_res=3, seed=0):
        import multiprocessing as mp
        self.procs = procs or (os.cpu_count() or 1)
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.procs, initializer=_init,
                             initargs=(state_path, filters, n_fc, n_res, sims, games_per_worker, seed))
        self.sims = sims
        self.games_per_worker = games_per_worker

    def step(self, budget_s):
        """every worker plays for budget_s seconds; returns dict(positions, evals, seconds, positions_per_sec)"""
        t0 = time.perf_counter()
        res = self.pool.map(_step, [budget_s] * self.procs)
        wall = time.perf_counter() - t0
        pos = sum(r[0] for r in res)
        ev = sum(r[1] for r in res)
        hits = sum(r[3] for r in res)
        busy = max(r[2] for r in res)
        return dict(positions=pos, evals=ev, seconds=busy, wall=wall, positions_per_sec=pos / busy,
                    evals_per_sec=ev / busy, memo_hit_rate=hits / max(1, hits + ev))

    def close(self):
        self.pool.terminate()
        self.pool.join()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10)
    ap.add_argument("--procs", type=int, default=None)
    ap.add_argument("--sims", type=int, default=800)
    a = ap.parse_args()
    pp = PortPool(a.procs, a.sims, state_path=os.path.join(ROOT, "tests", "golden", "example_net_state.npz"))
    pp.step(1.0)
    print(pp.step(a.seconds))
    pp.close()

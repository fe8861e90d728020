#!/usr/bin/env python3
"""Timings of the rows the path was widened into: Match in lock step, the evaluation pass, the generation
sink.  CUDA events / wall clock around the public calls; prints one JSON line per row."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from functools import partial
from connect4_b200 import evaluators as evl
from connect4_b200.board import BoardBatch
from connect4_b200.engine import augment_pack
from connect4_b200.match import Match
from connect4_b200.mcts import MCTS, MCTSConfig
from connect4_b200.neural.data import Connect4Dataset
from connect4_b200.neural.model import ModelWrapper

z = np.load("tests/golden/example_net_state.npz")
model = ModelWrapper(state_dict={k: z[k] for k in z.files})


def timed(fn, reps=3):
    fn()                                            # warm (engine creation, first launches)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best


def centre(name, sims, init=1.25):
    return MCTS(name, MCTSConfig(sims, 19652, init), evl.Evaluator(evl.evaluate_centre_with_prior))


import contextlib, io
def quiet(fn):
    def g():
        with contextlib.redirect_stdout(io.StringIO()):
            return fn()
    return g

t = timed(quiet(lambda: Match(False, centre("one", 200), centre("two", 100, 2.5), plies=1, switch=True).play()))
print(json.dumps({"row": "Match 14 games, centre evaluator, 200 vs 100 simulations (golden m0)", "seconds": t,
                  "reference_cpu_seconds": 15.81}))
az = MCTS("AlphaZero", MCTSConfig(800, 19652, 1.25, 0.0, 0.0, 0), evl.Evaluator(partial(evl.evaluate_nn, model=model)))
t = timed(quiet(lambda: Match(False, az, centre("centre", 800), plies=2, switch=True).play()), reps=2)
print(json.dumps({"row": "Match 98 games (2-ply openings, switched), network MCTS vs centre MCTS, 800 simulations", "seconds": t}))

rng = np.random.default_rng(8)
n = 67557                                            # BASELINE configs[4]: 8-ply-shaped set
c0 = np.zeros(n, np.uint64); c1 = np.zeros(n, np.uint64)
bb = BoardBatch(c0, c1)
for ply in range(8):                                 # 8 random legal plies on the device (terminal ones are kept: timing only)
    mask = bb.legal_mask().cpu().numpy()
    mv = np.array([rng.choice(np.flatnonzero([(m >> c) & 1 for c in range(7)])) if m else 0 for m in mask], np.int8)
    bb.drop(mv)
planes = bb.to_planes("float32").cpu()
ds = Connect4Dataset(planes, torch.as_tensor(rng.choice(np.array([0, .5, 1], np.float32), n)), torch.full((n, 7), 1 / 7.))
t = timed(lambda: model.evaluate(ds, shuffle=False))
print(json.dumps({"row": "ModelWrapper.evaluate over 67,557 positions (host planes in, CombinedStats out)", "seconds": t,
                  "positions_per_sec": n / t, "reference_cpu_positions_per_sec_8_threads": 21049}))
v, p = model.evaluate_bitboards(bb.c0, bb.c1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    model.evaluate_bitboards(bb.c0, bb.c1)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(json.dumps({"row": "network kernel alone, 67,557 resident positions", "ms": ms, "positions_per_sec": n / ms * 1e3,
                  "tflops": n * model.flops_per_position / ms / 1e9}))
rec = torch.zeros((500000, 64), dtype=torch.uint8, device="cuda")
augment_pack(rec)                                    # warm: output tensors come from the caching allocator afterwards
torch.cuda.synchronize()
ev0.record()
for _ in range(10):
    augment_pack(rec)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
out_bytes = 2 * 500000 * (126 + 1 + 7) * 4
print(json.dumps({"row": "c4_records_augment_pack, 500k records -> 1M rows of data.pth tensors", "ms": ms,
                  "GB_per_s_written": out_bytes / ms / 1e6}))

# ---- batched bitboard kernels (c4_board.cu): HBM-bound elementwise passes over resident positions
N = 1 << 24
big = BoardBatch(torch.randint(0, 1 << 40, (N,), dtype=torch.int64, device="cuda"),
                 torch.zeros(N, dtype=torch.int64, device="cuda"))
def dev_time(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps
for name, fn, nbytes in (("c4_board_legal_mask", lambda: big.legal_mask(), 17), ("c4_board_result", lambda: big.result(), 17),
                         ("c4_board_fliplr", lambda: big.fliplr(), 32), ("c4_board_evaluate_centre", lambda: big.evaluate_centre(), 24),
                         ("c4_board_to_planes(uint8)", lambda: big.to_planes("uint8"), 16 + 126)):
    ms = dev_time(fn)
    print(json.dumps({"row": "%s over %d resident positions" % (name, N), "ms": ms, "GB_per_s": N * nbytes / ms / 1e6,
                      "bytes_per_position": nbytes}))

# ---- BASELINE configs[1]: the 10,000-position x 800-simulation deterministic sweep, one fused launch
from connect4_b200.engine import Engine
m = np.load("tests/golden/mcts_sweep_800.npz")
eng = Engine(len(m["c0"]), MCTSConfig(800))
def sweep():
    eng.begin(m["c0"], m["c1"]); eng.run("centre")
t = timed(sweep, reps=2)
print(json.dumps({"row": "10,000 searches x 800 simulations, centre evaluator, one launch (BASELINE configs[1])", "seconds": t,
                  "simulations_per_sec": 1e4 * 800 / t, "reference_cpu": "0.35-0.6 s per search and core (SURVEY 8d)"}))

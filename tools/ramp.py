"""cold-start ramp of a self-play pool (evaluation memo starts empty): cumulative positions/s at fixed times and the
rate / hit rate / launch durations of the last block of passes.  usage: ramp.py [seconds] [-v]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
T = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
verbose = "-v" in sys.argv
z = np.load("tests/golden/example_net_state.npz")
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=4096, seed=1)
tot_ms = 0.0; tot_pos = 0; marks = [0.25, 0.5, 1.0, 2.0, 4.0, 8.0]; out = []; n = 0
while tot_ms < T * 1e3:
    r = pool.throughput(500); n += 500
    tot_ms += r["device_ms"]; tot_pos += r["positions"]
    line = "passes %5d t %.3f s pos/s %8.0f cum %8.0f hit %.3f evals/pass %5.0f tree %.1f us net %.1f us" % (
        n, tot_ms / 1e3, r["positions"] / r["device_ms"] * 1e3, tot_pos / tot_ms * 1e3,
        r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), r["evals"] / 500, r["tree_ms"] * 1e3, r["net_ms"] * 1e3)
    if verbose: print(line)
    while marks and tot_ms >= marks[0] * 1e3:
        out.append("cum@%.2gs %.0f" % (marks.pop(0), tot_pos / tot_ms * 1e3))
print(" ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("C4_")), "|", "  ".join(out), "| last:", line)

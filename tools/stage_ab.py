"""A/B of the shared-memory staging of the top of each tree (split engine, c4_split.cu / GameS in c4_tree.cuh):
cold-memo config-3 generation with C4_SP_STAGE = 0 (off), a few caps, and the default (all the shared memory a tree CTA has).
usage: stage_ab.py [--games N] [--stages 0,16,64,-1] [--reps R]      (-1 = default)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("C4_FZ_TIMEOUT_S", "30")
os.environ["C4_ENGINE"] = "split"
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
slots = int(arg("--games", "4096"))
stages = [int(x) for x in arg("--stages", "0,16,32,64,-1").split(",")]
reps = int(arg("--reps", "3"))
pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=1)
for rep in range(reps):
    for s in stages:
        if s < 0:
            os.environ.pop("C4_SP_STAGE", None)
        else:
            os.environ["C4_SP_STAGE"] = str(s)
        r = pool.stream(stop_games=slots, reset=True, cold_memo=True)
        print("stage %4s nodes/game: cold generation until %d games: %.3f s  %.0f positions/s  evals %d hit %.3f engine %s" % (
            "max" if s < 0 else s, slots, r["device_ms"] / 1e3, r["positions"] / r["device_ms"] * 1e3, r["evals"],
            r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), r["engine"]), flush=True)
pool.engine.close()

"""how many network evaluations of a cold generation are duplicates (the same position evaluated more than once because
several games asked for it before the first answer reached the memo): evals vs occupied memo entries"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
for engine in ("fused", "lockstep"):
    os.environ["C4_ENGINE"] = engine
    pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=4096, seed=1)
    tot = 0
    for stop in (256, 1024, 4096):
        r = pool.stream(stop_games=stop - (0 if stop == 256 else prev), reset=(stop == 256), cold_memo=(stop == 256))
        prev = stop
        tot += r["evals"]
        occ = pool.engine.lib.c4_ctx_get(pool.engine.h, 7)
        print("%s: after %5d games: evals %9d  occupied memo entries %9d  -> %.2f evaluations per distinct position" % (
            engine, stop, tot, occ, tot / max(1, occ)), flush=True)
    pool.engine.close()

"""one cold stream of the fused engine (profiling / knob sweeps): fused_prof.py [games_to_finish] [slots]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
stop = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=1)
r = pool.stream(stop_games=stop, reset=True, cold_memo=True)
print(" ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("C4_")), "|",
      "%s cold until %d games on %d slots: %.3f s %.0f positions/s hit %.3f | positions %d evals %d hits %d games %d" % (
          r["engine"], stop, slots, r["device_ms"] / 1e3, r["positions"] / r["device_ms"] * 1e3,
          r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), r["positions"], r["evals"], r["memo_hits"], r["games"]), flush=True)
if "--warm" in sys.argv:
    for _ in range(3):
        r = pool.stream(max_ms=400.0)
        print("   warm: %.0f positions/s hit %.3f" % (r["positions"] / r["device_ms"] * 1e3, r["memo_hits"] / max(1, r["memo_hits"] + r["evals"])), flush=True)

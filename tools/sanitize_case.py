"""the workload compute-sanitizer is run on (memcheck / racecheck, one tool per gpurun call): a small NET self-play
generation on all three engines, a batch of stand-alone NET searches and a 1,536-position c4_net_forward.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py [games] [sims]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from connect4_b200.engine import Engine
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
z = np.load(os.path.join(ROOT, "tests/golden/example_net_state.npz"))
g = np.load(os.path.join(ROOT, "tests/golden/net_outputs.npz"))
games = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 64
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
v, p = model.evaluate_bitboards(g["c0"][:1536], g["c1"][:1536])
print("c4_net_forward 1536 positions: max |dv| %.2e" % np.abs(v.cpu().numpy() - g["value"][:1536]).max(), flush=True)
out = {}
for engine in ("split", "fused", "lockstep"):
    os.environ["C4_ENGINE"] = engine
    pool = SelfPlayPool(model, MCTSConfig(sims, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=games, seed=4)
    rec = pool.generate_records(games + games // 2)
    out[engine] = rec[np.lexsort((rec["ply"], rec["game_id"]))]
    print(engine, "self-play:", len(rec), "records", flush=True)
    pool.engine.close()
    eng = Engine(32, MCTSConfig(sims))
    eng.set_net(model)
    eng.begin(g["c0"][:24], g["c1"][:24])
    eng.run("net")
    print(engine, "searches: root visits", eng.readout()["root_visits"][:4], flush=True)
    eng.close()
assert all(out[e][f].tobytes() == out["lockstep"][f].tobytes() for e in ("split", "fused") for f in out[e].dtype.names)
print("engines agree", flush=True)

import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("C4_FZ_TIMEOUT_S", "60")
from connect4_b200.engine import Engine
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.model import ModelWrapper
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
z = np.load(os.path.join(ROOT, "tests/golden/example_net_state.npz"))
g = np.load(os.path.join(ROOT, "tests/golden/net_outputs.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
outs = {}
for name, env in (("lockstep", {"C4_ENGINE": "lockstep"}), ("one", {"C4_ENGINE": "split"}), ("two", {"C4_ENGINE": "split", "C4_SP_LAUNCH": "two"})):
    os.environ.pop("C4_SP_LAUNCH", None)
    os.environ.update(env)
    for n, sims in ((1, 8), (24, 64)):
        eng = Engine(32, MCTSConfig(sims))
        eng.set_net(model)
        eng.begin(g["c0"][:n], g["c1"][:n])
        eng.run("net")
        r = eng.readout()
        outs[(name, n)] = r
        eng.close()
        print(name, n, sims, "visits[0]", r["visits"][0].tolist(), "root_visits", r["root_visits"][:4].tolist(), "vsum[0]", r["vsum"][0].tolist()[:3], flush=True)
for n in (1, 24):
    for name in ("one", "two"):
        a, b = outs[(name, n)], outs[("lockstep", n)]
        print(name, n, {k: bool(a[k].tobytes() == b[k].tobytes()) for k in a})

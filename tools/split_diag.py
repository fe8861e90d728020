"""which split-engine variant disagrees with the lock-step engine (used with the -DC4_CHECKED build)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("C4_FZ_TIMEOUT_S", "60")
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
KNOBS = ("C4_SP_NET_CTAS", "C4_MEMO_LOG2", "C4_MEMO_NO_DEDUP", "C4_SP_LAUNCH")


def gen(engine, env, slots=200, sims=64, n=320, seed=11):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["C4_ENGINE"] = engine
    pool = SelfPlayPool(model, MCTSConfig(sims, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=seed)
    rec = pool.generate_records(n)
    pool.engine.close()
    rec = rec[np.lexsort((rec["ply"], rec["game_id"]))]
    return rec


def same(a, b):
    return len(a) == len(b) and all(a[f].tobytes() == b[f].tobytes() for f in a.dtype.names)


ref = gen("lockstep", {})
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for env in ({}, {"C4_SP_NET_CTAS": "3"}, {"C4_SP_NET_CTAS": "120"}, {"C4_MEMO_LOG2": "0"}, {"C4_MEMO_NO_DEDUP": "1"},
                {"C4_MEMO_LOG2": "8"}, {"C4_SP_LAUNCH": "two"}, {"C4_SP_LAUNCH": "two", "C4_MEMO_NO_DEDUP": "1"}):
        r = gen("split", env)
        ok = same(r, ref)
        msg = ""
        if not ok:
            k = min(len(r), len(ref))
            bad = [i for i in range(k) if any(r[f][i].tobytes() != ref[f][i].tobytes() for f in r.dtype.names)]
            msg = " first bad record %d (game %d ply %d) of %d bad; len %d vs %d" % (
                bad[0] if bad else -1, int(r["game_id"][bad[0]]) if bad else -1, int(r["ply"][bad[0]]) if bad else -1, len(bad), len(r), len(ref))
        print(rep, env, "OK" if ok else "DIFFERENT" + msg, flush=True)

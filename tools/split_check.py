"""split persistent engine (csrc/c4_split.cu) vs lock-step pass engine: identical records on the same generations, then the
cold-memo config-3 throughput for several tower-CTA counts.  usage: split_check.py [--quick] [--games N] [--ctas a,b,c]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("C4_FZ_TIMEOUT_S", "30")
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper

z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})


def gen(engine, slots, sims, n_games, seed=3):
    os.environ["C4_ENGINE"] = engine
    pool = SelfPlayPool(model, MCTSConfig(sims, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=seed)
    t0 = time.perf_counter()
    rec = pool.generate_records(n_games)
    dt = time.perf_counter() - t0
    pool.engine.close()
    order = np.lexsort((rec["ply"], rec["game_id"]))
    raw = rec.view(np.uint8).reshape(-1, 64)[order]
    return raw.copy().view(rec.dtype).reshape(-1), dt


if "--perf-only" not in sys.argv:
    for slots, sims, n in ((8, 16, 8), (64, 64, 200), (300, 200, 700), (1000, 100, 1500), (1024, 64, 3000)):
        a, ta = gen("split", slots, sims, n)
        b, tb = gen("lockstep", slots, sims, n)
        same = len(a) == len(b) and a.tobytes() == b.tobytes()
        print("slots %4d sims %4d games %4d: split %6d records %.3f s | lockstep %6d records %.3f s | identical %s" % (
            slots, sims, n, len(a), ta, len(b), tb, same), flush=True)
        if not same:
            sys.exit(1)
if "--quick" in sys.argv:
    sys.exit(0)
slots = int(sys.argv[sys.argv.index("--games") + 1]) if "--games" in sys.argv else 4096
ctas = [int(x) for x in sys.argv[sys.argv.index("--ctas") + 1].split(",")] if "--ctas" in sys.argv else [48, 56, 64, 72]
os.environ["C4_ENGINE"] = "split"
pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=1)
for n in ctas:
    os.environ["C4_SP_NET_CTAS"] = str(n)
    r = pool.stream(stop_games=slots, reset=True, cold_memo=True)
    print("split %3d tower CTAs: cold generation until %d games: %.3f s  %.0f positions/s  evals %d hit %.3f engine %s" % (
        n, slots, r["device_ms"] / 1e3, r["positions"] / r["device_ms"] * 1e3, r["evals"],
        r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), r["engine"]), flush=True)
if "--warm" in sys.argv:
    for rep in range(3):
        r = pool.stream(max_ms=500.0)
        print("   warm 0.5 s: %.0f positions/s  hit %.3f" % (r["positions"] / r["device_ms"] * 1e3,
              r["memo_hits"] / max(1, r["memo_hits"] + r["evals"])), flush=True)
pool.engine.close()

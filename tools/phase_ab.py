"""split engine, cold benchmark generation (4,096 slots until 4,096 games are done), three measurements in one process:
 1. PUCT tables in the tree CTAs' shared memory off / on (C4_SP_SMEM_TABLES), alternating;
 2. the time profile of one generation in slices of --slice-ms (positions/s, network evaluations/s, memo hit rate per slice);
 3. a PHASED tower count: the first T1 ms with n1 tower CTAs, the rest with n2 (the engine reads C4_SP_NET_CTAS per launch and a
    launch continues the pool the previous one left), against the constant tower counts.
usage: phase_ab.py [--games N] [--slice-ms 25] [--reps 2]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("C4_FZ_TIMEOUT_S", "30")
os.environ["C4_ENGINE"] = "split"
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


slots = arg("--games", 4096)
slice_ms = arg("--slice-ms", 25.0)
reps = arg("--reps", 2)
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=1)


def generation(phases):
    """phases = [(tower CTAs, ms or None), ...]; the last phase runs until `slots` games are done.  -> positions/s, hit rate"""
    pos = ms = ev = hit = games = 0
    for i, (n, t) in enumerate(phases):
        os.environ["C4_SP_NET_CTAS"] = str(n)
        last = i == len(phases) - 1
        r = pool.stream(stop_games=slots - games if last else 0, max_ms=0.0 if last else float(t), reset=(i == 0), cold_memo=(i == 0))
        pos += r["positions"]; ms += r["device_ms"]; ev += r["evals"]; hit += r["memo_hits"]; games += r["games"]
        if games >= slots:
            break
    return pos / ms * 1e3, hit / max(1, hit + ev), ms


pool.stream(stop_games=slots, reset=True, cold_memo=True)          # warm-up (module load, allocator)
print("== 1. PUCT tables in shared memory (72 towers)", flush=True)
best = {}
for rep in range(reps + 1):
    for tab in ("0", "1"):
        os.environ["C4_SP_SMEM_TABLES"] = tab
        v, h, ms = generation([(72, None)])
        best[tab] = max(best.get(tab, 0.0), v)
        print("tables %s: %.0f positions/s  hit %.3f  %.1f ms" % (tab, v, h, ms), flush=True)
tab = "1" if best["1"] > best["0"] else "0"
os.environ["C4_SP_SMEM_TABLES"] = tab
print("-> tables %s for the rest" % tab, flush=True)

if "--only-tables" in sys.argv:
    sys.exit(0)
print("== 2. time profile of one generation, 72 towers, slices of %.0f ms" % slice_ms, flush=True)
os.environ["C4_SP_NET_CTAS"] = "72"
games = 0
t = 0.0
first = True
while games < slots:
    r = pool.stream(max_ms=slice_ms, reset=first, cold_memo=first)
    first = False
    games += r["games"]; t += r["device_ms"]
    print("t %6.1f ms: %7.0f positions/s  %8.0f evals/s  hit %.3f  games done %d" % (
        t, r["positions"] / r["device_ms"] * 1e3, r["evals"] / r["device_ms"] * 1e3,
        r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), games), flush=True)

if "--signals" in sys.argv:
    # per slice and tower count: throughput next to the engine's own cycle accounting (run with C4_FZ_DEBUG=1: the engine
    # prints boards per strip, tower busy share and tree-warp idle share of every launch to stderr)
    for n in (56, 64, 72, 80, 88):
        os.environ["C4_SP_NET_CTAS"] = str(n)
        games, t, first = 0, 0.0, True
        while games < slots:
            r = pool.stream(max_ms=slice_ms, reset=first, cold_memo=first)
            first = False
            games += r["games"]; t += r["device_ms"]
            print("towers %d t %6.1f ms: %7.0f positions/s  %8.0f evals/s  hit %.3f  games done %d" % (
                n, t, r["positions"] / r["device_ms"] * 1e3, r["evals"] / r["device_ms"] * 1e3,
                r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), games), flush=True)
    sys.exit(0)
if "--tune" in sys.argv:
    # controller variants (C4_SP_ADAPT_TUNE = fast_ms,gain_up,gain_down,max_step): cold generation, 16,384-game generation
    import time
    for tune in ("25,1,1,16", "10,1,1,16", "25,1,1.5,16", "10,1,1.5,16", "10,1,1.5,24", "10,1.25,1.5,24", "12,1,2,24", "25,1,1,16"):
        os.environ.pop("C4_SP_NET_CTAS", None)
        os.environ["C4_SP_ADAPT_TUNE"] = tune
        res = []
        for _ in range(3):
            r = pool.stream(stop_games=slots, reset=True, cold_memo=True)
            res.append(r["positions"] / r["device_ms"] * 1e3)
        pool.engine.clear_memo()
        t0 = time.perf_counter()
        rec = pool.generate_records(4 * slots)
        dt = time.perf_counter() - t0
        print("tune %-16s cold generation %s (launches %d) | generate_records(%d) %.0f positions/s" % (
            tune, ", ".join("%.0f" % v for v in res), r["launches"], 4 * slots, len(rec) / dt), flush=True)
    sys.exit(0)
if "--adapt" in sys.argv:
    # adaptive tower count (the engine's default for 32-filter networks from 1,024 games) against the constant 72
    import time
    def run(label, env):
        for k in ("C4_SP_NET_CTAS", "C4_SP_ADAPT", "C4_SP_ADAPT_LOG"):
            os.environ.pop(k, None)
        os.environ.update(env)
        res = []
        for _ in range(max(1, reps)):
            r = pool.stream(stop_games=slots, reset=True, cold_memo=True)
            res.append(r["positions"] / r["device_ms"] * 1e3)
        print("%-34s cold generation %s positions/s  (launches %d)" % (label, ", ".join("%.0f" % v for v in res), r["launches"]), flush=True)
    run("constant 72 (C4_SP_ADAPT=0)", {"C4_SP_ADAPT": "0"})
    for ms in ("10", "15", "25", "40", "60"):
        run("adaptive, slices of %s ms" % ms, {"C4_SP_ADAPT": ms})
    run("adaptive, default", {})
    run("constant 72 (C4_SP_ADAPT=0)", {"C4_SP_ADAPT": "0"})
    os.environ.pop("C4_SP_ADAPT", None)
    os.environ["C4_SP_ADAPT_LOG"] = "1"
    pool.stream(stop_games=slots, reset=True, cold_memo=True)
    os.environ.pop("C4_SP_ADAPT_LOG", None)
    # a whole generation of 4 pool-fulls from host start positions to host records (the bench's e2e leg)
    for label, env in (("constant 72", {"C4_SP_ADAPT": "0"}), ("adaptive", {}), ("constant 72", {"C4_SP_ADAPT": "0"}), ("adaptive", {})):
        os.environ.pop("C4_SP_ADAPT", None)
        os.environ.update(env)
        pool.engine.clear_memo()
        t0 = time.perf_counter()
        rec = pool.generate_records(4 * slots)
        dt = time.perf_counter() - t0
        print("%-12s generate_records(%d): %d records in %.3f s = %.0f positions/s" % (label, 4 * slots, len(rec), dt, len(rec) / dt), flush=True)
    # warm steady state
    for label, env in (("constant 72", {"C4_SP_ADAPT": "0"}), ("adaptive", {})):
        os.environ.pop("C4_SP_ADAPT", None)
        os.environ.update(env)
        pool.stream(stop_games=slots, reset=True, cold_memo=True)
        for _ in range(3):
            r = pool.stream(max_ms=500.0)
            print("%-12s warm 0.5 s: %.0f positions/s  hit %.3f" % (label, r["positions"] / r["device_ms"] * 1e3,
                  r["memo_hits"] / max(1, r["memo_hits"] + r["evals"])), flush=True)
    sys.exit(0)
print("== 3. constant and phased tower counts", flush=True)
plans = [[(72, None)]]
if "--early" in sys.argv:                                            # more towers at the cold start: measured worse than constant 72
    for n1 in (88, 104):
        for t1 in (25.0, 50.0, 100.0):
            for n2 in (56, 64, 72):
                plans.append([(n1, t1), (n2, None)])
# the time profile says: towers saturated (42 M evaluations/s) from ~60 to ~270 ms, trees the limit after that
for t1 in (250.0, 275.0, 300.0):
    for n2 in (48, 56, 64):
        plans.append([(72, t1), (n2, None)])
for n_mid in (80, 88):
    for n2 in (56, 64):
        plans.append([(72, 50.0), (n_mid, 225.0), (n2, None)])
plans += [[(64, 50.0), (80, 225.0), (56, None)], [(56, 50.0), (72, 225.0), (56, None)], [(72, None)]]
for plan in plans:
    res = [generation(plan) for _ in range(reps)]
    v = max(x[0] for x in res)
    print("%-40s best %.0f positions/s (all: %s)  hit %.3f" % (
        " -> ".join("%d%s" % (n, "" if t is None else " for %.0f ms" % t) for n, t in plan), v,
        ", ".join("%.0f" % x[0] for x in res), res[0][1]), flush=True)
pool.engine.close()

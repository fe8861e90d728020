"""determinism probes: (1) network outputs of the same boards in different batch positions / sizes, bit for bit;
(2) the same generation several times, field-by-field diff of the records"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
z = np.load("tests/golden/example_net_state.npz")
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
g = np.load("tests/golden/net_outputs.npz")
c0, c1 = g["c0"], g["c1"]
n = len(c0)
v0, p0 = model.evaluate_bitboards(c0, c1)
ref = torch.cat([p0, v0[:, None]], 1).cpu().numpy()
rng = np.random.default_rng(0)
bad = 0
for trial in range(40):
    m = int(rng.integers(1, n + 1))
    idx = rng.permutation(n)[:m]
    v, p = model.evaluate_bitboards(c0[idx], c1[idx])
    out = torch.cat([p, v[:, None]], 1).cpu().numpy()
    d = (out.view(np.uint32) != ref[idx].view(np.uint32)).any(1)
    if d.any():
        bad += 1
        j = np.flatnonzero(d)[:3]
        print("net: batch of %d: %d rows differ, e.g. pos-in-batch %s maxabs %.3g" % (m, d.sum(), j.tolist(), np.abs(out - ref[idx]).max()))
print("net determinism: %d of 40 shuffled batches differ" % bad)

def gen(slots):
    pool = SelfPlayPool(model, MCTSConfig(200, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=slots, seed=0)
    rec = pool.generate_records(300)
    pool.engine.close()
    return rec[np.lexsort((rec["ply"], rec["game_id"]))]
base = gen(300)
for t in range(8):
    r = gen(300)
    if len(r) != len(base):
        print("run %d: different record count" % t); continue
    diff = {f: int((r[f].reshape(len(r), -1).view(np.uint8) != base[f].reshape(len(r), -1).view(np.uint8)).any(1).sum()) for f in r.dtype.names}
    diff = {k: v for k, v in diff.items() if v}
    if diff:
        f = "policy" if "policy" in diff else list(diff)[0]
        rows = np.flatnonzero((r[f].reshape(len(r), -1).view(np.uint8) != base[f].reshape(len(r), -1).view(np.uint8)).any(1))
        gids = sorted(set(r["game_id"][rows].tolist()))
        k = rows[0]
        print("run %d: differing fields %s; games %s; first: game %d ply %d  %s vs %s" % (t, diff, gids[:8], r["game_id"][k], r["ply"][k], r[f][k], base[f][k]))
    else:
        print("run %d: identical" % t)

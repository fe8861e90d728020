"""steady-state self-play throughput with the reference's example_config network (64 filters / 6 residual / 6 fc,
random init) on 4096 slots, plus the raw network kernel rate"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from connect4_b200 import _lib
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.config import ModelConfig, NetConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper
torch.manual_seed(0)
m = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
g = np.load("tests/golden/net_outputs.npz")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
L = _lib.load()
for n in (1900, 4096, 16384):
    c0 = torch.as_tensor(np.resize(g["c0"], n).view(np.int64)).cuda(); c1 = torch.as_tensor(np.resize(g["c1"], n).view(np.int64)).cuda()
    out = torch.empty((n, 8), dtype=torch.float32, device="cuda"); st = _lib.stream_ptr()
    call = lambda: L.c4_net_forward(m.c4_net, _lib.ptr(c0), _lib.ptr(c1), n, None, _lib.ptr(out), st)
    for _ in range(3): call()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print("64f net: n %6d  %.1f us  %.0f TFLOP/s" % (n, us, n * m.flops_per_position / us / 1e6))
pool = SelfPlayPool(m, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=4096, seed=1)
tot_ms = tot_pos = 0
while tot_ms < 6000:
    r = pool.throughput(500); tot_ms += r["device_ms"]; tot_pos += r["positions"]
    line = "t %.2f s pos/s %8.0f cum %8.0f hit %.3f evals/pass %5.0f tree %.1f us net %.1f us" % (
        tot_ms / 1e3, r["positions"] / r["device_ms"] * 1e3, tot_pos / tot_ms * 1e3,
        r["memo_hits"] / max(1, r["memo_hits"] + r["evals"]), r["evals"] / 500, r["tree_ms"] * 1e3, r["net_ms"] * 1e3)
print("64f self-play (half pools: %d):" % r["pools"], line)

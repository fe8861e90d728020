"""BASELINE configs[3] on one GPU: the example_config generation (N games, 64f/6r/6fc network, cold memo) per engine and
tower-CTA count.  usage: gen1200.py [games] [slots] [--ctas a,b,c]"""
import os, sys, time, hashlib
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("C4_FZ_TIMEOUT_S", "60")
import torch
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.config import ModelConfig, NetConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper

games = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1200
slots = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else games
ctas = [int(x) for x in sys.argv[sys.argv.index("--ctas") + 1].split(",")] if "--ctas" in sys.argv else [96, 112, 128]
torch.manual_seed(0)
model = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
cfg = MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6)


def run(engine, n_net=None):
    os.environ["C4_ENGINE"] = engine
    if n_net is None:
        os.environ.pop("C4_SP_NET_CTAS", None)
    else:
        os.environ["C4_SP_NET_CTAS"] = str(n_net)
    pool = SelfPlayPool(model, cfg, concurrent_games=slots, seed=0)
    best = None
    for rep in range(2):
        pool.engine.clear_memo()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rec = pool.generate_records(games)
        torch.cuda.synchronize()
        best = time.perf_counter() - t0
    pool.engine.close()
    order = np.lexsort((rec["ply"], rec["game_id"]))
    raw = rec.view(np.uint8).reshape(-1, 64)[order]
    dig = hashlib.sha256(b"".join(rec[f][order].tobytes() for f in rec.dtype.names)).hexdigest()[:16]
    print("%-8s towers %-4s: %d games on %d slots: %.3f s  %d records  digest %s" % (engine, n_net, games, slots, best, len(rec), dig), flush=True)


run("fused")
for n in ctas:
    run("split", n)

"""phases of dist.generate_sharded at N ranks (torchrun): generation / all-gather / sort / host copy, seconds (max over ranks)
usage: torchrun ... tools/dist_profile.py [--games-per-gpu 16384]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from connect4_b200 import dist as c4d
from connect4_b200.mcts import MCTSConfig
from connect4_b200.neural.game_pool import SelfPlayPool
from connect4_b200.neural.model import ModelWrapper

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = int(sys.argv[sys.argv.index("--games-per-gpu") + 1]) if "--games-per-gpu" in sys.argv else 16384
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests/golden/example_net_state.npz"))
model = ModelWrapper(state_dict={k: z[k] for k in z.files})
pool = SelfPlayPool(model, MCTSConfig(800, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=4096, seed=7)
n = per * world


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter()


for rep in range(2):
    pool.engine.clear_memo()
    t0 = sync()
    n_local, base, stride = c4d.shard_games(n, rank, world)
    pool.generate_records(n_local, game_id_base=base, game_id_stride=stride, to_host=False)
    rec = pool.engine.last_records_device
    t1 = sync()
    if world > 1:
        rec = c4d.all_gather_records(rec)
    t2 = sync()
    rec = c4d.sort_records_device(rec)
    t3 = sync()
    host = rec.cpu().numpy()
    t4 = sync()
    pool.engine.clear_memo()
    t5 = sync()
    out = c4d.generate_sharded(pool, n)
    t6 = sync()
    pool.engine.clear_memo(); sync(); t6 = time.perf_counter()
    out0 = c4d.generate_sharded(pool, n, dst=0)
    t7 = sync()
    if rank == 0:
        print("rep %d world %d records %d (%.0f MB): generate %.3f  all-gather %.3f  sort %.3f  host copy %.3f | generate_sharded %.3f s"
              "  (dst=0: %.3f s)" % (rep, world, len(host), len(host) * 64 / 1e6, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t6 - t5, t7 - t6), flush=True)
pool.engine.close()
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python3
"""One self-play generation sharded over the GPUs of a box (BASELINE.json configs[3]): games g -> rank g % world, no
collective on the data path, one NCCL all-gather of the 64-byte position records at the end; rank 0 optionally writes the
reference's `games.pkl` / `data.pth`.  The Philox streams are keyed by the GLOBAL game id, so the records (and the
checksum printed here) do not depend on the number of GPUs.

    python tools/generate.py --games 1200 --net example_config
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/generate.py --games 1200 --net example_config
"""
import argparse
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=1200)               # example_config.py:15
    ap.add_argument("--net", default="example_config", choices=["default", "example_config"])
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--slots", type=int, default=4096, help="concurrent game slots per GPU")
    ap.add_argument("--out", default=None, help="folder for games.pkl / data.pth (rank 0)")
    ap.add_argument("--dump", default=None, help="write the sorted record array to this .npy (rank 0)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from connect4_b200.dist import generate_sharded, shard_games
    from connect4_b200.mcts import MCTSConfig
    from connect4_b200.neural.config import ModelConfig, NetConfig
    from connect4_b200.neural.game_pool import SelfPlayPool
    from connect4_b200.neural.model import ModelWrapper

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)                                             # the same random-init network on every rank
    nc = NetConfig(filters=64, n_fc_layers=6, n_residuals=6) if args.net == "example_config" else NetConfig()
    model = ModelWrapper(ModelConfig(net_config=nc))
    n_local = shard_games(args.games, rank, world)[0]
    pool = SelfPlayPool(model, MCTSConfig(args.sims, 19652, 1.25, 0.3, 0.25, 6),
                        concurrent_games=max(1, min(args.slots, n_local)), seed=0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = generate_sharded(pool, args.games)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        h = hashlib.sha256()                   # field by field: numpy leaves the 4 padding bytes of a record undefined
        for f in rec.dtype.names:
            h.update(np.ascontiguousarray(rec[f]).tobytes())
        digest = h.hexdigest()[:16]
        line = {"games": args.games, "n_gpus": world, "net": args.net, "simulations": args.sims, "positions": int(len(rec)),
                "seconds": dt, "positions_per_sec": len(rec) / dt, "games_on_rank0": n_local,
                "records_sha256_16": digest}
        if args.dump:
            np.save(args.dump, rec)
        if args.out:
            from connect4_b200.neural.data import TrainingDataStorage
            from connect4_b200.neural.training_game import games_from_records
            os.makedirs(args.out, exist_ok=True)
            TrainingDataStorage().save(games_from_records(rec), args.out)
            line["written"] = sorted(os.listdir(args.out))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

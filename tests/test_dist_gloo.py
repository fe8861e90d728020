"""world_size-2 gloo test of the multi-GPU host logic (game sharding + all-gather-v of position records)."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from connect4_b200.dist import all_gather_records, shard_games
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local, base, stride = shard_games(7, rank, world)
    rec = torch.zeros((n_local * 3, 64), dtype=torch.uint8)     # 3 positions per game, byte 0 = global game id
    for i in range(n_local):
        rec[3 * i:3 * i + 3, 0] = base + i * stride
        rec[3 * i:3 * i + 3, 1] = torch.arange(3, dtype=torch.uint8)
    out = all_gather_records(rec)
    q.put((rank, n_local, out.numpy().copy()))
    dist.destroy_process_group()


def test_shard_and_allgather_world2():
    from connect4_b200.dist import shard_games
    assert [shard_games(7, r, 2)[0] for r in range(2)] == [4, 3]
    assert sum(shard_games(1200, r, 8)[0] for r in range(8)) == 1200
    assert shard_games(3, 5, 8)[0] == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [4, 3]
    a, b = res[0][2], res[1][2]
    assert np.array_equal(a, b) and a.shape == (21, 64)
    assert sorted(set(a[:, 0].tolist())) == list(range(7))       # every game exactly once
    assert a[:12, 0].tolist() == [0, 0, 0, 2, 2, 2, 4, 4, 4, 6, 6, 6]  # rank order preserved


def test_sort_records_device_orders_by_game_and_ply():
    """the gathered generation is put in (game_id, ply) order where it lives (dist.sort_records_device): same order as the
    host lexsort, whole 64-byte records moved"""
    import numpy as np
    import torch
    from connect4_b200.dist import sort_records_device
    from connect4_b200.engine import RECORD_DTYPE
    rng = np.random.default_rng(3)
    n = 5000
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    pairs = rng.permutation(np.array([(g, p) for g in range(250) for p in range(20)]))
    rec["game_id"], rec["ply"] = pairs[:, 0], pairs[:, 1]
    rec["c0"] = rng.integers(0, 1 << 49, n)
    raw = torch.as_tensor(rec.view(np.uint8).reshape(n, 64).copy())
    out = sort_records_device(raw).numpy().view(RECORD_DTYPE).reshape(-1)
    ref = rec[np.lexsort((rec["ply"], rec["game_id"]))]
    for f in ("game_id", "ply", "c0"):
        assert (out[f] == ref[f]).all()
    assert sort_records_device(raw[:0]).shape == (0, 64)

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


class _FakeEngine:
    last_records_device = None
    last_generation_device = None


class _FakePool:
    """stands in for SelfPlayPool on a machine without a GPU: `plays` a rank's share by writing records whose game ids follow
    the (base, stride) numbering and whose number of plies depends on the game (ragged shares)"""
    def __init__(self):
        self.engine = _FakeEngine()

    def generate_records(self, n_games, game_id_base=0, game_id_stride=1, start=None, to_host=True):
        from connect4_b200.engine import RECORD_DTYPE
        gids = [game_id_base + i * game_id_stride for i in range(n_games)]
        rows = [(g, p) for g in reversed(gids) for p in range(3 + g % 4)]          # games finish out of order
        rec = np.zeros(len(rows), dtype=RECORD_DTYPE)
        rec["game_id"] = [r[0] for r in rows]
        rec["ply"] = [r[1] for r in rows]
        rec["c0"] = [1000 * r[0] + r[1] for r in rows]
        self.engine.last_records_device = torch.as_tensor(rec.view(np.uint8).reshape(len(rows), 64).copy())
        assert not to_host


def _worker_sharded(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from connect4_b200.dist import generate_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pool = _FakePool()
    everyone = generate_sharded(pool, 11)                                           # every rank gets the generation
    timing = {}
    only0 = generate_sharded(pool, 11, dst=0, timing=timing)                        # only rank 0 copies it to its host
    held = pool.engine.last_generation_device.shape[0]
    q.put((rank, everyone.copy(), None if only0 is None else only0.copy(), sorted(timing), held))
    dist.destroy_process_group()


def test_generate_sharded_world2_dst_and_timing():
    """dist.generate_sharded over two gloo ranks with a stand-in pool: game g is played by rank g % 2, the gathered generation
    is complete, in (game_id, ply) order and identical on both ranks; with dst=0 only rank 0 gets a host copy while every
    rank keeps the whole generation where its records live; the per-phase timing names the four phases"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    want = [(g, p) for g in range(11) for p in range(3 + g % 4)]
    for rank, everyone, only0, phases, held in res:
        assert list(zip(everyone["game_id"].tolist(), everyone["ply"].tolist())) == want
        assert (everyone["c0"] == 1000 * everyone["game_id"].astype(np.int64) + everyone["ply"]).all()
        assert phases == sorted(["generate", "all_gather", "sort"] + (["host_copy"] if rank == 0 else []))
        assert held == len(want)
        assert (only0 is None) == (rank != 0)
    for f in res[0][1].dtype.names:                       # field by field (numpy leaves a record's padding bytes undefined)
        assert res[0][1][f].tobytes() == res[1][1][f].tobytes() == res[0][2][f].tobytes(), f

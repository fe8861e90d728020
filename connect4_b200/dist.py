"""Multi-GPU generation: games are independent, so they are sharded round-robin over ranks (one process per GPU,
torch.distributed) with NO collective on the data path; the only exchange is an all-gather of the generated position
records at the end of a generation.  Global game ids (rank + i * world) key the RNG streams, so a
generation is identical for every world size.
"""
import numpy as np


def shard_games(n_games, rank, world):
    """number of games rank plays and its (base, stride) in the global game numbering: game g -> rank g % world"""
    n_local = (n_games - rank + world - 1) // world if n_games > rank else 0
    return n_local, rank, world


def all_gather_records(records, group=None):
    """records: uint8 tensor [n_local, 64] (CUDA for the NCCL backend, CPU for gloo). Returns every rank's records
    concatenated in rank order: one all-gather of the counts, one all-gather of the payloads padded to the largest share
    into ONE buffer, and one index_select that drops the padding."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return records
    dev = records.device
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.tolist()
    m = max(max(counts), 1)
    padded = torch.zeros((m, 64), dtype=torch.uint8, device=dev)
    padded[:records.shape[0]] = records
    out = torch.empty((world * m, 64), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(c == m for c in counts):
        return out
    keep = torch.cat([torch.arange(r * m, r * m + c, device=dev) for r, c in enumerate(counts)])
    return out.index_select(0, keep)


def sort_records_device(rec):
    """records (uint8 tensor [n, 64], any device) ordered by (game_id, ply): the key is read out of the raw bytes and
    sorted where the records live, so a gathered generation is put in order on the GPU before it goes to the host"""
    import torch
    if rec.shape[0] == 0:
        return rec
    gid = rec[:, 52:56].contiguous().view(torch.int32).reshape(-1).to(torch.int64)
    ply = rec[:, 57].to(torch.int64)
    order = torch.argsort(gid * 64 + ply)
    return rec.index_select(0, order)


def generate_sharded(pool, n_games, group=None, dst=None, timing=None):
    """Play this rank's share of `n_games` on `pool` (SelfPlayPool) and all-gather the records of all ranks (every GPU ends
    with the whole generation in HBM: `pool.engine.last_generation_device`, sorted by (game_id, ply)).
    Returns them as a numpy record array -- identical on every rank; with `dst` = a rank, only that rank copies the
    generation to its host (the reference collects the games in ONE process, neural/training.py:112-133) and the other
    ranks return None.  `timing`: a dict that receives this rank's seconds per phase (adds a device synchronisation
    after each phase)."""
    import time
    import torch
    import torch.distributed as dist
    from .engine import records_to_host
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)

    def mark(name, t0):
        if timing is None:
            return t0
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        timing[name] = t1 - t0
        return t1

    t = time.perf_counter()
    n_local, base, stride = shard_games(n_games, rank, world)
    pool.generate_records(n_local, game_id_base=base, game_id_stride=stride, to_host=False)
    rec = pool.engine.last_records_device
    t = mark("generate", t)
    if world > 1:
        rec = all_gather_records(rec, group)
    t = mark("all_gather", t)
    rec = sort_records_device(rec)
    pool.engine.last_generation_device = rec
    t = mark("sort", t)
    if dst is not None and rank != dst:
        return None
    out = records_to_host(rec)
    mark("host_copy", t)
    return out

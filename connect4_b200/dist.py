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
    concatenated in rank order: one all-gather of the counts, one all-gather of the padded payloads."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return records
    dev = records.device
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    padded = torch.zeros((m, 64), dtype=torch.uint8, device=dev)
    padded[:records.shape[0]] = records
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


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


def generate_sharded(pool, n_games, group=None):
    """Play this rank's share of `n_games` on `pool` (SelfPlayPool) and all-gather the records of all ranks.
    Returns a numpy record array sorted by (game_id, ply) -- identical on every rank."""
    import torch.distributed as dist
    from .engine import RECORD_DTYPE
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    n_local, base, stride = shard_games(n_games, rank, world)
    pool.generate_records(n_local, game_id_base=base, game_id_stride=stride, to_host=False)
    rec = pool.engine.last_records_device
    if world > 1:
        rec = all_gather_records(rec, group)
    return sort_records_device(rec).cpu().numpy().view(RECORD_DTYPE).reshape(-1)

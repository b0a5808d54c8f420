"""Multi-GPU plumbing for self-play: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch on the
GPU box, gloo in CPU tests).  Games are independent, so the simulation path has NO collective: rank r owns the game
ids [r*G, (r+1)*G).  Collectives appear exactly twice per generation:
  broadcast_weights   a weight blob from the trainer rank into every rank's device buffer: either the FOLDED blob the
                      kernels read (Engine.net_folded_tensor(): tower bf16 with BatchNorm folded, 49.6 MB for the reference
                      net; receivers call net_adopt_folded(), no fold there — what bench.py times), or the fp32 state_dict
                      blob (Engine.net_blob_tensor(), 101.6 MB; every rank then folds with net_commit())
  gather_records      variable-length gather of packed records (96 B bitboards + move + reward per position)
The reference's only multi-GPU construct is nn.DataParallel (ai/model_utils.py:26-28); this replaces it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_games_total: int, rank: int, world: int):
    """Contiguous block of game ids for a rank (the last ranks take one game less when it does not divide)."""
    base, rem = divmod(n_games_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights(blob: torch.Tensor, src: int = 0):
    """In-place broadcast of the weight blob (device staging tensor on NCCL ranks, CPU tensor under gloo)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def gather_records(lines: torch.Tensor, move: torch.Tensor, reward: torch.Tensor, game_id_base: int, game: torch.Tensor,
                   dst: int = 0):
    """Gather every rank's records on `dst` in global game order.  Inputs are this rank's packed records
    (lines int64 [n,16], move int32 [n], reward float32 [n], local game index int32 [n]).
    Returns (lines, move, reward, global_game) on dst, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return lines, move, reward, game.to(torch.int64) + game_id_base
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = lines.device
    n = torch.tensor([lines.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)

    def padded(t, width, dtype):
        out = torch.zeros((cap,) + ((width,) if width else ()), dtype=dtype, device=dev)
        out[: t.shape[0]] = t
        return out
    payload = (padded(lines, 16, torch.int64), padded(move, 0, torch.int32), padded(reward, 0, torch.float32),
               padded(game.to(torch.int64) + game_id_base, 0, torch.int64))
    outs = []
    for t in payload:
        bufs = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        outs.append(bufs)
    if rank != dst:
        return None
    cat = [torch.cat([b[:c] for b, c in zip(bufs, counts)]) for bufs in outs]
    return tuple(cat)

"""Frame sharding over ranks and the single collective of the path.

The reference shards frames with a DistributedSampler, writes one pickle per rank and meets at
a barrier (/root/reference/src/optimize.py:78-89, src/utils/init_utils.py:10-18).  Frames are
independent, so here every rank refines a contiguous block of frame ids with no communication
inside the loop, and the refined parameters plus per-frame loss statistics are exchanged with
ONE all-gather (NCCL on GPUs, gloo in the CPU tests).  SURVEY.md §8(e).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist

STATS_DIM = 2     # per-frame [collision_loss, joints_3d_loss_p]


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """torchrun-style initialisation; returns (rank, world_size, local_rank). A plain
    single-process run (no WORLD_SIZE) needs no process group and returns (0, 1, 0)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    rank, local_rank = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device(f"cuda:{local_rank}")
        dist.init_process_group(backend=backend, **kwargs)
    return rank, world, local_rank


def shard_range(total_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, start+count) of rank `rank`; blocks differ by at most one frame."""
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def pack_results(params: torch.Tensor, collision_loss: torch.Tensor, joints_3d_loss_p: torch.Tensor) -> torch.Tensor:
    """(b,122) refined parameters + (b,2) loss statistics -> (b,124)."""
    return torch.cat([params, collision_loss.view(-1, 1), joints_3d_loss_p.view(-1, 1)], dim=1).contiguous()


def all_gather_results(local: torch.Tensor, total_frames: int) -> torch.Tensor:
    """All ranks receive the (total_frames, 124) matrix in frame-id order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_range(total_frames, r, world)[1] for r in range(world)]
    if len(set(counts)) == 1:
        out = torch.empty(total_frames, local.shape[1], dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out
    # uneven blocks: pad every shard to the largest, gather once, drop the padding rows
    cmax = max(counts)
    padded = torch.zeros(cmax, local.shape[1], dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty(world * cmax, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * cmax:r * cmax + c] for r, c in enumerate(counts)], dim=0)

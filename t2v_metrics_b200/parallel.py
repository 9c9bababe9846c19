"""Data-parallel scoring across the GPUs of one box: one process per GPU, a full model replica each, the flattened pair
list split contiguously, and ONE all-gather of the fp32 scores at the end. The reference has no
multi-GPU data path at all (only HF device_map="auto", qwen2vl_model.py:120,128).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_pairs: int, world_size: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous split: ceil(N / world) pairs per rank (the tail ranks may be short or empty).
    Returns (start, end, per_rank)."""
    per = (num_pairs + world_size - 1) // world_size
    start = min(rank * per, num_pairs)
    end = min(start + per, num_pairs)
    return start, end, per


def gather_scores(local_scores: torch.Tensor, num_pairs: int, group=None) -> torch.Tensor:
    """All-gather the per-rank score shards (padded to ceil(N/world)) and trim to N. Works with NCCL (cuda tensors)
    and gloo (cpu tensors)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_scores[:num_pairs]
    per = (num_pairs + world - 1) // world
    padded = torch.zeros(per, dtype=torch.float32, device=local_scores.device)
    padded[: local_scores.numel()] = local_scores.float()
    out = torch.empty(per * world, dtype=torch.float32, device=local_scores.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:num_pairs]

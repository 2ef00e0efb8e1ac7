"""Data-parallel training over the GPUs of one box.

Tasks (dates x variables) are independent (nzdownscale/downscaler/train.py:308-317), so rank r steps
its own shard of every equal-station-count group and the only exchange is ONE all-reduce of the flat
fp32 gradient bucket (4.58 MB) per step -- NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests.
The reference itself is single-process; this is the generalisation of ``outputs/infer.py --gpu``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_tasks(tasks: Sequence, rank: int, world: int) -> List:
    """Round-robin shard that keeps the reference's grouping rule (train.py:448-475): tasks are first
    grouped by number of target stations, then every group is dealt out i = r (mod W)."""
    groups: Dict[int, list] = {}
    for t in tasks:
        groups.setdefault(int(t["X_t"][0].shape[-1]), []).append(t)
    out = []
    for k in sorted(groups):
        g = groups[k]
        n = (len(g) // world) * world  # equal work per rank; the remainder is dropped like train_epoch does
        out.extend(g[rank:n:world])
    return out


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=group)


def allreduce_mean_(flat: torch.Tensor, world: int, group=None) -> torch.Tensor:
    dist.all_reduce(flat, group=group)
    return flat.mul_(1.0 / world)


def enable_data_parallel(model, group=None) -> None:
    """Identical initial weights on every rank + gradient all-reduce inside the engine's backward."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    world = dist.get_world_size(group)
    broadcast_parameters(model.model, 0, group)
    model.engine.allreduce_group = group if group is not None else dist.group.WORLD
    model.engine.world_size = world

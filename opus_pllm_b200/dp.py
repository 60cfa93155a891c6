"""Data-parallel sharding of the prompt list — one process per GPU, each with a full replica.

Mirrors what the reference does through HF accelerate (multi_modality_v1/eval/run_opus_ddp.py:50,77-79,138):
`Accelerator.split_between_processes` = contiguous, order-preserving chunks (the first `len % world` ranks get one extra
item) and ONE collective at the end of the run. The reference gathers pickled strings (`gather_object`); here the
payload is the fixed-shape int64 token matrix, gathered with a single all_gather (NCCL on GPUs, gloo in CPU tests), and
detokenisation happens on the main rank. There is no collective on the model path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    per, extra = divmod(n, world)
    start = rank * per + min(rank, extra)
    return start, start + per + (1 if rank < extra else 0)


def split_between_processes(items, rank: int | None = None, world: int | None = None):
    """Contiguous shard of `items` for this rank (accelerate semantics)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard_bounds(len(items), rank, world)
    return items[a:b]


def gather_token_ids(local_ids: torch.Tensor, pad_id: int, group=None) -> torch.Tensor:
    """local_ids int64 [n_local, n_new] (n_local and n_new may differ between ranks) -> [N_total, max n_new] on every
    rank, rows in rank order == original prompt order for contiguous shards. One size exchange + one all_gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_ids
    world = dist.get_world_size(group)
    dev = local_ids.device
    shape = torch.tensor([local_ids.shape[0], local_ids.shape[1]], dtype=torch.int64, device=dev)
    shapes = torch.empty((world * 2,), dtype=torch.int64, device=dev)     # flat buffers: valid for NCCL and gloo
    dist.all_gather_into_tensor(shapes, shape, group=group)
    shapes_h = shapes.cpu().view(world, 2)
    n_max, t_max = int(shapes_h[:, 0].max()), int(shapes_h[:, 1].max())
    padded = torch.full((n_max, t_max), pad_id, dtype=torch.int64, device=dev)
    padded[: local_ids.shape[0], : local_ids.shape[1]] = local_ids
    out = torch.empty((world * n_max * t_max,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, padded.view(-1), group=group)
    out = out.view(world * n_max, t_max)
    rows = [out[r * n_max: r * n_max + int(shapes_h[r, 0])] for r in range(world)]
    return torch.cat(rows, 0)

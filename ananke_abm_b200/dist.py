"""Multi-GPU plumbing: agents shard across ranks, parameters and the zone graph are replicated.

No agent-agent term exists in either drift, so inference needs NO collective and training needs
exactly one: a sum all-reduce of the flat fp32 gradient buffer per optimiser step (NCCL over NVLink; ~0.84 M floats at
Z=10k -- latency bound).  `clip_grad_norm_` must run AFTER the all-reduce to match the single-process semantics of
/root/reference/src/ananke_abm/models/mode_sep/train/train.py:163.  Losses that are means over masked elements
(mode_sep/architecture/losses.py:14-43) are combined with `global_mean`: numerator and denominator are all-reduced.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of agents owned by `rank`: global index = lo + local index (batch indexing is unchanged)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_agents(tensors: Iterable[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    out = []
    for t in tensors:
        lo, hi = shard_bounds(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return out


def flatten_grads(params: Iterable[torch.nn.Parameter]) -> torch.Tensor:
    ps = list(params)
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])


def unflatten_to_grads(flat: torch.Tensor, params: Iterable[torch.nn.Parameter]) -> None:
    off = 0
    for p in params:
        n = p.numel()
        p.grad = flat[off:off + n].view_as(p).clone()
        off += n


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> torch.Tensor:
    """One sum all-reduce of the flat gradient buffer; writes the reduced gradients back.  Returns the flat buffer."""
    ps = list(params)
    flat = flatten_grads(ps)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unflatten_to_grads(flat, ps)
    return flat


def global_mean(local_sum: torch.Tensor, local_count: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over all ranks of a masked quantity: all-reduce numerator and denominator together."""
    buf = torch.stack([local_sum.reshape(()).float(), local_count.reshape(()).float()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[0] / buf[1].clamp_min(1.0)


def allreduce_error_sumsq(sumsq: torch.Tensor, group=None) -> torch.Tensor:
    """dopri5 with norm="global": torchdiffeq's RMS error norm runs over ALL agents, so a sharded
    run needs the sum of squares reduced across ranks once per step attempt (1 float)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sumsq, op=dist.ReduceOp.SUM, group=group)
    return sumsq

"""Parameter synchronisation around an expert-parallel MoE layer ("Gradient sync").

Under expert parallelism the expert tensors of a rank are ITS experts: they are never broadcast and their gradients are
never reduced (each expert's gradient already sums over every rank's tokens).  Everything else - the replicated router
and the rest of the backbone - starts identical on all ranks and averages its gradients over the world.  The reference
does this with

  * `sync_weights(model, except_key_words=["mlp.experts.h4toh", "mlp.experts.htoh4"])`
    (/root/reference/utils/moe_utils.py:310-324, called at train_fastmoe.py:461): broadcast every state-dict entry whose
    key contains none of the keywords from rank 0, then load the state dict back;
  * `model.allreduce_params()` of `fmoe.DistributedGroupedDataParallel` (train/train_utils.py:288,323,414,461), which
    reduces the gradients of parameters by their `dp_comm` tag and skips the ones tagged "none" (the experts:
    `mark_parallel_comm`, origin/custom_moe_layer.py:159).

Pure torch.distributed host code (any backend: NCCL on the GPUs, gloo in the CPU tests); the layer's data path does not
use it."""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist

EXPERT_KEY_WORDS = ("mlp.experts.h4toh", "mlp.experts.htoh4")


def sync_weights(model: torch.nn.Module, except_key_words: Iterable[str] = EXPERT_KEY_WORDS, group=None, src: int = 0) -> None:
    """Reference `sync_weights`: every state-dict tensor (parameters AND buffers) whose key contains none of
    `except_key_words` is overwritten with rank `src`'s copy; expert tensors keep their rank-local values."""
    state_dict = model.state_dict()
    words = tuple(except_key_words)
    for key, item in state_dict.items():
        if any(w in key for w in words):
            continue
        dist.broadcast(item, src, group=group)
    model.load_state_dict(state_dict)


def allreduce_replicated_grads(model: torch.nn.Module, group=None, average: bool = True,
                               skip_tags: Iterable[str] = ("none",)) -> int:
    """Gradient half of the contract: all-reduce (mean by default) the gradients of every parameter whose `dp_comm` tag is
    not in `skip_tags` - untagged parameters count as replicated ("dp"), the layer tags its router "gate" and its experts
    "none".  Parameters without a gradient on this rank (a task gate that was not used in this step) contribute zeros, so
    that every rank issues the same collectives.  Gradients are reduced in ONE flat buffer per dtype.  Returns the number
    of parameters reduced."""
    skip = set(skip_tags)
    world = dist.get_world_size(group)
    by_dtype = {}
    for p in model.parameters():
        if not p.requires_grad or getattr(p, "dp_comm", "dp") in skip:
            continue
        by_dtype.setdefault(p.dtype, []).append(p)
    n = 0
    for dtype, params in by_dtype.items():
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
        off = 0
        for p, g in zip(params, grads):
            k = g.numel()
            if p.grad is None:
                p.grad = flat[off:off + k].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + k].view_as(p))
            off += k
        n += len(params)
    return n

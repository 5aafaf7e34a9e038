"""Checkpoint / wire format of the MoE layer's parameters under expert parallelism.

State-dict contract: expert tensors `...mlp.experts.htoh4.{weight,bias}` and
`...mlp.experts.h4toh.{weight,bias}` carry the expert index in dim 0; rank r of an expert-parallel
job of W ranks owns the contiguous slice [r*E_loc, (r+1)*E_loc).  The reference

  * writes one `{rank}.pth` per rank into a directory: rank 0 the full state, other ranks the
    expert tensors only (/root/reference/utils/moe_utils.py:164-189),
  * merges them back by concatenating expert tensors along dim 0 (train_fastmoe.py:559-597),
  * slices a global state dict for a rank with `item[rank*E_loc:(rank+1)*E_loc]`
    (utils/moe_utils.py:191-198), and
  * tags single-file checkpoints with meta.expert_format in {"global","local"} and refuses
    rank-local ones (utils/moe_utils.py:34-106, pretrain/utils/moe_checkpoint.py:57-112).

These helpers reproduce that format so real M3ViT checkpoints load into the B200 layer at any W.
Pure host code (torch CPU tensors); no CUDA involved.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

EXPERT_KEYWORDS = ("mlp.experts.htoh4", "mlp.experts.h4toh")


def is_expert_key(key: str) -> bool:
    return any(p in key for p in EXPERT_KEYWORDS)


def strip_prefixes(key: str) -> str:
    for p in ("module.", "encoder."):
        if key.startswith(p):
            key = key[len(p):]
    return key


def first_expert_dim0(state_dict) -> Optional[int]:
    for k, v in state_dict.items():
        if is_expert_key(strip_prefixes(k)) and torch.is_tensor(v):
            return int(v.shape[0])
    return None


def shard_expert_state_dict(global_sd: Dict[str, torch.Tensor], rank: int, num_local: int) -> "OrderedDict":
    """global -> this rank's view: expert tensors sliced on dim 0, everything else unchanged."""
    out = OrderedDict()
    for k, v in global_sd.items():
        out[k] = v[rank * num_local:(rank + 1) * num_local] if is_expert_key(k) else v
    return out


def filter_expert_state(state_dict) -> "OrderedDict":
    """what ranks != 0 write: expert tensors only"""
    return OrderedDict((k, v) for k, v in state_dict.items() if is_expert_key(k))


def merge_expert_shards(shards: List[Dict[str, torch.Tensor]]) -> "OrderedDict":
    """[rank0 full state, rank1 experts, ...] -> global state dict (expert tensors concatenated on dim 0)."""
    if not shards:
        raise ValueError("no shards")
    out = OrderedDict(shards[0])
    for k in list(out.keys()):
        if is_expert_key(k):
            parts = [s[k] for s in shards]
            out[k] = torch.cat(parts, dim=0)
    return out


def expert_format(checkpoint: dict, state_dict, local_experts: int, world_size: int) -> str:
    """Classify a single-file checkpoint as "global" (usable at any W) or raise ValueError for a
    rank-local one, following the reference's rules (utils/moe_utils.py:34-106)."""
    dim0 = first_expert_dim0(state_dict)
    if dim0 is None or int(world_size) <= 1:
        return "global"
    expected_global = int(local_experts) * int(world_size)
    meta = checkpoint.get("meta", {}) if isinstance(checkpoint, dict) else {}
    fmt = meta.get("expert_format") if isinstance(meta, dict) else None
    if fmt == "global":
        if dim0 != expected_global:
            raise ValueError(f"meta says global experts but dim0={dim0}, expected {expected_global}")
        return "global"
    if fmt == "local":
        raise ValueError("checkpoint holds rank-local experts only; merge the shard directory first")
    if dim0 == expected_global:
        return "global"
    raise ValueError(f"cannot verify global expert format: expert dim0={dim0}, expected {expected_global} "
                     f"(local_experts={local_experts}, world_size={world_size})")


def save_ep_shard(state: dict, dirname: str, rank: int) -> str:
    """`{dirname}/{rank}.pth`; ranks != 0 keep only the expert tensors of state["state_dict"]."""
    os.makedirs(dirname, exist_ok=True)
    if rank != 0:
        state = dict(state)
        state["state_dict"] = filter_expert_state(state["state_dict"])
    path = os.path.join(dirname, f"{rank}.pth")
    torch.save(state, path)
    return path


def load_ep_dir(dirname: str, world_size: int, map_location="cpu") -> dict:
    """merge `{0..W-1}.pth` into one global checkpoint tagged meta.expert_format = "global"."""
    states = [torch.load(os.path.join(dirname, f"{r}.pth"), map_location=map_location, weights_only=False)
              for r in range(world_size)]
    out = dict(states[0])
    out["state_dict"] = merge_expert_shards([s["state_dict"] for s in states])
    meta = dict(out.get("meta", {}))
    meta["expert_format"] = "global"
    out["meta"] = meta
    return out

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

These helpers reproduce that format so real M3ViT checkpoints load into the B200 layer at any W, and
`upcycle_dense_mlp` / `inject_experts_from_dense_mlp` build the expert tensors from a dense DeiT / ViT MLP the way
the reference's upcycling does (utils/helpers.py:481-713; golden: oracle/make_upcycle_golden.py), with the router side of it:
`convert_gate_keys` (utils/common_config.py:47-68; oracle/make_gatekeys_golden.py) and the virtual-group router
initialisation (utils/helpers.py:715-866; oracle/make_vgi_golden.py).
Pure host code (torch CPU tensors); no CUDA involved.

Parity: `oracle/make_ckpt_golden.py` executes the reference's own functions (utils/moe_utils.py:34-198 under the
fmoe shim, pretrain/utils/moe_checkpoint.py:57-212 as is, `save_moe_model_to_dir` under a 2-rank gloo group) on a
fixed set of state dicts and records what they return / raise in tests/golden/ckpt_reference.pt;
tests/test_checkpoint.py holds every helper here to those records.
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
    rank-local one, following the reference's rules in their order
    (validate_single_file_moe_checkpoint_or_raise, utils/moe_utils.py:34-106): world size 1 or no expert tensors ->
    fine; meta.expert_format; the checkpoint's own args (world_size x dim0 == moe_experts -> rank-local); dim0 against
    local_experts x world_size."""
    if int(world_size) <= 1:
        return "global"
    dim0 = first_expert_dim0(state_dict)
    if dim0 is None:
        return "global"
    local_experts, world_size = int(local_experts), int(world_size)
    expected_global = local_experts * world_size
    meta = checkpoint.get("meta", {}) if isinstance(checkpoint, dict) else {}
    fmt = meta.get("expert_format") if isinstance(meta, dict) else None
    if fmt == "global":
        if dim0 != expected_global:
            raise ValueError(f"meta says global experts but dim0={dim0}, expected {expected_global}")
        return "global"
    if fmt == "local":
        raise ValueError("checkpoint holds rank-local experts only; merge the shard directory first")
    args = checkpoint.get("args", {}) if isinstance(checkpoint, dict) else {}
    ck_world = args.get("world_size") if isinstance(args, dict) else None
    ck_global = args.get("moe_experts") if isinstance(args, dict) else None
    if ck_world is not None and ck_global is not None:
        if int(ck_world) > 1 and dim0 * int(ck_world) == int(ck_global):
            raise ValueError(f"single-file checkpoint holds rank-local experts only (dim0={dim0}, written by "
                             f"{int(ck_world)} ranks for {int(ck_global)} experts); merge the shard directory first")
    if dim0 == expected_global:
        return "global"
    if dim0 == local_experts:
        raise ValueError(f"single-file checkpoint holds rank-local experts only (dim0={dim0} == local_experts); "
                         "merge the shard directory first")
    raise ValueError(f"cannot verify global expert format: expert dim0={dim0}, expected {expected_global} "
                     f"(local_experts={local_experts}, world_size={world_size})")


def infer_expert_format(checkpoint, state_dict, expected_global_experts=None, expected_world_size=None) -> str:
    """'global' | 'local' | 'dense' | 'unknown' for one state dict (pretrain/utils/moe_checkpoint.py:137-180): the meta tag
    wins, a state dict without expert tensors is dense, otherwise dim 0 is held against the expected (or the
    checkpoint's own args.moe_experts / args.world_size) expert count."""
    if isinstance(checkpoint, dict):
        meta = checkpoint.get("meta", {})
        if isinstance(meta, dict) and meta.get("expert_format") in ("global", "local"):
            return meta["expert_format"]
    dim0 = None
    for k, v in state_dict.items():                      # NOT prefix-stripped here, like the reference
        if is_expert_key(k) and torch.is_tensor(v):
            dim0 = int(v.shape[0])
            break
    if dim0 is None:
        return "dense"
    if expected_global_experts is None and isinstance(checkpoint, dict):
        args = checkpoint.get("args", {})
        if isinstance(args, dict):
            expected_global_experts = args.get("moe_experts")
            if expected_world_size is None:
                expected_world_size = args.get("world_size")
    if expected_global_experts is not None:
        g = int(expected_global_experts)
        if dim0 == g:
            return "global"
        if expected_world_size is not None and int(expected_world_size) > 1 and dim0 * int(expected_world_size) == g:
            return "local"
    return "unknown"


def to_backbone_state_dict(state_dict):
    """wrapper / DDP key space -> backbone key space of the multi-task loader (pretrain/utils/moe_checkpoint.py:23-47):
    `module.` and a leading `encoder.` are stripped, wrapper-only top-level `head.*` / `norm.*` are dropped.
    Returns (state, dropped keys)."""
    out, dropped = OrderedDict(), []
    for k, v in state_dict.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.startswith("encoder."):
            out[k[len("encoder."):]] = v
        elif k.startswith("head.") or k.startswith("norm."):
            dropped.append(k)
        else:
            out[k] = v
    return out, dropped


def build_meta(state_dict, source, world_size: int = 1, moe_experts_global=None, moe_experts_local=None) -> dict:
    """meta block of a global-expert checkpoint (pretrain/utils/moe_checkpoint.py:82-113)"""
    dim0 = None
    for k, v in state_dict.items():
        if is_expert_key(k) and torch.is_tensor(v):
            dim0 = int(v.shape[0])
            break
    if moe_experts_global is None and dim0 is not None:
        moe_experts_global = dim0
    if moe_experts_local is None:
        if dim0 is None:
            moe_experts_local = 0
        elif world_size > 0 and dim0 % world_size == 0:
            moe_experts_local = dim0 // world_size
        else:
            moe_experts_local = dim0
    return {"expert_format": "global", "moe_experts_global": int(moe_experts_global) if moe_experts_global is not None else 0,
            "moe_experts_local": int(moe_experts_local), "world_size": int(world_size), "source": str(source)}


def save_ep_shard(state: dict, dirname: str, rank: int) -> str:
    """`{dirname}/{rank}.pth`; ranks != 0 keep only the expert tensors of state["state_dict"]."""
    os.makedirs(dirname, exist_ok=True)
    if rank != 0:
        state = dict(state)
        state["state_dict"] = filter_expert_state(state["state_dict"])
    path = os.path.join(dirname, f"{rank}.pth")
    torch.save(state, path)
    return path


def _state_of(ckpt):
    if isinstance(ckpt, dict) and "state_dict" in ckpt:
        return ckpt["state_dict"]
    if isinstance(ckpt, dict) and "model" in ckpt:
        return ckpt["model"]
    return ckpt


def merge_shard_dir(dirname: str, map_location="cpu"):
    """`{0,1,..}.pth` of a train_fastmoe-style shard directory -> (rank-0 checkpoint, merged global state dict, number
    of shards), as pretrain/utils/moe_checkpoint.py:196-225: files are taken in numeric rank order, rank 0 must exist,
    expert tensors are concatenated on dim 0, other keys a later shard brings are added."""
    files = sorted((int(os.path.splitext(n)[0]), os.path.join(dirname, n)) for n in os.listdir(dirname)
                   if n.endswith(".pth") and os.path.splitext(n)[0].isdigit())
    if not files:
        raise ValueError(f"No rank shard '*.pth' files found in: {dirname}")
    if files[0][0] != 0:
        raise ValueError("Shard directory must contain rank-0 checkpoint file '0.pth'")
    base = torch.load(files[0][1], map_location=map_location, weights_only=False)
    if not isinstance(base, dict):
        raise ValueError(f"Checkpoint at {files[0][1]} must be a dict, got {type(base)}")
    merged = OrderedDict(_state_of(base))
    for _, path in files[1:]:
        shard = _state_of(torch.load(path, map_location=map_location, weights_only=False))
        for k, v in shard.items():
            if is_expert_key(k):
                merged[k] = torch.cat([merged[k], v], dim=0) if k in merged else v
            elif k not in merged:
                merged[k] = v
    return base, merged, len(files)


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


# ----------------------------------------------------------------------------- dense MLP -> experts ("upcycling")
def upcycle_dense_mlp(fc1_w: torch.Tensor, fc1_b: torch.Tensor, fc2_w: torch.Tensor, fc2_b: torch.Tensor, *,
                      local_experts: int, total_experts: Optional[int] = None, expert_hidden: Optional[int] = None,
                      top_k: int = 4, split: Optional[bool] = None, weight_scaling: bool = False,
                      require_granularity_4: bool = False):
    """Expert parameters of ONE MoE block from the dense DeiT / ViT MLP it replaces
    (/root/reference/utils/helpers.py:481-713, `_inject_moe_expert_from_deit_mlp`, the per-block body).

      fc1_w [Hd, D], fc1_b [Hd], fc2_w [D, Hd], fc2_b [D]   ->
      (htoh4.weight [E_loc, He, D], htoh4.bias [E_loc, He], h4toh.weight [E_loc, D, He], h4toh.bias [E_loc, D])

    * copy mode (`split=False`; the reference's moe_mlp_ratio != 1 path): every local expert is a copy of the dense MLP.
    * split mode (`split=True`; moe_mlp_ratio == 1 or deit_init_mode == "deit_warm_start"): the dense hidden dimension is
      cut into G = Hd / He groups - expert j of a group owns rows [j He, (j+1) He) of fc1 and the matching columns of fc2,
      fc2's bias is repeated per expert.  E_loc a multiple of G: the group template is repeated; otherwise the first E_loc
      experts of the template are used.  `weight_scaling` multiplies fc1 (weight and bias) and fc2 (weight) by
      sqrt(E G^2 / top_k) with E = total_experts / G (the reference's GELU / softmax-then-top-k rule).
    `split=None` picks split mode iff the expert hidden size differs from the dense one.  `require_granularity_4`
    reproduces the refusal of the warm-start mode for any other split."""
    Hd = int(fc1_w.shape[0])
    if total_experts is None or total_experts <= 0:
        total_experts = local_experts
    if expert_hidden is None:
        expert_hidden = Hd
    if split is None:
        split = expert_hidden != Hd
    if not split:
        rep3 = lambda t: t.unsqueeze(0).repeat(local_experts, 1, 1).contiguous()
        rep2 = lambda t: t.unsqueeze(0).repeat(local_experts, 1).contiguous()
        return rep3(fc1_w), rep2(fc1_b), rep3(fc2_w), rep2(fc2_b)
    G = Hd // int(expert_hidden)
    if G <= 0 or Hd % G != 0:
        raise AssertionError(f"invalid granularity {G} for dense hidden size {Hd}")
    if total_experts % G != 0:
        raise AssertionError(f"total_experts={total_experts} must be divisible by granularity={G}")
    if require_granularity_4 and G != 4:
        raise ValueError(f"deit_warm_start requires dense_hidden / expert_hidden == 4, got {G}")
    scale = 1.0
    if weight_scaling:
        scale = (((total_experts // G) * G * G) / float(max(int(top_k), 1))) ** 0.5
    e1_w = torch.stack((fc1_w * scale).chunk(G, dim=0), dim=0)          # [G, He, D]
    e1_b = torch.stack((fc1_b * scale).chunk(G, dim=0), dim=0)          # [G, He]
    e2_w = torch.stack((fc2_w * scale).chunk(G, dim=1), dim=0)          # [G, D, He]
    if local_experts % G == 0:
        reps = local_experts // G
        return (e1_w.repeat(reps, 1, 1).contiguous(), e1_b.repeat(reps, 1).contiguous(),
                e2_w.repeat(reps, 1, 1).contiguous(), fc2_b.unsqueeze(0).repeat(local_experts, 1).contiguous())
    return (e1_w[:local_experts].contiguous(), e1_b[:local_experts].contiguous(), e2_w[:local_experts].contiguous(),
            fc2_b.unsqueeze(0).repeat(local_experts, 1).contiguous())


def inject_experts_from_dense_mlp(state_dict, moe_blocks: Dict[int, dict], *, moe_mlp_ratio: float = 4.0,
                                  mlp_ratio: float = 4.0, mode: str = "deit_upcycling", weight_scaling: bool = False,
                                  default_top_k: int = 4):
    """State-dict level form of the same helper: for every MoE block index i in `moe_blocks` whose dense keys
    `blocks.{i}.mlp.fc1/fc2.{weight,bias}` are present, ADD `blocks.{i}.mlp.experts.htoh4/h4toh.{weight,bias}` (the dense
    keys stay, as in the reference; blocks without dense keys are skipped).  `moe_blocks[i]` describes the layer the
    tensors are for: {"local_experts", "world_size" (1), "total_experts" (local x world), "top_k" (default_top_k),
    "expert_hidden" (dense hidden)} - e.g. from a constructed model:
        {i: dict(local_experts=b.mlp.num_expert, world_size=b.mlp.world_size, top_k=b.mlp.top_k,
                 expert_hidden=b.mlp.experts.htoh4.weight.shape[1]) for i, b in enumerate(model.blocks) if b.moe}
    A negative `moe_mlp_ratio` means "same as the dense MLP", i.e. `mlp_ratio` (reference :497-499)."""
    mode = str(mode).strip().lower()
    if mode not in ("scratch", "deit_warm_start", "deit_upcycling"):
        raise ValueError(f"Unsupported deit_init_mode '{mode}'")
    force_split = mode == "deit_warm_start"
    ratio = float(mlp_ratio) if moe_mlp_ratio < 0 else float(moe_mlp_ratio)
    for i, info in moe_blocks.items():
        keys = [f"blocks.{i}.mlp.{n}" for n in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")]
        if keys[0] not in state_dict or keys[2] not in state_dict:
            continue
        fc1_w, fc1_b, fc2_w, fc2_b = (state_dict[k] for k in keys)
        e_loc = int(info["local_experts"])
        world = max(int(info.get("world_size", 1)), 1)
        total = int(info.get("total_experts", e_loc * world))
        e1w, e1b, e2w, e2b = upcycle_dense_mlp(
            fc1_w, fc1_b, fc2_w, fc2_b, local_experts=e_loc, total_experts=total if total > 0 else e_loc * world,
            expert_hidden=info.get("expert_hidden") if (force_split or ratio == 1.0) else None,
            top_k=int(info.get("top_k", default_top_k)), split=force_split or ratio == 1.0,
            weight_scaling=weight_scaling, require_granularity_4=force_split)
        state_dict[f"blocks.{i}.mlp.experts.htoh4.weight"] = e1w
        state_dict[f"blocks.{i}.mlp.experts.htoh4.bias"] = e1b
        state_dict[f"blocks.{i}.mlp.experts.h4toh.weight"] = e2w
        state_dict[f"blocks.{i}.mlp.experts.h4toh.bias"] = e2b
    return state_dict


# ----------------------------------------------------------------------------- router keys of a pretrained checkpoint
def convert_gate_keys(state_dict, *, multi_gate: bool, num_tasks: int, task_one_hot: bool = False,
                      gate_task_specific_dim: int = -1, regu_experts_fromtask: bool = False,
                      replicate_all_tasks: bool = False):
    """The router part of the reference's `cvt_state_dict` (/root/reference/utils/common_config.py:47-68): a checkpoint
    trained with ONE shared router per MoE block (`...mlp.gate.w_gate [D, E]`) is adapted, in place, to the layer it is
    loaded into:

      * shared router fed a task vector (`task_one_hot`, not `multi_gate`, not `regu_experts_fromtask`): zero rows are
        appended for the extra gate inputs - `num_tasks` rows if `gate_task_specific_dim < 0`, else that many
        (the new inputs start without influence on the logits);
      * `multi_gate`: every `...gate.w_gate` becomes `...gate.{t}.w_gate`, one copy per task gate, and the shared key is
        removed.  REFERENCE QUIRK kept by default: copies exist for tasks 0 and 1, plus 2 and 3 when `num_tasks == 4`, plus
        2, 3, 4 when `num_tasks == 5` - any other task count gets TWO gates (the rest keep their initialisation under
        `strict=False`).  `replicate_all_tasks=True` writes all `num_tasks` copies instead.
    Other keys are untouched.  Returns the same dict."""
    if task_one_hot and not multi_gate and not regu_experts_fromtask:
        rows = num_tasks if gate_task_specific_dim < 0 else gate_task_specific_dim
        for k in list(state_dict.keys()):
            if "mlp.gate.w_gate" in k:
                w = state_dict[k]
                state_dict[k] = torch.cat((w, torch.zeros((rows, w.shape[-1]))), 0)
    if multi_gate:
        n = num_tasks if (replicate_all_tasks or num_tasks in (4, 5)) else 2
        for k in list(state_dict.keys()):
            if "mlp.gate.w_gate" in k:
                stem = k[:-len("w_gate")]
                for t in range(n):
                    state_dict[f"{stem}{t}.w_gate"] = state_dict[k]
                del state_dict[k]
    return state_dict


# ----------------------------------------------------------------------------- router init for upcycled experts
def auto_virtual_group_size(tot_experts: int, *, local_experts=None, world_size=None, dense_hidden=None,
                            expert_hidden=None) -> int:
    """Size G of a "virtual group" of router columns (/root/reference/utils/helpers.py:715-754): the split granularity
    dense_hidden / expert_hidden when it is whole (else the local expert count, else tot / world, else 1), reduced to a
    common divisor of the local and the total expert count."""
    import math
    tot_experts = int(tot_experts)
    if tot_experts <= 0:
        return 1
    primary = None
    if dense_hidden is not None and expert_hidden is not None and expert_hidden > 0 and dense_hidden % expert_hidden == 0:
        primary = int(dense_hidden // expert_hidden)
    if primary is None or primary <= 0:
        if local_experts is not None and int(local_experts) > 0:
            primary = int(local_experts)
        elif world_size is not None and int(world_size) > 0 and tot_experts % int(world_size) == 0:
            primary = int(tot_experts // int(world_size))
        else:
            primary = 1
    g = int(primary)
    if local_experts is not None and int(local_experts) > 0:
        g = math.gcd(g, int(local_experts))
    g = math.gcd(g, tot_experts)
    if g <= 0 or tot_experts % g != 0:
        g = 1
    return g


def virtual_group_gate_init(like: torch.Tensor, group_size: int, std: float = 0.02) -> torch.Tensor:
    """A router matrix [D_g, E_tot] for experts that were upcycled in groups of `group_size` (reference
    `build_grouped_w_gate`, utils/helpers.py:783-803): N(0, std) everywhere, then the first group's columns repeated for
    every group - the copies of one dense-MLP slice start with the same logit.  Draws from torch's global CPU generator
    exactly like the reference (one normal_ over the full matrix), so a fixed torch seed gives the same bits."""
    d_model, tot = like.shape
    if group_size < 1 or tot % group_size != 0:
        raise AssertionError(f"group_size={group_size} must divide tot_experts={tot}")
    w = torch.empty((d_model, tot), device=like.device, dtype=like.dtype)
    torch.nn.init.normal_(w, mean=0.0, std=std)
    if group_size == 1:
        return w
    proto = torch.tensor_split(w, tot // group_size, dim=1)[0]
    return torch.cat([proto] * (tot // group_size), dim=1).contiguous()


def inject_virtual_group_gate_init(state_dict, model_state: Dict[str, torch.Tensor], moe_blocks: Dict[int, dict],
                                   std: float = 0.02):
    """State-dict level form (reference `_inject_virtual_group_init_for_gates`, utils/helpers.py:757-866): every router key
    of `model_state` (`blocks.{i}[.mlp].gate[.{t}].w_gate`, `blocks.{i}.shared_gate.w_gate`), in `model_state` order, is
    (re)initialised with `virtual_group_gate_init`; G comes from `auto_virtual_group_size` with the block's
    {"local_experts", "world_size"} of `moe_blocks[i]`, the dense hidden size of `blocks.{i}.mlp.fc1.weight` in `state_dict`
    and the expert hidden size of `blocks.{i}.mlp.experts.htoh4.weight` in `model_state` (else `state_dict`)."""
    import re
    pat = re.compile(r"^blocks\.\d+\.(?:mlp\.)?(?:gate(?:\.\d+)?|shared_gate)\.w_gate$")
    keys = [k for k in model_state.keys() if pat.search(k)]
    if not keys:
        raise KeyError("no router w_gate keys (blocks.{i}[.mlp].gate[.{j}].w_gate | blocks.{i}.shared_gate.w_gate)")
    for k in keys:
        ref = state_dict.get(k, model_state[k])
        i = int(re.search(r"^blocks\.(\d+)\.", k).group(1))
        info = moe_blocks.get(i, {})
        world = info.get("world_size", 1)
        world = 1 if world is None or int(world) < 1 else int(world)
        dense_hidden = expert_hidden = None
        if f"blocks.{i}.mlp.fc1.weight" in state_dict:
            dense_hidden = int(state_dict[f"blocks.{i}.mlp.fc1.weight"].shape[0])
        ek = f"blocks.{i}.mlp.experts.htoh4.weight"
        if ek in model_state:
            expert_hidden = int(model_state[ek].shape[1])
        elif ek in state_dict:
            expert_hidden = int(state_dict[ek].shape[1])
        g = auto_virtual_group_size(int(ref.shape[1]), local_experts=info.get("local_experts"), world_size=world,
                                    dense_hidden=dense_hidden, expert_hidden=expert_hidden)
        state_dict[k] = virtual_group_gate_init(ref, g, std).cpu()
    return state_dict

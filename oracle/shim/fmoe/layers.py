"""TEST INFRASTRUCTURE ONLY.  Restatement of fmoe/layers.py (FastMoE @4edeccd):
the FMoE base class attribute set and _fmoe_general_global_forward."""
import tree
import torch
import torch.nn as nn

from .functions import prepare_forward, MOEScatter, MOEGather
from .gates import NaiveGate


def mark_module_parallel_comm(module, comm):
    for p in module.parameters():
        setattr(p, "dp_comm", comm)


def _fmoe_general_global_forward(inp, gate, expert_fn, num_expert, world_size, **kwargs):
    (pos, local_expert_count, global_expert_count, fwd_expert_count, fwd_batch_size) = prepare_forward(
        gate, num_expert, world_size)
    topk = 1
    if len(gate.shape) == 2:
        topk = gate.shape[1]

    def scatter_func(tensor):
        return MOEScatter.apply(
            tensor, torch.div(pos, topk, rounding_mode="floor"),
            local_expert_count, global_expert_count, fwd_batch_size, world_size)

    x = tree.map_structure(scatter_func, inp)
    x = expert_fn(x, fwd_expert_count)

    out_batch_size = tree.flatten(inp)[0].shape[0]
    if len(gate.shape) == 2:
        out_batch_size *= gate.shape[1]

    def gather_func(tensor):
        return MOEGather.apply(
            tensor, pos, local_expert_count, global_expert_count, out_batch_size, world_size)

    return tree.map_structure(gather_func, x)


class FMoE(nn.Module):
    def __init__(self, num_expert=32, d_model=1024, world_size=1, mp_group=None, slice_group=None,
                 moe_group=None, top_k=2, gate=NaiveGate, expert=None, gate_hook=None,
                 mask=None, mask_dict=None):
        super().__init__()
        self.num_expert = num_expert
        self.d_model = d_model
        self.world_size = world_size
        self.slice_group = slice_group if slice_group is not None else mp_group
        if self.slice_group is None:
            self.slice_size, self.slice_rank = 1, 0
        else:
            self.slice_size = self.slice_group.size()
            self.slice_rank = self.slice_group.rank()
        self.top_k = top_k
        if type(expert) is list:
            self.experts = nn.ModuleList([e(d_model) for e in expert])
            self.experts_fused = False
            self.num_expert = num_expert = len(expert)
        elif expert is not None:
            self.experts = nn.ModuleList([expert(d_model) for _ in range(num_expert)])
            self.experts_fused = False
        else:
            self.experts_fused = True
        self.gate = gate(d_model, num_expert, world_size, top_k)
        self.gate_hook = gate_hook
        self.mask = mask
        self.mask_dict = mask_dict
        self.moe_group = moe_group

    def expert_fn(self, inp, fwd_expert_count):
        if self.experts_fused:
            return self.experts(inp, fwd_expert_count)
        raise NotImplementedError("unfused experts are not used by the reference")

    def mark_parallel_comm(self, expert_dp_comm="none"):
        if self.experts is not None:
            comm = expert_dp_comm
            if isinstance(self.experts, list):
                for e in self.experts:
                    mark_module_parallel_comm(e, comm)
            else:
                mark_module_parallel_comm(self.experts, comm)
        mark_module_parallel_comm(self.gate, "gate")

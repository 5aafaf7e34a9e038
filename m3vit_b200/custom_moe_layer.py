"""FMoETransformerMLP -- the M3ViT MoE layer as a drop-in module.

Mirrors /root/reference/models/moe/origin/custom_moe_layer.py:66-314 (returns the
output tensor; the gate keeps `loss`) and models/moe/ckpt/custom_moe_layer.py
(returns `(out, clean_logits, noisy_logits, noise_stddev, top_logits, gates)` so
the Block computes the cv-loss outside torch.utils.checkpoint).  Same constructor
keywords, same `forward(inp, gate_inp=None, task_id=None,
task_specific_feature=None, sem=None)`, same attribute names and state-dict keys:

    experts.htoh4.weight [E_loc, H, D]   experts.htoh4.bias [E_loc, H]
    experts.h4toh.weight [E_loc, D, H]   experts.h4toh.bias [E_loc, D]
    gate.w_gate [D_g, E_tot]   or   gate.{t}.w_gate  (multi_gate)

What changes is everything underneath: fmoe's python ops and `fmoe_cuda` kernels
are replaced by the C-ABI library (include/m3vit_moe.h) through one autograd node
(m3vit_b200/functions.py).  CUDA tensors on a B200 only; no CPU fallback.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import functions as F_
from .noisy_gate_vmoe import NoisyGate_VMoE, balance_loss


class NoisyGate:  # placeholder type: `--moe_gate_type noisy` is outside the named hot path
    def __init__(self, *a, **k):
        raise NotImplementedError(
            "gate=NoisyGate (learned-noise gate, models/moe/noisy_gate.py) is outside the B200 hot path; "
            "use NoisyGate_VMoE (moe_gate_type='noisy_vmoe')")


class NaiveGate:  # fmoe.gates.NaiveGate default of the reference signature; never used by M3ViT configs
    def __init__(self, *a, **k):
        raise NotImplementedError("NaiveGate is not part of the M3ViT hot path; pass gate=NoisyGate_VMoE")


class FMoELinear(nn.Module):
    """Parameter holder with fmoe.linear.FMoELinear's layout and init
    (weight [E, out, in] kaiming_uniform(a=sqrt(5)) per expert, bias [E, out] zeros).
    The arithmetic happens in the grouped-FFN kernels, never in this module."""

    def __init__(self, num_expert, in_feat, out_feat, bias=True, rank=0):
        super().__init__()
        self.num_expert, self.in_feat, self.out_feat, self.rank = num_expert, in_feat, out_feat, rank
        self.weight = nn.Parameter(torch.empty(num_expert, out_feat, in_feat))
        self.bias = nn.Parameter(torch.zeros(num_expert, out_feat)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        for i in range(self.num_expert):
            torch.nn.init.kaiming_uniform_(self.weight[i], a=math.sqrt(5))

    def extra_repr(self):
        return f"num_expert={self.num_expert}, in_features={self.in_feat}, out_features={self.out_feat}, rank={self.rank}"


class _Expert(nn.Module):
    """origin/custom_moe_layer.py:25-44: htoh4 -> activation -> h4toh."""

    def __init__(self, num_expert, d_model, d_hidden, activation, rank=0):
        super().__init__()
        self.htoh4 = FMoELinear(num_expert, d_model, d_hidden, bias=True, rank=rank)
        self.h4toh = FMoELinear(num_expert, d_hidden, d_model, bias=True, rank=rank)
        self.activation = activation


def _check_activation(activation):
    """The fused FFN implements exact-erf GELU followed by Dropout(p=0), which is what
    every reference config builds (origin/vision_transformer_moe.py:248-251)."""
    mods = list(activation.modules()) if isinstance(activation, nn.Module) else []
    gelu = [m for m in mods if isinstance(m, nn.GELU)]
    drop = [m for m in mods if isinstance(m, nn.Dropout)]
    other = [m for m in mods if not isinstance(m, (nn.GELU, nn.Dropout, nn.Sequential, nn.Identity))]
    if len(gelu) != 1 or other or getattr(gelu[0], "approximate", "none") != "none":
        raise NotImplementedError("expert activation must be nn.GELU() (exact erf), optionally followed by Dropout")
    return max([d.p for d in drop], default=0.0)


def _next_dropout_state(layer, device):
    """Expert dropout (the Dropout(p) behind the experts' GELU, origin/vision_transformer_moe.py:248-251): None in eval or
    at p = 0, else (p, rng) with rng = a SNAPSHOT {seed, call counter} int64[2] of the layer's device-side generator state,
    whose counter is then bumped by a stream-ordered add - no host read-back, CUDA-graph capturable (a replay draws new
    masks).  The seed follows torch.manual_seed at the time of the layer's first dropout call."""
    if not (layer.drop_p > 0 and layer.training):
        return None
    st = getattr(layer, "_drop_rng", None)
    if st is None or st.device != device:
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + id(layer)) & 0x7FFFFFFFFFFFFFFF
        st = layer._drop_rng = torch.tensor([seed, 0], dtype=torch.int64, device=device)
    snap = st.clone()
    st[1:].add_(1)
    return (float(layer.drop_p), snap)


class FMoETransformerMLP(nn.Module):
    # flipped by the ckpt subclass
    RETURN_SUMMARIES = False

    def __init__(self, num_expert=32, d_model=1024, d_gate=1024, d_hidden=4096, activation=torch.nn.GELU(),
                 expert_dp_comm="none", expert_rank=0, gate=NaiveGate, world_size=1, top_k=2, vmoe_noisy_std=1,
                 gate_return_decoupled_activation=False, gate_task_specific_dim=-1, multi_gate=False,
                 regu_experts_fromtask=False, num_experts_pertask=-1, num_tasks=-1, regu_sem=False, sem_force=False,
                 regu_subimage=False, expert_prune=False, prune_threshold=0.1,
                 # fmoe.layers.FMoE keywords accepted through **kwargs by the reference
                 mp_group=None, slice_group=None, moe_group=None, gate_hook=None, mask=None, mask_dict=None,
                 # B200-specific: None = follow the input dtype (fp32 in -> fp32 SIMT parity path,
                 # bf16 in -> tcgen05 path); torch.bfloat16 forces the tensor-core path for fp32 inputs
                 compute_dtype: Optional[torch.dtype] = None, **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        for name, val in (("regu_experts_fromtask", regu_experts_fromtask), ("regu_sem", regu_sem),
                          ("sem_force", sem_force), ("regu_subimage", regu_subimage), ("expert_prune", expert_prune),
                          ("gate_return_decoupled_activation", gate_return_decoupled_activation)):
            if val:
                raise NotImplementedError(f"{name}=True is a research branch outside the B200 hot path (SURVEY.md 8b)")
        if slice_group is not None or mp_group is not None:
            raise NotImplementedError("fmoe slice/model parallelism is never enabled by the reference")
        if mask is not None or mask_dict is not None:
            raise NotImplementedError("token masking (mask/mask_dict) is not implemented")
        # ---- fmoe.layers.FMoE attribute set
        self.num_expert = num_expert
        self.d_model = d_model
        self.world_size = world_size
        self.slice_group, self.slice_size, self.slice_rank = None, 1, 0
        self.top_k = top_k
        self.experts_fused = True
        self.gate_hook = gate_hook
        self.mask, self.mask_dict = None, None
        self.moe_group = moe_group
        # ---- reference layer attributes (origin:100-112)
        self.our_d_gate = d_gate
        self.our_d_model = d_model
        self.regu_experts_fromtask = False
        self.num_experts_pertask = num_experts_pertask
        self.num_tasks = num_tasks
        self.regu_sem = self.sem_force = self.regu_subimage = self.expert_prune = False
        self.prune_threshold = prune_threshold
        self.d_hidden = d_hidden
        self.drop_p = _check_activation(activation)
        self.experts = _Expert(num_expert, d_model, d_hidden, activation, rank=expert_rank)
        self.gate_task_specific_dim = gate_task_specific_dim
        self.multi_gate = multi_gate
        d_gate_eff = d_model if gate_task_specific_dim < 0 else d_model + gate_task_specific_dim   # origin:127-130
        if gate is NoisyGate_VMoE or (isinstance(gate, type) and issubclass(gate, NoisyGate_VMoE)):
            def mk():
                return gate(d_gate_eff, num_expert, world_size, top_k, noise_std=vmoe_noisy_std,
                            num_experts_pertask=num_experts_pertask, num_tasks=num_tasks,
                            return_summaries=self.RETURN_SUMMARIES)
            if multi_gate:
                # the reference derives the number of task gates as d_gate - d_model (origin:145-150)
                self.gate = nn.ModuleList([mk() for _ in range(self.our_d_gate - self.our_d_model)])
            else:
                self.gate = mk()
        elif gate is NoisyGate:
            NoisyGate()
        else:
            raise ValueError("No such gating type")                                                   # origin:157-158
        self.mark_parallel_comm(expert_dp_comm)
        env = os.environ.get("M3VIT_MOE_DTYPE", "").lower()
        if compute_dtype is None and env in ("bf16", "bfloat16"):
            compute_dtype = torch.bfloat16
        elif compute_dtype is None and env in ("fp32", "float32"):
            compute_dtype = torch.float32
        self.compute_dtype = compute_dtype
        self._wcache = F_.WeightCache()
        self._ep = None        # set by m3vit_b200.ep.attach() when world_size > 1

    # fmoe.layers.FMoE.mark_parallel_comm: tags read by fmoe.DistributedGroupedDataParallel
    def mark_parallel_comm(self, expert_dp_comm="none"):
        for p in self.experts.parameters():
            setattr(p, "dp_comm", expert_dp_comm)
        for p in self.gate.parameters():
            setattr(p, "dp_comm", "gate")

    def expert_fn(self, inp, fwd_expert_count):
        raise NotImplementedError("experts run inside the fused grouped-FFN kernels; call the layer")

    # ------------------------------------------------------------------------------
    def _select_gate(self, task_id):
        if (task_id is not None) and self.multi_gate:
            return self.gate[task_id]
        # reference: `self.gate(gate_inp, task_id=..)` - with multi_gate and no task_id this calls a
        # ModuleList, which raises TypeError (origin:216-217).  Same here.
        if isinstance(self.gate, nn.ModuleList):
            raise TypeError("'ModuleList' object is not callable: multi_gate=True requires a task_id")
        return self.gate

    def forward(self, inp: torch.Tensor, gate_inp=None, task_id=None, task_specific_feature=None, sem=None,
                fused_norm: Optional[nn.LayerNorm] = None):
        """Reference signature (origin:161).  `fused_norm` is the B200 Block-level extension:
        when given, `inp` is the RAW residual stream and the call returns  inp + MoE(fused_norm(inp))
        with the LayerNorm and the residual add fused into the layer's kernels (MoEBlockMlp uses it)."""
        original_shape = inp.shape
        x = inp.reshape(-1, self.d_model)
        gx = None
        if gate_inp is not None and gate_inp is not inp:
            gx = gate_inp.reshape(-1, gate_inp.shape[-1])
        tf = None
        if (task_id is not None) and (task_specific_feature is not None):
            assert self.multi_gate is False                                                       # origin:177
            tf = task_specific_feature.reshape(-1)
        gate = self._select_gate(task_id)
        out, summaries = self.forward_moe(gate, x, gx, tf, fused_norm)
        out = out.reshape(original_shape)
        if self.RETURN_SUMMARIES:
            return (out, *summaries)
        return out

    def forward_moe(self, gate: NoisyGate_VMoE, x, gx, tf, fused_norm=None):
        if gate.select_idx is not None:
            raise NotImplementedError("gate.select_idx (pruning research path) is not implemented")
        T = x.shape[0]
        cdt = self.compute_dtype or (torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32)
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError(f"unsupported input dtype {x.dtype}")
        nstd = float(gate.noise_stddev())
        noise = gate.draw_noise(T, x.device, as_tensor=fused_norm is not None)     # (the LN-fused gate entry takes a tensor)
        drop = _next_dropout_state(self, x.device)
        if fused_norm is not None:
            if self.world_size > 1 or gx is not None or x.dtype != torch.float32:
                raise NotImplementedError("fused_norm: single-GPU, fp32 residual stream, gate_inp is inp")
            D = x.shape[1]
            ln_w = fused_norm.weight if fused_norm.weight is not None else x.new_ones(D)
            ln_b = fused_norm.bias if fused_norm.bias is not None else x.new_zeros(D)
            res = F_.MoEBlockFunction.apply(
                x, ln_w, ln_b, gate.w_gate, tf, self.experts.htoh4.weight, self.experts.htoh4.bias,
                self.experts.h4toh.weight, self.experts.h4toh.bias, noise, float(fused_norm.eps), self.top_k, nstd,
                cdt, self.RETURN_SUMMARIES, self._wcache, drop)
        elif self.world_size > 1:
            if self._ep is None:
                raise RuntimeError("world_size > 1 needs m3vit_b200.ep.attach(layer, group) before the first forward")
            res = self._ep.forward(self, gate, x, gx, tf, noise, nstd, cdt, drop)
        else:
            res = F_.MoEFunction.apply(
                x, gx, gate.w_gate, tf, self.experts.htoh4.weight, self.experts.htoh4.bias,
                self.experts.h4toh.weight, self.experts.h4toh.bias, noise, self.top_k, nstd, cdt,
                self.RETURN_SUMMARIES, self._wcache, drop)
        out, score, top_vals, clean, noisy, gates, importance, load, idx, counts, cv_loss = res
        self.last_counts = counts                 # device tensor, no sync: for monitoring / tests
        if self.gate_hook is not None:
            self.gate_hook(idx, score, None)                                                       # origin:240-241
        gate._record(clean, noisy, importance, load, top_vals, cv_loss)
        if self.RETURN_SUMMARIES and self.multi_gate:
            # ckpt:214-217: keep the unused task gates in the autograd graph (DDP "marked ready twice")
            others = [p for g in self.gate if g is not gate for p in g.parameters()]
            if others and torch.is_grad_enabled():
                clean = clean + 0.0 * sum(p.sum() for p in others)
        return out, (clean, noisy, gate.noise_stddev(), top_vals, gates)


class FMoETransformerMLPCkpt(FMoETransformerMLP):
    """models/moe/ckpt/custom_moe_layer.py: returns
    (out, clean_logits, noisy_logits, noise_stddev, top_logits, gates)."""
    RETURN_SUMMARIES = True


class TokenFMoETransformerMLP(nn.Module):
    """Experts-only entry of the reference's token-MoE variant
    (/root/reference/models/moe/token/custom_moe_layer.py:55-156): routing (`gate_top_k_idx [T,K]` int64,
    `gate_score [T,K]`) is computed by the caller; the layer runs dispatch -> expert FFN -> combine.
    Same state-dict keys for the experts; no gate parameters."""

    def __init__(self, num_expert=32, d_model=1024, d_gate=1024, d_hidden=4096, activation=torch.nn.GELU(),
                 expert_dp_comm="none", expert_rank=0, world_size=1, top_k=2,
                 compute_dtype: Optional[torch.dtype] = None, **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        if world_size != 1:
            raise NotImplementedError("TokenFMoETransformerMLP: expert parallelism is wired for FMoETransformerMLP only")
        self.num_expert, self.d_model, self.world_size, self.top_k = num_expert, d_model, world_size, top_k
        self.our_d_model = d_model
        self.d_hidden = d_hidden
        self.drop_p = _check_activation(activation)
        self.experts = _Expert(num_expert, d_model, d_hidden, activation, rank=expert_rank)
        for p in self.experts.parameters():
            setattr(p, "dp_comm", expert_dp_comm)
        self.compute_dtype = compute_dtype
        self._wcache = F_.WeightCache()

    def forward(self, inp: torch.Tensor, gate_top_k_idx: torch.Tensor, gate_score: torch.Tensor):
        shape = inp.shape
        if inp.numel() == 0:          # an empty token subset (the token Block routes masked subsets, token/vision_transformer_moe.py:753)
            return inp.clone()
        x = inp.reshape(-1, self.d_model)
        idx = gate_top_k_idx.reshape(-1, self.top_k)
        score = gate_score.reshape(-1, self.top_k)
        cdt = self.compute_dtype or (torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32)
        out = F_.ExpertsFunction.apply(x, idx, score, self.experts.htoh4.weight, self.experts.htoh4.bias,
                                       self.experts.h4toh.weight, self.experts.h4toh.bias, cdt, self._wcache, None,
                                       _next_dropout_state(self, x.device))
        return out.reshape(shape)

"""Block-level adapter: builds the MoE layer from the reference's user-facing
`Block.__init__` vocabulary and reproduces the ckpt Block's cv-loss bookkeeping.

Reference: /root/reference/models/moe/origin/vision_transformer_moe.py:225-283
(Block) and models/moe/ckpt/vision_transformer_moe.py:380-560 (checkpointed Block,
cv-loss helpers :23-87).  Only the MoE half of the Block is in scope (SURVEY.md
section 8): attention / norm / patch-embed stay whatever the caller uses.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .custom_moe_layer import FMoETransformerMLP, FMoETransformerMLPCkpt
from .noisy_gate_vmoe import NoisyGate_VMoE, _gates_to_load, _prob_in_top_k, cv_squared  # noqa: F401


def build_moe_mlp(dim, mlp_ratio=4.0, drop=0.0, act_layer=nn.GELU, moe_mlp_ratio=-1, moe_experts=64, moe_top_k=2,
                  moe_gate_dim=-1, world_size=1, gate_return_decoupled_activation=False, moe_gate_type="noisy",
                  vmoe_noisy_std=1, gate_task_specific_dim=-1, multi_gate=False, regu_experts_fromtask=False,
                  num_experts_pertask=-1, num_tasks=-1, regu_sem=False, sem_force=False, regu_subimage=False,
                  expert_prune=False, variant="origin", compute_dtype=None):
    """Same argument mapping as Block.__init__ (origin/vision_transformer_moe.py:247-271):
    num_expert=moe_experts (already divided by world_size under EP, utils/common_config.py:179-185),
    d_gate=moe_gate_dim (YAML gate_dim = embed_dim + num_tasks for multi_gate),
    d_hidden=int(dim*moe_mlp_ratio), top_k=moe_top_k, gate by moe_gate_type."""
    activation = nn.Sequential(act_layer(), nn.Dropout(drop))
    if moe_gate_dim < 0:
        moe_gate_dim = dim
    if moe_mlp_ratio < 0:
        moe_mlp_ratio = mlp_ratio
    moe_hidden_dim = int(dim * moe_mlp_ratio)
    if moe_gate_type == "noisy_vmoe":
        gate_fun = NoisyGate_VMoE
    elif moe_gate_type == "noisy":
        raise NotImplementedError("moe_gate_type='noisy' (learned-noise NoisyGate) is outside the B200 hot path")
    else:
        raise ValueError("unknow gate type of {}".format(moe_gate_type))
    cls = FMoETransformerMLPCkpt if variant == "ckpt" else FMoETransformerMLP
    return cls(num_expert=moe_experts, d_model=dim, d_gate=moe_gate_dim, d_hidden=moe_hidden_dim,
               world_size=world_size, top_k=moe_top_k, activation=activation, gate=gate_fun,
               gate_return_decoupled_activation=gate_return_decoupled_activation, vmoe_noisy_std=vmoe_noisy_std,
               gate_task_specific_dim=gate_task_specific_dim, multi_gate=multi_gate,
               regu_experts_fromtask=regu_experts_fromtask, num_experts_pertask=num_experts_pertask,
               num_tasks=num_tasks, regu_sem=regu_sem, sem_force=sem_force, regu_subimage=regu_subimage,
               expert_prune=expert_prune, compute_dtype=compute_dtype)


class MoEBlockMlp(nn.Module):
    """The MoE half of a reference Block:  x + drop_path(mlp_drop(mlp(norm2(x), ...))).

    variant="origin": returns x;  the cv-loss lives on the gate (collect with
        collect_noisy_gating_loss, utils/moe_utils.py:201-207).
    variant="ckpt"  : returns (x, cv_loss) computed like ckpt/vision_transformer_moe.py:452-459,538-542.
    """

    def __init__(self, dim, norm_layer=nn.LayerNorm, drop=0.0, drop_path=0.0, variant="origin", fuse=True,
                 **moe_kwargs):
        super().__init__()
        self.variant = variant
        self.fuse = fuse
        self.norm2 = norm_layer(dim)
        self.mlp = build_moe_mlp(dim, drop=drop, variant=variant, **moe_kwargs)
        self.mlp_drop = nn.Dropout(drop)
        if drop_path > 0.0:
            raise NotImplementedError("drop_path > 0: wrap with the caller's DropPath")
        self.drop_path = nn.Identity()
        self.moe_top_k = self.mlp.top_k
        self.tot_expert = self.mlp.num_expert * self.mlp.world_size

    def _fusable(self, x, gate_inp):
        """Block-level fusion: norm2 + residual inside the layer's kernels."""
        drop_active = self.training and self.mlp_drop.p > 0
        return (self.fuse and isinstance(self.norm2, nn.LayerNorm) and x.is_cuda and x.dtype == torch.float32
                and gate_inp is None and self.mlp.world_size == 1 and not drop_active
                and tuple(self.norm2.normalized_shape) == (x.shape[-1],))

    def forward(self, x, gate_inp=None, task_id=None, task_specific_feature=None, sem=None):
        fused = self._fusable(x, gate_inp)
        if fused:
            # x + mlp(norm2(x)) in one pass: ret already contains the residual
            ret = self.mlp(x, None, task_id, task_specific_feature, sem, fused_norm=self.norm2)
        else:
            ret = self.mlp(self.norm2(x), gate_inp, task_id, task_specific_feature, sem)
        if self.variant != "ckpt":
            return ret if fused else x + self.drop_path(self.mlp_drop(ret))
        moe_output, clean_logits, noisy_logits, noise_stddev, top_logits, gates = ret
        x = moe_output if fused else x + self.drop_path(self.mlp_drop(moe_output))
        importance = gates.sum(0)
        if self.moe_top_k < self.tot_expert and abs(noise_stddev) > 1e-6:
            load = _prob_in_top_k(clean_logits, noisy_logits, noise_stddev, top_logits, self.moe_top_k).sum(0)
        else:
            load = _gates_to_load(gates)
        cv_loss = cv_squared(importance) + cv_squared(load) if self.training else 0
        return x, cv_loss


def collect_noisy_gating_loss(model, weight):
    """utils/moe_utils.py:201-207"""
    loss = 0
    for module in model.modules():
        if isinstance(module, NoisyGate_VMoE) and module.has_loss:
            loss += module.get_loss()
    return loss * weight


def collect_moe_activation(model, batch_size, activation_suppress="pool", return_name=False):
    """utils/moe_utils.py:226-250: the router probabilities every gate cached in its last forward, as [B, E] ("pool":
    mean over the tokens of an image), [B, N, E] ("origin") or [B, N*E] ("concat" - the reference's line for it,
    `torch.reshape(activation.shape[0], -1)`, raises a TypeError; the documented intent is implemented here)."""
    gate_activations, names = [], []
    for name, module in model.named_modules():
        if isinstance(module, NoisyGate_VMoE) and module.has_activation:
            activation = module.get_activation()
            c = activation.shape[-1]
            activation = activation.reshape(batch_size, -1, c)
            if activation_suppress == "pool":
                activation = activation.mean(dim=1)
            elif activation_suppress == "concat":
                activation = activation.reshape(activation.shape[0], -1)
            elif activation_suppress != "origin":
                raise ValueError("No activation_suppress of {}".format(activation_suppress))
            gate_activations.append(activation)
            names.append(name)
    return (gate_activations, names) if return_name else gate_activations


def set_moe_layer_train_mode(model):
    """utils/moe_utils.py:303-306: put every MoE layer (and nothing else) into train mode."""
    from .custom_moe_layer import FMoETransformerMLP
    for module in model.modules():
        if isinstance(module, FMoETransformerMLP):
            module.train()

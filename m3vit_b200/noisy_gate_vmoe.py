"""NoisyGate_VMoE -- drop-in for the reference's router module.

Mirrors /root/reference/models/moe/origin/noisy_gate_vmoe.py:15-306 (class
NoisyGate_VMoE(fmoe BaseGate)) and the ckpt twin
(models/moe/ckpt/noisy_gate_vmoe.py:15-274): same constructor arguments, same
parameter (`w_gate` [d_model, tot_expert]), same attributes other reference code
touches (`loss/has_loss/get_loss/set_loss`, `activation/has_activation/
get_activation`, `select_idx`, `top_k`, `noise_std`, `tot_expert`), same
`forward(inp, task_id=None, sem=None)` return conventions.  The arithmetic runs
in ONE fused CUDA kernel (csrc/gate.cu) instead of 6-8 ATen launches.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .functions import GateFunction


class BaseGate(nn.Module):
    """fmoe.gates.base_gate.BaseGate (FastMoE @4edeccd): loss bookkeeping +
    tot_expert = num_expert * world_size."""

    def __init__(self, num_expert, world_size):
        super().__init__()
        self.world_size = world_size
        self.num_expert = num_expert
        self.tot_expert = world_size * num_expert
        self.loss = None

    def forward(self, x):
        raise NotImplementedError("Base gate cannot be directly used for fwd")

    def set_loss(self, loss):
        self.loss = loss

    def get_loss(self, clear=True):
        loss = self.loss
        if clear:
            self.loss = None
        return loss

    @property
    def has_loss(self):
        return self.loss is not None


def cv_squared(x: torch.Tensor) -> torch.Tensor:
    """origin/noisy_gate_vmoe.py:127-141 (also ckpt/vision_transformer_moe.py:73-87)."""
    eps = 1e-10
    if x.shape[0] == 1:
        return torch.Tensor([0])
    return x.float().var() / (x.float().mean() ** 2 + eps)


def _gates_to_load(gates: torch.Tensor) -> torch.Tensor:
    """ckpt/vision_transformer_moe.py:23-31"""
    return (gates > 0).sum(0)


def _prob_in_top_k(clean_values, noisy_values, noise_stddev, noisy_top_values, top_k):
    """origin/noisy_gate_vmoe.py:82-125 / ckpt/vision_transformer_moe.py:33-71.  Kept in torch
    ops (noise>0 training only); bug-compatible: thresholds are top-(k+1)
    PROBABILITIES compared against LOGITS."""
    from torch.distributions.normal import Normal

    batch = clean_values.size(0)
    m = noisy_top_values.size(1)
    top_values_flat = noisy_top_values.flatten()
    pos_in = torch.arange(batch, device=clean_values.device) * m + top_k
    thr_in = torch.unsqueeze(torch.gather(top_values_flat, 0, pos_in), 1)
    is_in = torch.gt(noisy_values, thr_in)
    thr_out = torch.unsqueeze(torch.gather(top_values_flat, 0, pos_in - 1), 1)
    normal = Normal(torch.tensor([0.0], device=clean_values.device), torch.tensor([1.0], device=clean_values.device))
    prob_if_in = normal.cdf((clean_values - thr_in) / noise_stddev)
    prob_if_out = normal.cdf((clean_values - thr_out) / noise_stddev)
    return torch.where(is_in, prob_if_in, prob_if_out)


def balance_loss(importance, load_hard, clean_logits, noisy_logits, noise_stddev, top_logits, top_k, tot_expert,
                 training=True):
    """cv^2(importance) + cv^2(load) exactly as origin/noisy_gate_vmoe.py:267-283."""
    if not training:
        return 0
    if top_k < tot_expert and abs(noise_stddev) > 1e-6:
        load = _prob_in_top_k(clean_logits, noisy_logits, noise_stddev, top_logits, top_k).sum(0)
    else:
        load = load_hard
    return cv_squared(importance) + cv_squared(load)


class NoisyGate_VMoE(BaseGate):
    _UNSUPPORTED = ("regu_experts_fromtask", "regu_sem", "sem_force", "regu_subimage", "return_decoupled_activation")

    def __init__(self, d_model, num_expert, world_size, top_k=2, noise_std=1, no_noise=False,
                 return_decoupled_activation=False, regu_experts_fromtask=False, num_experts_pertask=-1,
                 num_tasks=-1, regu_sem=False, sem_force=False, regu_subimage=False, group_size=4,
                 return_summaries=False):
        super().__init__(num_expert, world_size)
        for name, val in (("return_decoupled_activation", return_decoupled_activation),
                          ("regu_experts_fromtask", regu_experts_fromtask), ("regu_sem", regu_sem),
                          ("regu_subimage", regu_subimage)):
            if val:
                raise NotImplementedError(
                    f"NoisyGate_VMoE({name}=True) is a research branch outside the B200 hot path "
                    "(SURVEY.md 8b); only the default noisy_vmoe routing is implemented")
        self.w_gate = nn.Parameter(torch.zeros(d_model, self.tot_expert), requires_grad=True)
        self.return_decoupled_activation = False
        self.top_k = top_k
        self.no_noise = no_noise
        self.noise_std = noise_std
        self.activation = None
        self.select_idx = None
        self.regu_experts_fromtask = False
        self.num_experts_pertask = num_experts_pertask
        self.num_tasks = num_tasks
        self.regu_sem = False
        self.regu_subimage = False
        self.sem_force = sem_force
        self.group_size = group_size
        self.patch_size = 16
        # origin returns (idx, score) and sets self.loss; ckpt returns the 6-tuple
        self.return_summaries = return_summaries
        self._last_logits = None
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.kaiming_uniform_(self.w_gate, a=math.sqrt(5))      # origin:63-70

    # ---- helpers shared with the fused layer path ---------------------------------
    def noise_stddev(self) -> float:
        """origin:180-181,221-222: a python float; 0 in eval."""
        s = (self.noise_std / self.tot_expert) * self.training
        if self.no_noise:
            s *= 0
        return s

    def draw_noise(self, T, device, as_tensor=False):
        """The router noise of origin:226 (`torch.randn_like(clean_logits) * noise_stddev`).  Nothing when the stddev is 0
        (eval, noise_std 0).  Otherwise, by default, a SNAPSHOT int64[2] = {seed, call counter} of this gate's device-side
        generator state: the gate kernel draws the normals itself (csrc/philox.cuh) and no [T, E] tensor exists; the
        counter is bumped by a stream-ordered add (no host read-back, CUDA-graph capturable).  `gate.strict_rng = True`
        (or as_tensor) draws them with torch.randn instead - the reference's own stream, consumed even at stddev 0 like
        the reference does."""
        strict = getattr(self, "strict_rng", False)
        if abs(self.noise_stddev()) > 0 and not (strict or as_tensor):
            st = getattr(self, "_noise_rng", None)
            if st is None or st.device != device:
                seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + id(self)) & 0x7FFFFFFFFFFFFFFF
                st = self._noise_rng = torch.tensor([seed, 0], dtype=torch.int64, device=device)
            snap = st.clone()
            st[1:].add_(1)
            return snap
        if abs(self.noise_stddev()) > 0 or strict:
            return torch.randn(T, self.tot_expert, device=device, dtype=torch.float32)
        return None

    def _record(self, clean_logits, noisy_logits, importance, load, top_vals, cv_loss=None):
        """origin:267-284: set self.loss (origin variant only) and remember the activation."""
        nstd = self.noise_stddev()
        if not self.return_summaries:
            hard_load = not (self.top_k < self.tot_expert and abs(nstd) > 1e-6)
            if not self.training:
                self.set_loss(0)
            elif hard_load and cv_loss is not None:
                # noise-free path: cv^2(importance) + cv^2(hard load) comes fused out of the
                # route-plan kernel (forward) and is chained inside the gate-backward kernel
                self.set_loss(cv_loss)
            else:
                self.set_loss(balance_loss(importance, load, clean_logits, noisy_logits, nstd, top_vals, self.top_k,
                                           self.tot_expert, self.training))
        self._last_logits = noisy_logits.detach()
        self._act_lead = None          # leading dimensions of the router input (stand-alone forward on an N-d input)

    def get_activation(self, clear=True):
        """origin:284,299-303: softmax probabilities of the last forward, shaped like the router input's leading
        dimensions + [E] (lazy here: the softmax runs only if somebody asks)."""
        if self.activation is None and self._last_logits is not None:
            act = torch.softmax(self._last_logits, dim=1)
            lead = getattr(self, "_act_lead", None)
            self.activation = act.reshape(list(lead) + [-1]).contiguous() if lead else act
        activation = self.activation
        if clear:
            self.activation = None
            self._last_logits = None
        return activation

    @property
    def has_activation(self):
        return self.activation is not None or self._last_logits is not None

    # ---- stand-alone forward ---------------------------------------------------------
    def forward(self, inp, task_id=None, sem=None):
        if self.select_idx is not None:
            raise NotImplementedError("select_idx (expert pruning research path) is not implemented")
        shape_input = list(inp.shape)
        other_dim = shape_input[:-1]
        inp2 = inp.reshape(-1, shape_input[-1])
        T = inp2.shape[0]
        noise = self.draw_noise(T, inp2.device)
        w_gate = self.w_gate
        pad = (-inp2.shape[1]) % 32
        if pad:
            # the kernel walks the router input in 32-column chunks; a width like d_model + task embedding (token-MoE,
            # 384 + 64 is fine, 64 + 16 is not) is zero-padded - trailing +0*0 terms leave every logit bit-identical
            inp2 = torch.nn.functional.pad(inp2, (0, pad))
            w_gate = torch.nn.functional.pad(w_gate, (0, 0, 0, pad))
        (score, top_vals, clean, noisy, gates, importance, cv_loss, load, idx, *_plan) = GateFunction.apply(
            inp2, w_gate, None, noise, self.top_k, float(self.noise_stddev()), self.return_summaries)
        self._record(clean, noisy, importance, load, top_vals, cv_loss)
        self._act_lead = other_dim
        self.last_plan = _plan
        top_k_indices = idx.reshape(other_dim + [self.top_k])
        top_k_gates = score.reshape(other_dim + [self.top_k])
        if self.return_summaries:
            return (top_k_indices, top_k_gates), clean, noisy, self.noise_stddev(), top_vals, gates
        return top_k_indices, top_k_gates


class TokenNoisyGate_VMoE(NoisyGate_VMoE):
    """Router of the reference's token-MoE variant (/root/reference/models/moe/token/noisy_gate_vmoe.py:13-113): same
    arithmetic as NoisyGate_VMoE (softmax over all experts, top-(K+1), the first K probabilities un-renormalised) but it
    ALWAYS returns `((top_k_indices, top_k_gates), clean_logits, noisy_logits, noise_stddev, top_logits, gates)` (:94-101),
    never sets `self.loss` (the token Block computes the balance loss itself from these, token/vision_transformer_moe.py
    :268-300) and records no activation.  The token Block calls it on a VARIABLE subset of the tokens (:753-768); any
    number of rows, including zero, is accepted."""

    def __init__(self, d_model, num_expert, world_size, top_k=2, noise_std=1, no_noise=False, num_experts_pertask=-1,
                 num_tasks=-1):
        super().__init__(d_model, num_expert, world_size, top_k=top_k, noise_std=noise_std, no_noise=no_noise,
                         num_experts_pertask=num_experts_pertask, num_tasks=num_tasks, return_summaries=True)

    def _record(self, *args, **kwargs):      # no loss, no activation (token/noisy_gate_vmoe.py:86-88)
        pass

    def forward(self, inp, task_id=None, sem=None):
        if inp.numel() == 0:                 # an empty subset: nothing to route
            other = list(inp.shape[:-1])
            E, K1 = self.tot_expert, min(self.top_k + 1, self.tot_expert)
            z = inp.new_zeros(0, E, dtype=torch.float32)
            return ((torch.zeros(other + [self.top_k], dtype=torch.int64, device=inp.device),
                     inp.new_zeros(other + [self.top_k], dtype=torch.float32)), z, z, self.noise_stddev(),
                    inp.new_zeros(0, K1, dtype=torch.float32), z)
        return super().forward(inp, task_id=task_id, sem=sem)

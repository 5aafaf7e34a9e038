"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file.  The product path (m3vit_b200/) never
does: it fails loudly when the CUDA extension is missing.

Self-contained PyTorch restatement (CPU, fp32 or fp64) of the reference's MoE
layer hot path.  Each function cites the reference lines it follows
(paths relative to /root/reference):

  gate_forward      models/moe/origin/noisy_gate_vmoe.py:168-297
                    (= models/moe/ckpt/noisy_gate_vmoe.py:80-264,
                       models/moe/gates.py:405-466)
  cv_squared        models/moe/origin/noisy_gate_vmoe.py:127-141
  prob_in_top_k     models/moe/origin/noisy_gate_vmoe.py:82-125
  cv_loss           models/moe/origin/noisy_gate_vmoe.py:267-283,
                    models/moe/ckpt/vision_transformer_moe.py:452-459,538-542
  route_plan        fmoe count_by_gate/assign_pos (FastMoE @4edeccd, external;
                    restated in oracle/shim/fmoe/functions.py)
  expert_ffn        models/moe/origin/custom_moe_layer.py:36-44 (+ FMoELinear)
  layer_forward     models/moe/origin/custom_moe_layer.py:161-314
  block_mlp_forward models/moe/origin/vision_transformer_moe.py:278-283 (MoE half of Block)

PARITY PIN: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  The oracle is pinned instead against outputs of the
reference's own files executed verbatim in the build container through
oracle/shim (oracle/make_golden.py -> tests/golden/*.pt); tests/test_oracle.py
asserts bit-equality of this restatement with those fixtures on CPU.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- gate
def cv_squared(x: torch.Tensor) -> torch.Tensor:
    """origin/noisy_gate_vmoe.py:127-141"""
    eps = 1e-10
    if x.shape[0] == 1:
        return torch.Tensor([0])
    return x.float().var() / (x.float().mean() ** 2 + eps)


def prob_in_top_k(clean_values, noisy_values, noise_stddev, noisy_top_values, top_k):
    """origin/noisy_gate_vmoe.py:82-125.  Bug-compatible: `noisy_top_values`
    are top-(k+1) softmax PROBABILITIES while clean/noisy values are LOGITS."""
    from torch.distributions.normal import Normal

    batch = clean_values.size(0)
    m = noisy_top_values.size(1)
    top_values_flat = noisy_top_values.flatten()
    pos_in = torch.arange(batch, device=clean_values.device) * m + top_k
    thr_in = torch.unsqueeze(torch.gather(top_values_flat, 0, pos_in), 1)
    is_in = torch.gt(noisy_values, thr_in)
    thr_out = torch.unsqueeze(torch.gather(top_values_flat, 0, pos_in - 1), 1)
    normal = Normal(torch.tensor([0.0]), torch.tensor([1.0]))
    prob_if_in = normal.cdf((clean_values - thr_in) / noise_stddev)
    prob_if_out = normal.cdf((clean_values - thr_out) / noise_stddev)
    return torch.where(is_in, prob_if_in, prob_if_out)


def gate_forward(
    gate_inp: torch.Tensor,          # [T, D_g]   (task feature already concatenated)
    w_gate: torch.Tensor,            # [D_g, E]
    top_k: int,
    noise_std: float = 0.0,
    training: bool = False,
    noise: Optional[torch.Tensor] = None,   # [T, E] standard normal; drawn if None
):
    """origin/noisy_gate_vmoe.py:168-297 (default flags).  Returns a dict with
    idx[T,K] int64, score[T,K], top_logits[T,K+1], clean_logits, noisy_logits,
    noise_stddev(float), probs[T,E], gates[T,E], loss (tensor or 0)."""
    E = w_gate.shape[1]
    clean_logits = gate_inp @ w_gate                                   # :179
    raw_noise_stddev = noise_std / E                                   # :180
    noise_stddev = raw_noise_stddev * training                         # :181
    if noise is None:
        noise = torch.randn_like(clean_logits)                         # :226 (RNG consumed even at std 0)
    noisy_logits = clean_logits + (noise * noise_stddev)               # :226
    probs = torch.softmax(noisy_logits, dim=1)                         # :255
    top_logits, top_indices = probs.topk(min(top_k + 1, E), dim=1)     # :256-258
    top_k_logits = top_logits[:, :top_k]                               # :260
    top_k_indices = top_indices[:, :top_k]                             # :261
    zeros = torch.zeros_like(probs, requires_grad=True)                # :264
    gates = zeros.scatter(1, top_k_indices, top_k_logits)              # :265
    if training:                                                       # :267-283
        if top_k < E and abs(noise_stddev) > 1e-6:
            load = prob_in_top_k(clean_logits, noisy_logits, noise_stddev, top_logits, top_k).sum(0)
        else:
            load = (gates > 0).sum(0)
        importance = gates.sum(0)
        loss = cv_squared(importance) + cv_squared(load)
    else:
        loss = 0
    return dict(
        idx=top_k_indices.contiguous(), score=top_k_logits.contiguous(), top_logits=top_logits,
        clean_logits=clean_logits, noisy_logits=noisy_logits, noise_stddev=noise_stddev,
        probs=probs, gates=gates, loss=loss,
    )


def cv_loss_from_summaries(gates, clean_logits, noisy_logits, noise_stddev, top_logits, top_k, training=True):
    """ckpt/vision_transformer_moe.py:452-459 + 538-542: the ckpt variant's
    cv-loss, computed by the Block from the tensors the layer hands back."""
    E = gates.shape[1]
    importance = gates.sum(0)
    if top_k < E and abs(noise_stddev) > 1e-6:
        load = prob_in_top_k(clean_logits, noisy_logits, noise_stddev, top_logits, top_k).sum(0)
    else:
        load = (gates > 0).sum(0)
    if not training:
        return 0, importance, load
    return cv_squared(importance) + cv_squared(load), importance, load


# --------------------------------------------------------------------------- routing
def route_plan(idx: torch.Tensor, num_expert: int, pad: int = 1):
    """Expert counts, (padded) exclusive offsets and the stable queue position of
    every flat (token,k) slot.  Semantics of fmoe's expert_count + assign_pos
    (counts) with a deterministic, stable in-queue order; `pad` rounds each
    expert's queue length up (the B200 layout; pad=1 is the reference layout).

    returns counts[E] int32, offsets[E+1] int32, pos[T*K] int32 (slot -> queue row),
            row_slot[offsets[E]] int32 (queue row -> slot, -1 for padding rows)"""
    flat = idx.reshape(-1).to(torch.int64)
    counts = torch.bincount(flat, minlength=num_expert)[:num_expert]
    padded = (counts + pad - 1) // pad * pad
    offsets = torch.zeros(num_expert + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(padded, 0)
    order = torch.sort(flat, stable=True).indices            # queue order (unpadded)
    start_unpadded = torch.zeros(num_expert + 1, dtype=torch.int64)
    start_unpadded[1:] = torch.cumsum(counts, 0)
    e_sorted = flat[order]
    rank_in_e = torch.arange(flat.numel()) - start_unpadded[e_sorted]
    rows = offsets[e_sorted] + rank_in_e
    pos = torch.empty(flat.numel(), dtype=torch.int64)
    pos[order] = rows
    row_slot = torch.full((int(offsets[-1]),), -1, dtype=torch.int64)
    row_slot[rows] = order
    return counts.int(), offsets.int(), pos.int(), row_slot.int()


def dispatch(x: torch.Tensor, pos: torch.Tensor, top_k: int, n_rows: int):
    """MOEScatter: queue row pos[t*K+k] <- x[t]; padding rows are zero."""
    xq = x.new_zeros(n_rows, x.shape[1])
    t_of_slot = torch.arange(pos.numel()) // top_k
    xq[pos.long()] = x[t_of_slot]
    return xq


# --------------------------------------------------------------------------- experts
def expert_ffn(xq, counts, offsets, w1, b1, w2, b2):
    """origin/custom_moe_layer.py:36-44: per expert e over its queue segment
    h = GELU_erf(x W1[e]^T + b1[e]);  y = h W2[e]^T + b2[e].   W1 [E,H,D], W2 [E,D,H]."""
    yq = xq.new_zeros(xq.shape[0], w2.shape[1])
    for e, n in enumerate(counts.tolist()):
        if n:
            s = int(offsets[e])
            seg = xq[s:s + n]
            h = F.gelu(seg @ w1[e].t() + b1[e])
            yq[s:s + n] = h @ w2[e].t() + b2[e]
    return yq


def combine(yq, pos, score):
    """origin/custom_moe_layer.py:283-297: out[t] = sum_k score[t,k] * y[t*K+k] (bmm)."""
    T, K = score.shape
    y = yq[pos.long()].view(T, K, -1)
    return torch.bmm(score.view(T, 1, K), y).reshape(T, -1)


# --------------------------------------------------------------------------- layer
def layer_forward(
    inp: torch.Tensor,                 # [..., D]
    w_gate: torch.Tensor,              # [D_g, E] (already selected: gate[task_id] if multi_gate)
    w1, b1, w2, b2,
    top_k: int,
    gate_inp: Optional[torch.Tensor] = None,
    task_specific_feature: Optional[torch.Tensor] = None,   # [D_t] or [1, D_t]
    noise_std: float = 0.0,
    training: bool = False,
    noise: Optional[torch.Tensor] = None,
):
    """origin/custom_moe_layer.py:161-314 for world_size == 1, default research
    flags.  Returns (out[..., D], gate dict)."""
    if gate_inp is None:
        gate_inp = inp
    shape = inp.shape
    D = shape[-1]
    x = inp.reshape(-1, D)
    g = gate_inp.reshape(-1, gate_inp.shape[-1])
    if task_specific_feature is not None:                              # :176-179
        g = torch.cat((g, task_specific_feature.reshape(1, -1).repeat(g.shape[0], 1)), dim=-1)
    gd = gate_forward(g, w_gate, top_k, noise_std, training, noise)
    E = w1.shape[0]
    counts, offsets, pos, _ = route_plan(gd["idx"], E, pad=1)
    xq = dispatch(x, pos, top_k, int(offsets[-1]))
    yq = expert_ffn(xq, counts, offsets, w1, b1, w2, b2)
    out = combine(yq, pos, gd["score"])
    gd["counts"] = counts
    return out.reshape(shape), gd


def block_mlp_forward(x, ln_w, ln_b, ln_eps, w_gate, w1, b1, w2, b2, top_k, task_specific_feature=None,
                      noise_std: float = 0.0, training: bool = False, noise=None):
    """The MoE half of the reference Block, origin/vision_transformer_moe.py:282-283 with drop = drop_path = 0:
        x = x + self.drop_path(self.mlp_drop(self.mlp(self.norm2(x), gate_inp, task_id, task_specific_feature, sem)))
    norm2 = nn.LayerNorm(dim, eps) (:246), gate_inp = None.  Returns (x_out, gate dict)."""
    normed = F.layer_norm(x, (x.shape[-1],), ln_w, ln_b, ln_eps)
    out, gd = layer_forward(normed, w_gate, w1, b1, w2, b2, top_k, None, task_specific_feature, noise_std,
                            training, noise)
    return x + out, gd


def min_topk_gap(probs64: torch.Tensor, k_plus_1: int) -> torch.Tensor:
    """Per-token minimum gap between adjacent values among the top-(K+1) (and the
    first excluded one) of fp64 probabilities: certifies that index equality is
    well defined for a fixture (hard part 1)."""
    v = probs64.topk(min(k_plus_1 + 1, probs64.shape[1]), dim=1).values
    return (v[:, :-1] - v[:, 1:]).min(dim=1).values

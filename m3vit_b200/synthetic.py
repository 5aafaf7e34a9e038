"""Synthetic NYUD / PASCAL-Context shaped inputs and weights for the MoE layer.

There is no network for datasets or checkpoints, so benchmarks and parity tests
use seeded synthetic patch tokens shaped like what `Block.norm2` feeds the layer
(/root/reference/models/moe/origin/vision_transformer_moe.py:282): per-token
zero-mean / unit-variance rows of width `d_model`.  Weight init follows the
reference: experts kaiming_uniform(a=sqrt(5)) per expert (fmoe FMoELinear),
`w_gate` kaiming_uniform(a=sqrt(5)) on the [D_g, E] tensor
(origin/noisy_gate_vmoe.py:63-70).  Biases are small non-zero U(-0.02, 0.02)
so that bias bugs are visible.

Everything is generated with an explicit CPU torch.Generator, so the same
(cfg, seed) gives the same tensors in the build container and on the GPU box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Optional

import torch


@dataclass
class MoECase:
    name: str
    batch: int
    tokens: int          # N per image (incl. cls token)
    d_model: int
    d_hidden: int
    num_expert: int
    top_k: int
    num_gates: int = 1   # >1 => multi_gate with that many task gates
    d_task: int = 0      # >0 => shared router with gate_task_specific_dim = d_task

    @property
    def T(self) -> int:
        return self.batch * self.tokens

    @property
    def d_gate(self) -> int:
        return self.d_model + self.d_task

    def dict(self):
        return asdict(self)


# SURVEY.md section 8 config sizes
C1 = MoECase("C1_vits_nyud_b2", batch=2, tokens=1201, d_model=384, d_hidden=384, num_expert=16, top_k=4, num_gates=2)
C3S = MoECase("C3_vitb_pascal_b1", batch=1, tokens=1025, d_model=768, d_hidden=768, num_expert=16, top_k=4, num_gates=5)
C4S = MoECase("C4_taskcond_b1", batch=1, tokens=513, d_model=384, d_hidden=384, num_expert=16, top_k=4, d_task=64)


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def make_weights(case: MoECase, seed: int):
    """Returns dict(w_gate=[G][D_g,E], w1[E,H,D], b1[E,H], w2[E,D,H], b2[E,D])."""
    gen = torch.Generator().manual_seed(1000 + seed)
    E, D, H = case.num_expert, case.d_model, case.d_hidden
    w1 = _uniform(gen, (E, H, D), 1.0 / math.sqrt(D))
    w2 = _uniform(gen, (E, D, H), 1.0 / math.sqrt(H))
    b1 = _uniform(gen, (E, H), 0.02)
    b2 = _uniform(gen, (E, D), 0.02)
    # kaiming_uniform_(a=sqrt(5)) on a [D_g, E] tensor: fan_in = E
    w_gate = [_uniform(gen, (case.d_gate, E), 1.0 / math.sqrt(E)) for _ in range(max(case.num_gates, 1))]
    return dict(w_gate=w_gate, w1=w1, b1=b1, w2=w2, b2=b2)


def make_tokens(case: MoECase, seed: int, gen: Optional[torch.Generator] = None):
    gen = gen or torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(case.batch, case.tokens, case.d_model, generator=gen, dtype=torch.float32)
    x = (x - x.mean(-1, keepdim=True)) / x.std(-1, keepdim=True, unbiased=False)
    return x


def make_case(case: MoECase, seed: int, min_gap: float = 1e-5, max_rounds: int = 50):
    """Inputs + weights for `case`.  Tokens whose fp64 top-(K+1) softmax
    probabilities (under ANY of the case's gates) have an adjacent gap below
    `min_gap` are resampled, so routing indices are well defined independent of
    fp32 summation order (SURVEY.md section 7 hard part 1)."""
    w = make_weights(case, seed)
    gen = torch.Generator().manual_seed(2000 + seed)
    x = make_tokens(case, seed, gen)
    tfeat = None
    if case.d_task > 0:
        tfeat = torch.randn(case.d_task, generator=gen, dtype=torch.float32)
    grad_out = torch.randn(case.batch, case.tokens, case.d_model, generator=gen, dtype=torch.float32)
    flat = x.view(-1, case.d_model)
    resampled = 0
    if min_gap > 0:
        kk = min(case.top_k + 2, case.num_expert)
        for _ in range(max_rounds):
            bad = torch.zeros(flat.shape[0], dtype=torch.bool)
            for wg in w["w_gate"]:
                g = flat.double()
                if tfeat is not None:
                    g = torch.cat((g, tfeat.double().view(1, -1).expand(g.shape[0], -1)), 1)
                p = torch.softmax(g @ wg.double(), 1)
                v = p.topk(kk, 1).values
                bad |= (v[:, :-1] - v[:, 1:]).min(1).values < min_gap
            n_bad = int(bad.sum())
            if n_bad == 0:
                break
            resampled += n_bad
            fresh = torch.randn(n_bad, case.d_model, generator=gen, dtype=torch.float32)
            fresh = (fresh - fresh.mean(-1, keepdim=True)) / fresh.std(-1, keepdim=True, unbiased=False)
            flat[bad] = fresh
        else:
            raise RuntimeError("could not certify routing gaps")
    return dict(x=x, grad_out=grad_out, task_feat=tfeat, resampled=resampled, **w)


def make_block_case(case: MoECase, seed: int, min_gap: float = 1e-5, max_rounds: int = 50):
    """Inputs for the Block-level path  x + mlp(norm2(x))  (SURVEY.md 8 f1): a RAW residual stream
    (per-token mean ~ N(0,1), scale ~ U(0.5,3)) and non-trivial LayerNorm affine parameters.  Routing gaps are
    certified in fp64 on the normalised gate input, like make_case."""
    d = make_case(case, seed, min_gap=0.0)
    gen = torch.Generator().manual_seed(3000 + seed)
    D = case.d_model
    ln_w = 1.0 + 0.1 * torch.randn(D, generator=gen, dtype=torch.float32)
    ln_b = 0.1 * torch.randn(D, generator=gen, dtype=torch.float32)
    eps = 1e-6                                      # partial(nn.LayerNorm, eps=1e-6) in the reference ViT

    def raw(n):
        z = torch.randn(n, D, generator=gen, dtype=torch.float32)
        scale = 0.5 + 2.5 * torch.rand(n, 1, generator=gen, dtype=torch.float32)
        shift = torch.randn(n, 1, generator=gen, dtype=torch.float32)
        return z * scale + shift
    flat = raw(case.T)
    tfeat = d["task_feat"]
    kk = min(case.top_k + 2, case.num_expert)
    resampled = 0
    for _ in range(max_rounds):
        g = torch.nn.functional.layer_norm(flat.double(), (D,), ln_w.double(), ln_b.double(), eps)
        if tfeat is not None:
            g = torch.cat((g, tfeat.double().view(1, -1).expand(g.shape[0], -1)), 1)
        bad = torch.zeros(flat.shape[0], dtype=torch.bool)
        for wg in d["w_gate"]:
            v = torch.softmax(g @ wg.double(), 1).topk(kk, 1).values
            bad |= (v[:, :-1] - v[:, 1:]).min(1).values < min_gap
        n_bad = int(bad.sum())
        if n_bad == 0 or min_gap <= 0:
            break
        resampled += n_bad
        flat[bad] = raw(n_bad)
    else:
        raise RuntimeError("could not certify routing gaps")
    d.update(x=flat.view(case.batch, case.tokens, D), ln_w=ln_w, ln_b=ln_b, ln_eps=eps, resampled=resampled)
    return d


def device_tokens(T: int, d_model: int, seed: int, device, dtype=torch.float32):
    """Benchmark-size tokens generated on the device (no gap certification)."""
    gen = torch.Generator(device=device).manual_seed(2000 + seed)
    x = torch.randn(T, d_model, generator=gen, device=device, dtype=torch.float32)
    x = (x - x.mean(-1, keepdim=True)) / x.std(-1, keepdim=True, unbiased=False)
    return x.to(dtype)

"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/token_*.pt.

Runs the reference's token-MoE files, unmodified, on CPU (fmoe / tree / timm shimmed as in make_golden.py):

    /root/reference/models/moe/token/noisy_gate_vmoe.py     TokenNoisyGate_VMoE.forward        (:41-102)
    /root/reference/models/moe/token/custom_moe_layer.py    TokenFMoETransformerMLP.forward    (:88-156)

in the way the token Block drives them (models/moe/token/vision_transformer_moe.py:753-790): a boolean mask picks a
VARIABLE subset of the B*N tokens, only those are routed and sent through the experts, and the result is added back
into the full tensor at their positions.

    python oracle/make_token_golden.py            # needs /root/reference
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

torch.set_num_threads(8)

CASES = [  # name, B, N, D, H, E, K, Dt (task embedding width; 0 = per-task gates), keep fraction
    ("token_multigate_e16k4", 2, 65, 64, 64, 16, 4, 0, 0.6),
    ("token_taskemb_e8k2", 3, 40, 64, 128, 8, 2, 16, 0.35),
    ("token_sparse_e16k1", 1, 97, 64, 64, 16, 1, 0, 0.05),
    ("token_all_e4k4", 1, 33, 64, 64, 4, 4, 0, 1.0),
]


def run(name, B, N, D, H, E, K, Dt, keep, seed=0):
    from models.moe.token.custom_moe_layer import TokenFMoETransformerMLP
    from models.moe.token.noisy_gate_vmoe import TokenNoisyGate_VMoE
    g = torch.Generator().manual_seed(seed + len(name))
    gate = TokenNoisyGate_VMoE(D + Dt, E, 1, top_k=K, noise_std=0)
    mlp = TokenFMoETransformerMLP(num_expert=E, d_model=D, d_gate=D, d_hidden=H,
                                  activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), top_k=K)
    with torch.no_grad():
        gate.w_gate.copy_((torch.rand(D + Dt, E, generator=g) * 2 - 1) * 0.3)
        for p in mlp.experts.parameters():
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (0.1 if p.dim() == 2 else 1.0 / p.shape[-1] ** 0.5))
    gate.train(); mlp.train()
    x = torch.randn(B, N, D, generator=g)
    x = ((x - x.mean(-1, keepdim=True)) / x.std(-1, keepdim=True)).requires_grad_(True)        # what norm2 hands the layer
    task_emb = torch.randn(Dt, generator=g).requires_grad_(True) if Dt else None
    mask = torch.rand(B * N, generator=g) < keep
    go = torch.randn(B * N, D, generator=g)
    # ---- token/vision_transformer_moe.py:753-790, attention / drop-path / cache left out
    x_flat = x.reshape(B * N, D)
    compute_idx = mask.nonzero(as_tuple=False).squeeze(1)
    Kc = int(compute_idx.numel())
    out_flat = x_flat.clone()
    rec = {}
    if Kc > 0:
        sub = x_flat[compute_idx]
        gate_inp = sub if task_emb is None else torch.cat([sub, task_emb.unsqueeze(0).expand(Kc, -1)], dim=-1)
        (idx, score), clean, noisy, nstd, top_logits, gates = gate(gate_inp, task_id=0)
        eo = mlp(sub, idx, score)
        out_flat[compute_idx] = out_flat[compute_idx] + eo
        rec.update(idx=idx.detach(), score=score.detach(), clean=clean.detach(), top_logits=top_logits.detach(),
                   gates=gates.detach(), noise_stddev=float(nstd), expert_out=eo.detach())
        # fp64 gap certificate for the routing comparison
        p64 = torch.softmax((gate_inp.detach().double() @ gate.w_gate.detach().double()), 1)
        tv = p64.topk(min(K + 1, E), 1).values
        rec["min_gap"] = float((tv[:, :-1] - tv[:, 1:]).min()) if tv.shape[1] > 1 else 1.0
    (out_flat * go).sum().backward()
    rec.update(name=name, shape=(B, N, D, H, E, K, Dt), x=x.detach(), mask=mask, grad_out=go, out=out_flat.detach(),
               dx=x.grad.detach(), task_emb=None if task_emb is None else task_emb.detach(),
               dtask_emb=None if task_emb is None else task_emb.grad.detach(),
               w_gate=gate.w_gate.detach(), dw_gate=gate.w_gate.grad.detach(),
               params={n: p.detach() for n, p in mlp.experts.named_parameters()},
               grads={n: p.grad.detach() for n, p in mlp.experts.named_parameters()})
    return rec


def main():
    for c in CASES:
        rec = run(*c)
        assert rec.get("min_gap", 1.0) > 1e-5, (c[0], rec["min_gap"])
        dst = os.path.join(ROOT, "tests", "golden", c[0] + ".pt")
        torch.save(rec, dst)
        print(c[0], "tokens kept", int(rec["mask"].sum()), "of", rec["mask"].numel(), "min gap", rec.get("min_gap"),
              os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()

"""GPU: the token-MoE entry points against fixtures produced by the reference's own
models/moe/token/{noisy_gate_vmoe,custom_moe_layer}.py (oracle/make_token_golden.py), driven the way the token Block
drives them (token/vision_transformer_moe.py:753-790): a mask selects a variable subset of the tokens, only those are
routed and run through the experts, the result is added back at their positions."""
import glob
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "token_*.pt")))


def nerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def build(rec, dev, cdt):
    import m3vit_b200 as M
    B, N, D, H, E, K, Dt = rec["shape"]
    gate = M.TokenNoisyGate_VMoE(D + Dt, E, 1, top_k=K, noise_std=0).to(dev).train()
    mlp = M.TokenFMoETransformerMLP(num_expert=E, d_model=D, d_gate=D, d_hidden=H,
                                    activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), top_k=K, compute_dtype=cdt).to(dev).train()
    with torch.no_grad():
        gate.w_gate.copy_(rec["w_gate"])
        for n, p in mlp.experts.named_parameters():
            p.copy_(rec["params"][n])
    return gate, mlp


def drive(rec, gate, mlp, dev):
    B, N, D, H, E, K, Dt = rec["shape"]
    x = rec["x"].to(dev).requires_grad_(True)
    emb = rec["task_emb"].to(dev).requires_grad_(True) if rec["task_emb"] is not None else None
    x_flat = x.reshape(B * N, D)
    compute_idx = rec["mask"].to(dev).nonzero(as_tuple=False).squeeze(1)
    Kc = int(compute_idx.numel())
    sub = x_flat[compute_idx]
    gate_inp = sub if emb is None else torch.cat([sub, emb.unsqueeze(0).expand(Kc, -1)], dim=-1)
    (idx, score), clean, noisy, nstd, top_logits, gates = gate(gate_inp, task_id=0)
    eo = mlp(sub, idx, score)
    out_flat = x_flat.clone()
    out_flat[compute_idx] = out_flat[compute_idx] + eo
    (out_flat * rec["grad_out"].to(dev)).sum().backward()
    return dict(idx=idx, score=score, clean=clean, top_logits=top_logits, gates=gates, nstd=nstd, eo=eo, out=out_flat,
                dx=x.grad, demb=None if emb is None else emb.grad)


@pytest.mark.parametrize("fname", FIXTURES)
def test_token_gate_and_experts_match_reference_fp32(fname):
    dev = torch.device("cuda:0")
    rec = torch.load(os.path.join(GOLD, fname), weights_only=False)
    gate, mlp = build(rec, dev, torch.float32)
    got = drive(rec, gate, mlp, dev)
    assert torch.equal(got["idx"].cpu(), rec["idx"])                                   # routing: bit-exact
    torch.testing.assert_close(got["score"].cpu(), rec["score"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["clean"].cpu(), rec["clean"], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(got["top_logits"].cpu(), rec["top_logits"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["gates"].cpu(), rec["gates"], rtol=1e-5, atol=1e-6)
    assert float(got["nstd"]) == rec["noise_stddev"] == 0.0
    assert not gate.has_loss                                                            # the token gate never sets one
    torch.testing.assert_close(got["eo"].detach().cpu(), rec["expert_out"], rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(got["out"].detach().cpu(), rec["out"], rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(got["dx"].cpu(), rec["dx"], rtol=2e-4, atol=2e-5)
    if rec["task_emb"] is not None:
        torch.testing.assert_close(got["demb"].cpu(), rec["dtask_emb"], rtol=2e-4, atol=2e-4)
    assert nerr(gate.w_gate.grad.cpu(), rec["dw_gate"]) <= 2e-4
    for n, p in mlp.experts.named_parameters():
        assert nerr(p.grad.cpu(), rec["grads"][n]) <= 2e-4, n


@pytest.mark.parametrize("fname", [f for f in FIXTURES if "e16k4" in f or "all" in f])
def test_token_path_bf16_within_tolerance(fname):
    dev = torch.device("cuda:0")
    rec = torch.load(os.path.join(GOLD, fname), weights_only=False)
    B, N, D, H, E, K, Dt = rec["shape"]
    if D % 128 or H % 128:
        pytest.skip("tensor-core path needs D, H multiples of 128")
    gate, mlp = build(rec, dev, torch.bfloat16)
    got = drive(rec, gate, mlp, dev)
    assert torch.equal(got["idx"].cpu(), rec["idx"])
    assert nerr(got["out"].detach().cpu(), rec["out"]) <= 3e-2


def test_empty_token_subset():
    """the token Block guards K > 0 (token/vision_transformer_moe.py:759); the modules accept an empty subset anyway"""
    import m3vit_b200 as M
    dev = torch.device("cuda:0")
    gate = M.TokenNoisyGate_VMoE(64, 8, 1, top_k=2, noise_std=0).to(dev)
    mlp = M.TokenFMoETransformerMLP(num_expert=8, d_model=64, d_hidden=64, activation=nn.GELU(), top_k=2).to(dev)
    sub = torch.zeros(0, 64, device=dev)
    (idx, score), clean, noisy, nstd, top, gates = gate(sub)
    assert idx.shape == (0, 2) and score.shape == (0, 2) and clean.shape == (0, 8) and top.shape == (0, 3)
    assert mlp(sub, idx, score).shape == (0, 64)

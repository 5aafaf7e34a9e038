"""GPU: the stochastic parts of the layer - router noise drawn in the gate kernel and expert dropout in
the FFN's GELU epilogue (csrc/philox.cuh).  The generator is not torch's, so parity with the reference
(noisy_gate_vmoe.py:226 torch.randn_like; nn.Dropout behind the experts' GELU, origin/vision_transformer_moe.py:248-251)
is STATISTICAL: moments / KS distance of the noise, keep-rate and scale of the mask, and exact algebra (same mask in
forward and backward, eval = identity) wherever the stream does not matter."""
import math

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def rng_state(seed, counter, dev):
    return torch.tensor([seed, counter], dtype=torch.int64, device=dev)


def test_router_noise_is_standard_normal_and_counter_based():
    from m3vit_b200 import ops
    from scipy import stats
    dev = torch.device("cuda:0")
    T, D, E, K = 20000, 64, 16, 4
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(T, D, device=dev, generator=g)
    wg = torch.randn(D, E, device=dev, generator=g) * 0.1
    std = 0.25
    a = ops.gate_fwd(x, wg, K, noise=rng_state(1234, 0, dev), noise_stddev=std)
    n = ((a.noisy_logits - a.clean_logits) / std).flatten().double().cpu()
    assert abs(float(n.mean())) < 5 / math.sqrt(n.numel()) and abs(float(n.var()) - 1) < 0.02
    assert abs(float((n ** 4).mean()) - 3) < 0.1                                  # kurtosis of a normal
    assert stats.kstest(n.numpy()[::7], "norm").statistic < 0.01
    m = n.view(T, E)
    assert abs(float((m[:, :-1] * m[:, 1:]).mean())) < 0.01 and abs(float((m[:-1] * m[1:]).mean())) < 0.01   # no lag-1 correlation
    # same state -> the same noise (backward / recompute see what the forward saw); next counter or other seed -> fresh noise
    b = ops.gate_fwd(x, wg, K, noise=rng_state(1234, 0, dev), noise_stddev=std)
    assert torch.equal(a.noisy_logits, b.noisy_logits) and torch.equal(a.idx, b.idx)
    for st in (rng_state(1234, 1, dev), rng_state(1235, 0, dev)):
        c = ops.gate_fwd(x, wg, K, noise=st, noise_stddev=std)
        n2 = ((c.noisy_logits - c.clean_logits) / std).flatten().double().cpu()
        assert abs(float((n * n2).mean())) < 0.01
    # routing statistics agree with the torch.randn path (the reference's stream): per-expert load within 4 sigma
    t = ops.gate_fwd(x, wg, K, noise=torch.randn(T, E, device=dev, generator=g), noise_stddev=std)
    la = torch.bincount(a.idx.flatten(), minlength=E).double().cpu()
    lt = torch.bincount(t.idx.flatten(), minlength=E).double().cpu()
    assert float(((la - lt).abs() / (la + lt).sqrt().clamp_min(1)).max()) < 4.5


def test_noisy_gate_module_draws_in_kernel_and_trains():
    import m3vit_b200 as M
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    gate = M.NoisyGate_VMoE(64, 16, 1, top_k=4, noise_std=1.0).to(dev).train()
    x = torch.randn(500, 64, device=dev, requires_grad=True)
    idx1, s1 = gate(x)
    loss = gate.get_loss()
    (s1.sum() + loss).backward()                                # the CDF load estimator (torch ops) back-propagates too
    assert torch.isfinite(x.grad).all() and torch.isfinite(gate.w_gate.grad).all()
    idx2, _ = gate(x)
    assert not torch.equal(idx1, idx2)                          # the counter moved: fresh noise
    gate.eval()
    i3, _ = gate(x)
    i4, _ = gate(x)
    assert torch.equal(i3, i4)                                  # no noise in eval (noise_stddev * training)


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_expert_dropout_mask_rate_scale_and_backward(cdt, p):
    """fc2 = identity, b2 = 0: yq IS the dropped hidden activation, which exposes the mask; the backward pass must use
    that very mask (torch autograd with the mask applied explicitly is the reference)."""
    from m3vit_b200 import ops
    dev = torch.device("cuda:0")
    T, K, E, D = 3000, 2, 8, 128
    H = D
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(T, D, generator=gen).to(dev)
    idx = torch.stack([torch.randperm(E, generator=gen)[:K] for _ in range(T)]).to(dev)
    w1 = (torch.randn(E, H, D, generator=gen) / D ** 0.5).to(dev)
    b1 = (torch.randn(E, H, generator=gen) * 0.1).to(dev)
    w2 = torch.eye(D).repeat(E, 1, 1).to(dev)
    b2 = torch.zeros(E, D, device=dev)
    plan = ops.route_plan(idx, E)
    n = int(plan.offsets[-1])
    xq = ops.dispatch_fwd(x, plan, K, out_dtype=cdt)
    if cdt == torch.bfloat16:
        w1c, w1t = ops.cast_weights_bf16(w1, True, True)
        w2c, w2t = ops.cast_weights_bf16(w2, True, True)
    else:
        w1c, w2c, w1t, w2t = w1, w2, None, None
    st = rng_state(77, 5, dev)
    h0, _ = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)                              # no dropout
    hd, saved = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, drop=(p, st))
    hd2, _ = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, drop=(p, st.clone()))
    assert torch.equal(hd[:n], hd2[:n])                                          # a pure function of (state, row, col)
    h0f, hdf = h0[:n].float(), hd[:n].float()
    live = h0f.abs() > 1e-3
    keep = (hdf != 0) & live
    rate = float(keep.sum()) / float(live.sum())
    assert abs(rate - (1 - p)) < 4 * math.sqrt(p * (1 - p) / float(live.sum())) + 1e-3
    tol = 1e-5 if cdt == torch.float32 else 1.6e-2
    ratio = (hdf / h0f)[keep]
    assert float((ratio - 1 / (1 - p)).abs().max()) <= tol * (1 / (1 - p)) + tol  # kept values are scaled by 1/(1-p)
    # a different call counter gives an independent mask
    hd3, _ = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, drop=(p, rng_state(77, 6, dev)))
    both = float((((hd3[:n].float() != 0) & keep).sum())) / float(live.sum())
    assert abs(both - (1 - p) ** 2) < 0.01
    # columns / rows of the mask are balanced (no stuck bits)
    colrate = keep.float().sum(0) / live.float().sum(0).clamp_min(1)
    assert float((colrate - (1 - p)).abs().max()) < 0.05
    # ---- backward with the same mask
    mask_scale = torch.where(hdf != 0, torch.full_like(hdf, 1 / (1 - p)), torch.zeros_like(hdf))
    dyq = torch.zeros_like(hd)
    dyq[:n] = (torch.randn(n, D, generator=gen) * 0.1).to(dev).to(cdt)
    pad = torch.ones(n, dtype=torch.bool, device=dev)
    pad[plan.pos.long()] = False
    dyq[:n][pad] = 0                              # padding rows carry zero gradients (combine_bwd writes them so)
    dxq, dw1, db1, dw2, db2 = ops.ffn_bwd(xq, saved, dyq, plan, w1c, w2c, w1t, w2t, drop=(p, st))
    xr = xq[:n].float().requires_grad_(True)
    w1r = (w1c.float() if cdt == torch.bfloat16 else w1).clone().requires_grad_(True)
    b1r = b1.clone().requires_grad_(True)
    off = plan.offsets.cpu().tolist()
    cnt = plan.counts.cpu().tolist()
    outs = []
    for e in range(E):
        r = slice(off[e], off[e] + cnt[e])
        z = xr[r] @ w1r[e].t() + b1r[e]
        outs.append((torch.nn.functional.gelu(z) * mask_scale[r], r))
    loss = sum((o * dyq[r].float()).sum() for o, r in outs)
    loss.backward()
    valid = torch.zeros(n, dtype=torch.bool, device=dev)
    for e in range(E):
        valid[off[e]:off[e] + cnt[e]] = True

    def nerr(a, b):
        return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))
    gtol = 2e-4 if cdt == torch.float32 else 3e-2
    assert nerr(dxq[:n][valid], xr.grad[valid]) < gtol
    assert nerr(dw1, w1r.grad) < gtol and nerr(db1, b1r.grad) < gtol
    # dW2 = dyq^T h_dropped
    want_dw2 = torch.stack([dyq[off[e]:off[e] + cnt[e]].float().t() @ hdf[off[e]:off[e] + cnt[e]] for e in range(E)])
    assert nerr(dw2, want_dw2) < gtol


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_layer_with_expert_dropout(cdt):
    """drop_rate 0.1 (configs/nyud/vit_moe/*drop0.1*.yml): trains, eval is the deterministic layer, the mask changes from
    call to call, and the expectation over masks is the eval output (dropout is unbiased up to fc2's linearity)."""
    import m3vit_b200 as M
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    act = nn.Sequential(nn.GELU(), nn.Dropout(0.1))
    layer = M.FMoETransformerMLP(num_expert=8, d_model=128, d_gate=128, d_hidden=128, activation=act, gate=M.NoisyGate_VMoE,
                                 top_k=2, vmoe_noisy_std=0, compute_dtype=cdt).to(dev)
    ref = M.FMoETransformerMLP(num_expert=8, d_model=128, d_gate=128, d_hidden=128,
                               activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE, top_k=2,
                               vmoe_noisy_std=0, compute_dtype=cdt).to(dev)
    ref.load_state_dict(layer.state_dict())
    x = torch.randn(4, 200, 128, device=dev)
    layer.eval(); ref.eval()
    with torch.no_grad():
        e1, e2 = layer(x), ref(x)
    assert torch.equal(e1, e2)                                                    # Dropout is the identity in eval
    layer.train()
    xg = x.clone().requires_grad_(True)
    o1 = layer(xg)
    o1.sum().backward()
    assert torch.isfinite(xg.grad).all() and layer.experts.htoh4.weight.grad is not None
    with torch.no_grad():
        o2 = layer(x)
        assert not torch.equal(o1.detach(), o2)                                   # a fresh mask per call
        acc = torch.zeros_like(e1, dtype=torch.float32)
        N = 64
        for _ in range(N):
            acc += layer(x).float()
    err = (acc / N - e2.float()).abs().mean() / e2.float().abs().mean()
    assert float(err) < 0.08

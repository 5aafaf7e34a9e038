"""GPU parity of the drop-in layer (m3vit_b200.FMoETransformerMLP[Ckpt]) against the
golden fixtures produced by the reference's own layer code (oracle/make_golden.py).

Tolerances: fp32 path rtol 2e-4 / atol 2e-5 on outputs and activation grads (different
fp32 summation order vs MKL), weight grads normalised error <= 2e-4; routing indices and
expert counts EXACT."""
import pytest
import torch
import torch.nn as nn

import m3vit_b200 as M
from helpers import all_fixtures, load_fixture

pytestmark = pytest.mark.gpu

ALL = all_fixtures()


def build_layer(case, data, variant, dev, compute_dtype=None):
    cls = M.FMoETransformerMLPCkpt if variant == "ckpt" else M.FMoETransformerMLP
    multi = case.num_gates > 1
    layer = cls(num_expert=case.num_expert, d_model=case.d_model,
                d_gate=case.d_model + (case.num_gates if multi else 0), d_hidden=case.d_hidden,
                activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE, world_size=1,
                top_k=case.top_k, vmoe_noisy_std=0,
                gate_task_specific_dim=(case.d_task if case.d_task > 0 else -1), multi_gate=multi,
                compute_dtype=compute_dtype).to(dev)
    with torch.no_grad():
        layer.experts.htoh4.weight.copy_(data["w1"]); layer.experts.htoh4.bias.copy_(data["b1"])
        layer.experts.h4toh.weight.copy_(data["w2"]); layer.experts.h4toh.bias.copy_(data["b2"])
        gates = layer.gate if multi else [layer.gate]
        for g, w in zip(gates, data["w_gate"]):
            g.w_gate.copy_(w)
    return layer


def nerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("fname", ALL)
def test_layer_fp32_matches_reference(fname):
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    layers = {v: build_layer(case, data, v, dev) for v in ("origin", "ckpt")}
    for (variant, task, mode), rec in fx["tasks"].items():
        layer = layers[variant]
        layer.train(mode == "train")
        layer.zero_grad(set_to_none=True)
        x = data["x"].to(dev).requires_grad_(True)
        kwargs = {}
        tf = None
        if data["task_feat"] is not None:
            tf = data["task_feat"].to(dev).requires_grad_(True)
            kwargs = dict(task_id=0, task_specific_feature=tf)
        elif task is not None:
            kwargs = dict(task_id=task)
        cap = {}
        layer.gate_hook = lambda idx, score, _: cap.update(idx=idx.detach(), score=score.detach())
        ret = layer(x, **kwargs)
        gate_mod = layer.gate[task] if task is not None else layer.gate
        if variant == "origin":
            out = ret
            loss = gate_mod.get_loss(clear=False)
        else:
            out, clean, noisy, nstd, top_logits, gates = ret
            importance = gates.sum(0)
            load = (gates > 0).sum(0)
            loss = M.cv_squared(importance) + M.cv_squared(load)
            torch.testing.assert_close(clean.detach().cpu()[::stride], rec["clean_logits"], rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(importance.detach().cpu(), rec["importance"], rtol=1e-5, atol=1e-5)
            assert torch.equal(load.cpu(), rec["load"])
            assert nstd == rec["noise_stddev"]
        # routing: bit exact
        assert torch.equal(cap["idx"].cpu().to(torch.int16), rec["idx"])
        assert torch.equal(layer.last_counts.cpu(), rec["counts"])
        torch.testing.assert_close(cap["score"].cpu(), rec["score"], rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(out.detach().reshape(case.T, -1).cpu()[::stride], rec["out"], rtol=2e-4, atol=2e-5)
        if mode != "train":
            assert loss == 0
            continue
        assert float(loss) == pytest.approx(rec["loss"], rel=1e-4)
        (out * data["grad_out"].to(dev)).sum().backward(retain_graph=True)
        torch.testing.assert_close(x.grad.reshape(case.T, -1).cpu()[::stride], rec["dx"], rtol=2e-4, atol=2e-5)
        if tf is not None:
            torch.testing.assert_close(tf.grad.cpu(), rec["dtask_feat"], rtol=2e-4, atol=2e-4)
        E = case.num_expert
        for name, p in layer.named_parameters():
            want = rec["grads"].get(name, "missing")
            if isinstance(want, str):
                continue
            if want is None:
                # unused task gates: no grad (origin) or an all-zero grad (ckpt's DDP trick)
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                continue
            got = p.grad.cpu()
            if got.shape != want.shape:
                got = got[[0, E - 1]][:, ::max(stride, 4)]
                s, a = rec["grads"][name + ".checksum"]
                assert abs(float(p.grad.double().sum()) - s) <= 2e-4 * a + 1e-6, name
            assert nerr(got, want) <= 2e-4, (name, nerr(got, want))
        # cv-loss gradient alone
        layer.zero_grad(set_to_none=True)
        x.grad = None
        loss.backward()
        torch.testing.assert_close(x.grad.reshape(case.T, -1).cpu()[::stride], rec["cv_dx"], rtol=1e-3, atol=1e-7)
        assert nerr(gate_mod.w_gate.grad.cpu(), rec["cv_dw_gate"]) <= 1e-3


@pytest.mark.parametrize("fname", [f for f in ALL if f.startswith(("S3", "S8", "C1", "C3", "C4"))])
def test_layer_bf16_within_tolerance(fname):
    """bf16 tensor-core path vs the fp32 reference fixture: routing exact (router stays fp32),
    outputs / activation grads within 3e-2 of the fixture's max magnitude."""
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    layer = build_layer(case, data, "origin", dev, compute_dtype=torch.bfloat16)
    for (variant, task, mode), rec in fx["tasks"].items():
        if variant != "origin" or mode != "train":
            continue
        layer.train(True)
        layer.zero_grad(set_to_none=True)
        x = data["x"].to(dev).requires_grad_(True)
        kwargs = dict(task_id=task) if task is not None else {}
        if data["task_feat"] is not None:
            kwargs = dict(task_id=0, task_specific_feature=data["task_feat"].to(dev))
        out = layer(x, **kwargs)
        assert torch.equal(layer.last_counts.cpu(), rec["counts"])
        assert nerr(out.detach().reshape(case.T, -1).cpu()[::stride], rec["out"]) <= 3e-2
        (out * data["grad_out"].to(dev)).sum().backward()
        assert nerr(x.grad.reshape(case.T, -1).cpu()[::stride], rec["dx"]) <= 3e-2
        gname = f"gate.{task}.w_gate" if task is not None else "gate.w_gate"
        assert nerr(dict(layer.named_parameters())[gname].grad.cpu(), rec["grads"][gname]) <= 3e-2
        assert nerr(layer.experts.h4toh.bias.grad.cpu(), rec["grads"]["experts.h4toh.bias"]) <= 3e-2
        assert nerr(layer.experts.htoh4.bias.grad.cpu(), rec["grads"]["experts.htoh4.bias"]) <= 3e-2
        # expert weight gradients of the tcgen05 wgrad kernels against the reference's own fp32 gradients
        for name in ("experts.htoh4.weight", "experts.h4toh.weight"):
            p, want = dict(layer.named_parameters())[name], rec["grads"][name]
            got = p.grad.cpu()
            if got.shape != want.shape:      # large fixtures keep experts 0 and E-1, every max(stride, 4)-th row, + a checksum
                got = got[[0, case.num_expert - 1]][:, ::max(stride, 4)]
                s, a = rec["grads"][name + ".checksum"]
                assert abs(float(p.grad.double().sum()) - s) <= 3e-2 * a + 1e-6, name
            assert nerr(got, want) <= 3e-2, (name, nerr(got, want))


@pytest.mark.parametrize("fname", [f for f in ALL if f.startswith(("S3", "S8", "C1", "C4"))])
def test_layer_bf16_tokens_match_fp32_tokens(fname):
    """A bf16 model hands the layer bf16 tokens (and expects bf16 out / dx).  Against the SAME layer fed the same
    (bf16-representable) values as fp32 tokens: routing identical (the router reads the tokens exactly and computes in
    fp32 either way), expert counts identical, outputs / dx equal up to the bf16 rounding of the results (2^-8 relative
    per element -> 1e-2 of the tensor's max magnitude), parameter gradients within 1e-2."""
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture(fname)
    layer = build_layer(case, data, "origin", dev, compute_dtype=torch.bfloat16).train()
    xb = data["x"].to(dev).bfloat16()
    gb = data["grad_out"].to(dev).bfloat16()
    task = 0 if case.num_gates > 1 else None
    res = {}
    for name, x0, g0 in (("f32", xb.float(), gb.float()), ("bf16", xb, gb)):
        layer.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        kwargs = dict(task_id=task) if task is not None else {}
        if data["task_feat"] is not None:
            kwargs = dict(task_id=0, task_specific_feature=data["task_feat"].to(dev))
        cap = {}
        layer.gate_hook = lambda idx, score, _: cap.update(idx=idx.detach().clone(), score=score.detach().clone())
        out = layer(x, **kwargs)
        assert out.dtype == x0.dtype
        (out * g0).sum().backward()
        assert x.grad.dtype == x0.dtype
        res[name] = dict(out=out.detach().float(), dx=x.grad.float(), idx=cap["idx"], score=cap["score"],
                         counts=layer.last_counts.clone(),
                         grads={n: p.grad.detach().clone() for n, p in layer.named_parameters() if p.grad is not None})
    layer.gate_hook = None
    a, b = res["f32"], res["bf16"]
    assert torch.equal(a["idx"], b["idx"]) and torch.equal(a["counts"], b["counts"])
    assert torch.equal(a["score"], b["score"])
    assert nerr(b["out"], a["out"]) <= 1e-2
    assert nerr(b["dx"], a["dx"]) <= 1e-2
    assert a["grads"].keys() == b["grads"].keys()
    for n in a["grads"]:
        assert nerr(b["grads"][n], a["grads"][n]) <= 1e-2, n


def test_full_size_properties_bf16_and_fp32():
    """BASELINE-size batch (ViT-S / NYUD, B=32): size-independent properties.
    (1) sum of counts == T*K, (2) permutation invariance: shuffling the tokens
    shuffles the outputs, (3) linearity of combine in the scores via two-task gates:
    eval-mode determinism across two calls (bit-identical)."""
    dev = torch.device("cuda:0")
    from m3vit_b200.synthetic import MoECase, make_weights, device_tokens
    case = MoECase("C2_b32", batch=32, tokens=1201, d_model=384, d_hidden=384, num_expert=16, top_k=4, num_gates=2)
    w = make_weights(case, 0)
    for cdt in (torch.float32, torch.bfloat16):
        layer = build_layer(case, dict(w, x=None), "origin", dev, compute_dtype=cdt).eval()
        x = device_tokens(case.T, case.d_model, 0, dev)
        with torch.no_grad():
            y1 = layer(x, task_id=0)
            c1 = layer.last_counts.clone()
            y2 = layer(x, task_id=0)
            assert torch.equal(y1, y2), "forward must be deterministic"
            assert int(c1.sum()) == case.T * case.top_k
            perm = torch.randperm(case.T, device=dev)
            yp = layer(x[perm], task_id=0)
            assert torch.equal(layer.last_counts, c1)
            torch.testing.assert_close(yp, y1[perm], rtol=0, atol=0)


def test_token_moe_experts_only_entry():
    """f2: TokenFMoETransformerMLP.forward(inp, gate_top_k_idx, gate_score) (models/moe/token/custom_moe_layer.py:88-156):
    externally supplied routing, compared with the oracle's dispatch / expert FFN / combine."""
    from oracle import moe_oracle as O
    dev = torch.device("cuda:0")
    E, K, D, H, B, N = 16, 2, 128, 256, 2, 61
    gen = torch.Generator().manual_seed(3)
    layer = M.TokenFMoETransformerMLP(num_expert=E, d_model=D, d_gate=D, d_hidden=H, top_k=K,
                                      activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0))).to(dev)
    with torch.no_grad():
        layer.experts.htoh4.bias.uniform_(-0.1, 0.1)
        layer.experts.h4toh.bias.uniform_(-0.1, 0.1)
    x = torch.randn(B, N, D, generator=gen)
    idx = torch.stack([torch.randperm(E, generator=gen)[:K] for _ in range(B * N)])
    score = torch.rand(B * N, K, generator=gen)
    g = torch.randn(B, N, D, generator=gen)
    xd, sd = x.to(dev).requires_grad_(True), score.to(dev).requires_grad_(True)
    out = layer(xd, idx.to(dev), sd)
    out.backward(g.to(dev))
    w1, b1 = layer.experts.htoh4.weight.detach().cpu(), layer.experts.htoh4.bias.detach().cpu()
    w2, b2 = layer.experts.h4toh.weight.detach().cpu(), layer.experts.h4toh.bias.detach().cpu()
    xr, sr = x.clone().requires_grad_(True), score.clone().requires_grad_(True)
    c, o, p, _ = O.route_plan(idx, E, 1)
    ref = O.combine(O.expert_ffn(O.dispatch(xr.view(-1, D), p, K, int(o[-1])), c, o, w1, b1, w2, b2), p, sr)
    ref.backward(g.view(-1, D))
    torch.testing.assert_close(out.detach().cpu().view(-1, D), ref.detach(), rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(xd.grad.cpu(), xr.grad, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(sd.grad.cpu(), sr.grad, rtol=2e-4, atol=2e-4)


def test_noisy_training_path_matches_oracle_with_same_noise():
    """f4: vmoe_noisy_std > 0 in training.  With gate.strict_rng the noise is drawn by torch on the device; replaying the
    same generator state gives the oracle the identical noise tensor, so routing and the normal-CDF load loss
    (origin/noisy_gate_vmoe.py:82-125,267-283) can be compared directly."""
    from oracle import moe_oracle as O
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture("S8_d128h256_g2_s0.pt")
    layer = build_layer(case, data, "origin", dev)
    for g in layer.gate:
        g.noise_std = 1.0
        g.strict_rng = True          # torch.randn, the reference's stream (default: drawn inside the gate kernel)
    layer.train()
    x = data["x"].to(dev).requires_grad_(True)
    torch.manual_seed(1234)
    out = layer(x, task_id=1)
    loss = layer.gate[1].get_loss()
    (out.sum() + loss).backward()
    torch.manual_seed(1234)
    noise = torch.randn(case.T, case.num_expert, device=dev).cpu()
    xr = data["x"].clone().requires_grad_(True)
    wg = data["w_gate"][1].clone().requires_grad_(True)
    ref, gd = O.layer_forward(xr, wg, data["w1"], data["b1"], data["w2"], data["b2"], case.top_k,
                              noise_std=1.0, training=True, noise=noise)
    (ref.sum() + gd["loss"]).backward()
    p64 = torch.softmax(gd["noisy_logits"].double(), 1)
    ok = O.min_topk_gap(p64, case.top_k + 1) > 1e-5
    assert int((~ok).sum()) <= 2
    torch.testing.assert_close(out.detach().cpu().view(case.T, -1)[ok], ref.detach().view(case.T, -1)[ok], rtol=2e-4, atol=2e-5)
    assert float(loss) == pytest.approx(float(gd["loss"]), rel=1e-3)
    if bool(ok.all()):
        assert nerr(layer.gate[1].w_gate.grad.cpu(), wg.grad) <= 1e-3
        torch.testing.assert_close(x.grad.cpu().view(case.T, -1), xr.grad.view(case.T, -1), rtol=1e-3, atol=1e-5)


def test_layer_is_cuda_graph_capturable():
    """No host synchronisation anywhere on the single-GPU path: forward + backward of the layer can be
    captured into a CUDA graph and replayed (the reference's fmoe path syncs every layer)."""
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture("S8_d128h256_g2_s0.pt")
    for cdt in (torch.float32, torch.bfloat16):
        layer = build_layer(case, data, "origin", dev, compute_dtype=cdt).train()
        x = data["x"].to(dev).requires_grad_(True)
        go = data["grad_out"].to(dev)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):                      # warm-up (allocator, weight cache) outside capture
                layer.zero_grad(set_to_none=True); x.grad = None
                layer(x, task_id=0).backward(go)
        torch.cuda.current_stream().wait_stream(s)
        want_dx, want_out = x.grad.clone(), layer(x, task_id=0).detach().clone()
        layer.zero_grad(set_to_none=True); x.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = layer(x, task_id=0)
            out.backward(go)
        x.grad.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want_out)
        torch.testing.assert_close(x.grad, want_dx, rtol=0, atol=0)


def test_layer_under_make_graphed_callables_is_bit_identical():
    """Small batches are launch-bound (reference config C1: B = 2): torch.cuda.make_graphed_callables captures the
    layer's forward AND backward (tools/graphed_c1.py: 579 -> 207 us per fwd+bwd call at T = 2402).  The graphed
    module must reproduce the eager results bit for bit, including the balance-loss gradient."""
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture("S8_d128h256_g2_s0.pt")      # bf16 tensor-core path needs D, H multiples of 128

    class TaskCall(nn.Module):
        def __init__(self, layer, task):
            super().__init__()
            self.layer, self.task = layer, task

        def forward(self, x):
            return self.layer(x, task_id=self.task), self.layer.gate[self.task].get_loss()

    layer = build_layer(case, data, "origin", dev, compute_dtype=torch.bfloat16).train()
    go = data["grad_out"].to(dev)
    w = torch.tensor(0.01, device=dev)

    def step(mod, x):
        layer.zero_grad(set_to_none=True)
        out, loss = mod(x)
        torch.autograd.backward([out, loss], [go, w])
        return out.detach().clone(), x.grad.clone(), layer.gate[1].w_gate.grad.clone(), layer.experts.h4toh.weight.grad.clone()

    xe = data["x"].to(dev).requires_grad_(True)
    want = step(TaskCall(layer, 1), xe)
    graphed = torch.cuda.make_graphed_callables(TaskCall(layer, 1), (data["x"].to(dev).requires_grad_(True),),
                                                allow_unused_input=True)      # gate[0] gets no gradient
    for _ in range(2):                                                         # replay twice: static buffers are reused
        xg = data["x"].to(dev).requires_grad_(True)
        got = step(graphed, xg)
        for a, b in zip(got, want):
            assert torch.equal(a, b)


def test_graphed_layer_follows_optimizer_steps():
    """a graph captured with a warm weight cache must not keep training on the bf16 expert weights it was
    captured with - the fp32 -> bf16 casts are part of the captured graph, so a replay after optimizer.step() equals the
    eager layer with the updated weights."""
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture("S8_d128h256_g2_s0.pt")
    layer = build_layer(case, data, "origin", dev, compute_dtype=torch.bfloat16).train()
    opt = torch.optim.SGD(layer.parameters(), lr=0.5)
    go = data["grad_out"].to(dev)

    class TaskCall(nn.Module):
        def __init__(self, layer):
            super().__init__()
            self.layer = layer

        def forward(self, x):
            return self.layer(x, task_id=1)

    graphed = torch.cuda.make_graphed_callables(TaskCall(layer), (data["x"].to(dev).requires_grad_(True),),
                                                allow_unused_input=True)
    for it in range(3):
        layer.zero_grad(set_to_none=True)
        xg = data["x"].to(dev).requires_grad_(True)
        out_g = graphed(xg)
        out_g.backward(go)
        got = (out_g.detach().clone(), xg.grad.clone(), layer.experts.htoh4.weight.grad.clone())
        layer.zero_grad(set_to_none=True)
        xe = data["x"].to(dev).requires_grad_(True)
        out_e = layer(xe, task_id=1)
        out_e.backward(go)
        want = (out_e.detach(), xe.grad, layer.experts.htoh4.weight.grad)
        for a, b in zip(got, want):
            assert torch.equal(a, b), it
        opt.step()                                   # changes the fp32 masters (and bumps their version)
    # in-place update through .data does not bump the version: invalidate() is the documented way
    w = layer.experts.htoh4.weight
    before = layer(data["x"].to(dev), task_id=1).detach().clone()
    w.data.mul_(1.5)
    layer._wcache.invalidate()
    after = layer(data["x"].to(dev), task_id=1).detach()
    assert not torch.equal(before, after)


def test_standalone_gate_activation_keeps_leading_dims():
    """origin/noisy_gate_vmoe.py:284,299-303: `get_activation()` returns the softmax probabilities shaped like the router
    input's leading dimensions + [E]; `clear=True` forgets them."""
    import m3vit_b200 as M
    dev = torch.device("cuda:0")
    gate = M.NoisyGate_VMoE(64, 8, 1, top_k=2, noise_std=0).to(dev).eval()
    x = torch.randn(3, 5, 64, device=dev)
    idx, score = gate(x)
    assert idx.shape == (3, 5, 2) and score.shape == (3, 5, 2)
    assert gate.has_activation
    act = gate.get_activation(clear=False)
    assert act.shape == (3, 5, 8)
    want = torch.softmax(x.reshape(-1, 64) @ gate.w_gate, dim=1).reshape(3, 5, 8)
    torch.testing.assert_close(act, want, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(act.gather(-1, idx), score, rtol=1e-5, atol=1e-6)
    assert gate.get_activation() is not None and not gate.has_activation

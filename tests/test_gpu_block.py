"""GPU parity of the Block-level fusion:  x + mlp(norm2(x))  with norm2 and the
residual add fused into the layer's kernels (m3vit_b200.MoEBlockMlp -> MoEBlockFunction ->
m3_ln_stats / m3_ln_fold_gate / m3_gate_fwd_ln / m3_dispatch_fwd_ln / m3_combine_fwd_res /
m3_gate_bwd_ln / m3_ln_bwd_res), against fixtures produced by the reference's own Block.forward
(origin/vision_transformer_moe.py:274-283, attention stubbed to zero; oracle/make_golden.py).

Tolerances: fp32 path rtol 2e-4 / atol 2e-5 (outputs, activation grads), parameter grads normalised
error <= 2e-4; routing indices and expert counts EXACT; bf16 path <= 3e-2 normalised."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import m3vit_b200 as M
from m3vit_b200 import ops
from m3vit_b200.synthetic import MoECase, make_block_case
from helpers import block_fixtures, load_fixture

pytestmark = pytest.mark.gpu

ALL = block_fixtures()


def nerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def build_block(case, data, dev, compute_dtype=None, fuse=True, variant="origin"):
    multi = case.num_gates > 1
    blk = M.MoEBlockMlp(case.d_model, norm_layer=lambda d: nn.LayerNorm(d, eps=data["ln_eps"]), variant=variant,
                        fuse=fuse, moe_mlp_ratio=case.d_hidden / case.d_model, moe_experts=case.num_expert,
                        moe_top_k=case.top_k, moe_gate_dim=case.d_model + (case.num_gates if multi else 0),
                        moe_gate_type="noisy_vmoe", vmoe_noisy_std=0,
                        gate_task_specific_dim=(case.d_task if case.d_task > 0 else -1), multi_gate=multi,
                        compute_dtype=compute_dtype).to(dev)
    with torch.no_grad():
        blk.norm2.weight.copy_(data["ln_w"]); blk.norm2.bias.copy_(data["ln_b"])
        e = blk.mlp.experts
        e.htoh4.weight.copy_(data["w1"]); e.htoh4.bias.copy_(data["b1"])
        e.h4toh.weight.copy_(data["w2"]); e.h4toh.bias.copy_(data["b2"])
        gates = blk.mlp.gate if multi else [blk.mlp.gate]
        for g, w in zip(gates, data["w_gate"]):
            g.w_gate.copy_(w)
    return blk


def run_block(blk, data, task, dev):
    x = data["x"].to(dev).requires_grad_(True)
    kwargs, tf = {}, None
    if data["task_feat"] is not None:
        tf = data["task_feat"].to(dev).requires_grad_(True)
        kwargs = dict(task_id=0, task_specific_feature=tf)
    elif task is not None:
        kwargs = dict(task_id=task)
    cap = {}
    blk.mlp.gate_hook = lambda idx, score, _: cap.update(idx=idx.detach(), score=score.detach())
    out = blk(x, **kwargs)
    return x, tf, out, cap


@pytest.mark.parametrize("fname", ALL)
def test_fused_block_fp32_matches_reference_block(fname):
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    blk = build_block(case, data, dev)
    launches0 = ops.launch_count
    for (task, mode), rec in fx["tasks"].items():
        blk.train(mode == "train")
        blk.zero_grad(set_to_none=True)
        x, tf, out, cap = run_block(blk, data, task, dev)
        gate_mod = blk.mlp.gate[task] if task is not None else blk.mlp.gate
        assert torch.equal(cap["idx"].cpu().to(torch.int16), rec["idx"])            # routing: bit exact
        assert torch.equal(blk.mlp.last_counts.cpu(), rec["counts"])
        torch.testing.assert_close(cap["score"].cpu(), rec["score"], rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(out.detach().reshape(case.T, -1).cpu()[::stride], rec["out"], rtol=2e-4, atol=2e-5)
        if mode != "train":
            continue
        loss = gate_mod.get_loss(clear=False)
        assert float(loss) == pytest.approx(rec["loss"], rel=1e-4)
        (out * data["grad_out"].to(dev)).sum().backward(retain_graph=True)
        torch.testing.assert_close(x.grad.reshape(case.T, -1).cpu()[::stride], rec["dx"], rtol=2e-4, atol=2e-5)
        if tf is not None:
            torch.testing.assert_close(tf.grad.cpu(), rec["dtask_feat"], rtol=2e-4, atol=2e-4)
        E = case.num_expert
        for name, p in blk.named_parameters():
            want = rec["grads"].get(name, "missing")
            if isinstance(want, str):
                continue
            if want is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                continue
            got = p.grad.cpu()
            if got.shape != want.shape:
                got = got[[0, E - 1]][:, ::max(stride, 4)]
                s, a = rec["grads"][name + ".checksum"]
                assert abs(float(p.grad.double().sum()) - s) <= 2e-4 * a + 1e-6, name
            assert nerr(got, want) <= 2e-4, (name, nerr(got, want))
        # cv-loss gradient alone (flows through the router into norm2 and x)
        blk.zero_grad(set_to_none=True)
        x.grad = None
        loss.backward()
        torch.testing.assert_close(x.grad.reshape(case.T, -1).cpu()[::stride], rec["cv_dx"], rtol=2e-3, atol=1e-7)
        assert nerr(blk.norm2.weight.grad.cpu(), rec["cv_dln_w"]) <= 2e-3
        assert nerr(blk.norm2.bias.grad.cpu(), rec["cv_dln_b"]) <= 2e-3
    assert ops.launch_count > launches0          # the CUDA library ran (no torch LayerNorm path)


@pytest.mark.parametrize("fname", [f for f in ALL if f.startswith(("B3", "B4"))])
def test_fused_block_bf16_within_tolerance(fname):
    dev = torch.device("cuda:0")
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    blk = build_block(case, data, dev, compute_dtype=torch.bfloat16)
    for (task, mode), rec in fx["tasks"].items():
        if mode != "train":
            continue
        blk.train(True)
        blk.zero_grad(set_to_none=True)
        x, tf, out, cap = run_block(blk, data, task, dev)
        assert torch.equal(blk.mlp.last_counts.cpu(), rec["counts"])               # router stays fp32
        assert nerr(out.detach().reshape(case.T, -1).cpu()[::stride], rec["out"]) <= 3e-2
        (out * data["grad_out"].to(dev)).sum().backward()
        assert nerr(x.grad.reshape(case.T, -1).cpu()[::stride], rec["dx"]) <= 3e-2
        assert nerr(blk.norm2.weight.grad.cpu(), rec["grads"]["norm2.weight"]) <= 3e-2
        assert nerr(blk.norm2.bias.grad.cpu(), rec["grads"]["norm2.bias"]) <= 3e-2


def test_block_ops_against_torch():
    """Each f1 entry point alone, against torch on the GPU (fp32)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for (T, D, E, K) in [(1, 64, 16, 4), (37, 64, 8, 2), (300, 384, 16, 4), (5000, 768, 64, 2), (129, 1024, 32, 1)]:
        x = (torch.randn(T, D, device=dev) * (0.5 + 2 * torch.rand(T, 1, device=dev)) + torch.randn(T, 1, device=dev))
        gamma, beta = 1 + 0.1 * torch.randn(D, device=dev), 0.1 * torch.randn(D, device=dev)
        wg = (torch.rand(D, E, device=dev) * 2 - 1) / E ** 0.5
        eps = 1e-6
        ln = ops.ln_prepare(x, gamma, beta, eps, wg)
        xn = F.layer_norm(x, (D,), gamma, beta, eps)
        mu = x.double().mean(1)
        var = x.double().var(1, unbiased=False)
        torch.testing.assert_close(ln.mean.double(), mu, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ln.rstd.double(), (var + eps).rsqrt(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ln.w_fold, gamma[:, None] * wg - (gamma[:, None] * wg).mean(0, keepdim=True),
                                   rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ln.gb[0].double(), (gamma.double()[:, None] * wg.double()).sum(0), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(ln.gb[1].double(), (beta.double()[:, None] * wg.double()).sum(0), rtol=1e-5, atol=1e-5)
        # gate on raw x == gate on LayerNorm(x)
        g = ops.gate_fwd_ln(x, ln, K)
        ref_logits = (xn.double() @ wg.double()).float()
        torch.testing.assert_close(g.clean_logits, ref_logits, rtol=2e-5, atol=2e-5)
        g0 = ops.gate_fwd(xn.contiguous(), wg, K)
        assert float((g.idx != g0.idx).float().mean()) <= 2e-3      # near-tie tokens may differ (random data)
        # dispatch with LN on the fly == dispatch of the normalised tokens
        plan = ops.route_plan(g.idx, E, M._lib.PAD_ROWS, g.imp_partial, g.load_partial)
        for dt in (torch.float32, torch.bfloat16):
            xq = ops.dispatch_fwd_ln(x, ln, plan, K, out_dtype=dt)
            xq0 = ops.dispatch_fwd(xn.contiguous(), plan, K, out_dtype=dt)
            n = int(plan.offsets[-1])
            torch.testing.assert_close(xq[:n].float(), xq0[:n].float(), rtol=1e-5 if dt == torch.float32 else 1e-2,
                                       atol=1e-5 if dt == torch.float32 else 1e-2)
            # combine + residual
            yq = torch.randn(plan.cap_rows, D, device=dev).to(dt)
            out = ops.combine_fwd_res(yq, plan, g.score, x)
            out0 = x + ops.combine_fwd(yq, plan, g.score)
            torch.testing.assert_close(out, out0, rtol=1e-6, atol=1e-6)
        # LayerNorm backward + residual
        dxn, dres = torch.randn(T, D, device=dev), torch.randn(T, D, device=dev)
        xr = x.clone().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        F.layer_norm(xr, (D,), gr, br, eps).backward(dxn)
        dx, dg, db = ops.ln_bwd_res(dxn, x, ln, dres)
        torch.testing.assert_close(dx, xr.grad + dres, rtol=1e-4, atol=1e-5)
        assert nerr(dg, gr.grad) <= 1e-4 and nerr(db, br.grad) <= 1e-4
        # deterministic
        dx2, dg2, db2 = ops.ln_bwd_res(dxn, x, ln, dres)
        assert torch.equal(dg, dg2) and torch.equal(db, db2) and torch.equal(dx, dx2)
        # gate backward with x normalised on load
        dscore = torch.randn(T, K, device=dev)
        dz, dw, _ = ops.gate_bwd_ln(x, ln, wg, g.clean_logits, g.idx_full, K, dscore=dscore)
        dz0, dw0, _, _ = ops.gate_bwd(xn.contiguous(), wg, g.clean_logits, g.idx_full, K, dscore=dscore)
        assert torch.equal(dz, dz0)
        assert nerr(dw, dw0) <= 1e-4


def test_fused_block_full_size_equals_unfused():
    """BASELINE-size batch (ViT-S / NYUD, B=32, T=38 432): the fused Block path against the same layer fed by
    torch's LayerNorm + residual add (fuse=False), fp32 and bf16."""
    dev = torch.device("cuda:0")
    case = MoECase("full", batch=32, tokens=1201, d_model=384, d_hidden=384, num_expert=16, top_k=4, num_gates=2)
    data = make_block_case(case, 0, min_gap=0.0)
    g = data["grad_out"].to(dev)
    for cdt, tol in ((torch.float32, 2e-4), (torch.bfloat16, 3e-2)):
        res = {}
        for fuse in (True, False):
            blk = build_block(case, data, dev, compute_dtype=cdt, fuse=fuse).train()
            x = data["x"].to(dev).requires_grad_(True)
            out = blk(x, task_id=1)
            loss = blk.mlp.gate[1].get_loss(clear=False)
            torch.autograd.backward([out, loss], [g, torch.tensor(0.01, device=dev)])
            res[fuse] = (out.detach(), x.grad, blk.norm2.weight.grad, blk.norm2.bias.grad, blk.mlp.gate[1].w_gate.grad,
                         blk.mlp.experts.htoh4.weight.grad, blk.mlp.last_counts, float(loss))
        a, b = res[True], res[False]
        # routing may differ on a handful of near-tie tokens (uncertified random data)
        assert int((a[6] - b[6]).abs().sum()) <= 8
        rows = (a[0] - b[0]).abs().amax(-1)
        assert float((rows > tol * float(b[0].abs().max())).float().mean()) <= 1e-3
        assert a[7] == pytest.approx(b[7], rel=1e-3)
        for i in (2, 3, 4, 5):
            assert nerr(a[i], b[i]) <= max(tol, 2e-3), i
        bad = ((a[1] - b[1]).abs().amax(-1) > tol * float(b[1].abs().max())).float().mean()
        assert float(bad) <= 1e-3

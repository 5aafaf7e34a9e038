"""GPU parity of every C-ABI op against the CPU oracle (oracle/moe_oracle.py) on the
golden-fixture inputs.  Integer / index work is bit-exact; fp32 within stated tolerance."""
import pytest
import torch

from helpers import all_fixtures, load_fixture
from oracle import moe_oracle as O

pytestmark = pytest.mark.gpu
from m3vit_b200._lib import PAD_ROWS as PAD  # noqa: E402

SMALL = [f for f in all_fixtures() if f.startswith("S")]
ALL = all_fixtures()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a B200"
    return torch.device("cuda:0")


def gate_input(case, data):
    g = data["x"].reshape(-1, case.d_model)
    return g


@pytest.mark.parametrize("fname", ALL)
def test_gate_fwd_matches_reference_fixture(fname, dev):
    """idx / counts bit-exact vs the reference's own run; probabilities to 2e-6."""
    from m3vit_b200 import ops
    fx, case, data = load_fixture(fname)
    x = gate_input(case, data).to(dev)
    tf = data["task_feat"].to(dev) if data["task_feat"] is not None else None
    for (variant, task, mode), rec in fx["tasks"].items():
        if variant != "ckpt":
            continue
        wg = data["w_gate"][task if task is not None else 0].to(dev)
        g = ops.gate_fwd(x, wg, case.top_k, tf, want_gates=True)
        plan = ops.route_plan(g.idx, case.num_expert, PAD, g.imp_partial, g.load_partial)
        assert torch.equal(g.idx.cpu().to(torch.int16), rec["idx"]), "routing indices differ"
        assert torch.equal(plan.counts.cpu(), rec["counts"]), "expert counts differ"
        torch.testing.assert_close(g.score.cpu(), rec["score"], rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(g.top_vals.cpu(), rec["top_logits"], rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(g.clean_logits.cpu()[::fx["row_stride"]], rec["clean_logits"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(plan.importance.cpu(), rec["importance"], rtol=1e-5, atol=1e-5)
        assert torch.equal(plan.load.cpu(), rec["load"].float())
        # dense gates = scatter(idx, score)
        dense = torch.zeros(case.T, case.num_expert, device=dev).scatter(1, g.idx, g.score)
        assert torch.equal(g.gates, dense)
        assert torch.equal(g.idx_full[:, :case.top_k].long(), g.idx)


@pytest.mark.parametrize("fname", SMALL)
def test_gate_fwd_bf16_input_and_noise(fname, dev):
    from m3vit_b200 import ops
    fx, case, data = load_fixture(fname)
    if data["task_feat"] is not None:
        pytest.skip("covered by the fp32 test")
    xb = gate_input(case, data).to(dev).bfloat16()
    wg = data["w_gate"][0].to(dev)
    gen = torch.Generator(device="cpu").manual_seed(7)
    noise = torch.randn(case.T, case.num_expert, generator=gen)
    std = 1.0 / case.num_expert
    g = ops.gate_fwd(xb, wg, case.top_k, None, noise.to(dev), std)
    ref = O.gate_forward(xb.float().cpu(), data["w_gate"][0], case.top_k, noise_std=1.0, training=True, noise=noise)
    torch.testing.assert_close(g.clean_logits.cpu(), ref["clean_logits"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(g.noisy_logits.cpu(), ref["noisy_logits"], rtol=1e-5, atol=1e-5)
    # indices may legitimately differ only where the oracle's own top-(K+1) gap is below fp32 noise
    p64 = torch.softmax(ref["noisy_logits"].double(), 1)
    gap = O.min_topk_gap(p64, min(case.top_k + 1, case.num_expert))
    ok = gap > 1e-5
    assert torch.equal(g.idx.cpu()[ok], ref["idx"][ok])
    torch.testing.assert_close(g.score.cpu()[ok], ref["score"][ok], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("pad", [1, 128, 256])
@pytest.mark.parametrize("T,K,E", [(1, 1, 4), (3, 2, 8), (513, 4, 16), (2402, 4, 16), (5000, 2, 64), (777, 1, 128)])
def test_route_plan_bit_exact(T, K, E, pad, dev):
    from m3vit_b200 import ops
    gen = torch.Generator().manual_seed(T + K + E)
    idx = torch.randint(0, E, (T, K), generator=gen)
    if E >= 8:
        idx[idx == 5] = 6                          # an EMPTY expert
    c, o, p, _ = O.route_plan(idx, E, pad)
    plan = ops.route_plan(idx.to(dev), E, pad)
    assert torch.equal(plan.counts.cpu(), c)
    assert torch.equal(plan.offsets.cpu(), o)
    assert torch.equal(plan.pos.cpu(), p)           # stable order: bit-exact positions
    if pad == 1:
        assert plan.tile_expert is None          # no tile map for the pad-1 (transport) plan
        return
    ntile = int(o[-1]) // pad
    te = plan.tile_expert.cpu()[:ntile]
    rows = torch.arange(ntile) * pad
    for i in range(ntile):
        e = int(te[i])
        assert int(o[e]) <= int(rows[i]) < int(o[e + 1])


def test_route_plan_dropped_slots(dev):
    """fmoe semantics: idx -1 means the slot is dropped."""
    from m3vit_b200 import ops
    idx = torch.tensor([[0, -1], [1, 0], [-1, -1], [1, 1]])
    plan = ops.route_plan(idx.to(dev), 4, 1)
    assert plan.counts.cpu().tolist() == [2, 3, 0, 0]
    assert plan.pos.cpu().tolist() == [0, -1, 2, 1, -1, -1, 3, 4]


@pytest.mark.parametrize("D", [64, 128, 384, 768])
@pytest.mark.parametrize("qdtype", [torch.float32, torch.bfloat16])
def test_dispatch_combine_roundtrip(D, qdtype, dev):
    from m3vit_b200 import ops
    T, K, E = 333, 4, 16
    gen = torch.Generator().manual_seed(D)
    x = torch.randn(T, D, generator=gen)
    idx = torch.stack([torch.randperm(E, generator=gen)[:K] for _ in range(T)])
    score = torch.rand(T, K, generator=gen)
    c, o, p, row_slot = O.route_plan(idx, E, PAD)
    plan = ops.route_plan(idx.to(dev), E, PAD)
    xq = ops.dispatch_fwd(x.to(dev), plan, K, out_dtype=qdtype)
    ref_xq = O.dispatch(x, p, K, int(o[-1]))
    n = int(o[-1])
    want = ref_xq.to(qdtype)
    assert torch.equal(xq[:n].cpu(), want), "dispatch must be an exact row copy (+cast), pads zero"
    # combine of the dispatched rows: out[t] = sum_k s[t,k] * x[t]
    out = ops.combine_fwd(xq, plan, score.to(dev), out_dtype=torch.float32)
    ref = O.combine(want.float(), p, score)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    # backward movers
    g = torch.randn(T, D, generator=gen)
    dyq, dscore = ops.combine_bwd(g.to(dev), xq, plan, score.to(dev))
    ref_dscore = (g.view(T, 1, D) * want.float()[p.long()].view(T, K, D)).sum(-1)
    torch.testing.assert_close(dscore.cpu(), ref_dscore, rtol=1e-4, atol=1e-3 if qdtype == torch.bfloat16 else 1e-4)
    ref_dyq = torch.zeros(n, D)
    ref_dyq[p.long()] = (score.view(T, K, 1) * g.view(T, 1, D)).reshape(T * K, D)
    torch.testing.assert_close(dyq[:n].float().cpu(), ref_dyq.to(qdtype).float(), rtol=1e-2 if qdtype == torch.bfloat16 else 1e-6, atol=1e-6)
    dx = ops.dispatch_bwd(dyq, plan, T, K, out_dtype=torch.float32)
    ref_dx = dyq[:n].float().cpu()[p.long()].view(T, K, D).sum(1)
    torch.testing.assert_close(dx.cpu(), ref_dx, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("fname", SMALL)
def test_ffn_f32_forward_backward(fname, dev):
    from m3vit_b200 import ops
    fx, case, data = load_fixture(fname)
    E, D, H, K = case.num_expert, case.d_model, case.d_hidden, case.top_k
    gd = O.gate_forward(O.torch.cat((data["x"].reshape(-1, D), data["task_feat"].view(1, -1).expand(case.T, -1)), 1)
                        if data["task_feat"] is not None else data["x"].reshape(-1, D), data["w_gate"][0], K)
    c, o, p, _ = O.route_plan(gd["idx"], E, PAD)
    n = int(o[-1])
    xq_ref = O.dispatch(data["x"].reshape(-1, D), p, K, n).requires_grad_(True)
    w = {k: data[k].clone().requires_grad_(True) for k in ("w1", "b1", "w2", "b2")}
    yq_ref = O.expert_ffn(xq_ref, c, o, w["w1"], w["b1"], w["w2"], w["b2"])
    gen = torch.Generator().manual_seed(3)
    dy = torch.randn(n, D, generator=gen)
    valid = torch.zeros(n, dtype=torch.bool)
    valid[p.long()] = True
    dy[~valid] = 0
    yq_ref.backward(dy)

    plan = ops.route_plan(gd["idx"].to(dev), E, PAD)
    xq = ops.dispatch_fwd(data["x"].reshape(-1, D).to(dev), plan, K)
    yq, hpre = ops.ffn_fwd(xq, plan, data["w1"].to(dev), data["b1"].to(dev), data["w2"].to(dev), data["b2"].to(dev))
    torch.testing.assert_close(yq[:n].cpu()[valid], yq_ref.detach()[valid], rtol=1e-4, atol=1e-5)
    dyq = torch.zeros(plan.cap_rows, D, device=dev)
    dyq[:n] = dy.to(dev)
    dxq, dw1, db1, dw2, db2 = ops.ffn_bwd(xq, hpre, dyq, plan, data["w1"].to(dev), data["w2"].to(dev))
    torch.testing.assert_close(dxq[:n].cpu()[valid], xq_ref.grad[valid], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dw1.cpu(), w["w1"].grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dw2.cpu(), w["w2"].grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(db1.cpu(), w["b1"].grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(db2.cpu(), w["b2"].grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("fname", SMALL)
def test_gate_bwd_matches_autograd_oracle(fname, dev):
    """softmax-Jacobian router backward vs torch autograd on the oracle gate, with
    gradients flowing into score, top_vals, gates, clean and noisy logits at once."""
    from m3vit_b200 import ops
    fx, case, data = load_fixture(fname)
    D, E, K = case.d_model, case.num_expert, case.top_k
    K1 = min(K + 1, E)
    x = data["x"].reshape(-1, D).clone().requires_grad_(True)
    wg = data["w_gate"][0].clone().requires_grad_(True)
    tf = data["task_feat"].clone().requires_grad_(True) if data["task_feat"] is not None else None
    gin = x if tf is None else torch.cat((x, tf.view(1, -1).repeat(case.T, 1)), 1)
    gd = O.gate_forward(gin, wg, K)
    gen = torch.Generator().manual_seed(11)
    r = {k: torch.randn(s, generator=gen) for k, s in
         dict(score=(case.T, K), top=(case.T, K1), gates=(case.T, E), clean=(case.T, E), imp=(E,)).items()}
    L = ((gd["score"] * r["score"]).sum() + (gd["top_logits"] * r["top"]).sum() + (gd["gates"] * r["gates"]).sum()
         + (gd["clean_logits"] * r["clean"]).sum() + (gd["gates"].sum(0) * r["imp"]).sum())
    L.backward()
    g = ops.gate_fwd(data["x"].reshape(-1, D).to(dev), data["w_gate"][0].to(dev), K,
                     data["task_feat"].to(dev) if tf is not None else None)
    assert torch.equal(g.idx.cpu(), gd["idx"])
    dz, dw, dtf, dxg = ops.gate_bwd(data["x"].reshape(-1, D).to(dev), data["w_gate"][0].to(dev), g.noisy_logits,
                                    g.idx_full, K, data["task_feat"].to(dev) if tf is not None else None,
                                    dscore=r["score"].to(dev), dtop_vals=r["top"].to(dev), dgates=r["gates"].to(dev),
                                    dimportance=r["imp"].to(dev), dclean=r["clean"].to(dev), want_dx_gate=True)
    torch.testing.assert_close(dw.cpu(), wg.grad, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(dxg.cpu(), x.grad, rtol=2e-4, atol=2e-5)
    if tf is not None:
        torch.testing.assert_close(dtf.cpu(), tf.grad, rtol=2e-4, atol=2e-4)


def test_cast_weights_bf16(dev):
    from m3vit_b200 import ops
    w = torch.randn(3, 70, 45, device=dev)
    o, ot = ops.cast_weights_bf16(w, True, True)
    assert torch.equal(o, w.bfloat16())
    assert torch.equal(ot, w.bfloat16().transpose(1, 2).contiguous())


def test_shape_errors_raise_value_error(dev):
    from m3vit_b200 import ops
    x = torch.randn(8, 48, device=dev)            # D % 32 != 0
    with pytest.raises(ValueError):
        ops.gate_fwd(x, torch.randn(48, 16, device=dev), 4)
    with pytest.raises(ValueError):
        ops.gate_fwd(torch.randn(8, 64, device=dev), torch.randn(64, 12, device=dev), 4)   # E=12 unsupported


@pytest.mark.parametrize("T,K,E,D,out_dtype", [
    (1000, 4, 16, 384, torch.float32),     # bench shape, ragged last 16-token tile
    (37, 2, 32, 128, torch.float32),       # two k-steps, tiny ragged T
    (513, 1, 64, 768, torch.float32),      # four k-steps, ViT-B width
    (300, 5, 16, 192, torch.bfloat16),     # K > 4 (two gather batches), bf16 dx
])
def test_dispatch_bwd_router_term_on_tensor_cores(T, K, E, D, out_dtype, dev):
    """dispatch_bwd with bf16 queues adds dz @ w_gate[:D]^T on mma.sync (bf16 operands, fp32 accumulate): against an fp64
    reference fed the same bf16-rounded dz / w_gate (tolerance: fp32 accumulation order only), and against the exact fp32
    SIMT kernel (tolerance: the bf16 rounding of dz and w_gate, 2^-8 relative per product).  Dropped slots stay out."""
    from m3vit_b200 import ops, _lib
    gen = torch.Generator().manual_seed(T + E)
    idx = torch.stack([torch.randperm(E, generator=gen)[:K] for _ in range(T)])
    idx[::7, 0] = -1                                             # dropped slots (capacity overflow / masked tokens)
    plan = ops.route_plan(idx.to(dev), E)
    n = int(plan.offsets[-1])
    dxq = torch.randn(plan.cap_rows, D, generator=gen).bfloat16()
    dz = torch.randn(T, E, generator=gen) * 0.3
    wg = torch.randn(D + 3, E, generator=gen) / D ** 0.5         # 3 task-feature rows that must be ignored
    pos = plan.pos.cpu().long().view(T, K)
    rows = dxq.double()[pos.clamp_min(0)] * (pos >= 0).unsqueeze(-1)
    ref_gather = rows.sum(1)
    ref = ref_gather + dz.bfloat16().double() @ wg[:D].bfloat16().double().t()
    got = ops.dispatch_bwd(dxq.to(dev), plan, T, K, out_dtype=out_dtype, dz=dz.to(dev), w_gate=wg.to(dev))
    tol = 1e-5 if out_dtype == torch.float32 else 1e-2
    assert float((got.double().cpu() - ref).abs().max() / ref.abs().max()) < tol
    exact = ref_gather + dz.double() @ wg[:D].double().t()
    assert float((got.double().cpu() - exact).norm() / exact.norm()) < (4e-3 if out_dtype == torch.float32 else 8e-3)
    if D * E * 4 > 100 * 1024:
        return                                                   # the SIMT kernel stages fp32 w_gate^T in smem: shape unsupported
    lib = _lib.load()
    old = lib.m3_set_knob(2, 9)                                  # M3_KNOB_MOVER_VARIANT = 9: exact fp32 SIMT router term
    try:
        simt = ops.dispatch_bwd(dxq.to(dev), plan, T, K, out_dtype=torch.float32, dz=dz.to(dev), w_gate=wg.to(dev))
    finally:
        lib.m3_set_knob(2, old)
    assert float((simt.double().cpu() - exact).abs().max() / exact.abs().max()) < 1e-5

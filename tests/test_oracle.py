"""CPU: the oracle restatement (oracle/moe_oracle.py) against the golden
fixtures produced by the reference's own files (oracle/make_golden.py)."""
import pytest
import torch

from helpers import all_fixtures, block_fixtures, load_fixture
from oracle import moe_oracle as O


@pytest.mark.parametrize("fname", all_fixtures())
def test_restatement_matches_reference_fixture(fname):
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    for (variant, task, mode), rec in fx["tasks"].items():
        if variant != "origin":
            continue
        wg = data["w_gate"][task if task is not None else 0]
        x = data["x"].clone().requires_grad_(mode == "train")
        w = {k: data[k].clone().requires_grad_(mode == "train") for k in ("w1", "b1", "w2", "b2")}
        wg = wg.clone().requires_grad_(mode == "train")
        out, gd = O.layer_forward(x, wg, w["w1"], w["b1"], w["w2"], w["b2"], case.top_k,
                                  task_specific_feature=data["task_feat"], noise_std=0.0,
                                  training=(mode == "train"))
        # integer work: bit exact
        assert torch.equal(gd["idx"].to(torch.int16), rec["idx"])
        assert torch.equal(gd["counts"], rec["counts"])
        # same ops in the same order on the same machine: bit exact as well
        assert torch.equal(gd["score"], rec["score"])
        assert torch.equal(out.reshape(case.T, -1)[::stride], rec["out"])
        if mode == "train":
            assert float(gd["loss"]) == pytest.approx(rec["loss"], rel=1e-6)
            (out * data["grad_out"]).sum().backward()
            torch.testing.assert_close(x.grad.reshape(case.T, -1)[::stride], rec["dx"], rtol=1e-5, atol=1e-6)
            gname = f"gate.{task}.w_gate" if task is not None else "gate.w_gate"
            torch.testing.assert_close(wg.grad, rec["grads"][gname], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(w["b2"].grad, rec["grads"]["experts.h4toh.bias"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(w["b1"].grad, rec["grads"]["experts.htoh4.bias"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("fname", all_fixtures())
def test_ckpt_variant_summaries(fname):
    """The ckpt variant's extra returns (ckpt/custom_moe_layer.py:181) and the
    Block-level cv-loss (ckpt/vision_transformer_moe.py:452-459,538-542)."""
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    for (variant, task, mode), rec in fx["tasks"].items():
        if variant != "ckpt":
            continue
        wg = data["w_gate"][task if task is not None else 0]
        _, gd = O.layer_forward(data["x"], wg, data["w1"], data["b1"], data["w2"], data["b2"], case.top_k,
                                task_specific_feature=data["task_feat"], training=True)
        assert torch.equal(gd["clean_logits"][::stride], rec["clean_logits"])
        assert torch.equal(gd["top_logits"], rec["top_logits"])
        loss, imp, load = O.cv_loss_from_summaries(gd["gates"], gd["clean_logits"], gd["noisy_logits"],
                                                   gd["noise_stddev"], gd["top_logits"], case.top_k)
        assert torch.equal(imp, rec["importance"])
        assert torch.equal(load, rec["load"])
        assert float(loss) == pytest.approx(rec["loss"], rel=1e-6)


@pytest.mark.parametrize("fname", block_fixtures())
def test_block_restatement_matches_reference_block_fixture(fname):
    """Block-level path: O.block_mlp_forward against the reference Block.forward
    (origin/vision_transformer_moe.py:274-283, attention stubbed to zero) run by oracle/make_golden.py."""
    fx, case, data = load_fixture(fname)
    stride = fx["row_stride"]
    for (task, mode), rec in fx["tasks"].items():
        train = mode == "train"
        x = data["x"].clone().requires_grad_(train)
        lw, lb = data["ln_w"].clone().requires_grad_(train), data["ln_b"].clone().requires_grad_(train)
        wg = data["w_gate"][task if task is not None else 0].clone().requires_grad_(train)
        out, gd = O.block_mlp_forward(x, lw, lb, data["ln_eps"], wg, data["w1"], data["b1"], data["w2"], data["b2"],
                                      case.top_k, task_specific_feature=data["task_feat"], training=train)
        assert torch.equal(gd["idx"].to(torch.int16), rec["idx"])
        assert torch.equal(gd["counts"], rec["counts"])
        assert torch.equal(gd["score"], rec["score"])
        assert torch.equal(out.reshape(case.T, -1)[::stride], rec["out"])
        if train:
            assert float(gd["loss"]) == pytest.approx(rec["loss"], rel=1e-6)
            (out * data["grad_out"]).sum().backward()
            torch.testing.assert_close(x.grad.reshape(case.T, -1)[::stride], rec["dx"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(lw.grad, rec["grads"]["norm2.weight"], rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(lb.grad, rec["grads"]["norm2.bias"], rtol=1e-4, atol=1e-5)
            gname = f"mlp.gate.{task}.w_gate" if task is not None else "mlp.gate.w_gate"
            torch.testing.assert_close(wg.grad, rec["grads"][gname], rtol=1e-5, atol=1e-6)


def test_fixture_gaps_are_certified():
    """Every fixture's routing must be well defined: fp64 adjacent gap among the
    top-(K+1) probabilities >= 1e-5 for every token (hard part 1)."""
    for fname in all_fixtures():
        fx, case, data = load_fixture(fname)
        for wg in data["w_gate"]:
            g = data["x"].reshape(-1, case.d_model).double()
            if data["task_feat"] is not None:
                g = torch.cat((g, data["task_feat"].double().view(1, -1).expand(g.shape[0], -1)), 1)
            p = torch.softmax(g @ wg.double(), 1)
            gap = O.min_topk_gap(p, min(case.top_k + 1, case.num_expert))
            if not case.name.startswith("S6"):
                assert float(gap.min()) >= 1e-5, fname


def test_route_plan_padding_and_stability():
    torch.manual_seed(0)
    idx = torch.randint(0, 8, (50, 2))
    idx[idx == 5] = 4                      # expert 5 is empty
    for pad in (1, 16, 128):
        counts, offsets, pos, row_slot = O.route_plan(idx, 8, pad)
        assert int(counts.sum()) == 100 and int(counts[5]) == 0
        assert all(int(o) % pad == 0 for o in offsets)
        flat = idx.reshape(-1)
        for e in range(8):
            slots = (flat == e).nonzero().flatten()
            rows = pos[slots].long()
            assert torch.equal(rows, torch.arange(len(slots)) + int(offsets[e]))   # stable, contiguous
        valid = row_slot >= 0
        assert int(valid.sum()) == 100
        assert torch.equal(pos[row_slot[valid].long()].long(), valid.nonzero().flatten())

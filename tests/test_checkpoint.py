"""CPU: expert-parallel checkpoint wire format."""
import pytest
import torch
import torch.nn as nn

import m3vit_b200 as M
from m3vit_b200 import checkpoint as C


def layer(num_expert, world=1):
    return M.FMoETransformerMLP(num_expert=num_expert, d_model=64, d_gate=64, d_hidden=128,
                                activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE,
                                world_size=world, top_k=2, vmoe_noisy_std=0)


def prefixed(sd, prefix="backbone.blocks.1.mlp."):
    return {prefix + k: v for k, v in sd.items()}


def test_shard_merge_roundtrip(tmp_path):
    torch.manual_seed(0)
    full = prefixed(layer(16).state_dict())
    W, E_loc = 4, 4
    for r in range(W):
        sd = C.shard_expert_state_dict(full, r, E_loc)
        assert tuple(sd["backbone.blocks.1.mlp.experts.htoh4.weight"].shape) == (E_loc, 128, 64)
        assert torch.equal(sd["backbone.blocks.1.mlp.experts.h4toh.bias"],
                           full["backbone.blocks.1.mlp.experts.h4toh.bias"][r * E_loc:(r + 1) * E_loc])
        assert torch.equal(sd["backbone.blocks.1.mlp.gate.w_gate"], full["backbone.blocks.1.mlp.gate.w_gate"])
        C.save_ep_shard({"state_dict": sd, "epoch": 3}, str(tmp_path), r)
    # ranks != 0 wrote experts only
    s1 = torch.load(tmp_path / "1.pth", weights_only=False)["state_dict"]
    assert all(C.is_expert_key(k) for k in s1) and len(s1) == 4
    merged = C.load_ep_dir(str(tmp_path), W)
    assert merged["meta"]["expert_format"] == "global" and merged["epoch"] == 3
    assert set(merged["state_dict"]) == set(full)
    for k in full:
        assert torch.equal(merged["state_dict"][k], full[k]), k


def test_sharded_state_loads_into_ep_layer():
    torch.manual_seed(1)
    full = layer(16)
    shard = layer(4, world=4)          # world_size=4: gate keeps 16 columns, experts 4 per rank
    pre = "blocks.1.mlp."                # expert keys are recognised by their "...mlp.experts.*" name
    sd = C.shard_expert_state_dict(prefixed(full.state_dict(), pre), 2, 4)
    shard.load_state_dict({k[len(pre):]: v for k, v in sd.items()})
    assert torch.equal(shard.experts.htoh4.weight, full.experts.htoh4.weight[8:12])
    assert torch.equal(shard.gate.w_gate, full.gate.w_gate)


def test_expert_format_rules():
    sd = prefixed(layer(4).state_dict())
    assert C.expert_format({}, sd, local_experts=4, world_size=1) == "global"
    with pytest.raises(ValueError):
        C.expert_format({}, sd, local_experts=4, world_size=4)              # dim0 = 4 looks rank-local
    with pytest.raises(ValueError):
        C.expert_format({"meta": {"expert_format": "local"}}, sd, 4, 4)
    g = prefixed(layer(16).state_dict())
    assert C.expert_format({"meta": {"expert_format": "global"}}, g, 4, 4) == "global"
    assert C.expert_format({}, {"module.encoder." + k: v for k, v in g.items()}, 4, 4) == "global"
    with pytest.raises(ValueError):
        C.expert_format({"meta": {"expert_format": "global"}}, g, 2, 4)


# ------------------------------------------------------------------------------------------------------------------
# Pinned to the reference: tests/golden/ckpt_reference.pt holds what the reference's OWN functions returned / raised
# (oracle/make_ckpt_golden.py executes utils/moe_utils.py and pretrain/utils/moe_checkpoint.py verbatim).
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_reference.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def same_state(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_filter_and_shard_match_reference(gold):
    g = gold["filter_state"]
    same_state(C.filter_expert_state(g["in"]), g["out"])                   # utils/moe_utils.py:128-134
    for c in gold["read_specific_group_experts"]:                           # utils/moe_utils.py:191-198
        same_state(C.shard_expert_state_dict(c["in"], c["rank"], c["num"]), c["out"])


def test_expert_format_matches_reference_validator(gold):
    """validate_single_file_moe_checkpoint_or_raise (utils/moe_utils.py:34-106): accepts exactly what ours accepts"""
    for c in gold["validate"]:
        kind = c["res"][0]
        if kind == "ok":
            assert C.expert_format(c["ckpt"], c["state"], c["local"], c["world"]) == "global"
        else:
            assert c["res"][1] == "ValueError"
            with pytest.raises(ValueError):
                C.expert_format(c["ckpt"], c["state"], c["local"], c["world"])


def test_pretrain_helpers_match_reference(gold):
    g = gold["to_backbone"]                                                  # pretrain/utils/moe_checkpoint.py:23-47
    st, dropped = C.to_backbone_state_dict(g["in"])
    same_state(st, g["out"])
    assert dropped == g["dropped"]
    for c in gold["build_meta"]:                                             # :82-113
        assert C.build_meta(c["state"], "unit", **c["kw"]) == c["out"]
    for c in gold["infer"]:                                                  # :137-180
        assert C.infer_expert_format(c["ckpt"], c["state"], c["g"], c["w"]) == c["out"], c


def test_shard_directory_written_by_the_reference(gold, tmp_path):
    """{rank}.pth files produced by the reference's save_moe_model_to_dir on two gloo ranks: ours writes the same files,
    and both mergers (ours, the reference's merge_moe_sharded_directory) give back the global state."""
    g = gold["shard_dir"]
    assert sorted(g["files"]) == ["0.pth", "1.pth"] and g["n"] == 2
    full = g["full"]
    for r in range(2):
        ours_dir = tmp_path / "ours"
        local = C.shard_expert_state_dict(dict(full), r, 4)
        C.save_ep_shard({"state_dict": local, "epoch": 7, "args": {"world_size": 2, "moe_experts": 8}}, str(ours_dir), r)
        mine = torch.load(ours_dir / f"{r}.pth", weights_only=False)
        ref = g["files"][f"{r}.pth"]
        assert set(mine) == set(ref) and mine["epoch"] == ref["epoch"] and mine["args"] == ref["args"]
        same_state(dict(mine["state_dict"]), dict(ref["state_dict"]))
    # merge the REFERENCE's files with our merger
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    for n, ck in g["files"].items():
        torch.save(ck, ref_dir / n)
    base, merged, n = C.merge_shard_dir(str(ref_dir))
    assert n == 2 and base["epoch"] == g["base_epoch"]
    same_state(dict(merged), g["merged"])
    same_state(dict(merged), full)
    same_state(dict(C.load_ep_dir(str(ref_dir), 2)["state_dict"]), full)
    for name, res in gold["merge_errors"]:
        d = tmp_path / name
        d.mkdir()
        if name == "no_rank0":
            torch.save({"state_dict": {}}, d / "1.pth")
        assert res == ("raise", "ValueError")
        with pytest.raises(ValueError):
            C.merge_shard_dir(str(d))

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

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


# ----------------------------------------------------------------------------- dense MLP -> experts (upcycling)
def _upcycle_golden():
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "upcycle_reference.pt")
    return torch.load(p, weights_only=False)


@pytest.mark.parametrize("name", ["ratio4_copy", "ratio1_split_repeat", "ratio1_split_scaled", "ratio1_split_truncate",
                                  "ratio2_split_g2_warm_refused", "ratio_minus1_uses_mlp_ratio"])
def test_upcycling_matches_reference_function(name):
    """inject_experts_from_dense_mlp against what the reference's own `_inject_moe_expert_from_deit_mlp`
    (utils/helpers.py:481-713, executed verbatim by oracle/make_upcycle_golden.py) returned or raised: bit-exact."""
    rec = _upcycle_golden()["cases"][name]
    c = rec["case"]
    sd = {k: v.clone() for k, v in rec["input"].items()}
    blocks = {i: dict(local_experts=c["E_local"], world_size=c["world"], total_experts=c["total"], top_k=c["top_k"],
                      expert_hidden=c["He"]) for i in (0, 2)}                 # block 1 is dense
    kw = dict(moe_mlp_ratio=c["ratio"], mode=c["mode"], weight_scaling=bool(c["cfg"].get("use_weight_scaling", False)))
    if "raises" in rec:
        with pytest.raises(ValueError if rec["raises"] == "ValueError" else AssertionError):
            C.inject_experts_from_dense_mlp(sd, blocks, **kw)
        return
    out = C.inject_experts_from_dense_mlp(sd, blocks, **kw)
    want = rec["output"]
    assert set(out) == set(want)
    for k in want:
        assert out[k].shape == want[k].shape and torch.equal(out[k], want[k]), k
    assert not any(k.startswith("blocks.1.mlp.experts") for k in out)


def test_upcycled_experts_reproduce_the_dense_mlp():
    """Function-level property (size independent): with every expert a full copy (copy mode) any routing whose scores sum
    to 1 reproduces the dense MLP; in split mode the G experts of a group together are the dense MLP minus (G-1) fc2 biases."""
    torch.manual_seed(0)
    D, Hd, G = 8, 32, 4
    fc1_w, fc1_b, fc2_w, fc2_b = torch.randn(Hd, D), torch.randn(Hd), torch.randn(D, Hd), torch.randn(D)
    x = torch.randn(5, D)
    dense = torch.nn.functional.gelu(x @ fc1_w.t() + fc1_b) @ fc2_w.t() + fc2_b
    w1, b1, w2, b2 = C.upcycle_dense_mlp(fc1_w, fc1_b, fc2_w, fc2_b, local_experts=3)
    for e in range(3):
        y = torch.nn.functional.gelu(x @ w1[e].t() + b1[e]) @ w2[e].t() + b2[e]
        assert torch.allclose(y, dense, atol=1e-5)
    w1, b1, w2, b2 = C.upcycle_dense_mlp(fc1_w, fc1_b, fc2_w, fc2_b, local_experts=8, total_experts=8, expert_hidden=Hd // G)
    assert w1.shape == (8, Hd // G, D) and w2.shape == (8, D, Hd // G)
    group = sum(torch.nn.functional.gelu(x @ w1[e].t() + b1[e]) @ w2[e].t() + b2[e] for e in range(G))
    assert torch.allclose(group, dense + (G - 1) * fc2_b, atol=1e-4)
    assert torch.equal(w1[:G], w1[G:]) and torch.equal(b2[0], fc2_b)


# ----------------------------------------------------------------------------- router keys (cvt_state_dict)
@pytest.mark.parametrize("name", ["shared_gate_untouched", "one_hot_pads_num_tasks_rows", "one_hot_pads_task_dim_rows",
                                  "one_hot_regu_untouched", "multi_gate_2", "multi_gate_4", "multi_gate_5",
                                  "multi_gate_3_gets_two"])
def test_gate_key_conversion_matches_reference_function(name):
    """convert_gate_keys against the state dict the reference's own `cvt_state_dict` (utils/common_config.py:31-100, its
    source segment executed verbatim by oracle/make_gatekeys_golden.py) handed to load_state_dict: same keys, same bits."""
    import os
    rec = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gatekeys_reference.pt"),
                     weights_only=False)["cases"][name]
    c = rec["case"]
    sd = {k: v.clone() for k, v in rec["input"].items()}
    out = C.convert_gate_keys(sd, multi_gate=c["multi_gate"], num_tasks=c["num_tasks"], task_one_hot=c["task_one_hot"],
                              gate_task_specific_dim=c["gtsd"], regu_experts_fromtask=c["regu"])
    want = rec["output"]
    assert set(out) == set(want)
    for k in want:
        assert out[k].shape == want[k].shape and torch.equal(out[k], want[k]), k


def test_gate_key_conversion_loads_into_the_layer():
    """The converted keys are exactly what the drop-in layer's state dict expects (multi-gate: one gate per task;
    task-conditioned shared router: D + D_t rows), and `replicate_all_tasks` covers task counts the reference forgets."""
    torch.manual_seed(0)
    shared = {"gate.w_gate": torch.randn(64, 8)}
    mg = M.FMoETransformerMLP(num_expert=8, d_model=64, d_gate=64 + 3, d_hidden=64, activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)),
                              gate=M.NoisyGate_VMoE, top_k=2, vmoe_noisy_std=0, multi_gate=True)
    sd = C.convert_gate_keys({"mlp." + k: v.clone() for k, v in shared.items()}, multi_gate=True, num_tasks=3,
                             replicate_all_tasks=True)
    sd = {k[len("mlp."):]: v for k, v in sd.items()}
    missing, unexpected = mg.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if "gate" in k]
    assert all(torch.equal(g.w_gate, shared["gate.w_gate"]) for g in mg.gate)
    tc = M.FMoETransformerMLP(num_expert=8, d_model=64, d_gate=64, d_hidden=64, activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)),
                              gate=M.NoisyGate_VMoE, top_k=2, vmoe_noisy_std=0, multi_gate=False, gate_task_specific_dim=6)
    sd = C.convert_gate_keys({"mlp.gate.w_gate": shared["gate.w_gate"].clone()}, multi_gate=False, num_tasks=2,
                             task_one_hot=True, gate_task_specific_dim=6)
    missing, unexpected = tc.load_state_dict({k[len("mlp."):]: v for k, v in sd.items()}, strict=False)
    assert not unexpected and not [k for k in missing if "gate" in k]
    assert tc.gate.w_gate.shape == (70, 8) and torch.equal(tc.gate.w_gate[:64], shared["gate.w_gate"])
    assert float(tc.gate.w_gate.detach()[64:].abs().max()) == 0.0


# ----------------------------------------------------------------------------- virtual-group router init
def _vgi_golden():
    import os
    return torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vgi_reference.pt"), weights_only=False)


def test_virtual_group_size_matches_reference_function():
    """auto_virtual_group_size on the 1 800-point grid the reference's `_auto_virtual_group_size` was run on."""
    for (tot, loc, world, dh, eh), want in _vgi_golden()["group_size"]:
        got = C.auto_virtual_group_size(tot, local_experts=loc, world_size=world, dense_hidden=dh, expert_hidden=eh)
        assert got == want, (tot, loc, world, dh, eh, got, want)


@pytest.mark.parametrize("name", ["split_g4_shared_gate", "split_g4_two_task_gates", "copy_g_is_local"])
def test_virtual_group_gate_init_matches_reference_function(name):
    """Same torch seed -> the same router matrices, bit for bit, as `_inject_virtual_group_init_for_gates` run verbatim;
    and the structure it is for: the columns of a group repeat across groups."""
    rec = _vgi_golden()["cases"][name]
    D, Hd, He, E_local, world, gates = rec["shape"]
    tot = E_local * world
    sd = {f"blocks.{i}.mlp.fc1.weight": torch.zeros(Hd, D) for i in (0, 2)}
    model_state = {}
    for i in (0, 2):
        model_state[f"blocks.{i}.mlp.experts.htoh4.weight"] = torch.empty(E_local, He, D)
        for t in ([None] if gates is None else range(gates)):
            model_state[f"blocks.{i}.mlp.gate.w_gate" if t is None else f"blocks.{i}.mlp.gate.{t}.w_gate"] = torch.zeros(D, tot)
    torch.manual_seed(rec["seed"])
    out = C.inject_virtual_group_gate_init(sd, model_state, {i: dict(local_experts=E_local, world_size=world) for i in (0, 2)})
    want = rec["gates"]
    assert {k for k in out if k.endswith("w_gate")} == set(want)
    for k in want:
        assert torch.equal(out[k], want[k]), k
    G = C.auto_virtual_group_size(tot, local_experts=E_local, world_size=world, dense_hidden=Hd, expert_hidden=He)
    for k in want:
        w = out[k]
        for grp in range(1, tot // G if G > 1 else 1):          # (G = 1: plain normal init, nothing repeats)
            assert torch.equal(w[:, :G], w[:, grp * G:(grp + 1) * G])

"""CPU: host-side mirror of the reference layer API (construction, state dict,
attributes, error behaviour).  No CUDA calls."""
import pytest
import torch
import torch.nn as nn

import m3vit_b200 as M
from m3vit_b200._lib import M3Error


def act():
    return nn.Sequential(nn.GELU(), nn.Dropout(0.0))


def make(**kw):
    base = dict(num_expert=16, d_model=64, d_gate=64, d_hidden=128, activation=act(), gate=M.NoisyGate_VMoE,
                top_k=4, vmoe_noisy_std=0)
    base.update(kw)
    return M.FMoETransformerMLP(**base)


def test_state_dict_contract_single_gate():
    # keys and shapes other reference code relies on
    sd = make().state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "experts.htoh4.weight": (16, 128, 64), "experts.htoh4.bias": (16, 128),
        "experts.h4toh.weight": (16, 64, 128), "experts.h4toh.bias": (16, 64),
        "gate.w_gate": (64, 16)}


def test_multi_gate_count_is_dgate_minus_dmodel():
    # origin/custom_moe_layer.py:138,150: number of task gates = d_gate - d_model
    layer = make(d_gate=64 + 5, multi_gate=True)
    assert isinstance(layer.gate, nn.ModuleList) and len(layer.gate) == 5
    assert "gate.4.w_gate" in layer.state_dict()


def test_task_conditioned_gate_width():
    layer = make(gate_task_specific_dim=16)          # origin:127-130
    assert tuple(layer.gate.w_gate.shape) == (80, 16)


def test_world_size_scales_gate_columns():
    layer = make(num_expert=4, world_size=4)         # BaseGate.tot_expert = num_expert * world_size
    assert tuple(layer.gate.w_gate.shape) == (64, 16)
    assert tuple(layer.experts.htoh4.weight.shape) == (4, 128, 64)


def test_dp_comm_tags():
    layer = make()
    assert all(p.dp_comm == "none" for p in layer.experts.parameters())
    assert all(p.dp_comm == "gate" for p in layer.gate.parameters())


def test_attributes_used_by_reference_utils():
    layer = make()
    for a in ("num_expert", "world_size", "top_k", "d_model", "gate", "experts", "gate_hook", "mask", "mask_dict"):
        assert hasattr(layer, a)
    g = layer.gate
    assert not g.has_loss and not g.has_activation and g.select_idx is None and g.tot_expert == 16


def test_bad_gate_type_raises_value_error():
    with pytest.raises(ValueError, match="No such gating type"):
        make(gate=nn.Linear)


@pytest.mark.parametrize("flag", ["regu_sem", "sem_force", "regu_subimage", "expert_prune", "regu_experts_fromtask"])
def test_research_flags_raise(flag):
    with pytest.raises(NotImplementedError):
        make(**{flag: True})


def test_cpu_tensor_is_refused_loudly():
    layer = make()
    with pytest.raises(M3Error, match="no CPU fallback"):
        layer(torch.randn(2, 3, 64))


def test_multi_gate_without_task_id_is_a_type_error():
    layer = make(d_gate=66, multi_gate=True)
    with pytest.raises(TypeError):
        layer(torch.randn(2, 3, 64))


def test_non_gelu_activation_rejected():
    with pytest.raises(NotImplementedError):
        make(activation=nn.ReLU())


def test_block_adapter_argument_mapping():
    m = M.build_moe_mlp(384, moe_mlp_ratio=1, moe_experts=16, moe_top_k=4, moe_gate_dim=386,
                        moe_gate_type="noisy_vmoe", vmoe_noisy_std=0, multi_gate=True)
    assert m.d_hidden == 384 and len(m.gate) == 2 and m.top_k == 4
    with pytest.raises(ValueError):
        M.build_moe_mlp(384, moe_gate_type="bogus")
    ck = M.build_moe_mlp(64, moe_mlp_ratio=1, moe_experts=8, moe_gate_type="noisy_vmoe", variant="ckpt")
    assert isinstance(ck, M.FMoETransformerMLPCkpt) and ck.gate.return_summaries


def test_cv_squared_matches_reference_formula():
    x = torch.tensor([1.0, 2.0, 3.0, 6.0])
    assert torch.allclose(M.cv_squared(x), x.var() / (x.mean() ** 2 + 1e-10))
    assert float(M.cv_squared(torch.tensor([5.0]))) == 0.0


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU oracle port on the host cores) prints ONE JSON line with the contract's keys,
    times exactly K steps, and needs no GPU.  Tiny batch: the workload shape is the bench's, the sample is bounded."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                        "--batch", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "moe_layer_tokens_per_s_fwd_bwd" and d["unit"] == "tokens/s"
    assert d["steps"] == 2 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["batch_per_gpu"] == 1 and d["config"]["layer_calls_per_step"] == 12


def test_model_walkers_of_the_reference_utils():
    """collect_noisy_gating_loss / collect_moe_activation / set_moe_layer_train_mode (utils/moe_utils.py:201-207,
    :226-250, :303-306) find the drop-in layer's gates and layers; activations and losses are planted by hand (no GPU)."""
    torch.manual_seed(0)
    B, N, E = 2, 5, 8
    model = nn.Module()
    model.blocks = nn.ModuleList()
    for _ in range(2):
        blk = nn.Module()
        blk.mlp = M.FMoETransformerMLP(num_expert=E, d_model=16, d_gate=16 + 2, d_hidden=16,
                                       activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE, top_k=2,
                                       vmoe_noisy_std=0, multi_gate=True)
        blk.norm = nn.LayerNorm(16)
        model.blocks.append(blk)
    model.eval()
    M.set_moe_layer_train_mode(model)
    assert all(b.mlp.training and not b.norm.training for b in model.blocks)
    probs = []
    for b in model.blocks:                                      # task gate 1 ran, task gate 0 did not
        p = torch.softmax(torch.randn(B * N, E), dim=1)
        b.mlp.gate[1].activation = p
        b.mlp.gate[1].set_loss(torch.tensor(0.25))
        probs.append(p)
    acts, names = M.collect_moe_activation(model, B, return_name=True)
    assert names == ["blocks.0.mlp.gate.1", "blocks.1.mlp.gate.1"]
    assert all(torch.allclose(a, p.reshape(B, N, E).mean(1)) for a, p in zip(acts, probs))
    assert not any(b.mlp.gate[1].has_activation for b in model.blocks)          # consumed, like the reference's getter
    for b, p in zip(model.blocks, probs):
        b.mlp.gate[1].activation = p
    assert M.collect_moe_activation(model, B, "origin")[0].shape == (B, N, E)
    for b, p in zip(model.blocks, probs):
        b.mlp.gate[1].activation = p
    assert M.collect_moe_activation(model, B, "concat")[1].shape == (B, N * E)
    for b, p in zip(model.blocks, probs):
        b.mlp.gate[1].activation = p
    with pytest.raises(ValueError):
        M.collect_moe_activation(model, B, "median")
    assert float(M.collect_noisy_gating_loss(model, 0.01)) == pytest.approx(0.005)
    assert float(M.collect_noisy_gating_loss(model, 0.01)) == 0.0               # get_loss clears


def test_committed_evidence_is_current_and_complete():
    """The committed measurement files stay usable: (1) the ncu traffic capture that `bench.py` reports as
    `roofline.traffic` was taken from the kernel sources as they are now (its sha256 matches, else the bench prints null);
    (2) the committed N = 1 bench line carries every key of the bench contract."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    key = (f"ffn_fwd_train:bf16:T{32 * bench.N_TOK}:D{bench.D_MODEL}:H{bench.D_HID}:E{bench.N_EXP}:K{bench.TOP_K}")
    traffic, src = bench.ncu_traffic(key)
    assert traffic is not None and traffic > 1e8, src
    line = json.loads(open(os.path.join(root, "profiles", "r2_bench_n1.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "stages"):
        assert k in line, k
    assert line["metric"] == bench.METRIC and line["n_gpus"] == 1 and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 12 * 32 * bench.N_TOK * bench.D_MODEL * 4 and line["e2e"]["value"] < line["value"]
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and not any(k in line["config"] for k in ("model", "seq_len"))

"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt.

Runs the REFERENCE'S OWN FILES, unmodified, on CPU:

    /root/reference/models/moe/origin/custom_moe_layer.py + origin/noisy_gate_vmoe.py
    /root/reference/models/moe/ckpt/custom_moe_layer.py   + ckpt/noisy_gate_vmoe.py

with `fmoe` / `tree` / `timm` replaced by the pure-PyTorch shims in oracle/shim
(FastMoE is an un-vendored third-party dependency, see oracle/shim/fmoe).  The
outputs are committed as small fixtures; /root/reference is not needed (and not
present) when the tests run on the GPU box.

    python oracle/make_golden.py            # needs /root/reference

Inputs and weights are NOT stored for the large cases: they are regenerated
from (case, seed) by m3vit_b200.synthetic.make_case, and the fixture keeps an
fp64 checksum of every regenerated tensor so RNG drift is detected.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from m3vit_b200.synthetic import MoECase, C1, C3S, C4S, make_block_case, make_case  # noqa: E402

torch.set_num_threads(8)

SMALL = [
    MoECase("S1_e16k4_g2", batch=2, tokens=65, d_model=64, d_hidden=64, num_expert=16, top_k=4, num_gates=2),
    MoECase("S2_e32k2_r4", batch=1, tokens=131, d_model=64, d_hidden=256, num_expert=32, top_k=2),
    MoECase("S3_e64k1", batch=3, tokens=50, d_model=128, d_hidden=128, num_expert=64, top_k=1),
    MoECase("S4_taskcond", batch=2, tokens=33, d_model=64, d_hidden=64, num_expert=16, top_k=4, d_task=16),
    MoECase("S5_e8k2_tiny", batch=1, tokens=3, d_model=64, d_hidden=64, num_expert=8, top_k=2),
    MoECase("S6_e16k4_starved", batch=1, tokens=97, d_model=64, d_hidden=128, num_expert=16, top_k=4, num_gates=3),
    MoECase("S7_e4k4_all", batch=1, tokens=40, d_model=64, d_hidden=64, num_expert=4, top_k=4),
    MoECase("S8_d128h256_g2", batch=2, tokens=77, d_model=128, d_hidden=256, num_expert=16, top_k=4, num_gates=2),
]
LARGE = [C1, C3S, C4S]


def checksum(t):
    return float(t.double().sum()), float(t.double().abs().sum())


def build_ref_layer(variant, case):
    if variant == "origin":
        from models.moe.origin.custom_moe_layer import FMoETransformerMLP
        from models.moe.origin.noisy_gate_vmoe import NoisyGate_VMoE
    else:
        from models.moe.ckpt.custom_moe_layer import FMoETransformerMLP
        from models.moe.ckpt.noisy_gate_vmoe import NoisyGate_VMoE
    multi = case.num_gates > 1
    layer = FMoETransformerMLP(
        num_expert=case.num_expert, d_model=case.d_model,
        # the reference derives the number of task gates as d_gate - d_model
        d_gate=case.d_model + (case.num_gates if multi else 0),
        d_hidden=case.d_hidden,
        activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)),
        gate=NoisyGate_VMoE, world_size=1, top_k=case.top_k, vmoe_noisy_std=0,
        gate_task_specific_dim=(case.d_task if case.d_task > 0 else -1), multi_gate=multi,
    )
    return layer


def load_weights(layer, case, data):
    with torch.no_grad():
        layer.experts.htoh4.weight.copy_(data["w1"])
        layer.experts.htoh4.bias.copy_(data["b1"])
        layer.experts.h4toh.weight.copy_(data["w2"])
        layer.experts.h4toh.bias.copy_(data["b2"])
        if case.num_gates > 1:
            for g, w in zip(layer.gate, data["w_gate"]):
                g.w_gate.copy_(w)
        else:
            layer.gate.w_gate.copy_(data["w_gate"][0])


def run_case(case, seed, full):
    data = make_case(case, seed)
    if case.name.startswith("S6"):
        # starve experts 3 and 11 under every gate: ragged / EMPTY expert queues
        for w in data["w_gate"]:
            w[:, 3] = 0.0
            w[0, 3] = -50.0
            w[:, 11] = 0.0
            w[0, 11] = -50.0
        data["x"][..., 0] = data["x"][..., 0].abs() + 0.1
    fx = dict(case=case.dict(), seed=seed, torch=torch.__version__, resampled=data["resampled"])
    fx["checksums"] = {k: checksum(v) for k, v in data.items()
                       if torch.is_tensor(v)} | {f"w_gate{i}": checksum(w) for i, w in enumerate(data["w_gate"])}
    stride = 1 if full else 32
    fx["row_stride"] = stride
    tasks = list(range(case.num_gates)) if case.num_gates > 1 else [None]
    if not full and len(tasks) > 2:
        tasks = [0, case.num_gates - 1]      # keep the large fixtures small
    T, K, E = case.T, case.top_k, case.num_expert
    fx["tasks"] = {}
    for variant in ("origin", "ckpt"):
        layer = build_ref_layer(variant, case)
        load_weights(layer, case, data)
        for task in tasks:
            for mode in ("eval", "train"):
                if variant == "ckpt" and mode == "eval":
                    continue
                layer.train(mode == "train")
                layer.zero_grad(set_to_none=True)
                x = data["x"].clone().requires_grad_(True)
                tf = data["task_feat"]
                kwargs = {}
                if tf is not None:
                    tf = tf.clone().requires_grad_(True)
                    kwargs = dict(task_id=0, task_specific_feature=tf)
                elif task is not None:
                    kwargs = dict(task_id=task)
                # capture the routing the layer used (gate_hook is part of the fmoe API,
                # origin/custom_moe_layer.py:240-241)
                cap = {}
                layer.gate_hook = lambda idx, score, _: cap.update(idx=idx.detach().clone(), score=score.detach().clone())
                ret = layer(x, **kwargs)
                rec = {}
                if variant == "origin":
                    out = ret
                    gate_mod = layer.gate[task] if task is not None else layer.gate
                    loss = gate_mod.get_loss(clear=False)
                else:
                    out, clean, noisy, nstd, top_logits, gates = ret
                    sys.path.insert(0, "/root/reference")
                    from models.moe.ckpt.vision_transformer_moe import cv_squared, _gates_to_load
                    importance = gates.sum(0)
                    load = _gates_to_load(gates)
                    loss = cv_squared(importance) + cv_squared(load)
                    rec.update(clean_logits=clean.detach()[::stride].clone(), noise_stddev=float(nstd),
                               top_logits=top_logits.detach().clone(), importance=importance.detach().clone(),
                               load=load.detach().clone())
                rec["idx"] = cap["idx"].reshape(T, K).to(torch.int16)
                rec["score"] = cap["score"].reshape(T, K).clone()
                rec["counts"] = torch.bincount(cap["idx"].reshape(-1), minlength=E).int()
                rec["out"] = out.detach().reshape(T, -1)[::stride].clone()
                if mode == "train":
                    rec["loss"] = float(loss)
                    # (1) task-loss gradient only
                    L = (out * data["grad_out"]).sum()
                    L.backward(retain_graph=True)
                    rec["dx"] = x.grad.reshape(T, -1)[::stride].clone()
                    if tf is not None:
                        rec["dtask_feat"] = tf.grad.clone()
                    g = {}
                    for name, p in layer.named_parameters():
                        if p.grad is None:
                            g[name] = None
                        elif p.grad.numel() <= 20000:
                            g[name] = p.grad.clone()
                        elif variant == "origin":
                            # expert weight grads: first/last expert, strided rows, + fp64 checksums
                            g[name] = p.grad[[0, E - 1]][:, ::max(stride, 4)].clone()
                            g[name + ".checksum"] = checksum(p.grad)
                    rec["grads"] = g
                    # (2) cv-loss gradient only
                    layer.zero_grad(set_to_none=True)
                    x.grad = None
                    loss.backward()
                    rec["cv_dx"] = x.grad.reshape(T, -1)[::stride].clone()
                    gname = f"gate.{task}.w_gate" if task is not None else "gate.w_gate"
                    rec["cv_dw_gate"] = dict(layer.named_parameters())[gname].grad.clone()
                fx["tasks"][(variant, task, mode)] = rec
    return fx


# ------------------------------------------------------------------ Block-level fixtures
BLOCK_SMALL = [
    MoECase("B1_e16k4_g2", batch=2, tokens=65, d_model=64, d_hidden=64, num_expert=16, top_k=4, num_gates=2),
    MoECase("B2_taskcond", batch=2, tokens=33, d_model=64, d_hidden=128, num_expert=16, top_k=2, d_task=16),
    MoECase("B3_e32k2_d128", batch=1, tokens=131, d_model=128, d_hidden=256, num_expert=32, top_k=2),
]
BLOCK_LARGE = [MoECase("B4_vits_nyud_b2", batch=2, tokens=1201, d_model=384, d_hidden=384, num_expert=16, top_k=4,
                       num_gates=2)]


class _ZeroAttention(nn.Module):
    """Stands in for Block.attn so that `x + drop_path(attn(norm1(x)))` leaves x unchanged and the
    reference Block.forward exercises exactly its MoE half."""

    def forward(self, x):
        return torch.zeros_like(x)


def build_ref_block(case):
    from models.moe.origin.vision_transformer_moe import Block
    from functools import partial
    multi = case.num_gates > 1
    blk = Block(dim=case.d_model, num_heads=2, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0, drop_path=0.0,
                norm_layer=partial(nn.LayerNorm, eps=1e-6), moe=True, moe_mlp_ratio=case.d_hidden / case.d_model,
                moe_experts=case.num_expert, moe_top_k=case.top_k,
                moe_gate_dim=case.d_model + (case.num_gates if multi else 0), world_size=1,
                moe_gate_type="noisy_vmoe", vmoe_noisy_std=0,
                gate_task_specific_dim=(case.d_task if case.d_task > 0 else -1), multi_gate=multi)
    blk.attn = _ZeroAttention()
    return blk


def run_block_case(case, seed, full):
    data = make_block_case(case, seed)
    fx = dict(case=case.dict(), seed=seed, torch=torch.__version__, resampled=data["resampled"], block=True)
    fx["checksums"] = {k: checksum(v) for k, v in data.items()
                       if torch.is_tensor(v)} | {f"w_gate{i}": checksum(w) for i, w in enumerate(data["w_gate"])}
    stride = 1 if full else 32
    fx["row_stride"] = stride
    T, K, E = case.T, case.top_k, case.num_expert
    tasks = list(range(case.num_gates)) if case.num_gates > 1 else [None]
    blk = build_ref_block(case)
    assert abs(blk.norm2.eps - data["ln_eps"]) < 1e-12
    load_weights(blk.mlp, case, data)
    with torch.no_grad():
        blk.norm2.weight.copy_(data["ln_w"])
        blk.norm2.bias.copy_(data["ln_b"])
    fx["tasks"] = {}
    for task in tasks:
        for mode in ("eval", "train"):
            blk.train(mode == "train")
            blk.zero_grad(set_to_none=True)
            x = data["x"].clone().requires_grad_(True)
            tf = data["task_feat"]
            kwargs = {}
            if tf is not None:
                tf = tf.clone().requires_grad_(True)
                kwargs = dict(task_id=0, task_specific_feature=tf)
            elif task is not None:
                kwargs = dict(task_id=task)
            cap = {}
            blk.mlp.gate_hook = lambda idx, score, _: cap.update(idx=idx.detach().clone(), score=score.detach().clone())
            out = blk(x, **kwargs)
            gate_mod = blk.mlp.gate[task] if task is not None else blk.mlp.gate
            rec = dict(idx=cap["idx"].reshape(T, K).to(torch.int16), score=cap["score"].reshape(T, K).clone(),
                       counts=torch.bincount(cap["idx"].reshape(-1), minlength=E).int(),
                       out=out.detach().reshape(T, -1)[::stride].clone())
            if mode == "train":
                loss = gate_mod.get_loss(clear=False)
                rec["loss"] = float(loss)
                (out * data["grad_out"]).sum().backward(retain_graph=True)
                rec["dx"] = x.grad.reshape(T, -1)[::stride].clone()
                if tf is not None:
                    rec["dtask_feat"] = tf.grad.clone()
                g = {}
                for name, p in blk.named_parameters():
                    if not (name.startswith("mlp.") or name.startswith("norm2.")):
                        continue
                    if p.grad is None:
                        g[name] = None
                    elif p.grad.numel() <= 20000:
                        g[name] = p.grad.clone()
                    else:
                        g[name] = p.grad[[0, E - 1]][:, ::max(stride, 4)].clone()
                        g[name + ".checksum"] = checksum(p.grad)
                rec["grads"] = g
                blk.zero_grad(set_to_none=True)
                x.grad = None
                loss.backward()
                rec["cv_dx"] = x.grad.reshape(T, -1)[::stride].clone()
                rec["cv_dln_w"] = blk.norm2.weight.grad.clone()
                rec["cv_dln_b"] = blk.norm2.bias.grad.clone()
            fx["tasks"][(task, mode)] = rec
    return fx


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for case, full in [(c, True) for c in BLOCK_SMALL] + [(c, False) for c in BLOCK_LARGE]:
        fx = run_block_case(case, 0, full=full)
        path = os.path.join(out_dir, f"{case.name}_s0.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", "resampled", fx["resampled"])
    if "--block-only" in sys.argv:
        return
    for case in SMALL:
        for seed in (0, 1):
            fx = run_case(case, seed, full=True)
            path = os.path.join(out_dir, f"{case.name}_s{seed}.pt")
            torch.save(fx, path)
            print(path, os.path.getsize(path) // 1024, "KiB", "resampled", fx["resampled"])
    for case in LARGE:
        fx = run_case(case, 0, full=False)
        path = os.path.join(out_dir, f"{case.name}_s0.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", "resampled", fx["resampled"])


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

Executes the REFERENCE'S OWN dense-MLP -> expert upcycling helper, unmodified, and records what it returns or raises:
    /root/reference/utils/helpers.py::_inject_moe_expert_from_deit_mlp   (:481-713; imported by file path, pure torch)
on small stub models (only the attributes the function reads: blocks[i].moe, .mlp.num_expert, .world_size, .tot_expert,
.moe_top_k, model.moe_mlp_ratio / mlp_ratio / moe_experts, model.state_dict()).  Run in THIS container only
(/root/reference does not exist on the GPU box):

    python oracle/make_upcycle_golden.py          ->  tests/golden/upcycle_reference.pt
"""
import importlib.util
import io
import os
import sys
from contextlib import redirect_stdout

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/utils/helpers.py"


def load_ref():
    spec = importlib.util.spec_from_file_location("m3vit_ref_helpers", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Mlp:
    def __init__(self, num_expert, world_size):
        self.num_expert, self.world_size = num_expert, world_size


class _Block:
    def __init__(self, moe, num_expert=0, world_size=1, tot_expert=None, top_k=None):
        self.moe = moe
        if moe:
            self.mlp = _Mlp(num_expert, world_size)
            self.world_size = world_size
            if tot_expert is not None:
                self.tot_expert = tot_expert
            if top_k is not None:
                self.moe_top_k = top_k


class _Model:
    def __init__(self, blocks, moe_mlp_ratio, mlp_ratio, sd):
        self.blocks, self.moe_mlp_ratio, self.mlp_ratio, self._sd = blocks, moe_mlp_ratio, mlp_ratio, sd

    def state_dict(self):
        return self._sd


def dense_sd(n_blocks, D, Hd, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(n_blocks):
        sd[f"blocks.{i}.mlp.fc1.weight"] = torch.randn(Hd, D, generator=g)
        sd[f"blocks.{i}.mlp.fc1.bias"] = torch.randn(Hd, generator=g)
        sd[f"blocks.{i}.mlp.fc2.weight"] = torch.randn(D, Hd, generator=g)
        sd[f"blocks.{i}.mlp.fc2.bias"] = torch.randn(D, generator=g)
    sd["blocks.0.attn.qkv.weight"] = torch.randn(3 * D, D, generator=g)      # an unrelated key must pass through untouched
    return sd


CASES = [
    # name, D, dense hidden, expert hidden, E_local, world, total, top_k, moe_mlp_ratio, cfg, mode
    dict(name="ratio4_copy", D=16, Hd=64, He=64, E_local=4, world=1, total=4, top_k=2, ratio=4.0, cfg={}, mode="deit_upcycling"),
    dict(name="ratio1_split_repeat", D=16, Hd=64, He=16, E_local=8, world=2, total=16, top_k=4, ratio=1.0, cfg={}, mode="deit_upcycling"),
    dict(name="ratio1_split_scaled", D=16, Hd=64, He=16, E_local=8, world=2, total=16, top_k=4, ratio=1.0,
         cfg={"use_weight_scaling": True}, mode="deit_upcycling"),
    dict(name="ratio1_split_truncate", D=16, Hd=64, He=16, E_local=2, world=4, total=8, top_k=2, ratio=1.0, cfg={}, mode="deit_upcycling"),
    dict(name="ratio2_split_g2_warm_refused", D=16, Hd=64, He=32, E_local=4, world=1, total=4, top_k=2, ratio=2.0, cfg={},
         mode="deit_warm_start"),
    dict(name="ratio_minus1_uses_mlp_ratio", D=16, Hd=64, He=64, E_local=2, world=1, total=2, top_k=1, ratio=-1.0, cfg={},
         mode="deit_upcycling"),
]


def main():
    R = load_ref()
    out = {"source": "aapdo/M3ViT utils/helpers.py::_inject_moe_expert_from_deit_mlp executed verbatim", "cases": {}}
    for ci, c in enumerate(CASES):
        sd = dense_sd(3, c["D"], c["Hd"], 100 + ci)
        # blocks 0 and 2 are MoE blocks, block 1 is dense (must be skipped)
        blocks = [_Block(True, c["E_local"], c["world"], c["total"], c["top_k"]), _Block(False),
                  _Block(True, c["E_local"], c["world"], c["total"], c["top_k"])]
        model_sd = {}
        for i in (0, 2):
            model_sd[f"blocks.{i}.mlp.experts.htoh4.weight"] = torch.empty(c["E_local"], c["He"], c["D"])
            model_sd[f"blocks.{i}.mlp.experts.htoh4.bias"] = torch.empty(c["E_local"], c["He"])
            model_sd[f"blocks.{i}.mlp.experts.h4toh.weight"] = torch.empty(c["E_local"], c["D"], c["He"])
            model_sd[f"blocks.{i}.mlp.experts.h4toh.bias"] = torch.empty(c["E_local"], c["D"])
        model = _Model(blocks, c["ratio"], 4.0, model_sd)
        inp = {k: v.clone() for k, v in sd.items()}
        rec = {"case": c, "input": inp}
        try:
            with redirect_stdout(io.StringIO()):
                res = R._inject_moe_expert_from_deit_mlp(sd, model, c["cfg"], deit_init_mode=c["mode"], verbose=False)
            rec["output"] = {k: v.clone() for k, v in res.items()}
        except Exception as e:          # the refusal is part of the contract
            rec["raises"] = type(e).__name__
        out["cases"][c["name"]] = rec
    path = os.path.join(ROOT, "tests", "golden", "upcycle_reference.pt")
    torch.save(out, path)
    for n, r in out["cases"].items():
        print(n, "raises " + r["raises"] if "raises" in r else sorted(k for k in r["output"] if "experts" in k)[:2])
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())

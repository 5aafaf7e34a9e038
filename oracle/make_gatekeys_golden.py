"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

Executes the REFERENCE'S OWN state-dict conversion for the router keys, unmodified:
    /root/reference/utils/common_config.py::cvt_state_dict   (:31-100)
The module itself cannot be imported here (it pulls in the whole model zoo: fmoe, timm, mmcv ...), so the function's
source segment is cut out of the file with `ast` and executed as is in a namespace holding torch / math / F and a stub
for `read_specific_group_experts` (not reached: moe_data_distributed=True).  Recorded: the state dict the function hands
to `model.load_state_dict`.  Run in THIS container only:

    python oracle/make_gatekeys_golden.py          ->  tests/golden/gatekeys_reference.pt
"""
import ast
import math
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/utils/common_config.py"


def load_fn():
    src = open(REF).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "cvt_state_dict")
    ns = {"torch": torch, "math": math, "F": F, "read_specific_group_experts": lambda sd, rank, n: sd, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns["cvt_state_dict"]


class _Model:
    def load_state_dict(self, sd, strict=False):
        self.loaded = {k: v.clone() for k, v in sd.items()}
        return "ok"


CASES = [
    dict(name="shared_gate_untouched", multi_gate=False, task_one_hot=False, num_tasks=2, gtsd=-1, regu=False),
    dict(name="one_hot_pads_num_tasks_rows", multi_gate=False, task_one_hot=True, num_tasks=2, gtsd=-1, regu=False),
    dict(name="one_hot_pads_task_dim_rows", multi_gate=False, task_one_hot=True, num_tasks=5, gtsd=6, regu=False),
    dict(name="one_hot_regu_untouched", multi_gate=False, task_one_hot=True, num_tasks=2, gtsd=-1, regu=True),
    dict(name="multi_gate_2", multi_gate=True, task_one_hot=True, num_tasks=2, gtsd=-1, regu=False),
    dict(name="multi_gate_4", multi_gate=True, task_one_hot=True, num_tasks=4, gtsd=-1, regu=False),
    dict(name="multi_gate_5", multi_gate=True, task_one_hot=True, num_tasks=5, gtsd=-1, regu=False),
    dict(name="multi_gate_3_gets_two", multi_gate=True, task_one_hot=True, num_tasks=3, gtsd=-1, regu=False),
]


def main():
    fn = load_fn()
    out = {"source": "aapdo/M3ViT utils/common_config.py::cvt_state_dict executed verbatim (function segment)", "cases": {}}
    for ci, c in enumerate(CASES):
        g = torch.Generator().manual_seed(200 + ci)
        sd = {"blocks.1.mlp.gate.w_gate": torch.randn(8, 4, generator=g), "blocks.3.mlp.gate.w_gate": torch.randn(8, 4, generator=g),
              "blocks.1.mlp.experts.htoh4.weight": torch.randn(4, 6, 8, generator=g), "blocks.0.attn.qkv.weight": torch.randn(24, 8, generator=g)}
        args = SimpleNamespace(pos_emb_from_pretrained=True, task_one_hot=c["task_one_hot"], multi_gate=c["multi_gate"],
                               regu_experts_fromtask=c["regu"], gate_task_specific_dim=c["gtsd"], num_tasks=c["num_tasks"],
                               moe_data_distributed=True, rank=0, moe_experts=4, start_epoch=-1)
        p = {"backbone_kwargs": {"pos_embed_interp": False}, "backbone": "VisionTransformer_moe"}
        model = _Model()
        inp = {k: v.clone() for k, v in sd.items()}
        fn(sd, model, p, args, "head")
        out["cases"][c["name"]] = {"case": c, "input": inp, "output": model.loaded}
        print(c["name"], sorted(k for k in model.loaded if "gate" in k), [tuple(v.shape) for k, v in sorted(model.loaded.items()) if "gate" in k][:1])
    path = os.path.join(ROOT, "tests", "golden", "gatekeys_reference.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())

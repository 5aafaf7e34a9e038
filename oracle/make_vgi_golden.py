"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

Executes the REFERENCE'S OWN virtual-group router initialisation, unmodified:
    /root/reference/utils/helpers.py::_auto_virtual_group_size (:715-754)
    /root/reference/utils/helpers.py::_inject_virtual_group_init_for_gates (:757-866)
(the companion of the split upcycling: experts that are copies of the same dense-MLP slice start with the SAME router
column).  The random draw is torch.nn.init.normal_ on a CPU tensor under a fixed torch seed, so the result is reproducible
bit for bit.  Run in THIS container only:

    python oracle/make_vgi_golden.py          ->  tests/golden/vgi_reference.pt
"""
import io
import itertools
import os
import sys
from contextlib import redirect_stdout

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_upcycle_golden import ROOT, _Block, _Model, load_ref  # noqa: E402


def main():
    R = load_ref()
    out = {"source": "aapdo/M3ViT utils/helpers.py::_auto_virtual_group_size / _inject_virtual_group_init_for_gates verbatim"}
    grid = []
    for tot, loc, world, dh, eh in itertools.product((0, 4, 8, 16, 12), (None, 0, 2, 4, 8, 6), (None, 1, 2, 4), (None, 64, 48),
                                                     (None, 16, 64, 0, 24)):
        grid.append(((tot, loc, world, dh, eh),
                     R._auto_virtual_group_size(tot, local_experts=loc, world_size=world, dense_hidden=dh, expert_hidden=eh)))
    out["group_size"] = grid
    cases = {}
    for name, (D, Hd, He, E_local, world, gates) in {
        "split_g4_shared_gate": (16, 64, 16, 8, 2, None),           # tot 16, G = 4
        "split_g4_two_task_gates": (16, 64, 16, 4, 4, 2),           # multi-gate keys blocks.i.mlp.gate.{t}.w_gate
        "copy_g_is_local": (16, 64, 64, 4, 1, None),                # dense == expert hidden -> primary 1 -> G = 1
    }.items():
        tot = E_local * world
        sd = {f"blocks.{i}.mlp.fc1.weight": torch.zeros(Hd, D) for i in (0, 2)}
        model_sd = {}
        for i in (0, 2):
            model_sd[f"blocks.{i}.mlp.experts.htoh4.weight"] = torch.empty(E_local, He, D)
            if gates is None:
                model_sd[f"blocks.{i}.mlp.gate.w_gate"] = torch.zeros(D, tot)
            else:
                for t in range(gates):
                    model_sd[f"blocks.{i}.mlp.gate.{t}.w_gate"] = torch.zeros(D, tot)
        model_sd["blocks.1.mlp.fc1.weight"] = torch.zeros(Hd, D)
        model = _Model([_Block(True, E_local, world, tot), _Block(False), _Block(True, E_local, world, tot)], 1.0, 4.0, model_sd)
        torch.manual_seed(1234)
        with redirect_stdout(io.StringIO()):
            res = R._inject_virtual_group_init_for_gates(dict(sd), model, cfg=None, std=0.02, verbose=False)
        cases[name] = {"shape": (D, Hd, He, E_local, world, gates), "seed": 1234,
                       "gates": {k: v.clone() for k, v in res.items() if k.endswith("w_gate")}}
        print(name, {k: tuple(v.shape) for k, v in cases[name]["gates"].items()})
    out["cases"] = cases
    path = os.path.join(ROOT, "tests", "golden", "vgi_reference.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes;", len(grid), "group-size points")


if __name__ == "__main__":
    sys.exit(main())

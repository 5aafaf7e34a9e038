"""Shared helpers for the parity tests: fixture loading and case regeneration."""
import os

import torch

from m3vit_b200.synthetic import MoECase, make_block_case, make_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(fname):
    fx = torch.load(os.path.join(GOLDEN, fname), weights_only=False)
    case = MoECase(**fx["case"])
    data = make_block_case(case, fx["seed"]) if fx.get("block") else make_case(case, fx["seed"])
    if case.name.startswith("S6"):   # same starvation edit as oracle/make_golden.py
        for w in data["w_gate"]:
            w[:, 3] = 0.0
            w[0, 3] = -50.0
            w[:, 11] = 0.0
            w[0, 11] = -50.0
        data["x"][..., 0] = data["x"][..., 0].abs() + 0.1
    # RNG drift check: regenerated tensors must be the ones the fixture was made from
    for k, (s, a) in fx["checksums"].items():
        t = data["w_gate"][int(k[6:])] if k.startswith("w_gate") and k != "w_gate" else data[k]
        assert abs(float(t.double().sum()) - s) <= 1e-9 * max(1.0, a), f"RNG drift in {k}"
    return fx, case, data


def all_fixtures(prefix=""):
    """MoE-layer fixtures (S*, C*); the Block-level ones (B*) are listed by block_fixtures()."""
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and f.startswith(prefix) and f[0] in "SC")


def block_fixtures():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and f.startswith("B"))

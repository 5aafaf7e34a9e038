"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/ckpt_reference.pt.

Executes the REFERENCE'S OWN checkpoint helpers, unmodified, and records what they return or raise:

    /root/reference/utils/moe_utils.py                   filter_state, read_specific_group_experts,
                                                         validate_single_file_moe_checkpoint_or_raise, save_moe_model_to_dir
                                                         (imported under the fmoe / tree / timm shims of oracle/shim)
    /root/reference/pretrain/utils/moe_checkpoint.py     to_mtl_backbone_state_dict, build_mtl_meta, infer_expert_format,
                                                         merge_moe_sharded_directory                (pure torch, imported as is)

    python oracle/make_ckpt_golden.py            # needs /root/reference; spawns a 2-rank gloo group for the shard writer

The state dicts are small and seeded; the fixture stores inputs AND outputs, so tests/test_checkpoint.py needs neither the
reference nor this script.
"""
import importlib.util
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_state(E, seed, prefix="backbone.blocks.1.mlp.", extra=()):
    g = torch.Generator().manual_seed(seed)
    sd = {
        prefix + "experts.htoh4.weight": torch.randn(E, 6, 4, generator=g),
        prefix + "experts.htoh4.bias": torch.randn(E, 6, generator=g),
        prefix + "experts.h4toh.weight": torch.randn(E, 4, 6, generator=g),
        prefix + "experts.h4toh.bias": torch.randn(E, 4, generator=g),
        prefix + "gate.w_gate": torch.randn(4, 8, generator=g),
        "backbone.blocks.0.mlp.fc1.weight": torch.randn(6, 4, generator=g),
    }
    for k in extra:
        sd[k] = torch.randn(3, generator=g)
    return sd


def outcome(fn, *a, **kw):
    try:
        return ("ok", fn(*a, **kw))
    except Exception as e:  # noqa: BLE001
        return ("raise", type(e).__name__, str(e))


def _shard_writer(rank, world, dirname, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(HERE, "shim"))
    sys.path.insert(0, "/root/reference")
    from utils import moe_utils as R
    full = make_state(8, 11)
    local = R.read_specific_group_experts(dict(full), rank, 8 // world)       # the reference's own slicing rule
    R.save_moe_model_to_dir({"state_dict": local, "epoch": 7, "args": {"world_size": world, "moe_experts": 8}}, dirname)
    dist.barrier()
    dist.destroy_process_group()


def main():
    from utils import moe_utils as R                                                    # shimmed fmoe
    P = load_by_path("ref_moe_checkpoint", "/root/reference/pretrain/utils/moe_checkpoint.py")
    out = {"source": "aapdo/M3ViT utils/moe_utils.py + pretrain/utils/moe_checkpoint.py executed verbatim"}

    # ---- filter_state / read_specific_group_experts
    full = make_state(8, 1)
    out["filter_state"] = {"in": full, "out": dict(R.filter_state(dict(full)))}
    out["read_specific_group_experts"] = [
        {"in": full, "rank": r, "num": n, "out": dict(R.read_specific_group_experts(dict(full), r, n))}
        for r, n in ((0, 4), (1, 4), (3, 2))]

    # ---- validate_single_file_moe_checkpoint_or_raise: (checkpoint, state, local_experts, world_size)
    g8, l2 = make_state(8, 2), make_state(2, 3)
    wrapped = {"module.encoder." + k: v for k, v in g8.items()}
    dense = {"backbone.blocks.0.mlp.fc1.weight": torch.zeros(2, 2)}
    cases = [
        ({}, g8, 8, 1), ({}, dense, 2, 4), ({}, g8, 2, 4), ({}, l2, 2, 4), ({}, g8, 4, 4),
        ({"meta": {"expert_format": "global"}}, g8, 2, 4), ({"meta": {"expert_format": "global"}}, g8, 4, 4),
        ({"meta": {"expert_format": "local"}}, g8, 2, 4),
        ({"args": {"world_size": 4, "moe_experts": 8}}, l2, 2, 4), ({"args": {"world_size": 4, "moe_experts": 8}}, g8, 2, 4),
        ({"args": {"world_size": 1, "moe_experts": 8}}, g8, 2, 4), ({}, wrapped, 2, 4), ({"meta": "oops"}, g8, 2, 4),
    ]
    out["validate"] = [{"ckpt": c, "state": s, "local": le, "world": w,
                        "res": outcome(R.validate_single_file_moe_checkpoint_or_raise, c, s, le, w, "x.pth")[:2]}
                       for c, s, le, w in cases]

    # ---- pretrain helpers
    wrapper_sd = {"module.encoder." + k: v for k, v in make_state(4, 4).items()}
    wrapper_sd.update({"module.head.weight": torch.ones(2), "norm.bias": torch.ones(2), "encoder.cls_token": torch.ones(1)})
    st, dropped = P.to_mtl_backbone_state_dict(wrapper_sd)
    out["to_backbone"] = {"in": wrapper_sd, "out": dict(st), "dropped": list(dropped)}
    out["build_meta"] = [{"state": s, "kw": kw, "out": P.build_mtl_meta(s, "unit", **kw)}
                         for s, kw in ((g8, dict(world_size=4)), (g8, dict(world_size=3)), (dense, dict(world_size=2)),
                                       (l2, dict(world_size=1, moe_experts_global=8)), (g8, dict(world_size=2, moe_experts_local=1)))]
    inf_cases = [
        ({"meta": {"expert_format": "local"}}, g8, None, None), ({"meta": {"expert_format": "global"}}, l2, None, None),
        ({}, dense, None, None), ({}, g8, 8, None), ({}, l2, 8, 4), ({}, l2, 8, None), ({}, l2, None, None),
        ({"args": {"moe_experts": 8, "world_size": 4}}, l2, None, None), ({"args": {"moe_experts": 8, "world_size": 4}}, g8, None, None),
        ({"args": {"moe_experts": 16}}, g8, None, 2), ({}, wrapped, 8, None),
    ]
    out["infer"] = [{"ckpt": c, "state": s, "g": g, "w": w, "out": P.infer_expert_format(c, s, g, w)} for c, s, g, w in inf_cases]

    # ---- shard directory written by the reference's save_moe_model_to_dir on 2 gloo ranks, merged by its own merger
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as td:
        d = os.path.join(td, "ckpt.pth.tar")
        mp.spawn(_shard_writer, args=(2, d, 29631), nprocs=2, join=True)
        files = {n: torch.load(os.path.join(d, n), weights_only=False) for n in sorted(os.listdir(d))}
        base, merged, n = P.merge_moe_sharded_directory(d)
        out["shard_dir"] = {"files": files, "merged": dict(merged), "n": n, "full": make_state(8, 11),
                            "base_epoch": base["epoch"]}
        out["merge_errors"] = []
        with tempfile.TemporaryDirectory() as e1:
            out["merge_errors"].append(("empty", outcome(P.merge_moe_sharded_directory, e1)[:2]))
            torch.save({"state_dict": {}}, os.path.join(e1, "1.pth"))
            out["merge_errors"].append(("no_rank0", outcome(P.merge_moe_sharded_directory, e1)[:2]))
    dst = os.path.join(ROOT, "tests", "golden", "ckpt_reference.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    for c in out["validate"]:
        print("validate", c["res"])
    print("infer", [c["out"] for c in out["infer"]])


if __name__ == "__main__":
    main()

"""Real multi-GPU expert-parallel check (one process per GPU, NCCL + CUDA-IPC peer queues).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/ep_multiproc_check.py

Every rank routes its own tokens; the EP layer output / input gradient must equal the replicated
single-GPU layer (all experts local) on the same tokens, and expert-weight gradients must equal the
sum over ranks of the replicated layer's gradients (all-reduced here only to build the expectation)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn

import m3vit_b200 as M
from m3vit_b200 import ep


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    E_tot, K, D, H, T = 16, 4, 384, 384, 2402
    E_loc = E_tot // world
    cdt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
    act = lambda: nn.Sequential(nn.GELU(), nn.Dropout(0.0))
    torch.manual_seed(0)                                     # identical weights on every rank
    full = M.FMoETransformerMLP(num_expert=E_tot, d_model=D, d_gate=D, d_hidden=H, activation=act(),
                                gate=M.NoisyGate_VMoE, top_k=K, vmoe_noisy_std=0, compute_dtype=cdt).to(dev)
    with torch.no_grad():
        full.experts.htoh4.bias.uniform_(-0.05, 0.05)
        full.experts.h4toh.bias.uniform_(-0.05, 0.05)
    shard = M.FMoETransformerMLP(num_expert=E_loc, d_model=D, d_gate=D, d_hidden=H, activation=act(),
                                 gate=M.NoisyGate_VMoE, world_size=world, top_k=K, vmoe_noisy_std=0,
                                 compute_dtype=cdt).to(dev)
    with torch.no_grad():                                    # utils/moe_utils.py:191-198 slicing rule
        shard.gate.w_gate.copy_(full.gate.w_gate)
        for n in ("htoh4", "h4toh"):
            getattr(shard.experts, n).weight.copy_(getattr(full.experts, n).weight[rank * E_loc:(rank + 1) * E_loc])
            getattr(shard.experts, n).bias.copy_(getattr(full.experts, n).bias[rank * E_loc:(rank + 1) * E_loc])
    ctx = ep.make_context(ep.TorchDistGroup(), dev, arena_bytes=1 << 30, capacity_factor=None)
    ep.attach(shard, ctx)
    full.train(); shard.train()
    gen = torch.Generator(device=dev).manual_seed(100 + rank)     # different tokens per rank
    ok = True
    for it in range(3):
        x = torch.randn(T, D, generator=gen, device=dev)
        g = torch.randn(T, D, generator=gen, device=dev)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        full.zero_grad(set_to_none=True); shard.zero_grad(set_to_none=True)
        ya = full(xa); la = full.gate.get_loss()
        torch.autograd.backward([ya, la], [g, torch.tensor(0.01, device=dev)])
        yb = shard(xb); lb = shard.gate.get_loss()
        torch.autograd.backward([yb, lb], [g, torch.tensor(0.01, device=dev)])
        ctx.check_overflow()
        tol = 1e-5 if cdt == torch.float32 else 2e-2
        def nerr(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
        errs = dict(out=nerr(yb, ya), dx=nerr(xb.grad, xa.grad), loss=abs(float(lb) - float(la)),
                    dwg=nerr(shard.gate.w_gate.grad, full.gate.w_gate.grad))
        for n in ("htoh4", "h4toh"):
            for pn in ("weight", "bias"):
                want = getattr(getattr(full.experts, n), pn).grad.clone()
                dist.all_reduce(want)                            # expectation: sum over ranks' tokens
                got = getattr(getattr(shard.experts, n), pn).grad
                errs[f"{n}.{pn}"] = nerr(got, want[rank * E_loc:(rank + 1) * E_loc])
        bad = {k: v for k, v in errs.items() if v > tol}
        ok &= not bad
        if rank == 0:
            print(f"iter {it} [{cdt}] max errs:", {k: f"{v:.2e}" for k, v in errs.items()}, "BAD" if bad else "ok", flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("EP_CHECK", "PASS" if int(t) == 1 else "FAIL", "world", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()

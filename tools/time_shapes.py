"""fwd+bwd time of one MoE layer call for other reference shapes (C3 ViT-B / PASCAL, ratio-4 experts, small batches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import m3vit_b200 as M
from m3vit_b200.synthetic import device_tokens


def run(name, B, N, D, H, E, K, gates, cdt=torch.bfloat16, iters=10):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    layer = M.FMoETransformerMLP(num_expert=E, d_model=D, d_gate=D + gates, d_hidden=H,
                                 activation=nn.Sequential(nn.GELU(), nn.Dropout(0.0)), gate=M.NoisyGate_VMoE, top_k=K,
                                 vmoe_noisy_std=0, multi_gate=True, compute_dtype=cdt).to(dev).train()
    T = B * N
    x = device_tokens(T, D, 0, dev).requires_grad_(True)
    g = torch.randn(T, D, device=dev) * 0.01
    w = torch.tensor(0.01, device=dev)

    def step(fwd_only):
        x.grad = None
        out = layer(x, task_id=0)
        if not fwd_only:
            torch.autograd.backward([out, layer.gate[0].get_loss()], [g, w])
    res = []
    for fo in (True, False):
        for _ in range(3):
            step(fo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step(fo)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / iters * 1000)
    flops = 3 * (2 * D * E + K * 4 * D * H) * T
    print(f"{name:28s} T={T:6d} D={D} H={H} E={E} K={K}: fwd {res[0]:7.1f} us  fwd+bwd {res[1]:7.1f} us  "
          f"{T / res[1]:6.1f} Mtok/s  {flops / res[1] / 1e6:6.1f} TF/s", flush=True)


if __name__ == "__main__":
    run("C2 ViT-S NYUD b32", 32, 1201, 384, 384, 16, 4, 2)
    run("C1 ViT-S NYUD b2", 2, 1201, 384, 384, 16, 4, 2)
    run("ViT-S NYUD b8", 8, 1201, 384, 384, 16, 4, 2)
    run("C3 ViT-B PASCAL b32", 32, 1025, 768, 768, 16, 4, 5)
    run("C3 ViT-B PASCAL b8", 8, 1025, 768, 768, 16, 4, 5)
    run("ViT-S ratio4 b32", 32, 1201, 384, 1536, 16, 4, 2)
    run("ViT-S E64 K2 b32", 32, 1201, 384, 384, 64, 2, 2)
    run("C4 taskcond-like N8193 b4", 4, 8193, 384, 384, 16, 4, 2)

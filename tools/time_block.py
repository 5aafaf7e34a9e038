"""Times the Block-level path  x + mlp(norm2(x))  fused (f1) vs unfused (torch LayerNorm + add) on one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import m3vit_b200 as M


def main():
    dev = torch.device("cuda:0")
    B, N, D, E, K = 32, 1201, 384, 16, 4
    x0 = torch.randn(B, N, D, device=dev) * 2 + 0.3
    g = torch.randn(B, N, D, device=dev)
    ncu = "ncu" in sys.argv          # short fused-only run for an ncu launch list
    for cdt in (torch.bfloat16,):
        for fuse in ((True,) if ncu else (False, True)):
            torch.manual_seed(0)
            blk = M.MoEBlockMlp(D, norm_layer=lambda d: nn.LayerNorm(d, eps=1e-6), fuse=fuse, moe_mlp_ratio=1,
                                moe_experts=E, moe_top_k=K, moe_gate_dim=D + 2, moe_gate_type="noisy_vmoe",
                                vmoe_noisy_std=0, multi_gate=True, compute_dtype=cdt).to(dev).train()

            def step(fwd_only=False):
                x = x0.clone().requires_grad_(True)
                out = blk(x, task_id=0)
                if not fwd_only:
                    loss = blk.mlp.gate[0].get_loss(clear=False)
                    torch.autograd.backward([out, loss], [g, torch.tensor(0.01, device=dev)])
            if ncu:
                for _ in range(3):
                    step(False)
                torch.cuda.synchronize()
                continue
            for fo in (True, False):
                for _ in range(5):
                    step(fo)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 20
                for _ in range(n):
                    step(fo)
                e1.record()
                torch.cuda.synchronize()
                print(f"dtype={cdt} fuse={fuse} {'fwd' if fo else 'fwd+bwd'}: {e0.elapsed_time(e1) / n * 1000:.1f} us/call")


if __name__ == "__main__":
    main()

"""Single-GPU probe of the overlapped push (csrc/ep_push.cuh): a world-1 "expert-parallel" context at bench size, so every
row the pusher CTAs move stays on this GPU (no NVLink) - isolates the pusher's own pipeline from the link.
    python tools/push_probe.py [push_ctas ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from m3vit_b200 import ep, ops
from m3vit_b200.synthetic import MoECase, device_tokens, make_weights


class SimGroup:
    def __init__(self):
        self.rank, self.world = 0, 1

    def barrier(self, device):
        pass


def main():
    dev = torch.device("cuda:0")
    B = 32
    T, D, H, K, E = B * bench.N_TOK, bench.D_MODEL, bench.D_HID, bench.TOP_K, bench.N_EXP
    cdt = torch.bfloat16
    w = make_weights(MoECase("C2", 1, bench.N_TOK, D, H, E, K, 2), 0)
    wg = w["w_gate"][0].to(dev)
    w1c, w1t = ops.cast_weights_bf16(w["w1"].to(dev), True, True)
    w2c, w2t = ops.cast_weights_bf16(w["w2"].to(dev), True, True)
    b1, b2 = w["b1"].to(dev), w["b2"].to(dev)
    x = device_tokens(T, D, 0, dev)
    go = torch.randn(T, D, device=dev) * 0.01
    arena = ep.Arena(3 << 30, dev)
    bases = torch.tensor([arena.base], dtype=torch.int64, device=dev)
    for pc in [int(a) for a in sys.argv[1:]] or [0, 8, 16, 32, 64]:
        os.environ["M3_EP_OVERLAP"] = "1" if pc else "0"
        ctx = ep.EPContext(0, 1, SimGroup(), arena, bases, None, torch.zeros(1, dtype=torch.int32, device=dev))
        ctx.push_ctas = max(pc, 2)
        acc = {}
        for it in range(6):
            evs = []

            def mark(n):
                e = torch.cuda.Event(enable_timing=True); e.record(); evs.append((n, e))
            mark("s")
            st = ep.phase_a_gate(x, wg, K, None, None, 0.0, False, E, ctx, cdt)
            cnt = st.plan_local.counts.view(1, -1).clone()
            ctx.apply_deferred_frees()
            ep.phase_b_dispatch(ctx, st, x, cnt, E, K, cdt); mark("B plan(+push)")
            ep.phase_c_ffn(ctx, st, w1c, b1, w2c, b2, True); mark("C fc1(+push) fc2")
            out = ep.phase_d_combine(ctx, st, T, D, K, torch.float32); mark("D combine")
            bs = ep.phase_e_combine_bwd(ctx, st, go, K); mark("E (push dy)")
            ep.phase_f_ffn_bwd(ctx, st, bs, w1c, w2c, w1t, w2t, parts=1); mark("F dgrad(+push dy)")
            ep.phase_f_ffn_bwd(ctx, st, bs, w1c, w2c, w1t, w2t, parts=2); mark("F wgrad")
            ep.release_bwd(ctx, st, bs); ep.release_fwd(ctx, st)
            torch.cuda.synchronize()
            if it >= 2:
                for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
                    acc.setdefault(n1, []).append(e0.elapsed_time(e1) * 1e3)
        print(f"push_ctas {pc:3d} (overlap {'on' if pc else 'off'}): " +
              "  ".join(f"{n} {sum(v) / len(v):7.1f}" for n, v in acc.items()), flush=True)
        arena.free(ctx.off_cnt, 8192)


main()

"""Per-phase CUDA-event timing of one expert-parallel layer call (fwd+bwd), rank 0 prints.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 tools/ep_phase_timing.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from m3vit_b200 import ep, ops
from m3vit_b200.synthetic import device_tokens, MoECase, make_weights


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T, D, H, K, E = B * bench.N_TOK, bench.D_MODEL, bench.D_HID, bench.TOP_K, bench.N_EXP
    E_loc = E // world
    cdt = torch.bfloat16
    q_bytes = ((int(2.0 * T * K) + E_loc * 255 + 255) // 256 * 256) * D * 2
    ctx = ep.make_context(ep.TorchDistGroup(), dev, arena_bytes=8 * (q_bytes + 4096), capacity_factor=2.0)
    w = make_weights(MoECase("C2", 1, bench.N_TOK, D, H, E, K, 2), 0)
    sl = slice(rank * E_loc, (rank + 1) * E_loc)
    wg = w["w_gate"][0].to(dev)
    w1c, w1t = ops.cast_weights_bf16(w["w1"][sl].to(dev), True, True)
    w2c, w2t = ops.cast_weights_bf16(w["w2"][sl].to(dev), True, True)
    b1, b2 = w["b1"][sl].to(dev), w["b2"][sl].to(dev)
    x = device_tokens(T, D, rank, dev)
    go = torch.randn(T, D, device=dev) * 0.01
    names, acc = [], {}

    def run(timed):
        evs = []

        def mark(n):
            if timed:
                e = torch.cuda.Event(enable_timing=True); e.record(); evs.append((n, e))
        grp = ctx.group
        mark("start")
        st = ep.phase_a_gate(x, wg, K, None, None, 0.0, False, E, ctx, cdt); mark("A gate+plan")
        cnt = grp.all_gather_counts(st.plan_local.counts); mark("  counts gather")
        ctx.apply_deferred_frees()
        ep.phase_b_dispatch(ctx, st, x, cnt, E_loc, K, cdt); mark("B ep_plan+push x")
        grp.barrier(dev); mark("  barrier")
        ep.phase_c_ffn(ctx, st, w1c, b1, w2c, b2, True); mark("C ffn fwd")
        grp.barrier(dev); mark("  barrier")
        out = ep.phase_d_combine(ctx, st, T, D, K, torch.float32); mark("D combine (pull y | local)")
        grp.barrier(dev); mark("  barrier")
        bs = ep.phase_e_combine_bwd(ctx, st, go, K); mark("E push dy")
        grp.barrier(dev); mark("  barrier")
        ep.phase_f_ffn_bwd(ctx, st, bs, w1c, w2c, w1t, w2t); mark("F ffn bwd")
        dz, dwg, _, _ = ops.gate_bwd(x, wg, st.g.noisy_logits, st.g.idx_full, K, dscore=bs.dscore); mark("  gate bwd")
        grp.barrier(dev); mark("  barrier")
        dx = ep.phase_g_dispatch_bwd(ctx, st, bs, T, D, K, dz, wg, torch.float32); mark("G dx (pull | local)")
        grp.barrier(dev); mark("  barrier")
        ep.release_bwd(ctx, st, bs); ep.release_fwd(ctx, st)
        if timed:
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
                acc.setdefault(n1, []).append(e0.elapsed_time(e1) * 1e3)
            if not names:
                names.extend(n for n, _ in evs[1:])
    for _ in range(3):
        run(False)
    torch.cuda.synchronize(); dist.barrier()
    for _ in range(10):
        run(True)
    if rank == 0:
        tot = 0
        seen = []
        for n in names:
            if n in seen and not n.startswith("  barrier"):
                continue
            seen.append(n)
        # names repeat for barriers; print in order with per-occurrence mean
        import collections
        order = names[: len(names)]
        occ = collections.Counter()
        per = {}
        for n in order:
            pass
        vals = {n: acc[n] for n in acc}
        nb = sum(1 for n in order if n == "  barrier")
        for n in dict.fromkeys(order):
            v = vals[n]
            k = nb if n == "  barrier" else 1
            mean = sum(v) / len(v) * (1 if n != "  barrier" else 1)
            print(f"{n:24s} {mean:8.1f} us" + (f"  (x{k} per call, mean each)" if n == "  barrier" else ""))
            tot += mean * k
        print(f"total ~{tot:.0f} us per layer call fwd+bwd, world {world}, T={T}")
    dist.destroy_process_group()


main()

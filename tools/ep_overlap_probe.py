"""How fast are the NVLink row movers when a persistent GEMM holds most SMs?  (design probe for EP overlap)
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/ep_overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from m3vit_b200 import ep, ops
from m3vit_b200._lib import load, check


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    T, D, H, K, E = 32 * bench.N_TOK, bench.D_MODEL, bench.D_HID, bench.TOP_K, bench.N_EXP
    E_loc = E // world
    cdt = torch.bfloat16
    q_bytes = ((int(2.0 * T * K) + E_loc * 255 + 255) // 256 * 256) * D * 2
    ctx = ep.make_context(ep.TorchDistGroup(), dev, arena_bytes=6 * (q_bytes + 4096), capacity_factor=2.0)
    from m3vit_b200.synthetic import device_tokens, MoECase, make_weights
    w = make_weights(MoECase("C2", 1, bench.N_TOK, D, H, E, K, 2), 0)
    wg = w["w_gate"][0].to(dev)
    x = device_tokens(T, D, rank, dev)
    lib = load()
    sink = torch.zeros(1, dtype=torch.int32, device=dev)
    side = torch.cuda.Stream(device=dev)
    grp = ctx.group
    for n_occ in (0, 112, 128, 136):
        ts = {"push": [], "pull": []}
        for it in range(4):
            st = ep.phase_a_gate(x, wg, K, None, None, 0.0, False, E)
            cnt = grp.all_gather_counts(st.plan_local.counts)
            torch.cuda.synchronize(); dist.barrier()
            if n_occ:
                with torch.cuda.stream(side):
                    check(lib.m3_debug_occupy(n_occ, 3_000_000, sink.data_ptr(), side.cuda_stream), "occupy")
                torch.cuda._sleep(200_000)          # let the occupier take its SMs first
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ep.phase_b_dispatch(ctx, st, x, cnt, E_loc, K, cdt)
            e1.record()
            torch.cuda.synchronize(); dist.barrier()
            ts["push"].append(e0.elapsed_time(e1) * 1000)
            # pull: read the x queue back as if it were y (same bytes, same pattern)
            st.off_yq = st.off_xq
            st.yq = st.xq
            if n_occ:
                with torch.cuda.stream(side):
                    check(lib.m3_debug_occupy(n_occ, 3_000_000, sink.data_ptr(), side.cuda_stream), "occupy")
                torch.cuda._sleep(200_000)
            e0.record()
            out = ep.phase_d_combine(ctx, st, T, D, K, torch.float32, keep_rows=True)
            e1.record()
            torch.cuda.synchronize(); dist.barrier()
            ts["pull"].append(e0.elapsed_time(e1) * 1000)
            st.off_yq = -1
            ep.release_fwd(ctx, st)
        if rank == 0:
            print(f"occupied SMs {n_occ:3d}: push x {min(ts['push']):7.1f} us   pull y+combine {min(ts['pull']):7.1f} us", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Quick CUDA-event timing of individual ops at bench size: python tools/time_ops.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from m3vit_b200 import ops
from m3vit_b200.synthetic import device_tokens, MoECase, make_weights

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, D, H, K, E = B * bench.N_TOK, 384, 384, 4, 16
w = make_weights(MoECase("C2", 1, bench.N_TOK, D, H, E, K, 2), 0)
wg = w["w_gate"][0].to(dev)
x = device_tokens(T, D, 0, dev)
g = ops.gate_fwd(x, wg, K)
plan = ops.route_plan(g.idx, E, imp_partial=g.imp_partial, load_partial=g.load_partial)
xq = ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16)
dxq = torch.randn_like(xq)
dz = torch.randn(T, E, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(name, fn, iters=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0
    for _ in range(iters):
        flush.zero_(); a.record(); fn(); b.record(); b.synchronize(); tot += a.elapsed_time(b)
    print(f"{name:40s} {tot / iters * 1e3:8.1f} us")

timed("dispatch_bwd (no router term)", lambda: ops.dispatch_bwd(dxq, plan, T, K))
timed("dispatch_bwd (+ dz @ w_gate^T)", lambda: ops.dispatch_bwd(dxq, plan, T, K, dz=dz, w_gate=wg))
timed("combine_fwd", lambda: ops.combine_fwd(xq, plan, g.score))
timed("gate_fwd fp32 x", lambda: ops.gate_fwd(x, wg, K))
timed("gate_fwd bf16 x", lambda: ops.gate_fwd(x.bfloat16(), wg, K))
xb = x.bfloat16()
timed("gate_fwd bf16 x (pre-cast)", lambda: ops.gate_fwd(xb, wg, K))
timed("gate_bwd", lambda: ops.gate_bwd(x, wg, g.noisy_logits, g.idx_full, K, dscore=g.score))
timed("route_plan", lambda: ops.route_plan(g.idx, E, imp_partial=g.imp_partial, load_partial=g.load_partial))
timed("torch copy x (read+write 118MB)", lambda: x.clone())

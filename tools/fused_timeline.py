"""Timeline of the fused FFN kernel (CTA 0): python tools/fused_timeline.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from m3vit_b200 import ops, _lib
from m3vit_b200.synthetic import device_tokens, MoECase, make_weights
dev = torch.device("cuda:0")
T, D, H, K, E = 32 * bench.N_TOK, 384, 384, 4, 16
w = make_weights(MoECase("C2", 1, bench.N_TOK, D, H, E, K, 2), 0)
x = device_tokens(T, D, 0, dev)
g = ops.gate_fwd(x, w["w_gate"][0].to(dev), K)
plan = ops.route_plan(g.idx, E)
xq = ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16)
w1c, _ = ops.cast_weights_bf16(w["w1"].to(dev), True, False)
w2c, _ = ops.cast_weights_bf16(w["w2"].to(dev), True, False)
b1, b2 = w["b1"].to(dev), w["b2"].to(dev)
for _ in range(2): ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
lib = _lib.load()
lib.m3_debug_trace(1, None, 0)
ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
buf = (C.c_ulonglong * 8192)()
n = lib.m3_debug_trace(0, buf, 4096)
ev = sorted(((buf[2*i+1], buf[2*i] >> 48, (buf[2*i] >> 32) & 0xffff, buf[2*i] & 0xffffffff) for i in range(n)))
t0 = ev[0][0]
names = {(0,0):"P wait x_empty",(0,1):"P got x_empty",(1,0):"M wait x_full",(1,1):"M got x_full",(1,2):"M got a1_empty",(1,3):"M issued G1",(1,4):"M got h_full",(1,5):"M issued G2",
         (2,0):"E wait a1_full",(2,1):"E got a1_full",(2,2):"E tmem read, a1 released",(2,3):"E math+hpre done",(2,4):"E h written",(2,5):"E wait a2_full",(2,6):"E got a2_full",(2,7):"E final done"}
for t, role, e, j in ev[:140]:
    print(f"{t - t0:8d}  {names.get((role, e), (role, e))} j={j}")

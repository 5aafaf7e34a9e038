"""clock64 timeline of CTA 0 of the expert-FFN chain kernel (producer / MMA / first epilogue warp):
    M3_GEMM_TRACE=1 python -m m3vit_b200.build          # separate trace library
    M3_LIB_PATH=m3vit_b200/lib/libm3vit_moe_trace.so python tools/chain_timeline.py [fwd|bwd] [dbg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from m3vit_b200 import ops, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dbg = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
lib = _lib.load()
dev = torch.device("cuda:0")
T, D, H, K, E = 32 * 1201, 384, 384, 4, 16
torch.manual_seed(0)
w1c, w1t = ops.cast_weights_bf16(torch.randn(E, H, D, device=dev) / D ** 0.5, True, True)
w2c, w2t = ops.cast_weights_bf16(torch.randn(E, D, H, device=dev) / H ** 0.5, True, True)
b1, b2 = torch.randn(E, H, device=dev) * 0.1, torch.randn(E, D, device=dev) * 0.1
x = torch.randn(T, D, device=dev)
idx = torch.rand(T, E, device=dev).topk(K, 1).indices
plan = ops.route_plan(idx, E)
xq = ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16)
yq, z = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
dyq = torch.randn_like(yq) * 0.05
fn = (lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)) if which == "fwd" else \
     (lambda: ops.ffn_bwd(xq, z, dyq, plan, w1c, w2c, w1t, w2t))
for _ in range(2):
    fn()
CAP = 6000
lib.m3_set_knob(4, dbg)
names = {0x00: "P  wait W1 slot", 0x01: "P  got W1 slot", 0x02: "P  x tile: wait boxes", 0x04: "P  wait W2 slot", 0x05: "P  got W2 slot",
         0x10: "M  wait acc1 free", 0x11: "M  got acc1", 0x12: "M  got W1 slot", 0x13: "M  G1 issued", 0x14: "M  wait h chunk",
         0x15: "M  got h chunk", 0x16: "M  got W2 slot", 0x17: "M  G2 issued",
         0x20: "E0 wait acc1 full", 0x21: "E0 got acc1", 0x22: "E0 tmem read, acc1 released", 0x23: "E0 math + stores done",
         0x24: "E0 got h buffer", 0x25: "E0 h written", 0x26: "E0 wait acc2", 0x27: "E0 got acc2", 0x28: "E0 tile stored"}
lim = int(os.environ.get("TL_EVENTS", "160"))
skip = int(os.environ.get("TL_SKIP", "300"))
buf = torch.zeros(4 + 2 * 3 * CAP, dtype=torch.int64, device=dev)
lib.m3_set_knob(5, 1)
torch.cuda.synchronize()
lib.m3_debug_trace_buffer(buf.data_ptr(), CAP)
fn()
torch.cuda.synchronize()
lib.m3_debug_trace_buffer(None, 0)
h = buf.cpu().tolist()
ev = []
for r in range(3):
    for i in range(min(h[r], CAP)):
        o = 4 + 2 * (r * CAP + i)
        ev.append((h[o + 1], h[o] >> 32, h[o] & 0xffffffff))
ev.sort()
t0 = ev[0][0]
print(f"--- chain {which} dbg={dbg:#x}: {len(ev)} events, {ev[-1][0] - t0} clk (CTA 0)")
for t, tag, j in ev[skip:skip + lim]:
    print(f"{t - t0:9d}  {names.get(tag, hex(tag)):28s} {j}")
lib.m3_set_knob(5, 0)
lib.m3_set_knob(4, 0)

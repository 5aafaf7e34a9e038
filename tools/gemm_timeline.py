"""clock64 timeline of CTA 0 of the tensor-core GEMM kernels (producer / MMA / first epilogue warp):
    M3_GEMM_TRACE=1 python -m m3vit_b200.build --force      # trace points are compiled in only on request
    python tools/gemm_timeline.py [fwd|bwd] [dbg]            (m3_debug_trace_buffer, include/m3vit_moe.h)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from m3vit_b200 import ops, _lib
from m3vit_b200.synthetic import device_tokens

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dbg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib = _lib.load()
dev = torch.device("cuda:0")
T, D, H, K, E = 32 * bench.N_TOK, bench.D_MODEL, bench.D_HID, bench.TOP_K, bench.N_EXP
layer = bench.build_layers(dev, torch.bfloat16)[0]
x = device_tokens(T, D, 0, dev)
wg = layer.gate[0].w_gate.detach()
b1, b2 = layer.experts.htoh4.bias.detach(), layer.experts.h4toh.bias.detach()
w1c, w2c, w1t, w2t = layer._wcache.get_bf16(layer.experts.htoh4.weight, layer.experts.h4toh.weight)
g = ops.gate_fwd(x, wg, K)
plan = ops.route_plan(g.idx, E, imp_partial=g.imp_partial, load_partial=g.load_partial)
xq = ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16)
yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
dyq, _ = ops.combine_bwd(torch.randn(T, D, device=dev), yq, plan, g.score)
fn = (lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)) if which == "fwd" else \
     (lambda: ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t))
for _ in range(2):
    fn()
CAP = 4000
lib.m3_set_knob(4, dbg)
kernels = ["fc1", "fc2"] if which == "fwd" else ["dgelu", "dx", "wgrad2", "wgrad1"]
# every kernel of the call overwrites the trace, so trace one kernel per run: SKIP mask leaves only kernel i
names = {0x00: "P  wait empty", 0x01: "P  got empty", 0x10: "M  wait", 0x11: "M  got tempty", 0x12: "M  got full",
         0x13: "M  issued+commit", 0x20: "E0 wait tfull", 0x21: "E0 got tfull", 0x22: "E0 tmem read", 0x23: "E0 block stored"}
lim = int(os.environ.get("TL_EVENTS", "120"))
skip = int(os.environ.get("TL_SKIP", "200"))
for ki, kname in enumerate(kernels):
    buf = torch.zeros(4 + 2 * 3 * CAP, dtype=torch.int64, device=dev)
    lib.m3_set_knob(5, ki + 1)                  # M3_KNOB_TRACE_KERNEL: only launch ki of the call writes the timeline
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
    if not ev:
        print(f"--- {kname}: no events"); continue
    t0 = ev[0][0]
    print(f"--- {which} dbg={dbg} kernel {kname}: {len(ev)} events, {ev[-1][0] - t0} clk (CTA 0)")
    for t, tag, j in ev[skip:skip + lim]:
        print(f"{t - t0:9d}  {names.get(tag, hex(tag)):18s} {j}")
lib.m3_set_knob(5, 0)
lib.m3_set_knob(4, 0)

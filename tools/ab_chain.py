"""A/B of the expert FFN: chain kernel (default) vs the two grouped GEMMs, CUDA events, L2 flushed, 4 operand sets.
    python tools/ab_chain.py [batch] [D] [H] [E] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from m3vit_b200 import ops, _lib

lib = _lib.load()
dev = torch.device("cuda:0")
a = [int(v) for v in sys.argv[1:]]
B, D, H, E, K = (a + [32, 384, 384, 16, 4][len(a):])[:5]
T = B * 1201
torch.manual_seed(0)
w1c, w1t = ops.cast_weights_bf16(torch.randn(E, H, D, device=dev) / D ** 0.5, True, True)
w2c, w2t = ops.cast_weights_bf16(torch.randn(E, D, H, device=dev) / H ** 0.5, True, True)
b1, b2 = torch.randn(E, H, device=dev) * 0.1, torch.randn(E, D, device=dev) * 0.1
sets = []
for i in range(4):
    x = torch.randn(T, D, device=dev)
    idx = torch.rand(T, E, device=dev).topk(K, 1).indices
    plan = ops.route_plan(idx, E)
    xq = ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16)
    sets.append((plan, xq, torch.randn_like(xq) * 0.05, i))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
R = T * K


def timed(fn, iters=10):
    for s in sets[:2]:
        fn(s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_(); e0.record()
        for s in sets:
            fn(s)
        e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters / len(sets) * 1e3


print(f"T={T} D={D} H={H} E={E} K={K}")
for chain_on, nm in ((0, "two grouped GEMMs"), (2, "chain kernel"), (0, "two grouped GEMMs"), (2, "chain kernel")):
    lib.m3_set_knob(6, chain_on)
    saved = [ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)[1] for plan, xq, _, _ in sets]
    tf = timed(lambda s: ops.ffn_fwd(s[1], s[0], w1c, b1, w2c, b2))
    ti = timed(lambda s: ops.ffn_fwd(s[1], s[0], w1c, b1, w2c, b2, save_hpre=False))
    tb = timed(lambda s: ops.ffn_bwd(s[1], saved[s[3]], s[2], s[0], w1c, w2c, w1t, w2t))
    print(f"[{nm:18s}] ffn_fwd train {tf:7.1f} us ({4.0 * R * D * H / tf / 1e6:6.0f} TF/s)  inference {ti:7.1f} us  "
          f"ffn_bwd {tb:7.1f} us ({8.0 * R * D * H / tb / 1e6:6.0f} TF/s)", flush=True)
lib.m3_set_knob(6, 1)
# measurement modes of the chain kernel (M3_KNOB_DEBUG: 1 no MMAs, 2 no TMA loads, 4 no epilogue work; results are garbage)
lib.m3_set_knob(6, 2)
if os.environ.get("AB_DEBUG_MODES", "1") == "1":
    for dbg, nm in ((0, "normal"), (2, "no loads"), (4, "no epilogue"), (6, "MMAs only"), (5, "loads only"), (3, "epilogue only"), (7, "barriers only")):
        lib.m3_set_knob(4, dbg)
        tf = timed(lambda s: ops.ffn_fwd(s[1], s[0], w1c, b1, w2c, b2))
        print(f"chain fwd [{nm:14s}] {tf:7.1f} us", flush=True)
    lib.m3_set_knob(4, 0)
    for dbg, nm in ((0, "staggered start (default)"), (16, "no stagger"), (0, "staggered start (default)"), (16, "no stagger")):
        lib.m3_set_knob(4, dbg)
        tf = timed(lambda s: ops.ffn_fwd(s[1], s[0], w1c, b1, w2c, b2))
        print(f"chain fwd [{nm:26s}] {tf:7.1f} us", flush=True)
    lib.m3_set_knob(4, 0)

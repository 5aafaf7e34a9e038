"""A/B of the gate_fwd tile configurations (M3_KNOB_GATE_CFG) at a given batch: bit-equality vs the default + CUDA-event time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from m3vit_b200 import ops, _lib
from m3vit_b200.synthetic import device_tokens
lib = _lib.load()
dev = torch.device("cuda:0")
a = [int(v) for v in sys.argv[1:]]
B, D, E, K = (a + [32, 384, 16, 4][len(a):])[:4]
T = B * 1201
xs = [device_tokens(T, D, i, dev) for i in range(4)]
wg = torch.randn(D, E, device=dev) * 0.05
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    for x in xs[:2]:
        fn(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_(); e0.record()
        for x in xs:
            fn(x)
        e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters / len(xs) * 1e3


lib.m3_set_knob(3, 2)          # cfg 1: the round-1 default at this size
ref = ops.gate_fwd(xs[0], wg, K)
for cfg in (0, 2, 3, 4, 5, 6, 0):
    lib.m3_set_knob(3, cfg)
    g = ops.gate_fwd(xs[0], wg, K)
    torch.cuda.synchronize()
    same = torch.equal(g.idx, ref.idx) and torch.equal(g.score, ref.score) and torch.equal(g.clean_logits, ref.clean_logits)
    print(f"gate_fwd knob {cfg} (cfg id {cfg - 1 if cfg else 'auto'}): bits equal to cfg 1: {same}   {timed(lambda x: ops.gate_fwd(x, wg, K)):6.1f} us", flush=True)
lib.m3_set_knob(3, 0)

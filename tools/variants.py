"""A/B timing of the tuning knobs (m3_set_knob) at bench size, in one process:
    python tools/variants.py [batch]
CUDA events, L2 flushed between iterations; also checks that every variant gives the same bits as the default."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from m3vit_b200 import ops, _lib

KNOB_PDL, KNOB_EPI, KNOB_MOVER = 0, 1, 2
lib = _lib.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, D, H, K, E = B * bench.N_TOK, bench.D_MODEL, bench.D_HID, bench.TOP_K, bench.N_EXP
layers = bench.build_layers(dev, torch.bfloat16)
layer = layers[0]
from m3vit_b200.synthetic import device_tokens
x = device_tokens(T, D, 0, dev)
wg = layer.gate[0].w_gate.detach()
b1, b2 = layer.experts.htoh4.bias.detach(), layer.experts.h4toh.bias.detach()
w1c, w2c, w1t, w2t = layer._wcache.get_bf16(layer.experts.htoh4.weight, layer.experts.h4toh.weight)
g = ops.gate_fwd(x, wg, K)
plan = ops.route_plan(g.idx, E, imp_partial=g.imp_partial, load_partial=g.load_partial)
xq = ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16)
yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
go = torch.randn(T, D, device=dev)
dyq, dscore = ops.combine_bwd(go, yq, plan, g.score)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, iters=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_(); a.record(); fn(); b.record(); b.synchronize(); tot += a.elapsed_time(b)
    print(f"{name:58s} {tot / iters * 1e3:8.1f} us", flush=True)


ROWS = int(plan.offsets[-1].item())      # queue rows in use (the tail of the capacity buffer is never written)


def same(a, b):
    if isinstance(a, (tuple, list)):
        return all(same(u, v) for u, v in zip(a, b))
    if a is None or b is None:
        return a is b
    if a.dim() >= 2 and a.shape[0] >= ROWS and a.dtype == torch.bfloat16:
        return torch.equal(a.reshape(a.shape[0], -1)[:ROWS], b.reshape(b.shape[0], -1)[:ROWS])
    return torch.equal(a, b)


print(f"T={T} D={D} H={H} E={E} K={K} rows={ROWS}")
KNOB_GATE = 3
ref_g = ops.gate_fwd(x, wg, K)
for cfg in (0, 1, 2, 3, 4):
    lib.m3_set_knob(KNOB_GATE, cfg)
    og = ops.gate_fwd(x, wg, K)
    torch.cuda.synchronize()
    print(f"  [gate cfg knob {cfg}] idx equal {torch.equal(og.idx, ref_g.idx)} score equal {torch.equal(og.score, ref_g.score)}")
    timed(f"gate_fwd [cfg knob {cfg}]", lambda: ops.gate_fwd(x, wg, K))
lib.m3_set_knob(KNOB_GATE, 0)
timed("dispatch_fwd", lambda: ops.dispatch_fwd(x, plan, K, out_dtype=torch.bfloat16))
dxq_ = torch.randn_like(xq)
dz_ = torch.randn(T, E, device=dev)
timed("dispatch_bwd (no router term)", lambda: ops.dispatch_bwd(dxq_, plan, T, K))
lib.m3_set_knob(KNOB_MOVER, 9)
d_simt = ops.dispatch_bwd(dxq_, plan, T, K, dz=dz_, w_gate=wg)
timed("dispatch_bwd (+ dz @ w_gate^T) [SIMT fp32 router term]", lambda: ops.dispatch_bwd(dxq_, plan, T, K, dz=dz_, w_gate=wg))
lib.m3_set_knob(KNOB_MOVER, 0)
d_mma = ops.dispatch_bwd(dxq_, plan, T, K, dz=dz_, w_gate=wg)
print(f"  mma.sync router term vs SIMT: max abs diff {(d_mma - d_simt).abs().max().item():.3e}, "
      f"normalised {((d_mma - d_simt).norm() / d_simt.norm()).item():.3e}")
timed("dispatch_bwd (+ dz @ w_gate^T) [mma.sync bf16 router term, 2 rows/batch, 6 CTAs/SM (default)]", lambda: ops.dispatch_bwd(dxq_, plan, T, K, dz=dz_, w_gate=wg))
timed("gate_bwd", lambda: ops.gate_bwd(x, wg, g.noisy_logits, g.idx_full, K, dscore=g.score))
timed("route_plan", lambda: ops.route_plan(g.idx, E, imp_partial=g.imp_partial, load_partial=g.load_partial))
timed("torch x.clone() (59 MB read + 59 MB write)", lambda: x.clone())

# ---------------- GEMM epilogue warps: value 0x100 | mask (bit EPI set -> 16 warps); EPI 0 store 1 bias 2 fc1 3 dgelu
ref_f = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
ref_b = ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t)
KNOB_DEBUG = 4
KNOB_CHAIN = 6
for chain_on, nm in ((0, "two grouped GEMMs"), (2, "chain kernel (training too)")):
    lib.m3_set_knob(KNOB_CHAIN, chain_on)
    yq_c, hpre_c = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
    timed(f"ffn_fwd training  [{nm}]", lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2))
    timed(f"ffn_fwd inference [{nm}]", lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, save_hpre=False))
    timed(f"ffn_bwd           [{nm}]", lambda: ops.ffn_bwd(xq, hpre_c, dyq, plan, w1c, w2c, w1t, w2t))
# the knob sweeps below exercise the grouped GEMM kernels
lib.m3_set_knob(KNOB_CHAIN, 0)
yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
ref_f = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
ref_b = ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t)
for pr, name in ((0, "L2 promotion 256 B (default)"), (3, "L2 promotion 128 B"), (1, "no L2 promotion"), (0, "L2 promotion 256 B (default)")):
    lib.m3_set_knob(KNOB_DEBUG, pr << 12)
    timed(f"ffn_fwd (fc1+fc2)      [{name}]", lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2))
    timed(f"ffn_bwd (2 dgrad+2 wg) [{name}]", lambda: ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t))
lib.m3_set_knob(KNOB_DEBUG, 0)
for dbg, name in ((0, "normal"), (5, "loads only"), (6, "MMAs only")):
    lib.m3_set_knob(KNOB_DEBUG, dbg)
    timed(f"ffn_fwd (fc1+fc2)      [{name}]", lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2))
    timed(f"ffn_bwd (2 dgrad+2 wg) [{name}]", lambda: ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t))
lib.m3_set_knob(KNOB_DEBUG, 0)
for mask, name in ((0, "default: fc1 16 warps, others 8 warps x 64-col blocks"), (0x200, "8-warp epilogues with 32-col blocks, 7 stages"), (0x100, "8 warps everywhere")):
    lib.m3_set_knob(KNOB_EPI, mask)
    of = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
    ob = ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t)
    torch.cuda.synchronize()
    print(f"  [{name}] bits equal to default: fwd {same(of, ref_f)}  bwd {same(ob, ref_b)}")
    timed(f"ffn_fwd (fc1+fc2)      [{name}]", lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2))
    timed(f"ffn_bwd (2 dgrad+2 wg) [{name}]", lambda: ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t))
lib.m3_set_knob(KNOB_EPI, 0)
lib.m3_set_knob(KNOB_CHAIN, 1)

# ---------------- combine fwd / bwd (the rows-in-flight / occupancy variants of round 1 were pruned: profiles/r1b_knob_variants.log)
timed("combine_fwd", lambda: ops.combine_fwd(yq, plan, g.score))
timed("combine_bwd", lambda: ops.combine_bwd(go, yq, plan, g.score))

# ---------------- whole step (12 layer calls fwd+bwd), PDL on / off
calls = [(li, t) for t in range(bench.N_TASK) for li in range(bench.N_LAYER)]
xs = [device_tokens(T, D, i, dev).requires_grad_(True) for i in range(len(calls))]
gs = [torch.randn(T, D, device=dev) * 0.01 for _ in range(2)]


def step():
    for l in layers:
        for p_ in l.parameters():
            p_.grad = None
    for i, (li, t) in enumerate(calls):
        xs[i].grad = None
        bench.one_call(layers[li], xs[i], gs[i & 1], t)


def time_step(name, n=10):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"step {name:52s} {ms:8.3f} ms  -> {len(calls) * T / ms / 1e3:7.2f} M tokens/s", flush=True)
    return [xs[0].grad.clone(), layers[0].experts.htoh4.weight.grad.clone(), layers[0].gate[0].w_gate.grad.clone()]


for rep in range(3):
    lib.m3_set_knob(KNOB_CHAIN, 0)
    time_step("two grouped GEMMs per direction")
    lib.m3_set_knob(KNOB_CHAIN, 2)
    time_step("chain kernel in training (saves z)")
lib.m3_set_knob(KNOB_PDL, 1)
time_step("PDL on")
lib.m3_set_knob(KNOB_PDL, 0)
time_step("PDL off (default)")

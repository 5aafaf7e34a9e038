#!/usr/bin/env python
"""Benchmark of the M3ViT MoE-layer hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the MoE layers of a ViT-S M3ViT backbone on
NYUD-shaped tokens: 6 MoE layers (MoE every other block) x 2 task passes
(multi_gate, one backbone pass per task), forward + backward, bf16 expert FFN,
D = H = 384, E = 16, top-4, N = 1201 tokens / image, B images / GPU.
One "step" = those 12 layer calls fwd+bwd over one batch of synthetic tokens.
Metric: MoE-layer tokens/s = (layer calls x B x N) / time, whole job.

value  : tokens resident in HBM, CUDA-event timed, max over ranks.
e2e    : same step through the public module API with HOST (pinned) token buffers:
         per layer call H2D of the tokens (prefetched one call ahead on a copy stream,
         double-buffered) and D2H of the loss scalars inside the timed region.
roofline: the grouped expert-FFN GEMM kernel (gg_kernel), tensor-bound, timed live.
cpu_baseline / --impl reference: the CPU oracle port of the reference layer
         (oracle/moe_oracle.py, fp32 PyTorch on the host cores) on a bounded
         sample (config C1: B = 2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_TOK, D_MODEL, D_HID, N_EXP, TOP_K, N_TASK, N_LAYER = 1201, 384, 384, 16, 4, 2, 6
METRIC = "moe_layer_tokens_per_s_fwd_bwd"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the sampler runs from
    before the warm-up, every sample carries nvidia-smi's own timestamp, and only samples between
    mark_start() and mark_end() are reported (all of them were taken under load)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark_start(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[1]), float(f[2]), f[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        used, window = (inside, "timed region") if inside else (rows[-5:], "nearest samples (region shorter than the sampling period)")
        reasons = set()
        for r in used:
            for n, v in zip(names, r[3]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm = [r[1] for r in used]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": used[-1][2] if used else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(steps, warmup, batch=2, threads=None):
    """The reference layer's CPU implementation (oracle port, fp32 PyTorch) on the host cores:
    config C1 sample (B=2, one task-gate layer call fwd+bwd per step)."""
    from oracle import moe_oracle as O
    from m3vit_b200.synthetic import MoECase, make_weights, make_tokens
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    case = MoECase("C1", batch=batch, tokens=N_TOK, d_model=D_MODEL, d_hidden=D_HID, num_expert=N_EXP, top_k=TOP_K,
                   num_gates=N_TASK)
    w = make_weights(case, 0)
    x0 = make_tokens(case, 0)
    g = torch.randn(x0.shape, generator=torch.Generator().manual_seed(1))
    params = [w[k].clone().requires_grad_(True) for k in ("w1", "b1", "w2", "b2")]
    gates = [wg.clone().requires_grad_(True) for wg in w["w_gate"]]
    times = []
    for it in range(warmup + steps):
        x = x0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out, gd = O.layer_forward(x, gates[it % N_TASK], *params, TOP_K, training=True)
        ((out * g).sum() + 0.01 * gd["loss"]).backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return dict(value=case.T / med, unit="tokens/s", cores=threads, kind="port",
                sample=f"config C1: 1 layer call fwd+bwd, B={batch} x {N_TOK} tokens, fp32 PyTorch CPU oracle "
                       f"(oracle/moe_oracle.py), median of {steps} after {warmup} warm-up",
                ms_per_call=med * 1e3)


# ----------------------------------------------------------------------------- our arm
def build_layers(dev, compute_dtype, rank=0, world=1, ep_ctx=None):
    """6 MoE layers.  world > 1: expert parallel exactly like the reference's get_backbone
    (utils/common_config.py:179-185): moe_experts //= world, rank r owns experts [r*E_loc, (r+1)*E_loc)."""
    import m3vit_b200 as M
    from m3vit_b200.synthetic import MoECase, make_weights
    e_loc = N_EXP // world
    layers = []
    for li in range(N_LAYER):
        case = MoECase("C2", batch=1, tokens=N_TOK, d_model=D_MODEL, d_hidden=D_HID, num_expert=N_EXP, top_k=TOP_K,
                       num_gates=N_TASK)
        w = make_weights(case, li)                     # same seed on every rank: replicated router
        layer = M.build_moe_mlp(D_MODEL, moe_mlp_ratio=D_HID / D_MODEL, moe_experts=e_loc, moe_top_k=TOP_K,
                                moe_gate_dim=D_MODEL + N_TASK, moe_gate_type="noisy_vmoe", vmoe_noisy_std=0,
                                multi_gate=True, world_size=world, compute_dtype=compute_dtype).to(dev)
        sl = slice(rank * e_loc, (rank + 1) * e_loc)
        with torch.no_grad():
            layer.experts.htoh4.weight.copy_(w["w1"][sl]); layer.experts.htoh4.bias.copy_(w["b1"][sl])
            layer.experts.h4toh.weight.copy_(w["w2"][sl]); layer.experts.h4toh.bias.copy_(w["b2"][sl])
            for g, wg in zip(layer.gate, w["w_gate"]):
                g.w_gate.copy_(wg)
        if ep_ctx is not None:
            from m3vit_b200 import ep
            ep.attach(layer, ep_ctx)
        layer.train()
        layers.append(layer)
    return layers


_CV_W = {}


def one_call(layer, x, g, task):
    """forward + backward of one layer call: d(out) = g (what the next block would hand back) and the
    balance loss with weight 0.01 (train/train_utils.py:437-449)."""
    out = layer(x, task_id=task)
    gate_loss = layer.gate[task].get_loss()
    w = _CV_W.get(x.device)
    if w is None:
        w = _CV_W[x.device] = torch.tensor(0.01, device=x.device)
    torch.autograd.backward([out, gate_loss], [g, w])
    return gate_loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--capacity-factor", type=float, default=2.0,
                    help="EP receive-queue rows as a multiple of the local T*K (overflow is detected and raised)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    config = {"workload": "ViT-S M3ViT backbone MoE layers (configs[1]): 6 MoE layers x 2 NYUD task passes, fwd+bwd",
              "d_model": D_MODEL, "d_hidden": D_HID, "experts": N_EXP, "top_k": TOP_K, "tokens_per_image": N_TOK,
              "batch_per_gpu": args.batch, "layer_calls_per_step": N_LAYER * N_TASK, "noise_std": 0,
              "l2": "inputs_exceed_l2 (12 distinct token buffers per step); per-stage timing: L2 flushed, then 4 launches on 4 "
                    "independent operand sets between one event pair"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference(max(args.steps, 3), warmup)
        line = {"metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": warmup, "ms_per_step": cb["ms_per_call"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from m3vit_b200 import ops
    cdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    T = args.batch * N_TOK
    ep_ctx = None
    if world > 1:
        # expert parallel over the GPUs of the box: peer-mapped queues (CUDA IPC over NVLink), see m3vit_b200/ep.py
        from m3vit_b200 import ep
        assert N_EXP % world == 0
        el = 2 if cdt == torch.bfloat16 else 4
        q_bytes = ((int(args.capacity_factor * T * TOP_K) + (N_EXP // world) * 255 + 255) // 256 * 256) * D_MODEL * el
        if os.environ.get("M3_EP_OVERLAP", "0") == "1":
            # opt-in: two half-batches on two streams, the NVLink movers of one half overlap the GEMMs of the other.
            # Correct (tools/ep_multiproc_check.py ... pipe) but host-launch-bound from Python today.
            ep_ctx = ep.make_pipelined_context(ep.TorchDistGroup(), dev, arena_bytes=6 * (q_bytes // 2 + (1 << 20)),
                                               capacity_factor=args.capacity_factor)
        else:
            ep_ctx = ep.make_context(ep.TorchDistGroup(), dev, arena_bytes=6 * (q_bytes + 4096),
                                     capacity_factor=args.capacity_factor)
    layers = build_layers(dev, cdt, rank, world, ep_ctx)
    from m3vit_b200.synthetic import device_tokens
    calls = [(li, t) for t in range(N_TASK) for li in range(N_LAYER)]      # per task: a full backbone pass
    xs = [device_tokens(T, D_MODEL, 100 * rank + i, dev).requires_grad_(True) for i in range(len(calls))]
    gs = [torch.randn(T, D_MODEL, device=dev) * 0.01 for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def zero_grads():
        # what optimizer.zero_grad(set_to_none=True) does at the top of every training step: the first task pass
        # then installs its gradients, the second one accumulates into them
        for l in layers:
            for p_ in l.parameters():
                p_.grad = None

    def step_resident():
        zero_grads()
        for i, (li, t) in enumerate(calls):
            xs[i].grad = None
            one_call(layers[li], xs[i], gs[i & 1], t)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        step_resident()
    for l in layers:
        l.zero_grad(set_to_none=True)
    barrier()
    sampler.mark_start()
    ops.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = None
    if os.environ.get("M3_BENCH_CPROFILE") and rank == 0:      # host-side profile of the timed loop (debug)
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    if prof is not None:
        prof.disable()
        import pstats
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(28)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count
    clocks = sampler.stop()
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / args.steps
    tokens_per_step = len(calls) * T * world
    value = tokens_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) token buffers, H2D per layer call, D2H of the loss scalars
    hx = [torch.empty(T, D_MODEL, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in hx:
        h.copy_(xs[0].detach().cpu())
    hloss = torch.empty(len(calls), dtype=torch.float32).pin_memory()
    dloss = torch.empty(len(calls), dtype=torch.float32, device=dev)

    # The next call's tokens are prefetched on a copy stream into a double-buffered device slot
    # while the current call computes (what an input pipeline does); every byte still crosses PCIe
    # inside the timed region and the compute stream waits for each copy's event.
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty(T, D_MODEL, device=dev) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])          # slot b's previous fwd+bwd has finished
            xbuf[b].copy_(hx[b], non_blocking=True)
            ev_ready[b].record(copy_stream)

    def step_e2e():
        zero_grads()
        cur = torch.cuda.current_stream()
        for b in range(2):
            ev_free[b].record(cur)
        prefetch(0)
        for i, (li, t) in enumerate(calls):
            b = i & 1
            if i + 1 < len(calls):
                prefetch(i + 1)
            cur.wait_event(ev_ready[b])
            x = xbuf[b].detach().requires_grad_(True)
            dloss[i] = one_call(layers[li], x, gs[b], t).detach()
            ev_free[b].record(cur)
        hloss.copy_(dloss, non_blocking=True)
        cur.synchronize()

    e2e_steps = max(2, min(args.steps, 10))
    step_e2e()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        step_e2e()
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1) / e2e_steps
    if world > 1:
        tms = torch.tensor([ems], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ems = float(tms.item())
    e2e = {"value": tokens_per_step / (ems * 1e-3), "unit": "tokens/s", "ms_per_step": ems,
           "h2d_bytes_per_step": len(calls) * T * D_MODEL * 4, "d2h_bytes_per_step": len(calls) * 4}

    # ---- roofline of the dominant kernel family, timed live on this stream
    pk = peaks()
    roof, stages = None, {}
    if ep_ctx is not None:
        ep_ctx.check_overflow()
    if rank == 0:
        ref_layer = layers[0] if world == 1 else build_layers(dev, cdt)[0]     # all-experts-local layer for the table
        roof, stages = kernel_rooflines(ref_layer, xs[0].detach(), dev, cdt, pk)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "stages": stages,
                "parallelism": (f"expert-parallel ep{world}: {N_EXP // world} experts/GPU, router replicated, token rows "
                                "pushed/pulled over NVLink peer queues (no NCCL on the data path)") if world > 1
                else "single GPU, all experts local"}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(5, 2)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_rooflines(layer, x, dev, cdt, pk, iters=10, nset=4):
    """CUDA-event timing of each stage of one layer call (T tokens), after warm-up, with the
    algorithmic bytes / flops of SURVEY.md 8(d).  Returns (dominant-kernel roofline, per-stage table).

    Every stage is launched back to back on `nset` INDEPENDENT sets of operands (different tokens, different
    buffers: each launch streams data no earlier launch touched, as in the real step where 12 layer calls follow each
    other) between ONE pair of events, after an L2 flush; the average launch duration is that time / launches.  A
    single launch between two events also measures ~6-8 us of launch latency, 15-25 % of a 30 us mover."""
    from m3vit_b200 import ops
    from m3vit_b200.synthetic import device_tokens
    T, D = x.shape
    K, E, H = TOP_K, N_EXP, D_HID
    el = 2 if cdt == torch.bfloat16 else 4
    wg = layer.gate[0].w_gate.detach()
    w1, b1 = layer.experts.htoh4.weight.detach(), layer.experts.htoh4.bias.detach()
    w2, b2 = layer.experts.h4toh.weight.detach(), layer.experts.h4toh.bias.detach()
    if cdt == torch.bfloat16:
        w1c, w2c, w1t, w2t = layer._wcache.get_bf16(layer.experts.htoh4.weight, layer.experts.h4toh.weight)
    else:
        w1c, w2c, w1t, w2t = w1, w2, None, None
    R = T * K

    class S:      # one independent set of operands
        pass
    sets = []
    for i in range(nset):
        o = S()
        o.x = x if i == 0 else device_tokens(T, D, 7000 + i, dev)
        o.g = ops.gate_fwd(o.x, wg, K)
        o.plan = ops.route_plan(o.g.idx, E, imp_partial=o.g.imp_partial, load_partial=o.g.load_partial)
        o.xq = ops.dispatch_fwd(o.x, o.plan, K, out_dtype=cdt)
        o.yq, o.hpre = ops.ffn_fwd(o.xq, o.plan, w1c, b1, w2c, b2)
        o.go = torch.randn(T, D, device=dev)
        o.dyq, o.dscore = ops.combine_bwd(o.go, o.yq, o.plan, o.g.score)
        o.dxq = ops.ffn_bwd(o.xq, o.hpre, o.dyq, o.plan, w1c, w2c, w1t, w2t)[0]
        o.dz = ops.gate_bwd(o.x, wg, o.g.noisy_logits, o.g.idx_full, K, dscore=o.dscore)[0]
        sets.append(o)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, nlaunch):
        for o in sets[:2]:
            fn(o)
        tot = 0.0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(iters):
            flush.zero_()                       # L2 flush between timed iterations (256 MiB > 126 MB L2)
            a.record()
            for o in sets:
                fn(o)
            b.record()
            b.synchronize()
            tot += a.elapsed_time(b)
        return tot / iters * 1e-3 / (nlaunch * len(sets))     # seconds per launch

    st = {}

    def add(name, fn, nlaunch, flops=None, nbytes=None):
        t = timed(fn, nlaunch)
        d = {"us_per_launch": t * 1e6, "launches": nlaunch}
        if flops is not None:
            d.update(bound="tensor", achieved=flops / nlaunch / t / 1e12, peak=pk["tf_burst"], unit="TFLOP/s")
        else:
            d.update(bound="hbm", achieved=nbytes / nlaunch / t / 1e9, peak=pk["hbm"], unit="GB/s")
        d["frac"] = d["achieved"] / d["peak"]
        st[name] = d

    add("gate_fwd", lambda o: ops.gate_fwd(o.x, wg, K), 1, nbytes=T * D * 4 + T * (E * 4 + K * 16 + (K + 1) * 8))
    add("route_plan", lambda o: ops.route_plan(o.g.idx, E, imp_partial=o.g.imp_partial, load_partial=o.g.load_partial), 2,
        nbytes=R * (8 + 8 + 4))
    add("dispatch_fwd", lambda o: ops.dispatch_fwd(o.x, o.plan, K, out_dtype=cdt), 1, nbytes=T * (D * 4 + K * D * el + K * 4))
    add("ffn_fwd", lambda o: ops.ffn_fwd(o.xq, o.plan, w1c, b1, w2c, b2), 2, flops=4.0 * R * D * H)
    add("combine_fwd", lambda o: ops.combine_fwd(o.yq, o.plan, o.g.score), 1, nbytes=T * (K * D * el + K * 8 + D * 4))
    add("combine_bwd", lambda o: ops.combine_bwd(o.go, o.yq, o.plan, o.g.score), 1, nbytes=T * (D * 4 + 2 * K * D * el + K * 12))
    nb = 4 if cdt == torch.bfloat16 else 6
    add("ffn_bwd", lambda o: ops.ffn_bwd(o.xq, o.hpre, o.dyq, o.plan, w1c, w2c, w1t, w2t), nb, flops=8.0 * R * D * H)
    add("gate_bwd", lambda o: ops.gate_bwd(o.x, wg, o.g.noisy_logits, o.g.idx_full, K, dscore=o.dscore), 3,
        nbytes=T * (D * 4 + E * 12))
    add("dispatch_bwd", lambda o: ops.dispatch_bwd(o.dxq, o.plan, T, K, dz=o.dz, w_gate=wg), 1,
        nbytes=T * (K * D * el + K * 4 + D * 4 + E * 4))
    f = st["ffn_fwd"]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this exact
    # configuration (profiles/r1b_ncu_all_kernels.md: fc1 124.2 + 186.0 MB, fc2 124.1 + 69.7 MB; writes still in
    # L2 at kernel end are not counted by ncu), averaged over the two launches like `achieved`; null for any other size.
    traffic = 252.0e6 if (cdt == torch.bfloat16 and T == 32 * N_TOK and D == 384 and H == 384 and K == 4) else None
    roof = {"kernel": "gg_kernel<192> (tcgen05 grouped GEMM; fc1+bias+GELU and fc2+bias launches of m3_ffn_fwd)"
            if cdt == torch.bfloat16 else "sgemm_grouped_kernel (fp32 SIMT)",
            "bound": "tensor", "achieved": f["achieved"], "peak": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": f["frac"], "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)", "peak_source": pk["src"] + " (burst, kernel timed alone)",
            "flops_per_launch": 2.0 * R * D * H, "us_per_launch": f["us_per_launch"]}
    return roof, st


if __name__ == "__main__":
    main()

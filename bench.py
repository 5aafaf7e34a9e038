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
roofline: the grouped expert-FFN GEMM kernel (gg_kernel: the fc1 / fc2 launches of a training forward), timed
         live.  Both rooflines are reported (algorithmic bytes / HBM copy peak and flops / cuBLAS bf16 peak); `bound`
         is the one that binds - at D = H = 384 with the activation planes kept for backward the pair runs at
         154 flop/B, below the ridge of 250 flop/B, so that is HBM.
cpu_baseline / --impl reference: the CPU oracle port of the reference layer
         (oracle/moe_oracle.py, fp32 PyTorch on the host cores) on the SAME workload
         (same B, same N, both task gates); each step is a bounded sample of the 12
         layer calls of a step (as many as fit the time budget; the line says how many).
configs: the other BASELINE.json configurations (C1 reference batch, C3 ViT-B 5-gate,
         C4 task-conditioned 8193 tokens, ratio-4 experts, E = 64 / top-2), one layer
         call fwd+bwd each, reported beside the headline line under "configs".
ep_parity (N > 1): an untimed comparison of the expert-parallel layer with the same layer
         holding all experts locally, on this rank's tokens (max over ranks).
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
def cpu_reference(steps, warmup, batch, budget_s=150.0, threads=None):
    """The reference layer's CPU implementation (oracle port of models/moe/origin/custom_moe_layer.py +
    noisy_gate_vmoe.py, fp32 PyTorch) on the host cores, on the bench workload itself: B = `batch` images x 1201
    tokens per layer call, both task gates, fwd+bwd with the balance loss.  One step = `calls` of the 12 layer calls
    of a step (distinct layers / gates in the step's order); `calls` is the largest number <= 12 for which
    (warmup + steps) steps fit `budget_s`, measured on the first call."""
    from oracle import moe_oracle as O
    from m3vit_b200.synthetic import MoECase, make_weights, make_tokens
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    case = MoECase("C2", batch=batch, tokens=N_TOK, d_model=D_MODEL, d_hidden=D_HID, num_expert=N_EXP, top_k=TOP_K,
                   num_gates=N_TASK)
    layers = []
    for li in range(N_LAYER):
        w = make_weights(case, li)
        layers.append(([w[k].clone().requires_grad_(True) for k in ("w1", "b1", "w2", "b2")],
                       [wg.clone().requires_grad_(True) for wg in w["w_gate"]]))
    x0 = make_tokens(case, 0)
    g = torch.randn(x0.shape, generator=torch.Generator().manual_seed(1)) * 0.01
    order = [(li, t) for t in range(N_TASK) for li in range(N_LAYER)]

    def call(i):
        li, t = order[i % len(order)]
        params, gates = layers[li]
        x = x0.clone().requires_grad_(True)
        out, gd = O.layer_forward(x, gates[t], *params, TOP_K, training=True)
        ((out * g).sum() + 0.01 * gd["loss"]).backward()

    t0 = time.perf_counter()
    call(0)                                   # also the first warm-up (page-in, thread pool start)
    t_first = time.perf_counter() - t0
    calls = int(max(1, min(len(order), budget_s / max(t_first * (steps + warmup), 1e-9))))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for i in range(calls):
            call(it * calls + i)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return dict(value=calls * case.T / med, unit="tokens/s", cores=threads, kind="port",
                sample=f"{calls} of the {len(order)} layer calls of a step per timed step, B={batch} x {N_TOK} tokens per call "
                       f"(the bench workload), fp32 PyTorch CPU oracle port (oracle/moe_oracle.py; pinned bit-exact to the "
                       f"reference's own files by tests/test_oracle.py), median of {steps} steps after {warmup} warm-up",
                calls_per_step=calls, ms_per_step=med * 1e3, ms_per_call=med * 1e3 / calls)


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


def other_configs(dev, cdt, pk, iters=8):
    """The other BASELINE.json configurations, one MoE layer call each (fwd and fwd+bwd, CUDA events): C1 (the
    reference's own batch of 2), C3 (ViT-B, PASCAL-Context 5 task gates), C4 (task-conditioned shared router on
    Cityscapes-sized 8193-token images) and the C5 sweep corners.  `tflops` counts the algorithmic flops of SURVEY 8(d):
    3 x (2 Dg E + 4 K D H) per token; `frac` is against the sustained bf16 peak (a whole call: GEMMs, movers, router)."""
    import m3vit_b200 as M
    from m3vit_b200.synthetic import device_tokens
    cases = [
        ("c1_vits_nyud_b2", dict(B=2, N=1201, D=384, H=384, E=16, K=4, gates=2, Dt=0)),
        ("c3_vitb_pascal_5gate_b8", dict(B=8, N=1025, D=768, H=768, E=16, K=4, gates=5, Dt=0)),
        ("c3_vitb_pascal_5gate_b32", dict(B=32, N=1025, D=768, H=768, E=16, K=4, gates=5, Dt=0)),
        ("c4_taskcond_cityscapes_b4", dict(B=4, N=8193, D=384, H=384, E=16, K=4, gates=0, Dt=64)),
        ("c5_ratio4_b32", dict(B=32, N=1201, D=384, H=1536, E=16, K=4, gates=2, Dt=0)),
        ("c5_e64_top2_b32", dict(B=32, N=1201, D=384, H=384, E=64, K=2, gates=2, Dt=0)),
        ("c5_e32_top1_b32", dict(B=32, N=1201, D=384, H=384, E=32, K=1, gates=2, Dt=0)),
    ]
    res = {}
    w01 = torch.tensor(0.01, device=dev)
    for name, c in cases:
        torch.manual_seed(0)
        B, N, D, H, E, K, G, Dt = (c[k] for k in ("B", "N", "D", "H", "E", "K", "gates", "Dt"))
        act = torch.nn.Sequential(torch.nn.GELU(), torch.nn.Dropout(0.0))
        if Dt > 0:      # one shared router conditioned on a task embedding (gate_task_specific_dim > 0, task_one_hot loop)
            layer = M.FMoETransformerMLP(num_expert=E, d_model=D, d_gate=D, d_hidden=H, activation=act, gate=M.NoisyGate_VMoE,
                                         top_k=K, vmoe_noisy_std=0, multi_gate=False, gate_task_specific_dim=Dt,
                                         compute_dtype=cdt).to(dev).train()
            feat = torch.randn(Dt, device=dev)
            gate = layer.gate
            fwd = lambda x: layer(x, task_id=1, task_specific_feature=feat)
        else:
            layer = M.FMoETransformerMLP(num_expert=E, d_model=D, d_gate=D + G, d_hidden=H, activation=act,
                                         gate=M.NoisyGate_VMoE, top_k=K, vmoe_noisy_std=0, multi_gate=True,
                                         compute_dtype=cdt).to(dev).train()
            gate = layer.gate[G - 1]
            fwd = lambda x: layer(x, task_id=G - 1)
        T = B * N
        x = device_tokens(T, D, 0, dev).requires_grad_(True)
        g = torch.randn(T, D, device=dev) * 0.01

        def step(bwd):
            x.grad = None
            out = fwd(x)
            if bwd:
                torch.autograd.backward([out, gate.get_loss()], [g, w01])
        t = []
        for bwd in (False, True):
            for _ in range(2):
                step(bwd)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                step(bwd)
            e1.record()
            e1.synchronize()
            t.append(e0.elapsed_time(e1) / iters * 1e-3)
        flops = 3.0 * (2 * (D + Dt) * E + 4 * K * D * H) * T
        res[name] = {**c, "tokens_per_call": T, "fwd_us": t[0] * 1e6, "fwd_bwd_us": t[1] * 1e6,
                     "tokens_per_s": T / t[1], "tflops": flops / t[1] / 1e12, "frac_of_sustained_bf16": flops / t[1] / 1e12 / pk["tf_sus"]}
        if cdt == torch.bfloat16:
            res[name]["gemm_roofline"] = config_gemm_roofline(layer, gate, x.detach(), K, E, D, H, Dt, feat if Dt > 0 else None,
                                                              cdt, pk, iters)
        del layer, x, g
    return res


def config_gemm_roofline(layer, gate, x, K, E, D, H, Dt, feat, cdt, pk, iters):
    """The expert GEMMs of one configuration on their own (training forward: fc1 + fc2; backward: dz, dxq, dW2, dW1), timed
    with CUDA events over `iters` back-to-back launches on this configuration's own dispatched queue, against BOTH rooflines
    (algorithmic bytes / HBM copy peak, flops / cuBLAS bf16 burst peak); `bound` is the one with the larger floor."""
    from m3vit_b200 import ops
    T = x.shape[0]
    R, el = T * K, 2
    with torch.no_grad():
        g = ops.gate_fwd(x, gate.w_gate.detach(), K, feat)
        plan = ops.route_plan(g.idx, E, imp_partial=g.imp_partial, load_partial=g.load_partial)
        xq = ops.dispatch_fwd(x, plan, K, out_dtype=cdt)
        w1c, w2c, w1t, w2t = layer._wcache.get_bf16(layer.experts.htoh4.weight, layer.experts.h4toh.weight)
        b1, b2 = layer.experts.htoh4.bias.detach(), layer.experts.h4toh.bias.detach()
        yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2)
        dyq = torch.randn_like(yq) * 0.01

        def timed(fn):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1) / iters * 1e-3
        tf = timed(lambda: ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2))
        tb = timed(lambda: ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t))
    wb = E * D * H * el
    out = {}
    for nm, t, flops, nbytes in (("ffn_fwd", tf, 4.0 * R * D * H, R * el * (D + 2 * H) + R * el * (H + D) + 2 * wb),
                                 ("ffn_bwd", tb, 8.0 * R * D * H, R * el * (D + 2 * H) + R * el * (H + D) + 2 * R * el * (D + H)
                                  + 2 * wb + 2 * E * D * H * 4)):
        ft, fh = flops / t / 1e12 / pk["tf_burst"], nbytes / t / 1e9 / pk["hbm"]
        out[nm] = {"us": t * 1e6, "tflops": flops / t / 1e12, "tensor_frac": ft, "hbm_frac": fh,
                   "tensor_frac_of_sustained_peak": flops / t / 1e12 / pk["tf_sus"],      # (informational: launches run back to back)
                   "bound": "tensor" if ft >= fh else "hbm", "frac": max(ft, fh), "flop_per_byte": flops / nbytes}
    return out


def ep_parity(layers, ep_layer_inputs, dev, cdt, rank, world, dist):
    """Untimed: the expert-parallel layer against the SAME layer with all experts local, on this rank's tokens.
    Returns the largest absolute differences over all ranks (out, dx) and the largest normalised difference of the
    local experts' weight gradients against the matching slice of the replicated layer's - which sees only this
    rank's tokens, so that check is done at world-summed level with an all-reduce of the replicated gradients."""
    ep_layer, x0, g0 = ep_layer_inputs
    full = build_layers(dev, cdt)[0]                       # all experts local, same seeds
    res = {}
    outs = []
    for layer in (ep_layer, full):
        for p_ in layer.parameters():
            p_.grad = None
        x = x0.detach().clone().requires_grad_(True)
        out = layer(x, task_id=1)
        torch.autograd.backward([out], [g0])
        outs.append((out.detach(), x.grad.detach(), layer.experts.htoh4.weight.grad.detach().clone(),
                     layer.gate[1].w_gate.grad.detach().clone()))
    (o_ep, dx_ep, dw_ep, dg_ep), (o_f, dx_f, dw_f, dg_f) = outs
    e_loc = N_EXP // world
    dist.all_reduce(dw_f)                                  # an expert's gradient sums over every rank's tokens
    dw_ref = dw_f[rank * e_loc:(rank + 1) * e_loc]
    vals = torch.stack([(o_ep - o_f).abs().max(), (dx_ep - dx_f).abs().max(),
                        (dw_ep - dw_ref).abs().max() / dw_ref.abs().max().clamp_min(1e-20),
                        (dg_ep - dg_f).abs().max() / dg_f.abs().max().clamp_min(1e-20)]).float()
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    v = [float(t) for t in vals]
    res = {"out_max_abs_diff": v[0], "dx_max_abs_diff": v[1], "expert_wgrad_max_rel_diff": v[2],
           "router_wgrad_max_rel_diff": v[3],
           "ok": bool(v[0] <= 1e-6 and v[1] <= 1e-6 and v[2] <= 1e-4 and v[3] <= 1e-4),
           "what": "EP layer vs the same layer with all 16 experts local on each rank's tokens (max over ranks); rows take the "
                   "same arithmetic in both, so out / dx are expected to be bit-identical"}
    for p_ in ep_layer.parameters():
        p_.grad = None
    del full
    return res


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture, valid only for
    the kernel sources it was taken from (profiles/ncu_traffic.json records their sha256): a stale capture reads as null."""
    import hashlib
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no capture"
    d = json.load(open(p))
    e = d.get(kernel_key)
    if e is None:
        return None, "no capture for this configuration"
    h = hashlib.sha256()
    for f in e["sources"]:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    if h.hexdigest() != e["sources_sha256"]:
        return None, "capture is older than the kernel sources"
    return e["dram_bytes_per_launch"], e["capture"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations (\"configs\" key)")
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
        cb = cpu_reference(max(args.steps, 1), warmup, args.batch)      # exactly K timed steps (median), W >= 3 warm-up
        config["reference_sample"] = (f"each timed step = {cb['calls_per_step']} of the {N_LAYER * N_TASK} layer calls (bounded "
                                      "sample of the same workload: same B, N, D, H, E, K, both task gates)")
        line = {"metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
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
        ep_ctx = ep.make_context(ep.TorchDistGroup(), dev, arena_bytes=10 * (q_bytes + 4096),
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
    def run_e2e(tok_dtype):
        """One e2e measurement with host token buffers of `tok_dtype`.  The next call's tokens are prefetched on a copy
        stream into a double-buffered device slot while the current call computes (what an input pipeline does); every
        byte still crosses PCIe inside the timed region and the compute stream waits for each copy's event."""
        hx = [torch.empty(T, D_MODEL, dtype=tok_dtype).pin_memory() for _ in range(2)]
        for h in hx:
            h.copy_(xs[0].detach().cpu().to(tok_dtype))
        hloss = torch.empty(len(calls), dtype=torch.float32).pin_memory()
        dloss = torch.empty(len(calls), dtype=torch.float32, device=dev)
        ge = [g_.to(tok_dtype) for g_ in gs]
        copy_stream = torch.cuda.Stream(device=dev)
        xbuf = [torch.empty(T, D_MODEL, device=dev, dtype=tok_dtype) for _ in range(2)]
        ev_ready = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[b])          # slot b's previous fwd+bwd has finished
                xbuf[b].copy_(hx[b], non_blocking=True)
                ev_ready[b].record(copy_stream)

        def step_e2e(first=False):
            zero_grads()
            cur = torch.cuda.current_stream()
            if first:                                       # prime the pipeline (untimed warm-up step only)
                for b in range(2):
                    ev_free[b].record(cur)
                prefetch(0)
            for i, (li, t) in enumerate(calls):
                b = i & 1
                # the NEXT call's tokens start crossing the bus now; after the last call of a step that is call 0 of the
                # next step (len(calls) is even, so the slot parity carries over): an input pipeline does not drain
                # between steps
                prefetch(i + 1)
                cur.wait_event(ev_ready[b])
                x = xbuf[b].detach().requires_grad_(True)
                dloss[i] = one_call(layers[li], x, ge[b], t).detach()
                ev_free[b].record(cur)
            hloss.copy_(dloss, non_blocking=True)
            cur.synchronize()                               # the step's result (12 balance losses) is on the host

        assert len(calls) % 2 == 0
        e2e_steps = max(2, min(args.steps, 10))
        step_e2e(first=True)
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
        nb = len(calls) * T * D_MODEL * hx[0].element_size()
        return {"value": tokens_per_step / (ems * 1e-3), "unit": "tokens/s", "ms_per_step": ems,
                "h2d_bytes_per_step": nb, "d2h_bytes_per_step": len(calls) * 4, "h2d_gb_per_s": nb / (ems * 1e-3) / 1e9}

    e2e = run_e2e(torch.float32)
    e2e["note"] = ("PCIe-bound: the fp32 token matrices of the 12 layer calls (what the reference's layer is handed) cross the bus "
                   "every step, prefetched one call ahead on a copy stream, across step boundaries too (each timed step issues 12 "
                   "copies, the last of which serves the following step); h2d_gb_per_s is the achieved H2D rate.  "
                   "`bf16_tokens` is the same measurement for a bf16 model (bf16 tokens in, bf16 out / dx): half the bytes on "
                   "the bus - informational, the reference arm and `value` use fp32 tokens")
    if cdt == torch.bfloat16 and world == 1:
        e2e["bf16_tokens"] = run_e2e(torch.bfloat16)

    # ---- roofline of the dominant kernel family, timed live on this stream
    pk = peaks()
    roof, stages = None, {}
    parity = None
    if ep_ctx is not None:
        ep_ctx.check_overflow()
        parity = ep_parity(layers, (layers[0], xs[0], gs[0]), dev, cdt, rank, world, dist)
    if rank == 0:
        ref_layer = layers[0] if world == 1 else build_layers(dev, cdt)[0]     # all-experts-local layer for the table
        roof, stages = kernel_rooflines(ref_layer, xs[0].detach(), dev, cdt, pk)
    extra = None
    if rank == 0 and world == 1 and not args.no_configs:
        extra = other_configs(dev, cdt, pk)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "stages": stages,
                "parallelism": (f"expert-parallel ep{world}: {N_EXP // world} experts/GPU, router replicated, token rows "
                                "pushed/pulled over NVLink peer queues (no NCCL on the data path)") if world > 1
                else "single GPU, all experts local"}
        if parity is not None:
            line["ep_parity"] = parity
        if extra is not None:
            line["configs"] = extra
        if not args.no_cpu_baseline and world == 1:      # (rank 0 at N = 1 only: the other ranks would idle behind it)
            line["cpu_baseline"] = cpu_reference(3, 1, args.batch, budget_s=25.0)
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
        """flops / nbytes: ALGORITHMIC work of the whole stage (all its launches).  A GEMM stage gets both: the roofline
        that bounds it is the one with the larger floor time (at D = H = 384 with the activation planes kept for backward
        the grouped GEMMs sit BELOW the ridge of 250 flop/B: HBM-bound by the model); the other fraction is kept too."""
        t = timed(fn, nlaunch)
        d = {"us_per_launch": t * 1e6, "launches": nlaunch}
        tens = hbm = None
        if flops is not None:
            tens = dict(achieved=flops / nlaunch / t / 1e12, peak=pk["tf_burst"], unit="TFLOP/s")
            tens["frac"] = tens["achieved"] / tens["peak"]
        if nbytes is not None:
            hbm = dict(achieved=nbytes / nlaunch / t / 1e9, peak=pk["hbm"], unit="GB/s")
            hbm["frac"] = hbm["achieved"] / hbm["peak"]
        if tens is not None and (hbm is None or tens["frac"] >= hbm["frac"]):
            d.update(bound="tensor", **tens)
        else:
            d.update(bound="hbm", **hbm)
        if tens is not None and hbm is not None:
            d["tensor"], d["hbm"] = tens, hbm
            d["flop_per_byte"] = flops / nbytes
        st[name] = d

    add("gate_fwd", lambda o: ops.gate_fwd(o.x, wg, K), 1, nbytes=T * D * 4 + T * (E * 4 + K * 16 + (K + 1) * 8))
    add("route_plan", lambda o: ops.route_plan(o.g.idx, E, imp_partial=o.g.imp_partial, load_partial=o.g.load_partial), 2,
        nbytes=R * (8 + 8 + 4))
    add("dispatch_fwd", lambda o: ops.dispatch_fwd(o.x, o.plan, K, out_dtype=cdt), 1, nbytes=T * (D * 4 + K * D * el + K * 4))
    # algorithmic bytes of the expert GEMMs (bf16 path; queue rows R x width x el, weights E x D x H x el per launch):
    #   fc1 (training): xq in, h and gelu'(z) out          fc2: h in, yq out
    #   dz = (dyq W2) * gelu': dyq + gelu' in, dz out      dxq = dz W1: dz in, dxq out
    #   dW2 = dyq^T h, dW1 = dz^T xq: two queue planes in each (fp32 dW out is E x D x H x 4)
    wb = E * D * H * el
    add("ffn_fwd", lambda o: ops.ffn_fwd(o.xq, o.plan, w1c, b1, w2c, b2), 2, flops=4.0 * R * D * H,
        nbytes=R * el * (D + 2 * H) + R * el * (H + D) + 2 * wb)
    from m3vit_b200 import _lib as L_
    chain = bool(L_.load().m3_ffn_uses_chain(1 if cdt == torch.bfloat16 else 0, D, H))
    add("ffn_fwd_inference", lambda o: ops.ffn_fwd(o.xq, o.plan, w1c, b1, w2c, b2, save_hpre=False), 1 if chain else 2,
        flops=4.0 * R * D * H, nbytes=(2 * R * el * D if chain else 2 * R * el * (D + H)) + 2 * wb)
    st["ffn_fwd_inference"]["kernel"] = "ffn_chain_kernel (fc1 -> GELU -> fc2 in one launch)" if chain else "gg_kernel x 2"
    add("combine_fwd", lambda o: ops.combine_fwd(o.yq, o.plan, o.g.score), 1, nbytes=T * (K * D * el + K * 8 + D * 4))
    add("combine_bwd", lambda o: ops.combine_bwd(o.go, o.yq, o.plan, o.g.score), 1, nbytes=T * (D * 4 + 2 * K * D * el + K * 12))
    nb = 4 if cdt == torch.bfloat16 else 6
    add("ffn_bwd", lambda o: ops.ffn_bwd(o.xq, o.hpre, o.dyq, o.plan, w1c, w2c, w1t, w2t), nb, flops=8.0 * R * D * H,
        nbytes=R * el * (D + 2 * H) + R * el * (H + D) + 2 * R * el * (D + H) + 2 * wb + 2 * E * D * H * 4)
    add("gate_bwd", lambda o: ops.gate_bwd(o.x, wg, o.g.noisy_logits, o.g.idx_full, K, dscore=o.dscore), 3,
        nbytes=T * (D * 4 + E * 12))
    add("dispatch_bwd", lambda o: ops.dispatch_bwd(o.dxq, o.plan, T, K, dz=o.dz, w_gate=wg), 1,
        nbytes=T * (K * D * el + K * 4 + D * 4 + E * 4))
    f = st["ffn_fwd"]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch (average of the fc1 and fc2 launches, like `achieved`) from
    # the committed `ncu --set full` capture of this configuration; null when there is none or the kernels changed since
    traffic, traffic_src = ncu_traffic(f"ffn_fwd_train:{'bf16' if cdt == torch.bfloat16 else 'f32'}:T{T}:D{D}:H{H}:E{E}:K{K}")
    roof = {"kernel": "gg_kernel<192> (tcgen05 grouped GEMM; fc1+bias+GELU and fc2+bias launches of m3_ffn_fwd)"
            if cdt == torch.bfloat16 else "sgemm_grouped_kernel (fp32 SIMT)",
            "bound": f["bound"], "achieved": f["achieved"], "peak": f["peak"], "unit": f["unit"],
            "frac": f["frac"], "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
            "traffic_source": traffic_src, "peak_source": pk["src"] + " (burst, kernel timed alone)",
            "flops_per_launch": 2.0 * R * D * H, "us_per_launch": f["us_per_launch"]}
    if "tensor" in f:      # both rooflines of the GEMM pair: the binding one above, the other for reference
        roof.update(tensor=f["tensor"], hbm=f["hbm"], flop_per_byte=f["flop_per_byte"],
                    algorithmic_bytes_per_launch=f["hbm"]["achieved"] * 1e9 * f["us_per_launch"] * 1e-6,
                    note="mean over the fc1 (+bias, GELU, saves gelu') and fc2 (+bias) launches of one training forward; at "
                         "%.0f flop/B the pair sits below the ridge (%.0f flop/B): the HBM roofline binds, the tensor "
                         "fraction is what the same time means against the cuBLAS bf16 peak"
                         % (f["flop_per_byte"], pk["tf_burst"] * 1e12 / (pk["hbm"] * 1e9)))
    return roof, st


if __name__ == "__main__":
    main()

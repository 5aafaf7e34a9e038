"""Tensor-level wrappers over the C ABI (one python function per C entry point).

Each wrapper only allocates the outputs/workspaces with torch (caching allocator)
and passes raw device pointers + the current torch stream to the library; no
arithmetic happens here.  The parity tests call these, so a green test means the
CUDA kernels produced the numbers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L
from ._lib import PAD_ROWS, check, dtype_code, load, ptr, require_device, stream_ptr


# kernels launched by each C entry point (our own kernels; used for bench.py's `gpu_launches`)
LAUNCHES = {"gate_fwd": 1, "gate_bwd": 3, "gate_bwd_dx": 1, "route_plan": 2, "dispatch_fwd": 1, "dispatch_bwd": 1,
            "combine_fwd": 1, "combine_bwd": 1, "cast_weights": 1, "ffn_fwd": 2, "ffn_bwd_f32": 6, "ffn_bwd_bf16": 4,
            "ln_stats": 1, "ln_fold_gate": 1, "ln_bwd_res": 2, "ffn_chain": 1}
launch_count = 0


def _count(name, n=1):
    global launch_count
    launch_count += LAUNCHES[name] * n


def _i32(n, dev):
    return torch.empty(n, dtype=torch.int32, device=dev)


def _f32(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


# ------------------------------------------------------------------------ router
@dataclass
class GateOut:
    idx: torch.Tensor          # [T,K] int64
    idx_full: torch.Tensor     # [T,K1] int32
    score: torch.Tensor        # [T,K]
    top_vals: torch.Tensor     # [T,K1]
    clean_logits: torch.Tensor  # [T,E]
    noisy_logits: torch.Tensor  # [T,E] (is clean_logits when no noise)
    gates: Optional[torch.Tensor]
    imp_partial: torch.Tensor
    load_partial: torch.Tensor


def gate_fwd(x, w_gate, top_k, task_feat=None, noise=None, noise_stddev=0.0, want_gates=False) -> GateOut:
    """x [T,D] fp32/bf16 (row-contiguous), w_gate [D+Dt,E] fp32, task_feat [Dt] fp32."""
    require_device(x)
    lib = load()
    T, D = x.shape
    Dg, E = w_gate.shape
    Dt = Dg - D
    assert x.stride(1) == 1 and w_gate.is_contiguous() and w_gate.dtype == torch.float32
    assert (Dt == 0) == (task_feat is None), "task_feat must be given iff w_gate has task rows"
    if task_feat is not None:
        task_feat = task_feat.reshape(-1).contiguous().float()
        assert task_feat.numel() == Dt
    dev = x.device
    K = top_k
    K1 = min(K + 1, E)
    n_part = lib.m3_gate_num_partials(T, E)
    if n_part < 0:
        check(n_part, "m3_gate_num_partials")
    idx = torch.empty(T, K, dtype=torch.int64, device=dev)
    idx_full = torch.empty(T, K1, dtype=torch.int32, device=dev)
    score = _f32((T, K), dev)
    top_vals = _f32((T, K1), dev)
    clean = _f32((T, E), dev)
    noisy = _f32((T, E), dev) if noise is not None else None
    gates = _f32((T, E), dev) if want_gates else None
    imp_p = _f32((max(n_part, 1), E), dev)
    load_p = torch.empty(max(n_part, 1), E, dtype=torch.int32, device=dev)
    if noise is not None and noise.dtype == torch.int64:
        # a generator state {seed, call counter} instead of a [T, E] tensor of normals: the kernel draws them itself
        assert noise.numel() == 2 and noise.is_contiguous()
        check(lib.m3_gate_fwd_rng(ptr(x), dtype_code(x), x.stride(0), ptr(task_feat), ptr(w_gate), ptr(noise),
                                  float(noise_stddev), T, D, Dt, E, K, ptr(idx), ptr(idx_full), ptr(score), ptr(top_vals),
                                  ptr(clean), ptr(noisy), ptr(gates), ptr(imp_p), ptr(load_p), stream_ptr()),
              "m3_gate_fwd_rng")
        _count("gate_fwd")
        return GateOut(idx, idx_full, score, top_vals, clean, noisy, gates, imp_p[:n_part], load_p[:n_part])
    if noise is not None:
        assert noise.shape == (T, E) and noise.dtype == torch.float32 and noise.is_contiguous()
    check(lib.m3_gate_fwd(ptr(x), dtype_code(x), x.stride(0), ptr(task_feat), ptr(w_gate), ptr(noise),
                          float(noise_stddev), T, D, Dt, E, K, ptr(idx), ptr(idx_full), ptr(score), ptr(top_vals),
                          ptr(clean), ptr(noisy), ptr(gates), ptr(imp_p), ptr(load_p), stream_ptr()), "m3_gate_fwd")
    _count("gate_fwd")
    return GateOut(idx, idx_full, score, top_vals, clean, noisy if noisy is not None else clean, gates,
                   imp_p[:n_part], load_p[:n_part])


def gate_bwd(x, w_gate, logits, idx_full, top_k, task_feat=None, dscore=None, dtop_vals=None, dgates=None,
             dimportance=None, dclean=None, dnoisy=None, want_dx_gate=False, importance=None, dcv_loss=None):
    """returns dz [T,E], dw_gate [Dg,E], dtask_feat [Dt] or None, dx_gate [T,D] or None"""
    require_device(x)
    lib = load()
    T, D = x.shape
    Dg, E = w_gate.shape
    Dt = Dg - D
    dev = x.device
    if task_feat is not None:
        task_feat = task_feat.reshape(-1).contiguous().float()

    def c(t):
        return None if t is None else t.contiguous().float()
    dscore, dtop_vals, dgates, dimportance, dclean, dnoisy = map(c, (dscore, dtop_vals, dgates, dimportance, dclean, dnoisy))
    if dcv_loss is not None:
        dcv_loss = dcv_loss.reshape(1).contiguous().float()
        assert importance is not None
    dz = _f32((T, E), dev)
    dw = _f32((Dg, E), dev)
    dtf = _f32((Dt,), dev) if Dt > 0 else None
    dxg = _f32((T, D), dev) if want_dx_gate else None
    nbytes = lib.m3_gate_bwd_workspace_bytes(T, D, Dt, E)
    ws = _ws(nbytes, dev)
    check(lib.m3_gate_bwd(ptr(x), dtype_code(x), x.stride(0), ptr(task_feat), ptr(w_gate), ptr(logits), ptr(idx_full),
                          T, D, Dt, E, top_k, ptr(dscore), ptr(dtop_vals), ptr(dgates), ptr(dimportance), ptr(dclean),
                          ptr(dnoisy), ptr(importance) if dcv_loss is not None else None, ptr(dcv_loss), ptr(dz),
                          ptr(dw), ptr(dtf), ptr(dxg), ptr(ws), ws.numel(), stream_ptr()),
          "m3_gate_bwd")
    _count("gate_bwd")
    if want_dx_gate:
        _count("gate_bwd_dx")
    return dz, dw, dtf, dxg


# -------------------------------------------------------------------- route plan
@dataclass
class Plan:
    counts: torch.Tensor        # [E] int32
    offsets: torch.Tensor       # [E+1] int32 (padded)
    pos: torch.Tensor           # [T*K] int32
    tile_expert: torch.Tensor   # [cap_rows / pad] int32
    cap_rows: int               # static queue capacity (rows)
    pad: int
    importance: Optional[torch.Tensor] = None
    load: Optional[torch.Tensor] = None
    cv_loss: Optional[torch.Tensor] = None      # 0-dim: cv^2(importance) + cv^2(load)


def route_plan(idx, num_expert, pad=PAD_ROWS, imp_partial=None, load_partial=None, inv_pos=None) -> Plan:
    """inv_pos: optional int32 [>= rows] buffer that receives the inverse map queue row -> slot (expert parallel)."""
    require_device(idx)
    lib = load()
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    T, K = idx.shape
    E = num_expert
    dev = idx.device
    cap_rows = lib.m3_route_max_rows(T, K, E, pad)
    counts, offsets, pos = _i32(E, dev), _i32(E + 1, dev), _i32(T * K, dev)
    tile_expert = _i32(max(cap_rows // pad, 1), dev) if pad >= 16 else None   # no tile map for pad-1 plans
    imp = load_v = cv = None
    n_part = 0
    if imp_partial is not None:
        n_part = imp_partial.shape[0]
        imp, load_v, cv = _f32((E,), dev), _f32((E,), dev), _f32((), dev)
    ws = _ws(lib.m3_route_plan_workspace_bytes(T, K, E), dev)
    check(lib.m3_route_plan(ptr(idx), T, K, E, pad, ptr(imp_partial), ptr(load_partial), n_part, ptr(counts),
                            ptr(offsets), ptr(pos), ptr(tile_expert), ptr(imp), ptr(load_v), ptr(cv), ptr(inv_pos),
                            ptr(ws), ws.numel(), stream_ptr()), "m3_route_plan")
    _count("route_plan")
    return Plan(counts, offsets, pos, tile_expert, cap_rows, pad, imp, load_v, cv)


# --------------------------------------------------------------- dispatch/combine
def dispatch_fwd(x, plan: Plan, top_k, out_dtype=None):
    require_device(x)
    T, D = x.shape
    x = x.contiguous()
    out_dtype = out_dtype or x.dtype
    xq = torch.empty(plan.cap_rows, D, dtype=out_dtype, device=x.device)
    E = plan.counts.numel()
    check(load().m3_dispatch_fwd(ptr(x), dtype_code(x), ptr(plan.pos), ptr(plan.counts), ptr(plan.offsets), T, top_k,
                                 D, E, ptr(xq), dtype_code(xq), stream_ptr()), "m3_dispatch_fwd")
    _count("dispatch_fwd")
    return xq


def dispatch_bwd(dxq, plan: Plan, T, top_k, out_dtype=torch.float32, dz=None, w_gate=None):
    require_device(dxq)
    D = dxq.shape[1]
    dx = torch.empty(T, D, dtype=out_dtype, device=dxq.device)
    E = w_gate.shape[1] if w_gate is not None else 0
    check(load().m3_dispatch_bwd(ptr(dxq), dtype_code(dxq), ptr(plan.pos), T, top_k, D, ptr(dz), ptr(w_gate), E,
                                 ptr(dx), dtype_code(dx), stream_ptr()), "m3_dispatch_bwd")
    _count("dispatch_bwd")
    return dx


def combine_fwd(yq, plan: Plan, score, out_dtype=torch.float32):
    require_device(yq)
    T, K = score.shape
    D = yq.shape[1]
    out = torch.empty(T, D, dtype=out_dtype, device=yq.device)
    check(load().m3_combine_fwd(ptr(yq), dtype_code(yq), ptr(plan.pos), ptr(score), T, K, D, ptr(out), dtype_code(out),
                                stream_ptr()), "m3_combine_fwd")
    _count("combine_fwd")
    return out


def combine_bwd(g, yq, plan: Plan, score):
    require_device(g)
    T, K = score.shape
    D = yq.shape[1]
    g = g.contiguous()
    dyq = torch.empty_like(yq)
    dscore = _f32((T, K), g.device)
    E = plan.counts.numel()
    check(load().m3_combine_bwd(ptr(g), dtype_code(g), ptr(yq), dtype_code(yq), ptr(plan.pos), ptr(score),
                                ptr(plan.counts), ptr(plan.offsets), T, K, D, E, ptr(dyq), dtype_code(dyq),
                                ptr(dscore), stream_ptr()), "m3_combine_bwd")
    _count("combine_bwd")
    return dyq, dscore


# -------------------------------------------------------------------- expert FFN
def cast_weights_bf16(w, want_plain=True, want_transposed=False, out=None):
    """fp32 [E,R,C] -> bf16 [E,R,C] and/or bf16 transposed [E,C,R]; `out` = (plain, transposed) buffers to overwrite"""
    require_device(w)
    E, R, Cc = w.shape
    w = w.contiguous()
    if out is not None:
        o, ot = out
        assert (o is None or (o.shape == (E, R, Cc) and o.dtype == torch.bfloat16 and o.is_contiguous()))
        assert (ot is None or (ot.shape == (E, Cc, R) and ot.dtype == torch.bfloat16 and ot.is_contiguous()))
    else:
        o = torch.empty(E, R, Cc, dtype=torch.bfloat16, device=w.device) if want_plain else None
        ot = torch.empty(E, Cc, R, dtype=torch.bfloat16, device=w.device) if want_transposed else None
    check(load().m3_cast_weights_bf16(ptr(w), E, R, Cc, ptr(o), ptr(ot), stream_ptr()), "m3_cast_weights_bf16")
    _count("cast_weights")
    return o, ot


def ffn_fwd(xq, plan: Plan, w1, b1, w2, b2, save_hpre=True, drop=None):
    """xq [cap,D] (fp32|bf16); w1 [E,H,D], w2 [E,D,H] same dtype as xq; b1,b2 fp32.
    Returns (yq, saved): `saved` is the library's opaque activation state for ffn_bwd (uint8).
    drop = (p, rng_state int64[2] = {seed, call counter} on the device): expert dropout behind the GELU."""
    require_device(xq)
    lib = load()
    cap, D = xq.shape
    E, H, _ = w1.shape
    dt = dtype_code(xq)
    assert w1.dtype == xq.dtype and w2.dtype == xq.dtype and b1.dtype == torch.float32
    # opaque activation state for the backward pass (fp32: pre-activation; bf16: z, or gelu'(z) and h planes)
    hpre = _ws(lib.m3_ffn_saved_bytes(dt, cap, D, H), xq.device) if save_hpre else None
    yq = torch.empty(cap, D, dtype=xq.dtype, device=xq.device)
    ws = _ws(lib.m3_ffn_workspace_bytes(dt, cap, D, H, E, 0), xq.device)
    if drop is not None and drop[0] > 0:
        assert save_hpre, "expert dropout is a training-time op"
        check(lib.m3_ffn_fwd_dropout(dt, ptr(xq), ptr(plan.offsets), ptr(plan.tile_expert), cap, E, D, H, ptr(w1), ptr(b1),
                                     ptr(w2), ptr(b2), ptr(hpre), ptr(yq), ptr(ws), ws.numel(), float(drop[0]), ptr(drop[1]),
                                     stream_ptr()), "m3_ffn_fwd_dropout")
    else:
        check(lib.m3_ffn_fwd(dt, ptr(xq), ptr(plan.offsets), ptr(plan.tile_expert), cap, E, D, H, ptr(w1), ptr(b1),
                             ptr(w2), ptr(b2), ptr(hpre), ptr(yq), ptr(ws), ws.numel(), stream_ptr()), "m3_ffn_fwd")
    _count("ffn_chain" if (not save_hpre and lib.m3_ffn_uses_chain(dt, D, H)) else "ffn_fwd")
    return yq, hpre


def ffn_bwd(xq, hpre, dyq, plan: Plan, w1, w2, w1t=None, w2t=None, drop=None):
    """returns dxq, dw1 [E,H,D] fp32, db1 [E,H], dw2 [E,D,H], db2 [E,D]; `drop` = what the forward call was given"""
    require_device(xq)
    lib = load()
    cap, D = xq.shape
    E, H, _ = w1.shape
    dt = dtype_code(xq)
    dev = xq.device
    dxq = torch.empty_like(xq)
    dw1, db1, dw2, db2 = _f32((E, H, D), dev), _f32((E, H), dev), _f32((E, D, H), dev), _f32((E, D), dev)
    ws = _ws(lib.m3_ffn_workspace_bytes(dt, cap, D, H, E, 1), dev)
    if drop is not None and drop[0] > 0:
        check(lib.m3_ffn_bwd_dropout(dt, ptr(xq), ptr(hpre), ptr(dyq), ptr(plan.counts), ptr(plan.offsets),
                                     ptr(plan.tile_expert), cap, E, D, H, ptr(w1), ptr(w2), ptr(w1t), ptr(w2t), ptr(dxq),
                                     ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(ws), ws.numel(), float(drop[0]),
                                     ptr(drop[1]), stream_ptr()), "m3_ffn_bwd_dropout")
    else:
        check(lib.m3_ffn_bwd(dt, ptr(xq), ptr(hpre), ptr(dyq), ptr(plan.counts), ptr(plan.offsets), ptr(plan.tile_expert),
                             cap, E, D, H, ptr(w1), ptr(w2), ptr(w1t), ptr(w2t), ptr(dxq), ptr(dw1), ptr(db1), ptr(dw2),
                             ptr(db2), ptr(ws), ws.numel(), stream_ptr()), "m3_ffn_bwd")
    if dt == L.M3_BF16:
        tiles = (D // 128) * (H // 128) * E          # mirrors wgrad_splits() in ffn_bf16.cu
        splits = min(16, max(1, 148 // tiles))
        _count("ffn_bwd_bf16")                       # 2 gg + 2 wgrad
        if splits > 1:
            _count("cast_weights", 4)                # 2 x (dW + db) split-K reduce launches
    else:
        _count("ffn_bwd_f32")
    return dxq, dw1, db1, dw2, db2


# ------------------------------------------------- Block-level fusion
@dataclass
class LnState:
    """LayerNorm(norm2) folded into the layer: statistics of the raw residual stream + folded router."""
    mean: torch.Tensor      # [T]
    rstd: torch.Tensor      # [T]
    gamma: torch.Tensor     # [D]
    beta: torch.Tensor      # [D]
    w_fold: torch.Tensor    # [Dg,E] gamma-scaled, column-centred router weights
    gb: torch.Tensor        # [2,E]  {gamma^T W, beta^T W}


def ln_prepare(x, gamma, beta, eps, w_gate) -> LnState:
    """x [T,D] raw fp32 residual stream; gamma/beta [D]; w_gate [Dg,E]."""
    require_device(x)
    lib = load()
    T, D = x.shape
    Dg, E = w_gate.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and w_gate.is_contiguous()
    gamma, beta = gamma.contiguous().float(), beta.contiguous().float()
    dev = x.device
    mean, rstd = _f32((T,), dev), _f32((T,), dev)
    w_fold, gb = _f32((Dg, E), dev), _f32((2, E), dev)
    check(lib.m3_ln_stats(ptr(x), T, D, float(eps), ptr(mean), ptr(rstd), stream_ptr()), "m3_ln_stats")
    check(lib.m3_ln_fold_gate(ptr(w_gate), ptr(gamma), ptr(beta), D, Dg, E, ptr(w_fold), ptr(gb), stream_ptr()),
          "m3_ln_fold_gate")
    _count("ln_stats")
    _count("ln_fold_gate")
    return LnState(mean, rstd, gamma, beta, w_fold, gb)


def gate_fwd_ln(x, ln: LnState, top_k, task_feat=None, noise=None, noise_stddev=0.0, want_gates=False) -> GateOut:
    """gate_fwd of LayerNorm(x) computed from the RAW x (never materialises the normalised tokens)."""
    require_device(x)
    lib = load()
    T, D = x.shape
    Dg, E = ln.w_fold.shape
    Dt = Dg - D
    assert x.dtype == torch.float32 and x.stride(1) == 1
    assert (Dt == 0) == (task_feat is None), "task_feat must be given iff w_gate has task rows"
    if task_feat is not None:
        task_feat = task_feat.reshape(-1).contiguous().float()
    dev = x.device
    K = top_k
    K1 = min(K + 1, E)
    n_part = lib.m3_gate_num_partials(T, E)
    if n_part < 0:
        check(n_part, "m3_gate_num_partials")
    idx = torch.empty(T, K, dtype=torch.int64, device=dev)
    idx_full = torch.empty(T, K1, dtype=torch.int32, device=dev)
    score, top_vals, clean = _f32((T, K), dev), _f32((T, K1), dev), _f32((T, E), dev)
    noisy = _f32((T, E), dev) if noise is not None else None
    gates = _f32((T, E), dev) if want_gates else None
    imp_p = _f32((max(n_part, 1), E), dev)
    load_p = torch.empty(max(n_part, 1), E, dtype=torch.int32, device=dev)
    check(lib.m3_gate_fwd_ln(ptr(x), x.stride(0), ptr(ln.mean), ptr(ln.rstd), ptr(ln.gb), ptr(task_feat),
                             ptr(ln.w_fold), ptr(noise), float(noise_stddev), T, D, Dt, E, K, ptr(idx), ptr(idx_full),
                             ptr(score), ptr(top_vals), ptr(clean), ptr(noisy), ptr(gates), ptr(imp_p), ptr(load_p),
                             stream_ptr()), "m3_gate_fwd_ln")
    _count("gate_fwd")
    return GateOut(idx, idx_full, score, top_vals, clean, noisy if noisy is not None else clean, gates,
                   imp_p[:n_part], load_p[:n_part])


def dispatch_fwd_ln(x, ln: LnState, plan: Plan, top_k, out_dtype=torch.float32):
    require_device(x)
    T, D = x.shape
    xq = torch.empty(plan.cap_rows, D, dtype=out_dtype, device=x.device)
    E = plan.counts.numel()
    check(load().m3_dispatch_fwd_ln(ptr(x), ptr(ln.mean), ptr(ln.rstd), ptr(ln.gamma), ptr(ln.beta), ptr(plan.pos),
                                    ptr(plan.counts), ptr(plan.offsets), T, top_k, D, E, ptr(xq), dtype_code(xq),
                                    stream_ptr()), "m3_dispatch_fwd_ln")
    _count("dispatch_fwd")
    return xq


def combine_fwd_res(yq, plan: Plan, score, residual):
    require_device(yq)
    T, K = score.shape
    D = yq.shape[1]
    assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == (T, D)
    out = _f32((T, D), yq.device)
    check(load().m3_combine_fwd_res(ptr(yq), dtype_code(yq), ptr(plan.pos), ptr(score), ptr(residual), T, K, D,
                                    ptr(out), stream_ptr()), "m3_combine_fwd_res")
    _count("combine_fwd")
    return out


def gate_bwd_ln(x, ln: LnState, w_gate, logits, idx_full, top_k, task_feat=None, dscore=None, dtop_vals=None,
                dgates=None, dimportance=None, dclean=None, dnoisy=None, importance=None, dcv_loss=None):
    """gate_bwd with x normalised on load.  returns dz [T,E], dw_gate [Dg,E], dtask_feat [Dt] or None"""
    require_device(x)
    lib = load()
    T, D = x.shape
    Dg, E = w_gate.shape
    Dt = Dg - D
    dev = x.device
    if task_feat is not None:
        task_feat = task_feat.reshape(-1).contiguous().float()

    def c(t):
        return None if t is None else t.contiguous().float()
    dscore, dtop_vals, dgates, dimportance, dclean, dnoisy = map(c, (dscore, dtop_vals, dgates, dimportance, dclean, dnoisy))
    if dcv_loss is not None:
        dcv_loss = dcv_loss.reshape(1).contiguous().float()
        assert importance is not None
    dz, dw = _f32((T, E), dev), _f32((Dg, E), dev)
    dtf = _f32((Dt,), dev) if Dt > 0 else None
    ws = _ws(lib.m3_gate_bwd_workspace_bytes(T, D, Dt, E), dev)
    check(lib.m3_gate_bwd_ln(ptr(x), x.stride(0), ptr(ln.mean), ptr(ln.rstd), ptr(ln.gamma), ptr(ln.beta),
                             ptr(task_feat), ptr(w_gate), ptr(logits), ptr(idx_full), T, D, Dt, E, top_k, ptr(dscore),
                             ptr(dtop_vals), ptr(dgates), ptr(dimportance), ptr(dclean), ptr(dnoisy),
                             ptr(importance) if dcv_loss is not None else None, ptr(dcv_loss), ptr(dz), ptr(dw),
                             ptr(dtf), ptr(ws), ws.numel(), stream_ptr()), "m3_gate_bwd_ln")
    _count("gate_bwd")
    return dz, dw, dtf


def ln_bwd_res(dxn, x, ln: LnState, dres):
    """dx = dres + LayerNorm'(dxn) wrt the raw x;  returns dx [T,D], dgamma [D], dbeta [D]"""
    require_device(dxn)
    lib = load()
    T, D = x.shape
    dev = x.device
    dxn, dres = dxn.contiguous(), dres.contiguous()
    assert dxn.dtype == torch.float32 and dres.dtype == torch.float32
    dx, dgamma, dbeta = _f32((T, D), dev), _f32((D,), dev), _f32((D,), dev)
    ws = _ws(lib.m3_ln_bwd_workspace_bytes(T, D), dev)
    check(lib.m3_ln_bwd_res(ptr(dxn), ptr(x), ptr(ln.mean), ptr(ln.rstd), ptr(ln.gamma), ptr(dres), T, D, ptr(dx),
                            ptr(dgamma), ptr(dbeta), ptr(ws), ws.numel(), stream_ptr()), "m3_ln_bwd_res")
    _count("ln_bwd_res")
    return dx, dgamma, dbeta

"""Autograd wiring of the MoE hot path over the C-ABI ops (m3vit_b200/ops.py).

`MoEFunction` is the whole layer of the reference in one autograd node
(/root/reference/models/moe/origin/custom_moe_layer.py:184-314):

    gate -> route plan -> dispatch -> expert FFN -> combine

replacing fmoe's prepare_forward / MOEScatter / MOELinear x2 / MOEGather
autograd Functions and the torch.bmm.  `GateFunction` and `ExpertsFunction` are
the same ops split at the routing boundary, for callers that compute or consume
routing on their own (NoisyGate_VMoE used stand-alone; the token-MoE experts-only
entry, models/moe/token/custom_moe_layer.py:88-156).

Nothing here touches the host: no .item(), no .cpu() - the reference's
per-layer D2H sync (fmoe prepare_forward) is gone.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import PAD_ROWS


def _c(t):
    return None if t is None else t.contiguous()


class MoEFunction(torch.autograd.Function):
    """inputs : x[T,D], gate_x[T,Dg0] or None (=x), w_gate[Dg,E], task_feat[Dt]|None,
                w1[E,H,D], b1[E,H], w2[E,D,H], b2[E,D] (fp32 masters),
                noise[T,E]|None, then non-tensor config.
       outputs: out[T,D], score[T,K], top_vals[T,K1], clean[T,E], noisy[T,E],
                gates[T,E]|empty, importance[E], load[E], idx[T,K], counts[E],
                cv_loss[] = cv^2(importance) + cv^2(load) (fused, differentiable)"""

    @staticmethod
    def forward(ctx, x, gate_x, w_gate, task_feat, w1, b1, w2, b2, noise, top_k, noise_stddev, compute_dtype,
                want_gates, wcache, drop=None):
        ctx.set_materialize_grads(False)      # undefined output grads stay None (no zero fills)
        T, D = x.shape
        E, H, _ = w1.shape
        x = _c(x)
        gx = x if gate_x is None else _c(gate_x)
        g = ops.gate_fwd(gx, w_gate, top_k, task_feat, noise, noise_stddev, want_gates)
        plan = ops.route_plan(g.idx, E, PAD_ROWS, g.imp_partial, g.load_partial)
        if compute_dtype == torch.bfloat16:
            w1c, w2c, w1t, w2t = wcache.get_bf16(w1, w2)
        else:
            w1c, w2c, w1t, w2t = w1, w2, None, None
        needs_grad = any(ctx.needs_input_grad)
        xq = ops.dispatch_fwd(x, plan, top_k, out_dtype=compute_dtype)
        yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, save_hpre=needs_grad, drop=drop if needs_grad else None)
        out = ops.combine_fwd(yq, plan, g.score, out_dtype=x.dtype)
        if needs_grad:
            ctx.save_for_backward(x, gate_x, w_gate, task_feat, w1c, w2c, w1t, w2t, xq, hpre, yq, g.score,
                                  g.noisy_logits, g.idx_full, plan.counts, plan.offsets, plan.pos, plan.tile_expert,
                                  plan.importance)
            ctx.cfg = (top_k, plan.cap_rows, gate_x is not None)
            ctx.drop = drop
        gates = g.gates if g.gates is not None else x.new_empty(0)
        ctx.mark_non_differentiable(g.idx, plan.load, plan.counts)
        if noise is None:
            # clean and noisy logits are the same tensor: hand out one differentiable view each
            noisy = g.clean_logits.view_as(g.clean_logits)
        else:
            noisy = g.noisy_logits
        return (out, g.score, g.top_vals, g.clean_logits, noisy, gates, plan.importance, plan.load, g.idx,
                plan.counts, plan.cv_loss)

    @staticmethod
    def backward(ctx, d_out, d_score, d_top, d_clean, d_noisy, d_gates, d_imp, _dl, _di, _dc, d_cv):
        (x, gate_x, w_gate, task_feat, w1c, w2c, w1t, w2t, xq, hpre, yq, score, logits, idx_full, counts, offsets,
         pos, tile_expert, importance) = ctx.saved_tensors
        top_k, cap_rows, separate_gate_inp = ctx.cfg
        T, D = x.shape
        plan = ops.Plan(counts, offsets, pos, tile_expert, cap_rows, PAD_ROWS)
        if d_out is None:
            d_out = torch.zeros_like(x)
        dyq, dscore = ops.combine_bwd(_c(d_out), yq, plan, score)
        dxq, dw1, db1, dw2, db2 = ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t, drop=ctx.drop)
        if d_score is not None:
            dscore = dscore + d_score
        if d_gates is not None and d_gates.numel() == 0:
            d_gates = None
        gx = x if gate_x is None else gate_x
        dz, dwg, dtf, dxg = ops.gate_bwd(gx, w_gate, logits, idx_full, top_k, task_feat, dscore, d_top, d_gates,
                                         d_imp, d_clean, d_noisy, want_dx_gate=separate_gate_inp,
                                         importance=importance, dcv_loss=d_cv)
        if separate_gate_inp:
            dx = ops.dispatch_bwd(dxq, plan, T, top_k, out_dtype=x.dtype)
            dgx = dxg.to(gate_x.dtype)
        else:
            dx = ops.dispatch_bwd(dxq, plan, T, top_k, out_dtype=x.dtype, dz=dz, w_gate=w_gate)
            dgx = None
        if dtf is not None and task_feat is not None:
            dtf = dtf.view_as(task_feat).to(task_feat.dtype)
        return dx, dgx, dwg, dtf, dw1, db1, dw2, db2, None, None, None, None, None, None, None


class MoEBlockFunction(torch.autograd.Function):
    """out = x + MoE(LayerNorm(x)): the MoE half of a reference Block
    (/root/reference/models/moe/origin/vision_transformer_moe.py:278-283) with norm2 and the residual add
    fused into the layer's kernels - the normalised tokens are never materialised.

       inputs : x[T,D] RAW fp32 residual stream, ln_w[D], ln_b[D], then as MoEFunction (gate input == LN(x))
       outputs: as MoEFunction (out already contains the residual)"""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w_gate, task_feat, w1, b1, w2, b2, noise, eps, top_k, noise_stddev,
                compute_dtype, want_gates, wcache, drop=None):
        ctx.set_materialize_grads(False)      # undefined output grads stay None (no zero fills)
        T, D = x.shape
        E = w1.shape[0]
        x = _c(x)
        ln = ops.ln_prepare(x, ln_w.detach(), ln_b.detach(), eps, w_gate.detach())
        g = ops.gate_fwd_ln(x, ln, top_k, task_feat, noise, noise_stddev, want_gates)
        plan = ops.route_plan(g.idx, E, PAD_ROWS, g.imp_partial, g.load_partial)
        xq = ops.dispatch_fwd_ln(x, ln, plan, top_k, out_dtype=compute_dtype)
        if compute_dtype == torch.bfloat16:
            w1c, w2c, w1t, w2t = wcache.get_bf16(w1, w2)
        else:
            w1c, w2c, w1t, w2t = w1, w2, None, None
        needs_grad = any(ctx.needs_input_grad)
        yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, save_hpre=needs_grad, drop=drop if needs_grad else None)
        out = ops.combine_fwd_res(yq, plan, g.score, x)
        if needs_grad:
            ctx.save_for_backward(x, ln.mean, ln.rstd, ln.gamma, ln.beta, w_gate, task_feat, w1c, w2c, w1t, w2t, xq,
                                  hpre, yq, g.score, g.noisy_logits, g.idx_full, plan.counts, plan.offsets, plan.pos,
                                  plan.tile_expert, plan.importance)
            ctx.cfg = (top_k, plan.cap_rows)
            ctx.drop = drop
        gates = g.gates if g.gates is not None else x.new_empty(0)
        ctx.mark_non_differentiable(g.idx, plan.load, plan.counts)
        noisy = g.clean_logits.view_as(g.clean_logits) if noise is None else g.noisy_logits
        return (out, g.score, g.top_vals, g.clean_logits, noisy, gates, plan.importance, plan.load, g.idx,
                plan.counts, plan.cv_loss)

    @staticmethod
    def backward(ctx, d_out, d_score, d_top, d_clean, d_noisy, d_gates, d_imp, _dl, _di, _dc, d_cv):
        (x, mean, rstd, gamma, beta, w_gate, task_feat, w1c, w2c, w1t, w2t, xq, hpre, yq, score, logits, idx_full,
         counts, offsets, pos, tile_expert, importance) = ctx.saved_tensors
        top_k, cap_rows = ctx.cfg
        T, D = x.shape
        plan = ops.Plan(counts, offsets, pos, tile_expert, cap_rows, PAD_ROWS)
        ln = ops.LnState(mean, rstd, gamma, beta, None, None)
        d_out = torch.zeros_like(x) if d_out is None else _c(d_out).float()
        dyq, dscore = ops.combine_bwd(d_out, yq, plan, score)
        dxq, dw1, db1, dw2, db2 = ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t, drop=ctx.drop)
        if d_score is not None:
            dscore = dscore + d_score
        if d_gates is not None and d_gates.numel() == 0:
            d_gates = None
        dz, dwg, dtf = ops.gate_bwd_ln(x, ln, w_gate, logits, idx_full, top_k, task_feat, dscore, d_top, d_gates,
                                       d_imp, d_clean, d_noisy, importance=importance, dcv_loss=d_cv)
        dxn = ops.dispatch_bwd(dxq, plan, T, top_k, out_dtype=torch.float32, dz=dz, w_gate=w_gate)
        dx, dgamma, dbeta = ops.ln_bwd_res(dxn, x, ln, d_out)
        if dtf is not None and task_feat is not None:
            dtf = dtf.view_as(task_feat).to(task_feat.dtype)
        return dx, dgamma, dbeta, dwg, dtf, dw1, db1, dw2, db2, None, None, None, None, None, None, None, None


class GateFunction(torch.autograd.Function):
    """The router alone (NoisyGate_VMoE.forward).  outputs as MoEFunction minus `out`."""

    @staticmethod
    def forward(ctx, gx, w_gate, task_feat, noise, top_k, noise_stddev, want_gates):
        ctx.set_materialize_grads(False)      # undefined output grads stay None (no zero fills)
        gx = _c(gx)
        E = w_gate.shape[1]
        g = ops.gate_fwd(gx, w_gate, top_k, task_feat, noise, noise_stddev, want_gates)
        plan = ops.route_plan(g.idx, E, PAD_ROWS, g.imp_partial, g.load_partial)
        ctx.save_for_backward(gx, w_gate, task_feat, g.noisy_logits, g.idx_full, plan.importance)
        ctx.top_k = top_k
        gates = g.gates if g.gates is not None else gx.new_empty(0)
        ctx.mark_non_differentiable(g.idx, plan.load, plan.counts, plan.offsets, plan.pos, plan.tile_expert)
        noisy = g.clean_logits.view_as(g.clean_logits) if noise is None else g.noisy_logits
        return (g.score, g.top_vals, g.clean_logits, noisy, gates, plan.importance, plan.cv_loss, plan.load, g.idx,
                plan.counts, plan.offsets, plan.pos, plan.tile_expert)

    @staticmethod
    def backward(ctx, d_score, d_top, d_clean, d_noisy, d_gates, d_imp, d_cv, *_):
        gx, w_gate, task_feat, logits, idx_full, importance = ctx.saved_tensors
        if d_gates is not None and d_gates.numel() == 0:
            d_gates = None
        dz, dwg, dtf, dxg = ops.gate_bwd(gx, w_gate, logits, idx_full, ctx.top_k, task_feat, d_score, d_top, d_gates,
                                         d_imp, d_clean, d_noisy, want_dx_gate=True, importance=importance,
                                         dcv_loss=d_cv)
        if dtf is not None and task_feat is not None:
            dtf = dtf.view_as(task_feat).to(task_feat.dtype)
        return dxg.to(gx.dtype), dwg, dtf, None, None, None, None


class ExpertsFunction(torch.autograd.Function):
    """dispatch -> expert FFN -> combine for externally supplied routing
    (idx[T,K] int64, score[T,K]); the token-MoE entry of the reference
    (models/moe/token/custom_moe_layer.py:88-156)."""

    @staticmethod
    def forward(ctx, x, idx, score, w1, b1, w2, b2, compute_dtype, wcache, plan: Optional[ops.Plan], drop=None):
        ctx.set_materialize_grads(False)      # undefined output grads stay None (no zero fills)
        T, D = x.shape
        E = w1.shape[0]
        K = idx.shape[1]
        x = _c(x)
        score = _c(score).float()
        if plan is None:
            plan = ops.route_plan(_c(idx), E, PAD_ROWS)
        xq = ops.dispatch_fwd(x, plan, K, out_dtype=compute_dtype)
        if compute_dtype == torch.bfloat16:
            w1c, w2c, w1t, w2t = wcache.get_bf16(w1, w2)
        else:
            w1c, w2c, w1t, w2t = w1, w2, None, None
        needs_grad = any(ctx.needs_input_grad)
        yq, hpre = ops.ffn_fwd(xq, plan, w1c, b1, w2c, b2, save_hpre=needs_grad, drop=drop if needs_grad else None)
        out = ops.combine_fwd(yq, plan, score, out_dtype=x.dtype)
        if needs_grad:
            ctx.save_for_backward(x, w1c, w2c, w1t, w2t, xq, hpre, yq, score, plan.counts, plan.offsets, plan.pos,
                                  plan.tile_expert)
            ctx.cfg = (K, plan.cap_rows)
            ctx.drop = drop
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, w1c, w2c, w1t, w2t, xq, hpre, yq, score, counts, offsets, pos, tile_expert = ctx.saved_tensors
        K, cap_rows = ctx.cfg
        plan = ops.Plan(counts, offsets, pos, tile_expert, cap_rows, PAD_ROWS)
        if d_out is None:
            d_out = torch.zeros_like(x)
        dyq, dscore = ops.combine_bwd(_c(d_out), yq, plan, score)
        dxq, dw1, db1, dw2, db2 = ops.ffn_bwd(xq, hpre, dyq, plan, w1c, w2c, w1t, w2t, drop=ctx.drop)
        dx = ops.dispatch_bwd(dxq, plan, x.shape[0], K, out_dtype=x.dtype)
        return dx, None, dscore, dw1, db1, dw2, db2, None, None, None, None


class WeightCache:
    """bf16 (and transposed bf16) copies of the fp32 expert weights.

    Eager mode: re-cast only when a parameter changed - optimizer steps and load_state_dict bump `_version`; in-place
    updates through `param.data` do NOT (`p.data.add_(..)`): call `invalidate()` after those.
    Under CUDA-graph capture (`torch.cuda.graph`, `make_graphed_callables`): the casts are ALWAYS launched, into the
    cache's persistent buffers, so that they are part of the captured graph and every replay re-reads the current fp32
    masters - a graph captured with a warm cache would otherwise keep training on the weights it was captured with."""

    def __init__(self):
        self._key = None
        self._val = None
        self._graph_val = None      # persistent buffers the captured graphs read (never replaced while the cache lives)
        self._graph_shapes = None

    def invalidate(self):
        self._key = None

    def get_bf16(self, w1, w2):
        if torch.cuda.is_current_stream_capturing():
            shapes = (w1.device, tuple(w1.shape), tuple(w2.shape))
            if self._graph_val is None or self._graph_shapes != shapes:
                self._graph_val = (torch.empty_like(w1, dtype=torch.bfloat16), torch.empty_like(w2, dtype=torch.bfloat16),
                                   torch.empty(w1.shape[0], w1.shape[2], w1.shape[1], dtype=torch.bfloat16, device=w1.device),
                                   torch.empty(w2.shape[0], w2.shape[2], w2.shape[1], dtype=torch.bfloat16, device=w2.device))
                self._graph_shapes = shapes
            w1c, w2c, w1t, w2t = self._graph_val
            ops.cast_weights_bf16(w1.detach(), out=(w1c, w1t))       # captured: every replay re-reads the fp32 masters
            ops.cast_weights_bf16(w2.detach(), out=(w2c, w2t))
            return self._graph_val
        key = (w1.data_ptr(), w1._version, w2.data_ptr(), w2._version, w1.device)
        if key != self._key:
            w1c, w1t = ops.cast_weights_bf16(w1.detach(), True, True)
            w2c, w2t = ops.cast_weights_bf16(w2.detach(), True, True)
            self._key, self._val = key, (w1c, w2c, w1t, w2t)
        return self._val

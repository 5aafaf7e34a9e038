/*
 * m3vit_moe.h -- C ABI of the B200-native M3ViT MoE-layer hot path.
 *
 * The reference (aapdo/M3ViT) has no native code and no FFI of its own: its MoE
 * layer (models/moe/origin/custom_moe_layer.py:161-314) reaches the arithmetic
 * through the third-party FastMoE python package (`fmoe`, pinned @4edeccd,
 * README.md:42-50), whose `fmoe_cuda` pybind ops are the seams this ABI replaces.
 * Every entry point below names the reference call site / fmoe op it stands in for.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  All pointers are DEVICE pointers
 *     unless stated otherwise.  The library never allocates, frees or retains
 *     device memory: every buffer (outputs, workspaces) is caller-owned.
 *   - every launch goes to the `stream` passed in (a cudaStream_t); no implicit
 *     device synchronisation, no host read-back.  The compute entry points keep no state between
 *     calls and may be called from several host threads on different streams.  The only
 *     process-wide state is the DIAGNOSIS interface at the end of this header (m3_set_knob,
 *     m3_set_gemm_sm_limit, m3_debug_trace_buffer): tuning / measurement switches read at launch
 *     time, never needed for results (every knob setting is bit-identical or documented as
 *     "measurement only") and not meant to be changed while other threads are launching.
 *   - return value: 0 ok; <0 argument error (m3_status); >0 a cudaError_t.
 *   - activations may be fp32 or bf16 (m3_dtype); router math is always fp32.
 *   - expert queues use the PADDED layout: expert e owns rows
 *     [offsets[e], offsets[e]+counts[e]) of the queue buffer, offsets[] is a
 *     multiple of `pad` (M3_PAD_ROWS for the tensor-core path); padding rows are zero.
 *
 * Symbols T tokens, K top-k, E total experts, D model dim, Dt task-feature dim
 * (0 if none), Dg = D + Dt router input dim, H expert hidden dim, K1 = min(K+1,E).
 */
#ifndef M3VIT_MOE_H_
#define M3VIT_MOE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3_ABI_VERSION 2
#define M3_PAD_ROWS 256 /* queue padding = M-tile of a CTA pair (cta_group::2) in the grouped GEMM */

typedef void* m3_stream_t; /* cudaStream_t */

typedef enum { M3_F32 = 0, M3_BF16 = 1 } m3_dtype;

typedef enum {
  M3_OK = 0,
  M3_ERR_ARG = -1,         /* null pointer / negative size */
  M3_ERR_SHAPE = -2,       /* unsupported shape (see each function) */
  M3_ERR_ALIGN = -3,       /* pointer or leading dimension not 16-byte aligned */
  M3_ERR_UNSUPPORTED = -4, /* dtype / feature not implemented */
  M3_ERR_DEVICE = -5,      /* current device is not sm_100 */
  M3_ERR_WORKSPACE = -6    /* workspace too small */
} m3_status;

int m3_abi_version(void);
const char* m3_status_string(int status);
/* 0 if the current CUDA device can run this library (compute capability 10.x). */
int m3_check_device(void);

/* ------------------------------------------------------------------ router --
 * Fused noisy_vmoe gate.  Replaces NoisyGate_VMoE.forward
 * (models/moe/origin/noisy_gate_vmoe.py:168-297): logits = [x, task_feat] @ w_gate,
 * (+ noise * noise_stddev), softmax over ALL E, top-K1 on the probabilities,
 * first K kept un-renormalised.  One kernel; per-CTA partial importance / load.
 *
 *   x            [T, D]  (x_dtype), row stride ldx elements
 *   task_feat    [Dt] fp32 or NULL (Dt == 0)       (custom_moe_layer.py:176-179)
 *   w_gate       [D+Dt, E] fp32
 *   noise        [T, E] fp32 standard-normal or NULL (eval / std 0)
 *   idx          [T, K]  int64  (what fmoe's prepare_forward consumes)
 *   idx_full     [T, K1] int32  (saved for backward)
 *   score        [T, K]  fp32;  top_vals [T, K1] fp32
 *   clean_logits [T, E] fp32;   noisy_logits [T, E] fp32 or NULL (then == clean)
 *   gates        [T, E] fp32 dense scatter(idx, score) or NULL
 *   imp_partial  [m3_gate_num_partials, E] fp32, load_partial same shape int32
 * Supported: E in {4,8,16,32,64,128}; D % 32 == 0; 1 <= K <= min(E, 8).
 */
int m3_gate_num_partials(int T, int E);
int m3_gate_fwd(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                const float* noise, float noise_stddev, int T, int D, int Dt, int E, int K,
                int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean_logits,
                float* noisy_logits, float* gates, float* imp_partial, int32_t* load_partial,
                m3_stream_t stream);
/* m3_gate_fwd with the router noise drawn IN the kernel (noisy_gate_vmoe.py:226 draws it with torch.randn_like):
 * rng_state = {uint64 seed, uint64 call counter} in device memory, 16-byte aligned (csrc/philox.cuh: Philox4x32-7 +
 * Box-Muller, four normals per lane); the [T, E] noise tensor is never materialised.  Not torch's stream: statistical
 * parity.  noisy_logits is required. */
int m3_gate_fwd_rng(const void* x, int x_dtype, int64_t ldx, const float* task_feat,
                    const float* w_gate, const void* rng_state, float noise_stddev, int T, int D, int Dt,
                    int E, int K, int64_t* idx, int32_t* idx_full, float* score, float* top_vals,
                    float* clean_logits, float* noisy_logits, float* gates, float* imp_partial,
                    int32_t* load_partial, m3_stream_t stream);

/* Router backward (autograd through noisy_gate_vmoe.py:179-265): softmax
 * Jacobian over all E from the gradients of every differentiable gate output,
 * then dw_gate = [x,tf]^T dz and dtask_feat.  dz is written for m3_dispatch_bwd,
 * which adds dz @ w_gate[:D]^T into dx.   Any of the d* inputs may be NULL (= 0).
 *   logits [T,E] = the (noisy) logits the forward soft-maxed
 *   importance[E] + dcv_loss[1] (both or neither): gradient of m3_route_plan's cv_loss,
 *            chained analytically through cv^2(importance) (the load term is piecewise constant)
 *   dz [T,E] out;  dw_gate [D+Dt, E] out (overwritten);  dtask_feat [Dt] out or NULL
 *   dx_gate [T,D] fp32 out or NULL: dz @ w_gate[:D]^T, for callers whose gate input
 *            is not the layer input (Block.gate_input_ahead)
 */
size_t m3_gate_bwd_workspace_bytes(int T, int D, int Dt, int E);
int m3_gate_bwd(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                const float* logits, const int32_t* idx_full, int T, int D, int Dt, int E, int K,
                const float* dscore, const float* dtop_vals, const float* dgates,
                const float* dimportance, const float* dclean, const float* dnoisy,
                const float* importance, const float* dcv_loss, float* dz, float* dw_gate,
                float* dtask_feat, float* dx_gate, void* workspace, size_t workspace_bytes,
                m3_stream_t stream);

/* -------------------------------------------------------------- route plan --
 * Replaces fmoe_cuda.expert_count + assign_pos (+ the host-side cumsum and the
 * D2H sync of fmoe.functions.prepare_forward; reference call site
 * custom_moe_layer.py:255-257).  Deterministic and stable: rows of one expert
 * keep flat-slot order.  Everything stays on the device.
 *   idx [T,K] int64 -> counts[E], offsets[E+1] (padded exclusive prefix),
 *   pos[T*K] (slot t*K+k -> queue row), tile_expert[offsets[E]/pad] (expert of
 *   every pad-row tile; capacity m3_route_max_tiles), and, if the partials of
 *   m3_gate_fwd are passed, importance[E] / load[E] fp32 (fixed-order sums) and
 *   cv_loss[1] = cv^2(importance) + cv^2(load), cv^2(u) = var_unbiased(u)/(mean(u)^2+1e-10)
 *   (noisy_gate_vmoe.py:127-141,278-283; the hard-count load of the noise-free path).
 *   inv_pos[offsets[E]] (may be NULL): the inverse map, queue row -> slot (rows of padding keep their old contents);
 *   with pad = 1 this is the slot list sorted by expert, the send order of the expert-parallel push.
 */
size_t m3_route_plan_workspace_bytes(int T, int K, int E);
int m3_route_max_rows(int T, int K, int E, int pad);   /* queue capacity in rows  */
int m3_route_max_tiles(int T, int K, int E, int pad);  /* = max_rows / pad        */
int m3_route_plan(const int64_t* idx, int T, int K, int E, int pad, const float* imp_partial,
                  const int32_t* load_partial, int n_partial, int32_t* counts, int32_t* offsets,
                  int32_t* pos, int32_t* tile_expert, float* importance, float* load,
                  float* cv_loss, int32_t* inv_pos, void* workspace, size_t workspace_bytes,
                  m3_stream_t stream);

/* -------------------------------------------------------- dispatch/combine --
 * HBM-bound row movers, 128-bit vectorised.  D % 8 == 0.
 * m3_dispatch_fwd : MOEScatter.forward (index_select by pos//K): xq[pos[t,k]] = x[t],
 *                   with dtype cast; zeroes the padding rows of every queue.
 * m3_dispatch_bwd : MOEScatter.backward (index_add): dx[t] = sum_k dxq[pos[t,k]]
 *                   (+ dz[t] @ w_gate[:D]^T when dz != NULL: router dx, gate_inp is inp)
 * m3_combine_fwd  : MOEGather.forward + torch.bmm (custom_moe_layer.py:283-297):
 *                   out[t] = sum_k score[t,k] * yq[pos[t,k]], fp32 accumulation in k order
 * m3_combine_bwd  : their backward: dscore[t,k] = <g[t], yq[pos[t,k]]>,
 *                   dyq[pos[t,k]] = score[t,k] * g[t]; zeroes dyq padding rows.
 */
int m3_dispatch_fwd(const void* x, int x_dtype, const int32_t* pos, const int32_t* counts,
                    const int32_t* offsets, int T, int K, int D, int E, void* xq, int xq_dtype,
                    m3_stream_t stream);
int m3_dispatch_bwd(const void* dxq, int dxq_dtype, const int32_t* pos, int T, int K, int D,
                    const float* dz, const float* w_gate, int E, void* dx, int dx_dtype,
                    m3_stream_t stream);
int m3_combine_fwd(const void* yq, int yq_dtype, const int32_t* pos, const float* score, int T,
                   int K, int D, void* out, int out_dtype, m3_stream_t stream);
int m3_combine_bwd(const void* g, int g_dtype, const void* yq, int yq_dtype, const int32_t* pos,
                   const float* score, const int32_t* counts, const int32_t* offsets, int T, int K,
                   int D, int E, void* dyq, int dyq_dtype, float* dscore, m3_stream_t stream);

/* --------------------------------------------------------------- expert FFN --
 * Replaces _Expert.forward (custom_moe_layer.py:36-44) = FMoELinear -> GELU(erf)
 * -> FMoELinear, i.e. fmoe_cuda.linear_forward/backward's per-expert cuBLAS loop,
 * as a grouped GEMM over the padded expert queues.
 *   dtype M3_F32 : fp32 SIMT path (parity mode; the reference trains in fp32)
 *   dtype M3_BF16: tcgen05/TMEM/TMA path (bf16 operands, fp32 accumulation)
 *   xq [rows, D], yq [rows, D]; rows = offsets[E] (<= cap_rows); w1 [E,H,D], w2 [E,D,H] in
 *   `dtype`; b1 [E,H], b2 [E,D] fp32.   tile_expert / offsets from m3_route_plan (pad M3_PAD_ROWS).
 *   saved: OPAQUE activation state of m3_ffn_saved_bytes(dtype, cap_rows, D, H) bytes that m3_ffn_fwd
 *   fills and m3_ffn_bwd consumes (NULL when not training).  fp32: the pre-activation z [rows, H];
 *   bf16: gelu'(z) and h = gelu(z) as two [rows, H] planes, so that the backward GEMM epilogue is a single multiply
 *   and h is not recomputed (autograd would keep z and re-evaluate erf/exp).
 *   bf16 backward additionally needs transposed weight copies w1t [E,D,H], w2t [E,H,D].
 *   Weight / bias gradients are fp32 and OVERWRITTEN (caller accumulates).
 */
size_t m3_ffn_workspace_bytes(int dtype, int cap_rows, int D, int H, int E, int backward);
size_t m3_ffn_saved_bytes(int dtype, int cap_rows, int D, int H);
/* 1 if m3_ffn_fwd with saved == NULL runs this (dtype, D, H) as the chain kernel (launch accounting, tests). */
int m3_ffn_uses_chain(int dtype, int D, int H);
/* Process-wide tuning knob: SMs the persistent tcgen05 GEMMs may occupy (default / out of range: all 148), for callers
 * that run other kernels beside them on another stream. */
int m3_set_gemm_sm_limit(int sms);
/* Process-wide tuning knobs (A/B measurement; defaults are the shipped configuration).
 *   M3_KNOB_PDL       1: kernels are launched with programmatic stream serialisation and start their
 *                     prologue (barrier init, TMEM allocation, descriptor prefetch) under the tail of the
 *                     previous kernel; every kernel executes griddepcontrol.wait before its first global access.
 *                     0 (default): plain stream-ordered launches (at bench size the kernels are 30-100 us long and
 *                     PDL measured no gain; it is meant for the launch-bound small-batch regime).
 *   M3_KNOB_EPI_WARPS 8 or 16 epilogue warps in the tcgen05 grouped GEMM (0 = per-epilogue default;
 *                     0x100 | mask: bit e of mask set -> 16 warps for epilogue e = 0 store, 1 bias, 2 fc1, 3 dgelu;
 *                     | 0x200: the 8-warp epilogues use 32-column register blocks / 2 KB staging boxes (one more smem
 *                     stage) instead of 64 / 4 KB).
 *   M3_KNOB_GATE_CFG  0 (default: chosen from T) or 1..4 = force gate_fwd tile configuration 0..3.
 *   M3_KNOB_DEBUG     measurement only (results are garbage): tcgen05 GEMMs run 1 = without MMAs, 2 = without TMA loads.
 *   M3_KNOB_TRACE_KERNEL  1 + index of the GEMM launch inside one m3_ffn_fwd / m3_ffn_bwd call that m3_debug_trace_buffer
 *                     records (0 = every launch).
 *   M3_KNOB_FFN_CHAIN 1 (default): a bf16 forward that keeps no state (saved = NULL) with D in {128, 256, 384} and
 *                     H <= 2 D runs as ONE chain kernel (fc1 -> GELU -> fc2, h never leaves the SM; ffn_chain.cu);
 *                     0: always the two grouped GEMMs.
 *   M3_KNOB_MOVER_VARIANT  0 (default); 9: dispatch_bwd keeps the exact SIMT fp32 router term for bf16 queues too
 *                     (default: mma.sync bf16; the test of that kernel compares the two).
 * Returns the previous value, or M3_ERR_ARG for an unknown knob. */
typedef enum { M3_KNOB_PDL = 0, M3_KNOB_EPI_WARPS = 1, M3_KNOB_MOVER_VARIANT = 2, M3_KNOB_GATE_CFG = 3, M3_KNOB_DEBUG = 4, M3_KNOB_TRACE_KERNEL = 5, M3_KNOB_FFN_CHAIN = 6, M3_KNOB_COUNT_ = 8 } m3_knob;
int m3_set_knob(int knob, int value);
int m3_ffn_fwd(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert,
               int cap_rows, int E, int D, int H, const void* w1, const float* b1, const void* w2,
               const float* b2, void* saved, void* yq, void* workspace, size_t workspace_bytes,
               m3_stream_t stream);
int m3_ffn_bwd(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
               const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
               const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq,
               float* dw1, float* db1, float* dw2, float* db2, void* workspace,
               size_t workspace_bytes, m3_stream_t stream);

/* The same pair with EXPERT DROPOUT behind the GELU: the reference's experts apply nn.Sequential(GELU, Dropout(p))
 * (models/moe/origin/vision_transformer_moe.py:248-251; drop_rate 0.1 in configs/nyud/vit_moe/*drop0.1*.yml), i.e.
 * h = m * gelu(z), m = 0 with probability p and 1/(1-p) otherwise.  The mask is a counter-based function (Philox4x32-7,
 * csrc/philox.cuh) of the element's queue coordinates and of `rng_state` = {uint64 seed, uint64 call counter} in DEVICE
 * memory (16-byte aligned; read by the kernels, so that a caller can bump the counter with a stream-ordered op and stay
 * CUDA-graph capturable).  The backward call must be given the rng_state VALUES of its forward call (keep a copy).
 * bf16: the saved planes carry the mask, the backward does not regenerate it; fp32: regenerated.  Not torch's random
 * stream: parity with the reference is statistical.  drop_p = 0 is m3_ffn_fwd / m3_ffn_bwd; saved must not be NULL. */
int m3_ffn_fwd_dropout(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert,
                       int cap_rows, int E, int D, int H, const void* w1, const float* b1, const void* w2,
                       const float* b2, void* saved, void* yq, void* workspace, size_t workspace_bytes,
                       float drop_p, const void* rng_state, m3_stream_t stream);
int m3_ffn_bwd_dropout(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
                       const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                       const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq,
                       float* dw1, float* db1, float* dw2, float* db2, void* workspace,
                       size_t workspace_bytes, float drop_p, const void* rng_state, m3_stream_t stream);

/* m3_ffn_bwd(_dropout) in two halves: parts = 1 the data gradients (dxq; the intermediate dz stays in `workspace`),
 * parts = 2 the weight / bias gradients (needs the workspace a parts = 1 call with the same arguments filled), parts = 3
 * both.  The expert-parallel backward issues the data gradients first - they are what the peers wait for - and runs the
 * router backward and the weight gradients before the last rendezvous, where they absorb rank skew. */
int m3_ffn_bwd_parts(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
                     const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                     const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq, float* dw1,
                     float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes,
                     float drop_p, const void* rng_state, int parts, m3_stream_t stream);

/* fp32 master weights [E,R,C] -> bf16 copy [E,R,C] and (optional) bf16 transpose [E,C,R]. */
int m3_cast_weights_bf16(const float* w, int E, int R, int C, void* w_bf16, void* wt_bf16,
                         m3_stream_t stream);

/* ------------------------------------------------- expert-parallel exchange --
 * Replace fmoe_cuda.global_scatter / global_gather (grouped ncclSend/ncclRecv with
 * host-side counts, reached from MOEScatter/MOEGather when world_size > 1) by
 * direct NVLink peer access: every rank writes its rows straight into the OWNER
 * rank's receive queue and reads results straight out of it (peer pointers
 * obtained once by the host through CUDA IPC), so the row movers ARE the
 * all-to-all and no host count read-back is needed.  See DESIGN.md, section EP.
 *   peer_*[W]     device array of W queue base pointers (one per rank, own rank too)
 *   dst_rank[T*K] owner rank of every local slot;  dst_row[T*K] its row there
 * Same arithmetic as the local row movers above; padding rows are zeroed by the
 * owner with m3_zero_pad_rows.
 *
 * m3_ep_plan: from the all-gathered count matrix cnt[W][E_tot] (E_tot = W*E_loc,
 * expert e lives on rank e / E_loc: utils/common_config.py:179-185) and this
 * rank's local stable plan (pos_local from m3_route_plan with pad = 1), derive
 *   dst_rank/dst_row for every local slot, and this rank's receive layout:
 *   recv_counts[E_loc], recv_offsets[E_loc+1] (padded), recv_tile_expert[].
 * Receive queue of local expert le: sources in rank order, each source's rows in
 * that source's slot order.  cap_rows = capacity of every rank's receive queue: slots
 * whose destination row would not fit are dropped (dst_row = -1) and *overflow_flag is
 * set to 1 (the host checks it lazily; size queues with capacity_factor, see ep.py).
 */
int m3_ep_plan(const int64_t* idx, const int32_t* pos_local, const int32_t* cnt_all, int rank, int W,
               int E_loc, int T, int K, int pad, int cap_rows, int32_t* dst_rank, int32_t* dst_row,
               int32_t* recv_counts, int32_t* recv_offsets, int32_t* recv_tile_expert,
               int32_t* overflow_flag, int32_t* pos_id, void* const* peer_inv, int32_t* meta,
               m3_stream_t stream);
/* Optional outputs of m3_ep_plan (each may be NULL) for the fused return store below:
 *   pos_id[T*K]   s for a live slot, -1 for a dropped one: the "queue position" of slot s in a slot-ordered return
 *                 buffer, to be handed to m3_combine_fwd / m3_dispatch_bwd.
 *   meta[cap_rows] + peer_inv[W] (both or neither; needs T*K <= 2^24, W <= 128): row origins of MY receive queue,
 *                 meta[r] = (source rank << 24) | slot, read from the sources' inverse plans (peer_inv[r] = rank r's
 *                 inv_pos from m3_route_plan(pad = 1), complete before the count exchange): coalesced 4-byte reads,
 *                 ~T*K*4 bytes per rank, instead of a separate 4-byte remote store with every pushed row.  Padding
 *                 rows are set to -1 by m3_zero_pad_rows. */
int m3_ep_dispatch_fwd(const void* x, int x_dtype, const int32_t* dst_rank, const int32_t* dst_row,
                       int T, int K, int D, void* const* peer_xq, int xq_dtype, m3_stream_t stream);
/* The expert FFN FUSED WITH THE RETURN HALF OF THE ALL-TO-ALL (bf16 / tcgen05 path only).  Same arithmetic as
 * m3_ffn_fwd_dropout / m3_ffn_bwd_parts over this rank's receive queue, but the epilogue of the LAST GEMM (fc2 in the
 * forward, dxq = dz W1 in the backward) stores every result row straight into the SOURCE rank's slot-ordered return
 * buffer over NVLink - row r goes to peer_ret[meta[r] >> 24] + (meta[r] & 0xffffff) * D, rows with meta[r] < 0
 * (padding) go nowhere - so the exchange that FastMoE runs as a separate global_gather after the expert GEMM
 * (fmoe MOEGather, reached from models/moe/origin/custom_moe_layer.py:255-257) overlaps the GEMM tile by tile and no
 * result queue exists on the owner.  After a rendezvous the source combines its return buffer locally
 * (m3_combine_fwd / m3_dispatch_bwd with pos = pos_id).
 *   ret_meta[cap_rows]  row origins written by the sources' m3_ep_dispatch_fwd (+ m3_zero_pad_rows for padding rows)
 *   peer_yret / peer_dxret[W]  device array: every rank's [T*K, D] bf16 return buffer */
int m3_ep_ffn_fwd(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows,
                  int E, int D, int H, const void* w1, const float* b1, const void* w2, const float* b2,
                  void* saved, const int32_t* ret_meta, void* const* peer_yret, void* workspace,
                  size_t workspace_bytes, float drop_p, const void* rng_state, m3_stream_t stream);
int m3_ep_ffn_bwd(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
                  const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                  const void* w1, const void* w2, const void* w1t, const void* w2t, const int32_t* ret_meta,
                  void* const* peer_dxret, float* dw1, float* db1, float* dw2, float* db2, void* workspace,
                  size_t workspace_bytes, float drop_p, const void* rng_state, int parts, m3_stream_t stream);
/* ysave [T*K, D] (queue dtype, may be NULL): m3_ep_combine_fwd keeps a LOCAL copy, in slot order, of
 * the result rows it pulls over NVLink; given to m3_ep_combine_bwd, dscore = <g, y> is computed from
 * that copy and the backward pass only PUSHES dyq (peer_yq may then be NULL). */
int m3_ep_combine_fwd(void* const* peer_yq, int yq_dtype, const int32_t* dst_rank,
                      const int32_t* dst_row, const float* score, int T, int K, int D, void* out,
                      int out_dtype, void* ysave, m3_stream_t stream);
int m3_ep_combine_bwd(const void* g, int g_dtype, void* const* peer_yq, void* const* peer_dyq,
                      int q_dtype, const int32_t* dst_rank, const int32_t* dst_row,
                      const float* score, int T, int K, int D, float* dscore, const void* ysave,
                      m3_stream_t stream);
int m3_ep_dispatch_bwd(void* const* peer_dxq, int dxq_dtype, const int32_t* dst_rank,
                       const int32_t* dst_row, int T, int K, int D, const float* dz,
                       const float* w_gate, int E, void* dx, int dx_dtype, m3_stream_t stream);
/* zeroes the padding rows of a receive queue; meta (may be NULL): also marks them -1 in the row-origin array */
int m3_zero_pad_rows(void* q, int dtype, const int32_t* counts, const int32_t* offsets, int E, int D,
                     int32_t* meta, m3_stream_t stream);
/* Device-side rendezvous of the W ranks over peer memory (replaces an NCCL barrier / the
 * fmoe expert_exchange count all-to-all): a 1-warp kernel stores `epoch` into slot `rank` of every
 * peer's flag array (system-scope release) and, if `payload` != NULL, first copies `payload_ints`
 * int32 into row `rank` of every peer's gather buffer; it then spins (acquire) until all W slots of ITS OWN flag array
 * have reached `epoch`.  Every rank must call it with the same epoch.  The spin gives up after 30 minutes
 * (M3_EP_BARRIER_TIMEOUT_S seconds, read once; 0 = never) with a device-side message and a trap, so that a dead rank
 * cannot hang a GPU for ever.
 *   peer_flags[W]  device array: base of every rank's flag array (W int32 each)
 *   peer_gather[W] device array: base of every rank's gather buffer [W][payload_ints] (or NULL)
 * Only valid with one process per GPU (ranks must be co-resident on different devices). */
int m3_ep_barrier(void* const* peer_flags, void* const* peer_gather, const int32_t* payload,
                  int payload_ints, int rank, int W, int epoch, m3_stream_t stream);

/* ---- Block-level fusion around the layer --------------------
 * Replaces, in the reference Block,  x + drop_path(mlp_drop(mlp(norm2(x), ...)))
 * (models/moe/origin/vision_transformer_moe.py:278-283; ckpt twin :441-451), the separate
 * LayerNorm (norm2) and residual-add passes.  x is the RAW fp32 residual stream [T, D]; the
 * normalised tokens are never written to memory.
 *   m3_ln_stats        mean[T], rstd[T] = 1/sqrt(var_biased + eps)      (torch.nn.LayerNorm)
 *   m3_ln_fold_gate    gb[2,E] = {G = gamma^T W, B = beta^T W};  w_fold[Dg,E] = gamma (.) w_gate - G/D
 *                      (column-centred, so the token mean cancels exactly; rows >= D copied)
 *   m3_gate_fwd_ln     m3_gate_fwd on raw x:  z = rstd * (x @ w_fold) + gb[1]
 *   m3_dispatch_fwd_ln m3_dispatch_fwd with LayerNorm applied on the fly
 *   m3_combine_fwd_res out[T,D] (fp32) = residual + sum_k score * yq[pos]
 *   m3_gate_bwd_ln     m3_gate_bwd with x normalised on load (w_gate = the ORIGINAL weights)
 *   m3_ln_bwd_res      dx = dres + LayerNorm'(dxn);  dgamma[D], dbeta[D]   (deterministic)           */
int m3_ln_stats(const float* x, int T, int D, float eps, float* mean, float* rstd, m3_stream_t stream);
int m3_ln_fold_gate(const float* w_gate, const float* gamma, const float* beta, int D, int Dg, int E,
                    float* w_fold, float* gb, m3_stream_t stream);
int m3_gate_fwd_ln(const float* x, int64_t ldx, const float* ln_mean, const float* ln_rstd,
                   const float* ln_gb, const float* task_feat, const float* w_gate_folded,
                   const float* noise, float noise_stddev, int T, int D, int Dt, int E, int K,
                   int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean_logits,
                   float* noisy_logits, float* gates, float* imp_partial, int32_t* load_partial,
                   m3_stream_t stream);
int m3_dispatch_fwd_ln(const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, const int32_t* pos, const int32_t* counts,
                       const int32_t* offsets, int T, int K, int D, int E, void* xq, int xq_dtype,
                       m3_stream_t stream);
int m3_combine_fwd_res(const void* yq, int yq_dtype, const int32_t* pos, const float* score,
                       const float* residual, int T, int K, int D, float* out, m3_stream_t stream);
int m3_gate_bwd_ln(const float* x, int64_t ldx, const float* ln_mean, const float* ln_rstd,
                   const float* ln_gamma, const float* ln_beta, const float* task_feat,
                   const float* w_gate, const float* logits, const int32_t* idx_full, int T, int D, int Dt,
                   int E, int K, const float* dscore, const float* dtop_vals, const float* dgates,
                   const float* dimportance, const float* dclean, const float* dnoisy,
                   const float* importance, const float* dcv_loss, float* dz, float* dw_gate,
                   float* dtask_feat, void* workspace, size_t workspace_bytes, m3_stream_t stream);
size_t m3_ln_bwd_workspace_bytes(int T, int D);
int m3_ln_bwd_res(const float* dxn, const float* x, const float* mean, const float* rstd,
                  const float* gamma, const float* dres, int T, int D, float* dx, float* dgamma,
                  float* dbeta, void* workspace, size_t workspace_bytes, m3_stream_t stream);

/* Debug only: clock64 timeline of CTA 0 of the tensor-core GEMM kernels (producer / MMA / one epilogue warp), appended
 * to a caller-owned DEVICE buffer of 4 + 6*max_events uint64 (buf[r] = event count of role r = 0 producer / 1 MMA /
 * 2 epilogue warp 0, whose {tag, clock} pairs start at buf[4 + 2*r*max_events]; zero it first); NULL = off. */
int m3_debug_trace_buffer(unsigned long long* dev_buf, int max_events);

/* CUDA IPC plumbing for the peer queues (host pointers in/out; 64-byte handles). */
int m3_ipc_alloc(size_t bytes, void** dev_ptr, void* handle64);
int m3_ipc_open(const void* handle64, void** dev_ptr);
int m3_ipc_close(void* dev_ptr);
int m3_ipc_free(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* M3VIT_MOE_H_ */

// C-ABI glue: status strings, device check, FFN dtype dispatch, weight casts,
// expert-parallel plan, CUDA-IPC plumbing.  See include/m3vit_moe.h.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

// implemented in ffn_f32.cu / ffn_bf16.cu
int m3_ffn_fwd_f32(const float* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                   int H, const float* w1, const float* b1, const float* w2, const float* b2, float* hpre,
                   float* yq, void* workspace, size_t workspace_bytes, float drop_p, const void* rng, cudaStream_t st);
int m3_ffn_bwd_f32(const float* xq, const float* hpre, const float* dyq, const int32_t* counts,
                   const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                   const float* w1, const float* w2, float* dxq, float* dw1, float* db1, float* dw2, float* db2,
                   void* workspace, size_t workspace_bytes, float drop_p, const void* rng, int parts, cudaStream_t st);
size_t m3_ffn_bf16_workspace_bytes(int cap_rows, int D, int H, int E, int backward);
size_t m3_ffn_bf16_saved_bytes(int cap_rows, int D, int H);
int m3_ffn_bf16_chain_mode(int D, int H);
int m3_ffn_bf16_set_sm_limit(int sms);
int m3_ffn_fwd_bf16(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                    int H, const void* w1, const float* b1, const void* w2, const float* b2, void* hpre, void* yq,
                    void* workspace, size_t workspace_bytes, float drop_p, const void* rng, const int32_t* ret_meta,
                    void* const* ret_bases, cudaStream_t st);
int m3_ffn_bwd_bf16(const void* xq, const void* hpre, const void* dyq, const int32_t* counts,
                    const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                    const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq, float* dw1,
                    float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes, int parts,
                    const int32_t* ret_meta, void* const* ret_bases, cudaStream_t st);

extern "C" int m3_abi_version(void) { return M3_ABI_VERSION; }

extern "C" const char* m3_status_string(int status) {
  switch (status) {
    case M3_OK: return "ok";
    case M3_ERR_ARG: return "invalid argument (null pointer or negative size)";
    case M3_ERR_SHAPE: return "unsupported shape";
    case M3_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case M3_ERR_UNSUPPORTED: return "unsupported dtype or feature";
    case M3_ERR_DEVICE: return "current CUDA device is not sm_100 (B200)";
    case M3_ERR_WORKSPACE: return "workspace too small";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown m3 status";
}

extern "C" int m3_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return (int)e;
  return major == 10 ? M3_OK : M3_ERR_DEVICE;
}

// ------------------------------------------------------------------ expert FFN
extern "C" size_t m3_ffn_workspace_bytes(int dtype, int cap_rows, int D, int H, int E, int backward) {
  if (dtype == M3_F32) return (size_t)cap_rows * H * sizeof(float);
  return m3_ffn_bf16_workspace_bytes(cap_rows, D, H, E, backward);
}

extern "C" int m3_set_gemm_sm_limit(int sms) { return m3_ffn_bf16_set_sm_limit(sms); }

namespace m3 {
int g_knobs[M3_KNOB_COUNT_] = {/*PDL*/ 0, /*EPI_WARPS*/ 0, /*MOVER_VARIANT*/ 0, /*GATE_CFG*/ 0, /*DEBUG*/ 0, /*TRACE_KERNEL*/ 0,
                                /*FFN_CHAIN*/ 1};      // 1: chain kernel for state-free forwards
}
namespace m3 { namespace tc {
unsigned long long* g_trace_buf = nullptr;
int g_trace_cap = 0;
} }
// Debug only: device buffer of 4 + 6*max_events uint64 filled by CTA 0 of the tensor-core GEMM kernels while set
// (layout: include/m3vit_moe.h); NULL switches tracing off.
extern "C" int m3_debug_trace_buffer(unsigned long long* dev_buf, int max_events) {
  m3::tc::g_trace_buf = dev_buf;
  m3::tc::g_trace_cap = dev_buf ? max_events : 0;
  return M3_OK;
}
extern "C" int m3_set_knob(int knob, int value) {
  if (knob < 0 || knob >= M3_KNOB_COUNT_) return M3_ERR_ARG;
  const int old = m3::g_knobs[knob];
  m3::g_knobs[knob] = value;
  return old;
}

extern "C" size_t m3_ffn_saved_bytes(int dtype, int cap_rows, int D, int H) {
  if (cap_rows < 0 || H <= 0 || D <= 0) return 0;
  if (dtype == M3_BF16) return m3_ffn_bf16_saved_bytes(cap_rows, D, H);
  return (size_t)cap_rows * H * sizeof(float);      // fp32 parity path: the pre-activation
}

extern "C" int m3_ffn_uses_chain(int dtype, int D, int H) {
  return dtype == M3_BF16 ? m3_ffn_bf16_chain_mode(D, H) : 0;
}

static int ffn_fwd_impl(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows,
                        int E, int D, int H, const void* w1, const float* b1, const void* w2, const float* b2,
                        void* hpre, void* yq, void* workspace, size_t workspace_bytes, float drop_p, const void* rng,
                        const int32_t* ret_meta, void* const* ret_bases, m3_stream_t stream) {
  M3_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f);
  if (rng) M3_CHECK_ALIGN16(rng);
  M3_CHECK_ARG(xq && offsets && tile_expert && w1 && b1 && w2 && b2 && (yq || ret_meta));
  M3_CHECK_ARG((ret_meta == nullptr) == (ret_bases == nullptr));
  if (ret_meta && dtype != M3_BF16) return M3_ERR_UNSUPPORTED;      // the return store lives in the tcgen05 epilogue
  M3_CHECK_ARG(cap_rows >= 0 && E >= 1 && D > 0 && H > 0);
  M3_CHECK_SHAPE(cap_rows % M3_PAD_ROWS == 0);
  M3_CHECK_ALIGN16(xq); M3_CHECK_ALIGN16(w1); M3_CHECK_ALIGN16(w2);
  if (yq) M3_CHECK_ALIGN16(yq);
  M3_CHECK_ALIGN16(b1); M3_CHECK_ALIGN16(b2);
  if (hpre) M3_CHECK_ALIGN16(hpre);
  if (workspace) M3_CHECK_ALIGN16(workspace);
  if (cap_rows == 0) return M3_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == M3_F32)
    return m3_ffn_fwd_f32((const float*)xq, offsets, tile_expert, cap_rows, E, D, H, (const float*)w1, b1,
                          (const float*)w2, b2, (float*)hpre, (float*)yq, workspace, workspace_bytes, drop_p, rng, st);
  if (dtype == M3_BF16)
    return m3_ffn_fwd_bf16(xq, offsets, tile_expert, cap_rows, E, D, H, w1, b1, w2, b2, hpre, yq, workspace,
                           workspace_bytes, drop_p, rng, ret_meta, ret_bases, st);
  return M3_ERR_UNSUPPORTED;
}

extern "C" int m3_ffn_fwd(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows,
                          int E, int D, int H, const void* w1, const float* b1, const void* w2, const float* b2,
                          void* hpre, void* yq, void* workspace, size_t workspace_bytes, m3_stream_t stream) {
  return ffn_fwd_impl(dtype, xq, offsets, tile_expert, cap_rows, E, D, H, w1, b1, w2, b2, hpre, yq, workspace,
                      workspace_bytes, 0.f, nullptr, nullptr, nullptr, stream);
}

extern "C" int m3_ffn_fwd_dropout(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert,
                                  int cap_rows, int E, int D, int H, const void* w1, const float* b1, const void* w2,
                                  const float* b2, void* saved, void* yq, void* workspace, size_t workspace_bytes,
                                  float drop_p, const void* rng_state, m3_stream_t stream) {
  return ffn_fwd_impl(dtype, xq, offsets, tile_expert, cap_rows, E, D, H, w1, b1, w2, b2, saved, yq, workspace,
                      workspace_bytes, drop_p, rng_state, nullptr, nullptr, stream);
}

extern "C" int m3_ep_ffn_fwd(int dtype, const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows,
                             int E, int D, int H, const void* w1, const float* b1, const void* w2, const float* b2,
                             void* saved, const int32_t* ret_meta, void* const* peer_yret, void* workspace,
                             size_t workspace_bytes, float drop_p, const void* rng_state, m3_stream_t stream) {
  M3_CHECK_ARG(ret_meta && peer_yret);
  return ffn_fwd_impl(dtype, xq, offsets, tile_expert, cap_rows, E, D, H, w1, b1, w2, b2, saved, nullptr, workspace,
                      workspace_bytes, drop_p, rng_state, ret_meta, peer_yret, stream);
}

static int ffn_bwd_impl(int dtype, const void* xq, const void* hpre, const void* dyq, const int32_t* counts,
                        const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                        const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq, float* dw1,
                        float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes, float drop_p,
                        const void* rng, int parts, const int32_t* ret_meta, void* const* ret_bases,
                        m3_stream_t stream) {
  M3_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && parts >= 1 && parts <= 3);
  M3_CHECK_ARG(xq && hpre && dyq && counts && offsets && tile_expert && w1 && w2 && (dxq || ret_meta) && dw1 && db1 && dw2 && db2);
  M3_CHECK_ARG((ret_meta == nullptr) == (ret_bases == nullptr));
  if (ret_meta && dtype != M3_BF16) return M3_ERR_UNSUPPORTED;
  M3_CHECK_ARG(cap_rows >= 0 && E >= 1 && D > 0 && H > 0 && workspace);
  M3_CHECK_SHAPE(cap_rows % M3_PAD_ROWS == 0);
  M3_CHECK_ALIGN16(xq); M3_CHECK_ALIGN16(hpre); M3_CHECK_ALIGN16(dyq); M3_CHECK_ALIGN16(dxq);
  M3_CHECK_ALIGN16(w1); M3_CHECK_ALIGN16(w2); M3_CHECK_ALIGN16(dw1); M3_CHECK_ALIGN16(dw2);
  M3_CHECK_ALIGN16(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == M3_F32)
    return m3_ffn_bwd_f32((const float*)xq, (const float*)hpre, (const float*)dyq, counts, offsets, tile_expert,
                          cap_rows, E, D, H, (const float*)w1, (const float*)w2, (float*)dxq, dw1, db1, dw2, db2,
                          workspace, workspace_bytes, drop_p, rng, parts, st);
  if (dtype == M3_BF16) {
    M3_CHECK_ARG(w1t && w2t);
    // bf16: the saved planes already carry the forward's keep-scale (h = m gelu(z), m gelu'(z)): nothing to regenerate
    return m3_ffn_bwd_bf16(xq, hpre, dyq, counts, offsets, tile_expert, cap_rows, E, D, H, w1, w2, w1t, w2t, dxq,
                           dw1, db1, dw2, db2, workspace, workspace_bytes, parts, ret_meta, ret_bases, st);
  }
  return M3_ERR_UNSUPPORTED;
}

extern "C" int m3_ffn_bwd(int dtype, const void* xq, const void* hpre, const void* dyq, const int32_t* counts,
                          const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                          const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq, float* dw1,
                          float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes,
                          m3_stream_t stream) {
  return ffn_bwd_impl(dtype, xq, hpre, dyq, counts, offsets, tile_expert, cap_rows, E, D, H, w1, w2, w1t, w2t, dxq, dw1,
                      db1, dw2, db2, workspace, workspace_bytes, 0.f, nullptr, 3, nullptr, nullptr, stream);
}

extern "C" int m3_ffn_bwd_parts(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
                                const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                                const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq, float* dw1,
                                float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes,
                                float drop_p, const void* rng_state, int parts, m3_stream_t stream) {
  return ffn_bwd_impl(dtype, xq, saved, dyq, counts, offsets, tile_expert, cap_rows, E, D, H, w1, w2, w1t, w2t, dxq, dw1,
                      db1, dw2, db2, workspace, workspace_bytes, drop_p, rng_state, parts, nullptr, nullptr, stream);
}

extern "C" int m3_ep_ffn_bwd(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
                             const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                             const void* w1, const void* w2, const void* w1t, const void* w2t, const int32_t* ret_meta,
                             void* const* peer_dxret, float* dw1, float* db1, float* dw2, float* db2, void* workspace,
                             size_t workspace_bytes, float drop_p, const void* rng_state, int parts, m3_stream_t stream) {
  M3_CHECK_ARG(ret_meta && peer_dxret);
  return ffn_bwd_impl(dtype, xq, saved, dyq, counts, offsets, tile_expert, cap_rows, E, D, H, w1, w2, w1t, w2t, nullptr,
                      dw1, db1, dw2, db2, workspace, workspace_bytes, drop_p, rng_state, parts, ret_meta, peer_dxret,
                      stream);
}

extern "C" int m3_ffn_bwd_dropout(int dtype, const void* xq, const void* saved, const void* dyq, const int32_t* counts,
                                  const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                                  const void* w1, const void* w2, const void* w1t, const void* w2t, void* dxq,
                                  float* dw1, float* db1, float* dw2, float* db2, void* workspace,
                                  size_t workspace_bytes, float drop_p, const void* rng_state, m3_stream_t stream) {
  return ffn_bwd_impl(dtype, xq, saved, dyq, counts, offsets, tile_expert, cap_rows, E, D, H, w1, w2, w1t, w2t, dxq, dw1,
                      db1, dw2, db2, workspace, workspace_bytes, drop_p, rng_state, 3, nullptr, nullptr, stream);
}


// ------------------------------------------------------------------ weight cast
namespace m3 {
// 32x32 tile: coalesced fp32 read, bf16 straight copy + bf16 transpose via smem
__global__ void cast_weights_kernel(const float* __restrict__ w, int R, int C, __nv_bfloat16* __restrict__ o,
                                    __nv_bfloat16* __restrict__ ot) {
  __shared__ float tile[32][33];
  const int e = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const float* we = w + (int64_t)e * R * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = we[(int64_t)r * C + c];
      if (o) o[(int64_t)e * R * C + (int64_t)r * C + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  if (ot == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) ot[(int64_t)e * R * C + (int64_t)c * R + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

__global__ void ep_plan_kernel(const int64_t* __restrict__ idx, const int32_t* __restrict__ pos_local,
                               const int32_t* __restrict__ cnt_all, int rank, int W, int E_loc, int R, int pad,
                               int cap_rows, int32_t* __restrict__ dst_rank, int32_t* __restrict__ dst_row,
                               int32_t* __restrict__ recv_counts, int32_t* __restrict__ recv_offsets,
                               int32_t* __restrict__ recv_tile_expert, int32_t* __restrict__ overflow_flag,
                               int32_t* __restrict__ pos_id, int32_t* const* __restrict__ peer_inv,
                               int32_t* __restrict__ meta) {
  extern __shared__ int sm[];
  const int E_tot = W * E_loc;
  int* loc_off = sm;            // [E_tot] exclusive prefix of this rank's counts (pad 1)
  int* base = loc_off + E_tot;  // [E_tot] first row of my segment in the owner's queue
  int* tot = base + E_tot;      // [E_tot] rows of global expert ge over all sources
  for (int ge = threadIdx.x; ge < E_tot; ge += blockDim.x) {
    int t = 0, before = 0;
    for (int s = 0; s < W; ++s) {
      const int c = cnt_all[s * E_tot + ge];
      t += c;
      if (s < rank) before += c;
    }
    tot[ge] = t;
    base[ge] = before;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int ge = 0; ge < E_tot; ++ge) { loc_off[ge] = run; run += cnt_all[rank * E_tot + ge]; }
    for (int o = 0; o < W; ++o) {
      int roff = 0;
      for (int le = 0; le < E_loc; ++le) {
        const int ge = o * E_loc + le;
        base[ge] += roff;
        if (o == rank && blockIdx.x == 0) { recv_counts[le] = tot[ge]; recv_offsets[le] = roff; }
        roff += (tot[ge] + pad - 1) / pad * pad;
      }
      if (o == rank && blockIdx.x == 0) recv_offsets[E_loc] = roff;
    }
  }
  __syncthreads();
  if (meta != nullptr && peer_inv != nullptr) {
    // Row origins of MY receive queue, for the return store: row i of the (local expert le, source src) sub-segment is
    // the i-th row of src's sorted send list for that expert, i.e. slot inv_src[loc_off_src[ge] + i].  Pulled from the
    // sources' inverse plans (coalesced 4-byte reads over NVLink, ~T*K*4 bytes in total) instead of having every
    // pushed row carry a separate 4-byte remote store.
    // every (local expert, source) sub-segment is spread over ALL blocks: the reads are latency-bound (NVLink round
    // trips), so they want every SM's worth of loads in flight, not one block per sub-segment
    const int NP = E_loc * W;
    int* pr = tot + E_tot;       // [3 * NP]: first sorted position at the source, first row in my queue, flat prefix
    for (int pair = threadIdx.x; pair < NP; pair += blockDim.x) {
      const int le = pair / W, src = pair % W, ge = rank * E_loc + le;
      int p0 = 0, r0 = 0;
      for (int g2 = 0; g2 < ge; ++g2) p0 += cnt_all[src * E_tot + g2];
      for (int l2 = 0; l2 < le; ++l2) r0 += (tot[rank * E_loc + l2] + pad - 1) / pad * pad;
      for (int s2 = 0; s2 < src; ++s2) r0 += cnt_all[s2 * E_tot + ge];
      pr[3 * pair] = p0; pr[3 * pair + 1] = r0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = 0;
      for (int pair = 0; pair < NP; ++pair) {
        pr[3 * pair + 2] = run;
        run += cnt_all[(pair % W) * E_tot + rank * E_loc + pair / W];
      }
      pr[3 * NP] = run;
    }
    __syncthreads();
    const int total = pr[3 * NP];
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
      int lo = 0, hi = NP;       // largest pair whose flat prefix <= f (empty pairs share their successor's prefix)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pr[3 * mid + 2] <= f) lo = mid; else hi = mid;
      }
      const int i = f - pr[3 * lo + 2], src = lo % W, row = pr[3 * lo + 1] + i;
      if (row < cap_rows) meta[row] = (src << 24) | peer_inv[src][pr[3 * lo] + i];
    }
  }
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < R; s += gridDim.x * blockDim.x) {
    const int64_t ge = idx[s];
    const int p = pos_local[s];
    if (ge < 0 || ge >= E_tot || p < 0) {
      dst_rank[s] = 0; dst_row[s] = -1;
      if (pos_id != nullptr) pos_id[s] = -1;
      continue;
    }
    const int row = base[ge] + (p - loc_off[ge]);
    dst_rank[s] = (int)(ge / E_loc);
    if (row >= cap_rows) {           // receive queue too small: drop the slot, tell the host
      dst_row[s] = -1;
      if (overflow_flag != nullptr) *overflow_flag = 1;
      if (pos_id != nullptr) pos_id[s] = -1;
    } else {
      dst_row[s] = row;
      if (pos_id != nullptr) pos_id[s] = s;      // return buffers are in slot order: row of slot s = s
    }
  }
  if (blockIdx.x == 0) {
    // tile map of my receive queue
    int roff = 0;
    for (int le = 0; le < E_loc; ++le) {
      const int n = (tot[rank * E_loc + le] + pad - 1) / pad;
      for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (roff + (i + 1) * pad <= cap_rows) recv_tile_expert[roff / pad + i] = le;
      roff += n * pad;
    }
    // if the padded total exceeds the capacity, clamp what the GEMM sees to whole tiles that fit
    if (threadIdx.x == 0 && roff > cap_rows) {
      if (overflow_flag != nullptr) *overflow_flag = 1;
      for (int le = 0; le <= E_loc; ++le)
        if (recv_offsets[le] > cap_rows) recv_offsets[le] = cap_rows / pad * pad;
    }
  }
}
// Flag barrier (+ optional small all-gather) over peer memory; see m3_ep_barrier in the header.
__global__ void ep_barrier_kernel(int32_t* const* __restrict__ peer_flags, int32_t* const* __restrict__ peer_gather,
                                  const int32_t* __restrict__ payload, int n, int rank, int W, int epoch,
                                  long long timeout_clk) {
  const int lane = threadIdx.x;
  if (peer_gather != nullptr) {
    for (int p = 0; p < W; ++p) {
      int32_t* dst = peer_gather[p] + (int64_t)rank * n;
      for (int i = lane; i < n; i += 32) dst[i] = payload[i];
    }
  }
  __threadfence_system();          // payload (and everything earlier in this stream) before the flag
  __syncwarp();
  if (lane < W) {
    volatile int32_t* f = peer_flags[lane] + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
  if (lane < W) {
    const int32_t* mine = peer_flags[rank] + lane;
    const long long t0 = clock64();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (timeout_clk > 0 && clock64() - t0 > timeout_clk) {   // a peer never arrived (a dead rank must not hang the GPU)
        printf("m3_ep_barrier: rank %d timed out waiting for rank %d (epoch %d, have %d)\n", rank, lane, epoch, v);
        __trap();
      }
    } while (v < epoch);
  }
  __threadfence_system();
}
}  // namespace m3

extern "C" int m3_ep_barrier(void* const* peer_flags, void* const* peer_gather, const int32_t* payload,
                             int payload_ints, int rank, int W, int epoch, m3_stream_t stream) {
  M3_CHECK_ARG(peer_flags && W >= 1 && W <= 32 && rank >= 0 && rank < W && epoch > 0);
  M3_CHECK_ARG((peer_gather == nullptr) == (payload == nullptr) && payload_ints >= 0);
  // Ranks are routinely far apart (data-loader start-up at epoch boundaries, rank-0-only evaluation or checkpoint writing,
  // first-iteration lazy initialisation): wait 30 minutes by default like a collective library would, not seconds.
  // M3_EP_BARRIER_TIMEOUT_S overrides (0 = wait for ever).
  static const long long timeout_clk = [] {
    const char* v = getenv("M3_EP_BARRIER_TIMEOUT_S");
    const double sec = v != nullptr ? atof(v) : 1800.0;
    return sec <= 0 ? 0LL : (long long)(sec * 2.0e9);
  }();
  m3::ep_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<int32_t* const*>(peer_flags), reinterpret_cast<int32_t* const*>(peer_gather), payload,
      payload_ints, rank, W, epoch, timeout_clk);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" int m3_cast_weights_bf16(const float* w, int E, int R, int C, void* w_bf16, void* wt_bf16,
                                    m3_stream_t stream) {
  M3_CHECK_ARG(w && (w_bf16 || wt_bf16) && E >= 1 && R >= 1 && C >= 1);
  dim3 grid(m3_ceil_div(C, 32), m3_ceil_div(R, 32), E), block(32, 8);
  m3::cast_weights_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      w, R, C, static_cast<__nv_bfloat16*>(w_bf16), static_cast<__nv_bfloat16*>(wt_bf16));
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" int m3_ep_plan(const int64_t* idx, const int32_t* pos_local, const int32_t* cnt_all, int rank, int W,
                          int E_loc, int T, int K, int pad, int cap_rows, int32_t* dst_rank, int32_t* dst_row,
                          int32_t* recv_counts, int32_t* recv_offsets, int32_t* recv_tile_expert,
                          int32_t* overflow_flag, int32_t* pos_id, void* const* peer_inv, int32_t* meta,
                          m3_stream_t stream) {
  M3_CHECK_ARG(idx && pos_local && cnt_all && dst_rank && dst_row && recv_counts && recv_offsets && recv_tile_expert);
  M3_CHECK_ARG(W >= 1 && rank >= 0 && rank < W && E_loc >= 1 && T >= 0 && K >= 1 && pad >= 1 && cap_rows >= 0);
  M3_CHECK_SHAPE(W * E_loc <= 1024);
  M3_CHECK_ARG((meta == nullptr) == (peer_inv == nullptr));
  M3_CHECK_ARG(meta == nullptr || (W <= 128 && (int64_t)T * K <= (1 << 24)));
  const int R = T * K;
  int grid = m3_ceil_div(R, 256);
  if (grid < 1) grid = 1;
  if (grid > 2 * m3::kNumSMs) grid = 2 * m3::kNumSMs;
  m3::ep_plan_kernel<<<grid, 256, (6 * W * E_loc + 1) * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
      idx, pos_local, cnt_all, rank, W, E_loc, R, pad, cap_rows, dst_rank, dst_row, recv_counts, recv_offsets,
      recv_tile_expert, overflow_flag, pos_id, reinterpret_cast<int32_t* const*>(peer_inv), meta);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

// ------------------------------------------------------------------ CUDA IPC
extern "C" int m3_ipc_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  M3_CHECK_ARG(dev_ptr && handle64 && bytes > 0);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), *dev_ptr);
  if (e != cudaSuccess) { cudaFree(*dev_ptr); *dev_ptr = nullptr; return (int)e; }
  return M3_OK;
}
extern "C" int m3_ipc_open(const void* handle64, void** dev_ptr) {
  M3_CHECK_ARG(handle64 && dev_ptr);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? M3_OK : (int)e;
}
extern "C" int m3_ipc_close(void* dev_ptr) {
  M3_CHECK_ARG(dev_ptr);
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  return e == cudaSuccess ? M3_OK : (int)e;
}
extern "C" int m3_ipc_free(void* dev_ptr) {
  M3_CHECK_ARG(dev_ptr);
  cudaError_t e = cudaFree(dev_ptr);
  return e == cudaSuccess ? M3_OK : (int)e;
}

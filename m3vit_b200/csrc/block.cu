// Block-level fusion around the MoE layer for sm_100a.
//
// The reference Block computes   x + drop_path(mlp_drop(mlp(norm2(x), ...)))
// (/root/reference/models/moe/origin/vision_transformer_moe.py:278-283): a LayerNorm in
// front of the layer and a residual add behind it, i.e. two more [T, D] fp32 round trips
// on each side of the hot path.  Here the normalised tokens are never materialised:
//
//   ln_stats_kernel         per-token mean / rstd of the raw residual stream (one read of x)
//   ln_fold_gate_kernel     W' = gamma (.) w_gate - colmean,  B = beta^T w_gate, so that the gate
//                           kernel runs on RAW x:  z = rstd * (x W') + B
//   dispatch_fwd_ln_kernel  normalises on the fly while copying a token to its K queue rows
//   combine_fwd_res_kernel  out = x + sum_k score*yq                 (residual fused)
//   ln_bwd_res_kernel       dx = d_out + LayerNorm'(dxn), per-CTA partial dgamma / dbeta
//   ln_bwd_reduce_kernel    fixed-order reduction of the partials (deterministic)
//
// All HBM-bound; 16 lanes own a token, 128-bit transactions, like permute.cu.
#include "common.cuh"

namespace m3 {

constexpr int kBlkLanes = 16;
constexpr int kBlkThreads = 256;
constexpr int kBlkTok = kBlkThreads / kBlkLanes;

__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = kBlkLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mean[t], rstd[t] = 1/sqrt(var_biased + eps): two passes over the row held in registers
// (torch.nn.LayerNorm semantics; no E[x^2]-mean^2 cancellation).
template <int NV>
__global__ void __launch_bounds__(kBlkThreads)
ln_stats_kernel(const float* __restrict__ x, int T, int D, float eps, float* __restrict__ mean,
                float* __restrict__ rstd) {
  const int sub = threadIdx.x % kBlkLanes;
  const int tt = blockIdx.x * kBlkTok + threadIdx.x / kBlkLanes;
  const int t = tt < T ? tt : T - 1;
  const int nvec = D / 8;
  Vec8 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kBlkLanes;
    if (c < nvec) {
      v[i] = load8<float>(x + (int64_t)t * D + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i].v[j];
    }
  }
  const float mu = group16_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kBlkLanes;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i].v[j] - mu; q = fmaf(d, d, q); }
    }
  }
  const float var = group16_sum(q) / (float)D;
  if (tt < T && sub == 0) {
    mean[t] = mu;
    rstd[t] = rsqrtf(var + eps);
  }
}

// gb[0,e] = G_e = sum_d gamma[d]*w_gate[d,e],  gb[1,e] = B_e = sum_d beta[d]*w_gate[d,e]
// w_fold[d,e] = gamma[d]*w_gate[d,e] - G_e/D  for d < D  (task-feature rows d >= D copied).
// The columns are CENTRED: since sum_d (x_d - mean) = 0,
//   LayerNorm(x) @ W = rstd * sum_d (x_d - mean) * gamma_d W_de + B_e = rstd * (x @ w_fold)_e + B_e
// holds with the raw x and no  "- mean * G"  correction term (which would cancel catastrophically
// for tokens whose |mean| is large against their spread).
__global__ void __launch_bounds__(1024)
ln_fold_gate_kernel(const float* __restrict__ w_gate, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int D, int Dg, int E, float* __restrict__ w_fold, float* __restrict__ gb) {
  __shared__ float cm[1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < E; e += 32) {     // one warp per column, fixed reduction order
    float g = 0.f, b = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float w = w_gate[(int64_t)d * E + e];
      g = fmaf(gamma[d], w, g);
      b = fmaf(beta[d], w, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      gb[e] = g;
      gb[E + e] = b;
      cm[e] = g / (float)D;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Dg * E; i += blockDim.x) {
    const int d = i / E, e = i % E;
    w_fold[i] = d < D ? fmaf(gamma[d], w_gate[i], -cm[e]) : w_gate[i];
  }
}

// xq[pos[t,k]] = cast(LayerNorm(x[t]));  trailing CTAs zero the padding rows of every queue.
template <typename TO, int NV>
__global__ void __launch_bounds__(kBlkThreads)
dispatch_fwd_ln_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       const int32_t* __restrict__ pos, const int32_t* __restrict__ counts,
                       const int32_t* __restrict__ offsets, int T, int K, int D, int tok_ctas, TO* __restrict__ xq) {
  const int nvec = D / 8;
  if ((int)blockIdx.x >= tok_ctas) {
    const int e = blockIdx.x - tok_ctas;
    const int r0 = offsets[e] + counts[e], r1 = offsets[e + 1];
    Vec8 z;
#pragma unroll
    for (int i = 0; i < 8; ++i) z.v[i] = 0.f;
    for (int64_t i = threadIdx.x; i < (int64_t)(r1 - r0) * nvec; i += kBlkThreads)
      store8<TO>(xq + ((int64_t)r0 + i / nvec) * D + (i % nvec) * 8, z);
    return;
  }
  const int sub = threadIdx.x % kBlkLanes;
  const int t = blockIdx.x * kBlkTok + threadIdx.x / kBlkLanes;
  if (t >= T) return;
  const float mu = __ldg(mean + t), rs = __ldg(rstd + t);
  Vec8 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kBlkLanes;
    if (c < nvec) {
      v[i] = load8<float>(x + (int64_t)t * D + c * 8);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c * 8 + 4));
      const float g8[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i].v[j] = fmaf((v[i].v[j] - mu) * rs, g8[j], b8[j]);
    }
  }
  for (int k = 0; k < K; ++k) {
    const int row = __ldg(pos + (int64_t)t * K + k);
    if (row < 0) continue;
    TO* dst = xq + (int64_t)row * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = sub + i * kBlkLanes;
      if (c < nvec) store8<TO>(dst + c * 8, v[i]);
    }
  }
}

// out[t] = res[t] + sum_k score[t,k] * yq[pos[t,k]]  (fp32; the MoE sum is formed first, k ascending)
template <typename TI, int NV>
__global__ void __launch_bounds__(kBlkThreads)
combine_fwd_res_kernel(const TI* __restrict__ yq, const int32_t* __restrict__ pos, const float* __restrict__ score,
                       const float* __restrict__ res, int T, int K, int D, float* __restrict__ out) {
  const int sub = threadIdx.x % kBlkLanes;
  const int t = blockIdx.x * kBlkTok + threadIdx.x / kBlkLanes;
  if (t >= T) return;
  const int nvec = D / 8;
  Vec8 acc[NV], r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kBlkLanes;
    if (c < nvec) r[i] = load8<float>(res + (int64_t)t * D + c * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i].v[j] = 0.f;
  }
  for (int k = 0; k < K; ++k) {
    const int row = __ldg(pos + (int64_t)t * K + k);
    if (row < 0) continue;
    const float s = __ldg(score + (int64_t)t * K + k);
    const TI* src = yq + (int64_t)row * D;
    Vec8 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = sub + i * kBlkLanes;
      if (c < nvec) v[i] = load8<TI>(src + c * 8);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i].v[j] = fmaf(s, v[i].v[j], acc[i].v[j]);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kBlkLanes;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i].v[j] += r[i].v[j];
      store8<float>(out + (int64_t)t * D + c * 8, acc[i]);
    }
  }
}

// LayerNorm backward + residual:
//   xh = (x-mean)*rstd, gg = dxn*gamma, dx = dres + rstd*(gg - mean_d(gg) - xh*mean_d(gg*xh))
//   part[cta][0][d] = sum_t dxn*xh (dgamma),  part[cta][1][d] = sum_t dxn (dbeta)
// One WARP per token, float4 slices (NQ per lane): the six per-token arrays stay at 4*NQ registers
// each, so several CTAs fit an SM.  Grid-stride over tokens; per-thread partials are combined warp
// by warp in a fixed order.
constexpr int kLnbWarps = kBlkThreads / 32;

__device__ __forceinline__ float4 ld4_stream(const float* p) {
  const uint4 u = ldg_stream(p);
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}
template <int NQ>
__global__ void __launch_bounds__(kBlkThreads)
ln_bwd_res_kernel(const float* __restrict__ dxn, const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dres,
                  int T, int D, float* __restrict__ dx, float* __restrict__ part) {
  extern __shared__ __align__(16) float red[];  // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = D / 4;
  float4 gm[NQ], dg[NQ], db[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int c = lane + i * 32;
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = dg[i];
    gm[i] = c < nq ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : dg[i];
  }
  const float invD = 1.f / (float)D;
  for (int t = blockIdx.x * kLnbWarps + warp; t < T; t += gridDim.x * kLnbWarps) {
    const float mu = __ldg(mean + t), rs = __ldg(rstd + t);
    float4 g[NQ], xh[NQ], dr[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int c = lane + i * 32;
      if (c < nq) {
        g[i] = ld4_stream(dxn + (int64_t)t * D + c * 4);
        xh[i] = ld4_stream(x + (int64_t)t * D + c * 4);
        dr[i] = ld4_stream(dres + (int64_t)t * D + c * 4);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int c = lane + i * 32;
      if (c < nq) {
        float* gp = &g[i].x; float* xp = &xh[i].x; const float* mp = &gm[i].x;
        float* dgp = &dg[i].x; float* dbp = &db[i].x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xp[j] = (xp[j] - mu) * rs;
          dgp[j] = fmaf(gp[j], xp[j], dgp[j]);
          dbp[j] += gp[j];
          gp[j] *= mp[j];                 // gg
          s1 += gp[j];
          s2 = fmaf(gp[j], xp[j], s2);
        }
      }
    }
    const float c1 = warp_sum(s1) * invD, c2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int c = lane + i * 32;
      if (c < nq) {
        uint4 o;
        o.x = __float_as_uint(fmaf(rs, g[i].x - c1 - xh[i].x * c2, dr[i].x));
        o.y = __float_as_uint(fmaf(rs, g[i].y - c1 - xh[i].y * c2, dr[i].y));
        o.z = __float_as_uint(fmaf(rs, g[i].z - c1 - xh[i].z * c2, dr[i].z));
        o.w = __float_as_uint(fmaf(rs, g[i].w - c1 - xh[i].w * c2, dr[i].w));
        stg_stream(dx + (int64_t)t * D + c * 4, o);
      }
    }
  }
  for (int i = threadIdx.x; i < 2 * D; i += kBlkThreads) red[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < kLnbWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int c = lane + i * 32;
        if (c < nq) {
          const float* dgp = &dg[i].x; const float* dbp = &db[i].x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            red[c * 4 + j] += dgp[j];
            red[D + c * 4 + j] += dbp[j];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * D; i += kBlkThreads) part[(int64_t)blockIdx.x * 2 * D + i] = red[i];
}

// out2[i] = sum_cta part[cta][i]  (i < 2*D): 32 warps stride over the CTAs, combined in a fixed order
__global__ void __launch_bounds__(1024)
ln_bwd_reduce_kernel(const float* __restrict__ part, int nparts, int D, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  __shared__ float red[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float a = 0.f;
  if (i < 2 * D)
    for (int c = w; c < nparts; c += 32) a += part[(int64_t)c * 2 * D + i];
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && i < 2 * D) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) s += red[q][lane];
    if (i < D) dgamma[i] = s;
    else dbeta[i - D] = s;
  }
}

static inline int blk_nv(int D) { return m3_ceil_div(D / 8, kBlkLanes); }
static inline int ln_bwd_ctas(int T) {
  int n = m3_ceil_div(T, kLnbWarps);
  if (n > 2 * kNumSMs) n = 2 * kNumSMs;
  return n < 1 ? 1 : n;
}

}  // namespace m3

using namespace m3;
typedef __nv_bfloat16 bf16;

static int blk_check(int T, int D) {
  if (T < 0 || D < 8) return M3_ERR_ARG;
  if (D % 8 != 0 || D > kBlkLanes * 8 * 8) return M3_ERR_SHAPE;
  const int nv = blk_nv(D);
  if (!(nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8)) return M3_ERR_SHAPE;
  return M3_OK;
}

#define M3_BLK_NV_SWITCH(...)                             \
  switch (nv) {                                           \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break; \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break; \
    case 8: { constexpr int NV = 8; __VA_ARGS__; } break; \
    default: return M3_ERR_SHAPE;                         \
  }

extern "C" int m3_ln_stats(const float* x, int T, int D, float eps, float* mean, float* rstd, m3_stream_t stream) {
  M3_CHECK_ARG(x && mean && rstd);
  int rc = blk_check(T, D);
  if (rc) return rc;
  M3_CHECK_ALIGN16(x);
  if (T == 0) return M3_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = blk_nv(D);
  const int grid = m3_ceil_div(T, kBlkTok);
  M3_BLK_NV_SWITCH((ln_stats_kernel<NV><<<grid, kBlkThreads, 0, st>>>(x, T, D, eps, mean, rstd)))
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" int m3_ln_fold_gate(const float* w_gate, const float* gamma, const float* beta, int D, int Dg, int E,
                               float* w_fold, float* gb, m3_stream_t stream) {
  M3_CHECK_ARG(w_gate && gamma && beta && w_fold && gb);
  M3_CHECK_ARG(D > 0 && Dg >= D && E > 0 && E <= 1024);
  ln_fold_gate_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(w_gate, gamma, beta, D, Dg, E, w_fold, gb);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" int m3_dispatch_fwd_ln(const float* x, const float* mean, const float* rstd, const float* gamma,
                                  const float* beta, const int32_t* pos, const int32_t* counts,
                                  const int32_t* offsets, int T, int K, int D, int E, void* xq, int xq_dtype,
                                  m3_stream_t stream) {
  M3_CHECK_ARG(x && mean && rstd && gamma && beta && pos && counts && offsets && xq);
  M3_CHECK_ARG(K >= 1 && E >= 1);
  int rc = blk_check(T, D);
  if (rc) return rc;
  M3_CHECK_ALIGN16(x); M3_CHECK_ALIGN16(xq); M3_CHECK_ALIGN16(gamma); M3_CHECK_ALIGN16(beta);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = blk_nv(D);
  const int tok_ctas = m3_ceil_div(T, kBlkTok);
  const int grid = tok_ctas + E;
  if (xq_dtype == M3_BF16) {
    M3_BLK_NV_SWITCH((dispatch_fwd_ln_kernel<bf16, NV><<<grid, kBlkThreads, 0, st>>>(x, mean, rstd, gamma, beta, pos, counts, offsets, T, K, D, tok_ctas, (bf16*)xq)))
  } else if (xq_dtype == M3_F32) {
    M3_BLK_NV_SWITCH((dispatch_fwd_ln_kernel<float, NV><<<grid, kBlkThreads, 0, st>>>(x, mean, rstd, gamma, beta, pos, counts, offsets, T, K, D, tok_ctas, (float*)xq)))
  } else {
    return M3_ERR_UNSUPPORTED;
  }
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" int m3_combine_fwd_res(const void* yq, int yq_dtype, const int32_t* pos, const float* score,
                                  const float* residual, int T, int K, int D, float* out, m3_stream_t stream) {
  M3_CHECK_ARG(yq && pos && score && residual && out && K >= 1);
  int rc = blk_check(T, D);
  if (rc) return rc;
  M3_CHECK_ALIGN16(yq); M3_CHECK_ALIGN16(residual); M3_CHECK_ALIGN16(out);
  if (T == 0) return M3_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = blk_nv(D);
  const int grid = m3_ceil_div(T, kBlkTok);
  if (yq_dtype == M3_BF16) {
    M3_BLK_NV_SWITCH((combine_fwd_res_kernel<bf16, NV><<<grid, kBlkThreads, 0, st>>>((const bf16*)yq, pos, score, residual, T, K, D, out)))
  } else if (yq_dtype == M3_F32) {
    M3_BLK_NV_SWITCH((combine_fwd_res_kernel<float, NV><<<grid, kBlkThreads, 0, st>>>((const float*)yq, pos, score, residual, T, K, D, out)))
  } else {
    return M3_ERR_UNSUPPORTED;
  }
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" size_t m3_ln_bwd_workspace_bytes(int T, int D) {
  return (size_t)ln_bwd_ctas(T) * 2 * (size_t)D * sizeof(float);
}

extern "C" int m3_ln_bwd_res(const float* dxn, const float* x, const float* mean, const float* rstd,
                             const float* gamma, const float* dres, int T, int D, float* dx, float* dgamma,
                             float* dbeta, void* workspace, size_t workspace_bytes, m3_stream_t stream) {
  M3_CHECK_ARG(dxn && x && mean && rstd && gamma && dres && dx && dgamma && dbeta && workspace);
  int rc = blk_check(T, D);
  if (rc) return rc;
  M3_CHECK_ARG(T > 0);
  M3_CHECK_ALIGN16(dxn); M3_CHECK_ALIGN16(x); M3_CHECK_ALIGN16(dres); M3_CHECK_ALIGN16(dx); M3_CHECK_ALIGN16(gamma);
  if (workspace_bytes < m3_ln_bwd_workspace_bytes(T, D)) return M3_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = ln_bwd_ctas(T);
  const size_t smem = 2 * (size_t)D * sizeof(float);
  float* part = static_cast<float*>(workspace);
  switch (m3_ceil_div(D / 4, 32)) {
#define M3_LNB_CASE(Q) \
  case Q: ln_bwd_res_kernel<Q><<<grid, kBlkThreads, smem, st>>>(dxn, x, mean, rstd, gamma, dres, T, D, dx, part); break;
    M3_LNB_CASE(1) M3_LNB_CASE(2) M3_LNB_CASE(3) M3_LNB_CASE(4) M3_LNB_CASE(5) M3_LNB_CASE(6) M3_LNB_CASE(7) M3_LNB_CASE(8)
    default: return M3_ERR_SHAPE;
  }
  M3_LAUNCH_CHECK();
  ln_bwd_reduce_kernel<<<m3_ceil_div(2 * D, 32), 1024, 0, st>>>(part, grid, D, dgamma, dbeta);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

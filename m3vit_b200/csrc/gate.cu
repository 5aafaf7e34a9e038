// Fused noisy_vmoe router (forward + backward) for sm_100a.
//
// Replaces NoisyGate_VMoE.forward of the reference
// (/root/reference/models/moe/origin/noisy_gate_vmoe.py:168-297; ckpt twin
// models/moe/ckpt/noisy_gate_vmoe.py:80-264): 6-8 tiny ATen kernels (GEMM with
// N=16, randn add, softmax, topk, slice, scatter, reductions) become ONE kernel.
//
// Forward layout: a warp owns TOK_W = (32/EG)*TM tokens, EG = E/4 lanes per token,
// every lane accumulates TM tokens x 4 experts in registers with sequential fp32
// FMAs over d (deterministic summation order).  x is staged through shared memory
// in 32-column chunks with cp.async double buffering; w_gate chunks are shared by
// the CTA.  Softmax / top-(K+1) run on the registers with warp-shuffle reductions
// across the EG lanes of a token (lowest index wins ties).
//
// Roofline: HBM-bound on reading x once (T*D*el bytes) for large T, FFMA-bound
// below that; 2*Dg*E flop per token.
#include "common.cuh"
#include "philox.cuh"

namespace m3 {

constexpr int kGateDC = 32;  // columns of x per smem chunk
constexpr int kGateBigCfg = 5;  // tile configuration for big batches and E >= 16 (see gate_cfg_id)
constexpr int kGateStages = 4;  // cp.async ring depth (chunks in flight per CTA)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename XT>
struct GateRow;  // smem row geometry of one x chunk
template <>
struct GateRow<float> {
  static constexpr int kBytes = kGateDC * 4 + 16;  // +16 B pad: conflict-free LDS.128 over 8 rows
  __device__ static __forceinline__ float4 ld4(const unsigned char* row, int d) {
    return *reinterpret_cast<const float4*>(row + d * 4);
  }
};
template <>
struct GateRow<__nv_bfloat16> {
  static constexpr int kBytes = kGateDC * 2 + 16;
  __device__ static __forceinline__ float4 ld4(const unsigned char* row, int d) {
    uint2 u = *reinterpret_cast<const uint2*>(row + d * 2);
    float2 a = bf16x2_to_float2(u.x), b = bf16x2_to_float2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

template <int E, int TM, int NW, int EPL = 4>
struct GateCfg {
  static constexpr int EG = E / EPL;        // lanes per token (EPL experts per lane: 4, or 8 for the big-batch tile)
  static constexpr int TG = 32 / EG;        // tokens per warp "row"
  static constexpr int TOK_W = TG * TM;     // tokens per warp
  static constexpr int TOK_CTA = TOK_W * NW;
};

template <int E, int TM, int NW, typename XT, int EPL = 4>
__global__ void __launch_bounds__(NW * 32)
gate_fwd_kernel(const XT* __restrict__ x, int64_t ldx, const float* __restrict__ task_feat,
                const float* __restrict__ w_gate, const float* __restrict__ noise, float noise_stddev,
                int T, int D, int Dt, int K, int K1, int64_t* __restrict__ idx,
                int32_t* __restrict__ idx_full, float* __restrict__ score, float* __restrict__ top_vals,
                float* __restrict__ clean_logits, float* __restrict__ noisy_logits,
                float* __restrict__ gates, float* __restrict__ imp_partial,
                int32_t* __restrict__ load_partial, const float* __restrict__ ln_mean,
                const float* __restrict__ ln_rstd, const float* __restrict__ ln_gb, const RngState* __restrict__ rng) {
  pdl_wait();
  pdl_trigger();
  using C = GateCfg<E, TM, NW, EPL>;
  using Row = GateRow<XT>;
  constexpr int EG = C::EG, TG = C::TG, TOK_W = C::TOK_W;
  constexpr int EQ = EPL / 4;                      // float4 groups of experts per lane
  static_assert(EPL == 4 || EPL == 8, "experts per lane");
  constexpr int ROWB = Row::kBytes;
  constexpr int XS_STAGE = TOK_W * ROWB;           // bytes per warp per stage
  constexpr int WS_STAGE = kGateDC * E * 4;        // bytes per stage
  constexpr int CHUNK_VECS = kGateDC * (int)sizeof(XT) / 16;  // 16-B vectors per x row chunk

  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int S = kGateStages;
  unsigned char* ws = smem;                                   // [S][DC][E] fp32
  unsigned char* xs = smem + S * WS_STAGE;                    // [NW][S][TOK_W][ROWB]
  float* red_imp = reinterpret_cast<float*>(xs + NW * S * XS_STAGE);  // [NW][E]
  int* red_load = reinterpret_cast<int*>(red_imp + NW * E);           // [NW][E]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int eg = lane % EG, tg = lane / EG;
  const int tok_w0 = blockIdx.x * C::TOK_CTA + warp * TOK_W;
  unsigned char* my_xs = xs + warp * S * XS_STAGE;

  const int NC = D / kGateDC;
  auto issue = [&](int c, int stage) {
    // x: TOK_W rows x CHUNK_VECS 16-B vectors, rows clamped to T-1 (never OOB)
    for (int v = lane; v < TOK_W * CHUNK_VECS; v += 32) {
      int r = v / CHUNK_VECS, q = v % CHUNK_VECS;
      int t = min(tok_w0 + r, T - 1);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(x + (int64_t)t * ldx + c * kGateDC) + q * 16;
      cp_async16(my_xs + stage * XS_STAGE + r * ROWB + q * 16, src);
    }
    // w_gate rows [c*DC, c*DC+DC) are contiguous: DC*E floats
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(w_gate + (int64_t)c * kGateDC * E);
    for (int v = threadIdx.x; v < WS_STAGE / 16; v += NW * 32) cp_async16(ws + stage * WS_STAGE + v * 16, wsrc + v * 16);
    cp_async_commit();
  };

  float acc[TM][EPL];
#pragma unroll
  for (int j = 0; j < TM; ++j)
#pragma unroll
    for (int c = 0; c < EPL; ++c) acc[j][c] = 0.f;

  // S-deep ring: chunks c .. c+S-2 are in flight while chunk c is consumed.  One barrier per chunk:
  // it publishes chunk c and proves that every warp has finished chunk c-1, whose stage is refilled.
#pragma unroll
  for (int p = 0; p < S - 1; ++p) {
    if (p < NC) issue(p, p);
    else cp_async_commit();
  }
  for (int c = 0; c < NC; ++c) {
    cp_async_wait<S - 2>();
    __syncthreads();
    if (c + S - 1 < NC) issue(c + S - 1, (c + S - 1) % S);
    else cp_async_commit();
    const unsigned char* xst = my_xs + (c % S) * XS_STAGE;
    const float* wst = reinterpret_cast<const float*>(ws + (c % S) * WS_STAGE);
#pragma unroll 2
    for (int d4 = 0; d4 < kGateDC; d4 += 4) {
      // EPL = 8, TM = 4: 8 + 4 LDS.128 per 128 FMAs instead of 4 + 2 per 32 - the kernel is bound by the 4 clk an
      // LDS.128 takes on the return path, not by the FMAs
      float4 w[4][EQ];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < EQ; ++h) w[i][h] = *reinterpret_cast<const float4*>(wst + (d4 + i) * E + eg * EPL + 4 * h);
#pragma unroll
      for (int j = 0; j < TM; ++j) {
        float4 xv = Row::ld4(xst + (tg + TG * j) * ROWB, d4);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int h = 0; h < EQ; ++h) {
            acc[j][4 * h + 0] = fmaf(xa[i], w[i][h].x, acc[j][4 * h + 0]);
            acc[j][4 * h + 1] = fmaf(xa[i], w[i][h].y, acc[j][4 * h + 1]);
            acc[j][4 * h + 2] = fmaf(xa[i], w[i][h].z, acc[j][4 * h + 2]);
            acc[j][4 * h + 3] = fmaf(xa[i], w[i][h].w, acc[j][4 * h + 3]);
          }
        }
      }
    }
  }

  // Block-level fusion (block.cu): x is the RAW residual stream and w_gate is gamma-folded and
  // column-centred, so  LayerNorm(x) @ W = rstd * (x @ W') + B,   B = ln_gb[E:2E]
  if (ln_mean != nullptr) {
    float Bv[EPL];
#pragma unroll
    for (int c = 0; c < EPL; ++c) Bv[c] = __ldg(ln_gb + E + eg * EPL + c);
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int t = min(tok_w0 + tg + TG * j, T - 1);
      const float rs = __ldg(ln_rstd + t);
#pragma unroll
      for (int c = 0; c < EPL; ++c) acc[j][c] = fmaf(rs, acc[j][c], Bv[c]);
    }
  }

  // task-conditioned router: constant contribution of the task feature rows
  // (the reference concatenates it onto every token, custom_moe_layer.py:176-179)
  if (Dt > 0) {
    float tb[EPL];
#pragma unroll
    for (int c = 0; c < EPL; ++c) tb[c] = 0.f;
    for (int j = 0; j < Dt; ++j) {
      float f = __ldg(task_feat + j);
#pragma unroll
      for (int c = 0; c < EPL; ++c) tb[c] = fmaf(f, __ldg(w_gate + (int64_t)(D + j) * E + eg * EPL + c), tb[c]);
    }
#pragma unroll
    for (int j = 0; j < TM; ++j)
#pragma unroll
      for (int c = 0; c < EPL; ++c) acc[j][c] += tb[c];
  }

  float imp[EPL];
  int ld[EPL];
#pragma unroll
  for (int c = 0; c < EPL; ++c) { imp[c] = 0.f; ld[c] = 0; }

#pragma unroll
  for (int j = 0; j < TM; ++j) {
    const int t = tok_w0 + tg + TG * j;
    const bool valid = t < T;
    const int64_t te = (int64_t)(valid ? t : 0) * E + eg * EPL;
    float z[EPL];
#pragma unroll
    for (int c = 0; c < EPL; ++c) z[c] = acc[j][c];
    if (valid) {
#pragma unroll
      for (int h = 0; h < EQ; ++h)
        *reinterpret_cast<float4*>(clean_logits + te + 4 * h) = make_float4(z[4 * h], z[4 * h + 1], z[4 * h + 2], z[4 * h + 3]);
    }
    if (noise != nullptr || rng != nullptr) {
      // noisy = clean + N(0,1) * stddev (noisy_gate_vmoe.py:226): the normals come from the caller (torch.randn_like, the
      // reference's stream) or are drawn here, four per expert quad, from the counter-based generator (philox.cuh)
#pragma unroll
      for (int h = 0; h < EQ; ++h) {
        float nn[4] = {0.f, 0.f, 0.f, 0.f};
        if (noise != nullptr) {
          if (valid) {
            const float4 n = __ldg(reinterpret_cast<const float4*>(noise + te + 4 * h));
            nn[0] = n.x; nn[1] = n.y; nn[2] = n.z; nn[3] = n.w;
          }
        } else {
          normal4(*rng, (uint32_t)(valid ? t : 0), (uint32_t)(eg * EQ + h), nn);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) z[4 * h + c] += nn[c] * noise_stddev;
      }
      if (valid && noisy_logits != nullptr) {
#pragma unroll
        for (int h = 0; h < EQ; ++h)
          *reinterpret_cast<float4*>(noisy_logits + te + 4 * h) = make_float4(z[4 * h], z[4 * h + 1], z[4 * h + 2], z[4 * h + 3]);
      }
    }
    // softmax over all E experts of this token (EG lanes x EPL)
    float m = z[0];
#pragma unroll
    for (int c = 1; c < EPL; ++c) m = fmaxf(m, z[c]);
#pragma unroll
    for (int o = EG / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float p[EPL];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < EPL; ++c) { p[c] = expf(z[c] - m); s += p[c]; }
#pragma unroll
    for (int o = EG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
    for (int c = 0; c < EPL; ++c) p[c] = p[c] / s;

    // top-K1 on the probabilities, descending, lowest index wins ties
    unsigned taken = 0, takenK = 0;
    for (int r = 0; r < K1; ++r) {
      float bv = -1.f;
      int bi = 0x7fffffff;
#pragma unroll
      for (int c = 0; c < EPL; ++c)
        if (!((taken >> c) & 1u) && p[c] > bv) { bv = p[c]; bi = eg * EPL + c; }
#pragma unroll
      for (int o = EG / 2; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (bi >= E) { bi = r; bv = 0.f; }  // NaN row: keep indices in range
      if ((bi / EPL) == eg) {
        taken |= 1u << (bi % EPL);
        if (r < K) takenK |= 1u << (bi % EPL);
      }
      if (valid && eg == 0) {
        idx_full[(int64_t)t * K1 + r] = bi;
        top_vals[(int64_t)t * K1 + r] = bv;
        if (r < K) {
          idx[(int64_t)t * K + r] = bi;
          score[(int64_t)t * K + r] = bv;
        }
      }
    }
    float g[EPL];
#pragma unroll
    for (int c = 0; c < EPL; ++c) {
      const bool sel = valid && ((takenK >> c) & 1u);
      g[c] = sel ? p[c] : 0.f;
      imp[c] += g[c];
      ld[c] += (sel && p[c] > 0.f) ? 1 : 0;
    }
    if (valid && gates != nullptr) {
#pragma unroll
      for (int h = 0; h < EQ; ++h)
        *reinterpret_cast<float4*>(gates + te + 4 * h) = make_float4(g[4 * h], g[4 * h + 1], g[4 * h + 2], g[4 * h + 3]);
    }
  }

  // importance / load partials: fixed-order reduction -> deterministic
#pragma unroll
  for (int c = 0; c < EPL; ++c) {
#pragma unroll
    for (int o = EG; o < 32; o <<= 1) {
      imp[c] += __shfl_xor_sync(0xffffffffu, imp[c], o);
      ld[c] += __shfl_xor_sync(0xffffffffu, ld[c], o);
    }
  }
  if (tg == 0) {
#pragma unroll
    for (int c = 0; c < EPL; ++c) {
      red_imp[warp * E + eg * EPL + c] = imp[c];
      red_load[warp * E + eg * EPL + c] = ld[c];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += NW * 32) {
    float a = 0.f;
    int b = 0;
    for (int w = 0; w < NW; ++w) { a += red_imp[w * E + e]; b += red_load[w * E + e]; }
    imp_partial[(int64_t)blockIdx.x * E + e] = a;
    load_partial[(int64_t)blockIdx.x * E + e] = b;
  }
}

template <int E, int TM, int NW, typename XT, int EPL = 4>
static size_t gate_fwd_smem() {
  using C = GateCfg<E, TM, NW, EPL>;
  return (size_t)kGateStages * kGateDC * E * 4 + (size_t)NW * kGateStages * C::TOK_W * GateRow<XT>::kBytes +
         (size_t)NW * E * 8;
}

// Three tile configurations; pick the largest that still puts >= 16 warps on every SM (the kernel
// hides its smem / HBM latency with warps, not with ILP).
//   0: 1 warp x TG*2 tokens (small T)   1: 4 warps x TG*2   2: 4 warps x TG*4   3: 4 warps x TG*8
//   4 / 5 (E >= 16): 8 experts per lane - 1 warp x TG*4 tokens / 2 warps x TG*2 tokens per CTA: half the LDS per FMA
template <int E>
static int gate_cfg_id(int T) {
  if (g_knobs[M3_KNOB_GATE_CFG] > 0) {                                       // forced (A/B measurement)
    const int id = g_knobs[M3_KNOB_GATE_CFG] - 1;
    return (id >= 4 && E < 16) ? 1 : id;
  }
  // measured at T = 38 432, E = 16 (tools/ab_gate.py): cfg 1 32.9 us, cfg 4 35.7 us, cfg 5 29.9 us
  if (E >= 16 && T >= 4 * kNumSMs * GateCfg<(E >= 16 ? E : 16), 2, 2, 8>::TOK_CTA) return kGateBigCfg;
  if (T >= 16 * kNumSMs * GateCfg<E, 8, 4>::TOK_W) return 3;
  if (T >= 16 * kNumSMs * GateCfg<E, 4, 4>::TOK_W) return 2;   // (T = 38 432, E = 16: TM = 2 35 us, TM = 4 43 us, TM = 8 60 us)
  if (T >= 16 * kNumSMs * GateCfg<E, 2, 4>::TOK_W) return 1;
  return 0;
}

template <int E, int TM, int NW, typename XT, int EPL = 4>
static int launch_gate_fwd(const void* x, int64_t ldx, const float* task_feat, const float* w_gate,
                           const float* noise, float noise_stddev, int T, int D, int Dt, int K, int K1,
                           int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean,
                           float* noisy, float* gates, float* imp_partial, int32_t* load_partial,
                           const float* ln_mean, const float* ln_rstd, const float* ln_gb, const RngState* rng,
                           cudaStream_t st) {
  using C = GateCfg<E, TM, NW, EPL>;
  size_t smem = gate_fwd_smem<E, TM, NW, XT, EPL>();
  auto kern = gate_fwd_kernel<E, TM, NW, XT, EPL>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int grid = m3_ceil_div(T, C::TOK_CTA);
  launch_k(kern, grid, NW * 32, smem, st, static_cast<const XT*>(x), ldx, task_feat, w_gate, noise, noise_stddev, T,
           D, Dt, K, K1, idx, idx_full, score, top_vals, clean, noisy, gates, imp_partial, load_partial, ln_mean,
           ln_rstd, ln_gb, rng);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

template <int E>
static int gate_tokens_per_cta(int T) {
  const int id = gate_cfg_id<E>(T);
  constexpr int E8 = E >= 16 ? E : 16;       // (the 8-experts-per-lane tiles exist for E >= 16 only)
  return id == 5 ? GateCfg<E8, 2, 2, 8>::TOK_CTA : id == 4 ? GateCfg<E8, 4, 1, 8>::TOK_CTA
       : id == 3 ? GateCfg<E, 8, 4>::TOK_CTA : id == 2 ? GateCfg<E, 4, 4>::TOK_CTA
       : id == 1 ? GateCfg<E, 2, 4>::TOK_CTA : GateCfg<E, 2, 1>::TOK_CTA;
}

// ------------------------------------------------------------------ backward
// dz[t,:] = softmax-Jacobian applied to the gradient of every selected probability.
template <int E>
__global__ void __launch_bounds__(256)
gate_bwd_dz_kernel(const float* __restrict__ logits, const int32_t* __restrict__ idx_full, int T, int K, int K1,
                   const float* __restrict__ dscore, const float* __restrict__ dtop,
                   const float* __restrict__ dgates, const float* __restrict__ dimp,
                   const float* __restrict__ dclean, const float* __restrict__ dnoisy,
                   const float* __restrict__ importance, const float* __restrict__ dcv,
                   float* __restrict__ dz) {
  constexpr int EG = E / 4;
  // gradient of cv^2(importance) w.r.t. importance[e], scaled by d(cv_loss) (added to dimp)
  pdl_wait();
  pdl_trigger();
  __shared__ float gimp[E];
  if (dcv != nullptr) {
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int e = 0; e < E; ++e) m += importance[e];
      m /= (float)E;
      float var = 0.f;
      for (int e = 0; e < E; ++e) { const float d = importance[e] - m; var = fmaf(d, d, var); }
      var /= (float)(E > 1 ? E - 1 : 1);
      const float den = m * m + 1e-10f;
      const float up = __ldg(dcv);
      for (int e = 0; e < E; ++e)
        gimp[e] = E > 1 ? up * (2.f * (importance[e] - m) / ((float)(E - 1) * den) - var * 2.f * m / ((float)E * den * den))
                        : 0.f;
    }
    __syncthreads();
  }
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int eg = (int)(gid % EG);
  const int64_t tt = gid / EG;
  const bool valid = tt < T;
  const int64_t t = valid ? tt : (T - 1);
  const int64_t te = t * E + eg * 4;
  float4 zv = __ldg(reinterpret_cast<const float4*>(logits + te));
  float z[4] = {zv.x, zv.y, zv.z, zv.w};
  float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
#pragma unroll
  for (int o = EG / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float p[4], s = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) { p[c] = expf(z[c] - m); s += p[c]; }
#pragma unroll
  for (int o = EG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
  for (int c = 0; c < 4; ++c) p[c] = p[c] / s;

  float dp[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < K1; ++r) {
    const int e = __ldg(idx_full + t * K1 + r);
    float g = 0.f;
    if (dtop != nullptr) g += __ldg(dtop + t * K1 + r);
    if (r < K) {
      if (dscore != nullptr) g += __ldg(dscore + t * K + r);
      if (dgates != nullptr) g += __ldg(dgates + t * E + e);
      if (dimp != nullptr) g += __ldg(dimp + e);
      if (dcv != nullptr) g += gimp[e];
    }
    if ((e >> 2) == eg) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if ((e & 3) == c) dp[c] += g;
    }
  }
  float dot = p[0] * dp[0] + p[1] * dp[1] + p[2] * dp[2] + p[3] * dp[3];
#pragma unroll
  for (int o = EG / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  float o4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) o4[c] = p[c] * (dp[c] - dot);
  if (dclean != nullptr) {
    float4 a = __ldg(reinterpret_cast<const float4*>(dclean + te));
    o4[0] += a.x; o4[1] += a.y; o4[2] += a.z; o4[3] += a.w;
  }
  if (dnoisy != nullptr) {
    float4 a = __ldg(reinterpret_cast<const float4*>(dnoisy + te));
    o4[0] += a.x; o4[1] += a.y; o4[2] += a.z; o4[3] += a.w;
  }
  if (valid) *reinterpret_cast<float4*>(dz + te) = make_float4(o4[0], o4[1], o4[2], o4[3]);
}

// partial dW[chunk][d][e] = sum_{t in chunk} x[t,d] * dz[t,e]; thread tile 4 d x EW experts.
constexpr int kDwSub = 160;   // tokens of dz staged in shared memory per pass (T = 38 432 on 296 chunks: 130 per chunk = one pass)

template <int EW, typename XT, bool LN, int TB>
__global__ void __launch_bounds__(TB == 16 ? 384 : TB == 8 ? 512 : 1024)
gate_bwd_dw_kernel(const XT* __restrict__ x, int64_t ldx, const float* __restrict__ dz, int T,
                                   int D, int E, int tok_per_chunk, float* __restrict__ part,
                                   float* __restrict__ cs_part, const float* __restrict__ ln_mean,
                                   const float* __restrict__ ln_rstd, const float* __restrict__ ln_gamma,
                                   const float* __restrict__ ln_beta) {
  const int ngrp = blockDim.x / (D / 4);          // expert groups per CTA (EB / EW)
  pdl_wait();
  pdl_trigger();
  const int dq = threadIdx.x % (D / 4);
  const int eh = threadIdx.x / (D / 4);
  const int e0 = blockIdx.y * (EW * ngrp) + eh * EW;
  const int t0 = blockIdx.x * tok_per_chunk;
  const int t1 = min(T, t0 + tok_per_chunk);
  float acc[4][EW];
  float cs[EW];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < EW; ++c) acc[i][c] = 0.f;
#pragma unroll
  for (int c = 0; c < EW; ++c) cs[c] = 0.f;
  // Block-level fusion: x is the raw residual stream, normalised on load
  float lg[4] = {1.f, 1.f, 1.f, 1.f}, lb[4] = {0.f, 0.f, 0.f, 0.f};
  if constexpr (LN) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { lg[i] = __ldg(ln_gamma + dq * 4 + i); lb[i] = __ldg(ln_beta + dq * 4 + i); }
  }
  // dz (and the LayerNorm statistics) of kDwSub tokens are staged in shared memory once per CTA pass: every thread of
  // an expert group reads the same values, and keeping them out of the register file leaves room for TB tokens of x in
  // flight per thread (the kernel is bound by the latency of those loads, not by bandwidth or FMAs).
  __shared__ __align__(16) float dzs[kDwSub * 16];
  __shared__ float lns[LN ? 2 * kDwSub : 2];
  const int EB = EW * ngrp;                           // experts handled by this CTA (<= 16)
  const int eb0 = blockIdx.y * EB;
  for (int ts = t0; ts < t1; ts += kDwSub) {
    const int n = min(kDwSub, t1 - ts);
    __syncthreads();                                  // the previous pass has been consumed
    for (int i = threadIdx.x; i < n * (EB / 4); i += blockDim.x) {
      const int tok = i / (EB / 4), q = i % (EB / 4);
      *reinterpret_cast<float4*>(dzs + tok * 16 + q * 4) =
          __ldg(reinterpret_cast<const float4*>(dz + (int64_t)(ts + tok) * E + eb0 + q * 4));
    }
    if constexpr (LN) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        lns[2 * i] = __ldg(ln_mean + ts + i);
        lns[2 * i + 1] = __ldg(ln_rstd + ts + i);
      }
    }
    __syncthreads();
    for (int tb = 0; tb < n; tb += TB) {
      float xv[TB][4];
#pragma unroll
      for (int u = 0; u < TB; ++u) {                  // all loads of the batch first
        const int t = ts + min(tb + u, n - 1);
        if constexpr (sizeof(XT) == 4) {
          float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)t * ldx + dq * 4));
          xv[u][0] = v.x; xv[u][1] = v.y; xv[u][2] = v.z; xv[u][3] = v.w;
        } else {
          uint2 w = __ldg(reinterpret_cast<const uint2*>(x + (int64_t)t * ldx + dq * 4));
          float2 a = bf16x2_to_float2(w.x), b = bf16x2_to_float2(w.y);
          xv[u][0] = a.x; xv[u][1] = a.y; xv[u][2] = b.x; xv[u][3] = b.y;
        }
      }
#pragma unroll
      for (int u = 0; u < TB; ++u) {                  // t ascending: deterministic summation order
        if (tb + u < n) {
          float dv[EW];
#pragma unroll
          for (int c = 0; c < EW; c += 4) {
            const float4 v = *reinterpret_cast<const float4*>(dzs + (tb + u) * 16 + eh * EW + c);
            dv[c] = v.x; dv[c + 1] = v.y; dv[c + 2] = v.z; dv[c + 3] = v.w;
          }
          if constexpr (LN) {
            const float mu = lns[2 * (tb + u)], rs = lns[2 * (tb + u) + 1];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[u][i] = fmaf((xv[u][i] - mu) * rs, lg[i], lb[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < EW; ++c) acc[i][c] = fmaf(xv[u][i], dv[c], acc[i][c]);
          if (dq == 0) {
#pragma unroll
            for (int c = 0; c < EW; ++c) cs[c] += dv[c];
          }
        }
      }
    }
  }
  float* dst = part + ((int64_t)blockIdx.x * D + dq * 4) * E + e0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < EW; c += 4)
      *reinterpret_cast<float4*>(dst + (int64_t)i * E + c) = make_float4(acc[i][c], acc[i][c + 1], acc[i][c + 2], acc[i][c + 3]);
  if (dq == 0) {
#pragma unroll
    for (int c = 0; c < EW; ++c) cs_part[(int64_t)blockIdx.x * E + e0 + c] = cs[c];
  }
}

// The same partial sums with x, dz (and the LayerNorm statistics) streamed through a cp.async ring of S stages x
// kDwTS tokens.  The register-batch kernel above is bound by LOAD LATENCY: all warps of a CTA issue a batch of loads,
// wait ~1.5 us for DRAM, then compute for ~0.4 us, 9 times per chunk, and two 6-warp CTAs per SM cannot hide that
// (27 us at T = 38 432 against 9 us of HBM time; two register half-batches were no better: a thread cannot hold the
// 4-5 batches in flight that the latency asks for).  A shared-memory ring can: S - 1 stages are always in flight,
// regardless of the register file, and both expert groups of the CTA read ONE copy of x.  Same thread tile, same
// summation order (t ascending) - bit-identical partials.
constexpr int kDwTS = 16;     // tokens per stage

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}

template <typename XT, bool LN>
__host__ __device__ constexpr int dw_stage_bytes(int D) {
  return kDwTS * D * (int)sizeof(XT) + kDwTS * 16 * 4 + (LN ? kDwTS * 2 * 4 : 0);
}

template <int EW, typename XT, bool LN, int S, int DPT>
__global__ void __launch_bounds__(512)
gate_bwd_dw_ring_kernel(const XT* __restrict__ x, int64_t ldx, const float* __restrict__ dz, int T, int D, int E,
                        int tok_per_chunk, float* __restrict__ part, float* __restrict__ cs_part,
                        const float* __restrict__ ln_mean, const float* __restrict__ ln_rstd,
                        const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  const int nthr = blockDim.x;
  static_assert(DPT == 2 || DPT == 4, "columns of x per thread");
  const int ngrp = nthr / (D / DPT);              // expert groups per CTA (EB / EW)
  pdl_wait();
  pdl_trigger();
  const int dq = threadIdx.x % (D / DPT);         // thread tile: DPT columns x EW experts (DPT = 2: twice the warps per
  const int eh = threadIdx.x / (D / DPT);         // SM behind the same stalls - the kernel is latency-, not issue-bound)
  const int EB = EW * ngrp;                       // experts handled by this CTA (<= 16)
  const int eb0 = blockIdx.y * EB;
  const int e0 = eb0 + eh * EW;
  const int t0 = blockIdx.x * tok_per_chunk;
  const int t1 = min(T, t0 + tok_per_chunk);
  const int rowb = D * (int)sizeof(XT);
  const int stage_bytes = dw_stage_bytes<XT, LN>(D);
  const int nst = (t1 - t0 + kDwTS - 1) / kDwTS;

  const int nvec = rowb / 16;
  const int xq = threadIdx.x % nvec, xr = threadIdx.x / nvec, xstep = nthr / nvec;
  auto issue = [&](int c) {
    unsigned char* st = dw_smem + (c % S) * stage_bytes;
    const int tb = t0 + c * kDwTS;
    // nthr is a multiple of the 16-byte vectors of a row (D / 4 threads per expert group): a thread keeps its column
    // vector xq and walks the rows xr, xr + xstep, ... - no divisions in the copy loop
    for (int r = xr; r < kDwTS; r += xstep) {
      const int t = min(tb + r, t1 - 1);          // rows past the chunk re-read its last token (never consumed)
      cp_async16(st + r * rowb + xq * 16, reinterpret_cast<const unsigned char*>(x + (int64_t)t * ldx) + xq * 16);
    }
    float* dzs = reinterpret_cast<float*>(st + kDwTS * rowb);
    for (int v = threadIdx.x; v < kDwTS * (EB / 4); v += nthr) {
      const int r = v / (EB / 4), q = v % (EB / 4);
      const int t = min(tb + r, t1 - 1);
      cp_async16(dzs + r * 16 + q * 4, dz + (int64_t)t * E + eb0 + q * 4);
    }
    if constexpr (LN) {
      float* lns = dzs + kDwTS * 16;
      for (int v = threadIdx.x; v < kDwTS; v += nthr) {
        const int t = min(tb + v, t1 - 1);
        cp_async4(lns + 2 * v, ln_mean + t);
        cp_async4(lns + 2 * v + 1, ln_rstd + t);
      }
    }
    cp_async_commit();
  };

  // accumulators as fp32 PAIRS of neighbouring experts: fma.rn.f32x2 is the same IEEE fma per half, at half the
  // instructions (the kernel is issue-bound: ncu counted 88 warp instructions per warp-token for 32 FMAs' worth of work)
  f32x2 acc2[DPT][EW / 2];
  float cs1 = 0.f;
#pragma unroll
  for (int i = 0; i < DPT; ++i)
#pragma unroll
    for (int c = 0; c < EW / 2; ++c) acc2[i][c] = pk2(0.f, 0.f);
  float lg[DPT], lb[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) { lg[i] = 1.f; lb[i] = 0.f; }
  if constexpr (LN) {
#pragma unroll
    for (int i = 0; i < DPT; ++i) { lg[i] = __ldg(ln_gamma + dq * DPT + i); lb[i] = __ldg(ln_beta + dq * DPT + i); }
  }

#pragma unroll
  for (int p0 = 0; p0 < S - 1; ++p0) {
    if (p0 < nst) issue(p0);
    else cp_async_commit();
  }
  for (int c = 0; c < nst; ++c) {
    cp_async_wait<S - 2>();
    __syncthreads();              // stage c has landed for every thread; stage c - 1 (refilled next) has been consumed
    if (c + S - 1 < nst) issue(c + S - 1);
    else cp_async_commit();
    const unsigned char* st = dw_smem + (c % S) * stage_bytes;
    const float* dzs = reinterpret_cast<const float*>(st + kDwTS * rowb);
    const int n = min(kDwTS, t1 - (t0 + c * kDwTS));
    auto token = [&](int u) {
      float xv[DPT];
      if constexpr (sizeof(XT) == 4 && DPT == 4) {
        const float4 v = *reinterpret_cast<const float4*>(st + u * rowb + dq * 16);
        xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
      } else if constexpr (sizeof(XT) == 4) {
        const float2 v = *reinterpret_cast<const float2*>(st + u * rowb + dq * 8);
        xv[0] = v.x; xv[1] = v.y;
      } else if constexpr (DPT == 4) {
        const uint2 w = *reinterpret_cast<const uint2*>(st + u * rowb + dq * 8);
        const float2 a = bf16x2_to_float2(w.x), b = bf16x2_to_float2(w.y);
        xv[0] = a.x; xv[1] = a.y; xv[2] = b.x; xv[3] = b.y;
      } else {
        const float2 a = bf16x2_to_float2(*reinterpret_cast<const uint32_t*>(st + u * rowb + dq * 4));
        xv[0] = a.x; xv[1] = a.y;
      }
      f32x2 dv2[EW / 2];
#pragma unroll
      for (int k = 0; k < EW; k += 4) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(dzs + u * 16 + eh * EW + k);
        dv2[k / 2] = v.x; dv2[k / 2 + 1] = v.y;
      }
      if constexpr (LN) {
        const float* lns = dzs + kDwTS * 16;
        const float mu = lns[2 * u], rs = lns[2 * u + 1];
#pragma unroll
        for (int i = 0; i < DPT; ++i) xv[i] = fmaf((xv[i] - mu) * rs, lg[i], lb[i]);
      }
#pragma unroll
      for (int i = 0; i < DPT; ++i) {
        const f32x2 xx = pk2(xv[i], xv[i]);
#pragma unroll
        for (int k = 0; k < EW / 2; ++k) acc2[i][k] = fma2(xx, dv2[k], acc2[i][k]);
      }
    };
    // t ascending: deterministic summation order.  Full stages run without a per-token branch.
    if (n == kDwTS) {
#pragma unroll
      for (int u = 0; u < kDwTS; ++u) token(u);
    } else {
      for (int u = 0; u < n; ++u) token(u);
    }
    // column sums of dz (task-feature rows of dW): thread dq < EW of each expert group owns one expert, away from the
    // FMA stream (as predicated adds inside it they cost every warp 8 issue slots per token)
    if (dq < EW)
      for (int u = 0; u < n; ++u) cs1 += dzs[u * 16 + eh * EW + dq];
  }
  float* dst = part + ((int64_t)blockIdx.x * D + dq * DPT) * E + e0;
#pragma unroll
  for (int i = 0; i < DPT; ++i)
#pragma unroll
    for (int c = 0; c < EW; c += 4) {
      float a0, a1, a2, a3;
      unpk2(acc2[i][c / 2], a0, a1);
      unpk2(acc2[i][c / 2 + 1], a2, a3);
      *reinterpret_cast<float4*>(dst + (int64_t)i * E + c) = make_float4(a0, a1, a2, a3);
    }
  if (dq < EW) cs_part[(int64_t)blockIdx.x * E + e0 + dq] = cs1;
}

// dW[d][e] = sum_chunks part (fixed order); task rows from the column sums of dz.
// A block owns 32 consecutive outputs; its 32 warps stride over the chunks (coalesced 128-B
// reads, many loads in flight) and are combined in a fixed order -> deterministic.
__global__ void __launch_bounds__(1024)
gate_bwd_reduce_kernel(const float* __restrict__ part, const float* __restrict__ cs_part, int nchunk, int D, int Dt,
                       int E, const float* __restrict__ task_feat, const float* __restrict__ w_gate,
                       float* __restrict__ dw, float* __restrict__ dtask) {
  __shared__ float red[32][33];
  __shared__ float cs_s[128];
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n_main = (int64_t)D * E;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  if ((int64_t)blockIdx.x * 32 < n_main) {
    float a0 = 0.f, a1 = 0.f;
    if (i < n_main) {
      int c = w;
      for (; c + 32 < nchunk; c += 64) {
        a0 += part[(int64_t)c * n_main + i];
        a1 += part[(int64_t)(c + 32) * n_main + i];
      }
      if (c < nchunk) a0 += part[(int64_t)c * n_main + i];
    }
    red[w][lane] = a0 + a1;
    __syncthreads();
    if (w == 0 && i < n_main) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) s += red[q][lane];
      dw[i] = s;
    }
    return;
  }
  // trailing block: task-feature rows.  column sums of dz first (E <= 128)
  if (threadIdx.x < E) {
    float cs = 0.f;
    for (int c = 0; c < nchunk; ++c) cs += cs_part[(int64_t)c * E + threadIdx.x];
    cs_s[threadIdx.x] = cs;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < Dt * E + Dt; k += 1024) {
    if (k < Dt * E) {
      dw[n_main + k] = __ldg(task_feat + k / E) * cs_s[k % E];
    } else if (dtask != nullptr) {
      const int j = k - Dt * E;
      float a = 0.f;
      for (int e = 0; e < E; ++e) a = fmaf(__ldg(w_gate + (int64_t)(D + j) * E + e), cs_s[e], a);
      dtask[j] = a;
    }
  }
}

// standalone router dx: dxg[t, d] = sum_e dz[t,e] * w_gate[d,e]
__global__ void gate_bwd_dx_kernel(const float* __restrict__ dz, const float* __restrict__ w_gate, int T, int D,
                                   int E, float* __restrict__ dxg) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * (D / 4)) return;
  const int64_t t = i / (D / 4);
  const int dq = (int)(i % (D / 4));
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int e = 0; e < E; e += 4) {
    float4 g = __ldg(reinterpret_cast<const float4*>(dz + t * E + e));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 w = __ldg(reinterpret_cast<const float4*>(w_gate + (int64_t)(dq * 4 + r) * E + e));
      a[r] = fmaf(g.x, w.x, a[r]); a[r] = fmaf(g.y, w.y, a[r]);
      a[r] = fmaf(g.z, w.z, a[r]); a[r] = fmaf(g.w, w.w, a[r]);
    }
  }
  *reinterpret_cast<float4*>(dxg + t * D + dq * 4) = make_float4(a[0], a[1], a[2], a[3]);
}

static inline int gate_bwd_chunks(int T, int E) {
  const int yb = (E > 16) ? E / 16 : 1;
  int n = m3_ceil_div(T, 64);
  int cap = (2 * kNumSMs) / yb;     // T = 38 432: 1 / 2 / 4 / 8 chunks per SM -> 54.9 / 46.3 / 48.3 / 52.3 us for the three launches
  if (cap < 1) cap = 1;
  return n < cap ? (n < 1 ? 1 : n) : cap;
}

}  // namespace m3

using namespace m3;

extern "C" int m3_gate_num_partials(int T, int E) {
  switch (E) {
    case 4: return m3_ceil_div(T, gate_tokens_per_cta<4>(T));
    case 8: return m3_ceil_div(T, gate_tokens_per_cta<8>(T));
    case 16: return m3_ceil_div(T, gate_tokens_per_cta<16>(T));
    case 32: return m3_ceil_div(T, gate_tokens_per_cta<32>(T));
    case 64: return m3_ceil_div(T, gate_tokens_per_cta<64>(T));
    case 128: return m3_ceil_div(T, gate_tokens_per_cta<128>(T));
    default: return M3_ERR_SHAPE;
  }
}

#define M3_GATE_ARGS x, ldx, task_feat, w_gate, noise, noise_stddev, T, D, Dt, K, K1, idx, idx_full, score, \
                     top_vals, clean_logits, noisy_logits, gates, imp_partial, load_partial, ln_mean, ln_rstd, \
                     ln_gb, rng, st
#define M3_GATE_CASE_T(EE, XT)                                               \
  switch (gate_cfg_id<EE>(T)) {                                              \
    case 5: if constexpr (EE >= 16) return launch_gate_fwd<EE, 2, 2, XT, 8>(M3_GATE_ARGS); else return M3_ERR_SHAPE; \
    case 4: if constexpr (EE >= 16) return launch_gate_fwd<EE, 4, 1, XT, 8>(M3_GATE_ARGS); else return M3_ERR_SHAPE; \
    case 3: return launch_gate_fwd<EE, 8, 4, XT>(M3_GATE_ARGS);              \
    case 2: return launch_gate_fwd<EE, 4, 4, XT>(M3_GATE_ARGS);              \
    case 1: return launch_gate_fwd<EE, 2, 4, XT>(M3_GATE_ARGS);              \
    default: return launch_gate_fwd<EE, 2, 1, XT>(M3_GATE_ARGS);             \
  }
#define M3_GATE_CASE(EE)                                    \
  case EE:                                                  \
    if (x_dtype == M3_F32) { M3_GATE_CASE_T(EE, float) }    \
    else { M3_GATE_CASE_T(EE, __nv_bfloat16) }

static int gate_fwd_impl(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                         const float* noise, float noise_stddev, int T, int D, int Dt, int E, int K,
                         int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean_logits,
                         float* noisy_logits, float* gates, float* imp_partial, int32_t* load_partial,
                         const float* ln_mean, const float* ln_rstd, const float* ln_gb, const void* rng_state,
                         m3_stream_t stream) {
  M3_CHECK_ARG(x && w_gate && idx && idx_full && score && top_vals && clean_logits && imp_partial && load_partial);
  M3_CHECK_ARG(!(noise && rng_state));
  if (rng_state) M3_CHECK_ALIGN16(rng_state);
  const RngState* rng = static_cast<const RngState*>(rng_state);
  M3_CHECK_ARG(T >= 0 && D > 0 && Dt >= 0 && (Dt == 0 || task_feat));
  M3_CHECK_SHAPE(D % kGateDC == 0 && K >= 1 && K <= E && K <= 8);
  M3_CHECK_SHAPE(x_dtype == M3_F32 || x_dtype == M3_BF16);
  const int el = x_dtype == M3_F32 ? 4 : 2;
  M3_CHECK_ALIGN16(x); M3_CHECK_ALIGN16(w_gate); M3_CHECK_ALIGN16(clean_logits);
  if ((ldx * el) % 16 != 0) return M3_ERR_ALIGN;
  if (noise) M3_CHECK_ALIGN16(noise);
  if (noisy_logits) M3_CHECK_ALIGN16(noisy_logits);
  if (gates) M3_CHECK_ALIGN16(gates);
  if (T == 0) return M3_OK;
  const int K1 = K + 1 < E ? K + 1 : E;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (E) {
    M3_GATE_CASE(4)
    M3_GATE_CASE(8)
    M3_GATE_CASE(16)
    M3_GATE_CASE(32)
    M3_GATE_CASE(64)
    M3_GATE_CASE(128)
    default: return M3_ERR_SHAPE;
  }
}

extern "C" int m3_gate_fwd(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                           const float* noise, float noise_stddev, int T, int D, int Dt, int E, int K,
                           int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean_logits,
                           float* noisy_logits, float* gates, float* imp_partial, int32_t* load_partial,
                           m3_stream_t stream) {
  return gate_fwd_impl(x, x_dtype, ldx, task_feat, w_gate, noise, noise_stddev, T, D, Dt, E, K, idx, idx_full, score,
                       top_vals, clean_logits, noisy_logits, gates, imp_partial, load_partial, nullptr, nullptr,
                       nullptr, nullptr, stream);
}

// The same kernel drawing the router noise itself: noisy = clean + N(0,1) * noise_stddev with the normals generated in
// registers from rng_state = {uint64 seed, uint64 call counter} (device memory, philox.cuh) - no [T, E] noise tensor is
// ever written or read.  noisy_logits must be given (the backward pass and the load estimator read them).
extern "C" int m3_gate_fwd_rng(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                               const void* rng_state, float noise_stddev, int T, int D, int Dt, int E, int K,
                               int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean_logits,
                               float* noisy_logits, float* gates, float* imp_partial, int32_t* load_partial,
                               m3_stream_t stream) {
  M3_CHECK_ARG(rng_state && noisy_logits);
  return gate_fwd_impl(x, x_dtype, ldx, task_feat, w_gate, nullptr, noise_stddev, T, D, Dt, E, K, idx, idx_full, score,
                       top_vals, clean_logits, noisy_logits, gates, imp_partial, load_partial, nullptr, nullptr,
                       nullptr, rng_state, stream);
}

// Block-level fusion: x = RAW fp32 residual stream, w_gate_folded / ln_gb from m3_ln_fold_gate,
// ln_mean / ln_rstd from m3_ln_stats.
extern "C" int m3_gate_fwd_ln(const float* x, int64_t ldx, const float* ln_mean, const float* ln_rstd,
                              const float* ln_gb, const float* task_feat, const float* w_gate_folded,
                              const float* noise, float noise_stddev, int T, int D, int Dt, int E, int K,
                              int64_t* idx, int32_t* idx_full, float* score, float* top_vals, float* clean_logits,
                              float* noisy_logits, float* gates, float* imp_partial, int32_t* load_partial,
                              m3_stream_t stream) {
  M3_CHECK_ARG(ln_mean && ln_rstd && ln_gb);
  M3_CHECK_ALIGN16(ln_gb);
  return gate_fwd_impl(x, M3_F32, ldx, task_feat, w_gate_folded, noise, noise_stddev, T, D, Dt, E, K, idx, idx_full,
                       score, top_vals, clean_logits, noisy_logits, gates, imp_partial, load_partial, ln_mean,
                       ln_rstd, ln_gb, nullptr, stream);
}

extern "C" size_t m3_gate_bwd_workspace_bytes(int T, int D, int Dt, int E) {
  (void)Dt;
  const size_t n = (size_t)gate_bwd_chunks(T, E);
  return n * ((size_t)D * E + E) * sizeof(float);
}

static int gate_bwd_impl(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                         const float* logits, const int32_t* idx_full, int T, int D, int Dt, int E, int K,
                         const float* dscore, const float* dtop_vals, const float* dgates,
                         const float* dimportance, const float* dclean, const float* dnoisy,
                         const float* importance, const float* dcv_loss, float* dz, float* dw_gate,
                         float* dtask_feat, float* dx_gate, void* workspace, size_t workspace_bytes,
                         const float* ln_mean, const float* ln_rstd, const float* ln_gamma, const float* ln_beta,
                         m3_stream_t stream) {
  if ((importance == nullptr) != (dcv_loss == nullptr)) return M3_ERR_ARG;
  M3_CHECK_ARG(x && w_gate && logits && idx_full && dz && dw_gate && workspace);
  M3_CHECK_ARG(T >= 0 && D > 0 && Dt >= 0 && (Dt == 0 || task_feat));
  M3_CHECK_SHAPE(D % 4 == 0 && D / 4 * 2 <= 1024 && K >= 1 && K <= E);
  M3_CHECK_SHAPE(x_dtype == M3_F32 || x_dtype == M3_BF16);
  M3_CHECK_ALIGN16(x); M3_CHECK_ALIGN16(w_gate); M3_CHECK_ALIGN16(logits); M3_CHECK_ALIGN16(dz);
  if (dclean) M3_CHECK_ALIGN16(dclean);
  if (dnoisy) M3_CHECK_ALIGN16(dnoisy);
  if (workspace_bytes < m3_gate_bwd_workspace_bytes(T, D, Dt, E)) return M3_ERR_WORKSPACE;
  const int K1 = K + 1 < E ? K + 1 : E;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (T == 0) {      // an empty token subset (m3_gate_fwd accepts it too): the weight gradients are zero
    cudaError_t e = cudaMemsetAsync(dw_gate, 0, (size_t)(D + Dt) * E * sizeof(float), st);
    if (e == cudaSuccess && dtask_feat != nullptr && Dt > 0) e = cudaMemsetAsync(dtask_feat, 0, (size_t)Dt * sizeof(float), st);
    return e == cudaSuccess ? M3_OK : (int)e;
  }
  {
    const int64_t nthr = (int64_t)T * (E / 4);
    const int grid = (int)((nthr + 255) / 256);
#define M3_DZ_CASE(EE) \
  case EE: launch_k(gate_bwd_dz_kernel<EE>, grid, 256, 0, st, logits, idx_full, T, K, K1, dscore, dtop_vals, dgates, dimportance, dclean, dnoisy, importance, dcv_loss, dz); break;
    switch (E) {
      M3_DZ_CASE(4) M3_DZ_CASE(8) M3_DZ_CASE(16) M3_DZ_CASE(32) M3_DZ_CASE(64) M3_DZ_CASE(128)
      default: return M3_ERR_SHAPE;
    }
    M3_LAUNCH_CHECK();
  }
  const int nchunk = gate_bwd_chunks(T, E);
  const int tok_per_chunk = m3_ceil_div(T, nchunk);
  float* part = static_cast<float*>(workspace);
  float* cs_part = part + (size_t)nchunk * D * E;
  {
    // EB experts per CTA (<=16), EW per thread (<=8)
    const int EB = E < 16 ? E : 16;
    const int EW = EB < 8 ? EB : 8;
    dim3 grid(nchunk, E / EB);
    const int threads = (D / 4) * (EB / EW);
#define M3_DW_ARGS(XT) static_cast<const XT*>(x), ldx, dz, T, D, E, tok_per_chunk, part, cs_part, ln_mean, ln_rstd, ln_gamma, ln_beta
    // token batch (loads in flight per thread) limited by the register file at large CTAs
#define M3_DW_LAUNCH(EWV, XT, LNV)                                                                                   \
  do {                                                                                                              \
    if (threads <= 384) launch_k(gate_bwd_dw_kernel<EWV, XT, LNV, 16>, grid, threads, 0, st, M3_DW_ARGS(XT));            \
    else if (threads <= 512) launch_k(gate_bwd_dw_kernel<EWV, XT, LNV, 8>, grid, threads, 0, st, M3_DW_ARGS(XT));        \
    else launch_k(gate_bwd_dw_kernel<EWV, XT, LNV, 4>, grid, threads, 0, st, M3_DW_ARGS(XT));                            \
  } while (0)
    // cp.async ring variant (see gate_bwd_dw_ring_kernel): 16-byte copies need 16-byte rows; >= 2 CTAs per SM want
    // S stages within ~100 KB
    const int el = x_dtype == M3_F32 ? 4 : 2;
    const bool ring_ok = (D * el) % 16 == 0 && (ldx * el) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                         EB % 4 == 0 && threads <= 512 && (reinterpret_cast<uintptr_t>(dz) & 15u) == 0;
#define M3_DW_RING_S(EWV, XT, LNV, SV, DPTV)                                                                        \
  do {                                                                                                              \
    auto kern = gate_bwd_dw_ring_kernel<EWV, XT, LNV, SV, DPTV>;                                                    \
    if (smem > 48 * 1024) {                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
      if (e != cudaSuccess) return (int)e;                                                                          \
    }                                                                                                               \
    launch_k(kern, grid, threads * (4 / DPTV), smem, st, M3_DW_ARGS(XT));                                           \
  } while (0)
#define M3_DW_RING(EWV, XT, LNV)                                                                                    \
  do {                                                                                                              \
    const int sb = dw_stage_bytes<XT, LNV>(D);                                                                      \
    const int S = 4 * sb <= 100 * 1024 ? 4 : 3 * sb <= 100 * 1024 ? 3 : 2;                                          \
    const size_t smem = (size_t)S * sb;                                                                             \
    if (2 * threads <= 512) {       /* two columns per thread: twice the warps */                                   \
      if (S == 4) M3_DW_RING_S(EWV, XT, LNV, 4, 2); else if (S == 3) M3_DW_RING_S(EWV, XT, LNV, 3, 2);              \
      else M3_DW_RING_S(EWV, XT, LNV, 2, 2);                                                                        \
    } else {                                                                                                        \
      if (S == 4) M3_DW_RING_S(EWV, XT, LNV, 4, 4); else if (S == 3) M3_DW_RING_S(EWV, XT, LNV, 3, 4);              \
      else M3_DW_RING_S(EWV, XT, LNV, 2, 4);                                                                        \
    }                                                                                                               \
  } while (0)
    if (ring_ok && 2 * dw_stage_bytes<float, true>(D) <= 200 * 1024) {
      if (ln_mean != nullptr) {
        if (EW == 8) M3_DW_RING(8, float, true); else M3_DW_RING(4, float, true);
      } else if (x_dtype == M3_F32) {
        if (EW == 8) M3_DW_RING(8, float, false); else M3_DW_RING(4, float, false);
      } else {
        if (EW == 8) M3_DW_RING(8, __nv_bfloat16, false); else M3_DW_RING(4, __nv_bfloat16, false);
      }
    } else if (ln_mean != nullptr) {
      if (EW == 8) M3_DW_LAUNCH(8, float, true); else M3_DW_LAUNCH(4, float, true);
    } else if (x_dtype == M3_F32) {
      if (EW == 8) M3_DW_LAUNCH(8, float, false); else M3_DW_LAUNCH(4, float, false);
    } else {
      if (EW == 8) M3_DW_LAUNCH(8, __nv_bfloat16, false); else M3_DW_LAUNCH(4, __nv_bfloat16, false);
    }
    M3_LAUNCH_CHECK();
  }
  {
    const int main_blocks = (int)(((int64_t)D * E + 31) / 32);
    launch_k(gate_bwd_reduce_kernel, main_blocks + (Dt > 0 ? 1 : 0), 1024, 0, st, part, cs_part, nchunk, D, Dt, E,
             task_feat, w_gate, dw_gate, dtask_feat);
    M3_LAUNCH_CHECK();
  }
  if (dx_gate != nullptr) {
    const int64_t n = (int64_t)T * (D / 4);
    launch_k(gate_bwd_dx_kernel, (int)((n + 255) / 256), 256, 0, st, dz, w_gate, T, D, E, dx_gate);
    M3_LAUNCH_CHECK();
  }
  return M3_OK;
}

extern "C" int m3_gate_bwd(const void* x, int x_dtype, int64_t ldx, const float* task_feat, const float* w_gate,
                           const float* logits, const int32_t* idx_full, int T, int D, int Dt, int E, int K,
                           const float* dscore, const float* dtop_vals, const float* dgates,
                           const float* dimportance, const float* dclean, const float* dnoisy,
                           const float* importance, const float* dcv_loss, float* dz, float* dw_gate,
                           float* dtask_feat, float* dx_gate, void* workspace, size_t workspace_bytes,
                           m3_stream_t stream) {
  return gate_bwd_impl(x, x_dtype, ldx, task_feat, w_gate, logits, idx_full, T, D, Dt, E, K, dscore, dtop_vals, dgates,
                       dimportance, dclean, dnoisy, importance, dcv_loss, dz, dw_gate, dtask_feat, dx_gate, workspace,
                       workspace_bytes, nullptr, nullptr, nullptr, nullptr, stream);
}

// Block-level fusion: x = RAW fp32 residual stream, normalised on load (dw_gate is w.r.t. the
// ORIGINAL w_gate, which is also what must be passed here).
extern "C" int m3_gate_bwd_ln(const float* x, int64_t ldx, const float* ln_mean, const float* ln_rstd,
                              const float* ln_gamma, const float* ln_beta, const float* task_feat,
                              const float* w_gate, const float* logits, const int32_t* idx_full, int T, int D, int Dt,
                              int E, int K, const float* dscore, const float* dtop_vals, const float* dgates,
                              const float* dimportance, const float* dclean, const float* dnoisy,
                              const float* importance, const float* dcv_loss, float* dz, float* dw_gate,
                              float* dtask_feat, void* workspace, size_t workspace_bytes, m3_stream_t stream) {
  M3_CHECK_ARG(ln_mean && ln_rstd && ln_gamma && ln_beta);
  return gate_bwd_impl(x, M3_F32, ldx, task_feat, w_gate, logits, idx_full, T, D, Dt, E, K, dscore, dtop_vals, dgates,
                       dimportance, dclean, dnoisy, importance, dcv_loss, dz, dw_gate, dtask_feat, nullptr, workspace,
                       workspace_bytes, ln_mean, ln_rstd, ln_gamma, ln_beta, stream);
}

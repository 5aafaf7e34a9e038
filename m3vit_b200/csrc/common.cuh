// Shared device/host helpers for the M3ViT MoE hot-path kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/m3vit_moe.h"

#define M3_CHECK_ARG(cond)      \
  do {                          \
    if (!(cond)) return M3_ERR_ARG; \
  } while (0)
#define M3_CHECK_SHAPE(cond)      \
  do {                            \
    if (!(cond)) return M3_ERR_SHAPE; \
  } while (0)
#define M3_CHECK_ALIGN16(p)                                   \
  do {                                                        \
    if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return M3_ERR_ALIGN; \
  } while (0)
#define M3_LAUNCH_CHECK()                      \
  do {                                         \
    cudaError_t e__ = cudaGetLastError();      \
    if (e__ != cudaSuccess) return (int)e__;   \
  } while (0)

static inline int m3_ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int m3_round_up(int a, int b) { return (a + b - 1) / b * b; }

namespace m3 {

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) ------------------------------------
// Every kernel of the path calls pdl_wait() before its first global-memory access (reads of a predecessor's
// output AND writes a predecessor might still read) and pdl_trigger() right away, so that the next kernel's
// CTAs are scheduled as this kernel's last CTAs drain and run their prologue (smem carve-up, mbarrier init,
// TMEM allocation, tensor-map prefetch) under this kernel's tail.  Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_knobs[M3_KNOB_COUNT_];   // abi.cu (m3_set_knob)

// kern<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialisation attribute (M3_KNOB_PDL)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_knobs[M3_KNOB_PDL] ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 16-byte vector access -------------------------------------------------
// Streaming (read-once) global load / store hints: the row movers touch every
// byte exactly once, so keep them out of L1.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// 256-bit variants (sm_100: LDG.256 / STG.256): one full 32-byte sector per lane
struct U8 { uint32_t v[8]; };
__device__ __forceinline__ U8 ldg_stream256(const void* p) {
  U8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream256(void* p, const U8& r) {
  asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]),
               "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t float2_to_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// An 8-element slice of a row, held as fp32 in registers, loadable/storable as
// fp32 (2 x 16 B) or bf16 (1 x 16 B).
struct Vec8 {
  float v[8];
};

template <typename T>
__device__ __forceinline__ Vec8 load8(const T* p);

template <>
__device__ __forceinline__ Vec8 load8<float>(const float* p) {
  Vec8 r;
  uint4 a = ldg_stream(p), b = ldg_stream(p + 4);
  r.v[0] = __uint_as_float(a.x); r.v[1] = __uint_as_float(a.y);
  r.v[2] = __uint_as_float(a.z); r.v[3] = __uint_as_float(a.w);
  r.v[4] = __uint_as_float(b.x); r.v[5] = __uint_as_float(b.y);
  r.v[6] = __uint_as_float(b.z); r.v[7] = __uint_as_float(b.w);
  return r;
}
template <>
__device__ __forceinline__ Vec8 load8<__nv_bfloat16>(const __nv_bfloat16* p) {
  Vec8 r;
  uint4 a = ldg_stream(p);
  float2 f;
  f = bf16x2_to_float2(a.x); r.v[0] = f.x; r.v[1] = f.y;
  f = bf16x2_to_float2(a.y); r.v[2] = f.x; r.v[3] = f.y;
  f = bf16x2_to_float2(a.z); r.v[4] = f.x; r.v[5] = f.y;
  f = bf16x2_to_float2(a.w); r.v[6] = f.x; r.v[7] = f.y;
  return r;
}

// The same slice as it sits in memory (bf16: 4 registers instead of 8).  Row gathers issue ALL their
// loads as Raw8 first and convert afterwards: twice the rows in flight for the same register budget.
template <typename T>
struct Raw8;
template <>
struct Raw8<float> {
  uint4 a, b;
};
template <>
struct Raw8<__nv_bfloat16> {
  uint4 a;
};
__device__ __forceinline__ void load_raw8(Raw8<float>& r, const float* p) {
  r.a = ldg_stream(p);
  r.b = ldg_stream(p + 4);
}
__device__ __forceinline__ void load_raw8(Raw8<__nv_bfloat16>& r, const __nv_bfloat16* p) { r.a = ldg_stream(p); }
__device__ __forceinline__ void store_raw8(float* p, const Raw8<float>& r) { stg_stream(p, r.a); stg_stream(p + 4, r.b); }
__device__ __forceinline__ void store_raw8(__nv_bfloat16* p, const Raw8<__nv_bfloat16>& r) { stg_stream(p, r.a); }
__device__ __forceinline__ void zero_raw8(Raw8<float>& r) { r.a = make_uint4(0, 0, 0, 0); r.b = r.a; }
__device__ __forceinline__ void zero_raw8(Raw8<__nv_bfloat16>& r) { r.a = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ Vec8 cvt8(const Raw8<float>& q) {
  Vec8 r;
  r.v[0] = __uint_as_float(q.a.x); r.v[1] = __uint_as_float(q.a.y);
  r.v[2] = __uint_as_float(q.a.z); r.v[3] = __uint_as_float(q.a.w);
  r.v[4] = __uint_as_float(q.b.x); r.v[5] = __uint_as_float(q.b.y);
  r.v[6] = __uint_as_float(q.b.z); r.v[7] = __uint_as_float(q.b.w);
  return r;
}
__device__ __forceinline__ Vec8 cvt8(const Raw8<__nv_bfloat16>& q) {
  Vec8 r;
  float2 f;
  f = bf16x2_to_float2(q.a.x); r.v[0] = f.x; r.v[1] = f.y;
  f = bf16x2_to_float2(q.a.y); r.v[2] = f.x; r.v[3] = f.y;
  f = bf16x2_to_float2(q.a.z); r.v[4] = f.x; r.v[5] = f.y;
  f = bf16x2_to_float2(q.a.w); r.v[6] = f.x; r.v[7] = f.y;
  return r;
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const Vec8& r);

template <>
__device__ __forceinline__ void store8<float>(float* p, const Vec8& r) {
  uint4 a, b;
  a.x = __float_as_uint(r.v[0]); a.y = __float_as_uint(r.v[1]);
  a.z = __float_as_uint(r.v[2]); a.w = __float_as_uint(r.v[3]);
  b.x = __float_as_uint(r.v[4]); b.y = __float_as_uint(r.v[5]);
  b.z = __float_as_uint(r.v[6]); b.w = __float_as_uint(r.v[7]);
  stg_stream(p, a);
  stg_stream(p + 4, b);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const Vec8& r) {
  uint4 a;
  a.x = float2_to_bf16x2(r.v[0], r.v[1]);
  a.y = float2_to_bf16x2(r.v[2], r.v[3]);
  a.z = float2_to_bf16x2(r.v[4], r.v[5]);
  a.w = float2_to_bf16x2(r.v[6], r.v[7]);
  stg_stream(p, a);
}

// exact-erf GELU and its derivative (nn.GELU() default, as the reference's experts use)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast erf for the tensor-core epilogues: odd rational minimax x*P(x^2)/Q(x^2) on [-4,4]
// (|err| <= 4.3e-7 vs erf, verified against torch.erf in fp64) - ~2x cheaper than erff and
// far below bf16 resolution.  GELU here is still the exact-erf GELU of the reference, not the
// tanh approximation.
__device__ __forceinline__ float erf_fast(float x) {
  x = fminf(fmaxf(x, -4.0f), 4.0f);
  const float x2 = x * x;
  float p = -2.72614225801306e-10f;
  p = fmaf(p, x2, 2.77068142495902e-08f);
  p = fmaf(p, x2, -2.10102402082508e-06f);
  p = fmaf(p, x2, -5.69250639462346e-05f);
  p = fmaf(p, x2, -7.34990630326855e-04f);
  p = fmaf(p, x2, -2.95459980854025e-03f);
  p = fmaf(p, x2, -1.60960333262415e-02f);
  p *= x;
  float q = -1.45660718464996e-05f;
  q = fmaf(q, x2, -2.13374055278905e-04f);
  q = fmaf(q, x2, -1.68282697438203e-03f);
  q = fmaf(q, x2, -7.37332916720468e-03f);
  q = fmaf(q, x2, -1.42647390514189e-02f);
  return __fdividef(p, q);
}
__device__ __forceinline__ float gelu_fast(float x) {
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));
}
// returns gelu'(x), and gelu(x) through *g
__device__ __forceinline__ float gelu_fast_grad(float x, float* g) {
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  *g = x * cdf;
  return fmaf(x, pdf, cdf);
}

// ---- packed fp32x2 math (sm_100 FFMA2): halves the instruction count of the GEMM epilogues,
// which are issue-bound at K = 384 (0.09 tensor-clk per output element).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define M3_K2(c) pk2((c), (c))
// erf of two values (same rational approximation as erf_fast)
__device__ __forceinline__ f32x2 erf_fast2(float a, float b) {
  a = fminf(fmaxf(a, -4.0f), 4.0f);
  b = fminf(fmaxf(b, -4.0f), 4.0f);
  const f32x2 x = pk2(a, b);
  const f32x2 x2 = mul2(x, x);
  f32x2 p = M3_K2(-2.72614225801306e-10f);
  p = fma2(p, x2, M3_K2(2.77068142495902e-08f));
  p = fma2(p, x2, M3_K2(-2.10102402082508e-06f));
  p = fma2(p, x2, M3_K2(-5.69250639462346e-05f));
  p = fma2(p, x2, M3_K2(-7.34990630326855e-04f));
  p = fma2(p, x2, M3_K2(-2.95459980854025e-03f));
  p = fma2(p, x2, M3_K2(-1.60960333262415e-02f));
  p = mul2(p, x);
  f32x2 q = M3_K2(-1.45660718464996e-05f);
  q = fma2(q, x2, M3_K2(-2.13374055278905e-04f));
  q = fma2(q, x2, M3_K2(-1.68282697438203e-03f));
  q = fma2(q, x2, M3_K2(-7.37332916720468e-03f));
  q = fma2(q, x2, M3_K2(-1.42647390514189e-02f));
  float q0, q1;
  unpk2(q, q0, q1);
  return mul2(p, pk2(rcp_approx(q0), rcp_approx(q1)));
}
// v <- gelu(v) for a pair
__device__ __forceinline__ f32x2 gelu_fast2(f32x2 v) {
  float a, b;
  unpk2(mul2(v, M3_K2(0.70710678118654752440f)), a, b);
  const f32x2 e = erf_fast2(a, b);
  return mul2(mul2(v, M3_K2(0.5f)), add2(e, M3_K2(1.0f)));
}
// returns gelu'(x) for a pair; *g = gelu(x).
// Both need Phi(x) = (1 + erf(x/sqrt2))/2 AND pdf(x) = exp(-x^2/2)/sqrt(2 pi), so erf comes from Abramowitz-Stegun 7.1.26
//   erf(u) = 1 - (a1 t + .. + a5 t^5) exp(-u^2),  t = 1/(1 + p u),  u = |x|/sqrt2  (|err| <= 1.5e-7)
// whose exponential IS the pdf's: one ex2 + one rcp per element, no clamps (exp underflows to the saturated values).
// Constant folding: the ex2 argument carries log2(1/sqrt(2 pi)) so that ex2 returns the pdf itself, and the
// coefficients carry -sqrt(2 pi)/2, so that  |Phi(x) - 1/2| = 1/2 + poly(t) * t * pdf  is one FMA.
// |gelu err| <= 5e-7, |gelu' err| <= 4e-7 against fp64.
#define M3_GELU_A5 (-1.3302744296f)
#define M3_GELU_A4 (1.8212559791f)
#define M3_GELU_A3 (-1.7814779366f)
#define M3_GELU_A2 (0.3565637812f)
#define M3_GELU_A1 (-0.3193815303f)
#define M3_GELU_TK (0.3275911f * 0.70710678118654752440f)
#define M3_GELU_EC (-0.7213475204444817f)      /* -0.5 * log2(e) */
#define M3_GELU_EL (-1.3257480647361592f)      /* log2(1 / sqrt(2 pi)) */
__device__ __forceinline__ f32x2 gelu_fast_grad2(f32x2 x, f32x2* g) {
  float x0, x1;
  unpk2(x, x0, x1);
  float d0, d1;
  unpk2(fma2(pk2(fabsf(x0), fabsf(x1)), M3_K2(M3_GELU_TK), M3_K2(1.0f)), d0, d1);
  const f32x2 t = pk2(rcp_approx(d0), rcp_approx(d1));
  float a0, a1;
  unpk2(fma2(mul2(x, x), M3_K2(M3_GELU_EC), M3_K2(M3_GELU_EL)), a0, a1);
  const f32x2 pdf = pk2(ex2_approx(a0), ex2_approx(a1));
  f32x2 p = fma2(M3_K2(M3_GELU_A5), t, M3_K2(M3_GELU_A4));
  p = fma2(p, t, M3_K2(M3_GELU_A3));
  p = fma2(p, t, M3_K2(M3_GELU_A2));
  p = fma2(p, t, M3_K2(M3_GELU_A1));
  float q0, q1;
  unpk2(fma2(mul2(p, t), pdf, M3_K2(0.5f)), q0, q1);                    // |Phi(x) - 1/2|
  const f32x2 cdf = add2(pk2(copysignf(q0, x0), copysignf(q1, x1)), M3_K2(0.5f));
  *g = mul2(x, cdf);
  return fma2(x, pdf, cdf);
}

// volatile variants: ptxas keeps volatile asm statements in program order, which is how the stage-by-stage
// order below survives into SASS (left alone, its scheduler re-serialises each pair's dependent chain).
__device__ __forceinline__ f32x2 fma2v(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2v(f32x2 a, f32x2 b) {
  f32x2 r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx_v(float x) {
  float r;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx_v(float x) {
  float r;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int NP>
__device__ __forceinline__ void gelu_fast_grad2_batch(const f32x2* x, f32x2* g, f32x2* gr) {
  f32x2 t[NP], e[NP], p[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float x0, x1, d0, d1;
    unpk2(x[i], x0, x1);
    unpk2(fma2(pk2(fabsf(x0), fabsf(x1)), M3_K2(M3_GELU_TK), M3_K2(1.0f)), d0, d1);
    t[i] = pk2(rcp_approx_v(d0), rcp_approx_v(d1));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float a0, a1;
    unpk2(fma2(mul2(x[i], x[i]), M3_K2(M3_GELU_EC), M3_K2(M3_GELU_EL)), a0, a1);
    e[i] = pk2(ex2_approx_v(a0), ex2_approx_v(a1));                       // the pdf itself
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = fma2v(M3_K2(M3_GELU_A5), t[i], M3_K2(M3_GELU_A4));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = fma2v(p[i], t[i], M3_K2(M3_GELU_A3));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = fma2v(p[i], t[i], M3_K2(M3_GELU_A2));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = fma2v(p[i], t[i], M3_K2(M3_GELU_A1));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = fma2(mul2v(p[i], t[i]), e[i], M3_K2(0.5f));    // |Phi(x) - 1/2|
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float x0, x1, q0, q1;
    unpk2(x[i], x0, x1);
    unpk2(p[i], q0, q1);
    const f32x2 cdf = add2(pk2(copysignf(q0, x0), copysignf(q1, x1)), M3_K2(0.5f));
    g[i] = mul2(x[i], cdf);
    gr[i] = fma2(x[i], e[i], cdf);
  }
}

}  // namespace m3

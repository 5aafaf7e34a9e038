// Dispatch / combine row movers (forward + backward) for sm_100a.
//
// Replace fmoe's MOEScatter / MOEGather (torch.index_select / index_copy_ /
// index_add_ on [R, D] buffers) and the torch.bmm gate-weighted combine of the
// reference (/root/reference/models/moe/origin/custom_moe_layer.py:255-257,283-297).
//
// All four kernels are HBM-bound.  A token is owned by 16 lanes (half a warp);
// each lane moves 8-element slices as single 128-bit transactions, so a queue
// row (D*2 B in bf16) is written/read as whole 128-byte lines.  Loads are issued
// before any store (memory-level parallelism), with L1::no_allocate streaming
// hints: every byte is touched exactly once.
//
//   algorithmic bytes per token (bf16 queues, el=2, x fp32 = 4):
//     dispatch_fwd  D*el_x + K*D*el + K*4            combine_fwd  K*D*el + K*8 + D*el_out
//     combine_bwd   D*el_g + 2*K*D*el + K*12         dispatch_bwd K*D*el + K*4 + D*el_dx
#include "common.cuh"

namespace m3 {

constexpr int kLanesPerTok = 16;
constexpr int kPermThreads = 256;
constexpr int kTokPerCta = kPermThreads / kLanesPerTok;
constexpr int kMaxVec = 8;  // per-lane 8-element slices: D <= 16*8*8 = 1024
constexpr int kGatherK = 4; // queue rows gathered per batch (loads in flight per lane = kGatherK * NV)

// Queue addressing.  Single GPU: every queue row lives in the local buffer.
// Expert parallel (EP): slot s belongs to rank slot_rank[s]; its row lives in that
// rank's queue, reached through a peer-mapped base pointer (NVLink load/store).
template <typename T>
struct Queue {
  T* local;
  T* const* bases;           // [W] device array of peer-mapped queue bases (EP only)
  const int32_t* slot_rank;  // [T*K] owner rank of every slot (EP only)
  template <bool EP>
  __device__ __forceinline__ T* row(int64_t slot, int r, int D) const {
    T* b = EP ? bases[__ldg(slot_rank + slot)] : local;
    return b + (int64_t)r * D;
  }
};

template <typename TO>
__device__ __forceinline__ void zero_pad_rows(TO* q, const int32_t* counts, const int32_t* offsets, int e, int D,
                                              int32_t* meta = nullptr) {
  const int r0 = offsets[e] + counts[e], r1 = offsets[e + 1];
  if (meta != nullptr)      // expert parallel, return store: a padding row has no home (m3_ep_ffn_fwd)
    for (int i = threadIdx.x; i < r1 - r0; i += kPermThreads) meta[r0 + i] = -1;
  const int nvec = D / 8;
  Vec8 z;
#pragma unroll
  for (int i = 0; i < 8; ++i) z.v[i] = 0.f;
  for (int64_t i = threadIdx.x; i < (int64_t)(r1 - r0) * nvec; i += kPermThreads)
    store8<TO>(q + ((int64_t)r0 + i / nvec) * D + (i % nvec) * 8, z);
}

template <typename TO>
__global__ void __launch_bounds__(kPermThreads)
zero_pad_rows_kernel(TO* __restrict__ q, const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets, int D,
                     int32_t* __restrict__ meta) {
  pdl_wait();
  pdl_trigger();
  zero_pad_rows<TO>(q, counts, offsets, blockIdx.x, D, meta);
}

// xq[pos[t,k]] = cast(x[t]); trailing CTAs zero the padding rows of every queue.
template <typename TI, typename TO, int NV, bool EP>
__global__ void __launch_bounds__(kPermThreads)
dispatch_fwd_kernel(const TI* __restrict__ x, const int32_t* __restrict__ pos, const int32_t* __restrict__ counts,
                    const int32_t* __restrict__ offsets, int T, int K, int D, int tok_ctas, Queue<TO> xq) {
  pdl_wait();
  pdl_trigger();
  if ((int)blockIdx.x >= tok_ctas) {
    // zero rows [off[e]+cnt[e], off[e+1]) of expert e (local queue only)
    zero_pad_rows<TO>(xq.local, counts, offsets, blockIdx.x - tok_ctas, D);
    return;
  }
  const int sub = threadIdx.x % kLanesPerTok;
  const int t = blockIdx.x * kTokPerCta + threadIdx.x / kLanesPerTok;
  if (t >= T) return;
  const int nvec = D / 8;
  Vec8 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kLanesPerTok;
    if (c < nvec) v[i] = load8<TI>(x + (int64_t)t * D + c * 8);
  }
  for (int k = 0; k < K; ++k) {
    const int row = __ldg(pos + (int64_t)t * K + k);
    if (row < 0) continue;  // dropped slot
    TO* dst = xq.template row<EP>((int64_t)t * K + k, row, D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = sub + i * kLanesPerTok;
      if (c < nvec) store8<TO>(dst + c * 8, v[i]);
    }
  }
}

// out[t] = sum_k score[t,k] * yq[pos[t,k]]   (fp32 accumulation, k ascending like bmm)
// Rows in flight per thread vs resident CTAs (tools/variants.py, T = 38 432, D = 384): 4 rows at 2 CTAs/SM (128 regs) 44 us,
// 2 rows at 4 CTAs/SM (<= 64 regs) 39 us - the gather wants warps, not deeper per-thread batches.
template <typename TI, typename TO, int NV, bool EP,
          int kGatherK = (NV <= 3 && sizeof(TI) == 2 && !EP ? 2 : m3::kGatherK),
          int MINB = (NV <= 3 && sizeof(TI) == 2 && !EP ? 4 : 1)>
__global__ void __launch_bounds__(kPermThreads, MINB)
combine_fwd_kernel(Queue<const TI> yq, const int32_t* __restrict__ pos, const float* __restrict__ score,
                   int T, int K, int D, TO* __restrict__ out, TI* __restrict__ ysave) {
  pdl_wait();
  pdl_trigger();
  const int sub = threadIdx.x % kLanesPerTok;
  const int t = blockIdx.x * kTokPerCta + threadIdx.x / kLanesPerTok;
  if (t >= T) return;
  const int nvec = D / 8;
  Vec8 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i].v[j] = 0.f;
  // rows are gathered kGatherK at a time, every load issued (raw, unconverted) before the first FMA
  for (int k0 = 0; k0 < K; k0 += kGatherK) {
    int row[kGatherK];
    float s[kGatherK];
    Raw8<TI> raw[kGatherK][NV];
#pragma unroll
    for (int u = 0; u < kGatherK; ++u) {
      const int k = k0 + u;
      row[u] = k < K ? __ldg(pos + (int64_t)t * K + k) : -1;
      s[u] = row[u] >= 0 ? __ldg(score + (int64_t)t * K + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kGatherK; ++u) {
#pragma unroll
      for (int i = 0; i < NV; ++i) zero_raw8(raw[u][i]);
      if (row[u] >= 0) {
        const TI* src = yq.template row<EP>((int64_t)t * K + k0 + u, row[u], D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = sub + i * kLanesPerTok;
          if (c < nvec) load_raw8(raw[u][i], src + c * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherK; ++u)    // k ascending, like bmm
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const Vec8 v = cvt8(raw[u][i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i].v[j] = fmaf(s[u], v.v[j], acc[i].v[j]);
      }
    // expert parallel: keep a LOCAL copy of the rows just pulled over NVLink (slot order), so that the
    // backward pass (dscore = <g, y>) does not have to pull them a second time
    if (EP && ysave != nullptr) {
#pragma unroll
      for (int u = 0; u < kGatherK; ++u) {
        if (row[u] < 0) continue;
        TI* dst = ysave + ((int64_t)t * K + k0 + u) * D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = sub + i * kLanesPerTok;
          if (c < nvec) store_raw8(dst + c * 8, raw[u][i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kLanesPerTok;
    if (c < nvec) store8<TO>(out + (int64_t)t * D + c * 8, acc[i]);
  }
}

// dscore[t,k] = <g[t], yq[pos[t,k]]>;  dyq[pos[t,k]] = score[t,k] * g[t];  zero dyq padding rows
template <typename TG, typename TQ, int NV, bool EP, int kGatherK = m3::kGatherK, int MINB = 1>
__global__ void __launch_bounds__(kPermThreads, MINB)
combine_bwd_kernel(const TG* __restrict__ g, Queue<const TQ> yq, const int32_t* __restrict__ pos,
                   const float* __restrict__ score, const int32_t* __restrict__ counts,
                   const int32_t* __restrict__ offsets, int T, int K, int D, int tok_ctas,
                   Queue<TQ> dyq, float* __restrict__ dscore, const TQ* __restrict__ ysave) {
  pdl_wait();
  pdl_trigger();
  if ((int)blockIdx.x >= tok_ctas) {
    zero_pad_rows<TQ>(dyq.local, counts, offsets, blockIdx.x - tok_ctas, D);
    return;
  }
  const int sub = threadIdx.x % kLanesPerTok;
  const int tt = blockIdx.x * kTokPerCta + threadIdx.x / kLanesPerTok;
  const bool valid = tt < T;
  const int t = valid ? tt : T - 1;  // keep all lanes alive for the shuffles
  const int nvec = D / 8;
  Vec8 gv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kLanesPerTok;
    if (c < nvec) gv[i] = load8<TG>(g + (int64_t)t * D + c * 8);
  }
  // the y rows of kGatherK slots are loaded (raw) together; the dyq stores do not depend on them and go first
  for (int k0 = 0; k0 < K; k0 += kGatherK) {
    int row[kGatherK];
    float s[kGatherK];
    Raw8<TQ> raw[kGatherK][NV];
#pragma unroll
    for (int u = 0; u < kGatherK; ++u) {
      const int k = k0 + u;
      row[u] = k < K ? __ldg(pos + (int64_t)t * K + k) : -1;
      s[u] = row[u] >= 0 ? __ldg(score + (int64_t)t * K + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kGatherK; ++u) {
#pragma unroll
      for (int i = 0; i < NV; ++i) zero_raw8(raw[u][i]);
      if (row[u] >= 0) {
        const int64_t slot = (int64_t)t * K + k0 + u;
        const TQ* ysrc = (EP && ysave != nullptr) ? ysave + slot * D        // local copy kept by combine_fwd
                                                  : yq.template row<EP>(slot, row[u], D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = sub + i * kLanesPerTok;
          if (c < nvec) load_raw8(raw[u][i], ysrc + c * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherK; ++u) {
      if (row[u] >= 0 && valid) {
        TQ* ddst = dyq.template row<EP>((int64_t)t * K + k0 + u, row[u], D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = sub + i * kLanesPerTok;
          if (c < nvec) {
            Vec8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = s[u] * gv[i].v[j];
            store8<TQ>(ddst + c * 8, o);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherK; ++u) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (sub + i * kLanesPerTok < nvec) {
          const Vec8 yv = cvt8(raw[u][i]);    // zero where nothing was loaded
#pragma unroll
          for (int j = 0; j < 8; ++j) dot = fmaf(gv[i].v[j], yv.v[j], dot);
        }
      }
#pragma unroll
      for (int o = kLanesPerTok / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (valid && sub == 0 && k0 + u < K) dscore[(int64_t)t * K + k0 + u] = dot;
    }
  }
}

// dx[t] = sum_k dxq[pos[t,k]]
template <typename TI, typename TO, int NV, bool EP>
__global__ void __launch_bounds__(kPermThreads)
dispatch_bwd_kernel(Queue<const TI> dxq, const int32_t* __restrict__ pos, int T, int K, int D, TO* __restrict__ dx) {
  pdl_wait();
  pdl_trigger();
  const int sub = threadIdx.x % kLanesPerTok;
  const int t = blockIdx.x * kTokPerCta + threadIdx.x / kLanesPerTok;
  if (t >= T) return;
  const int nvec = D / 8;
  Vec8 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i].v[j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const int row = __ldg(pos + (int64_t)t * K + k);
    if (row < 0) continue;
    const TI* src = dxq.template row<EP>((int64_t)t * K + k, row, D);
    Vec8 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = sub + i * kLanesPerTok;
      if (c < nvec) v[i] = load8<TI>(src + c * 8);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i].v[j] += v[i].v[j];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = sub + i * kLanesPerTok;
    if (c < nvec) store8<TO>(dx + (int64_t)t * D + c * 8, acc[i]);
  }
}

// Same, plus the router's dx:  dx[t] += dz[t] @ w_gate[:D]^T  (gate input == layer input).
// w_gate[:D] is staged ONCE per CTA in shared memory, transposed to [E][D] so that the 16 lanes of
// a token read consecutive 32-byte slices (conflict-free).  Every 16-lane group owns TWO tokens, so
// each weight slice read from smem feeds two tokens (the kernel was smem-bound with one).
template <typename TI, typename TO, int NV, bool EP>
__global__ void __launch_bounds__(kPermThreads, 2)
dispatch_bwd_gate_kernel(Queue<const TI> dxq, const int32_t* __restrict__ pos, int T, int K, int D,
                         const float* __restrict__ dz, const float* __restrict__ w_gate, int E,
                         TO* __restrict__ dx) {
  pdl_wait();
  pdl_trigger();
  // wt[e][h][c][4]: the 8 columns c*8..c*8+7 of a lane's slice are kept as two float4 in separate
  // planes h = 0/1, so that the 16 lanes of a group read CONSECUTIVE 16-byte words (a plain [E][D]
  // layout makes lanes 32 B apart: 2-way bank conflicts on every read).
  extern __shared__ __align__(16) float wt[];
  const int half = D / 2;
  for (int i = threadIdx.x; i < D * E; i += kPermThreads) {
    const int e = i / D, d = i % D;             // strided (L2-resident) reads, once per CTA
    const int c = d >> 3, h = (d >> 2) & 1, j = d & 3;
    wt[e * D + h * half + c * 4 + j] = __ldg(w_gate + (int64_t)d * E + e);
  }
  __syncthreads();
  const int sub = threadIdx.x % kLanesPerTok;
  const int grp = threadIdx.x / kLanesPerTok;
  const int nvec = D / 8;
  for (int tb = blockIdx.x * (2 * kTokPerCta); tb < T; tb += gridDim.x * (2 * kTokPerCta)) {
    const int t0 = tb + 2 * grp;
    if (t0 >= T) continue;
    const bool has1 = t0 + 1 < T;
    const int t1 = has1 ? t0 + 1 : t0;
    Vec8 a0[NV], a1[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0[i].v[j] = 0.f; a1[i].v[j] = 0.f; }
    // Software pipeline inside the thread: each phase ISSUES the (raw) loads of KB queue rows per token,
    // then does its share of the router-term FMAs (which only touch dz and shared memory) while those
    // loads are in flight, and only then adds the rows.
    constexpr int KB = sizeof(TI) == 2 ? 2 : 1;
    const int P = (K + KB - 1) / KB;
    for (int ph = 0; ph < P; ++ph) {
      Raw8<TI> q0[KB][NV], q1[KB][NV];
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        const int k = ph * KB + u;
        const int r0 = k < K ? __ldg(pos + (int64_t)t0 * K + k) : -1;
        const int r1 = k < K ? __ldg(pos + (int64_t)t1 * K + k) : -1;
#pragma unroll
        for (int i = 0; i < NV; ++i) { zero_raw8(q0[u][i]); zero_raw8(q1[u][i]); }
        if (r0 >= 0) {
          const TI* src = dxq.template row<EP>((int64_t)t0 * K + k, r0, D);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = sub + i * kLanesPerTok;
            if (c < nvec) load_raw8(q0[u][i], src + c * 8);
          }
        }
        if (r1 >= 0) {
          const TI* src = dxq.template row<EP>((int64_t)t1 * K + k, r1, D);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = sub + i * kLanesPerTok;
            if (c < nvec) load_raw8(q1[u][i], src + c * 8);
          }
        }
      }
      const int e_lo = ((E * ph) / P) & ~3;
      const int e_hi = ph == P - 1 ? E : (((E * (ph + 1)) / P) & ~3);
      for (int e = e_lo; e < e_hi; e += 4) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(dz + (int64_t)t0 * E + e));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(dz + (int64_t)t1 * E + e));
        const float z0[4] = {g0.x, g0.y, g0.z, g0.w};
        const float z1[4] = {g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int ee = 0; ee < 4; ++ee) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = sub + i * kLanesPerTok;
            if (c < nvec) {
              const float4 wa = *reinterpret_cast<const float4*>(wt + (e + ee) * D + c * 4);
              const float4 wb = *reinterpret_cast<const float4*>(wt + (e + ee) * D + half + c * 4);
              const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                a0[i].v[j] = fmaf(z0[ee], w8[j], a0[i].v[j]);
                a1[i].v[j] = fmaf(z1[ee], w8[j], a1[i].v[j]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < KB; ++u)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const Vec8 v0 = cvt8(q0[u][i]), v1 = cvt8(q1[u][i]);
#pragma unroll
          for (int j = 0; j < 8; ++j) { a0[i].v[j] += v0.v[j]; a1[i].v[j] += v1.v[j]; }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = sub + i * kLanesPerTok;
      if (c < nvec) {
        store8<TO>(dx + (int64_t)t0 * D + c * 8, a0[i]);
        if (has1) store8<TO>(dx + (int64_t)t1 * D + c * 8, a1[i]);
      }
    }
  }
}

// bf16 queues: the same kernel with the router term on the tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32).
// The SIMT version above spends 2 B of shared-memory reads per FMA on w_gate and is bound by them (59 us for a
// 27 us gather at T = 38 432); here a warp owns 16 tokens, dz[16 x E] is the A operand (fp32 -> bf16 on load), the
// bf16 copy of w_gate[:D] sits in shared memory as the B operand, and the accumulator fragments of every 64-column
// group are permuted so that lane (g, t) ends up with 16 CONTIGUOUS columns of tokens g and g+8: exactly the 32-byte
// sector it gathers (one LDG.256) from each of the K queue rows of a token; the four lanes of a row cover one full
// 128-byte line.  dz and w_gate are rounded to bf16 (like every other operand of the bf16 path: the queue rows being
// summed are bf16 already); fp32 queues keep the exact SIMT kernel.
//   MMA column n of block (J, q), q = 0..7   <->   d = 64 J + 16 (n >> 1) + 2 q + (n & 1)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kMmaThreads = 128;   // 4 warps x 16 tokens
constexpr int kMmaKC = 2;          // queue rows per token gathered per batch: 2 at 6 CTAs/SM (<= 80 registers) measured 43 us,
                                   // 4 at 4 CTAs/SM 51 us (tools/variants.py) - like combine_fwd, the gather wants warps

template <typename TO, bool EP, int KS, int kMmaKC = m3::kMmaKC, int MINB = 6>     // KS = E / 16 k-steps
__global__ void __launch_bounds__(kMmaThreads, MINB)
dispatch_bwd_gate_mma_kernel(Queue<const __nv_bfloat16> dxq, const int32_t* __restrict__ pos, int T, int K, int D,
                             const float* __restrict__ dz, const float* __restrict__ w_gate, TO* __restrict__ dx) {
  constexpr int E = 16 * KS;
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint32_t wb[];          // [D][E/2] bf16x2: w_gate[d][2i], w_gate[d][2i+1]
  for (int i = threadIdx.x; i < D * (E / 2); i += kMmaThreads) {
    const float2 w = __ldg(reinterpret_cast<const float2*>(w_gate) + i);      // rows 0..D-1 of [Dg][E] are contiguous
    wb[i] = float2_to_bf16x2(w.x, w.y);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_tiles = (T + 15) / 16;
  for (int tile = blockIdx.x * (kMmaThreads / 32) + warp; tile < n_tiles; tile += gridDim.x * (kMmaThreads / 32)) {
    const int ta_ = tile * 16 + g, tb_ = ta_ + 8;
    const bool va = ta_ < T, vb = tb_ < T;
    const int ta = va ? ta_ : T - 1, tb = vb ? tb_ : T - 1;
    // A operand: dz rows of tokens g / g+8, k = expert
    uint32_t a[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float2 x0 = __ldg(reinterpret_cast<const float2*>(dz + (int64_t)ta * E + 16 * ks + 2 * t));
      const float2 x1 = __ldg(reinterpret_cast<const float2*>(dz + (int64_t)tb * E + 16 * ks + 2 * t));
      const float2 x2 = __ldg(reinterpret_cast<const float2*>(dz + (int64_t)ta * E + 16 * ks + 2 * t + 8));
      const float2 x3 = __ldg(reinterpret_cast<const float2*>(dz + (int64_t)tb * E + 16 * ks + 2 * t + 8));
      a[ks][0] = float2_to_bf16x2(x0.x, x0.y); a[ks][1] = float2_to_bf16x2(x1.x, x1.y);
      a[ks][2] = float2_to_bf16x2(x2.x, x2.y); a[ks][3] = float2_to_bf16x2(x3.x, x3.y);
    }
    for (int J = 0; J < D / 64; ++J) {
      const int col = 64 * J + 16 * t;   // this lane's 16 columns
      float acc[2][16];                  // [token g / g+8][16 contiguous columns]
      for (int k0 = 0; k0 < K; k0 += kMmaKC) {
        // gather: every load of this batch is issued before anything is consumed
        U8 raw[kMmaKC][2];
#pragma unroll
        for (int u = 0; u < kMmaKC; ++u) {
          const int k = k0 + u;
          const int ra = k < K ? __ldg(pos + (int64_t)ta * K + k) : -1;
          const int rb = k < K ? __ldg(pos + (int64_t)tb * K + k) : -1;
#pragma unroll
          for (int i = 0; i < 8; ++i) { raw[u][0].v[i] = 0u; raw[u][1].v[i] = 0u; }
          if (ra >= 0) raw[u][0] = ldg_stream256(dxq.template row<EP>((int64_t)ta * K + k, ra, D) + col);
          if (rb >= 0) raw[u][1] = ldg_stream256(dxq.template row<EP>((int64_t)tb * K + k, rb, D) + col);
        }
        if (k0 == 0) {
          // router term of these 64 columns while the rows are in flight
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            const int d = 64 * J + 16 * (g >> 1) + 2 * q + (g & 1);     // B column g of block (J, q)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              mma_bf16_16816(c, a[ks], wb[d * (E / 2) + 8 * ks + t], wb[d * (E / 2) + 8 * ks + t + 4]);
            acc[0][2 * q] = c[0]; acc[0][2 * q + 1] = c[1];
            acc[1][2 * q] = c[2]; acc[1][2 * q + 1] = c[3];
          }
        }
#pragma unroll
        for (int u = 0; u < kMmaKC; ++u)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = bf16x2_to_float2(raw[u][h].v[i]);
              acc[h][2 * i] += f.x;
              acc[h][2 * i + 1] += f.y;
            }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 0 ? va : vb) {
          TO* dst = dx + (int64_t)(h == 0 ? ta : tb) * D + col;
          if (sizeof(TO) == 4) {
            U8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = __float_as_uint(acc[h][i]);
            stg_stream256(dst, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = __float_as_uint(acc[h][8 + i]);
            stg_stream256(dst + 8, o);
          } else {
            U8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = float2_to_bf16x2(acc[h][2 * i], acc[h][2 * i + 1]);
            stg_stream256(dst, o);
          }
        }
      }
    }
  }
}

static inline int perm_nv(int D) { return m3_ceil_div(D / 8, kLanesPerTok); }

}  // namespace m3

using namespace m3;
typedef __nv_bfloat16 bf16;

static int perm_common_check(int T, int K, int D) {
  if (T < 0 || K < 1 || D < 8) return M3_ERR_ARG;
  if (D % 8 != 0 || D > kLanesPerTok * kMaxVec * 8) return M3_ERR_SHAPE;
  const int nv = perm_nv(D);
  if (!(nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8)) return M3_ERR_SHAPE;
  return M3_OK;
}

#define M3_NV_SWITCH(...)                          \
  switch (nv) {                                    \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break; \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break; \
    case 8: { constexpr int NV = 8; __VA_ARGS__; } break; \
    default: return M3_ERR_SHAPE;                  \
  }
// runs BODY with type aliases TA / TB bound to the two dtypes
#define M3_DTYPE2_SWITCH(da, db, ...)                                                           \
  if (da == M3_F32 && db == M3_F32) { using TA = float; using TB = float; __VA_ARGS__ }         \
  else if (da == M3_F32 && db == M3_BF16) { using TA = float; using TB = bf16; __VA_ARGS__ }    \
  else if (da == M3_BF16 && db == M3_F32) { using TA = bf16; using TB = float; __VA_ARGS__ }    \
  else if (da == M3_BF16 && db == M3_BF16) { using TA = bf16; using TB = bf16; __VA_ARGS__ }    \
  else return M3_ERR_UNSUPPORTED;

template <bool EP>
static int dispatch_fwd_impl(const void* x, int x_dtype, const int32_t* pos, const int32_t* counts,
                             const int32_t* offsets, int T, int K, int D, int E, void* xq, void* const* peer,
                             const int32_t* slot_rank, int xq_dtype, cudaStream_t st) {
  int rc = perm_common_check(T, K, D);
  if (rc) return rc;
  const int nv = perm_nv(D);
  const int tok_ctas = m3_ceil_div(T, kTokPerCta);
  const int grid = tok_ctas + (EP ? 0 : E);
  if (grid == 0) return M3_OK;
  M3_DTYPE2_SWITCH(x_dtype, xq_dtype, {
    Queue<TB> q{(TB*)xq, (TB* const*)peer, slot_rank};
    M3_NV_SWITCH((launch_k(dispatch_fwd_kernel<TA, TB, NV, EP>, grid, kPermThreads, 0, st, (const TA*)x, pos, counts, offsets, T, K, D, tok_ctas, q)))
  })
  M3_LAUNCH_CHECK();
  return M3_OK;
}

template <bool EP>
static int combine_fwd_impl(const void* yq, void* const* peer, const int32_t* slot_rank, int yq_dtype,
                            const int32_t* pos, const float* score, int T, int K, int D, void* out, int out_dtype,
                            void* ysave, cudaStream_t st) {
  int rc = perm_common_check(T, K, D);
  if (rc) return rc;
  if (T == 0) return M3_OK;
  const int nv = perm_nv(D);
  const int grid = m3_ceil_div(T, kTokPerCta);
  M3_DTYPE2_SWITCH(yq_dtype, out_dtype, {
    Queue<const TA> q{(const TA*)yq, (const TA* const*)peer, slot_rank};
    M3_NV_SWITCH((launch_k(combine_fwd_kernel<TA, TB, NV, EP>, grid, kPermThreads, 0, st, q, pos, score, T, K, D, (TB*)out, (TA*)ysave)))
  })
  M3_LAUNCH_CHECK();
  return M3_OK;
}

template <bool EP>
static int combine_bwd_impl(const void* g, int g_dtype, const void* yq, void* const* peer_yq, void* dyq,
                            void* const* peer_dyq, const int32_t* slot_rank, int q_dtype, const int32_t* pos,
                            const float* score, const int32_t* counts, const int32_t* offsets, int T, int K, int D,
                            int E, float* dscore, const void* ysave, cudaStream_t st) {
  int rc = perm_common_check(T, K, D);
  if (rc) return rc;
  const int nv = perm_nv(D);
  const int tok_ctas = m3_ceil_div(T, kTokPerCta);
  const int grid = tok_ctas + (EP ? 0 : E);
  if (grid == 0) return M3_OK;
  M3_DTYPE2_SWITCH(g_dtype, q_dtype, {
    Queue<const TB> qy{(const TB*)yq, (const TB* const*)peer_yq, slot_rank};
    Queue<TB> qd{(TB*)dyq, (TB* const*)peer_dyq, slot_rank};
    M3_NV_SWITCH((launch_k(combine_bwd_kernel<TA, TB, NV, EP>, grid, kPermThreads, 0, st, (const TA*)g, qy, pos, score, counts, offsets, T, K, D, tok_ctas, qd, dscore, (const TB*)ysave)))
  })
  M3_LAUNCH_CHECK();
  return M3_OK;
}

template <bool EP>
static int dispatch_bwd_impl(const void* dxq, void* const* peer, const int32_t* slot_rank, int dxq_dtype,
                             const int32_t* pos, int T, int K, int D, const float* dz, const float* w_gate, int E,
                             void* dx, int dx_dtype, cudaStream_t st) {
  int rc = perm_common_check(T, K, D);
  if (rc) return rc;
  if (T == 0) return M3_OK;
  const int nv = perm_nv(D);
  if (dz == nullptr) {
    const int grid = m3_ceil_div(T, kTokPerCta);
    M3_DTYPE2_SWITCH(dxq_dtype, dx_dtype, {
      Queue<const TA> q{(const TA*)dxq, (const TA* const*)peer, slot_rank};
      M3_NV_SWITCH((launch_k(dispatch_bwd_kernel<TA, TB, NV, EP>, grid, kPermThreads, 0, st, q, pos, T, K, D, (TB*)dx)))
    })
  } else if (dxq_dtype == M3_BF16 && (E == 16 || E == 32 || E == 64) && D % 64 == 0 && (size_t)D * E * 2 <= 96 * 1024 &&
             (reinterpret_cast<uintptr_t>(dxq) & 31u) == 0 && (reinterpret_cast<uintptr_t>(dx) & 31u) == 0 &&
             g_knobs[M3_KNOB_MOVER_VARIANT] != 9) {
    // bf16 queues: router term on the tensor cores (mma.sync), see dispatch_bwd_gate_mma_kernel
    const size_t smem = (size_t)D * E * 2;
    int grid = m3_ceil_div(m3_ceil_div(T, 16), kMmaThreads / 32);
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    Queue<const bf16> q{(const bf16*)dxq, (const bf16* const*)peer, slot_rank};
#define M3_MMA_LAUNCH(TOV, KSV)                                                                              \
  do {                                                                                                       \
    auto kern = dispatch_bwd_gate_mma_kernel<TOV, EP, KSV>;                                                  \
    if (smem > 48 * 1024) {                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
      if (e != cudaSuccess) return (int)e;                                                                   \
    }                                                                                                        \
    launch_k(kern, grid, kMmaThreads, smem, st, q, pos, T, K, D, dz, w_gate, (TOV*)dx);                      \
  } while (0)
#define M3_MMA_KS(TOV) do { if (E == 16) M3_MMA_LAUNCH(TOV, 1); else if (E == 32) M3_MMA_LAUNCH(TOV, 2); else M3_MMA_LAUNCH(TOV, 4); } while (0)
    if (dx_dtype == M3_F32) M3_MMA_KS(float);
    else if (dx_dtype == M3_BF16) M3_MMA_KS(bf16);
    else return M3_ERR_UNSUPPORTED;
#undef M3_MMA_KS
#undef M3_MMA_LAUNCH
  } else {
    // router term: w_gate[:D]^T staged once per CTA -> few, grid-striding CTAs (2 per SM)
    const size_t smem = (size_t)D * E * sizeof(float);
    if (smem > 100 * 1024) return M3_ERR_SHAPE;
    int grid = m3_ceil_div(T, 2 * kTokPerCta);
    if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
    M3_DTYPE2_SWITCH(dxq_dtype, dx_dtype, {
      Queue<const TA> q{(const TA*)dxq, (const TA* const*)peer, slot_rank};
      M3_NV_SWITCH({
        auto kern = dispatch_bwd_gate_kernel<TA, TB, NV, EP>;
        if (smem > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return (int)e;
        }
        launch_k(kern, grid, kPermThreads, smem, st, q, pos, T, K, D, dz, w_gate, E, (TB*)dx);
      })
    })
  }
  M3_LAUNCH_CHECK();
  return M3_OK;
}

extern "C" int m3_dispatch_fwd(const void* x, int x_dtype, const int32_t* pos, const int32_t* counts,
                               const int32_t* offsets, int T, int K, int D, int E, void* xq, int xq_dtype,
                               m3_stream_t stream) {
  M3_CHECK_ARG(x && pos && counts && offsets && xq && E >= 1);
  M3_CHECK_ALIGN16(x); M3_CHECK_ALIGN16(xq);
  return dispatch_fwd_impl<false>(x, x_dtype, pos, counts, offsets, T, K, D, E, xq, nullptr, nullptr, xq_dtype,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" int m3_combine_fwd(const void* yq, int yq_dtype, const int32_t* pos, const float* score, int T, int K,
                              int D, void* out, int out_dtype, m3_stream_t stream) {
  M3_CHECK_ARG(yq && pos && score && out);
  M3_CHECK_ALIGN16(yq); M3_CHECK_ALIGN16(out);
  return combine_fwd_impl<false>(yq, nullptr, nullptr, yq_dtype, pos, score, T, K, D, out, out_dtype, nullptr,
                                 static_cast<cudaStream_t>(stream));
}

extern "C" int m3_combine_bwd(const void* g, int g_dtype, const void* yq, int yq_dtype, const int32_t* pos,
                              const float* score, const int32_t* counts, const int32_t* offsets, int T, int K,
                              int D, int E, void* dyq, int dyq_dtype, float* dscore, m3_stream_t stream) {
  M3_CHECK_ARG(g && yq && pos && score && counts && offsets && dyq && dscore && E >= 1);
  if (yq_dtype != dyq_dtype) return M3_ERR_UNSUPPORTED;
  M3_CHECK_ALIGN16(g); M3_CHECK_ALIGN16(yq); M3_CHECK_ALIGN16(dyq);
  return combine_bwd_impl<false>(g, g_dtype, yq, nullptr, dyq, nullptr, nullptr, yq_dtype, pos, score, counts,
                                 offsets, T, K, D, E, dscore, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int m3_dispatch_bwd(const void* dxq, int dxq_dtype, const int32_t* pos, int T, int K, int D,
                               const float* dz, const float* w_gate, int E, void* dx, int dx_dtype,
                               m3_stream_t stream) {
  M3_CHECK_ARG(dxq && pos && dx);
  M3_CHECK_ARG(dz == nullptr || (w_gate != nullptr && E >= 4 && E % 4 == 0));
  M3_CHECK_ALIGN16(dxq); M3_CHECK_ALIGN16(dx);
  if (dz) { M3_CHECK_ALIGN16(dz); M3_CHECK_ALIGN16(w_gate); }
  return dispatch_bwd_impl<false>(dxq, nullptr, nullptr, dxq_dtype, pos, T, K, D, dz, w_gate, E, dx, dx_dtype,
                                  static_cast<cudaStream_t>(stream));
}

// ---- expert-parallel variants: rows live in the owner rank's queue (peer pointers)
extern "C" int m3_ep_dispatch_fwd(const void* x, int x_dtype, const int32_t* dst_rank, const int32_t* dst_row, int T,
                                  int K, int D, void* const* peer_xq, int xq_dtype, m3_stream_t stream) {
  M3_CHECK_ARG(x && dst_rank && dst_row && peer_xq);
  M3_CHECK_ALIGN16(x);
  return dispatch_fwd_impl<true>(x, x_dtype, dst_row, nullptr, nullptr, T, K, D, 0, nullptr, peer_xq, dst_rank,
                                 xq_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int m3_ep_combine_fwd(void* const* peer_yq, int yq_dtype, const int32_t* dst_rank, const int32_t* dst_row,
                                 const float* score, int T, int K, int D, void* out, int out_dtype, void* ysave,
                                 m3_stream_t stream) {
  M3_CHECK_ARG(peer_yq && dst_rank && dst_row && score && out);
  M3_CHECK_ALIGN16(out);
  if (ysave) M3_CHECK_ALIGN16(ysave);
  return combine_fwd_impl<true>(nullptr, peer_yq, dst_rank, yq_dtype, dst_row, score, T, K, D, out, out_dtype, ysave,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int m3_ep_combine_bwd(const void* g, int g_dtype, void* const* peer_yq, void* const* peer_dyq, int q_dtype,
                                 const int32_t* dst_rank, const int32_t* dst_row, const float* score, int T, int K,
                                 int D, float* dscore, const void* ysave, m3_stream_t stream) {
  M3_CHECK_ARG(g && (peer_yq || ysave) && peer_dyq && dst_rank && dst_row && score && dscore);
  M3_CHECK_ALIGN16(g);
  if (ysave) M3_CHECK_ALIGN16(ysave);
  return combine_bwd_impl<true>(g, g_dtype, nullptr, peer_yq, nullptr, peer_dyq, dst_rank, q_dtype, dst_row, score,
                                nullptr, nullptr, T, K, D, 0, dscore, ysave, static_cast<cudaStream_t>(stream));
}

extern "C" int m3_ep_dispatch_bwd(void* const* peer_dxq, int dxq_dtype, const int32_t* dst_rank,
                                  const int32_t* dst_row, int T, int K, int D, const float* dz, const float* w_gate,
                                  int E, void* dx, int dx_dtype, m3_stream_t stream) {
  M3_CHECK_ARG(peer_dxq && dst_rank && dst_row && dx);
  M3_CHECK_ARG(dz == nullptr || (w_gate != nullptr && E >= 4 && E % 4 == 0));
  M3_CHECK_ALIGN16(dx);
  return dispatch_bwd_impl<true>(nullptr, peer_dxq, dst_rank, dxq_dtype, dst_row, T, K, D, dz, w_gate, E, dx,
                                 dx_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int m3_zero_pad_rows(void* q, int dtype, const int32_t* counts, const int32_t* offsets, int E, int D,
                                int32_t* meta, m3_stream_t stream) {
  M3_CHECK_ARG(q && counts && offsets && E >= 1 && D >= 8 && D % 8 == 0);
  M3_CHECK_ALIGN16(q);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == M3_F32) zero_pad_rows_kernel<float><<<E, kPermThreads, 0, st>>>((float*)q, counts, offsets, D, meta);
  else if (dtype == M3_BF16) zero_pad_rows_kernel<bf16><<<E, kPermThreads, 0, st>>>((bf16*)q, counts, offsets, D, meta);
  else return M3_ERR_UNSUPPORTED;
  M3_LAUNCH_CHECK();
  return M3_OK;
}


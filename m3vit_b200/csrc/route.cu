// Route plan: expert counts, padded prefix sum, stable queue positions, tile map.
//
// Replaces fmoe_cuda.expert_count + assign_pos and the host-side cumsum / D2H sync
// of fmoe.functions.prepare_forward (reached from the reference at
// /root/reference/models/moe/origin/custom_moe_layer.py:255-257).  Differences by
// design: (1) positions are STABLE (rows of one expert keep flat-slot order; fmoe's
// atomicSub order is nondeterministic), (2) every expert queue is padded to a
// multiple of `pad` rows so the grouped GEMM only ever sees full 128-row tiles,
// (3) nothing is read back to the host.
//
// Two launches: per-block histograms, then a block-local stable rank + the
// cross-block prefix (each block re-derives its own prefix from the histogram
// table, so there is no inter-block dependency and no atomics).  One extra block
// of the second launch writes counts / offsets / tile map and reduces the router's
// importance / load partials into the cv^2 balance loss.
// Integer work only; HBM traffic = 8+4 B per slot.  Bit-exact vs the oracle.
#include "common.cuh"

namespace m3 {

#ifndef M3_ROUTE_CHUNK
#define M3_ROUTE_CHUNK 2048
#endif
constexpr int kRouteChunk = M3_ROUTE_CHUNK;   // slots per block
static_assert(kRouteChunk % 1024 == 0, "8 warps x 32 slots per pass, whole batches per lane in the scan");
constexpr int kRouteThreads = 256;
constexpr int kRouteBatches = kRouteChunk / 32;

__global__ void __launch_bounds__(kRouteThreads)
route_count_kernel(const int64_t* __restrict__ idx, int R, int E, int32_t* __restrict__ block_hist) {
  extern __shared__ int hist[];      // [warps][E]: one private histogram per warp (shared-memory atomics on 16 bins from
  constexpr int NW = kRouteThreads / 32;                       // 256 threads serialise on the bins; 32 lanes do not)
  for (int i = threadIdx.x; i < NW * E; i += kRouteThreads) hist[i] = 0;
  pdl_wait();
  pdl_trigger();
  __syncthreads();
  const int base = blockIdx.x * kRouteChunk;
  int* mine = hist + (threadIdx.x >> 5) * E;
  int64_t e[kRouteChunk / kRouteThreads];
#pragma unroll
  for (int i = 0; i < kRouteChunk / kRouteThreads; ++i) {      // all loads first
    const int s = base + i * kRouteThreads + threadIdx.x;
    e[i] = s < R ? idx[s] : -1;
  }
#pragma unroll
  for (int i = 0; i < kRouteChunk / kRouteThreads; ++i)
    if (e[i] >= 0 && e[i] < E) atomicAdd(&mine[(int)e[i]], 1);
  __syncthreads();
  for (int x = threadIdx.x; x < E; x += kRouteThreads) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) c += hist[w * E + x];
    block_hist[(int64_t)blockIdx.x * E + x] = c;
  }
}

__global__ void __launch_bounds__(kRouteThreads)
route_assign_kernel(const int64_t* __restrict__ idx, int R, int E, int pad, int nblk,
                    const int32_t* __restrict__ block_hist, const float* __restrict__ imp_partial,
                    const int32_t* __restrict__ load_partial, int n_partial, int32_t* __restrict__ counts,
                    int32_t* __restrict__ offsets, int32_t* __restrict__ pos,
                    int32_t* __restrict__ tile_expert, float* __restrict__ importance,
                    float* __restrict__ load, float* __restrict__ cv_loss, int32_t* __restrict__ inv_pos) {
  extern __shared__ int sm[];
  int* tot = sm;                     // [E] total count per expert
  int* pre = tot + E;                // [E] count in blocks before this one
  int* off = pre + E;                // [E+1] padded offsets
  int* bh = off + E + 1;             // [kRouteBatches][E] per-warp-batch histogram / prefix
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;

  // this block's slots first: their latency hides behind the histogram-table pass
  const int base = b * kRouteChunk;
  int64_t ee[kRouteBatches / 8];
#pragma unroll
  for (int i = 0; i < kRouteBatches / 8; ++i) {
    const int s = base + (warp + 8 * i) * 32 + lane;
    ee[i] = s < R ? idx[s] : -1;
  }
  for (int i = tid; i < kRouteBatches * E; i += kRouteThreads) bh[i] = 0;
  // totals and this block's cross-block prefix: thread (part, e) sums a strided set of histogram rows in registers, the
  // parts are combined through shared memory (integer adds: order-free; no atomics - 256 threads on 2 E bins serialise)
  int* red_t = bh + kRouteBatches * E;          // [parts][E], the scratch the summary block later reuses
  int* red_p = red_t + kRouteThreads;
  const int parts = kRouteThreads / E;          // >= 2 (E <= 128)
  if (tid < parts * E) {
    const int e = tid % E, part = tid / E;
    int t = 0, q = 0;
    for (int bb = part; bb < nblk; bb += parts) {
      const int v = block_hist[bb * E + e];
      t += v;
      if (bb < b) q += v;
    }
    red_t[tid] = t;
    red_p[tid] = q;
  }
  __syncthreads();
  for (int e = tid; e < E; e += kRouteThreads) {
    int t = 0, q = 0;
    for (int part = 0; part < parts; ++part) { t += red_t[part * E + e]; q += red_p[part * E + e]; }
    tot[e] = t;
    pre[e] = q;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int e = 0; e < E; ++e) {
      off[e] = run;
      run += (tot[e] + pad - 1) / pad * pad;
    }
    off[E] = run;
  }
  // Block `nblk` (one past the last chunk) has no slots: it writes the plan's per-expert outputs and reduces the router's
  // importance / load partials, beside the position blocks instead of as the tail of one of them.
  const bool summary_block = b == nblk;
  // stable rank inside the block: pass 1, per-batch histograms via match_any
  int my_e[kRouteBatches / 8], my_rank[kRouteBatches / 8];
#pragma unroll
  for (int i = 0; i < kRouteBatches / 8; ++i) {
    const int wb = warp + 8 * i;
    const int e = (ee[i] >= 0 && ee[i] < E) ? (int)ee[i] : -1;
    const unsigned mask = __match_any_sync(0xffffffffu, e);
    const int rank = __popc(mask & ((1u << lane) - 1u));
    if (e >= 0 && rank == 0) bh[wb * E + e] = __popc(mask);
    my_e[i] = e;
    my_rank[i] = rank;
  }
  __syncthreads();
  // exclusive scan over batches, per expert: a warp per expert, BPL consecutive batches per lane, shuffle scan over lanes
  constexpr int BPL = kRouteBatches / 32;
  for (int e = warp; e < E; e += kRouteThreads / 32) {
    int v[BPL], sum = 0;
#pragma unroll
    for (int j = 0; j < BPL; ++j) { v[j] = bh[(lane * BPL + j) * E + e]; sum += v[j]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int run = incl - sum;
#pragma unroll
    for (int j = 0; j < BPL; ++j) { bh[(lane * BPL + j) * E + e] = run; run += v[j]; }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRouteBatches / 8; ++i) {
    const int wb = warp + 8 * i;
    const int s = base + wb * 32 + lane;
    if (s < R) {
      const int e = my_e[i];
      const int p = e >= 0 ? off[e] + pre[e] + bh[wb * E + e] + my_rank[i] : -1;
      pos[s] = p;
      if (inv_pos != nullptr && p >= 0) inv_pos[p] = s;      // queue row -> slot (expert parallel: the sorted send order)
    }
  }
  if (summary_block) {
    for (int e = tid; e < E; e += kRouteThreads) counts[e] = tot[e];
    for (int e = tid; e <= E; e += kRouteThreads) offsets[e] = off[e];
    const int ntile = tile_expert != nullptr ? off[E] / pad : 0;
    for (int i = tid; i < ntile; i += kRouteThreads) {
      const int row = i * pad;
      int lo = 0, hi = E;  // largest e with off[e] <= row
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= row) lo = mid; else hi = mid;
      }
      // skip empty experts sharing the same offset: off[lo] <= row < off[lo+1] must hold
      while (lo + 1 < E && off[lo + 1] <= row) ++lo;
      tile_expert[i] = lo;
    }
    if (importance != nullptr && imp_partial != nullptr) {
      // fixed assignment (thread -> partials part, part + P, ...), fixed-order combination: deterministic.  The block
      // is the critical path of the launch (n_partial x E values behind one L2 / DRAM latency per dependent round), so
      // it reads 16-byte quads of experts with up to 8 partial rows in flight per thread.
      const int Q = E / 4;
      const bool quads = E % 4 == 0 && (Q & (Q - 1)) == 0 && Q <= 32 &&
                         ((reinterpret_cast<uintptr_t>(imp_partial) | reinterpret_cast<uintptr_t>(load_partial)) & 15u) == 0;
      int P;      // rows of fs / is to combine
      // [P][E] each: one row per warp in the batch table (this block has no slots, and the scan above is behind a
      // barrier), or one row per part in the scratch behind it (P E <= 256)
      float* fs = reinterpret_cast<float*>(quads ? bh : bh + kRouteBatches * E);
      int* is = quads ? bh + (kRouteThreads / 32) * E : bh + kRouteBatches * E + kRouteThreads;
      if (quads) {
        const int ppw = 32 / Q, q = lane % Q;               // a warp: ppw parts x Q quads
        const int NP = ppw * (kRouteThreads / 32);
        const int part = warp * ppw + lane / Q;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 l = make_int4(0, 0, 0, 0);
        constexpr int U = 8;
        for (int i0 = part; i0 < n_partial; i0 += U * NP) {
          float4 v[U];
          int4 u[U];
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int i = i0 + j * NP;
            const bool ok = i < n_partial;
            v[j] = ok ? __ldg(reinterpret_cast<const float4*>(imp_partial + (int64_t)i * E) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            u[j] = ok ? __ldg(reinterpret_cast<const int4*>(load_partial + (int64_t)i * E) + q) : make_int4(0, 0, 0, 0);
          }
#pragma unroll
          for (int j = 0; j < U; ++j) {
            a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w;
            l.x += u[j].x; l.y += u[j].y; l.z += u[j].z; l.w += u[j].w;
          }
        }
        for (int o = Q; o < 32; o <<= 1) {                  // the parts of this warp (xor tree: fixed order)
          a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
          a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
          l.x += __shfl_xor_sync(0xffffffffu, l.x, o); l.y += __shfl_xor_sync(0xffffffffu, l.y, o);
          l.z += __shfl_xor_sync(0xffffffffu, l.z, o); l.w += __shfl_xor_sync(0xffffffffu, l.w, o);
        }
        if (lane < Q) {
          float* f = fs + warp * E + 4 * q;       // (the scratch is only 4-byte aligned)
          int* g = is + warp * E + 4 * q;
          f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
          g[0] = l.x; g[1] = l.y; g[2] = l.z; g[3] = l.w;
        }
        P = kRouteThreads / 32;
      } else {
        P = kRouteThreads / E > 0 ? kRouteThreads / E : 1;
        if (tid < P * E) {
          const int e = tid % E, part = tid / E;
          float a = 0.f;
          int l = 0;
          for (int i = part; i < n_partial; i += P) {
            a += imp_partial[(int64_t)i * E + e];
            l += load_partial[(int64_t)i * E + e];
          }
          fs[part * E + e] = a;
          is[part * E + e] = l;
        }
      }
      __syncthreads();
      for (int e = tid; e < E; e += kRouteThreads) {
        float a = 0.f;
        int l = 0;
        for (int q = 0; q < P; ++q) { a += fs[q * E + e]; l += is[q * E + e]; }
        importance[e] = a;
        if (load != nullptr) load[e] = (float)l;
        fs[e] = a;                  // P*E >= E: reuse the first row for the cv^2 pass
        is[e] = l;
      }
      __syncthreads();
      if (tid == 0 && cv_loss != nullptr) {
        // cv^2(u) = var_unbiased(u) / (mean(u)^2 + 1e-10); 0 for a single expert
        float loss = 0.f;
        if (E > 1) {
          float mi = 0.f, ml = 0.f;
          for (int e = 0; e < E; ++e) { mi += fs[e]; ml += (float)is[e]; }
          mi /= (float)E; ml /= (float)E;
          float vi = 0.f, vl = 0.f;
          for (int e = 0; e < E; ++e) {
            const float di = fs[e] - mi, dl = (float)is[e] - ml;
            vi = fmaf(di, di, vi); vl = fmaf(dl, dl, vl);
          }
          vi /= (float)(E - 1); vl /= (float)(E - 1);
          loss = vi / (mi * mi + 1e-10f) + vl / (ml * ml + 1e-10f);
        }
        *cv_loss = loss;
      }
    }
  }
}

}  // namespace m3

using namespace m3;

extern "C" size_t m3_route_plan_workspace_bytes(int T, int K, int E) {
  const int64_t R = (int64_t)T * K;
  const int64_t nblk = (R + kRouteChunk - 1) / kRouteChunk;
  return (size_t)(nblk > 0 ? nblk : 1) * E * sizeof(int32_t);
}

extern "C" int m3_route_max_rows(int T, int K, int E, int pad) {
  // sum_e roundup(c_e, pad) <= R + E*(pad-1), rounded up to a whole tile
  const int64_t R = (int64_t)T * K;
  const int64_t cap = (R + (int64_t)E * (pad - 1) + pad - 1) / pad * pad;
  return (int)cap;
}

extern "C" int m3_route_max_tiles(int T, int K, int E, int pad) { return m3_route_max_rows(T, K, E, pad) / pad; }

extern "C" int m3_route_plan(const int64_t* idx, int T, int K, int E, int pad, const float* imp_partial,
                             const int32_t* load_partial, int n_partial, int32_t* counts, int32_t* offsets,
                             int32_t* pos, int32_t* tile_expert, float* importance, float* load, float* cv_loss,
                             int32_t* inv_pos, void* workspace, size_t workspace_bytes, m3_stream_t stream) {
  M3_CHECK_ARG(idx && counts && offsets && pos && workspace);   // tile_expert may be NULL (pad-1 plans)
  M3_CHECK_ARG(T >= 0 && K >= 1 && E >= 1 && pad >= 1);
  M3_CHECK_SHAPE(E <= 128);
  if ((int64_t)T * K > (int64_t)1 << 30) return M3_ERR_SHAPE;
  if (workspace_bytes < m3_route_plan_workspace_bytes(T, K, E)) return M3_ERR_WORKSPACE;
  if (importance != nullptr && (imp_partial == nullptr || load_partial == nullptr)) return M3_ERR_ARG;
  if (cv_loss != nullptr && importance == nullptr) return M3_ERR_ARG;
  const int R = T * K;
  int nblk = m3_ceil_div(R, kRouteChunk);
  if (nblk < 1) nblk = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* block_hist = static_cast<int32_t*>(workspace);
  launch_k(route_count_kernel, nblk, kRouteThreads, (kRouteThreads / 32) * E * sizeof(int), st, idx, R, E, block_hist);
  M3_LAUNCH_CHECK();
  const size_t smem = (size_t)(3 * E + 1 + kRouteBatches * E + 2 * kRouteThreads) * sizeof(int);
  launch_k(route_assign_kernel, nblk + 1, kRouteThreads, smem, st, idx, R, E, pad, nblk, block_hist, imp_partial,
           load_partial, n_partial, counts, offsets, pos, tile_expert, importance, load, cv_loss, inv_pos);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

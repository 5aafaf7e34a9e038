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
static_assert(kRouteChunk % 256 == 0, "8 warps x 32 slots per pass");
constexpr int kRouteThreads = 256;
constexpr int kRouteBatches = kRouteChunk / 32;

__global__ void __launch_bounds__(kRouteThreads)
route_count_kernel(const int64_t* __restrict__ idx, int R, int E, int32_t* __restrict__ block_hist) {
  extern __shared__ int hist[];
  for (int e = threadIdx.x; e < E; e += kRouteThreads) hist[e] = 0;
  pdl_wait();
  pdl_trigger();
  __syncthreads();
  const int base = blockIdx.x * kRouteChunk;
  for (int i = threadIdx.x; i < kRouteChunk; i += kRouteThreads) {
    const int s = base + i;
    if (s < R) {
      const int64_t e = idx[s];
      if (e >= 0 && e < E) atomicAdd(&hist[(int)e], 1);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += kRouteThreads) block_hist[(int64_t)blockIdx.x * E + e] = hist[e];
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

  for (int i = tid; i < 2 * E; i += kRouteThreads) sm[i] = 0;
  for (int i = tid; i < kRouteBatches * E; i += kRouteThreads) bh[i] = 0;
  __syncthreads();
  // totals and this block's cross-block prefix (integer adds: order-free)
  for (int i = tid; i < nblk * E; i += kRouteThreads) {
    const int bb = i / E, e = i % E;
    const int v = block_hist[i];
    if (v) {
      atomicAdd(&tot[e], v);
      if (bb < b) atomicAdd(&pre[e], v);
    }
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
  const int base = b * kRouteChunk;
  int my_e[kRouteBatches / 8], my_rank[kRouteBatches / 8];
#pragma unroll
  for (int i = 0; i < kRouteBatches / 8; ++i) {
    const int wb = warp + 8 * i;
    const int s = base + wb * 32 + lane;
    int e = -1;
    if (s < R) {
      const int64_t ee = idx[s];
      if (ee >= 0 && ee < E) e = (int)ee;
    }
    const unsigned mask = __match_any_sync(0xffffffffu, e);
    const int rank = __popc(mask & ((1u << lane) - 1u));
    if (e >= 0 && rank == 0) bh[wb * E + e] = __popc(mask);
    my_e[i] = e;
    my_rank[i] = rank;
  }
  __syncthreads();
  // exclusive scan over batches, per expert
  for (int e = tid; e < E; e += kRouteThreads) {
    int run = 0;
    for (int wb = 0; wb < kRouteBatches; ++wb) {
      const int v = bh[wb * E + e];
      bh[wb * E + e] = run;
      run += v;
    }
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
      // fixed assignment (thread -> partials p = part, part+P, ...) and fixed-order final sum:
      // deterministic, but 256 threads wide instead of E
      float* fs = reinterpret_cast<float*>(bh + kRouteBatches * E);   // [P][E] scratch behind bh
      int* is = bh + kRouteBatches * E + kRouteThreads;               // [P][E]
      const int P = kRouteThreads / E > 0 ? kRouteThreads / E : 1;
      if (tid < P * E) {
        const int e = tid % E, part = tid / E;
        float a = 0.f;
        int l = 0;
        int i = part;
        for (; i + 3 * P < n_partial; i += 4 * P) {      // four partials in flight; summed in the same order as one by one
          float v[4];
          int u[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[q] = __ldg(imp_partial + (int64_t)(i + q * P) * E + e);
            u[q] = __ldg(load_partial + (int64_t)(i + q * P) * E + e);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) { a += v[q]; l += u[q]; }
        }
        for (; i < n_partial; i += P) {
          a += imp_partial[(int64_t)i * E + e];
          l += load_partial[(int64_t)i * E + e];
        }
        fs[part * E + e] = a;
        is[part * E + e] = l;
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
  launch_k(route_count_kernel, nblk, kRouteThreads, E * sizeof(int), st, idx, R, E, block_hist);
  M3_LAUNCH_CHECK();
  const size_t smem = (size_t)(3 * E + 1 + kRouteBatches * E + 2 * kRouteThreads) * sizeof(int);
  launch_k(route_assign_kernel, nblk + 1, kRouteThreads, smem, st, idx, R, E, pad, nblk, block_hist, imp_partial,
           load_partial, n_partial, counts, offsets, pos, tile_expert, importance, load, cv_loss, inv_pos);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

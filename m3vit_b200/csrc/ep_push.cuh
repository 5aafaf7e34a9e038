// Expert-parallel PUSH half of the all-to-all, run by a few CTAs of the grouped-GEMM kernel itself (ffn_bf16.cu) while
// the other CTAs of the same launch run the GEMM on the rows as they arrive.
//
// Replaces fmoe's global_scatter (grouped ncclSend/ncclRecv sized by host counts, reached from MOEScatter in
// /root/reference/models/moe/origin/custom_moe_layer.py:255-257 when world_size > 1) and the barrier that used to
// separate it from the first expert GEMM.
//
//   * The rows leave in SORTED order (by destination expert), not in token order: segment v of the send schedule
//     (m3_ep_plan) is "my rows for global expert ge", contiguous in the owner's receive queue.  Segments are ordered by
//     LOCAL expert index first and rotated by rank, so that at any moment every rank feeds a different owner and every
//     owner's queue fills expert by expert.
//   * A chunk of kPushChunk consecutive rows is stored by the data warps; after a CTA barrier the signalling warp
//     publishes it: fence.sys + red.release.sys.add on the owner's arrival counter of that local expert (one add per
//     segment the chunk overlaps).  Counters only ever grow; the owner compares them with a running target
//     (wrap-safe), so nothing is reset between calls.
//   * The GEMM CTAs' TMA producer polls counter[e] >= target[e] (ld.acquire.sys) before the first tile of expert e.
//
// mode 0 (forward):  xq[row] = bf16(x[t])                               (same conversion as dispatch_fwd_kernel)
// mode 1 (backward): dyq[row] = score[s] * g[t],  dscore[s] = <g[t], y[s]>   (same arithmetic, lane layout and
//                    reduction order as combine_bwd_kernel: bit-identical dscore)
#pragma once
#include "common.cuh"

namespace m3 {

constexpr int kPushChunk = 1024;    // rows per published chunk (and per unit of work of a pusher CTA)
constexpr int kSegInts = 6;         // vstart, src0, dst_rank, dst_row0, j (local expert at the owner), n

struct EpPush {
  const void* src;                     // x [T][D] (mode 0) or g [T][D] (mode 1)
  int src_f32;                         // 1: fp32 source, 0: bf16
  int mode;
  const int32_t* inv;                  // [R] sorted position -> slot (m3_route_plan, pad 1)
  const int32_t* seg;                  // send schedule: {V, R_live, V x kSegInts} (m3_ep_plan)
  __nv_bfloat16* const* dst_bases;     // [W] peer-mapped receive queues (xq / dyq)
  int32_t* const* cnt_bases;           // [W] peer-mapped arrival counters [E_loc]
  int K, D, cap_rows;
  const float* score;                  // mode 1: [R]
  const __nv_bfloat16* ysave;          // mode 1: [R][D] result rows in slot order
  float* dscore;                       // mode 1: [R]
};

__device__ __forceinline__ void red_release_sys_add(int32_t* p, int v) {
  asm volatile("red.release.sys.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Owner side: wait until `counter` has reached `target` (both only ever grow; the difference is wrap-safe).
static __device__ __noinline__ void ep_wait_arrival(const int32_t* counter, int target) {
  const long long t0 = clock64();
  while (ld_acquire_sys(counter) - target < 0) {
    __nanosleep(200);       // ~130 producers poll the same few lines the peers' reds have to reach
    if (clock64() - t0 > 60000000000LL) {      // ~30 s: a peer died
      if ((threadIdx.x & 31) == 0)
        printf("m3 EP: rows never arrived (block %d, have %d, want %d)\n", blockIdx.x, ld_acquire_sys(counter), target);
      __trap();
    }
  }
  // the rows were written by generic-proxy stores of a peer; the TMA loads that follow read through the async proxy
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Largest segment index whose vstart <= v (empty segments share their successor's vstart and are skipped this way).
__device__ __forceinline__ int seg_find(const int* vstart, int V, int v) {
  int lo = 0, hi = V;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (vstart[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One chunk of rows, NB row buffers per half-warp.
//   prologue: all data threads resolve the chunk's rows in parallel (schedule lookup -> slot -> destination pointer,
//             score) into a 16-byte header per row in shared memory: ONE exposed global-load latency per chunk instead of
//             one per row;
//   steady state: every lane is its own software pipeline over the 16-byte pieces it owns (piece sub + 16 i of every row
//             of its half-warp): cp.async (LDGSTS) of the pieces of row i + NB - 1 into shared memory, then the pieces of
//             row i out of shared memory -> cast / scale / dot -> one 16-byte store each into the owner's queue.  No CTA
//             barrier, no shuffle (except the dscore reduction) and no dependent global load inside the loop.
// Shared memory is the "register file" that keeps ~100 KB of loads in flight per SM.  Measured dead ends (tools/
// push_probe.py): per-lane register loads (one row per half-warp in flight) 5x slower than the stand-alone dispatch kernel
// on all SMs; row-sized TMA bulk copies in either direction cost ~100 clk of TMA issue each (15 GB/s per SM); staging
// whole passes for one large bulk store per contiguous run adds two CTA barriers per 68 rows and lost to plain stores.
template <int NB>
static __device__ __forceinline__ void ep_push_chunk(const EpPush& q, const int* sseg, int V, uint32_t mybuf, int bufb,
                                                     int v0, int v1, int hw, int HW, int sub, uint32_t info0) {
  const int D = q.D, nvec = D / 8;
  const int elb = q.src_f32 ? 32 : 16;        // bytes of an 8-element piece of the source row
  const int srcb = D * (q.src_f32 ? 4 : 2);
  const int ndata = HW * 16;
  {
    const int* vstart = sseg;
    const int* src0 = sseg + V;
    const int* drank = sseg + 2 * V;
    const int* drow0 = sseg + 3 * V;
    named_bar_sync(1, ndata);                  // the previous chunk's headers are no longer read
    for (int r = threadIdx.x; r < v1 - v0; r += ndata) {
      const int v = v0 + r;
      const int sg = seg_find(vstart, V, v);
      const int within = v - vstart[sg];
      const int s = __ldg(q.inv + src0[sg] + within);     // slot
      const int row = drow0[sg] + within;
      // a row beyond the owner's capacity is dropped (but still counted): destination 0
      const unsigned long long dst =
          row < q.cap_rows ? reinterpret_cast<unsigned long long>(q.dst_bases[drank[sg]] + (int64_t)row * D) : 0ull;
      const float sc = q.mode ? __ldg(q.score + s) : 0.f;
      tc::sts128(info0 + r * 16, make_uint4((uint32_t)dst, (uint32_t)(dst >> 32), (uint32_t)s, __float_as_uint(sc)));
    }
    named_bar_sync(1, ndata);
  }
  auto issue = [&](int i) {
    const int r = hw + i * HW;
    if (v0 + r < v1) {
      const uint32_t buf = mybuf + (i % NB) * bufb;
      const int s = (int)tc::lds128(info0 + r * 16).z;
      const int t = s / q.K;
      const uint8_t* src = static_cast<const uint8_t*>(q.src) + (int64_t)t * srcb;
      for (int cidx = sub; cidx < nvec; cidx += 16) {
        cp_async16(buf + cidx * elb, src + cidx * elb);
        if (q.src_f32) cp_async16(buf + cidx * elb + 16, src + cidx * elb + 16);
      }
      if (q.mode) {
        const __nv_bfloat16* ysrc = q.ysave + (int64_t)s * D;
        for (int cidx = sub; cidx < nvec; cidx += 16) cp_async16(buf + srcb + cidx * 16, ysrc + cidx * 8);
      }
    }
    cp_async_commit();
  };
  auto consume = [&](int i) {
    cp_async_wait<NB - 1>();                   // (every lane reads back only the pieces it fetched itself)
    const int r = hw + i * HW;
    const bool valid = v0 + r < v1;
    const uint32_t buf = mybuf + (i % NB) * bufb;
    uint4 info = make_uint4(0, 0, 0, 0);
    if (valid) info = tc::lds128(info0 + r * 16);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(((unsigned long long)info.y << 32) | info.x);
    const bool live = valid && dst != nullptr;
    if (q.mode == 0) {
      if (live)
        for (int cidx = sub; cidx < nvec; cidx += 16) {
          if (q.src_f32) {
            Raw8<float> a;
            a.a = tc::lds128(buf + cidx * 32);
            a.b = tc::lds128(buf + cidx * 32 + 16);
            store8<__nv_bfloat16>(dst + cidx * 8, cvt8(a));
          } else {
            stg_stream(dst + cidx * 8, tc::lds128(buf + cidx * 16));
          }
        }
    } else {
      const float sc = __uint_as_float(info.w);
      float dot = 0.f;
      if (valid)
        for (int cidx = sub; cidx < nvec; cidx += 16) {
          Vec8 gv;
          if (q.src_f32) {
            Raw8<float> a;
            a.a = tc::lds128(buf + cidx * 32);
            a.b = tc::lds128(buf + cidx * 32 + 16);
            gv = cvt8(a);
          } else {
            Raw8<__nv_bfloat16> a;
            a.a = tc::lds128(buf + cidx * 16);
            gv = cvt8(a);
          }
          Raw8<__nv_bfloat16> yr;
          yr.a = tc::lds128(buf + srcb + cidx * 16);
          if (live) {
            Vec8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = sc * gv.v[j];
            store8<__nv_bfloat16>(dst + cidx * 8, o);
          }
          const Vec8 yv = cvt8(yr);
#pragma unroll
          for (int j = 0; j < 8; ++j) dot = fmaf(gv.v[j], yv.v[j], dot);
        }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (valid && sub == 0) q.dscore[info.z] = dot;
    }
  };
  const int iters = (v1 - v0 + HW - 1) / HW;   // warp-uniform (a half-warp without a row in the last pass idles)
  for (int i = 0; i < NB - 1; ++i) issue(i);
  for (int i = 0; i < iters; ++i) {
    issue(i + NB - 1);
    consume(i);
  }
  cp_async_wait<0>();
}

// Runs on ALL threads of a pusher CTA.  The last warp only publishes finished chunks, the others move rows.
static __device__ void ep_push_rows(const EpPush& q, int cta, int nctas, uint8_t* smem, int smem_bytes) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int V = q.seg[0], R_live = q.seg[1];
  int* sseg = reinterpret_cast<int*>(smem);
  for (int i = tid; i < V * kSegInts; i += nthr) {      // struct-of-arrays copy of the schedule
    const int sg = i / kSegInts, f = i % kSegInts;
    sseg[f * V + sg] = q.seg[2 + i];
  }
  __syncthreads();
  const int* vstart = sseg;
  const int* drank = sseg + 2 * V;
  const int* dj = sseg + 4 * V;
  const int seg_bytes = (V * kSegInts * 4 + 127) & ~127;
  const int info_bytes = kPushChunk * 16;     // per-row headers of the chunk in flight
  const int D = q.D;
  const int bufb = D * (q.src_f32 ? 4 : 2) + (q.mode ? D * 2 : 0);      // source row (+ y row)
  const int nwarps = nthr >> 5, warp = tid >> 5, lane = tid & 31;
  const int dw = nwarps - 1;                  // data warps
  const int HW = dw * 2;                      // half-warps = rows per pass
  int NB = (smem_bytes - seg_bytes - info_bytes) / (HW * bufb);
  NB = NB > 4 ? 4 : NB;
  if (NB < 2) return;      // (rows too wide for the stage budget: D > ~1400 with fp32 sources; the host checks)
  auto chunk_rows = [&](int c) { const int r = R_live - c * kPushChunk; return r < kPushChunk ? r : kPushChunk; };
  if (warp == dw) {        // signaller
    for (int c = cta; c * kPushChunk < R_live; c += nctas) {
      named_bar_sync(2, nthr);           // every row of chunk c has been stored by the data warps
      const int v0 = c * kPushChunk, v1 = v0 + chunk_rows(c);
      const int sg0 = seg_find(vstart, V, v0);
      for (int sg = sg0 + lane; sg < V && vstart[sg] < v1; sg += 32) {
        const int a = vstart[sg] > v0 ? vstart[sg] : v0;
        const int e1 = sg + 1 < V ? vstart[sg + 1] : R_live;
        const int b = e1 < v1 ? e1 : v1;
        if (b > a) {
          __threadfence_system();        // cumulative: covers the data warps' stores ordered before the CTA barrier
          red_release_sys_add(q.cnt_bases[drank[sg]] + dj[sg], b - a);
        }
      }
    }
    return;
  }
  const int sub = lane & 15, hw = warp * 2 + (lane >> 4);
  const uint32_t info0 = tc::smem_u32(smem + seg_bytes);
  const uint32_t mybuf = info0 + info_bytes + hw * NB * bufb;
  for (int c = cta; c * kPushChunk < R_live; c += nctas) {
    const int v0 = c * kPushChunk, v1 = v0 + chunk_rows(c);
    if (NB == 2) ep_push_chunk<2>(q, sseg, V, mybuf, bufb, v0, v1, hw, HW, sub, info0);
    else if (NB == 3) ep_push_chunk<3>(q, sseg, V, mybuf, bufb, v0, v1, hw, HW, sub, info0);
    else ep_push_chunk<4>(q, sseg, V, mybuf, bufb, v0, v1, hw, HW, sub, info0);
    named_bar_sync(2, nthr);             // hand the chunk to the signaller
  }
}

}  // namespace m3

// bf16 expert FFN on the 5th-generation tensor cores (sm_100a): TMA-fed tcgen05.mma
// with fp32 accumulators in TMEM, grouped over the padded expert queues.
//
// Replaces fmoe's FMoELinear pair (per-expert cuBLAS GEMM loop, reached from
// /root/reference/models/moe/origin/custom_moe_layer.py:36-44) and its backward.
//
//   gg_kernel    C[256-row pair tile, BN] = A[rows,Kd] * B_e[N,Kd]^T   (both K-major), one tile per CTA PAIR
//                (cluster of 2, tcgen05.mma.cta_group::2: each CTA loads its 128 rows of A and half of B).
//                Persistent pairs over (m_pair_tile, n_tile); warp 0 = TMA producer, warp 1 = MMA issuer
//                (leader CTA, one thread) + TMEM owner, warps 2-9 = epilogue in two groups that alternate
//                accumulator buffers (tcgen05.ld -> bias / GELU (+GELU') / x saved GELU' -> bf16 -> private
//                swizzled smem box -> coalesced 128-byte row segments).  Smem ring of {A 128x64, B (BN/2)x64}
//                stages (SWIZZLE_128B); TMEM accumulator double-buffered so the epilogue of tile i overlaps
//                the MMAs of tile i+1.
//   wgrad_kernel dW_e[M,N] = sum_rows X1[rows,M]^T X2[rows,N]     (both MN-major operands,
//                read straight from the row-major queues - no transposes in memory)
//
// Expert queues are padded to M3_PAD_ROWS = 256 rows (route plan), so every tile is full and the expert
// of a tile is a table lookup; padding rows are zero, so they add nothing to dW.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace m3 {
namespace tc {

constexpr int BM = 128;  // UMMA M (cta_group::1): accumulator row i <-> TMEM lane i
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;        // wgrad kernel: 2 + 4 warps
constexpr int kEpiWarps = 8;          // gg kernel: two warps per TMEM lane quarter
constexpr int kGGThreads = 64 + kEpiWarps * 32;
constexpr int BOX_BYTES = BM * 64 * 2;  // one [128 rows][64 bf16] swizzle-128B box = 16 KB

enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_FC1 = 2, EPI_DGELU = 3 };

struct GGParams {
  const int32_t* offsets;      // [E+1] padded queue offsets (offsets[E] = rows in use)
  const int32_t* tile_expert;  // [rows/128]
  int E, N, Kd;
  const float* bias;           // [E][N]   (EPI_BIAS, EPI_FC1)
  int save_out2;               // EPI_FC1: also store gelu'(pre-activation) for the backward pass (training)
  __nv_bfloat16* out;          // [rows][N]
  __nv_bfloat16* out2;         // EPI_FC1: gelu'(pre-activation)
};

constexpr int WBOX_BYTES = 32 * 64 * 2;  // one warp's [32 rows][64 bf16] swizzle-128B box = 4 KB

template <int BN, int EPI, int NCTA>
struct GGCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / NCTA) * BK * 2;      // a CTA pair splits the B tile
  static constexpr int STAGE = A_BYTES + B_BYTES;
  // epilogue staging: one private 4 KB box per epilogue warp per output tensor (+ per TMA-loaded aux input)
  static constexpr int N_OUT = 1;                            // outputs share one transpose box, flushed in turn
  static constexpr int N_AUX = (EPI == EPI_DGELU) ? 1 : 0;   // EPI_DGELU: gelu'(pre-activation) saved by fc1
  static constexpr int WARP_STAGING = (N_OUT + N_AUX) * WBOX_BYTES;
  static constexpr int STAGING = kEpiWarps * WARP_STAGING;
  static constexpr int BUDGET = 227 * 1024 - 1024 - 512 - STAGING;
  static constexpr int STAGES = (BUDGET / STAGE) < 6 ? (BUDGET / STAGE) : 6;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM = STAGES * STAGE + STAGING + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(STAGES >= 3, "smem ring too shallow");
};

__device__ __forceinline__ uint32_t pack_bf16x2(f32x2 v) {
  float a, b;
  unpk2(v, a, b);
  return float2_to_bf16x2(a, b);
}
__device__ __forceinline__ f32x2 unpack_bf16x2(uint32_t u) {
  const float2 f = bf16x2_to_float2(u);
  return pk2(f.x, f.y);
}
// 16-byte chunk c (0..7) of row r inside a [rows][64 bf16] swizzle-128B box (what TMA reads / writes)
__device__ __forceinline__ uint32_t box_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

template <int BN, int EPI, int NCTA>
__global__ void __launch_bounds__(kGGThreads, 1)
gg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
          const __grid_constant__ CUtensorMap tmAux, GGParams p) {
  using Cfg = GGCfg<BN, EPI, NCTA>;
  // NCTA == 2: CTA pair (cluster of 2) sharing one 256 x BN MMA tile, see tc_common.cuh
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader_cta = cta_rank == 0;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NB = BN / 64;  // 64-column boxes per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + STAGES * Cfg::STAGE;                 // staging boxes (1024-aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + Cfg::STAGING);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* aux_full = tempty + 2;                            // [kEpiWarps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NCTA); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], NCTA * 4 * 32); }
      for (int w = 0; w < kEpiWarps; ++w) mbar_init(&aux_full[w], 1);
      fence_barrier_init();
    }
    __syncwarp();
    if (NCTA == 2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  if (NCTA == 2) cluster_sync(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tile schedule: a "unit" (CTA or CTA pair) walks pair-tiles pt = unit, unit + n_units, ...;
  // pair-tile pt covers M-tiles (pt / n_tiles) * NCTA + cta_rank (queues are padded to NCTA*128 rows,
  // so both halves of a pair belong to the same expert) and N-tile pt % n_tiles.
  const int n_tiles = p.N / BN;
  const int m_tiles = p.offsets[p.E] / BM;
  const int total = (m_tiles / NCTA) * n_tiles;
  const int kchunks = p.Kd / BK;
  const int unit = blockIdx.x / NCTA, n_units = gridDim.x / NCTA;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < total; tile += n_units) {
        const int m_blk = (tile / n_tiles) * NCTA + (int)cta_rank, n_blk = tile % n_tiles;
        const int e = p.tile_expert[(m_blk * BM) / M3_PAD_ROWS];
        const int b_row = e * p.N + n_blk * BN + (int)cta_rank * (BN / NCTA);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE;
          if (NCTA == 2) {
            // both CTAs' bytes are accounted on the LEADER's full barrier
            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
            if (leader_cta) mbar_expect_tx(&full[stage], Cfg::STAGE * NCTA);
            else mbar_arrive_remote(bar);
            tma_load_2d_2sm(sa, &tmA, bar, kc * BK, m_blk * BM);
            tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmB, bar, kc * BK, b_row);
          } else {
            mbar_expect_tx(&full[stage], Cfg::STAGE);
            tma_load_2d(sa, &tmA, &full[stage], kc * BK, m_blk * BM);
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[stage], kc * BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader_cta) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * NCTA, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (int tile = unit; tile < total; tile += n_units) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE);
          const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(a_base + k * UMMA_K * 2, 0, 1024);
            const uint64_t bdesc = make_smem_desc(b_base + k * UMMA_K * 2, 0, 1024);
            if (NCTA == 2) umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kc | k) != 0);
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (kc | k) != 0);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (NCTA == 2) umma_commit_2sm(&empty[stage], 3); else umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (NCTA == 2) umma_commit_2sm(&tfull[acc], 3); else umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---- epilogue.  Every warp is an independent pipeline: TMEM -> registers -> bias / GELU / GELU'
    // (packed f32x2) -> bf16 -> its private swizzled [32 x 64] smem box -> its own TMA store.  No
    // CTA-wide barrier.  Warp w reads TMEM lanes 32*(w%4).. (hardware rule); warps 2-5 serve
    // accumulator buffer 0 (even tiles of this CTA), warps 6-9 buffer 1 (odd tiles), so two tiles
    // are always in flight in the epilogue.
    const int q = warp & 3;
    const int ew = warp - 2;                     // 0..7
    const uint32_t grp = (uint32_t)ew >> 2;      // accumulator buffer served
    uint8_t* my = stg + ew * Cfg::WARP_STAGING;
    uint8_t* box = my;                     // [32 rows][64 bf16] transpose box (swizzled, conflict-free both ways)
    uint8_t* ax = my + WBOX_BYTES;
    uint64_t* my_aux = &aux_full[ew];
    // registers -> swizzled box (one row per lane) -> coalesced 128-B row segments in global memory.
    // Plain st.global: fire-and-forget, so the box is reusable right away (a TMA store would have to
    // drain first, and that latency set the tile period through the 2-deep TMEM pipeline).
    auto flush = [&](__nv_bfloat16* dst, int row0, int col) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), ch = lane & 7;
        const uint4 u = *reinterpret_cast<const uint4*>(box + box_off(rr, ch));
        stg_stream(dst + (int64_t)(row0 + rr) * p.N + col + ch * 8, u);
      }
      __syncwarp();
    };
    uint32_t aux_uses = 0;
    const int first = unit + (int)grp * n_units;
    const uint32_t tempty_remote = NCTA == 2 ? mapa_u32(smem_u32(&tempty[grp]), 0) : 0u;
    if (EPI == EPI_DGELU && lane == 0 && first < total) {
      mbar_expect_tx(my_aux, WBOX_BYTES);
      tma_load_2d(ax, &tmAux, my_aux, (first % n_tiles) * BN, ((first / n_tiles) * NCTA + (int)cta_rank) * BM + q * 32);
    }
    uint32_t it = grp;
    for (int tile = first; tile < total; tile += 2 * n_units, it += 2) {
      const int m_blk = (tile / n_tiles) * NCTA + (int)cta_rank, n_blk = tile % n_tiles;
      const int e = p.tile_expert[(m_blk * BM) / M3_PAD_ROWS];
      const int row0 = m_blk * BM + q * 32;
      mbar_wait(&tfull[grp], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < NB; ++cb) {
        const int col = n_blk * BN + cb * 64;
        float v[64];
        // bias slice of this column block: issued BEFORE the TMEM read so that its latency hides behind it
        // (loaded after, the first add stalled on it: 14 % of the fc1 kernel's stall samples)
        float4 bb[16];
        if (EPI == EPI_BIAS || EPI == EPI_FC1) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + (int64_t)e * p.N + col);
#pragma unroll
          for (int j = 0; j < 16; ++j) bb[j] = __ldg(b4 + j);
        }
        const uint32_t taddr = tmem_base + grp * BN + cb * 64 + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        if (cb == NB - 1) {  // last TMEM read of this accumulator: hand it back to the MMA warp early
          tcgen05_fence_before();
          if (NCTA == 2 && !leader_cta) mbar_arrive_remote(tempty_remote);   // the MMA issuer lives in the leader CTA
          else mbar_arrive(&tempty[grp]);
        }
        f32x2 w2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) w2[j] = pk2(v[2 * j], v[2 * j + 1]);
        if (EPI == EPI_BIAS || EPI == EPI_FC1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 b = bb[j];
            w2[2 * j] = add2(w2[2 * j], pk2(b.x, b.y));
            w2[2 * j + 1] = add2(w2[2 * j + 1], pk2(b.z, b.w));
          }
        }
        if (EPI == EPI_FC1) {
          if (p.save_out2) {
            // training: h = gelu(z) goes on to fc2 (and is kept for dW2); the backward pass only ever
            // needs gelu'(z), so THAT is saved instead of z: dgelu becomes one multiply per element
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {      // 8 pairs = two 16-byte chunks per batch
              f32x2 gl[8], gr[8];
              gelu_fast_grad2_batch<8>(&w2[8 * c2], gl, gr);
#pragma unroll
              for (int i = 0; i < 8; ++i) w2[8 * c2 + i] = gl[i];
              *reinterpret_cast<uint4*>(box + box_off(lane, 2 * c2)) =
                  make_uint4(pack_bf16x2(gr[0]), pack_bf16x2(gr[1]), pack_bf16x2(gr[2]), pack_bf16x2(gr[3]));
              *reinterpret_cast<uint4*>(box + box_off(lane, 2 * c2 + 1)) =
                  make_uint4(pack_bf16x2(gr[4]), pack_bf16x2(gr[5]), pack_bf16x2(gr[6]), pack_bf16x2(gr[7]));
            }
            flush(p.out2, row0, col);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) w2[j] = gelu_fast2(w2[j]);
          }
        }
        if (EPI == EPI_DGELU) {
          mbar_wait(my_aux, aux_uses & 1);
          ++aux_uses;
          uint4 hraw[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) hraw[c] = *reinterpret_cast<const uint4*>(ax + box_off(lane, c));
          __syncwarp();
          if (lane == 0) {  // aux box consumed into registers: prefetch the next one behind the math
            int nt = tile, ncb = cb + 1;
            if (ncb == NB) { ncb = 0; nt += 2 * n_units; }
            if (nt < total) {
              mbar_expect_tx(my_aux, WBOX_BYTES);
              tma_load_2d(ax, &tmAux, my_aux, (nt % n_tiles) * BN + ncb * 64,
                          ((nt / n_tiles) * NCTA + (int)cta_rank) * BM + q * 32);
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t hw[4] = {hraw[c].x, hraw[c].y, hraw[c].z, hraw[c].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) w2[4 * c + i] = mul2(w2[4 * c + i], unpack_bf16x2(hw[i]));   // * gelu'(z)
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 u;
          u.x = pack_bf16x2(w2[4 * c]); u.y = pack_bf16x2(w2[4 * c + 1]);
          u.z = pack_bf16x2(w2[4 * c + 2]); u.w = pack_bf16x2(w2[4 * c + 3]);
          *reinterpret_cast<uint4*>(box + box_off(lane, c)) = u;
        }
        flush(p.out, row0, col);
      }
    }
  }
  tcgen05_fence_before();
  if (NCTA == 2) cluster_sync(); else __syncthreads();     // a pair's smem / TMEM stay alive until both are done
  if (warp == 1) {
    __syncwarp();
    if (NCTA == 2) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------- wgrad
// dW[e][m0..m0+128][n0..n0+BN] = sum over the expert's rows r of X1[r][m] * X2[r][n]
struct WGParams {
  const int32_t* offsets;
  int M, N;       // dW_e is [M][N]; X1 is [rows][M], X2 is [rows][N]
  float* dW;      // [S][E][M][N]  (S = gridDim.z / E row-splits; S == 1: the final gradient)
  float* db;      // [S][E][M] column sums of X1 over the split's rows (bias gradient)
  int E;
};

template <int BN>
struct WGCfg {
  static constexpr int A_BYTES = BK * BM * 2;   // 2 boxes of [64 rows][64 cols]
  static constexpr int B_BYTES = BK * BN * 2;   // BN/64 boxes
  static constexpr int BOX = BK * 64 * 2;       // 8192 B
  static constexpr int ONES_BYTES = BK * 64 * 2;  // [64 k][64 n] box of bf16 ones, one per stage, right behind B:
  static constexpr int TMA_BYTES = A_BYTES + B_BYTES;          // what TMA delivers per stage
  static constexpr int STAGE = TMA_BYTES + ONES_BYTES;         // B | ones is ONE contiguous MN-major operand
  static constexpr int STAGES = (BN <= 128) ? 5 : 3;
  static constexpr int TMEM_COLS = 256;         // BN accumulator columns + 16 bias-gradient columns
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
  static_assert(BN + 16 <= TMEM_COLS, "TMEM");
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2, WGParams p) {
  using Cfg = WGCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-K over the expert's rows: blockIdx.z = split * E + e, split s owns a contiguous range of
  // 64-row chunks, partial results are reduced in a fixed order afterwards (deterministic)
  const int e = blockIdx.z % p.E, split = blockIdx.z / p.E, nsplit = gridDim.z / p.E;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // db = X1^T * 1: the n0 == 0 CTAs widen their MMA from N = BN to N = BN + 16, the extra 16
  // B columns being a box of ones that sits right behind the B boxes of every stage (an extra MMA
  // would re-read the whole A tile from smem and cost as much as the main one).
  const bool with_db = (blockIdx.y == 0) && (p.db != nullptr);
  if (with_db) {
    for (int st = 0; st < STAGES; ++st) {
      uint32_t* ones = reinterpret_cast<uint32_t*>(smem + st * Cfg::STAGE + Cfg::TMA_BYTES);
      for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += kThreads) ones[i] = 0x3F803F80u;
    }
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm1); tma_prefetch_desc(&tm2); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int all_chunks = (p.offsets[e + 1] - p.offsets[e]) / BK;   // queues are padded to M3_PAD_ROWS (a multiple of BK)
  const int per_split = (all_chunks + nsplit - 1) / nsplit;
  const int c_begin = min(split * per_split, all_chunks), c_end = min(c_begin + per_split, all_chunks);
  const int r0 = p.offsets[e] + c_begin * BK;
  const int kchunks = c_end - c_begin;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], Cfg::TMA_BYTES);
        uint8_t* sa = smem + stage * Cfg::STAGE;
        const int r = r0 + kc * BK;
#pragma unroll
        for (int b = 0; b < BM / 64; ++b) tma_load_2d(sa + b * Cfg::BOX, &tm1, &full[stage], m0 + b * 64, r);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_load_2d(sa + Cfg::A_BYTES + b * Cfg::BOX, &tm2, &full[stage], n0 + b * 64, r);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = with_db ? make_idesc_bf16(BM, BN + 16, 1, 1) : make_idesc_bf16(BM, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // MN-major: 16 k-rows = 2 swizzle groups of 8 rows (1024 B each)
          const uint64_t adesc = make_smem_desc(a_base + k * 2048, Cfg::BOX, 1024);
          const uint64_t bdesc = make_smem_desc(b_base + k * 2048, Cfg::BOX, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (kc | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const int64_t se = (int64_t)split * p.E + e;
    const int64_t sbe = se;                                           // db partial index [S][E][M]
    float* dst = p.dW + (se * p.M + row) * p.N + n0;
    if (kchunks == 0) {  // expert received no rows: dW_e = 0
#pragma unroll 1
      for (int c = 0; c < BN / 4; ++c) *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (with_db) p.db[sbe * p.M + row] = 0.f;
    } else {
      mbar_wait(tfull, 0);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + c * 32 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(dst + c * 32 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (with_db) {
        float v[32];   // 16 identical columns (+16 unused): every column of X1^T * ones is the column sum
        tmem_ld_32x32(tmem_base + BN + ((uint32_t)(q * 32) << 16), v);
        p.db[sbe * p.M + row] = v[0];
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2-D bf16 row-major tensor [rows][cols], box [box_rows][64 cols] (128 B inner), SWIZZLE_128B
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return M3_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M3_OK : M3_ERR_ARG;
}

// BN = 256 leaves too little smem for the two-output epilogues: those use 192 or 128
static int pick_bn(int N, bool heavy_epilogue) {
  if (N % 192 == 0) return 192;
  if (N % 256 == 0 && !heavy_epilogue) return 256;
  return N % 128 == 0 ? 128 : 0;
}

constexpr int kGGNcta = 2;   // CTA pairs: halves the per-SM weight traffic (the GEMMs are L2 -> SM bound at K = 384)

// SMs the persistent grouped GEMMs may occupy.  Expert parallelism with overlap (ep.py) lowers it so that the NVLink row
// movers of the other half-batch find free SMs while a GEMM runs (they need ~12-20 SMs: tools/ep_overlap_probe.py).
static int g_gemm_sms = kNumSMs;

template <int BN, int EPI>
static int launch_gg_t(const CUtensorMap* maps, const GGParams& p, int max_tiles, cudaStream_t st) {
  using Cfg = GGCfg<BN, EPI, kGGNcta>;
  auto kern = gg_kernel<BN, EPI, kGGNcta>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return (int)e;
  int grid = max_tiles < g_gemm_sms ? max_tiles : g_gemm_sms;
  grid = grid / kGGNcta * kGGNcta;
  if (grid < kGGNcta) grid = kGGNcta;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGGThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kGGNcta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], p);
  if (e != cudaSuccess) return (int)e;
  M3_LAUNCH_CHECK();
  return M3_OK;
}

// A [cap_rows][Kd] bf16, B [E*N][Kd] bf16 -> out [cap_rows][N] (+ out2 / aux of the same shape)
template <int EPI>
static int launch_gg(const void* A, const void* B, void* out, void* out2, const void* aux, GGParams p, int cap_rows,
                     cudaStream_t st) {
  constexpr bool heavy = (EPI == EPI_FC1 || EPI == EPI_DGELU);
  const int BN = pick_bn(p.N, heavy);
  if (BN == 0 || p.Kd % BK != 0 || cap_rows % (BM * kGGNcta) != 0) return M3_ERR_SHAPE;
  CUtensorMap maps[5];
  int rc = make_map(&maps[0], A, (uint64_t)cap_rows, (uint64_t)p.Kd, BM);
  if (rc) return rc;
  rc = make_map(&maps[1], B, (uint64_t)p.E * p.N, (uint64_t)p.Kd, (uint32_t)(BN / kGGNcta));
  if (rc) return rc;
  rc = make_map(&maps[2], out, (uint64_t)cap_rows, (uint64_t)p.N, 32);      // per-warp [32 x 64] boxes
  if (rc) return rc;
  rc = make_map(&maps[3], out2 ? out2 : out, (uint64_t)cap_rows, (uint64_t)p.N, 32);
  if (rc) return rc;
  rc = make_map(&maps[4], aux ? aux : out, (uint64_t)cap_rows, (uint64_t)p.N, 32);
  if (rc) return rc;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out2 = static_cast<__nv_bfloat16*>(out2);
  const int max_tiles = (cap_rows / BM) * (p.N / BN);
  switch (BN) {
    case 128: return launch_gg_t<128, EPI>(maps, p, max_tiles, st);
    case 192: return launch_gg_t<192, EPI>(maps, p, max_tiles, st);
    default:
      if constexpr (!heavy) return launch_gg_t<256, EPI>(maps, p, max_tiles, st);
      return M3_ERR_SHAPE;
  }
}

// out[i] = sum_s part[s][i] (fixed order), float4-wide
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int nsplit, int64_t n4, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<const float4*>(part)[i];
  for (int s = 1; s < nsplit; ++s) {
    const float4 b = reinterpret_cast<const float4*>(part)[(int64_t)s * n4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  reinterpret_cast<float4*>(out)[i] = a;
}

// row-splits per expert so that the wgrad grid fills the GPU once (few local experts under EP)
static int wgrad_splits(int E, int M, int N) {
  const int tiles = (M / BM) * (N / 128) * E;
  int s = kNumSMs / tiles;
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}
static size_t wgrad_ws_bytes(int E, int M, int N) {
  const int S = wgrad_splits(E, M, N);
  return S > 1 ? (size_t)S * E * ((size_t)M * N + M) * sizeof(float) : 0;
}

// dW [E][M][N] fp32 = X1[rows][M]^T X2[rows][N] per expert (+ db [E][M] = column sums of X1)
static int launch_wgrad(const void* X1, const void* X2, const int32_t* offsets, int cap_rows, int E, int M, int N,
                        float* dW, float* db, float* ws, cudaStream_t st) {
  constexpr int BN = 128;
  if (M % BM != 0 || N % BN != 0 || (M * (int64_t)N) % 4 != 0) return M3_ERR_SHAPE;
  CUtensorMap t1, t2;
  int rc = make_map(&t1, X1, (uint64_t)cap_rows, (uint64_t)M, BK);
  if (rc) return rc;
  rc = make_map(&t2, X2, (uint64_t)cap_rows, (uint64_t)N, BK);
  if (rc) return rc;
  using Cfg = WGCfg<BN>;
  auto kern = wgrad_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int S = wgrad_splits(E, M, N);
  float* pW = S > 1 ? ws : dW;
  float* pb = S > 1 ? ws + (size_t)S * E * M * N : db;
  WGParams p{offsets, M, N, pW, pb, E};
  kern<<<dim3(M / BM, N / BN, E * S), kThreads, Cfg::SMEM, st>>>(t1, t2, p);
  M3_LAUNCH_CHECK();
  if (S > 1) {
    const int64_t n4 = (int64_t)E * M * N / 4;
    splitk_reduce_kernel<<<(int)((n4 + 255) / 256), 256, 0, st>>>(pW, S, n4, dW);
    M3_LAUNCH_CHECK();
    const int64_t b4 = (int64_t)E * M / 4;
    splitk_reduce_kernel<<<(int)((b4 + 255) / 256), 256, 0, st>>>(pb, S, b4, db);
    M3_LAUNCH_CHECK();
  }
  return M3_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace tc
}  // namespace m3

using namespace m3;
using namespace m3::tc;
typedef __nv_bfloat16 bf16;

// ffn_fused.cu
int m3_ffn_fused_supported(int D, int H);
int m3_ffn_fused_fwd(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                     int H, const void* w1, const float* b1, const void* w2, const float* b2, void* hpre, void* yq,
                     cudaStream_t st);
int m3_ffn_fused_bwd(const void* dyq, const void* hpre, const int32_t* offsets, const int32_t* tile_expert,
                     int cap_rows, int E, int D, int H, const void* w2t, const void* w1t, void* dhpre, void* h,
                     void* dxq, cudaStream_t st);
// M3_FFN_FUSED=1 selects the single-kernel fc1->GELU->fc2 chain (ffn_fused.cu) where the shape allows.
// It is parity-tested but currently slower than the two CTA-pair GEMMs at D = H = 384 (its N = 64
// chunk MMAs re-read the whole resident x tile from shared memory), so the
// two-kernel path stays the default.
static bool use_fused(int D, int H) {
  static const bool on = [] { const char* v = getenv("M3_FFN_FUSED"); return v != nullptr && v[0] == '1'; }();
  return on && m3_ffn_fused_supported(D, H);
}

// Opaque activation state handed from m3_ffn_fwd to m3_ffn_bwd (bf16): two [cap][H] planes,
//   plane 0 = gelu'(z) (two-kernel path) or z (fused chain kernel),  plane 1 = h = gelu(z) (two-kernel path).
size_t m3_ffn_bf16_saved_bytes(int cap_rows, int H) { return 2 * align256((size_t)cap_rows * H * 2); }

// workspace: forward  : h [cap][H] bf16 (inference only; training keeps h in the saved state)
//            backward : dz [cap][H] bf16 | h [cap][H] bf16 (fused path only) | wgrad split-K partials fp32
size_t m3_ffn_bf16_workspace_bytes(int cap_rows, int D, int H, int E, int backward) {
  (void)E;
  const size_t hbytes = align256((size_t)cap_rows * H * 2);
  if (!backward) return hbytes;
  const size_t wg = wgrad_ws_bytes(E, D, H) > wgrad_ws_bytes(E, H, D) ? wgrad_ws_bytes(E, D, H) : wgrad_ws_bytes(E, H, D);
  return 2 * hbytes + align256(wg);
}

int m3_ffn_fwd_bf16(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                    int H, const void* w1, const float* b1, const void* w2, const float* b2, void* saved, void* yq,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (workspace == nullptr || workspace_bytes < m3_ffn_bf16_workspace_bytes(cap_rows, D, H, E, 0)) return M3_ERR_WORKSPACE;
  if (use_fused(D, H))
    return m3_ffn_fused_fwd(xq, offsets, tile_expert, cap_rows, E, D, H, w1, b1, w2, b2, saved, yq, st);
  const size_t hbytes = align256((size_t)cap_rows * H * 2);
  bf16* gp = static_cast<bf16*>(saved);
  bf16* h = saved ? reinterpret_cast<bf16*>(static_cast<uint8_t*>(saved) + hbytes) : static_cast<bf16*>(workspace);
  GGParams p{};
  p.offsets = offsets; p.tile_expert = tile_expert; p.E = E;
  // fc1: h = gelu(xq W1^T + b1); gelu'(.) saved for backward
  p.N = H; p.Kd = D; p.bias = b1; p.save_out2 = saved != nullptr;
  int rc = launch_gg<EPI_FC1>(xq, w1, h, gp, nullptr, p, cap_rows, st);
  if (rc) return rc;
  // fc2: yq = h W2^T + b2
  p.N = D; p.Kd = H; p.bias = b2; p.save_out2 = 0;
  return launch_gg<EPI_BIAS>(h, w2, yq, nullptr, nullptr, p, cap_rows, st);
}

int m3_ffn_bwd_bf16(const void* xq, const void* saved, const void* dyq, const int32_t* counts, const int32_t* offsets,
                    const int32_t* tile_expert, int cap_rows, int E, int D, int H, const void* w1, const void* w2,
                    const void* w1t, const void* w2t, void* dxq, float* dw1, float* db1, float* dw2, float* db2,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  (void)counts; (void)w1; (void)w2;
  if (workspace_bytes < m3_ffn_bf16_workspace_bytes(cap_rows, D, H, E, 1)) return M3_ERR_WORKSPACE;
  const size_t hbytes = align256((size_t)cap_rows * H * 2);
  bf16* dhpre = static_cast<bf16*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 2 * hbytes);
  const bf16* h;
  int rc;
  if (use_fused(D, H)) {
    // one kernel: dz = (dyq W2) * gelu'(z), h = gelu(z), dxq = dz W1      (saved plane 0 = z)
    bf16* hw = reinterpret_cast<bf16*>(static_cast<uint8_t*>(workspace) + hbytes);
    rc = m3_ffn_fused_bwd(dyq, saved, offsets, tile_expert, cap_rows, E, D, H, w2t, w1t, dhpre, hw, dxq, st);
    if (rc) return rc;
    h = hw;
  } else {
    h = reinterpret_cast<const bf16*>(static_cast<const uint8_t*>(saved) + hbytes);
    GGParams p{};
    p.offsets = offsets; p.tile_expert = tile_expert; p.E = E;
    // dz = (dyq W2) * gelu'(z)                              B = W2^T [E][H][D] (K-major in D)
    p.N = H; p.Kd = D;
    rc = launch_gg<EPI_DGELU>(dyq, w2t, dhpre, nullptr, saved, p, cap_rows, st);
    if (rc) return rc;
    // dxq = dz W1                                           B = W1^T [E][D][H] (K-major in H)
    p.N = D; p.Kd = H;
    rc = launch_gg<EPI_STORE>(dhpre, w1t, dxq, nullptr, nullptr, p, cap_rows, st);
    if (rc) return rc;
  }
  // dW2[e] = dyq_e^T h_e  [D][H];   dW1[e] = dz_e^T xq_e  [H][D]
  // the bias gradients ride along in the same MMA against a tile of ones
  rc = launch_wgrad(dyq, h, offsets, cap_rows, E, D, H, dw2, db2, part, st);
  if (rc) return rc;
  return launch_wgrad(dhpre, xq, offsets, cap_rows, E, H, D, dw1, db1, part, st);
}

int m3_ffn_bf16_set_sm_limit(int sms) {
  if (sms < 2 || sms > kNumSMs) sms = kNumSMs;
  g_gemm_sms = sms;
  return M3_OK;
}

// bf16 expert FFN on the 5th-generation tensor cores (sm_100a): TMA-fed tcgen05.mma
// with fp32 accumulators in TMEM, grouped over the padded expert queues.
//
// Replaces fmoe's FMoELinear pair (per-expert cuBLAS GEMM loop, reached from
// /root/reference/models/moe/origin/custom_moe_layer.py:36-44) and its backward.
//
//   gg_kernel    C[256-row pair tile, BN] = A[rows,Kd] * B_e[N,Kd]^T   (both K-major), one tile per CTA PAIR
//                (cluster of 2, tcgen05.mma.cta_group::2: each CTA loads its 128 rows of A and half of B).
//                Persistent pairs over (m_pair_tile, n_tile); warp 0 = TMA producer, warp 1 = MMA issuer (leader
//                CTA) + TMEM owner - both run their loops CONVERGED, every TMA / MMA / commit predicated on
//                elect.sync inside its asm block (tc_common.cuh) - warps 2.. = 8 or 16 epilogue warps in two groups
//                that alternate accumulator buffers (tcgen05.ld -> bias / GELU (+GELU') / x saved GELU' -> bf16 ->
//                private swizzled smem box -> coalesced row segments).  Smem ring of {A 128x64, B (BN/2)x64}
//                stages (SWIZZLE_128B); TMEM accumulator double-buffered so the epilogue of tile i overlaps
//                the MMAs of tile i+1.  Opt-in variant (bit-identical): 32-column epilogue blocks.
//   wgrad_kernel dW_e[M,N] = sum_rows X1[rows,M]^T X2[rows,N]     (both MN-major operands,
//                read straight from the row-major queues - no transposes in memory); same converged producer / MMA
//                warps
//
// Expert queues are padded to M3_PAD_ROWS = 256 rows (route plan), so every tile is full and the expert
// of a tile is a table lookup; padding rows are zero, so they add nothing to dW.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "philox.cuh"
#include "tc_common.cuh"

namespace m3 {
namespace tc {

constexpr int BM = 128;  // UMMA M (cta_group::1): accumulator row i <-> TMEM lane i
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;        // wgrad kernel: 2 + 4 warps
// gg kernel epilogue warps: EW = 8 (one warp per TMEM lane quarter and accumulator buffer, 64-column register
// blocks) or EW = 16 (two warps per quarter and buffer, interleaved 32-column blocks: twice the warps to hide
// the dependent erf/exp chains of the GELU epilogues behind, at half the registers per thread)
constexpr int gg_threads(int EW) { return 64 + EW * 32; }
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
  int dbg;                     // M3_KNOB_DEBUG (measurement only, results are garbage): 1 = no MMAs, 2 = no TMA loads
  unsigned long long* trace;   // m3_debug_trace_buffer
  int trace_cap;
  // EPI_FC1, training: expert dropout behind the GELU (philox.cuh).  Both saved planes carry the keep-scale, h = m gelu(z)
  // and m gelu'(z) with m in {0, 1/(1-p)}, so the backward pass needs neither the mask nor the random stream again.
  uint32_t drop_thr;           // p * 2^32 (0 = no dropout)
  float drop_inv_keep;         // 1 / (1 - p)
  const RngState* rng;
  // Expert parallel, EPI_BIAS / EPI_STORE ("return store"): the output row of queue row r does not go to out[r] but straight
  // to the rank the row came from, over NVLink, from this epilogue: ret_meta[r] = (source rank << 24) | slot (or -1: padding
  // row, nothing stored), destination = ret_bases[source rank] + slot * N.  The exchange that FastMoE runs as a separate
  // global_gather after the expert GEMM overlaps the GEMM tile by tile; see m3_ep_ffn_fwd in the header.
  const int32_t* ret_meta;     // [rows] or nullptr
  __nv_bfloat16* const* ret_bases;   // [W] peer-mapped [T*K][N] return buffers
};

template <int BN, int EPI, int NCTA, int EW, int CWP>
struct GGCfg {
  static constexpr int CW = CWP;                              // columns per epilogue register block (32 or 64)
  static_assert(CWP == 32 || (CWP == 64 && EW == 8), "epilogue block width");
  static constexpr int WBOX_BYTES = 32 * CW * 2;              // one warp's [32 rows][CW bf16] swizzled box (4 / 2 KB)
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / NCTA) * BK * 2;      // a CTA pair splits the B tile
  static constexpr int STAGE = A_BYTES + B_BYTES;
  // epilogue staging: one private 4 KB box per epilogue warp per output tensor (+ per TMA-loaded aux input)
  static constexpr int N_OUT = 1;                            // outputs share one transpose box, flushed in turn
  static constexpr int N_AUX = (EPI == EPI_DGELU) ? 1 : 0;   // EPI_DGELU: gelu'(pre-activation) saved by fc1
  static constexpr int WARP_STAGING = (N_OUT + N_AUX) * WBOX_BYTES;
  static constexpr int STAGING = EW * WARP_STAGING;
  static constexpr int BUDGET = 227 * 1024 - 1024 - 512 - STAGING;
  static constexpr int MAX_STAGES = 8;
  static constexpr int STAGES = (BUDGET / STAGE) < MAX_STAGES ? (BUDGET / STAGE) : MAX_STAGES;
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
// 16-byte chunk c of row r inside a [rows][CW bf16] box as TMA lays it out: CW = 64 -> 128-byte rows, SWIZZLE_128B
// (chunk ^= row & 7); CW = 32 -> 64-byte rows, SWIZZLE_64B (chunk ^= (row >> 1) & 3).  Conflict-free for the
// row-per-lane writes and for the row-segment reads of the flush alike.
template <int CW>
__device__ __forceinline__ uint32_t box_off(int r, int c) {
  if (CW == 64) return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4);
}

template <int BN, int EPI, int NCTA, int EW, int CWP>
__global__ void __launch_bounds__(gg_threads(EW), 1)
gg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmAux, GGParams p) {
  using Cfg = GGCfg<BN, EPI, NCTA, EW, CWP>;
  constexpr int kEpiWarps = EW;
  constexpr int CW = Cfg::CW, WBOX_BYTES = Cfg::WBOX_BYTES;
  constexpr int NQ = EW / 8;            // warps sharing one (TMEM lane quarter, accumulator buffer)
  constexpr int NCH = CW / 8;           // 16-byte chunks per box row
  // NCTA == 2: CTA pair (cluster of 2) sharing one 256 x BN MMA tile, see tc_common.cuh
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader_cta = cta_rank == 0;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NBW = BN / CW / NQ;  // CW-column blocks of a tile handled by one epilogue warp
  static_assert(BN % (CW * NQ) == 0, "tile width vs epilogue blocks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + STAGES * Cfg::STAGE;                 // staging boxes (1024-aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + Cfg::STAGING);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* aux_full = tempty + 2;                            // [kEpiWarps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + kEpiWarps);
  static_assert((2 * STAGES + 4 + EW) * 8 + 4 <= 512, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NCTA); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], NCTA * (EW / 2) * 32); }
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
  // everything above is private to this CTA (smem, mbarriers, TMEM): under programmatic dependent launch it
  // runs beneath the tail of the previous kernel; global memory is only touched from here on
  pdl_wait();
  pdl_trigger();

  // Tile schedule.  A "unit" (CTA or CTA pair) computes `my_tiles` pair-tiles, local index i = 0 .. my_tiles-1:
  // pair-tile pt = unit + i * n_units covers M-tiles (pt / n_tiles) * NCTA + cta_rank (queues are padded to NCTA*128
  // rows, so both halves of a pair belong to the same expert) and N-tile pt % n_tiles.
  const int n_tiles = p.N / BN;
  const int m_tiles = p.offsets[p.E] / BM;
  const int kchunks = p.Kd / BK;
  const int unit = blockIdx.x / NCTA, n_units = gridDim.x / NCTA;
  const int total = (m_tiles / NCTA) * n_tiles;
  const int my_tiles = unit < total ? (total - unit + n_units - 1) / n_units : 0;
  // local tile index -> (M-tile of this CTA, N-tile)
  auto tile_mblk = [&](int i) { return ((unit + i * n_units) / n_tiles) * NCTA + (int)cta_rank; };
  auto tile_nblk = [&](int i) { return (unit + i * n_units) % n_tiles; };

  // Producer and MMA warps run their loops CONVERGED (all 32 lanes); every TMA / MMA / commit / arrive is issued by one
  // elected lane inside its asm block (tc_common.cuh: the single-thread `if (lane == 0)` form cost ~700 clk of scalar
  // code per 4-MMA k-chunk and bounded every GEMM of the layer).
  // Both loops walk the smem ring in rounds of STAGES k-chunks with the stage index a compile-time constant (every
  // barrier address and descriptor is base + immediate); the flat chunk index f runs over all tiles of this unit.
  const int F = my_tiles * kchunks;
  if (warp == 0) {
    const uint32_t smem_base = smem_u32(smem);
    const bool no_tma = (p.dbg & 2) != 0;
    const uint32_t full0 = NCTA == 2 ? mapa_u32(smem_u32(&full[0]), 0) : smem_u32(&full[0]);   // leader's barriers
    Tracer trc(p.trace, p.trace_cap, 0);
    uint32_t phase = 0;
    int ti = 0, kc = 0, m_blk = 0, b_row = 0;
    for (int f0 = 0; f0 < F; f0 += STAGES) {
#pragma unroll
      for (int st = 0; st < STAGES; ++st) {
        if (f0 + st < F) {
          if (kc == 0) {
            m_blk = tile_mblk(ti);
            const int e = p.tile_expert[(m_blk * BM) / M3_PAD_ROWS];
            b_row = e * p.N + tile_nblk(ti) * BN + (int)cta_rank * (BN / NCTA);
          }
          trc.ev(0x00, f0 + st);
          mbar_wait(&empty[st], phase ^ 1);
          trc.ev(0x01, f0 + st);
          __syncwarp();
          const uint32_t sa = smem_base + st * Cfg::STAGE, bar = full0 + st * 8;
          if (NCTA == 2) {
            // both CTAs' bytes are accounted on the LEADER's full barrier
            if (no_tma) {   // measurement only: no loads, the MMAs chew on whatever the stage holds
              if (leader_cta) mbar_arrive_elect(&full[st]); else mbar_arrive_remote_elect(bar);
            } else {
              if (leader_cta) mbar_expect_tx_elect(&full[st], Cfg::STAGE * NCTA);
              else mbar_arrive_remote_elect(bar);
              tma_load_2d_2sm_elect(sa + Cfg::A_BYTES, &tmB, bar, kc * BK, b_row);
              tma_load_2d_2sm_elect(sa, &tmA, bar, kc * BK, m_blk * BM);
            }
          } else {
            mbar_expect_tx_elect(&full[st], Cfg::STAGE);
            tma_load_2d_elect(sa + Cfg::A_BYTES, &tmB, bar, kc * BK, b_row);
            tma_load_2d_elect(sa, &tmA, bar, kc * BK, m_blk * BM);
          }
          if (++kc == kchunks) {
            kc = 0;
            ++ti;
          }
        }
      }
      phase ^= 1;
    }
    trc.done();
  } else if (warp == 1) {
    if (leader_cta) {
      Tracer trc(p.trace, p.trace_cap, 1);
      constexpr uint32_t idesc = make_idesc_bf16(BM * NCTA, BN, 0, 0);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      // K-major operand tiles [rows][64 bf16], SWIZZLE_128B: 8-row groups 1024 B apart (SBO); K advances 32 B per UMMA_K
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 0), hi = smem_desc_hi(1024);
      const bool no_mma = (p.dbg & 1) != 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0, d_tmem = tm;
      int kc = 0;
      for (int f0 = 0; f0 < F; f0 += STAGES) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) {
          if (f0 + st < F) {
            if (kc == 0) {   // first k-chunk of a tile: its accumulator buffer must have been drained
              trc.ev(0x10, f0 + st);
              mbar_wait(&tempty[acc], acc_phase ^ 1);
              trc.ev(0x11, f0 + st);
              d_tmem = tm + acc * BN;
            }
            mbar_wait(&full[st], phase);
            trc.ev(0x12, f0 + st);
            __syncwarp();
            tcgen05_fence_after();
            const uint32_t a_lo = a_lo0 + (uint32_t)st * (Cfg::STAGE >> 4);
            const uint32_t b_lo = a_lo + (Cfg::A_BYTES >> 4);
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                if (NCTA == 2) umma_bf16_2sm_elect(d_tmem, a_lo + k * (UMMA_K * 2 >> 4), hi, b_lo + k * (UMMA_K * 2 >> 4), hi, idesc, (kc | k) != 0);
                else umma_bf16_elect(d_tmem, a_lo + k * (UMMA_K * 2 >> 4), hi, b_lo + k * (UMMA_K * 2 >> 4), hi, idesc, (kc | k) != 0);
              }
            }
            // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
            if (NCTA == 2) umma_commit_2sm_elect(&empty[st], 3); else umma_commit_elect(&empty[st]);
            trc.ev(0x13, f0 + st);
            if (++kc == kchunks) {
              // accumulator complete -> epilogue warps (of both CTAs)
              if (NCTA == 2) umma_commit_2sm_elect(&tfull[acc], 3); else umma_commit_elect(&tfull[acc]);
              kc = 0;
              acc ^= 1;
              if (acc == 0) acc_phase ^= 1;
            }
          }
        }
        phase ^= 1;
      }
      trc.done();
    }
  } else if (warp < 2 + EW) {
    // ---- epilogue.  Every warp is an independent pipeline: TMEM -> registers -> bias / GELU / GELU'
    // (packed f32x2) -> bf16 -> its private swizzled [32 x 64] smem box -> its own TMA store.  No
    // CTA-wide barrier.  Warp w reads TMEM lanes 32*(w%4).. (hardware rule); warps 2-5 serve
    // accumulator buffer 0 (even tiles of this CTA), warps 6-9 buffer 1 (odd tiles), so two tiles
    // are always in flight in the epilogue.
    Tracer trc(p.trace, p.trace_cap, 2);
    const int q = warp & 3;
    const int ew = warp - 2;                     // 0..EW-1
    const uint32_t grp = (uint32_t)ew / (EW / 2);   // accumulator buffer served
    const int sub = (ew >> 2) % NQ;              // which of the interleaved column blocks (EW = 16)
    uint8_t* my = stg + ew * Cfg::WARP_STAGING;
    const uint32_t box = smem_u32(my);     // [32 rows][CW bf16] transpose box (swizzled, conflict-free both ways)
    uint8_t* ax = my + WBOX_BYTES;
    const uint32_t ax32 = smem_u32(ax);
    uint64_t* my_aux = &aux_full[ew];
    // registers -> swizzled box (one row per lane) -> coalesced row segments (CW * 2 bytes) in global memory.
    // Plain st.global: fire-and-forget, so the box is reusable right away (a TMA store would have to
    // drain first, and that latency set the tile period through the 2-deep TMEM pipeline).
    auto flush = [&](__nv_bfloat16* dst, int row0, int col) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int rr = i * (32 / NCH) + lane / NCH, ch = lane % NCH;
        const uint4 u = lds128(box + box_off<CW>(rr, ch));
        stg_stream(dst + (int64_t)(row0 + rr) * p.N + col + ch * 8, u);
      }
      __syncwarp();
    };
    // return store (expert parallel): every lane carries the destination of ITS row of the warp's 32-row quarter
    auto flush_ret = [&](unsigned long long rowdst, int col) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int rr = i * (32 / NCH) + lane / NCH, ch = lane % NCH;
        const uint4 u = lds128(box + box_off<CW>(rr, ch));
        const unsigned long long d = __shfl_sync(0xffffffffu, rowdst, rr);
        if (d != 0ull) stg_stream(reinterpret_cast<__nv_bfloat16*>(d) + col + ch * 8, u);
      }
      __syncwarp();
    };
    constexpr bool kCanRet = (EPI == EPI_BIAS || EPI == EPI_STORE);
    const bool ret = kCanRet && p.ret_meta != nullptr;
    uint32_t aux_uses = 0;
    const uint32_t tempty_remote = NCTA == 2 ? mapa_u32(smem_u32(&tempty[grp]), 0) : 0u;
    if (EPI == EPI_DGELU && lane == 0 && (int)grp < my_tiles) {
      mbar_expect_tx(my_aux, WBOX_BYTES);
      tma_load_2d(ax, &tmAux, my_aux, tile_nblk(grp) * BN + sub * CW, tile_mblk(grp) * BM + q * 32);
    }
    for (uint32_t it = grp; (int)it < my_tiles; it += 2) {       // local tile index: this group serves every other tile
      const int m_blk = tile_mblk(it), n_blk = tile_nblk(it);
      const int e = p.tile_expert[(m_blk * BM) / M3_PAD_ROWS];
      const int row0 = m_blk * BM + q * 32;
      unsigned long long rowdst = 0ull;
      if (kCanRet && ret) {
        const int meta = __ldg(p.ret_meta + row0 + lane);
        if (meta >= 0)
          rowdst = reinterpret_cast<unsigned long long>(p.ret_bases[meta >> 24] + (int64_t)(meta & 0xffffff) * p.N);
      }
      if (ew == 0) trc.ev(0x20, it);
      mbar_wait(&tfull[grp], (it >> 1) & 1);
      if (ew == 0) trc.ev(0x21, it);
      tcgen05_fence_after();
      if (p.dbg & 4) {   // measurement only: no epilogue work, the accumulator goes straight back
        tcgen05_fence_before();
        if (NCTA == 2 && !leader_cta) mbar_arrive_remote(tempty_remote);
        else mbar_arrive(&tempty[grp]);
        continue;
      }
#pragma unroll 1
      for (int cbi = 0; cbi < NBW; ++cbi) {
        const int cb = cbi * NQ + sub;                 // CW-column block of the tile
        const int col = n_blk * BN + cb * CW;
        float v[CW];
        // bias slice of this column block: issued BEFORE the TMEM read so that its latency hides behind it
        // (loaded after, the first add stalled on it: 14 % of the fc1 kernel's stall samples)
        float4 bb[CW / 4];
        if (EPI == EPI_BIAS || EPI == EPI_FC1) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + (int64_t)e * p.N + col);
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) bb[j] = __ldg(b4 + j);
        }
        const uint32_t taddr = tmem_base + grp * BN + cb * CW + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32(taddr, v);
        if (CW == 64) tmem_ld_32x32(taddr + 32, v + (CW == 64 ? 32 : 0));
        if (ew == 0) trc.ev(0x22, cbi);
        if (cbi == NBW - 1) {  // this warp's last TMEM read of the accumulator: hand it back to the MMA warp early
          tcgen05_fence_before();
          if (NCTA == 2 && !leader_cta) mbar_arrive_remote(tempty_remote);   // the MMA issuer lives in the leader CTA
          else mbar_arrive(&tempty[grp]);
        }
        f32x2 w2[CW / 2];
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) w2[j] = pk2(v[2 * j], v[2 * j + 1]);
        if (EPI == EPI_BIAS || EPI == EPI_FC1) {
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
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
            for (int c2 = 0; c2 < CW / 16; ++c2) {      // 8 pairs = two 16-byte chunks per batch
              f32x2 gl[8], gr[8];
              gelu_fast_grad2_batch<8>(&w2[8 * c2], gl, gr);
              if (p.drop_thr != 0u) {
                const RngState rs = *p.rng;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  float sc[4];
                  dropout_scale4(rs, (uint32_t)(row0 + lane), (uint32_t)((col + 16 * c2) / 4 + q4), p.drop_thr, p.drop_inv_keep, sc);
                  const f32x2 s01 = pk2(sc[0], sc[1]), s23 = pk2(sc[2], sc[3]);
                  gl[2 * q4] = mul2(gl[2 * q4], s01); gl[2 * q4 + 1] = mul2(gl[2 * q4 + 1], s23);
                  gr[2 * q4] = mul2(gr[2 * q4], s01); gr[2 * q4 + 1] = mul2(gr[2 * q4 + 1], s23);
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) w2[8 * c2 + i] = gl[i];
              sts128(box + box_off<CW>(lane, 2 * c2),
                     make_uint4(pack_bf16x2(gr[0]), pack_bf16x2(gr[1]), pack_bf16x2(gr[2]), pack_bf16x2(gr[3])));
              sts128(box + box_off<CW>(lane, 2 * c2 + 1),
                     make_uint4(pack_bf16x2(gr[4]), pack_bf16x2(gr[5]), pack_bf16x2(gr[6]), pack_bf16x2(gr[7])));
            }
            flush(p.out2, row0, col);
          } else {
#pragma unroll
            for (int j = 0; j < CW / 2; ++j) w2[j] = gelu_fast2(w2[j]);
          }
        }
        if (EPI == EPI_DGELU) {
          mbar_wait(my_aux, aux_uses & 1);
          ++aux_uses;
          uint4 hraw[NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c) hraw[c] = lds128(ax32 + box_off<CW>(lane, c));
          __syncwarp();
          if (lane == 0) {  // aux box consumed into registers: prefetch the next one behind the math
            int nt = (int)it, ncbi = cbi + 1;
            if (ncbi == NBW) { ncbi = 0; nt += 2; }
            if (nt < my_tiles) {
              mbar_expect_tx(my_aux, WBOX_BYTES);
              tma_load_2d(ax, &tmAux, my_aux, tile_nblk(nt) * BN + (ncbi * NQ + sub) * CW, tile_mblk(nt) * BM + q * 32);
            }
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const uint32_t hw[4] = {hraw[c].x, hraw[c].y, hraw[c].z, hraw[c].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) w2[4 * c + i] = mul2(w2[4 * c + i], unpack_bf16x2(hw[i]));   // * gelu'(z)
          }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint4 u;
          u.x = pack_bf16x2(w2[4 * c]); u.y = pack_bf16x2(w2[4 * c + 1]);
          u.z = pack_bf16x2(w2[4 * c + 2]); u.w = pack_bf16x2(w2[4 * c + 3]);
          sts128(box + box_off<CW>(lane, c), u);
        }
        if (kCanRet && ret) flush_ret(rowdst, col);
        else flush(p.out, row0, col);
        if (ew == 0) trc.ev(0x23, cbi);
      }
    }
    if (ew == 0) trc.done();
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
  int dbg;        // M3_KNOB_DEBUG (measurement only): 1 = no MMAs, 2 = no TMA loads
  unsigned long long* trace;   // m3_debug_trace_buffer
  int trace_cap;
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
      for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += blockDim.x) ones[i] = 0x3F803F80u;
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
  pdl_wait();      // prologue above overlaps the previous kernel's tail (programmatic dependent launch)
  pdl_trigger();
  const int all_chunks = (p.offsets[e + 1] - p.offsets[e]) / BK;   // queues are padded to M3_PAD_ROWS (a multiple of BK)
  const int per_split = (all_chunks + nsplit - 1) / nsplit;
  const int c_begin = min(split * per_split, all_chunks), c_end = min(c_begin + per_split, all_chunks);
  const int r0 = p.offsets[e] + c_begin * BK;
  const int kchunks = c_end - c_begin;

  if (warp == 0) {
    // converged warp, elected lane issues; ring walked in rounds with a compile-time stage index (see gg_kernel)
    uint32_t phase = 0;
    const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(&full[0]);
    const bool no_tma = (p.dbg & 2) != 0;
    Tracer trc(p.trace, p.trace_cap, 0);
    for (int kc0 = 0; kc0 < kchunks; kc0 += STAGES) {
#pragma unroll
      for (int st = 0; st < STAGES; ++st) {
        if (kc0 + st < kchunks) {
          trc.ev(0x00, kc0 + st);
          mbar_wait(&empty[st], phase ^ 1);
          trc.ev(0x01, kc0 + st);
          __syncwarp();
          if (no_tma) {   // measurement only: no loads
            mbar_arrive_elect(&full[st]);
          } else {
            mbar_expect_tx_elect(&full[st], Cfg::TMA_BYTES);
            const uint32_t sa = smem_base + st * Cfg::STAGE, bar = full0 + st * 8;
            const int r = r0 + (kc0 + st) * BK;
#pragma unroll
            for (int b = 0; b < BM / 64; ++b) tma_load_2d_elect(sa + b * Cfg::BOX, &tm1, bar, m0 + b * 64, r);
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) tma_load_2d_elect(sa + Cfg::A_BYTES + b * Cfg::BOX, &tm2, bar, n0 + b * 64, r);
          }
        }
      }
      phase ^= 1;
    }
    trc.done();
  } else if (warp == 1) {
    Tracer trc(p.trace, p.trace_cap, 1);
    // whole warp, converged; the MMAs / commits are issued by one elected lane (see tc_common.cuh)
    const uint32_t idesc = with_db ? make_idesc_bf16(BM, BN + 16, 1, 1) : make_idesc_bf16(BM, BN, 1, 1);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    // MN-major: 16 k-rows = 2 swizzle groups of 8 rows (1024 B each, SBO); next 64 M/N elements one box further (LBO)
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), Cfg::BOX), hi = smem_desc_hi(1024);
    const bool no_mma = (p.dbg & 1) != 0;
    // The ring is walked in rounds of STAGES chunks with the stage index a compile-time constant, so that every
    // descriptor is base + immediate and every barrier address an immediate offset (a run-time stage index costs
    // ~25 SASS instructions per MMA in index arithmetic and vector-to-uniform register moves).
    uint32_t phase = 0;
    for (int kc0 = 0; kc0 < kchunks; kc0 += STAGES) {
#pragma unroll
      for (int st = 0; st < STAGES; ++st) {
        if (kc0 + st < kchunks) {
          trc.ev(0x10, kc0 + st);
          mbar_wait(&full[st], phase);
          trc.ev(0x12, kc0 + st);
          __syncwarp();
          tcgen05_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)st * (Cfg::STAGE >> 4), b_lo = a_lo + (Cfg::A_BYTES >> 4);
          if (!no_mma) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_elect(tm, a_lo + k * (2048 >> 4), hi, b_lo + k * (2048 >> 4), hi, idesc, (kc0 | st | k) != 0);
          }
          umma_commit_elect(&empty[st]);
          trc.ev(0x13, kc0 + st);
        }
      }
      phase ^= 1;
    }
    umma_commit_elect(tfull);
    trc.done();
  } else if (warp < 6) {
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

// L2 sector promotion of the TMA loads (M3_KNOB_DEBUG bits 12-13: 1 = none, 2 = 64 B, 3 = 128 B; default 256 B)
static CUtensorMapL2promotion l2_promotion() {
  switch ((g_knobs[M3_KNOB_DEBUG] >> 12) & 3) {
    case 1: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
    case 2: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    case 3: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    default: return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }
}

// 2-D bf16 row-major tensor [rows][cols], box [box_rows][64 cols] (128 B inner), SWIZZLE_128B
// (box_cols = 32: 64 B inner, SWIZZLE_64B - the 32-column epilogue blocks of the 16-warp epilogue)
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                    uint32_t box_cols = 64) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return M3_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M3_OK : M3_ERR_ARG;
}

// BN = 256 leaves too little smem for the two-output epilogues: those use 192 or 128.  The single-output epilogues take
// the widest tile that divides N (fewer L2 -> SM operand bytes per flop): at N = 768 (ViT-B), 256 against 192 measured
// 307 against 318 us for fc1 + fc2 and 622 against 632 us for the backward in one process.
static int pick_bn(int N, bool heavy_epilogue) {
  if (N % 256 == 0 && !heavy_epilogue) return 256;
  if (N % 192 == 0) return 192;
  return N % 128 == 0 ? 128 : 0;
}

// measured at T = 38 432, D = H = 384 (tools/variants.py): fc1 (bias + GELU + GELU', two outputs) gains ~8 % from 16
// warps, the other three epilogues are within noise of each other
constexpr int kDefaultEpiWarps[4] = {/*STORE*/ 8, /*BIAS*/ 8, /*FC1*/ 16, /*DGELU*/ 8};
// m3_debug_trace_buffer: M3_KNOB_TRACE_KERNEL = 1 + index of the GEMM launch (within one m3_ffn_fwd / m3_ffn_bwd call) that
// writes the timeline; 0 = every launch (each overwrites the previous one's events)
static int g_trace_launch_idx = 0;
static unsigned long long* trace_buf_for_this_launch() {
  const int want = g_knobs[M3_KNOB_TRACE_KERNEL], idx = g_trace_launch_idx++;
  return (want == 0 || want - 1 == idx) ? g_trace_buf : nullptr;
}
constexpr int kGGNcta = 2;   // CTA pairs: halves the per-SM weight traffic (the GEMMs are L2 -> SM bound at K = 384)

// SMs the persistent grouped GEMMs may occupy (m3_set_gemm_sm_limit): a caller that runs other kernels beside them on
// another stream (e.g. NVLink row movers, which need ~12-20 SMs) lowers it.
static int g_gemm_sms = kNumSMs;

// epilogue warps per GEMM epilogue: the knob (8 / 16) wins, otherwise the per-epilogue default
template <int EPI>
static int epi_warps() {
  const int k = g_knobs[M3_KNOB_EPI_WARPS];
  if (k & 0x100) return ((k >> EPI) & 1) ? 16 : 8;     // 0x100 | mask: bit EPI set -> 16 warps for that epilogue
  if ((k & 0xff) == 8 || (k & 0xff) == 16) return k & 0xff;
  return kDefaultEpiWarps[EPI];                        // (bit 0x200 selects the block width, see launch_gg)
}

template <int BN, int EPI, int EW, int CWP>
static int launch_gg_t(const CUtensorMap* maps, const GGParams& p, int max_tiles, cudaStream_t st) {
  using Cfg = GGCfg<BN, EPI, kGGNcta, EW, CWP>;
  auto kern = gg_kernel<BN, EPI, kGGNcta, EW, CWP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return (int)e;
  int grid = max_tiles < g_gemm_sms ? max_tiles : g_gemm_sms;
  grid = grid / kGGNcta * kGGNcta;
  if (grid < kGGNcta) grid = kGGNcta;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(gg_threads(EW));
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kGGNcta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_knobs[M3_KNOB_PDL] ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], p);
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
  CUtensorMap maps[3];      // A, B, aux (saved gelu' of the dgelu epilogue); outputs use plain stores
  int rc = make_map(&maps[0], A, (uint64_t)cap_rows, (uint64_t)p.Kd, BM);
  if (rc) return rc;
  rc = make_map(&maps[1], B, (uint64_t)p.E * p.N, (uint64_t)p.Kd, (uint32_t)(BN / kGGNcta));
  if (rc) return rc;
  const int ew = epi_warps<EPI>();
  // epilogue register-block / staging-box width of the 8-warp epilogues: 32 columns (2 KB boxes) leave room for a 7th
  // smem stage, but measured 10 us slower per fc1+fc2 than 64 columns / 6 stages (tools/variants.py): opt-in (0x200)
  const bool narrow = ew == 16 || (g_knobs[M3_KNOB_EPI_WARPS] & 0x200) != 0;
  rc = make_map(&maps[2], aux ? aux : out, (uint64_t)cap_rows, (uint64_t)p.N, 32, narrow ? 32 : 64);   // per-warp [32 x CW] boxes
  if (rc) return rc;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out2 = static_cast<__nv_bfloat16*>(out2);
  p.dbg = g_knobs[M3_KNOB_DEBUG];
  p.trace = trace_buf_for_this_launch();
  p.trace_cap = g_trace_cap;
  const int max_tiles = (cap_rows / BM) * (p.N / BN);
#define M3_GG_EW(BNV)                                                                   \
  (ew == 16 ? launch_gg_t<BNV, EPI, 16, 32>(maps, p, max_tiles, st)                      \
            : narrow ? launch_gg_t<BNV, EPI, 8, 32>(maps, p, max_tiles, st)              \
                     : launch_gg_t<BNV, EPI, 8, 64>(maps, p, max_tiles, st))
  switch (BN) {
    case 128: return M3_GG_EW(128);
    case 192: return M3_GG_EW(192);
    default:
      if constexpr (!heavy) return M3_GG_EW(256);
      return M3_ERR_SHAPE;
  }
#undef M3_GG_EW
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
  const int S = wgrad_splits(E, M, N);
  float* pW = S > 1 ? ws : dW;
  float* pb = S > 1 ? ws + (size_t)S * E * M * N : db;
  WGParams p{offsets, M, N, pW, pb, E, g_knobs[M3_KNOB_DEBUG], trace_buf_for_this_launch(), g_trace_cap};
  auto kern = wgrad_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return (int)e;
  launch_k(kern, dim3(M / BM, N / BN, E * S), dim3(kThreads), Cfg::SMEM, st, t1, t2, p);
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

// ffn_chain.cu
int m3_ffn_chain_supported(int D, int H);
int m3_ffn_chain_fwd(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                     int H, const void* w1, const float* b1, const void* w2, const float* b2, void* yq, int max_ctas,
                     cudaStream_t st);
// The single-kernel fc1 -> GELU -> fc2 chain (ffn_chain.cu) runs the forwards that keep NO state (inference), where its
// [128 x D] output accumulator fits TMEM (D = 128 / 256 / 384) and H <= 2 D: 2 queue-sized planes of DRAM traffic instead
// of 4, 121 vs 125 us at the bench shape (at H = 4 D the two GEMMs win: 395 vs 403 us).  M3_KNOB_FFN_CHAIN = 0 switches it
// off.  Training keeps the two grouped GEMMs: a chain forward that saves z (one plane instead of two) measured 128 vs
// 143 us, but the backward then has to rebuild h = gelu(z) AND gelu'(z) - in the first GEMM's epilogue 258 vs 230 us, as a
// backward chain kernel 285 us - so every variant lost over forward + backward and none is kept.
static bool chain_inference(int D, int H) {
  return g_knobs[M3_KNOB_FFN_CHAIN] != 0 && m3_ffn_chain_supported(D, H) && H <= 2 * D;
}
int m3_ffn_bf16_chain_mode(int D, int H) { return chain_inference(D, H) ? 1 : 0; }

// Opaque activation state handed from m3_ffn_fwd to m3_ffn_bwd (bf16): two [cap][H] planes,
//   plane 0 = gelu'(z),  plane 1 = h = gelu(z)
size_t m3_ffn_bf16_saved_bytes(int cap_rows, int D, int H) { (void)D; return 2 * align256((size_t)cap_rows * H * 2); }

// workspace: forward  : h [cap][H] bf16 (two-kernel inference only; training keeps h in the saved state)
//            backward : dz [cap][H] bf16 | (unused [cap][H]) | wgrad split-K partials fp32
size_t m3_ffn_bf16_workspace_bytes(int cap_rows, int D, int H, int E, int backward) {
  const size_t hbytes = align256((size_t)cap_rows * H * 2);
  if (!backward) return hbytes;
  const size_t wg = wgrad_ws_bytes(E, D, H) > wgrad_ws_bytes(E, H, D) ? wgrad_ws_bytes(E, D, H) : wgrad_ws_bytes(E, H, D);
  return 2 * hbytes + align256(wg);
}

int m3_ffn_fwd_bf16(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                    int H, const void* w1, const float* b1, const void* w2, const float* b2, void* saved, void* yq,
                    void* workspace, size_t workspace_bytes, float drop_p, const void* rng, const int32_t* ret_meta,
                    void* const* ret_bases, cudaStream_t st) {
  if (drop_p > 0.f && (saved == nullptr || rng == nullptr)) return M3_ERR_ARG;     // dropout is a training-time op
  if ((ret_meta == nullptr) != (ret_bases == nullptr) || (ret_meta == nullptr && yq == nullptr)) return M3_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < m3_ffn_bf16_workspace_bytes(cap_rows, D, H, E, 0)) return M3_ERR_WORKSPACE;
  g_trace_launch_idx = 0;
  if (saved == nullptr && ret_meta == nullptr && chain_inference(D, H))
    return m3_ffn_chain_fwd(xq, offsets, tile_expert, cap_rows, E, D, H, w1, b1, w2, b2, yq, g_gemm_sms, st);
  const size_t hbytes = align256((size_t)cap_rows * H * 2);
  bf16* gp = static_cast<bf16*>(saved);
  bf16* h = saved ? reinterpret_cast<bf16*>(static_cast<uint8_t*>(saved) + hbytes) : static_cast<bf16*>(workspace);
  GGParams p{};
  p.offsets = offsets; p.tile_expert = tile_expert; p.E = E;
  // fc1: h = gelu(xq W1^T + b1); gelu'(.) saved for backward
  p.N = H; p.Kd = D; p.bias = b1; p.save_out2 = saved != nullptr;
  if (drop_p > 0.f) {
    p.drop_thr = dropout_threshold(drop_p);
    p.drop_inv_keep = 1.0f / (1.0f - drop_p);
    p.rng = static_cast<const RngState*>(rng);
  }
  int rc = launch_gg<EPI_FC1>(xq, w1, h, gp, nullptr, p, cap_rows, st);
  if (rc) return rc;
  // fc2: yq = h W2^T + b2
  p.N = D; p.Kd = H; p.bias = b2; p.save_out2 = 0; p.drop_thr = 0u;
  p.ret_meta = ret_meta; p.ret_bases = reinterpret_cast<bf16* const*>(ret_bases);      // expert parallel: rows go home
  return launch_gg<EPI_BIAS>(h, w2, ret_meta ? static_cast<void*>(h) : yq, nullptr, nullptr, p, cap_rows, st);
}

int m3_ffn_bwd_bf16(const void* xq, const void* saved, const void* dyq, const int32_t* counts, const int32_t* offsets,
                    const int32_t* tile_expert, int cap_rows, int E, int D, int H, const void* w1, const void* w2,
                    const void* w1t, const void* w2t, void* dxq, float* dw1, float* db1, float* dw2, float* db2,
                    void* workspace, size_t workspace_bytes, int parts, const int32_t* ret_meta, void* const* ret_bases,
                    cudaStream_t st) {
  (void)counts; (void)w1; (void)w2;
  if ((ret_meta == nullptr) != (ret_bases == nullptr) || (ret_meta == nullptr && dxq == nullptr)) return M3_ERR_ARG;
  if (workspace_bytes < m3_ffn_bf16_workspace_bytes(cap_rows, D, H, E, 1)) return M3_ERR_WORKSPACE;
  g_trace_launch_idx = 0;
  const size_t hbytes = align256((size_t)cap_rows * H * 2);
  bf16* dz = static_cast<bf16*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 2 * hbytes);
  const bf16* h;
  int rc;
  h = reinterpret_cast<const bf16*>(static_cast<const uint8_t*>(saved) + hbytes);
  if (parts & 1) {      // data gradients: dz (kept in the workspace for the weight-gradient part) and dxq
    GGParams p{};
    p.offsets = offsets; p.tile_expert = tile_expert; p.E = E;
    // dz = (dyq W2) * gelu'(z)                              B = W2^T [E][H][D] (K-major in D)
    p.N = H; p.Kd = D;
    rc = launch_gg<EPI_DGELU>(dyq, w2t, dz, nullptr, saved, p, cap_rows, st);
    if (rc) return rc;
    // dxq = dz W1                                           B = W1^T [E][D][H] (K-major in H)
    p.N = D; p.Kd = H;
    p.ret_meta = ret_meta; p.ret_bases = reinterpret_cast<bf16* const*>(ret_bases);    // expert parallel: rows go home
    rc = launch_gg<EPI_STORE>(dz, w1t, ret_meta ? static_cast<void*>(dz) : dxq, nullptr, nullptr, p, cap_rows, st);
    if (rc) return rc;
  }
  if (!(parts & 2)) return M3_OK;
  // dW2[e] = dyq_e^T h_e  [D][H];   dW1[e] = dz_e^T xq_e  [H][D]
  // the bias gradients ride along in the same MMA against a tile of ones
  rc = launch_wgrad(dyq, h, offsets, cap_rows, E, D, H, dw2, db2, part, st);
  if (rc) return rc;
  return launch_wgrad(dz, xq, offsets, cap_rows, E, H, D, dw1, db1, part, st);
}

int m3_ffn_bf16_set_sm_limit(int sms) {
  if (sms < 2 || sms > kNumSMs) sms = kNumSMs;
  g_gemm_sms = sms;
  return M3_OK;
}

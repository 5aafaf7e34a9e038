// Expert FFN chain on CTA pairs:  y = GELU(x W1^T + b1) W2^T + b2  in ONE kernel (sm_100a) - the hidden
// activation never leaves the SM.  The inference forward of the layer (no state kept for a backward pass).
//
// Replaces _Expert.forward of the reference (FMoELinear -> GELU -> FMoELinear,
// /root/reference/models/moe/origin/custom_moe_layer.py:36-44).  The two-kernel path (ffn_bf16.cu) moves 4 queue-sized
// planes through HBM per inference forward (x read, h written, h read, y written); this kernel moves 2.
//
// A CTA pair (cluster of 2, tcgen05 cta_group::2) owns a 256-row tile of ONE expert's queue (queues are padded to
// 256 rows); each CTA keeps its own [128 x D] slice of x resident in shared memory and walks the hidden dimension in
// chunks of 64:
//     G1(c):  acc1[128 x 64]  = x       * W1[e][64c .. 64c+64, :]^T     (TMEM, double-buffered: columns 384 / 448)
//     mid(c): h = gelu(acc1 + b1) -> bf16 -> h chunk in smem (K-major, SWIZZLE_128B, written by the epilogue warps,
//             fence.proxy.async, mbarrier arrive)
//     G2(c):  acc2[128 x D]  += hchunk  * W2[e][:, 64c .. 64c+64]^T     (TMEM columns 0 .. D-1)
// and finally  y = acc2 + b2.  The MMA warp issues G1(c+2) BEFORE G2(c), so mid(c) runs under G1(c+1) and G1(c+2).
// Weight tiles stream through a ring of TMA slots; the pair splits every weight tile (each CTA loads half).
// TMEM: 384 + 2 * 64 = 512 columns.
//
// Warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA) + TMEM owner, warps 2-17 = epilogue (two sets of 8).  Producer and MMA
// warps run CONVERGED with every TMA / MMA / commit predicated on elect.sync inside its asm block (tc_common.cuh:
// the `if (lane == 0)` form of the first version of this kernel cost ~175 clk of scalar code per MMA - 192 MMAs per
// tile - and was the whole reason it lost against the two-kernel path).  The schedule is FLAT over all chunks of all
// tiles of the pair: step f runs G1(chunk f) and then G2(chunk f - 2).  The step loop is NOT unrolled and every
// mbarrier wait keeps its time-out path out of line: the unrolled form (6 steps, ~90 KB of SASS in the MMA warp alone)
// missed the instruction cache at every branch target and spent ~400 clk per k-box on instruction fetch.
//
// Training does not use it.  Variants that did were built and measured at the bench shape (T = 38 432, D = H = 384):
// a forward that also stores z as the saved state (one plane instead of gelu'(z) and h) ran in 128 us against 143 us for
// the two GEMMs, but the backward then has to rebuild h AND gelu'(z): as a backward chain kernel (mid-epilogue with
// twice the math and twice the stores) 180 us against 2 x 59 us for the GEMM pair, in the first backward GEMM's epilogue
// +28 us.  Every combination lost over forward + backward; the per-chunk budget explains why: 1728 clk of MMA, but
// ~1500 clk of FP32 pipe for one GELU over [128 x 64] and two barrier hops of latency around it.
#include <cstdio>

#include "tc_common.cuh"

namespace m3 {
namespace tc {

constexpr int CBM = 128, CBK = 64, CUK = 16, CHC = 64;   // rows per CTA, k-box, UMMA K, hidden chunk
constexpr int kChainEpiWarps = 16;    // two sets of 8 (one warp per TMEM lane quarter and column half), alternating chunks
constexpr int kChainThreads = 64 + kChainEpiWarps * 32;
constexpr int CXBOX = CBM * CBK * 2;                      // 16 KB: [128 rows][64 bf16]

struct ChainParams {
  const int32_t* offsets;      // [E+1]
  const int32_t* tile_expert;  // per 256-row tile
  int E, H;
  const float* b1;             // [E][H]
  const float* b2;             // [E][D]
  __nv_bfloat16* out;          // y [rows][D]
  int dbg;                     // M3_KNOB_DEBUG (measurement only, results are garbage): see m3vit_moe.h
  unsigned long long* trace;   // m3_debug_trace_buffer (compiled in with M3_GEMM_TRACE only)
  int trace_cap;
};

template <int D>
struct ChainCfg {
  static constexpr int KD = D / CBK;                        // k-boxes of the resident A tile
  static constexpr int NPART = (D > 256) ? 2 : 1;           // G2 output split into <= 256-column MMAs
  static constexpr int NP = D / NPART;                      // columns per G2 MMA (pair-wide N)
  static constexpr int W2_PART = (NP / 2) * CBK * 2;        // this CTA's half of a [NP x 64] W2 part
  static constexpr int W1_BOX = (CHC / 2) * CBK * 2;        // this CTA's half of a [64 x 64] W1 k-box = 4 KB
  static constexpr int SLOT = KD * W1_BOX;                  // one ring slot = one weight chunk: W1 [32 x D] or W2 [D/2 x 64]
  static_assert(SLOT == NPART * W2_PART, "W1 and W2 chunks are the same size");
  static constexpr int X_BYTES = KD * CXBOX;
  static constexpr int H_BYTES = 2 * CXBOX;                 // two h chunks [128 x 64]
  static constexpr int B1_RING = 8;                         // bias chunks [64] fp32 staged per hidden chunk (FWD)
  static constexpr int B1_BYTES = B1_RING * CHC * 4;
  static constexpr int B2_BYTES = 2 * D * 4;                // b2[e] of the current and the next tile (FWD)
  static constexpr int FIXED = X_BYTES + H_BYTES + B1_BYTES + B2_BYTES;
  static constexpr int NSLOT_RAW = (227 * 1024 - 1024 - 1024 - FIXED) / SLOT;
  static constexpr int NSLOT = NSLOT_RAW > 8 ? 8 : NSLOT_RAW;
  static constexpr int SMEM = FIXED + NSLOT * SLOT + 1024 + 1024;
  static_assert(D % 64 == 0 && D <= 384, "the chain kernel keeps the [128 x D] output accumulator in TMEM");
  static_assert(NP % 32 == 0 && NP <= 256, "G2 MMA width");
  static_assert(NSLOT >= 3, "weight ring too shallow");
  static_assert((2 * 8 + 2 * 6 + 10 + 2 * 8 + 4) * 8 + 4 <= 1024, "barrier block");
};

// 16-byte chunk c (0..7) of row r inside a [rows][64 bf16] swizzle-128B box
__device__ __forceinline__ uint32_t cbox_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

__device__ __forceinline__ uint32_t cpk_bf16x2(f32x2 v) {
  float a, b;
  unpk2(v, a, b);
  return float2_to_bf16x2(a, b);
}
__device__ __forceinline__ f32x2 cunpk_bf16x2(uint32_t u) {
  const float2 f = bf16x2_to_float2(u);
  return pk2(f.x, f.y);
}

template <int D>
__global__ void __launch_bounds__(kChainThreads, 1)
ffn_chain_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, ChainParams p) {
  using Cfg = ChainCfg<D>;
  constexpr int KD = Cfg::KD, NPART = Cfg::NPART, NP = Cfg::NP, NSLOT = Cfg::NSLOT;
    const uint32_t cta_rank = cluster_ctarank();
  const bool leader_cta = cta_rank == 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                   // [KD][128][64] resident A tile
  uint8_t* hs = xs + Cfg::X_BYTES;                      // [2][128][64]   h chunks (A operand of G2)
  uint8_t* ring = hs + Cfg::H_BYTES;                    // [NSLOT][SLOT]  weight ring
  uint8_t* b1s = ring + NSLOT * Cfg::SLOT;              // [8][64] fp32: b1 of the hidden chunks in flight
  uint8_t* b2s = b1s + Cfg::B1_BYTES;                   // [2][D] fp32:  b2 of this / the next tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(b2s + Cfg::B2_BYTES);
  uint64_t* w_full = bars;                  // [8] one per ring slot; leader-waited, count 2 (+ tx bytes of both CTAs)
  uint64_t* w_empty = w_full + 8;           // [8] count 1 (multicast commit)
  uint64_t* x_full = w_empty + 8;           // [6] per k-box of the resident tile; leader-waited, count 2
  uint64_t* x_empty = x_full + 6;           // [6] count 1 (multicast commit after the tile's LAST G1 has read the box)
  uint64_t* a1_full = x_empty + 6;          // [2] count 1 (multicast commit)
  uint64_t* a1_empty = a1_full + 2;         // [2] leader-waited, count 2 CTAs * 8 warps (the set that owns the buffer)
  uint64_t* h_full = a1_empty + 2;          // [2] leader-waited, count 2 CTAs * 8 warps
  uint64_t* h_empty = h_full + 2;           // [2] count 1 (multicast commit)
  uint64_t* a2_full = h_empty + 2;          // count 1 (multicast commit)
  uint64_t* a2_empty = a2_full + 1;         // leader-waited, count 2 CTAs * 16 warps
  uint64_t* b1_full = a2_empty + 1;         // [8] this CTA's own bulk copy of a b1 chunk (count 1 + tx)
  uint64_t* b1_empty = b1_full + 8;         // [8] count 8: the epilogue set that used it (this CTA)
  uint64_t* b2_full = b1_empty + 8;         // [2] by tile parity (count 1 + tx)
  uint64_t* b2_empty = b2_full + 2;         // [2] count 16: every epilogue warp of this CTA after the tile's final epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b2_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 2); mbar_init(&w_empty[s], 1); }
      for (int k = 0; k < KD; ++k) { mbar_init(&x_full[k], 2); mbar_init(&x_empty[k], 1); }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&a1_full[b], 1); mbar_init(&a1_empty[b], kChainEpiWarps);     // one set of 8 warps in each CTA
        mbar_init(&h_full[b], kChainEpiWarps); mbar_init(&h_empty[b], 1);
      }
      mbar_init(a2_full, 1); mbar_init(a2_empty, 2 * kChainEpiWarps);         // all 16 warps of both CTAs
      for (int i = 0; i < 8; ++i) { mbar_init(&b1_full[i], 1); mbar_init(&b1_empty[i], kChainEpiWarps / 2); }
      for (int i = 0; i < 2; ++i) { mbar_init(&b2_full[i], 1); mbar_init(&b2_empty[i], kChainEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm<512>(tmem_slot);
  }
  tcgen05_fence_before();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t acc2_col = 0, acc1_col = 384;   // acc1 buffers at columns 384 and 448
  pdl_wait();       // everything above is private to the CTA pair: under PDL it runs beneath the previous kernel's tail
  pdl_trigger();

  const int NH = p.H / CHC;
  const int total = p.offsets[p.E] / (2 * CBM);         // 256-row pair tiles in use
  const int unit = blockIdx.x / 2, n_units = gridDim.x / 2;
  const int my_tiles = unit < total ? (total - unit + n_units - 1) / n_units : 0;
  // flat schedule: chunk c = ti * NH + j (local tile ti, hidden chunk j); step f = 0 .. C + LAG - 1 runs G1(chunk f)
  // [f < C] and then G2(chunk f - LAG) [f >= LAG], taking its W1 slots and then its W2 slots from the ring.  LAG = 2
  // (both acc1 buffers in flight): the mid-epilogue of a chunk - barrier hops, TMEM read, the erf / rcp chains, the h
  // store - is ~3000 clk of LATENCY and has two whole steps to finish before the MMA warp asks for its h chunk; with
  // LAG = 1 that latency, not the tensor pipe, set the pace (162 vs 100 us with the epilogue switched off).
  constexpr int LAG = 2;
  const int C = my_tiles * NH;

  if (warp == 0) {
    // ===================================================== TMA producer (converged warp, elected lane issues)
    if (C > 0) {
      const uint32_t xs_a = smem_u32(xs), ring_a = smem_u32(ring);
      const uint32_t wfull_r = mapa_u32(smem_u32(&w_full[0]), 0), xfull_r = mapa_u32(smem_u32(&x_full[0]), 0);
      int j1 = 0, ti1 = 0, j2 = 0, ti2 = 0;     // (chunk-in-tile, local tile) of the G1 / G2 chunk of the step
      int slot = 0;                             // ring position of the next slot use, and the parity of its wrap count
      uint32_t wpar = 0;
      Tracer trc(p.trace, p.trace_cap, 0);
      const bool no_tma = (p.dbg & 2) != 0;     // measurement only: no loads, the MMAs chew on whatever smem holds
      // All pairs start together and run the same schedule, so without help they all store their y tile (and load the
      // next x tile) in the same ~2 us window: a 14 MB burst every tile period that the memory system takes 2.3 us to
      // absorb while every tensor pipe waits.  Pairs that have a tile fewer than the busiest ones start up to one tile
      // period late, in four phases, which spreads the bursts at no cost to the critical path.
      const int max_tiles = (total + n_units - 1) / n_units;
      if (my_tiles < max_tiles && !(p.dbg & 16)) {
        const long long delay = (long long)((unit * 7) % 4 + 1) * NH * 480;
        const long long t0 = clock64();
        while (clock64() - t0 < delay) {}
        __syncwarp();
      }
#pragma unroll 1
      for (int f = 0; f < C + LAG; ++f) {
        if (f < C) {
          const int pt = unit + ti1 * n_units;
          const int e1 = p.tile_expert[pt];
          if (j1 == 0) {
            // resident A tile of the next tile, box by box as the previous tile's last G1 lets go of them
            const int row0 = (pt * 2 + (int)cta_rank) * CBM;
            const uint32_t xpar = ((uint32_t)ti1 & 1u) ^ 1u;
            trc.ev(0x02, f);
            {      // b2 of the tile's expert -> smem (own copy in each CTA)
              const uint32_t tb = (uint32_t)ti1 & 1u;
              mbar_wait(&b2_empty[tb], (((uint32_t)ti1 >> 1) & 1u) ^ 1u);
              __syncwarp();
              mbar_expect_tx_elect(&b2_full[tb], D * 4);
              bulk_load_1d_elect(smem_u32(b2s) + tb * (D * 4), p.b2 + (int64_t)e1 * D, D * 4, &b2_full[tb]);
            }
#pragma unroll
            for (int kc = 0; kc < KD; ++kc) {
              mbar_wait(&x_empty[kc], xpar);
              __syncwarp();
              if (no_tma) {
                if (leader_cta) mbar_arrive_elect(&x_full[kc]); else mbar_arrive_remote_elect(xfull_r + kc * 8);
              } else {
                if (leader_cta) mbar_expect_tx_elect(&x_full[kc], 2 * CXBOX); else mbar_arrive_remote_elect(xfull_r + kc * 8);
                tma_load_2d_2sm_elect(xs_a + kc * CXBOX, &tmX, xfull_r + kc * 8, kc * CBK, row0);
              }
            }
          }
          {        // b1 of the chunk -> smem ring entry f & 7 (own copy in each CTA)
            const uint32_t be = (uint32_t)f & 7u;
            mbar_wait(&b1_empty[be], (((uint32_t)f >> 3) & 1u) ^ 1u);
            __syncwarp();
            mbar_expect_tx_elect(&b1_full[be], CHC * 4);
            bulk_load_1d_elect(smem_u32(b1s) + be * (CHC * 4), p.b1 + (int64_t)e1 * p.H + j1 * CHC, CHC * 4, &b1_full[be]);
          }
          // W1 chunk j1: this CTA's 32 of the 64 hidden rows, all KD k-boxes, ONE ring slot
          const int wrow = e1 * p.H + j1 * CHC + (int)cta_rank * (CHC / 2);
          trc.ev(0x00, f);
          mbar_wait(&w_empty[slot], wpar ^ 1u);
          trc.ev(0x01, f);
          __syncwarp();
          {
            const uint32_t bar = wfull_r + slot * 8, dst = ring_a + slot * Cfg::SLOT;
            if (no_tma) {
              if (leader_cta) mbar_arrive_elect(&w_full[slot]); else mbar_arrive_remote_elect(bar);
            } else {
              if (leader_cta) mbar_expect_tx_elect(&w_full[slot], 2 * Cfg::SLOT); else mbar_arrive_remote_elect(bar);
#pragma unroll
              for (int kc = 0; kc < KD; ++kc) tma_load_2d_2sm_elect(dst + kc * Cfg::W1_BOX, &tmW1, bar, kc * CBK, wrow);
            }
          }
          if (++slot == NSLOT) { slot = 0; wpar ^= 1u; }
          if (++j1 == NH) { j1 = 0; ++ti1; }
        }
        if (f >= LAG) {
          // W2 chunk j2: this CTA's half of every NP-row part, k = the hidden chunk, ONE ring slot
          const int e2 = p.tile_expert[unit + ti2 * n_units];
          trc.ev(0x04, f);
          mbar_wait(&w_empty[slot], wpar ^ 1u);
          trc.ev(0x05, f);
          __syncwarp();
          {
            const uint32_t bar = wfull_r + slot * 8, dst = ring_a + slot * Cfg::SLOT;
            if (no_tma) {
              if (leader_cta) mbar_arrive_elect(&w_full[slot]); else mbar_arrive_remote_elect(bar);
            } else {
              if (leader_cta) mbar_expect_tx_elect(&w_full[slot], 2 * Cfg::SLOT); else mbar_arrive_remote_elect(bar);
#pragma unroll
              for (int part = 0; part < NPART; ++part)
                tma_load_2d_2sm_elect(dst + part * Cfg::W2_PART, &tmW2, bar, j2 * CHC, e2 * D + part * NP + (int)cta_rank * (NP / 2));
            }
          }
          if (++slot == NSLOT) { slot = 0; wpar ^= 1u; }
          if (++j2 == NH) { j2 = 0; ++ti2; }
        }
      }
      trc.done();
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA; converged warp, elected lane issues)
    if (leader_cta && C > 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(2 * CBM, CHC, 0, 0);   // 256 x 64
      constexpr uint32_t idesc2 = make_idesc_bf16(2 * CBM, NP, 0, 0);    // 256 x NP
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      // K-major operand tiles [rows][64 bf16], SWIZZLE_128B: 8-row groups 1024 B apart (SBO); K advances 32 B per UMMA_K
      const uint32_t hi = smem_desc_hi(1024);
      const uint32_t xs_lo0 = smem_desc_lo(smem_u32(xs), 0), hs_lo0 = smem_desc_lo(smem_u32(hs), 0);
      const uint32_t ring_lo0 = smem_desc_lo(smem_u32(ring), 0);
      int j1 = 0, ti1 = 0, j2 = 0, ti2 = 0;
      int slot = 0;
      uint32_t wpar = 0;
      Tracer trc(p.trace, p.trace_cap, 1);
      const bool no_mma = (p.dbg & 1) != 0;     // measurement only
      // The issuing warp does little besides waiting: a try_wait on an already completed barrier still costs ~90 clk and
      // they serialise, so every phase waits for ALL its barriers with the try_waits in flight together, and a weight
      // chunk is ONE ring slot (one wait, one commit) - with 4 slots, 13 waits and 9 commits per step the bare
      // synchronisation skeleton of this loop took 2250 clk per step against 1728 clk of MMAs.
#pragma unroll 1
      for (int f = 0; f < C + LAG; ++f) {
        if (f < C) {                                     // ---- G1(chunk f): acc1[f & 1] = x * W1_j^T
          const uint32_t b = (uint32_t)f & 1u;
          const uint32_t cpar = ((uint32_t)f >> 1) & 1u; // use parity of acc1[b]
          trc.ev(0x10, f);
          mbar_wait2(&a1_empty[b], cpar ^ 1u, &w_full[slot], wpar);
          trc.ev(0x12, f);
          if (j1 == 0) {                                 // first chunk of a tile: its x boxes
            const uint32_t xp = (uint32_t)ti1 & 1u;
            bool ok[KD];
#pragma unroll
            for (int kc = 0; kc < KD; ++kc) ok[kc] = mbar_try_wait(&x_full[kc], xp);
#pragma unroll
            for (int kc = 0; kc < KD; ++kc) if (!ok[kc]) mbar_wait_slow(smem_u32(&x_full[kc]), xp);
          }
          __syncwarp();
          tcgen05_fence_after();
          const uint32_t d1 = tm + acc1_col + b * CHC;
          const uint32_t b_lo0 = ring_lo0 + slot * (Cfg::SLOT >> 4);
          if (!no_mma && j1 != NH - 1) {
            // all KD k-boxes (4 * KD MMAs) under one election
            umma_bf16_2sm_elect_boxes<KD>(d1, xs_lo0, b_lo0, hi, idesc1, 0, CXBOX >> 4, Cfg::W1_BOX >> 4);
          } else if (!no_mma) {
            // the tile's last G1 lets go of the resident tile box by box: the next tile's x streams in behind it
#pragma unroll
            for (int kc = 0; kc < KD; ++kc) {
              umma_bf16_2sm_elect_x4(d1, xs_lo0 + kc * (CXBOX >> 4), b_lo0 + kc * (Cfg::W1_BOX >> 4), hi, idesc1, kc != 0);
              umma_commit_2sm_elect(&x_empty[kc], 3);
            }
          } else if (j1 == NH - 1) {
#pragma unroll
            for (int kc = 0; kc < KD; ++kc) umma_commit_2sm_elect(&x_empty[kc], 3);
          }
          umma_commit_2sm_elect(&w_empty[slot], 3);
          umma_commit_2sm_elect(&a1_full[b], 3);
          if (++slot == NSLOT) { slot = 0; wpar ^= 1u; }
          trc.ev(0x13, f);
          if (++j1 == NH) { j1 = 0; ++ti1; }
        }
        if (f >= LAG) {                                  // ---- G2(chunk f - LAG): acc2 += h * W2_j^T
          const uint32_t b = (uint32_t)(f - LAG) & 1u;
          const uint32_t cpar = ((uint32_t)(f - LAG) >> 1) & 1u;
          trc.ev(0x14, f);
          if (j2 == 0) mbar_wait(a2_empty, ((uint32_t)ti2 & 1u) ^ 1u);   // drained by the previous tile's epilogue
          mbar_wait2(&h_full[b], cpar, &w_full[slot], wpar);             // h chunk written by the epilogue warps of BOTH CTAs
          trc.ev(0x15, f);
          __syncwarp();
          tcgen05_fence_after();
          const uint32_t a_lo = hs_lo0 + b * (CXBOX >> 4);
          const uint32_t b_lo0 = ring_lo0 + slot * (Cfg::SLOT >> 4);
          if (!no_mma) {
#pragma unroll
            for (int part = 0; part < NPART; ++part)
              umma_bf16_2sm_elect_x4(tm + acc2_col + part * NP, a_lo, b_lo0 + part * (Cfg::W2_PART >> 4), hi, idesc2, j2 != 0);
          }
          umma_commit_2sm_elect(&w_empty[slot], 3);
          umma_commit_2sm_elect(&h_empty[b], 3);
          if (j2 == NH - 1) umma_commit_2sm_elect(a2_full, 3);
          if (++slot == NSLOT) { slot = 0; wpar ^= 1u; }
          trc.ev(0x17, f);
          if (++j2 == NH) { j2 = 0; ++ti2; }
        }
      }
      trc.done();
    }
  } else {
    // ===================================================== epilogue warps (both CTAs)
    // 16 warps in two sets: set s = ew / 8 owns the chunks c with c & 1 == s, i.e. accumulator acc1[s] and h buffer s, so
    // two mid-epilogues are always in flight (the dependent erf / rcp chains need the warps, not the issue slots); all
    // 16 share the final epilogue of a tile.  Every lane owns one row and 32 consecutive columns = 64 contiguous bytes of
    // each output row, which it stores with two 256-bit st.global (full 32-byte sectors, no staging through smem).
    const int q = warp & 3;                      // TMEM lane quarter
    const int ew = warp - 2;                     // 0 .. 15
    const uint32_t set = (uint32_t)ew >> 3;      // which chunk parity / acc1 buffer / h buffer
    const int half = (ew >> 2) & 1;              // which 32 of a chunk's 64 columns
    const int grp = ew >> 2;                     // final epilogue: which quarter of the D output columns
    const uint32_t hs_a = smem_u32(hs);
    const uint32_t a1e_remote = mapa_u32(smem_u32(&a1_empty[0]), 0);
    const uint32_t hf_remote = mapa_u32(smem_u32(&h_full[0]), 0);
    const uint32_t a2e_remote = mapa_u32(smem_u32(a2_empty), 0);
    auto arrive_leader = [&](uint64_t* local, uint32_t remote) {     // one arrival per warp, on the LEADER's barrier
      __syncwarp();
      if (lane == 0) { if (leader_cta) mbar_arrive(local); else mbar_arrive_remote(remote); }
    };
    auto store_row32 = [&](__nv_bfloat16* dst, const uint32_t* w) {   // 32 bf16 = 64 bytes of this lane's row
      U8 lo, hi8;
#pragma unroll
      for (int i = 0; i < 8; ++i) { lo.v[i] = w[i]; hi8.v[i] = w[8 + i]; }
      stg_stream256(dst, lo);
      stg_stream256(dst + 16, hi8);
    };
    Tracer trc(ew == 0 ? p.trace : nullptr, p.trace_cap, 2);
    const uint32_t b = set;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int pt = unit + ti * n_units;
      const int64_t grow = (int64_t)(pt * 2 + (int)cta_rank) * CBM + q * 32 + lane;   // this lane's global row
      const int r = q * 32 + lane;                                                   // row inside the CTA tile
#pragma unroll 1
      for (int j = 0; j < NH; ++j) {
        const uint32_t c = (uint32_t)(ti * NH + j);                  // flat chunk index
        if ((c & 1u) != set) continue;
        const uint32_t par = (c >> 1) & 1u;
        const int hcol = j * CHC + half * 32;
        trc.ev(0x20, c);
        mbar_wait(&a1_full[b], par);
        trc.ev(0x21, c);
        tcgen05_fence_after();
        float v[32];
        tmem_ld_32x32(tmem_base + acc1_col + b * CHC + half * 32 + ((uint32_t)(q * 32) << 16), v);
        tcgen05_fence_before();
        arrive_leader(&a1_empty[b], a1e_remote + b * 8);             // acc1[b] may be overwritten by G1(c + 2)
        trc.ev(0x22, c);
        if (p.dbg & 4) {   // measurement only: no epilogue math / stores, the (unwritten) h chunk is handed over at once
          mbar_wait(&b1_full[c & 7u], (c >> 3) & 1u);
          if (lane == 0) mbar_arrive(&b1_empty[c & 7u]);
          mbar_wait(&h_empty[b], par ^ 1u);
          arrive_leader(&h_full[b], hf_remote + b * 8);
          continue;
        }
        f32x2 w2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w2[i] = pk2(v[2 * i], v[2 * i + 1]);
        uint32_t hq[16];                                             // what goes into the h chunk (A of G2)
        {
          // + b1: the chunk's 64 bias values sit in smem ring entry c & 7 (bulk copy issued with the W1 chunk)
          const uint32_t be = c & 7u;
          mbar_wait(&b1_full[be], (c >> 3) & 1u);
          const uint32_t ba = smem_u32(b1s) + be * (CHC * 4) + half * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 u = lds128(ba + i * 16);
            w2[2 * i] = add2(w2[2 * i], pk2(__uint_as_float(u.x), __uint_as_float(u.y)));
            w2[2 * i + 1] = add2(w2[2 * i + 1], pk2(__uint_as_float(u.z), __uint_as_float(u.w)));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&b1_empty[be]);
#pragma unroll
          for (int i = 0; i < 16; ++i) hq[i] = cpk_bf16x2(gelu_fast2(w2[i]));
        }
        // h chunk (A operand of G2): wait until G2(c - 2) has finished reading this buffer
        trc.ev(0x23, c);
        mbar_wait(&h_empty[b], par ^ 1u);
        trc.ev(0x24, c);
        const uint32_t hb = hs_a + b * CXBOX;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          sts128(hb + cbox_off(r, half * 4 + cc), make_uint4(hq[4 * cc], hq[4 * cc + 1], hq[4 * cc + 2], hq[4 * cc + 3]));
        // generic-proxy writes -> async proxy (drains this thread's st.shared), then a plain arrive: release / acquire at
        // CLUSTER scope compile to MEMBAR.ALL.GPU (behind the outstanding z stores: ~2 us per chunk) and CCTL.IVALL
        fence_proxy_async_smem();
        arrive_leader(&h_full[b], hf_remote + b * 8);
        trc.ev(0x25, c);
      }
      // ---- final epilogue: out = acc2 (+ b2); this warp's quarter of the D columns
      trc.ev(0x26, ti);
      mbar_wait(a2_full, (uint32_t)ti & 1u);
      trc.ev(0x27, ti);
      tcgen05_fence_after();
      if (p.dbg & 4) {
        tcgen05_fence_before();
        arrive_leader(a2_empty, a2e_remote);
        mbar_wait(&b2_full[(uint32_t)ti & 1u], ((uint32_t)ti >> 1) & 1u);
        if (lane == 0) mbar_arrive(&b2_empty[(uint32_t)ti & 1u]);
        continue;
      }
      constexpr int CPH = D / 4 / 32;              // 32-column blocks per warp
      const uint32_t tb = (uint32_t)ti & 1u;
      mbar_wait(&b2_full[tb], ((uint32_t)ti >> 1) & 1u);
#pragma unroll 1
      for (int cb = 0; cb < CPH; ++cb) {
        const int col = grp * (D / 4) + cb * 32;
        float v[32];
        tmem_ld_32x32(tmem_base + acc2_col + col + ((uint32_t)(q * 32) << 16), v);
        if (cb == CPH - 1) {
          tcgen05_fence_before();
          arrive_leader(a2_empty, a2e_remote);
        }
        f32x2 w2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w2[i] = pk2(v[2 * i], v[2 * i + 1]);
        {
          const uint32_t ba = smem_u32(b2s) + tb * (D * 4) + col * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 u = lds128(ba + i * 16);
            w2[2 * i] = add2(w2[2 * i], pk2(__uint_as_float(u.x), __uint_as_float(u.y)));
            w2[2 * i + 1] = add2(w2[2 * i + 1], pk2(__uint_as_float(u.z), __uint_as_float(u.w)));
          }
        }
        uint32_t yq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) yq[i] = cpk_bf16x2(w2[i]);
        store_row32(p.out + grow * D + col, yq);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b2_empty[tb]);
      trc.ev(0x28, ti);
    }
    trc.done();
  }
  tcgen05_fence_before();
  cluster_sync();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnC chain_encode() {
  static EncodeTiledFnC fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFnC>(f);
    return static_cast<EncodeTiledFnC>(nullptr);
  }();
  return fn;
}
static int cmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFnC enc = chain_encode();
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

template <int D>
static int launch_chain_t(const void* A, const void* B1, const void* B2, const ChainParams& p, int cap_rows,
                          int max_ctas, cudaStream_t st) {
  using Cfg = ChainCfg<D>;
  CUtensorMap tx, t1, t2;
  int rc = cmap(&tx, A, (uint64_t)cap_rows, D, CBM);                              // A tile rows
  if (rc) return rc;
  rc = cmap(&t1, B1, (uint64_t)p.E * p.H, D, CHC / 2);                            // [E*H][D], 32-row half boxes
  if (rc) return rc;
  rc = cmap(&t2, B2, (uint64_t)p.E * D, (uint64_t)p.H, Cfg::NP / 2);              // [E*D][H], NP/2-row half boxes
  if (rc) return rc;
  auto kern = ffn_chain_kernel<D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return (int)e;
  int grid = cap_rows / (2 * CBM) * 2;
  if (grid > max_ctas / 2 * 2) grid = max_ctas / 2 * 2;
  if (grid < 2) grid = 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_knobs[M3_KNOB_PDL] ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, tx, t1, t2, p);
  if (e != cudaSuccess) return (int)e;
  M3_LAUNCH_CHECK();
  return M3_OK;
}

}  // namespace tc
}  // namespace m3

using namespace m3::tc;

// measurement knobs / timeline of the chain launch (launch 0 of its m3_ffn_fwd / m3_ffn_bwd call: M3_KNOB_TRACE_KERNEL 0 or 1)
static void chain_debug(ChainParams& p) {
  p.dbg = m3::g_knobs[M3_KNOB_DEBUG];
  const int want = m3::g_knobs[M3_KNOB_TRACE_KERNEL];
  p.trace = (want == 0 || want == 1) ? g_trace_buf : nullptr;
  p.trace_cap = g_trace_cap;
}

// 1 if the chain kernel supports this shape (else the two grouped GEMMs of ffn_bf16.cu are used)
int m3_ffn_chain_supported(int D, int H) { return (D == 128 || D == 256 || D == 384) && H % 64 == 0 && H >= 64; }

// forward, no state kept: y = gelu(x W1^T + b1) W2^T + b2
int m3_ffn_chain_fwd(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                     int H, const void* w1, const float* b1, const void* w2, const float* b2, void* yq, int max_ctas,
                     cudaStream_t st) {
  if (cap_rows % (2 * CBM) != 0) return M3_ERR_SHAPE;
  ChainParams p{};
  p.offsets = offsets; p.tile_expert = tile_expert; p.E = E; p.H = H; p.b1 = b1; p.b2 = b2;
  p.out = static_cast<__nv_bfloat16*>(yq);
  chain_debug(p);
  switch (D) {
    case 128: return launch_chain_t<128>(xq, w1, w2, p, cap_rows, max_ctas, st);
    case 256: return launch_chain_t<256>(xq, w1, w2, p, cap_rows, max_ctas, st);
    case 384: return launch_chain_t<384>(xq, w1, w2, p, cap_rows, max_ctas, st);
    default: return M3_ERR_SHAPE;
  }
}

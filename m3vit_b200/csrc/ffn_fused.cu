// Fused expert FFN on CTA pairs:  y = GELU(x W1^T + b1) W2^T + b2  in ONE kernel (sm_100a).
//
// Replaces _Expert.forward of the reference (FMoELinear -> GELU -> FMoELinear,
// /root/reference/models/moe/origin/custom_moe_layer.py:36-44): the hidden activation h never
// leaves the SM.  At D = H = 384 the un-fused pair moves 5 x R*D*2 bytes through HBM (x, hpre, h
// written, h read, y) and is HBM-bound (192 flop/B < ridge); fused it moves 3 x (x, hpre, y) - or
// 2 x without the saved pre-activation (inference) - and becomes tensor/epilogue-bound.
//
// A CTA pair (cluster of 2, cta_group::2) owns a 256-row tile of one expert's queue; each CTA keeps
// its own 128 x D slice of x resident in shared memory and walks the hidden dimension in chunks of 64:
//     G1(j):  acc1[128 x 64]   = x      * W1[e][64j..64j+64, :]^T      (TMEM, double-buffered)
//     mid(j): +b1, store hpre chunk (training), GELU -> bf16 -> h chunk in smem (K-major, swizzled)
//     G2(j):  acc2[128 x D]   += hchunk * W2[e][:, 64j..64j+64]^T      (TMEM, D <= 384 columns)
// and finally  y = acc2 + b2.  Weight tiles stream through a ring of TMA slots; a pair splits every
// weight tile (each CTA loads half), which halves the per-SM weight traffic.  TMEM: 384 + 2*64 = 512.
// Warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only), warps 2-9 = epilogue.
// The same kernel runs the data-gradient chain of the backward pass (MODE_BWD):
//     G1: dh = dy * W2   (B = W2^T),  mid: dhpre = dh * GELU'(hpre) (+ h = GELU(hpre) for wgrad),
//     G2: dx += dhpre * W1   (B = W1^T).
#include <cstdio>

#include "tc_common.cuh"

namespace m3 {
namespace tc {

constexpr int FBM = 128, FBK = 64, FUK = 16, FHC = 64;   // rows per CTA, k-box, UMMA K, hidden chunk
constexpr int kFusedEpiWarps = 8;
constexpr int kFusedThreads = 64 + kFusedEpiWarps * 32;
constexpr int XBOX = FBM * FBK * 2;                       // 16 KB: [128 rows][64 bf16]
constexpr int TBOX = 32 * 32 * 2;                         // 2 KB per-warp transpose box [32 rows][32 bf16]

enum { MODE_FWD = 0, MODE_BWD = 1 };

// Debug timeline (M3_FUSED_TRACE=1): clock64 stamps written by CTA 0; read with m3_debug_read_trace.
__device__ unsigned long long g_trace[8192];
__device__ unsigned int g_trace_n;
__device__ int g_trace_on;
__device__ __forceinline__ void trace(int role, int ev, int j) {
  if (g_trace_on && blockIdx.x == 0) {
    const unsigned int i = atomicAdd(&g_trace_n, 1u);
    if (i < 4096) { g_trace[2 * i] = ((unsigned long long)role << 48) | ((unsigned long long)ev << 32) | (unsigned)j; g_trace[2 * i + 1] = clock64(); }
  }
}

struct FusedParams {
  const int32_t* offsets;      // [E+1]
  const int32_t* tile_expert;  // per 256-row tile
  int E, H;
  const float* b1;             // [E][H]   (FWD)
  const float* b2;             // [E][D]   (FWD)
  const __nv_bfloat16* aux;    // BWD: hpre [rows][H]
  __nv_bfloat16* mid_out;      // FWD: hpre [rows][H] or null   BWD: dhpre [rows][H]
  __nv_bfloat16* mid_out2;     // BWD: h = gelu(hpre) [rows][H]
  __nv_bfloat16* out;          // FWD: y [rows][D]   BWD: dx [rows][D]
};

template <int D>
struct FusedCfg {
  static constexpr int KD = D / FBK;                        // k-boxes of the resident A tile
  static constexpr int NPART = (D > 256) ? 2 : 1;           // G2 output split into <= 256-column MMAs
  static constexpr int NP = D / NPART;                      // columns per G2 MMA (pair-wide N)
  static constexpr int W2_SLOT = (NP / 2) * FBK * 2;        // this CTA's half of a [NP x 64] W2 tile
  static constexpr int W1_BOX = (FHC / 2) * FBK * 2;        // this CTA's half of a [64 x 64] W1 k-box = 4 KB
  static constexpr int W1_BOXES_PER_SLOT = W2_SLOT / W1_BOX;
  static constexpr int W1_SLOTS = (KD + W1_BOXES_PER_SLOT - 1) / W1_BOXES_PER_SLOT;
  static constexpr int SLOT = W2_SLOT;
  static constexpr int X_BYTES = KD * XBOX;
  static constexpr int H_BYTES = 2 * XBOX;                  // two h chunks [128 x 64]
  static constexpr int T_BYTES = kFusedEpiWarps * TBOX;     // one transpose box per epilogue warp
  static constexpr int FIXED = X_BYTES + H_BYTES + T_BYTES;
  static constexpr int NSLOT_RAW = (227 * 1024 - 1024 - 1024 - FIXED) / SLOT;
  static constexpr int NSLOT = NSLOT_RAW > 8 ? 8 : NSLOT_RAW;
  static constexpr int SMEM = FIXED + NSLOT * SLOT + 1024 + 1024;
  static_assert(D % 64 == 0 && D <= 384, "fused FFN keeps the [128 x D] output accumulator in TMEM");
  static_assert(NP % 32 == 0 && NP <= 256, "G2 MMA width");
  static_assert(NSLOT >= 4, "weight ring too shallow");
};

// 16-byte chunk c (0..7) of row r inside a [rows][64 bf16] swizzle-128B box
__device__ __forceinline__ uint32_t fbox_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }
// 16-byte chunk c (0..3) of row r inside a per-warp [32 rows][32 bf16] transpose box (conflict-free both ways)
__device__ __forceinline__ uint32_t tbox_off(int r, int c) { return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ uint32_t pk_bf16x2(f32x2 v) {
  float a, b;
  unpk2(v, a, b);
  return float2_to_bf16x2(a, b);
}
__device__ __forceinline__ f32x2 unpk_bf16x2(uint32_t u) {
  const float2 f = bf16x2_to_float2(u);
  return pk2(f.x, f.y);
}

template <int D, int MODE>
__global__ void __launch_bounds__(kFusedThreads, 1)
ffn_chain_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, FusedParams p) {
  using Cfg = FusedCfg<D>;
  constexpr int KD = Cfg::KD, NPART = Cfg::NPART, NP = Cfg::NP, NSLOT = Cfg::NSLOT;
  constexpr int W1S = Cfg::W1_SLOTS, BPS = Cfg::W1_BOXES_PER_SLOT;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader_cta = cta_rank == 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                   // [KD][128][64] resident A tile
  uint8_t* hs = xs + Cfg::X_BYTES;                      // [2][128][64]   h chunks (A operand of G2)
  uint8_t* ring = hs + Cfg::H_BYTES;                    // [NSLOT][SLOT]  weight ring
  uint8_t* tb = ring + NSLOT * Cfg::SLOT;               // per-warp transpose + aux boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(tb + Cfg::T_BYTES);
  uint64_t* w_full = bars;                  // [NSLOT] leader-waited, count 2
  uint64_t* w_empty = w_full + NSLOT;       // [NSLOT] count 1 (multicast commit)
  uint64_t* x_full = w_empty + NSLOT;       // leader-waited, count 2
  uint64_t* x_empty = x_full + 1;           // count 1 (multicast commit)
  uint64_t* a1_full = x_empty + 1;          // [2] count 1 (multicast commit)
  uint64_t* a1_empty = a1_full + 2;         // [2] leader-waited, count 2*8 warps
  uint64_t* h_full = a1_empty + 2;          // [2] leader-waited, count 2*8 warps
  uint64_t* h_empty = h_full + 2;           // [2] count 1 (multicast commit)
  uint64_t* a2_full = h_empty + 2;          // count 1 (multicast commit)
  uint64_t* a2_empty = a2_full + 1;         // leader-waited, count 2*8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 2); mbar_init(&w_empty[s], 1); }
      mbar_init(x_full, 2); mbar_init(x_empty, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&a1_full[b], 1); mbar_init(&a1_empty[b], 2 * kFusedEpiWarps);
        mbar_init(&h_full[b], 2 * kFusedEpiWarps); mbar_init(&h_empty[b], 1);
      }
      mbar_init(a2_full, 1); mbar_init(a2_empty, 2 * kFusedEpiWarps);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm<512>(tmem_slot);
  }
  tcgen05_fence_before();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc2_col = 0, acc1_col = 384;   // acc1 buffers at columns 384 and 448

  const int NH = p.H / FHC;
  const int total = p.offsets[p.E] / (2 * FBM);         // 256-row pair tiles
  const int unit = blockIdx.x / 2, n_units = gridDim.x / 2;

  if (warp == 0) {
    // ===================================================== TMA producer (one thread per CTA)
    if (lane == 0) {
      int slot = 0;
      uint32_t sphase = 0, xuse = 0;
      auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; sphase ^= 1; } };
      for (int pt = unit; pt < total; pt += n_units, ++xuse) {
        const int e = p.tile_expert[pt];
        const int row0 = (pt * 2 + (int)cta_rank) * FBM;
        // resident A tile (free once the previous tile's last G1 has been read)
        trace(0, 0, xuse);
        mbar_wait(x_empty, (xuse & 1) ^ 1);
        trace(0, 1, xuse);
        {
          const uint32_t bar = mapa_u32(smem_u32(x_full), 0);
          if (leader_cta) mbar_expect_tx(x_full, 2 * Cfg::X_BYTES); else mbar_arrive_remote(bar);
          for (int kc = 0; kc < KD; ++kc) tma_load_2d_2sm(xs + kc * XBOX, &tmX, bar, kc * FBK, row0);
        }
        for (int j = 0; j <= NH; ++j) {
          if (j < NH) {  // W1 chunk j: this CTA's 32 of the 64 hidden rows, all KD k-boxes
            const int wrow = e * p.H + j * FHC + (int)cta_rank * (FHC / 2);
            for (int s = 0; s < W1S; ++s) {
              mbar_wait(&w_empty[slot], sphase ^ 1);
              const int nb = (KD - s * BPS) < BPS ? (KD - s * BPS) : BPS;
              const uint32_t bar = mapa_u32(smem_u32(&w_full[slot]), 0);
              if (leader_cta) mbar_expect_tx(&w_full[slot], 2 * nb * Cfg::W1_BOX); else mbar_arrive_remote(bar);
              for (int b = 0; b < nb; ++b)
                tma_load_2d_2sm(ring + slot * Cfg::SLOT + b * Cfg::W1_BOX, &tmW1, bar, (s * BPS + b) * FBK, wrow);
              next_slot();
            }
          }
          if (j > 0) {   // W2 chunk j-1: this CTA's half of every NP-row part, k = hidden chunk j-1
            for (int part = 0; part < NPART; ++part) {
              mbar_wait(&w_empty[slot], sphase ^ 1);
              const uint32_t bar = mapa_u32(smem_u32(&w_full[slot]), 0);
              if (leader_cta) mbar_expect_tx(&w_full[slot], 2 * Cfg::W2_SLOT); else mbar_arrive_remote(bar);
              tma_load_2d_2sm(ring + slot * Cfg::SLOT, &tmW2, bar, (j - 1) * FHC,
                              e * D + part * NP + (int)cta_rank * (NP / 2));
              next_slot();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA, one thread)
    if (lane == 0 && leader_cta) {
      constexpr uint32_t idesc1 = make_idesc_bf16(2 * FBM, FHC, 0, 0);   // 256 x 64
      constexpr uint32_t idesc2 = make_idesc_bf16(2 * FBM, NP, 0, 0);    // 256 x NP
      int slot = 0;
      uint32_t sphase = 0, tile_it = 0;
      uint32_t use1[2] = {0, 0}, useh[2] = {0, 0};
      auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; sphase ^= 1; } };
      const uint32_t xs_a = smem_u32(xs), hs_a = smem_u32(hs), ring_a = smem_u32(ring);
      for (int pt = unit; pt < total; pt += n_units, ++tile_it) {
        trace(1, 0, tile_it);
        mbar_wait(x_full, tile_it & 1);
        tcgen05_fence_after();
        trace(1, 1, tile_it);
        for (int j = 0; j <= NH; ++j) {
          if (j < NH) {                                    // ---- G1(j): acc1[j&1] = x * W1_j^T
            const int b = j & 1;
            mbar_wait(&a1_empty[b], (use1[b] & 1) ^ 1);
            ++use1[b];
            tcgen05_fence_after();
            trace(1, 2, j);
            const uint32_t d1 = tmem_base + acc1_col + b * FHC;
            for (int s = 0; s < W1S; ++s) {
              mbar_wait(&w_full[slot], sphase);
              tcgen05_fence_after();
              const int nb = (KD - s * BPS) < BPS ? (KD - s * BPS) : BPS;
              for (int bb = 0; bb < nb; ++bb) {
                const int kc = s * BPS + bb;
#pragma unroll
                for (int k = 0; k < FBK / FUK; ++k) {
                  const uint64_t ad = make_smem_desc(xs_a + kc * XBOX + k * FUK * 2, 0, 1024);
                  const uint64_t bd = make_smem_desc(ring_a + slot * Cfg::SLOT + bb * Cfg::W1_BOX + k * FUK * 2, 0, 1024);
                  umma_bf16_2sm(d1, ad, bd, idesc1, (kc | k) != 0);
                }
              }
              umma_commit_2sm(&w_empty[slot], 3);
              next_slot();
            }
            umma_commit_2sm(&a1_full[b], 3);
            trace(1, 3, j);
            if (j == NH - 1) umma_commit_2sm(x_empty, 3);   // A tile no longer needed: next tile may load
          }
          if (j > 0) {                                     // ---- G2(j-1): acc2 += h_{j-1} * W2_{j-1}^T
            const int jj = j - 1, b = jj & 1;
            if (jj == 0) {                                 // acc2 must have been drained by the last epilogue
              mbar_wait(a2_empty, (tile_it & 1) ^ 1);
              tcgen05_fence_after();
            }
            mbar_wait(&h_full[b], useh[b] & 1);
            ++useh[b];
            tcgen05_fence_after();
            trace(1, 4, jj);
            for (int part = 0; part < NPART; ++part) {
              mbar_wait(&w_full[slot], sphase);
              tcgen05_fence_after();
              const uint32_t d2 = tmem_base + acc2_col + part * NP;
#pragma unroll
              for (int k = 0; k < FHC / FUK; ++k) {
                const uint64_t ad = make_smem_desc(hs_a + b * XBOX + k * FUK * 2, 0, 1024);
                const uint64_t bd = make_smem_desc(ring_a + slot * Cfg::SLOT + k * FUK * 2, 0, 1024);
                umma_bf16_2sm(d2, ad, bd, idesc2, (jj | k) != 0);
              }
              umma_commit_2sm(&w_empty[slot], 3);
              next_slot();
            }
            umma_commit_2sm(&h_empty[b], 3);
            trace(1, 5, jj);
            if (jj == NH - 1) umma_commit_2sm(a2_full, 3);
          }
        }
      }
    }
  } else {
    // ===================================================== epilogue warps (both CTAs)
    const int q = warp & 3;                      // TMEM lane quarter
    const int half = (warp - 2) >> 2;            // which 32 of a chunk's 64 columns / which half of D
    const int ew = warp - 2;
    uint8_t* tbox = tb + ew * TBOX;
    const uint32_t a1e_remote = mapa_u32(smem_u32(&a1_empty[0]), 0), a1e_remote1 = mapa_u32(smem_u32(&a1_empty[1]), 0);
    const uint32_t hf_remote = mapa_u32(smem_u32(&h_full[0]), 0), hf_remote1 = mapa_u32(smem_u32(&h_full[1]), 0);
    const uint32_t a2e_remote = mapa_u32(smem_u32(a2_empty), 0);
    auto arrive_leader = [&](uint64_t* local, uint32_t remote) {     // one arrival per warp
      __syncwarp();
      if (lane == 0) { if (leader_cta) mbar_arrive(local); else mbar_arrive_remote(remote); }
    };
    // registers (one row per lane, 32 columns) -> transpose box -> coalesced 64-B row segments
    auto flush = [&](__nv_bfloat16* dst, int ld, int64_t grow0, int col) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + (lane >> 2), ch = lane & 3;
        const uint4 u = *reinterpret_cast<const uint4*>(tbox + tbox_off(rr, ch));
        stg_stream(dst + (grow0 + rr) * ld + col + ch * 8, u);
      }
      __syncwarp();
    };
    uint32_t use1[2] = {0, 0}, useh[2] = {0, 0}, tile_it = 0;
    for (int pt = unit; pt < total; pt += n_units, ++tile_it) {
      const int e = p.tile_expert[pt];
      const int64_t grow0 = (int64_t)(pt * 2 + (int)cta_rank) * FBM + q * 32;   // this warp's first global row
      const int r = q * 32 + lane;                                             // row inside the CTA tile
#pragma unroll 1
      for (int j = 0; j < NH; ++j) {
        const int b = j & 1;
        const int hcol = j * FHC + half * 32;
        // BWD: this lane's 32 hpre values (64 contiguous bytes) - issue before waiting on the MMA
        uint4 hraw[4];
        if (MODE == MODE_BWD) {
          const uint4* src = reinterpret_cast<const uint4*>(p.aux + (grow0 + lane) * p.H + hcol);
#pragma unroll
          for (int c = 0; c < 4; ++c) hraw[c] = __ldg(src + c);
        }
        if (warp == 2 && lane == 0) trace(2, 0, j);
        mbar_wait(&a1_full[b], use1[b] & 1);
        ++use1[b];
        tcgen05_fence_after();
        if (warp == 2 && lane == 0) trace(2, 1, j);
        float v[32];
        tmem_ld_32x32(tmem_base + acc1_col + b * FHC + half * 32 + ((uint32_t)(q * 32) << 16), v);
        tcgen05_fence_before();
        arrive_leader(&a1_empty[b], b ? a1e_remote1 : a1e_remote);   // acc1[b] may be overwritten by G1(j+2)
        if (warp == 2 && lane == 0) trace(2, 2, j);
        f32x2 w2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w2[i] = pk2(v[2 * i], v[2 * i + 1]);
        uint32_t hq[16];                                             // what goes into the h chunk (A of G2)
        if (MODE == MODE_FWD) {
          const float4* b4 = reinterpret_cast<const float4*>(p.b1 + (int64_t)e * p.H + hcol);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i);
            w2[2 * i] = add2(w2[2 * i], pk2(bb.x, bb.y));
            w2[2 * i + 1] = add2(w2[2 * i + 1], pk2(bb.z, bb.w));
          }
          if (p.mid_out != nullptr) {                                // save the pre-activation for backward
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(tbox + tbox_off(lane, c)) =
                  make_uint4(pk_bf16x2(w2[4 * c]), pk_bf16x2(w2[4 * c + 1]), pk_bf16x2(w2[4 * c + 2]), pk_bf16x2(w2[4 * c + 3]));
            flush(p.mid_out, p.H, grow0, hcol);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) hq[i] = pk_bf16x2(gelu_fast2(w2[i]));
        } else {
          uint32_t hact[16];
          const uint32_t hw[16] = {hraw[0].x, hraw[0].y, hraw[0].z, hraw[0].w, hraw[1].x, hraw[1].y, hraw[1].z, hraw[1].w,
                                   hraw[2].x, hraw[2].y, hraw[2].z, hraw[2].w, hraw[3].x, hraw[3].y, hraw[3].z, hraw[3].w};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f32x2 gl;
            const f32x2 gr = gelu_fast_grad2(unpk_bf16x2(hw[i]), &gl);
            hq[i] = pk_bf16x2(mul2(w2[i], gr));                      // dhpre
            hact[i] = pk_bf16x2(gl);                                 // h = gelu(hpre), for wgrad
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(tbox + tbox_off(lane, c)) = make_uint4(hact[4 * c], hact[4 * c + 1], hact[4 * c + 2], hact[4 * c + 3]);
          flush(p.mid_out2, p.H, grow0, hcol);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(tbox + tbox_off(lane, c)) = make_uint4(hq[4 * c], hq[4 * c + 1], hq[4 * c + 2], hq[4 * c + 3]);
          flush(p.mid_out, p.H, grow0, hcol);
        }
        // h chunk (A operand of G2): wait until G2(j-2) has finished reading this buffer
        if (warp == 2 && lane == 0) trace(2, 3, j);
        mbar_wait(&h_empty[b], (useh[b] & 1) ^ 1);
        ++useh[b];
        uint8_t* hb = hs + b * XBOX;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(hb + fbox_off(r, half * 4 + c)) = make_uint4(hq[4 * c], hq[4 * c + 1], hq[4 * c + 2], hq[4 * c + 3]);
        fence_proxy_async_smem();
        arrive_leader(&h_full[b], b ? hf_remote1 : hf_remote);
        if (warp == 2 && lane == 0) trace(2, 4, j);
      }
      // ---- final epilogue: out = acc2 (+ b2)
      mbar_wait(a2_full, tile_it & 1);
      tcgen05_fence_after();
      constexpr int CPH = D / 2 / 32;              // 32-column chunks per warp (its half of D)
#pragma unroll 1
      for (int c = 0; c < CPH; ++c) {
        const int col = half * (D / 2) + c * 32;
        float v[32];
        tmem_ld_32x32(tmem_base + acc2_col + col + ((uint32_t)(q * 32) << 16), v);
        if (c == CPH - 1) {
          tcgen05_fence_before();
          arrive_leader(a2_empty, a2e_remote);
        }
        f32x2 w2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w2[i] = pk2(v[2 * i], v[2 * i + 1]);
        if (MODE == MODE_FWD) {
          const float4* b4 = reinterpret_cast<const float4*>(p.b2 + (int64_t)e * D + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i);
            w2[2 * i] = add2(w2[2 * i], pk2(bb.x, bb.y));
            w2[2 * i + 1] = add2(w2[2 * i + 1], pk2(bb.z, bb.w));
          }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          *reinterpret_cast<uint4*>(tbox + tbox_off(lane, cc)) =
              make_uint4(pk_bf16x2(w2[4 * cc]), pk_bf16x2(w2[4 * cc + 1]), pk_bf16x2(w2[4 * cc + 2]), pk_bf16x2(w2[4 * cc + 3]));
        flush(p.out, D, grow0, col);
      }
      if (warp == 2 && lane == 0) trace(2, 7, tile_it);
    }
  }
  tcgen05_fence_before();
  cluster_sync();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnF)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnF fused_encode() {
  static EncodeTiledFnF fn = nullptr;
  if (fn == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnF>(f);
  }
  return fn;
}
static int fmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFnF enc = fused_encode();
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

template <int D, int MODE>
static int launch_chain_t(const void* A, const void* B1, const void* B2, const FusedParams& p, int cap_rows,
                          cudaStream_t st) {
  using Cfg = FusedCfg<D>;
  CUtensorMap tx, t1, t2;
  int rc = fmap(&tx, A, (uint64_t)cap_rows, D, FBM);                              // A tile rows
  if (rc) return rc;
  rc = fmap(&t1, B1, (uint64_t)p.E * p.H, D, FHC / 2);                            // [E*H][D], 32-row half boxes
  if (rc) return rc;
  rc = fmap(&t2, B2, (uint64_t)p.E * D, (uint64_t)p.H, Cfg::NP / 2);              // [E*D][H], NP/2-row half boxes
  if (rc) return rc;
  auto kern = ffn_chain_kernel<D, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return (int)e;
  int grid = cap_rows / (2 * FBM) * 2;
  if (grid > kNumSMs / 2 * 2) grid = kNumSMs / 2 * 2;
  if (grid < 2) grid = 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, tx, t1, t2, p);
  if (e != cudaSuccess) return (int)e;
  M3_LAUNCH_CHECK();
  return M3_OK;
}

}  // namespace tc
}  // namespace m3

using namespace m3::tc;

// debug: enable / read the clock64 timeline of CTA 0 (not part of the public ABI contract)
extern "C" int m3_debug_trace(int enable, unsigned long long* host_out, int max_events) {
  int on = enable;
  unsigned int n = 0;
  if (host_out != nullptr) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n));
    if ((int)n > max_events) n = max_events;
    cudaMemcpyFromSymbol(host_out, g_trace, (size_t)n * 16);
  }
  unsigned int zero = 0;
  cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero));
  cudaMemcpyToSymbol(g_trace_on, &on, sizeof(on));
  return (int)n;
}

// 1 if the fused chain kernel supports this shape (else the un-fused gg kernels are used)
int m3_ffn_fused_supported(int D, int H) { return (D == 128 || D == 256 || D == 384) && H % 64 == 0 && H >= 64; }

// forward: y = gelu(x W1^T + b1) W2^T + b2;  hpre (nullable) saved for backward
int m3_ffn_fused_fwd(const void* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                     int H, const void* w1, const float* b1, const void* w2, const float* b2, void* hpre, void* yq,
                     cudaStream_t st) {
  if (cap_rows % (2 * FBM) != 0) return M3_ERR_SHAPE;
  FusedParams p{};
  p.offsets = offsets; p.tile_expert = tile_expert; p.E = E; p.H = H; p.b1 = b1; p.b2 = b2;
  p.mid_out = static_cast<__nv_bfloat16*>(hpre);
  p.out = static_cast<__nv_bfloat16*>(yq);
  switch (D) {
    case 128: return launch_chain_t<128, MODE_FWD>(xq, w1, w2, p, cap_rows, st);
    case 256: return launch_chain_t<256, MODE_FWD>(xq, w1, w2, p, cap_rows, st);
    case 384: return launch_chain_t<384, MODE_FWD>(xq, w1, w2, p, cap_rows, st);
    default: return M3_ERR_SHAPE;
  }
}

// backward data-gradient chain: dhpre = (dy W2) * gelu'(hpre), h = gelu(hpre), dx = dhpre W1
//   w2t = W2^T [E][H][D],  w1t = W1^T [E][D][H]
int m3_ffn_fused_bwd(const void* dyq, const void* hpre, const int32_t* offsets, const int32_t* tile_expert,
                     int cap_rows, int E, int D, int H, const void* w2t, const void* w1t, void* dhpre, void* h,
                     void* dxq, cudaStream_t st) {
  if (cap_rows % (2 * FBM) != 0) return M3_ERR_SHAPE;
  FusedParams p{};
  p.offsets = offsets; p.tile_expert = tile_expert; p.E = E; p.H = H;
  p.aux = static_cast<const __nv_bfloat16*>(hpre);
  p.mid_out = static_cast<__nv_bfloat16*>(dhpre);
  p.mid_out2 = static_cast<__nv_bfloat16*>(h);
  p.out = static_cast<__nv_bfloat16*>(dxq);
  switch (D) {
    case 128: return launch_chain_t<128, MODE_BWD>(dyq, w2t, w1t, p, cap_rows, st);
    case 256: return launch_chain_t<256, MODE_BWD>(dyq, w2t, w1t, p, cap_rows, st);
    case 384: return launch_chain_t<384, MODE_BWD>(dyq, w2t, w1t, p, cap_rows, st);
    default: return M3_ERR_SHAPE;
  }
}

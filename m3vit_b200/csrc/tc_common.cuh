// Blackwell (sm_100a) building blocks as inline PTX: mbarrier, TMA
// (cp.async.bulk.tensor), TMEM allocation, tcgen05.mma / commit / ld, and the
// shared-memory / instruction descriptors.  Hand-written; bit layouts follow the
// PTX ISA tcgen05 descriptor tables.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the driver entry point is resolved at run time)

#include "common.cuh"

namespace m3 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Debug timeline (m3_debug_trace_buffer): {tag, clock64} pairs written by CTA 0 with plain stores (no atomics: an
// atomic round trip per event costs ~1000 clk and drowns what is being measured).  Role r (0 producer, 1 MMA,
// 2 epilogue warp 0) owns events [r*cap, (r+1)*cap) and counts them in a register; buf[r] = its final count.
extern unsigned long long* g_trace_buf;   // abi.cu
extern int g_trace_cap;
// Compiled in only with -DM3_GEMM_TRACE (M3_GEMM_TRACE=1 python -m m3vit_b200.build --force): even switched off at run
// time, the per-lane `buf` test is a potentially divergent branch inside the converged producer / MMA loops and cost
// 5-10 us per GEMM launch.
#ifdef M3_GEMM_TRACE
struct Tracer {
  unsigned long long* buf;
  int cap, role, n;
  __device__ __forceinline__ Tracer(unsigned long long* b, int c, int r)
      : buf((blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0) ? b : nullptr),
        cap(c), role(r), n(0) {}
  __device__ __forceinline__ void ev(uint32_t tag, uint32_t j) {
    if (buf != nullptr && n < cap) {
      unsigned long long* e = buf + 4 + 2 * ((size_t)role * cap + n);
      e[0] = ((unsigned long long)tag << 32) | j;
      e[1] = (unsigned long long)clock64();
      ++n;
    }
  }
  __device__ __forceinline__ void done() { if (buf != nullptr) buf[role] = (unsigned long long)n; }
};
#else
struct Tracer {
  __device__ __forceinline__ Tracer(unsigned long long*, int, int) {}
  __device__ __forceinline__ void ev(uint32_t, uint32_t) {}
  __device__ __forceinline__ void done() {}
};
#endif

// explicit shared-space 128-bit accesses: pointers carved out of the dynamic smem block by integer arithmetic lose
// their address space and compile to GENERIC LD.E / ST.E (long-scoreboard latency); these stay LDS / STS
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ----------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must fault the kernel (trap), never hang the GPU.  The polling loop with its time-out
// lives OUT OF LINE: inlined at every wait site (~35 SASS instructions each, printf call included) it bloated the
// warp-specialised loops past the instruction cache.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      printf("m3 tcgen05: mbarrier timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(smem_u32(bar), parity);
}
// two barriers at once: both try_waits are in flight together (a try_wait on a completed barrier still takes ~90 clk,
// and the issuing warps of the chain kernel do little else)
__device__ __forceinline__ void mbar_wait2(uint64_t* a, uint32_t pa, uint64_t* b, uint32_t pb) {
  const bool oa = mbar_try_wait(a, pa), ob = mbar_try_wait(b, pb);
  if (!oa) mbar_wait_slow(smem_u32(a), pa);
  if (!ob) mbar_wait_slow(smem_u32(b), pb);
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// --------------------------------------------------------------------- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------ tcgen05.mma
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulation, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------ CTA pairs (cta_group::2)
// Two CTAs of a cluster (same TPC) run ONE tcgen05.mma over a 256-row tile: each CTA holds its own
// 128 rows of A and HALF of the B tile; accumulators live in both CTAs' TMEM.  Only the leader
// (cluster rank 0) issues MMAs; both issue TMA, whose transaction bytes land on the LEADER's mbarrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issued MMAs completed) on the mbarrier at this smem offset in EVERY CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
//   [4,6) D format (1 = f32)   [7,10) A format (1 = bf16)   [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1:
//   [0,14) start address >> 4   [16,30) leading-dim byte offset >> 4   [32,46) stride-dim byte offset >> 4
//   [46,48) version = 1         [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand tile [rows][64 bf16] (128 B per row, TMA SWIZZLE_128B): 8-row groups are
//   1024 B apart (SBO); LBO unused.  K advances by +32 B (>>4 = 2) per UMMA_K = 16.
// MN-major operand tile: [k rows][64 bf16 of M/N] per TMA box; 8-k-row groups 1024 B apart
//   (SBO), next 64 M/N elements one box further (LBO = box bytes).  K advances by 16 rows = 2048 B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 1-D bulk copy global -> shared (this CTA), completion on a local mbarrier; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load_1d_elect(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}\n"
      ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------ lean MMA issue (whole warp converged, one elected lane issues)
// The issuing warp runs its loop CONVERGED (all 32 lanes wait on the mbarriers and do the - warp-uniform - descriptor
// arithmetic), and only the tcgen05 instruction itself is predicated on elect.sync.  Inside an `if (lane == 0)` region
// ptxas cannot prove uniformity and wraps EVERY uniform-datapath instruction (UTCHMMA, UTCBAR, UTMALDG) in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop and recomputes the 64-bit descriptors from scratch: ~17 SASS instructions
// per MMA, 700 clk per 4-MMA k-chunk - the single issuing thread, not HBM or shared memory, bounded every GEMM here.
// Descriptors are passed as {lo, hi} halves: hi is loop-invariant, lo = base_lo + ((byte offset) >> 4).
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);     // version 1, SWIZZLE_128B
}
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 da, db;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                    uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 da, db;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Several k-boxes of MMAs under ONE election: the descriptor halves reach the uniform datapath once and advance there
// (~3 SASS instructions per MMA instead of ~16 with one asm block per MMA).  It matters for the N = 64 MMAs of the
// chain kernel, whose floor is 32 clk each: issued one by one they cost ~60-100 clk apiece and the issuing thread,
// not the tensor pipe, set the pace.
//   NBOX boxes of 4 MMAs (UMMA_K = 16: descriptors 32 B = 2 units apart inside a 64-wide SWIZZLE_128B k-box); between
//   boxes A advances by a_box and B by b_box (descriptor units of 16 bytes); all accumulate into d_tmem, the very first
//   MMA with `accumulate_first`.
#define M3_MMA2(ACC) "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, " ACC ";\n"
#define M3_ADV2 "add.s64 da, da, 2;\nadd.s64 db, db, 2;\n"
#define M3_BOX_REST M3_ADV2 M3_MMA2("1") M3_ADV2 M3_MMA2("1") M3_ADV2 M3_MMA2("1")
#define M3_NEXT_BOX "add.s64 da, da, %6;\nadd.s64 db, db, %7;\n" M3_MMA2("1") M3_BOX_REST
template <int NBOX>
__device__ __forceinline__ void umma_bf16_2sm_elect_boxes(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                          uint32_t idesc, uint32_t accumulate_first, uint32_t a_box,
                                                          uint32_t b_box) {
  static_assert(NBOX >= 1 && NBOX <= 6, "boxes per block");
  // after a box the descriptors stand 6 units past its start
  const unsigned long long sa = (unsigned long long)a_box - 6ull, sb = (unsigned long long)b_box - 6ull;
#define M3_BLOCK(BODY)                                                                                                  \
  asm volatile("{\n.reg .pred p, q;\n.reg .b64 da, db;\nelect.sync _|q, 0xffffffff;\nmov.b64 da, {%1, %3};\n"           \
               "mov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n" M3_MMA2("p") M3_BOX_REST BODY "}\n" ::"r"(d_tmem),       \
               "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate_first), "l"(sa), "l"(sb)                       \
               : "memory")
  if constexpr (NBOX == 1) M3_BLOCK("");
  if constexpr (NBOX == 2) M3_BLOCK(M3_NEXT_BOX);
  if constexpr (NBOX == 3) M3_BLOCK(M3_NEXT_BOX M3_NEXT_BOX);
  if constexpr (NBOX == 4) M3_BLOCK(M3_NEXT_BOX M3_NEXT_BOX M3_NEXT_BOX);
  if constexpr (NBOX == 5) M3_BLOCK(M3_NEXT_BOX M3_NEXT_BOX M3_NEXT_BOX M3_NEXT_BOX);
  if constexpr (NBOX == 6) M3_BLOCK(M3_NEXT_BOX M3_NEXT_BOX M3_NEXT_BOX M3_NEXT_BOX M3_NEXT_BOX);
#undef M3_BLOCK
}
__device__ __forceinline__ void umma_bf16_2sm_elect_x4(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                       uint32_t idesc, uint32_t accumulate_first) {
  umma_bf16_2sm_elect_boxes<1>(d_tmem, a_lo, b_lo, hi, idesc, accumulate_first, 6, 6);
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_elect(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// elected-lane variants of the producer-side instructions (same converged-warp scheme)
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_elect(uint64_t* bar) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q mbarrier.arrive.shared::cta.b64 _, [%0];\n}\n" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_elect(uint32_t cluster_addr) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q mbarrier.arrive.shared::cluster.b64 _, [%0];\n}\n" ::"r"(
                   cluster_addr)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n}\n"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                      int c0, int c1) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n}\n"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------- tcgen05.ld
// 32 lanes x 32 columns of fp32: thread i of the warp gets TMEM lane (lane_base + i), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc
}  // namespace m3

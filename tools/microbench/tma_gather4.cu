// Feasibility test of TMA tile::gather4 (sm_100a): 4 arbitrary rows of a row-major [T][D] bf16 matrix per instruction,
// landing as 4 consecutive 128-byte rows of a SWIZZLE_128B box; an out-of-range row index must read as zeros.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu -lcuda && ./tma_gather4
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../../m3vit_b200/csrc/tc_common.cuh"
namespace m3 { int g_knobs[M3_KNOB_COUNT_] = {0}; namespace tc { unsigned long long* g_trace_buf = nullptr; int g_trace_cap = 0; } }
using namespace m3::tc;

__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               :: "r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// one CTA gathers 128 rows x 64 columns (column block cb) listed in idx[] and writes them out un-swizzled
__global__ void k(const __grid_constant__ CUtensorMap tm, const int* idx, int cb, __nv_bfloat16* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 16384);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 16384);
    for (int i = 0; i < 32; ++i)
      gather4(smem_u32(smem + i * 512), &tm, smem_u32(bar), cb * 64, idx[4 * i], idx[4 * i + 1], idx[4 * i + 2], idx[4 * i + 3]);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i / 8, c = i % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(smem + r * 128 + ((c ^ (r & 7)) << 4));
    *reinterpret_cast<uint4*>(out + r * 64 + c * 8) = v;
  }
}

int main() {
  const int T = 1000, D = 384;
  std::vector<__nv_bfloat16> h((size_t)T * D);
  for (int t = 0; t < T; ++t) for (int d = 0; d < D; ++d) h[(size_t)t * D + d] = __float2bfloat16((float)(t % 251) + d * 0.001f * (d % 7));
  std::vector<int> idx(128);
  for (int i = 0; i < 128; ++i) idx[i] = (i * 37 + 11) % T;
  idx[5] = T; idx[77] = T + 100; idx[127] = T;          // out of range -> zeros
  __nv_bfloat16 *dx, *dout; int* didx;
  cudaMalloc(&dx, h.size() * 2); cudaMalloc(&dout, 128 * 64 * 2); cudaMalloc(&didx, 128 * 4);
  cudaMemcpy(dx, h.data(), h.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(didx, idx.data(), 512, cudaMemcpyHostToDevice);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))f;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)T}; cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {64, 1}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  int bad_total = 0;
  for (int cb = 0; cb < D / 64; cb += 5) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    k<<<1, 128, 20000>>>(tm, didx, cb, dout);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<__nv_bfloat16> o(128 * 64);
    cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128; ++i) for (int c = 0; c < 64; ++c) {
      const float want = idx[i] < T ? __bfloat162float(h[(size_t)idx[i] * D + cb * 64 + c]) : 0.f;
      if (__bfloat162float(o[i * 64 + c]) != want) ++bad;
    }
    printf("cb=%d cuda=%s mismatches=%d\n", cb, cudaGetErrorString(e), bad);
    bad_total += bad;
  }
  printf(bad_total == 0 ? "GATHER4 OK\n" : "GATHER4 MISMATCH\n");
  return bad_total != 0;
}

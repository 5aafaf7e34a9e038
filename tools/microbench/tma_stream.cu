// Microbenchmark: how fast can TMA stream a [rows][384] bf16 activation matrix into smem
//  mode 0: [128 rows x 64 cols] swizzle-128B boxes of a row-major matrix (row stride 768 B)  <- what gg_kernel does
//  mode 1: the same bytes as contiguous 16 KB chunks (1-D bulk copies)                        <- "blocked" layout
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda && ./tma_stream
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../m3vit_b200/csrc/tc_common.cuh"
using namespace m3::tc;
constexpr int STAGES = 8, BOXB = 16384;
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int m_tiles, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * BOXB);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x)
      for (int kc = 0; kc < 6; ++kc) {
        mbar_wait(&empty[stage], ph ^ 1);
        mbar_expect_tx(&full[stage], BOXB);
        if (mode == 0) tma_load_2d(smem + stage * BOXB, &tm, &full[stage], kc * 64, t * 128);
        else {
          const uint8_t* src = base + ((size_t)t * 6 + kc) * BOXB;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       :: "r"(smem_u32(smem + stage * BOXB)), "l"(src), "r"(BOXB), "r"(smem_u32(&full[stage])) : "memory");
        }
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x)
      for (int kc = 0; kc < 6; ++kc) {
        mbar_wait(&full[stage], ph);
        mbar_arrive(&empty[stage]);
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
  }
}
int main() {
  const int rows = 155648 * 4, cols = 384;   // 4 x the B=32 queue (478 MB) to amortise launch/ramp
  size_t bytes = (size_t)rows * cols * 2;
  uint8_t* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
  uint8_t* flush; cudaMalloc(&flush, 256 << 20);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
  int smem = STAGES * BOXB + 1024 + 256;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {148, 296, 444}) {
      float best = 1e9;
      for (int it = 0; it < 5; ++it) {
        cudaMemsetAsync(flush, it, 256 << 20);
        cudaEventRecord(a);
        stream_kernel<<<grid, 64, smem>>>(tm, d, rows / 128, mode);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      printf("mode %d grid %d (smem %d KB/CTA): %.1f us  %.2f TB/s  err=%s\n", mode, grid, smem / 1024, best * 1e3, bytes / (best * 1e-3) / 1e12,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

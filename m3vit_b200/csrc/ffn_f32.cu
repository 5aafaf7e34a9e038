// fp32 expert FFN over the padded expert queues (SIMT FFMA grouped GEMM).
//
// This is the PARITY-MODE path: the reference trains its MoE experts in fp32
// (no autocast in /root/reference/train/train_utils.py) through fmoe's
// FMoELinear = one cuBLAS SGEMM per expert (models/moe/origin/custom_moe_layer.py:36-44).
// Here every linear of the pair is ONE grouped launch over all experts; bias, exact
// GELU and GELU' are fused into the epilogues.  The tcgen05 bf16 path (ffn_bf16.cu)
// is the performance path; this one exists so outputs/gradients can be compared
// with the fp32 oracle at 1e-5.
//
// 64x64x16 tiles, 256 threads, 4x4 register tile, sequential-k fp32 FMA.
#include "common.cuh"
#include "philox.cuh"

namespace m3 {

constexpr int BM = 64, BN = 64, BK = 16, SG_THREADS = 256;

enum { LAY_NT = 0, LAY_NN = 1, LAY_TN = 2 };
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU_SAVE = 2, EPI_GELU_GRAD = 3 };

struct SgemmParams {
  const float* A;      // rows-grouped: [rows, Kdim];  TN: [rows, M]
  const float* B;      // NT: [E][N][Kdim]; NN: [E][Kdim][N]; TN: [rows, N]
  float* C;            // rows-grouped: [rows, N];  TN: [E][M][N]
  const float* bias;   // [E][N]
  const float* aux;    // EPI_GELU_GRAD: hpre [rows, N];   TN with BGELU: unused
  float* aux_out;      // EPI_BIAS_GELU_SAVE: hpre out [rows, N] or null
  const int32_t* offsets;
  const int32_t* counts;
  const int32_t* tile_expert;
  int pad;
  int M, N, Kdim;      // rows-grouped: M unused;  TN: M x N output, K = rows of expert
  int E;
  // expert dropout behind the GELU (philox.cuh): the keep-scale m(row, col) in {0, 1/(1-p)} is a pure function of the
  // element's queue coordinates, regenerated wherever gelu(z) or gelu'(z) is formed (fc1 epilogue, dgelu epilogue, the
  // B operand of dW2) - only z is stored
  uint32_t drop_thr;   // p * 2^32 (0 = no dropout)
  float drop_inv_keep;
  const RngState* rng;
};

template <int LAY, int EPI, bool BGELU>
__global__ void __launch_bounds__(SG_THREADS) sgemm_grouped_kernel(SgemmParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;

  int e, m0, n0, kbeg, kend, kvalid_end;
  if (LAY == LAY_TN) {
    e = blockIdx.z;
    m0 = blockIdx.x * BM;
    n0 = blockIdx.y * BN;
    kbeg = p.offsets[e];
    kvalid_end = kbeg + p.counts[e];
    kend = kbeg + (p.counts[e] + BK - 1) / BK * BK;
  } else {
    m0 = blockIdx.x * BM;
    if (m0 >= p.offsets[p.E]) return;
    e = p.tile_expert[m0 / p.pad];
    n0 = blockIdx.y * BN;
    kbeg = 0;
    kend = p.Kdim;
    kvalid_end = kend;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* Bexp = (LAY == LAY_NT)   ? p.B + (int64_t)e * p.N * p.Kdim
                      : (LAY == LAY_NN) ? p.B + (int64_t)e * p.Kdim * p.N
                                        : p.B;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (LAY == LAY_TN) {
      const int k = tid / 16, i4 = tid % 16;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < kvalid_end) v = *reinterpret_cast<const float4*>(p.A + (int64_t)(k0 + k) * p.M + m0 + i4 * 4);
      *reinterpret_cast<float4*>(&As[k][i4 * 4]) = v;
    } else {
      const int m = tid / 4, k4 = tid % 4;
      float4 v = *reinterpret_cast<const float4*>(p.A + (int64_t)(m0 + m) * p.Kdim + k0 + k4 * 4);
      As[k4 * 4 + 0][m] = v.x; As[k4 * 4 + 1][m] = v.y; As[k4 * 4 + 2][m] = v.z; As[k4 * 4 + 3][m] = v.w;
    }
    // ---- B tile -> Bs[k][n]
    if (LAY == LAY_NT) {
      const int n = tid / 4, k4 = tid % 4;
      float4 v = *reinterpret_cast<const float4*>(Bexp + (int64_t)(n0 + n) * p.Kdim + k0 + k4 * 4);
      Bs[k4 * 4 + 0][n] = v.x; Bs[k4 * 4 + 1][n] = v.y; Bs[k4 * 4 + 2][n] = v.z; Bs[k4 * 4 + 3][n] = v.w;
    } else {
      const int k = tid / 16, n4 = tid % 16;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (LAY == LAY_NN || k0 + k < kvalid_end)
        v = *reinterpret_cast<const float4*>(Bexp + (int64_t)(k0 + k) * p.N + n0 + n4 * 4);
      if (BGELU) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        if (p.drop_thr != 0u) {
          float sc[4];
          dropout_scale4(*p.rng, (uint32_t)(k0 + k), (uint32_t)((n0 + n4 * 4) / 4), p.drop_thr, p.drop_inv_keep, sc);
          v.x *= sc[0]; v.y *= sc[1]; v.z *= sc[2]; v.w *= sc[3];
        }
      }
      if (BGELU && !(k0 + k < kvalid_end)) v = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&Bs[k][n4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU_SAVE) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + (int64_t)e * p.N + n0 + tx * 4);
    bias[0] = b.x; bias[1] = b.y; bias[2] = b.z; bias[3] = b.w;
  }
  float* Cbase = (LAY == LAY_TN) ? p.C + (int64_t)e * p.M * p.N : p.C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t off = (int64_t)(m0 + ty * 4 + i) * p.N + n0 + tx * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    float sc[4] = {1.f, 1.f, 1.f, 1.f};
    if ((EPI == EPI_BIAS_GELU_SAVE || EPI == EPI_GELU_GRAD) && p.drop_thr != 0u)
      dropout_scale4(*p.rng, (uint32_t)(m0 + ty * 4 + i), (uint32_t)((n0 + tx * 4) / 4), p.drop_thr, p.drop_inv_keep, sc);
    if (EPI == EPI_BIAS_GELU_SAVE) {
      if (p.aux_out != nullptr) *reinterpret_cast<float4*>(p.aux_out + off) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]) * sc[j];
    }
    if (EPI == EPI_GELU_GRAD) {
      const float4 h = *reinterpret_cast<const float4*>(p.aux + off);
      v[0] *= gelu_erf_grad(h.x) * sc[0]; v[1] *= gelu_erf_grad(h.y) * sc[1];
      v[2] *= gelu_erf_grad(h.z) * sc[2]; v[3] *= gelu_erf_grad(h.w) * sc[3];
    }
    *reinterpret_cast<float4*>(Cbase + off) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// db[e][n] = sum over the expert's valid rows of G[row][n]
__global__ void colsum_grouped_kernel(const float* __restrict__ G, const int32_t* __restrict__ offsets,
                                      const int32_t* __restrict__ counts, int N, float* __restrict__ db) {
  const int e = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int r0 = offsets[e], r1 = r0 + counts[e];
  float a = 0.f;
  for (int r = r0; r < r1; ++r) a += G[(int64_t)r * N + n];
  db[(int64_t)e * N + n] = a;
}

}  // namespace m3

using namespace m3;

int m3_ffn_fwd_f32(const float* xq, const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D,
                   int H, const float* w1, const float* b1, const float* w2, const float* b2, float* hpre,
                   float* yq, void* workspace, size_t workspace_bytes, float drop_p, const void* rng, cudaStream_t st) {
  M3_CHECK_SHAPE(D % BN == 0 && H % BN == 0 && cap_rows % BM == 0);
  if (drop_p > 0.f && (hpre == nullptr || rng == nullptr)) return M3_ERR_ARG;      // dropout is a training-time op
  if (workspace_bytes < (size_t)cap_rows * H * sizeof(float)) return M3_ERR_WORKSPACE;
  float* h = static_cast<float*>(workspace);
  SgemmParams p{};
  p.offsets = offsets; p.tile_expert = tile_expert; p.pad = M3_PAD_ROWS; p.E = E;
  // fc1: h = gelu(xq W1^T + b1)
  p.A = xq; p.B = w1; p.C = h; p.bias = b1; p.aux_out = hpre; p.N = H; p.Kdim = D;
  if (drop_p > 0.f) {
    p.drop_thr = dropout_threshold(drop_p);
    p.drop_inv_keep = 1.0f / (1.0f - drop_p);
    p.rng = static_cast<const RngState*>(rng);
  }
  sgemm_grouped_kernel<LAY_NT, EPI_BIAS_GELU_SAVE, false><<<dim3(cap_rows / BM, H / BN), SG_THREADS, 0, st>>>(p);
  M3_LAUNCH_CHECK();
  // fc2: yq = h W2^T + b2
  p.A = h; p.B = w2; p.C = yq; p.bias = b2; p.aux_out = nullptr; p.N = D; p.Kdim = H;
  sgemm_grouped_kernel<LAY_NT, EPI_BIAS, false><<<dim3(cap_rows / BM, D / BN), SG_THREADS, 0, st>>>(p);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

int m3_ffn_bwd_f32(const float* xq, const float* hpre, const float* dyq, const int32_t* counts,
                   const int32_t* offsets, const int32_t* tile_expert, int cap_rows, int E, int D, int H,
                   const float* w1, const float* w2, float* dxq, float* dw1, float* db1, float* dw2, float* db2,
                   void* workspace, size_t workspace_bytes, float drop_p, const void* rng, int parts, cudaStream_t st) {
  M3_CHECK_SHAPE(D % BN == 0 && H % BN == 0 && cap_rows % BM == 0);
  if (drop_p > 0.f && rng == nullptr) return M3_ERR_ARG;
  if (workspace_bytes < (size_t)cap_rows * H * sizeof(float)) return M3_ERR_WORKSPACE;
  float* dhpre = static_cast<float*>(workspace);
  SgemmParams p{};
  p.offsets = offsets; p.counts = counts; p.tile_expert = tile_expert; p.pad = M3_PAD_ROWS; p.E = E;
  if (drop_p > 0.f) {      // the forward's mask again, from the same (seed, counter) the caller kept
    p.drop_thr = dropout_threshold(drop_p);
    p.drop_inv_keep = 1.0f / (1.0f - drop_p);
    p.rng = static_cast<const RngState*>(rng);
  }
  if (parts & 1) {      // data gradients (dhpre stays in the workspace for the weight-gradient part)
    // dhpre = (dyq W2) * gelu'(hpre)                     [rows, D] x [D, H]
    p.A = dyq; p.B = w2; p.C = dhpre; p.aux = hpre; p.N = H; p.Kdim = D;
    sgemm_grouped_kernel<LAY_NN, EPI_GELU_GRAD, false><<<dim3(cap_rows / BM, H / BN), SG_THREADS, 0, st>>>(p);
    M3_LAUNCH_CHECK();
    // dxq = dhpre W1                                     [rows, H] x [H, D]
    p.A = dhpre; p.B = w1; p.C = dxq; p.aux = nullptr; p.N = D; p.Kdim = H;
    sgemm_grouped_kernel<LAY_NN, EPI_NONE, false><<<dim3(cap_rows / BM, D / BN), SG_THREADS, 0, st>>>(p);
    M3_LAUNCH_CHECK();
  }
  if (!(parts & 2)) return M3_OK;
  // dW2[e] = dyq_e^T gelu(hpre_e)                      [D, H]
  p.A = dyq; p.B = hpre; p.C = dw2; p.M = D; p.N = H;
  sgemm_grouped_kernel<LAY_TN, EPI_NONE, true><<<dim3(D / BM, H / BN, E), SG_THREADS, 0, st>>>(p);
  M3_LAUNCH_CHECK();
  // dW1[e] = dhpre_e^T xq_e                            [H, D]
  p.A = dhpre; p.B = xq; p.C = dw1; p.M = H; p.N = D;
  sgemm_grouped_kernel<LAY_TN, EPI_NONE, false><<<dim3(H / BM, D / BN, E), SG_THREADS, 0, st>>>(p);
  M3_LAUNCH_CHECK();
  colsum_grouped_kernel<<<dim3(m3_ceil_div(D, 128), E), 128, 0, st>>>(dyq, offsets, counts, D, db2);
  M3_LAUNCH_CHECK();
  colsum_grouped_kernel<<<dim3(m3_ceil_div(H, 128), E), 128, 0, st>>>(dhpre, offsets, counts, H, db1);
  M3_LAUNCH_CHECK();
  return M3_OK;
}

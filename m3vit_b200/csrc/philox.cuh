// Counter-based random numbers for the stochastic parts of the layer: the router noise of
// NoisyGate_VMoE (/root/reference/models/moe/origin/noisy_gate_vmoe.py:226, torch.randn_like) and the expert dropout
// inside the FFN (nn.Dropout after the GELU, models/moe/origin/vision_transformer_moe.py:248-251, drop_rate 0.1 in
// configs/nyud/vit_moe/*drop0.1*.yml).  Philox4x32 (Salmon et al., SC'11) with 7 rounds - the smallest round count
// that passes BigCrush - keyed by a 64-bit seed and counted by (row, column group, call counter): every element's
// random bits are a pure function of its coordinates, so the backward pass regenerates the forward's dropout mask
// without storing it and nothing depends on the launch geometry.  The streams are NOT torch's: parity with the
// reference is statistical (tests/test_gpu_stochastic.py), never bit-wise.
#pragma once

#include <stdint.h>

namespace m3 {

struct RngState {        // device memory, 16 bytes: {seed, per-call counter}; bumped by the host side between calls
  unsigned long long seed;
  unsigned long long counter;
};

__device__ __forceinline__ uint4 philox4x32_7(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// 4 x 32 random bits for elements (row, 4*quad .. 4*quad+3) of stream `stream` (0 = dropout mask, 1 = router noise)
__device__ __forceinline__ uint4 rng_bits4(const RngState& s, uint32_t stream, uint32_t row, uint32_t quad) {
  const uint2 key = make_uint2((uint32_t)s.seed, (uint32_t)(s.seed >> 32) ^ (stream * 0x85EBCA6Bu));
  return philox4x32_7(make_uint4(row, quad, (uint32_t)s.counter, (uint32_t)(s.counter >> 32)), key);
}

// keep-scale factors of 4 consecutive columns: 1/(1-p) with probability 1-p, else 0   (thr = p * 2^32)
__device__ __forceinline__ void dropout_scale4(const RngState& s, uint32_t row, uint32_t quad, uint32_t thr, float inv_keep,
                                               float out[4]) {
  const uint4 b = rng_bits4(s, 0u, row, quad);
  out[0] = b.x >= thr ? inv_keep : 0.f;
  out[1] = b.y >= thr ? inv_keep : 0.f;
  out[2] = b.z >= thr ? inv_keep : 0.f;
  out[3] = b.w >= thr ? inv_keep : 0.f;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

// 4 standard normals (Box-Muller on two pairs of uniforms in (0, 1])
__device__ __forceinline__ void normal4(const RngState& s, uint32_t row, uint32_t quad, float out[4]) {
  const uint4 b = rng_bits4(s, 1u, row, quad);
  const float u0 = ((float)(b.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u1 = (float)(b.y >> 8) * (1.0f / 16777216.0f);
  const float u2 = ((float)(b.z >> 8) + 1.0f) * (1.0f / 16777216.0f), u3 = (float)(b.w >> 8) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  out[0] = r0 * c0; out[1] = r0 * s0; out[2] = r1 * c1; out[3] = r1 * s1;
}

}  // namespace m3

// placeholder until the tcgen05 path lands (replaced in the next commit)
#include "common.cuh"
size_t m3_ffn_bf16_workspace_bytes(int, int, int, int, int) { return 0; }
int m3_ffn_fwd_bf16(const void*, const int32_t*, const int32_t*, int, int, int, int, const void*, const float*,
                    const void*, const float*, void*, void*, void*, size_t, cudaStream_t) { return M3_ERR_UNSUPPORTED; }
int m3_ffn_bwd_bf16(const void*, const void*, const void*, const int32_t*, const int32_t*, const int32_t*, int, int,
                    int, int, const void*, const void*, const void*, const void*, void*, float*, float*, float*,
                    float*, void*, size_t, cudaStream_t) { return M3_ERR_UNSUPPORTED; }

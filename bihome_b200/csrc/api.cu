// Library-level entry points: version, error text, launch counter, MACE.
#include <string.h>

#include "bh_common.cuh"

namespace bh {
unsigned long long g_launch_count = 0;
int g_tune[kTuneCount] = {0};

// train.py:401-404 / eval.py:133-134: mean over B*4 corners of the Euclidean corner error
__global__ void __launch_bounds__(256) mace_kernel(const float* __restrict__ gt, const float* __restrict__ hat,
                                                   float* __restrict__ out, int n_corners) {
    __shared__ float red[8];
    float s[1] = {0.0f};
    for (int i = threadIdx.x; i < n_corners; i += blockDim.x) {
        const float dx = gt[2 * i] - hat[2 * i], dy = gt[2 * i + 1] - hat[2 * i + 1];
        s[0] += sqrtf(fmaf(dx, dx, dy * dy));
    }
    block_sum<1>(s, red);
    if (threadIdx.x == 0) out[0] = s[0] / static_cast<float>(n_corners);
}
}  // namespace bh

extern "C" int bh_version(void) { return BH_VERSION; }

extern "C" unsigned long long bh_launch_count(void) { return bh::g_launch_count; }

extern "C" const char* bh_strerror(int code) {
    switch (code) {
        case BH_OK: return "ok";
        case BH_E_NULL: return "bihome_b200: required pointer is NULL";
        case BH_E_SHAPE: return "bihome_b200: non-positive or unsupported dimension";
        case BH_E_ALIGN: return "bihome_b200: pointer is not 16-byte aligned";
        case BH_E_WORKSPACE: return "bihome_b200: workspace missing or too small";
        case BH_E_UNSUPPORTED: return "bihome_b200: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "bihome_b200: unknown error code";
}

extern "C" int bh_tune_set(const char* key, int value) {
    if (!key) return BH_E_NULL;
    static const char* const names[] = {"warp_path", "loss_variant", "loss_cluster", "warp_variant", "fieldhead_variant"};
    for (int k = 0; k < 5; ++k)
        if (strcmp(key, names[k]) == 0) {
            bh::g_tune[k] = value;
            return BH_OK;
        }
    return BH_E_UNSUPPORTED;
}

extern "C" int bh_mace(const float* delta_gt, const float* delta_hat, float* out, int B, bh_stream_t stream) {
    if (!delta_gt || !delta_hat || !out) return BH_E_NULL;
    if (B <= 0) return BH_E_SHAPE;
    bh::mace_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(delta_gt, delta_hat, out, B * 4);
    return bh::launch_status();
}

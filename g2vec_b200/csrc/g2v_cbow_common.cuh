// g2v_cbow_common.cuh -- device helpers shared by the CBOW kernels (g2v_cbow.cu, g2v_cbow_slab.cu).
#pragma once
#include "g2v_common.cuh"

namespace g2v {

constexpr int kCbowWarps = 8;

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ void red_add4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float sigmoid_stable(float x) {
    if (x >= 0.f) { const float z = expf(-x); return 1.f / (1.f + z); }
    const float z = expf(x);
    return z / (1.f + z);
}

struct CtaAcc {   // per-CTA accumulators in shared memory
    double loss;
    unsigned long long correct;
};

// grid of a one-warp-per-window kernel: whole chip resident (SMs x occupancy), never more CTAs than windows
int rows_grid(const void *kernel, size_t smem, int64_t n_win, int *grid_out);

}  // namespace g2v

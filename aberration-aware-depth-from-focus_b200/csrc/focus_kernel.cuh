// select_focus_dist(depth, num, mode='linear') on the device (dff/utils.py:4-51 of the reference):
// per image, min over valid depths (> 0) and max over all depths, then num focus distances
// depth_min + i * (depth_max - depth_min) / (num - 1).  One CTA per image, no host round trip
// (the reference loops over the batch in Python and indexes with a boolean mask).
#pragma once
#include <cuda_runtime.h>

namespace aadff {

constexpr int FOCUS_NT = 1024;

__global__ void __launch_bounds__(FOCUS_NT)
select_focus_kernel(const float* __restrict__ depth, long long hw, int num, float* __restrict__ out) {
    __shared__ float s_min[FOCUS_NT / 32], s_max[FOCUS_NT / 32];
    const float* d = depth + (long long)blockIdx.x * hw;
    float vmin = __int_as_float(0x7f800000), vmax = -__int_as_float(0x7f800000);   // +inf, -inf
    for (long long i = threadIdx.x; i < hw; i += FOCUS_NT) {
        const float v = __ldg(d + i);
        vmax = fmaxf(vmax, v);
        if (v > 0.f) vmin = fminf(vmin, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = vmin; s_max[threadIdx.x >> 5] = vmax; }
    __syncthreads();
    if (threadIdx.x < 32) {
        vmin = s_min[threadIdx.x];
        vmax = s_max[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        // reference arithmetic order: depth_min + (i * (depth_max - depth_min)) / (num - 1)
        // An image without a single valid (> 0) depth: the reference raises (torch.min of an empty selection,
        // dff/utils.py:22) and its training loop skips such batches (2_aber_aware_dff_aif.py:103-105).  Here every
        // focus distance of that image is a quiet NaN -- an explicit sentinel the caller can test without a sync.
        const bool none_valid = !(vmin < __int_as_float(0x7f800000));
        const float span = vmax - vmin;
        for (int i = threadIdx.x; i < num; i += 32)
            out[(long long)blockIdx.x * num + i] =
                none_valid ? __int_as_float(0x7fc00000) : vmin + __fdiv_rn((float)i * span, (float)(num - 1));
    }
}

}  // namespace aadff

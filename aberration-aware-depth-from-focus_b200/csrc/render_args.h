// Arguments shared by the render kernels + the network-input helpers that restate
// deeplens/psfnet.py:426-437 (coordinate grid, depth2z) in device code.
#pragma once
#include <cuda_runtime.h>

namespace aadff {

constexpr int MAX_LAYERS = 16;

struct RenderArgs {
    const float* img;     // [N,C,H,W] fp32, contiguous
    const float* depth;   // [N,H,W]   fp32, mm (<= 0)
    const float* foc;     // [N,S]     fp32, mm (< 0); row stride foc_stride (a slice window of a wider array)
    float* out;           // element strides below
    long long os_n, os_c, os_s, os_h, os_w;
    int N, C, S, H, W, ks;  // C = channels rendered by this launch (<= 4)
    int Ctot, c0;           // img/out have Ctot channels; this launch covers [c0, c0+C)
    int foc_stride;         // elements between consecutive images in foc (>= S)
    float d_min, d_range; // d_min = -200, d_range = d_max - d_min = -19800
    float step_x, step_y; // torch.linspace steps: 2/(W-1), -2/(H-1)
};

// torch.linspace(-1, 1, W)[w]: fma(step, w, start) in the first half, fma(-step, W-1-w, end) after
__host__ __device__ __forceinline__ float coord_x(int w, int W, float step) {
    if (W == 1) return -1.0f;                                   // linspace(a, b, 1) = [a]
    return (w < W / 2) ? fmaf(step, (float)w, -1.0f) : fmaf(-step, (float)(W - 1 - w), 1.0f);
}
// torch.linspace(1, -1, H)[h]
__host__ __device__ __forceinline__ float coord_y(int h, int H, float step) {
    if (H == 1) return 1.0f;
    return (h < H / 2) ? fmaf(step, (float)h, 1.0f) : fmaf(-step, (float)(H - 1 - h), -1.0f);
}
// PSFNet.depth2z (deeplens/psfnet.py:447-450): clamp((d - d_min) / (d_max - d_min), 0, 1), IEEE division
__device__ __forceinline__ float depth_to_z(float d, float d_min, float d_range) {
    const float z = __fdiv_rn(d - d_min, d_range);
    return fminf(fmaxf(z, 0.0f), 1.0f);
}

}  // namespace aadff

// AutoAgument's rotation on the device: scipy.ndimage.rotate(x, degree, reshape=False) with its defaults order = 3,
// mode = 'constant', cval = 0, prefilter = True (dff/dataset.py:275-284), i.e. scipy.ndimage.affine_transform on
// cubic B-spline coefficients.  Three steps on fp32 planes [B, P, H, W] (P = RGB planes and/or the depth plane):
//
//   1. prepare_planes_kernel   uint8 BGR / uint16 depth -> fp32 planes at full resolution with AutoAgument's colour
//                              jitter and flips applied (the same 256-entry table as preprocess_rgbd_kernel);
//   2. spline_prefilter_kernel the B-spline prefilter.  scipy runs the recursive filter  c+[i] = x[i] + z c+[i-1],
//                              c[i] = z (c[i+1] - c+[i]),  z = sqrt(3) - 2, gain 6, with "mirror" initial values, in
//                              float64.  With mirror boundaries that recursion IS the two-sided sum
//                              c[i] = sqrt(3) * sum_k z^|k| x~[i+k] over the mirror-extended signal x~ (checked against
//                              scipy.ndimage.spline_filter to 3e-15 by the tests), and z^17 = 2e-10:
//                              a 33-tap symmetric FIR per axis is exact to fp32 rounding and has no serial dependence;
//   3. spline_affine_kernel    per output pixel: input coordinate = M o + offset in double, in scipy's operation order
//                              (the in/out-of-image decision `0 <= c <= n-1` must not flip at the border); outside ->
//                              0; inside -> 4 x 4 coefficients (indices mirrored like scipy's edge offsets) with the
//                              cubic weights of ni_interpolation.c; the depth plane is clamped at 0
//                              (`depth[depth<0] = 0`).  Samples whose matrix entry is NaN are copied (no rotation drawn).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace aadff {

constexpr int SPL_K = 16;                          // FIR half-width: z^17 = 1.9e-10
// sqrt(3) * z^k, z = sqrt(3) - 2
__constant__ float c_spline_taps[SPL_K + 1] = {
    1.7320508075688772f, -0.4641016151377547f, 0.12435565298214114f, -0.033320996790809666f, 0.008928334181097486f,
    -0.0023923399335802615f, 0.000641025553223557f, -0.00017176227931396584f, 4.602356403230609e-05f,
    -1.2331976815258487e-05f, 3.3043432287278414e-06f, -8.85396099652874e-07f, 2.3724116988365352e-07f,
    -6.356857988173978e-08f, 1.7033149643305495e-08f, -4.564018691482174e-09f, 1.2229251226231986e-09f};

// scipy's mirror extension of an index (ni_interpolation.c, NI_EXTEND_MIRROR): whole-sample symmetric, any distance
__device__ __forceinline__ int spline_mirror(int i, int n) {
    if (n <= 1) return 0;
    const int s2 = 2 * n - 2;
    if (i < 0) {
        i = s2 * (-i / s2) + i;
        i = (i <= 1 - n) ? i + s2 : -i;
    } else if (i >= n) {
        i -= s2 * (i / s2);
        if (i >= n) i = s2 - i;
    }
    return i;
}

struct PlanesArgs {
    const uint8_t* bgr;       // [B,H,W,3] or null
    const uint16_t* depth;    // [B,H,W] or null
    float* planes;            // [B,P,H,W]: RGB planes first (if bgr), then the depth plane (if depth)
    const float* jitter;      // [B,2] or null (contrast < 0 = none)
    const uint8_t* flips;     // [B] or null
    int B, H, W, P;
    float depth_div;
};

__global__ void __launch_bounds__(256) prepare_planes_kernel(const PlanesArgs a) {
    __shared__ float lut[256];
    const int b = blockIdx.y;
    {
        double con = -1.0, bri = 0.0;
        if (a.jitter) { con = (double)a.jitter[2 * b]; bri = (double)a.jitter[2 * b + 1]; }
        double v = (double)threadIdx.x / 255.0;
        if (con >= 0.0) v = fmin(fmax(0.5 + con * (v - 0.5) + bri, 0.0), 1.0);
        lut[threadIdx.x] = (float)v;
    }
    __syncthreads();
    const int fl = a.flips ? a.flips[b] : 0;
    const bool flx = fl & 1, fly = fl & 2;
    const long long total = (long long)a.H * a.W;
    float* dst = a.planes + (long long)b * a.P * total;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(id % a.W), y = (int)(id / a.W);
        const int sx = flx ? a.W - 1 - x : x, sy = fly ? a.H - 1 - y : y;
        const long long s = ((long long)b * a.H + sy) * a.W + sx;
        int p = 0;
        if (a.bgr) {
            const uint8_t* px = a.bgr + s * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) dst[(long long)c * total + id] = lut[px[2 - c]];        // BGR -> RGB, /255, jitter
            p = 3;
        }
        if (a.depth) dst[(long long)p * total + id] = (float)a.depth[s] / a.depth_div;
    }
}

// one axis of the prefilter: AXIS 1 = along W, 0 = along H.  One thread per element; neighbouring threads read
// neighbouring addresses in both cases, the 33 taps come through L1.
template <int AXIS>
__global__ void __launch_bounds__(256) spline_prefilter_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              long long planes, int H, int W) {
    const long long total = planes * H * W;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(id % W), y = (int)((id / W) % H);
        const long long base = id - (AXIS == 1 ? x : (long long)y * W);
        const int n = AXIS == 1 ? W : H, i = AXIS == 1 ? x : y;
        const long long stride = AXIS == 1 ? 1 : W;
        float acc = c_spline_taps[0] * src[id];
        if (i >= SPL_K && i + SPL_K < n) {
#pragma unroll
            for (int k = 1; k <= SPL_K; ++k)
                acc = fmaf(c_spline_taps[k], src[base + (long long)(i - k) * stride] + src[base + (long long)(i + k) * stride], acc);
        } else {
#pragma unroll 1
            for (int k = 1; k <= SPL_K; ++k)
                acc = fmaf(c_spline_taps[k], src[base + (long long)spline_mirror(i - k, n) * stride] +
                                                 src[base + (long long)spline_mirror(i + k, n) * stride], acc);
        }
        dst[id] = acc;
    }
}

struct AffineArgs {
    const float* coef;        // [B,P,H,W] prefiltered
    const float* raw;         // [B,P,H,W] un-filtered (copied where a sample is not transformed)
    float* out;               // [B,P,H,W]
    const double* xform;      // [B,6] = m00, m01, m10, m11, off0, off1 (input row/col = M (out row, out col) + off); m00 NaN = copy
    int B, P, H, W;
    int clamp_plane;          // plane index clamped at >= 0 after the transform (the depth plane), or -1
};

// cubic B-spline weights, ni_interpolation.c get_spline_interpolation_weights(order 3): t = x - floor(x)
__device__ __forceinline__ void spline_weights3(float t, float (&w)[4]) {
    const float u = 1.f - t;
    w[1] = (t * t * (t - 2.f) * 3.f + 4.f) / 6.f;
    w[2] = (u * u * (u - 2.f) * 3.f + 4.f) / 6.f;
    w[0] = u * u * u / 6.f;
    w[3] = 1.f - w[0] - w[1] - w[2];
}

__global__ void __launch_bounds__(256) spline_affine_kernel(const AffineArgs a) {
    const int b = blockIdx.y;
    const long long total = (long long)a.H * a.W;
    const double* xf = a.xform + 6 * b;
    const double m00 = xf[0], m01 = xf[1], m10 = xf[2], m11 = xf[3], o0 = xf[4], o1 = xf[5];
    const bool copy = (m00 != m00);
    const float* coef = a.coef + (long long)b * a.P * total;
    const float* raw = a.raw + (long long)b * a.P * total;
    float* out = a.out + (long long)b * a.P * total;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        if (copy) {
            for (int p = 0; p < a.P; ++p) out[(long long)p * total + id] = raw[(long long)p * total + id];
            continue;
        }
        const int q = (int)(id % a.W), r = (int)(id / a.W);
        // scipy: cc = shift; cc += matrix[.][0] * row; cc += matrix[.][1] * col   (double, no contraction)
        const double yy = __dadd_rn(__dadd_rn(o0, __dmul_rn(m00, (double)r)), __dmul_rn(m01, (double)q));
        const double xx = __dadd_rn(__dadd_rn(o1, __dmul_rn(m10, (double)r)), __dmul_rn(m11, (double)q));
        if (yy < 0.0 || yy > (double)(a.H - 1) || xx < 0.0 || xx > (double)(a.W - 1)) {          // mode='constant', cval=0
            for (int p = 0; p < a.P; ++p) out[(long long)p * total + id] = 0.f;
            continue;
        }
        const double fy = floor(yy), fx = floor(xx);
        float wy[4], wx[4];
        spline_weights3((float)(yy - fy), wy);
        spline_weights3((float)(xx - fx), wx);
        const int sy = (int)fy - 1, sx = (int)fx - 1;
        int iy[4], ix[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            iy[t] = spline_mirror(sy + t, a.H);
            ix[t] = spline_mirror(sx + t, a.W);
        }
        for (int p = 0; p < a.P; ++p) {
            const float* cp = coef + (long long)p * total;
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float* row = cp + (long long)iy[t] * a.W;
                const float rsum = wx[0] * row[ix[0]] + wx[1] * row[ix[1]] + wx[2] * row[ix[2]] + wx[3] * row[ix[3]];
                acc = fmaf(wy[t], rsum, acc);
            }
            if (p == a.clamp_plane) acc = fmaxf(acc, 0.f);
            out[(long long)p * total + id] = acc;
        }
    }
}

}  // namespace aadff

// Fused thin-lens render: the CUDA counterpart of ThinLens.render + ThinLens.coc
// (deeplens/psfnet.py:503-570).  Per output pixel: circle of confusion from depth and focus
// distance -> clipped Gaussian k x k PSF (sigma = coc/2, taps outside the coc radius are zero)
// -> normalise -> gather (deeplens/render_psf.py:76-107).  The reference materialises the
// [N,H,W,k,k] PSF tensor through ~15 elementwise torch kernels; here taps live in registers.
//
// CTA tile = 8 rows x 32 columns, replicate-clamped image halo in shared memory, one thread
// per pixel.  Bound by shared-memory reads / issue (3 LDS + ~9 ALU per tap), far above the
// 28 B/pixel HBM traffic.
#pragma once
#include <cuda_runtime.h>

namespace aadff {

constexpr int TL_TILE_H = 8;
constexpr int TL_TILE_W = 32;
constexpr int TL_MAXC = 4;

struct ThinLensArgs {
    const float* img;     // [N,C,H,W]
    const float* depth;   // [N,H,W] mm (sign as given by the caller)
    const float* foc;     // [N] mm
    float* out;           // [N,C,H,W]
    int N, C, H, W, ks, c0, cn;
    float k1;             // foc_len / fnum
    float foc_len, ps;    // focal length [mm], pixel size [mm]
    float d_lo, d_hi;     // depth clamp: 200, 20000 mm
    int flip;             // 1: depth and foc are negated first (reference: `if (depth < 0).any()`)
    const unsigned char* flip_dev;   // if not null, the decision is read from device memory (see any_negative_kernel):
                                     // the reference's data-dependent branch without a device->host round trip
};

// flag[0] |= any(x < 0)   (the reference's `if (depth < 0).any()`, psfnet.py:504, decided on the device)
__global__ void __launch_bounds__(256) any_negative_kernel(const float* __restrict__ x, long long n, unsigned char* flag) {
    bool neg = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        neg |= (__ldg(x + i) < 0.f);
    if (__syncthreads_or(neg) && threadIdx.x == 0) *flag = 1;     // benign race: every writer stores 1
}

__global__ void __launch_bounds__(TL_TILE_H * 32)
thinlens_render_kernel(ThinLensArgs a) {
    extern __shared__ float tl_img[];            // [cn][HH][pitch]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ks = a.ks, r = (ks - 1) / 2;
    const int HH = TL_TILE_H + ks - 1, HW = TL_TILE_W + ks - 1;
    const int pitch = HW | 1;
    const int cstride = HH * pitch;
    const int tiles_x = (a.W + TL_TILE_W - 1) / TL_TILE_W, tiles_y = (a.H + TL_TILE_H - 1) / TL_TILE_H;
    const long long n_tiles = (long long)a.N * tiles_x * tiles_y;
    const bool flip = a.flip_dev ? (*a.flip_dev != 0) : (a.flip != 0);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
        const int n = (int)(tile / ((long long)tiles_x * tiles_y));
        const int h0 = ty * TL_TILE_H, w0 = tx * TL_TILE_W;
        const int h = h0 + warp, w = w0 + lane;
        const bool ok = (h < a.H) && (w < a.W);
        // circle of confusion (ThinLens.coc, psfnet.py:503-511), operation order of the reference
        float inv_2r2_log2e = 0.f, r2 = 0.f;
        if (ok) {
            float d = __ldg(a.depth + ((long long)n * a.H + h) * a.W + w);
            float f = __ldg(a.foc + n);
            if (flip) { d = -d; f = -f; }
            d = fminf(fmaxf(d, a.d_lo), a.d_hi);
            float coc = a.k1 * fabsf(d - f);
            coc = __fdiv_rn(coc, d);
            coc = coc * a.foc_len;
            coc = __fdiv_rn(coc, f - a.foc_len);
            const float coc_px = fmaxf(__fdiv_rn(coc, a.ps), 0.1f);
            const float rad = coc_px * 0.5f;
            r2 = rad * rad;
            inv_2r2_log2e = __fdiv_rn(-0.5f * 1.4426950408889634f, r2);    // exp(-d2/2/r2) = 2^(d2 * this)
        }
        __syncthreads();
        {
            const int total = a.cn * HH * HW;
            for (int base = threadIdx.x; base < total; base += 4 * TL_TILE_H * 32) {
                float v[4];
                int slot[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + u * TL_TILE_H * 32;
                    slot[u] = -1;
                    if (idx < total) {
                        const int c = idx / (HH * HW), rem = idx - c * (HH * HW);
                        const int yy = rem / HW, xx = rem - yy * HW;
                        const int gy = min(max(h0 + yy - r, 0), a.H - 1), gx = min(max(w0 + xx - r, 0), a.W - 1);
                        v[u] = __ldg(a.img + ((long long)(n * a.C + a.c0 + c) * a.H + gy) * a.W + gx);
                        slot[u] = (c * HH + yy) * pitch + xx;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (slot[u] >= 0) tl_img[slot[u]] = v[u];
            }
        }
        __syncthreads();
        if (!ok) continue;
        float acc[TL_MAXC] = {0.f, 0.f, 0.f, 0.f}, wsum = 0.f;
        const float* ib = tl_img + warp * pitch + lane;
        for (int i = 0; i < ks; ++i) {
            const int dy2 = (i - r) * (i - r);
            const float* prow = ib + i * pitch;
#pragma unroll 4
            for (int j = 0; j < ks; ++j) {
                const float d2 = (float)(dy2 + (j - r) * (j - r));
                float wt = exp2f(d2 * inv_2r2_log2e);
                wt = (d2 < r2) ? wt : 0.f;                 // psf_mask = (x^2 + y^2 < radius^2)
                wsum += wt;
                acc[0] = fmaf(prow[j], wt, acc[0]);
                if (a.cn > 1) acc[1] = fmaf(prow[j + cstride], wt, acc[1]);
                if (a.cn > 2) acc[2] = fmaf(prow[j + 2 * cstride], wt, acc[2]);
                if (a.cn > 3) acc[3] = fmaf(prow[j + 3 * cstride], wt, acc[3]);
            }
        }
        const float inv = __fdiv_rn(1.0f, wsum);            // the centre tap always passes the mask: wsum >= 1
#pragma unroll
        for (int c = 0; c < TL_MAXC; ++c)
            if (c < a.cn) a.out[((long long)(n * a.C + a.c0 + c) * a.H + h) * a.W + w] = acc[c] * inv;
    }
}

}  // namespace aadff

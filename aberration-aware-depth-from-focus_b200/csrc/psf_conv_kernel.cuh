// Spatially-(piecewise-)invariant PSF convolution: the CUDA counterpart of
//   render_psf      (deeplens/render_psf.py:12-28)  one [C,ks,ks] PSF for the whole image, and
//   render_psf_map  (deeplens/render_psf.py:31-73)  a grid x grid map of PSFs, one per image patch,
// both of which the reference evaluates as grouped conv2d on a reflect-padded image with the PSF flipped
// (a true convolution; note that local_psf_render, the per-pixel gather, does NOT flip).  For every output pixel
//
//   out[b,c,h,w] = sum_{i,j < ks} img[b,c,refl(h+i-p),refl(w+j-p)] * psf[c][cell(h,w)][ks-1-i][ks-1-j],   p = (ks-1)/2
//
// where cell(h,w) is the patch (gi,gj) with hb[gi] <= h < hb[gi+1], wb[gj] <= w < wb[gj+1] and the bounds are the
// reference's int(i/grid*H) (computed on the host in Python arithmetic).  render_psf is grid = 1.
//
// CTA = 256 threads = a 32 x 64 output tile of one (image, channel, cell); a thread owns 8 consecutive pixels of a
// row and slides a (8+KS-1)-wide register window over the shared-memory halo tile: per tile row (KS+7) LDS for
// 8*KS FFMA.  Bound by the fp32 FFMA rate (KS^2 FMA per pixel against 8 B of HBM traffic), so tiles are sized for
// register reuse, not for HBM.
#pragma once
#include <cuda_runtime.h>

namespace aadff {

constexpr int PC_TILE_H = 32;
constexpr int PC_TILE_W = 64;
constexpr int PC_PX = 8;                 // pixels per thread
constexpr int PC_NT = 256;
constexpr int PC_MAX_GRID = 32;

struct PsfConvArgs {
    const float* img;        // [B,C,H,W]
    const float* psf_map;    // [C, grid*ks, grid*ks]
    float* out;              // [B,C,H,W]
    int B, C, H, W, grid;
    int hb[PC_MAX_GRID + 1], wb[PC_MAX_GRID + 1];    // patch bounds, hb[grid] = H, wb[grid] = W
    int max_ch, max_cw;      // largest patch height / width (tile counts per patch are derived from these)
};

__device__ __forceinline__ int reflect_index(int p, int n) {      // F.pad(mode='reflect'): -1 -> 1, n -> n-2
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * (n - 1) - p : p;
    return min(max(p, 0), n - 1);          // rows/columns of a tile that lie beyond the image (never used) stay in range
}

template <int KS>
__global__ void __launch_bounds__(PC_NT) psf_conv_kernel(const PsfConvArgs a) {
    constexpr int P = (KS - 1) / 2;
    constexpr int HH = PC_TILE_H + KS - 1, HW = PC_TILE_W + KS - 1;
    constexpr int PITCH = HW | 1;
    __shared__ float s_img[HH * PITCH];
    __shared__ float s_psf[KS * KS];
    const int tiles_y = (a.max_ch + PC_TILE_H - 1) / PC_TILE_H, tiles_x = (a.max_cw + PC_TILE_W - 1) / PC_TILE_W;
    const long long per_plane = (long long)a.grid * a.grid * tiles_y * tiles_x;
    const long long n_tiles = per_plane * a.B * a.C;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;            // 8 threads x 8 px = 64 columns, 32 rows
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int bc = (int)(t / per_plane);
        int rem = (int)(t % per_plane);
        const int cell = rem / (tiles_y * tiles_x);
        rem -= cell * tiles_y * tiles_x;
        const int gi = cell / a.grid, gj = cell % a.grid;
        const int h_lo = a.hb[gi], h_hi = a.hb[gi + 1], w_lo = a.wb[gj], w_hi = a.wb[gj + 1];
        const int h0 = h_lo + (rem / tiles_x) * PC_TILE_H, w0 = w_lo + (rem % tiles_x) * PC_TILE_W;
        if (h0 >= h_hi || w0 >= w_hi) continue;                       // this patch is smaller than the largest one
        const int c = bc % a.C;
        const float* plane = a.img + (long long)bc * a.H * a.W;
        __syncthreads();                                              // previous tile done with smem
        for (int i = threadIdx.x; i < KS * KS; i += PC_NT) {          // flipped PSF of this (channel, cell)
            const int pi = i / KS, pj = i - pi * KS;
            s_psf[i] = __ldg(a.psf_map + ((long long)c * a.grid * KS + gi * KS + (KS - 1 - pi)) * (a.grid * KS) +
                             gj * KS + (KS - 1 - pj));
        }
        for (int i = threadIdx.x; i < HH * HW; i += PC_NT) {
            const int yy = i / HW, xx = i - yy * HW;
            s_img[yy * PITCH + xx] = __ldg(plane + (long long)reflect_index(h0 + yy - P, a.H) * a.W +
                                           reflect_index(w0 + xx - P, a.W));
        }
        __syncthreads();
        float acc[PC_PX];
#pragma unroll
        for (int p = 0; p < PC_PX; ++p) acc[p] = 0.f;
        const float* base = s_img + ty * PITCH + tx * PC_PX;
#pragma unroll 1
        for (int i = 0; i < KS; ++i) {
            float v[PC_PX + KS - 1];
#pragma unroll
            for (int u = 0; u < PC_PX + KS - 1; ++u) v[u] = base[i * PITCH + u];
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                const float w = s_psf[i * KS + j];
#pragma unroll
                for (int p = 0; p < PC_PX; ++p) acc[p] = fmaf(v[p + j], w, acc[p]);
            }
        }
        const int h = h0 + ty;
        if (h < h_hi) {
            float* o = a.out + (long long)bc * a.H * a.W + (long long)h * a.W + w0 + tx * PC_PX;
#pragma unroll
            for (int p = 0; p < PC_PX; ++p)
                if (w0 + tx * PC_PX + p < w_hi) o[p] = acc[p];
        }
    }
}

}  // namespace aadff

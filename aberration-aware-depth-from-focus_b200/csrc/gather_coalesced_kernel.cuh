// Stand-alone per-pixel PSF gather, register-streaming version: the CUDA counterpart of
// deeplens/render_psf.py:76-107 (local_psf_render) for a PSF tensor in HBM ([N,H,W,ks,ks] fp32).
//
// The PSF read (4*ks^2 B per pixel) is the whole cost, so the taps never touch shared memory:
//   * G consecutive pixels of a row are one flat run of G*ks^2 floats.  A warp reads the run with
//     NL = ceil(G*ks^2/32) fully coalesced 128-byte loads (lane l takes element 32m + l) straight into
//     registers; the loads of the NEXT group are issued while the current one is consumed, so every warp
//     keeps NL*128 B (3.9 KB at ks = 11) in flight the whole time, across tile boundaries.
//   * Which pixel / tap element (m, lane) is depends only on (m, lane): the image-halo offset of every
//     element is computed once per kernel and kept in registers (16-bit byte offsets, two per register); pixel boundaries inside a load are
//     compile-time lane thresholds.
//   * The image halo of the CTA tile (16 rows x 32 columns) sits in shared memory, planar per channel, with
//     pitch = 32 + ks: consecutive taps -> consecutive banks even across a PSF row wrap (an interleaved
//     float4 halo -- one LDS.128 per tap -- measured 3-14 % slower: 4 instead of 3 wavefronts per tap).  It is double-
//     buffered: the next tile's halo arrives by cp.async (LDGSTS, no registers) under this tile's work, so a
//     tile costs one __syncthreads and no exposed L2 latency.
//   * Each lane owns partial sums of the G pixels; a transposing butterfly (G-1 + log2(32/G) shuffles
//     per channel) leaves pixel q, channel c in lane 4q + c (G = 8), which stores it.
// G is the largest of {8,4,2,1} with NL <= 31: ks = 11 -> G 8, NL 31 (97.6 % of the loaded bytes are
// taps); ks = 31 -> G 1, NL 31 (96.9 %).  Any odd ks in 3..31, any W, any alignment.
#pragma once
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"

namespace aadff {

// Predicated streaming load that leaves 0 in the register when `pred` is false.  Written in PTX so that the
// destination is the load's own register (a C++ select made ptxas copy every result with a MOV right behind
// the LDG, which serialises on the load latency).
__device__ __forceinline__ float ldg_or_zero(const float* p, bool pred) {
    float v;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@p ld.global.nc.f32 %0, [%1];\n\t"
        "}"
        : "=f"(v)
        : "l"(p), "r"((int)pred));
    return v;
}

#ifndef AADFF_GC_WARPS
#define AADFF_GC_WARPS 16
#endif
constexpr int GC_WARPS = AADFF_GC_WARPS;        // = tile rows
constexpr int GC_TW = 32;                       // tile columns

template <int KS>
struct GatherCfg {
    static constexpr int KK = KS * KS;
    static constexpr int G = (8 * KK + 31) / 32 <= 31 ? 8 : (4 * KK + 31) / 32 <= 31 ? 4 : (2 * KK + 31) / 32 <= 31 ? 2 : 1;
    static constexpr int NL = (G * KK + 31) / 32;
    static constexpr int LOG2G = G == 8 ? 3 : G == 4 ? 2 : G == 2 ? 1 : 0;
    static constexpr int HH = GC_WARPS + KS - 1;
    static constexpr int HW = GC_TW + KS - 1;
    static constexpr int PITCH = GC_TW + KS;
    static constexpr int CSTRIDE = HH * PITCH;
    static constexpr int SMEM_FLOATS(int cn) { return 2 * cn * CSTRIDE; }   // double-buffered halo
};

template <int KS, int CN>
__global__ void __launch_bounds__(GC_WARPS * 32, 1)
local_psf_coalesced_kernel(const float* __restrict__ img, const float* __restrict__ psf, float* __restrict__ out,
                           int N, int C, int H, int W, int c0) {
    using Cfg = GatherCfg<KS>;
    constexpr int KK = Cfg::KK, G = Cfg::G, NL = Cfg::NL, R = (KS - 1) / 2;
    constexpr int HH = Cfg::HH, HW = Cfg::HW, PITCH = Cfg::PITCH, CSTRIDE = Cfg::CSTRIDE;
    constexpr int SUB = 32 >> Cfg::LOG2G;                      // lanes per pixel after the butterfly (>= 4)
    static_assert(CN <= SUB, "one lane per (pixel, channel) in the store");
    // The group loop stays rolled: unrolling it (g * G as an LDS immediate) quadruples the code to ~90 KB and the
    // sixteen warps, all at different places in it, thrash the instruction cache (3x slower, measured).
    extern __shared__ float s_img[];                           // 2 x [CN][HH][PITCH]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // byte offset (< 64 KB) of element (m, lane) in the halo, relative to the group; two per register
    uint32_t offp[(NL + 1) / 2];
#pragma unroll
    for (int m = 0; m < NL; ++m) {
        const int e = 32 * m + lane;
        const int q = e / KK, t = e - q * KK, i = t / KS, j = t - i * KS;
        const uint32_t o = 4u * (uint32_t)(((e < G * KK) ? i * PITCH + q + j : 0) + warp * PITCH);   // + this warp's tile row
        if (m & 1) offp[m >> 1] |= o << 16;
        else offp[m >> 1] = o;
    }

    const int tiles_x = (W + GC_TW - 1) / GC_TW, tiles_y = (H + GC_WARPS - 1) / GC_WARPS;
    const int n_tiles = N * tiles_x * tiles_y;                 // < 2^31 (checked by the host)

    // this warp's groups, in processing order: position (t, g) on a group that exists, at or after (t, g)
    auto find = [&](int& t, int& g, const float*& ptr, int& lim) -> bool {
        while (t < n_tiles) {
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
            const int h = ty * GC_WARPS + warp, w0 = tx * GC_TW + g * G;
            if (h < H && g * G < GC_TW && w0 < W) {
                ptr = psf + ((long long)(n * H + h) * W + w0) * KK;
                lim = min(G, W - w0) * KK - lane;              // element 32m + lane is a tap iff 32m < lim
                return true;
            }
            t += gridDim.x;
            g = 0;
        }
        return false;
    };

    float buf[NL];
    int pt = blockIdx.x, pg = 0, plim = 0;
    const float* pptr = nullptr;
    bool pvalid = find(pt, pg, pptr, plim);
#pragma unroll
    for (int m = 0; m < NL; ++m) buf[m] = ldg_or_zero(pvalid ? pptr + 32 * m + lane : psf, pvalid && 32 * m < plim);

    // image halo, double-buffered: the copies of the NEXT tile (cp.async, no registers) run under this tile's groups
    auto issue_halo = [&](int tile, float* dst) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int h0 = ty * GC_WARPS, w0 = tx * GC_TW;
        const uint32_t d0 = smem_u32(dst);
        for (int idx = threadIdx.x; idx < CN * HH * HW; idx += GC_WARPS * 32) {
            const int c = idx / (HH * HW), rem = idx - c * (HH * HW);
            const int yy = rem / HW, xx = rem - yy * HW;
            const int gy = min(max(h0 + yy - R, 0), H - 1), gx = min(max(w0 + xx - R, 0), W - 1);   // render_psf.py:96
            cp_async4(d0 + 4u * (uint32_t)(c * CSTRIDE + yy * PITCH + xx),
                      img + ((long long)(n * C + c0 + c) * H + gy) * W + gx);
        }
    };
    float* s_cur = s_img;
    float* s_nxt = s_img + CN * CSTRIDE;
    if ((int)blockIdx.x < n_tiles) issue_halo(blockIdx.x, s_cur);
    cp_async_wait_all();
    __syncthreads();

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int h0 = ty * GC_WARPS, w0 = tx * GC_TW, h = h0 + warp;
        if (tile + (int)gridDim.x < n_tiles) issue_halo(tile + gridDim.x, s_nxt);
        const int ngr = (h < H) ? (min(GC_TW, W - w0) + G - 1) / G : 0;
#pragma unroll 1
        for (int g = 0; g < ngr; ++g) {
            // buf holds the taps of (tile, g) == (pt, pg); find the group after it
            int nt = pt, ng = pg + 1, nlim;
            const float* nptr;
            bool nvalid = true;
            if (g + 1 < ngr) {                                  // same row strip: the next G pixels follow directly
                nptr = pptr + G * KK;
                nlim = min(G, W - w0 - ng * G) * KK - lane;
            } else {
                nvalid = find(nt, ng, nptr, nlim);
                if (!nvalid) { nlim = -(1 << 30); nptr = psf; }
            }
            const float* nsrc = nptr + (nvalid ? lane : 0);
            const int npx = min(G, W - w0 - g * G);
            const float* ib = s_cur + g * G;
            float acc[G][CN];
#pragma unroll
            for (int q = 0; q < G; ++q)
#pragma unroll
                for (int c = 0; c < CN; ++c) acc[q][c] = 0.f;
#pragma unroll
            for (int m = 0; m < NL; ++m) {
                const float v = buf[m];
                buf[m] = ldg_or_zero(nsrc + 32 * m, 32 * m < nlim);
                const uint32_t ob = (m & 1) ? (offp[m >> 1] >> 16) : (offp[m >> 1] & 0xffffu);
                const float* px = reinterpret_cast<const float*>(reinterpret_cast<const char*>(ib) + ob);
                const int q_lo = (32 * m) / KK;
                const int q_hi = (32 * m + 31) / KK < G - 1 ? (32 * m + 31) / KK : G - 1;
                float iv[CN];
#pragma unroll
                for (int c = 0; c < CN; ++c) iv[c] = px[c * CSTRIDE];
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    if (q < q_lo || q > q_hi) continue;         // compile-time after unrolling m
                    // lanes of this load that belong to pixel q: q*KK <= 32m + lane < (q+1)*KK
                    const bool mine = (q == q_lo || lane >= q * KK - 32 * m) && (q == q_hi || lane < (q + 1) * KK - 32 * m);
                    const float vq = (q_lo == q_hi || mine) ? v : 0.f;
#pragma unroll
                    for (int c = 0; c < CN; ++c) acc[q][c] = fmaf(iv[c], vq, acc[q][c]);
                }
            }
            pt = nt;
            pg = ng;
            pptr = nptr;
            // transposing butterfly: after step s the lane keeps the pixels whose bit (LOG2G-1-s) equals lane bit (4-s)
            float res[CN];
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                float a[G];
#pragma unroll
                for (int q = 0; q < G; ++q) a[q] = acc[q][c];
#pragma unroll
                for (int s = 0; s < Cfg::LOG2G; ++s) {
                    const int d = 16 >> s, half = (G >> s) >> 1;
                    const bool upper = (lane & d) != 0;
#pragma unroll
                    for (int k = 0; k < half; ++k) {
                        const float send = upper ? a[k] : a[k + half];
                        const float keep = upper ? a[k + half] : a[k];
                        a[k] = keep + __shfl_xor_sync(0xffffffffu, send, d);
                    }
                }
#pragma unroll
                for (int d = SUB >> 1; d > 0; d >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], d);
                res[c] = a[0];
            }
            const int q = lane / SUB, cc = lane - q * SUB;     // lane -> (pixel, channel)
            float r = res[0];
#pragma unroll
            for (int c = 1; c < CN; ++c) r = (cc == c) ? res[c] : r;
            if (cc < CN && q < npx)
                out[((long long)(n * C + c0 + cc) * H + h) * W + w0 + g * G + q] = r;
        }
        cp_async_wait_all();                                   // this thread's share of the next halo has landed
        __syncthreads();                                       // ... everyone's has, and nobody still reads s_cur
        float* t = s_cur;
        s_cur = s_nxt;
        s_nxt = t;
    }
}

}  // namespace aadff

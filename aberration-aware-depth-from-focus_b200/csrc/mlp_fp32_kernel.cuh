// CUDA-core (fp32 FFMA) implementation of PSFNet evaluation fused with the gather.
// It follows the reference arithmetic operation for operation in fp32
// (deeplens/psfnet_arch.py:31-47 + deeplens/render_psf.py:76-107 +
// deeplens/psfnet.py:424-441) and is the "exact" mode (mode 2) of the C-ABI:
// slow (bounded by the fp32 FFMA rate, ~60 Mpix/s) but free of any reduced
// precision operand, so it doubles as the on-device cross-check for the
// tensor-core kernel and serves PSFNet.pred() for arbitrary [M,4] probes.
//
// One CTA = 256 threads = 64 pixels.  Activations ping-pong between two
// [256 features][64 pixels] fp32 shared-memory buffers; every thread owns an
// 8-feature x 8-pixel register tile; the weights W^T[k][n] of a layer stream through shared memory in tiles of
// 16 k-rows, double-buffered with cp.async (the first version read them through L1 straight from L2 and was bound by
// that latency: 8 warps, ~300 cycles per k-row, 0.27 of the FFMA rate).
#pragma once
#include <cuda_runtime.h>
#include "render_args.h"
#include "ptx_sm100.cuh"

namespace aadff {

constexpr int F32_TP = 64;        // pixels per CTA
constexpr int F32_NT = 256;       // threads per CTA
constexpr int F32_KT = 32;        // k-rows per weight tile
constexpr int F32_SMEM = (2 * 256 * F32_TP + 4 * F32_TP * 5 + 2 * F32_KT * 256) * 4;

struct Fp32Net {
    const float* wt[MAX_LAYERS];   // W^T, [K][npad] fp32
    const float* bias[MAX_LAYERS]; // [npad]
    int k[MAX_LAYERS];
    int npad[MAX_LAYERS];          // multiple of 8
    int n_layers;
    int kk;                        // ks*ks
};

__device__ __forceinline__ void f32_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void f32_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void f32_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// out[f][p] = sum_k WT[k][nb + f] * in[k][p] + b[nb + f] for the 8x8 tile owned by this thread (f0 = its first feature
// inside the block of ncols <= 256 columns starting at nb).  Called by ALL threads of the CTA (block-wide barriers);
// threads whose tile lies beyond ncols only help with the copies.  wbuf: 2 x [F32_KT][256] floats.
__device__ __forceinline__ void f32_tile(const float* __restrict__ wt, const float* __restrict__ bias, int K,
                                         int npad, int nb, int ncols, int f0, const float* in, int pg, float* wbuf,
                                         float (&acc)[8][8]) {
    const bool active = f0 < ncols;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const float b = active ? __ldg(bias + nb + f0 + a) : 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[a][p] = b;
    }
    const int nkt = (K + F32_KT - 1) / F32_KT;
    const int c4 = ncols >> 2;                                   // float4 per tile row (ncols is a multiple of 8)
    auto issue = [&](int kt, int b) {
        const int rows = min(F32_KT, K - kt * F32_KT);
        const uint32_t d0 = smem_u32(wbuf + b * (F32_KT * 256));
        for (int idx = threadIdx.x; idx < rows * c4; idx += F32_NT) {
            const int rr = idx / c4, cc = idx - rr * c4;
            f32_cp_async16(d0 + 4u * (uint32_t)(rr * 256 + cc * 4), wt + (size_t)(kt * F32_KT + rr) * npad + nb + cc * 4);
        }
        f32_cp_commit();
    };
    issue(0, 0);
    for (int kt = 0; kt < nkt; ++kt) {
        if (kt + 1 < nkt) {
            issue(kt + 1, (kt + 1) & 1);
            f32_cp_wait<1>();
        } else {
            f32_cp_wait<0>();
        }
        __syncthreads();                                         // tile kt has landed (everyone's copies)
        if (active) {
            const float* wrow = wbuf + (kt & 1) * (F32_KT * 256) + f0;
            const float* xrow = in + (size_t)kt * F32_KT * F32_TP;
            const int rows = min(F32_KT, K - kt * F32_KT);
#pragma unroll 4
            for (int k = 0; k < rows; ++k) {
                const float4 w0 = *reinterpret_cast<const float4*>(wrow + k * 256);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + k * 256 + 4);
                const float4 a0 = *reinterpret_cast<const float4*>(xrow + k * F32_TP + pg * 4);
                const float4 a1 = *reinterpret_cast<const float4*>(xrow + k * F32_TP + 32 + pg * 4);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int p = 0; p < 8; ++p) acc[a][p] = fmaf(w[a], x[p], acc[a][p]);
            }
        }
        __syncthreads();                                         // tile buffer kt & 1 may be refilled (by tile kt + 2)
    }
}

template <bool RENDER>
__global__ void __launch_bounds__(F32_NT, 1)
mlp_fp32_kernel(Fp32Net net, RenderArgs ra, const float* __restrict__ probes, float* __restrict__ psf_out,
                long long M0, long long M) {     // flat pixel*slice (or probe) ids [M0, M)
    extern __shared__ __align__(16) float sm[];
    float* buf0 = sm;
    float* buf1 = sm + 256 * F32_TP;
    float* red = sm + 2 * 256 * F32_TP;       // [4 parts][64 px][5]
    float* wbuf = red + 4 * F32_TP * 5;       // 2 x [F32_KT][256] weight tiles
    const int t = threadIdx.x;
    const int fg = t >> 3, pg = t & 7;
    const int gp = t & 63, part = t >> 6;     // gather-phase mapping
    const int kk = net.kk;

    for (long long base = M0 + (long long)blockIdx.x * F32_TP; base < M; base += (long long)gridDim.x * F32_TP) {
        // ---- network input (x, y, z, foc_z) -> buf0[0..3][p]
        if (t < F32_TP) {
            const long long id = base + t;
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            if (id < M) {
                if (RENDER) {
                    const int w = (int)(id % ra.W);
                    const int h = (int)((id / ra.W) % ra.H);
                    const int s = (int)((id / ((long long)ra.W * ra.H)) % ra.S);
                    const int n = (int)(id / ((long long)ra.W * ra.H * ra.S));
                    o[0] = coord_x(w, ra.W, ra.step_x);
                    o[1] = coord_y(h, ra.H, ra.step_y);
                    o[2] = depth_to_z(__ldg(ra.depth + ((long long)n * ra.H + h) * ra.W + w), ra.d_min, ra.d_range);
                    o[3] = depth_to_z(__ldg(ra.foc + (long long)n * ra.foc_stride + s), ra.d_min, ra.d_range);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[c] = __ldg(probes + id * 4 + c);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) buf0[c * F32_TP + t] = o[c];
        }
        __syncthreads();

        float* in = buf0;
        float* ob = buf1;
        // ---- hidden layers (ReLU)
        for (int l = 0; l < net.n_layers - 1; ++l) {
            const int f0 = fg * 8;
            float acc[8][8];
            f32_tile(net.wt[l], net.bias[l], net.k[l], net.npad[l], 0, net.npad[l], f0, in, pg, wbuf, acc);
            if (f0 < net.npad[l]) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    float4 v0 = make_float4(fmaxf(acc[a][0], 0.f), fmaxf(acc[a][1], 0.f), fmaxf(acc[a][2], 0.f),
                                            fmaxf(acc[a][3], 0.f));
                    float4 v1 = make_float4(fmaxf(acc[a][4], 0.f), fmaxf(acc[a][5], 0.f), fmaxf(acc[a][6], 0.f),
                                            fmaxf(acc[a][7], 0.f));
                    *reinterpret_cast<float4*>(ob + (f0 + a) * F32_TP + pg * 4) = v0;
                    *reinterpret_cast<float4*>(ob + (f0 + a) * F32_TP + 32 + pg * 4) = v1;
                }
            }
            __syncthreads();
            float* tmp = in; in = ob; ob = tmp;
        }

        // ---- head (Sigmoid) in blocks of 256 taps, each followed by its share of the gather
        const int L = net.n_layers - 1;
        const long long gid = base + gp;
        int gn = 0, gs = 0, gh = 0, gw = 0;
        if (RENDER && gid < M) {
            gw = (int)(gid % ra.W);
            gh = (int)((gid / ra.W) % ra.H);
            gs = (int)((gid / ((long long)ra.W * ra.H)) % ra.S);
            gn = (int)(gid / ((long long)ra.W * ra.H * ra.S));
        }
        const int r = (ra.ks - 1) / 2;
        float ssum = 0.f, cacc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int nb = 0; nb < net.npad[L]; nb += 256) {
            const int f0 = nb + fg * 8;
            float acc[8][8];
            f32_tile(net.wt[L], net.bias[L], net.k[L], net.npad[L], nb, min(256, net.npad[L] - nb), fg * 8, in, pg, wbuf, acc);
            if (f0 < net.npad[L]) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    float v[8];
#pragma unroll
                    for (int p = 0; p < 8; ++p) v[p] = 1.0f / (1.0f + expf(-acc[a][p]));
                    *reinterpret_cast<float4*>(ob + (fg * 8 + a) * F32_TP + pg * 4) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(ob + (fg * 8 + a) * F32_TP + 32 + pg * 4) =
                        make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            __syncthreads();
            if (gid < M) {
                const int nt = min(256, kk - nb);
                for (int tl = part; tl < nt; tl += 4) {
                    const float s = ob[tl * F32_TP + gp];
                    ssum += s;
                    if (RENDER) {
                        const int tap = nb + tl;
                        const int i = tap / ra.ks, j = tap - i * ra.ks;
                        const int yy = min(max(gh + i - r, 0), ra.H - 1);
                        const int xx = min(max(gw + j - r, 0), ra.W - 1);
                        const float* px = ra.img + (((long long)gn * ra.Ctot + ra.c0) * ra.H + yy) * ra.W + xx;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < ra.C) cacc[c] = fmaf(__ldg(px + (long long)c * ra.H * ra.W), s, cacc[c]);
                    } else {
                        psf_out[gid * kk + nb + tl] = s;
                    }
                }
            }
            __syncthreads();
        }
        // ---- reduce the 4 tap-partitions, normalise (F.normalize p=1, eps 1e-12), write
        red[(part * F32_TP + gp) * 5 + 0] = ssum;
#pragma unroll
        for (int c = 0; c < 4; ++c) red[(part * F32_TP + gp) * 5 + 1 + c] = cacc[c];
        __syncthreads();
        if (gid < M) {
            float tot = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) tot += red[(q * F32_TP + gp) * 5];
            const float denom = fmaxf(tot, 1e-12f);
            if (RENDER) {
                if (part == 0) {
                    for (int c = 0; c < ra.C; ++c) {
                        float v = 0.f;
#pragma unroll
                        for (int q = 0; q < 4; ++q) v += red[(q * F32_TP + gp) * 5 + 1 + c];
                        ra.out[gn * ra.os_n + (ra.c0 + c) * ra.os_c + gs * ra.os_s + gh * ra.os_h + gw * ra.os_w] = v / denom;
                    }
                }
            } else {
                for (int tl = part; tl < kk; tl += 4) psf_out[gid * kk + tl] = psf_out[gid * kk + tl] / denom;
            }
        }
        __syncthreads();
    }
}

}  // namespace aadff

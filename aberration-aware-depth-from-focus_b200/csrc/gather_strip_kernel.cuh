// Stand-alone per-pixel PSF gather, strip-walking version for ks <= 15: the CUDA
// counterpart of deeplens/render_psf.py:76-107 (local_psf_render) for a PSF tensor in HBM ([N,H,W,ks,ks] fp32).
//
// The register-streaming kernel (gather_coalesced_kernel.cuh) spends ~83 warp instructions per pixel at ks = 11
// (per-element halo addressing + a transposing butterfly) and is issue-bound at 0.66 of the HBM peak.  This one is
// built so that the arithmetic is nothing but LDS + FFMA with compile-time indices (~16 warp instructions per pixel):
//   * no CTA-wide synchronisation at all: every WARP is its own pipeline.  It walks down a vertical strip of
//     64 pixel columns; the PSFs of one strip row are ONE contiguous run of 64*ks^2 floats in HBM which lane 0 pulls
//     into the warp's private shared-memory slot with a single cp.async.bulk (mbarrier complete_tx).  SL slots per
//     warp, refilled as soon as the warp has consumed them; NW warps per SM keep NW*SL chunks in flight or in use.
//   * lane j owns the two pixels (2j, 2j+1) of the strip row.  Its taps are the flat run [2j*ks^2, (2j+2)*ks^2) of
//     the chunk, read as ks^2 LDS.64: the lane stride is 2*ks^2 floats with ks^2 odd, so the sixteen lanes of a
//     half-warp hit sixteen different even banks -- conflict-free without any padding or transposition.
//   * the image halo is a circular buffer of ks+1 rows per warp (planar, 64+ks-1 columns): walking down one row
//     costs ONE new halo row (cp.async, prefetched a row ahead).  Per PSF row the lane loads a (ks+1)-wide window
//     per channel with LDS.64 and both pixels take their taps from it: (ks+1)/2 * C + ks LDS.64 for 2*ks*C FFMA.
//   * the flattened list of (image, strip, row) is cut into one contiguous run per warp (balanced to +-1 row), so
//     there is no work counter and no tail.
// Requirements (else the host falls back to the register-streaming kernel): W % 4 == 0, 16-byte aligned psf,
// 8-byte aligned out, ks odd in 3..15 (larger kernels: see StripPlan in aadff_api.cu).
#pragma once
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"

namespace aadff {

constexpr int GSW_PX = 64;                      // pixel columns per pass over a strip row = 2 per lane

// SPX = strip width: 64, or 64*U taken in U passes over ONE bulk copy (small kernels: the memory system likes long contiguous requests --
// 12.5 KB chunks at ks = 7 reach 0.79 of the HBM peak, 25 KB 0.84, 50 KB 0.87)
template <int KS, int CN, int SPX_ = GSW_PX>
struct StripCfg {
    static constexpr int KK = KS * KS;
    static constexpr int SPX = SPX_;                                    // strip width
    static constexpr int U = (SPX + GSW_PX - 1) / GSW_PX;               // 64-column passes per strip row
    static_assert(SPX % 32 == 0, "strip width");
    static constexpr int HR = KS + 1;                                   // circular halo rows
    static constexpr int PITCH = SPX + KS - 1;                          // even
    static constexpr int CHUNK_BYTES = SPX * KK * 4;                    // multiple of 256
    static constexpr int HALO_BYTES = ((HR * CN * PITCH * 4 + 127) / 128) * 128;
    __host__ __device__ static constexpr int WARP_BYTES(int sl) { return sl * CHUNK_BYTES + HALO_BYTES; }
    __host__ __device__ static constexpr int SMEM_BYTES(int sl, int nw) { return nw * WARP_BYTES(sl) + nw * sl * 8; }
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory");
}

template <int KS, int CN, int SL, int NW, int SPX_ = GSW_PX>
__global__ void __launch_bounds__(NW * 32, 1)
local_psf_strip_kernel(const float* __restrict__ img, const float* __restrict__ psf, float* __restrict__ out,
                       int N, int C, int H, int W, int c0) {
    using Cfg = StripCfg<KS, CN, SPX_>;
    constexpr int KK = Cfg::KK, R = (KS - 1) / 2, HR = Cfg::HR, PITCH = Cfg::PITCH, SPX = Cfg::SPX, U = Cfg::U;
    constexpr int WB = Cfg::WARP_BYTES(SL);
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_psf = reinterpret_cast<float*>(s_raw + warp * WB);
    float* s_halo = reinterpret_cast<float*>(s_raw + warp * WB + SL * Cfg::CHUNK_BYTES);
    const uint32_t bar0 = smem_u32(s_raw + NW * WB) + 8u * (uint32_t)(warp * SL);
    const uint32_t halo_u32 = smem_u32(s_halo);

    const int nstrips = (W + SPX - 1) / SPX;
    // The flattened (image, strip, row) list is cut into one run per warp, balanced by ROWS: a warp's time per row is
    // the latency of its chunk, not the chunk's width (balancing by pixels made the warps of a narrower last strip
    // the slowest: measured 0.61 against 0.74 of the HBM peak at ks = 7, W = 640, 256-column strips) -- the host
    // instead picks a strip width that divides W well.
    const long long RT = (long long)N * nstrips * H;                    // strip rows in the launch
    const long long TW = (long long)gridDim.x * NW, gw = (long long)blockIdx.x * NW + warp;
    const long long q0 = RT * gw / TW, q1 = RT * (gw + 1) / TW;
    if (q0 >= q1) return;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SL; ++s) mbar_init(bar0 + 8u * s, 1);
        fence_mbar_init();
    }
    __syncwarp();

    // producer cursor (lane 0 issues): SL chunks ahead of the consumer
    const int nq = (int)(q1 - q0);                                      // this warp's strip rows
    int ip = 0;
    int py = (int)(q0 % H), ps = (int)((q0 / H) % nstrips), pn = (int)(q0 / ((long long)H * nstrips));
    auto issue_chunk = [&]() {
        if (ip < nq) {
            if (lane == 0) {
                const int slot = ip % SL;
                const uint32_t bytes = (uint32_t)min(SPX, W - SPX * ps) * (uint32_t)(KK * 4);
                mbar_arrive_expect_tx(bar0 + 8u * slot, bytes);
                bulk_g2s(smem_u32(s_psf + slot * SPX * KK),
                         psf + ((long long)(pn * H + py) * W + SPX * ps) * KK, bytes, bar0 + 8u * slot);
            }
            ++ip;
            if (++py == H) {
                py = 0;
                if (++ps == nstrips) { ps = 0; ++pn; }
            }
        }
    };
#pragma unroll
    for (int s = 0; s < SL; ++s) issue_chunk();

    // consumer cursor
    int y = (int)(q0 % H), s = (int)((q0 / H) % nstrips), n = (int)(q0 / ((long long)H * nstrips));
    // one halo row: image row clamp(r), columns clamp(64 s - R + i), i < PITCH  (replicate pad, render_psf.py:96)
    auto issue_halo_row = [&](int r, int hslot) {
        const int gy = min(max(r, 0), H - 1);
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            const float* src = img + ((long long)(n * C + c0 + c) * H + gy) * W;
            const uint32_t dst = halo_u32 + 4u * (uint32_t)((hslot * CN + c) * PITCH);
#pragma unroll
            for (int i0 = 0; i0 < PITCH; i0 += 32) {
                const int i = i0 + lane;
                if (i < PITCH) cp_async4(dst + 4u * i, src + min(max(SPX * s - R + i, 0), W - 1));
            }
        }
    };

    bool fresh = true;          // the halo of this strip has not been loaded yet
    int hs = 0;                 // circular slot of image row y - R
    for (int it = 0; it < nq; ++it) {
        if (fresh) {
            // nobody reads the halo here (program order + the __syncwarp that ended the previous row)
            cp_async_wait_group<0>();
#pragma unroll 1
            for (int d = 0; d < KS; ++d) issue_halo_row(y - R + d, d);
            cp_async_commit();
            issue_halo_row(y + R + 1, KS);
            cp_async_commit();
            hs = 0;
            fresh = false;
        }
        cp_async_wait_group<1>();                                   // rows y-R .. y+R have landed (this lane's part)
        const int slot = it % SL;
        mbar_wait(bar0 + 8u * slot, (uint32_t)((it / SL) & 1));
        __syncwarp();                                               // ... every lane's part

        const int npx = min(SPX, W - SPX * s);
#pragma unroll 1
        for (int u = 0; u < U; ++u) {                               // 64 columns per pass
            const int col = GSW_PX * u + 2 * lane;                  // this lane's first column inside the strip
            float acc[2][CN];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int c = 0; c < CN; ++c) acc[p][c] = 0.f;
            if (col < npx) {
                const float2* tp = reinterpret_cast<const float2*>(s_psf + slot * SPX * KK) + (col >> 1) * KK;
#pragma unroll
                for (int dy = 0; dy < KS; ++dy) {
                    int hsl = hs + dy;
                    hsl -= (hsl >= HR) ? HR : 0;
                    const float2* hrow = reinterpret_cast<const float2*>(s_halo + hsl * CN * PITCH) + (col >> 1);
                    float win[CN][KS + 1];
#pragma unroll
                    for (int c = 0; c < CN; ++c)
#pragma unroll
                        for (int i = 0; i < (KS + 1) / 2; ++i) {
                            const float2 v = hrow[c * (PITCH / 2) + i];
                            win[c][2 * i] = v.x;
                            win[c][2 * i + 1] = v.y;
                        }
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const int e0 = p * KK + dy * KS;            // compile-time after unrolling
#pragma unroll
                        for (int dx = 0; dx < KS; ++dx) {
                            const int e = e0 + dx;
                            const float2 tv = tp[e >> 1];
                            const float t = (e & 1) ? tv.y : tv.x;
#pragma unroll
                            for (int c = 0; c < CN; ++c) acc[p][c] = fmaf(win[c][p + dx], t, acc[p][c]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < CN; ++c)
                    *reinterpret_cast<float2*>(out + ((long long)(n * C + c0 + c) * H + y) * W + SPX * s + col) =
                        make_float2(acc[0][c], acc[1][c]);
            }
        }
        __syncwarp();                                               // every lane is done with the chunk and row y-R
        issue_chunk();                                              // refill this slot (chunk q + SL)
        if (++y == H) {
            y = 0;
            fresh = true;
            if (++s == nstrips) { s = 0; ++n; }
        } else {
            // rows y-R .. y+R of the new y: the missing one (old y+R+1) is in flight; fetch the one after it into the
            // slot of the row that just left the window
            issue_halo_row(y + R + 1, hs);
            cp_async_commit();
            hs = (hs + 1 == HR) ? 0 : hs + 1;
        }
    }
    cp_async_wait_group<0>();
}

}  // namespace aadff

// Single-term ("fast") fused PSFNet + render with TWO TILES IN FLIGHT per CTA (sm_100a, tcgen05 / TMEM).
//
// Same path as fused_tc_kernel.cuh -- PSFNet.render -> MLP.forward -> local_psf_render, S slices, one launch
// (deeplens/psfnet.py:393-441, psfnet_arch.py:24-47, render_psf.py:76-107, 2_aber_aware_dff_aif.py:108-114) -- for
// the arithmetic mode in which every layer is ONE fp16 MMA term (AADFF_MODE_FAST).  In that mode the one-tile kernel
// keeps the tensor pipe only 50 % busy (round-2 ncu): a layer is 2048 cycles of MMAs and ~1700 cycles of epilogue on
// the same tile, back to back.  Nothing needs the "A lo" operand here, so its 64 KB hold a SECOND tile's activations
// and the two tiles alternate, layer by layer:
//
//      tensor pipe :  MMA(t0, l)   MMA(t1, l)   MMA(t0, l+1)   MMA(t1, l+1) ...
//      epilogue    :               epi(t0, l)   epi(t1, l)     epi(t0, l+1) ...
//
// Tile slot s owns accumulator columns [256 s, 256 s + 256) of TMEM and activation buffer A[s] (updated in place: the
// MMAs of (s, l) have all retired when its accumulator is reported full).  Hand-offs are whole layers (one mbarrier per
// slot and direction) -- the other slot's MMAs cover the latency that the one-tile kernel had to hide with 32-column
// chunks.  Weights are streamed once per (layer, slot) through a 4 x 16 KB ring, two stages (four MMAs) per commit.
// Restrictions (the host falls back to the one-tile kernel otherwise): single head block (k*k <= 256), C <= 4, and the
// shared-memory budget (k <= 13 at 227 KB).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "fused_tc_kernel.cuh"

namespace aadff {

struct F2Params {
    RenderArgs ra;
    const uint8_t* wpack;   // fp16 slabs, consumption order (hi, lo per K-slab; only the hi slabs are read)
    const float* bias;      // L1..L9 biases then the padded head bias (pre-scaled by -log2 e)
    const float* w0b0;      // W0 [64][4] then b0 [64]
    TcGroup g[TC_MAX_GROUPS];
    int n_groups, n_hidden, n_bias, bias_skip, kk;
    int tiles_x, tiles_y;
    long long n_tiles, tile0;
    int halo_pitch;         // float4 per halo row (= 16 + ks - 1)
    uint32_t halo_bytes;    // per slot
    uint32_t off_stage, off_bias, off_w0, off_halo, off_red, off_bar, off_ones, off_bslab;
};

constexpr int F2_STAGES = 4;

__global__ void __launch_bounds__(TC_NT, 1) fused_fast2_kernel(const __grid_constant__ F2Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const RenderArgs& ra = P.ra;
    const uint32_t stage0 = sbase + P.off_stage;
    float* s_bias = reinterpret_cast<float*>(smem + P.off_bias);
    float* s_w0 = reinterpret_cast<float*>(smem + P.off_w0);
    const uint32_t bar0 = sbase + P.off_bar;
    // the ring is used as TWO pairs of 16 KB stages: one "full" and one "empty" barrier per pair, so that the MMA warp --
    // whose own instruction stream is the critical path -- waits and commits once per four MMAs
    auto bar_full = [&](int pr) { return bar0 + 8u * pr; };
    auto bar_empty = [&](int pr) { return bar0 + 8u * (F2_STAGES + pr); };
    auto bar_aready = [&](int t) { return bar0 + 8u * (2 * F2_STAGES + t); };
    auto bar_accfull = [&](int t) { return bar0 + 8u * (2 * F2_STAGES + 2 + t); };
    auto bar_accfree = [&](int t) { return bar0 + 8u * (2 * F2_STAGES + 4 + t); };
    auto bar_bfull = [&]() { return bar0 + 8u * (2 * F2_STAGES + 6); };
    auto bar_bempty = [&]() { return bar0 + 8u * (2 * F2_STAGES + 7); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.off_bar + 8 * (2 * F2_STAGES + 8));
    static_assert(8 * (2 * F2_STAGES + 8) + 4 <= TC_BAR_BYTES, "barrier area too small");

    // tile pairs of this CTA
    const long long units = (P.n_tiles + 1) / 2;
    const long long my_iters = (units > (long long)blockIdx.x) ? (units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int pr = 0; pr < 2; ++pr) { mbar_init(bar_full(pr), 1); mbar_init(bar_empty(pr), 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_aready(t), TC_EPI_WARPS);
            mbar_init(bar_accfull(t), 1);
            mbar_init(bar_accfree(t), TC_EPI_WARPS);
        }
        mbar_init(bar_bfull(), 1);
        mbar_init(bar_bempty(), 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < P.n_bias - P.bias_skip; i += TC_NT) s_bias[i] = __ldg(P.bias + P.bias_skip + i);
    for (int i = threadIdx.x; i < 2 * TC_M; i += TC_NT) {                      // constant A operand of the bias slabs
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i < TC_M) v.x = 0x3C003C00u;
        *reinterpret_cast<uint4*>(smem + P.off_ones + (i < TC_M ? 0 : TC_A_LBO) + (i % TC_M) * 16) = v;
    }
    for (int i = threadIdx.x; i < TC_BSLAB_BYTES / 16; i += TC_NT)             // K columns 8..15 of every bias slab: zeros
        *reinterpret_cast<uint4*>(smem + P.off_bslab + TC_BSLAB_BYTES + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    for (int i = threadIdx.x; i < 320; i += TC_NT) s_w0[i] = __ldg(P.w0b0 + i);
    if (warp == TC_WARP_PRODUCER) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC_WARP_PRODUCER) {
        // =========================================================== weight producer
        int pr = 0;
        uint32_t phase = 0, bphase = 0;
        for (long long iter = 0; iter < my_iters; ++iter) {
            for (int gi = 0; gi < P.n_groups; ++gi) {
                const uint32_t bytes = (uint32_t)P.g[gi].N * (TC_SLAB_K * 2);
                const int npairs = P.g[gi].K / (2 * TC_SLAB_K);          // 1 (K = 64) or 4 (K = 256)
                for (int slot = 0; slot < 2; ++slot) {
                    const uint8_t* src = P.wpack + P.g[gi].w_off;
                    if (gi < P.n_hidden) {
                        mbar_wait(bar_bempty(), bphase ^ 1);
                        if (elect_one_sync()) {
                            mbar_arrive_expect_tx(bar_bfull(), TC_BSLAB_BYTES);
                            bulk_g2s(sbase + P.off_bslab, src, TC_BSLAB_BYTES, bar_bfull());
                        }
                        __syncwarp();
                        bphase ^= 1;
                        src += (uint32_t)P.g[gi].N * 32;
                    }
                    for (int j = 0; j < npairs; ++j) {
                        mbar_wait(bar_empty(pr), phase ^ 1);
                        if (elect_one_sync()) {
                            mbar_arrive_expect_tx(bar_full(pr), 2 * bytes);
                            // hi slabs of K-slabs 2j and 2j+1 (the packed stream interleaves hi, lo)
                            bulk_g2s(stage0 + (2 * pr) * TC_STAGE_BYTES, src + (size_t)(4 * j) * bytes, bytes, bar_full(pr));
                            bulk_g2s(stage0 + (2 * pr + 1) * TC_STAGE_BYTES, src + (size_t)(4 * j + 2) * bytes, bytes, bar_full(pr));
                        }
                        __syncwarp();
                        pr ^= 1;
                        if (pr == 0) phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // =========================================================== MMA issuer
        // The work is a flat sequence of steps (group g, slot, pair j), four MMAs each.  Everything step k+1 has to wait
        // for -- its weight pair, and at the start of a (group, slot) the accumulator, the bias slab and the activations
        // -- is waited for right after the MMAs of step k are queued and BEFORE step k's commits (a commit blocks the
        // issuing thread until the pipe has drained to it), so the waits overlap MMA execution.  No wait of step k+1 can
        // depend on a commit of step k: its weight pair was released by step k-1, and its slot's hand-offs belong to the
        // other slot's previous layer.
        uint32_t fph = 0, aph = 0, frph = 0, bph = 0, used = 0;
        const uint64_t da0 = umma_smem_desc(sbase, TC_A_LBO, 128);
        const uint64_t da_ones = umma_smem_desc(sbase + P.off_ones, TC_A_LBO, 128);
        constexpr uint32_t KSTEP_A = (2 * TC_A_LBO) >> 4;
        auto prewait = [&](int gi, int slot, int j, int pr) {
            if (j == 0) {
                if ((used >> slot) & 1) {                            // the epilogue has drained this slot's accumulator
                    mbar_wait(bar_accfree(slot), (frph >> slot) & 1);
                    frph ^= 1u << slot;
                }
                used |= 1u << slot;
                if (gi < P.n_hidden) { mbar_wait(bar_bfull(), bph); bph ^= 1; }
                if (P.g[gi].new_a) { mbar_wait(bar_aready(slot), (aph >> slot) & 1); aph ^= 1u << slot; }
            }
            mbar_wait(bar_full(pr), (fph >> pr) & 1);
            fph ^= 1u << pr;
        };
        int pr = 0;
        if (my_iters > 0) prewait(0, 0, 0, 0);
        for (long long iter = 0; iter < my_iters; ++iter) {
            for (int gi = 0; gi < P.n_groups; ++gi) {
                const uint32_t gN = P.g[gi].N;
                const int npairs = P.g[gi].K / (2 * TC_SLAB_K);
                const uint32_t idesc = umma_idesc_f16_f32(TC_M, gN);
                const uint32_t kstep_b = (2 * gN * 16) >> 4;
                const uint64_t db0 = umma_smem_desc(stage0, gN * 16, 128);
                const bool has_bias = gi < P.n_hidden;
                for (int slot = 0; slot < 2; ++slot) {
                    const uint32_t d_tmem = tmem_base + slot * 256;
                    const uint64_t da = da0 + (uint64_t)((uint32_t)(slot * TC_A_PART_BYTES) >> 4);
                    for (int j = 0; j < npairs; ++j) {
                        const bool last = (j + 1 == npairs);
                        tc_fence_after_sync();
                        if (elect_one_sync()) {
                            if (j == 0 && has_bias)                  // accumulator := bias (constant A operand of ones)
                                umma_f16_ss(d_tmem, da_ones, umma_smem_desc(sbase + P.off_bslab, TC_BSLAB_BYTES, 128), idesc, 0);
                            const uint64_t ah = da + (uint32_t)j * (4 * KSTEP_A);
                            const uint64_t dbA = (db0 & ~0x3FFFull) | (((stage0 + (2 * pr) * TC_STAGE_BYTES) & 0x3FFFFu) >> 4);
                            const uint64_t dbB = (db0 & ~0x3FFFull) | (((stage0 + (2 * pr + 1) * TC_STAGE_BYTES) & 0x3FFFFu) >> 4);
                            umma_f16_ss(d_tmem, ah, dbA, idesc, has_bias || j != 0);
                            umma_f16_ss(d_tmem, ah + KSTEP_A, dbA + kstep_b, idesc, 1);
                            umma_f16_ss(d_tmem, ah + 2 * KSTEP_A, dbB, idesc, 1);
                            umma_f16_ss(d_tmem, ah + 3 * KSTEP_A, dbB + kstep_b, idesc, 1);
                        }
                        __syncwarp();
                        // ---- the next step's waits, under the MMAs just queued
                        // (exception: a one-step group, K = 64 -- the step still running owns the bias-slab slot, its release
                        //  is among the commits below, and the producer loads the next bias slab and the weights behind
                        //  it only after that: those waits come after the commits)
                        bool late = false;
                        int ng = gi, nslot = slot, nj = j + 1;
                        {
                            bool any = true;
                            if (nj == npairs) {
                                nj = 0;
                                if (++nslot == 2) {
                                    nslot = 0;
                                    if (++ng == P.n_groups) { ng = 0; any = (iter + 1 < my_iters); }
                                }
                            }
                            late = any && j == 0 && has_bias && nj == 0;
                            if (any && !late) prewait(ng, nslot, nj, pr ^ 1);
                        }
                        // ---- now the draining commits
                        if (elect_one_sync()) {
                            if (last) umma_commit(bar_accfull(slot));
                            umma_commit(bar_empty(pr));
                            if (j == 0 && has_bias) umma_commit(bar_bempty());
                        }
                        __syncwarp();
                        if (late) prewait(ng, nslot, nj, pr ^ 1);
                        pr ^= 1;
                    }
                }
            }
        }
    } else {
        // =========================================================== epilogue / compute warps 0..7
        const int e = warp, q = warp & 3, hh = e >> 2;
        const int et = e * 32 + lane;
        const int row = q * 32 + lane;
        const int ty = row >> 4, tx = row & 15;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int ks = ra.ks, r = (ks - 1) / 2, kk = P.kk;
        const int HH = TC_TILE_H + ks - 1, HW = TC_TILE_W + ks - 1, HP = P.halo_pitch;
        const int tiles_xy = P.tiles_x * P.tiles_y;
        const uint32_t a_row = (uint32_t)row * 16;
        const float NEG_LOG2E = -1.4426950408889634f;
        uint32_t afph = 0;

        auto tile_coords = [&](long long tile, int& n, int& s, int& h0, int& w0) {
            const long long gt = tile + P.tile0;
            const int txy = (int)(gt % tiles_xy);
            const long long ns = gt / tiles_xy;
            s = (int)(ns % ra.S);
            n = (int)(ns / ra.S);
            h0 = (txy / P.tiles_x) * TC_TILE_H;
            w0 = (txy % P.tiles_x) * TC_TILE_W;
        };
        float nx_depth[2] = {0.f, 0.f}, nx_foc[2] = {0.f, 0.f};
        auto fetch_dz = [&](long long tile, int slot) {
            nx_depth[slot] = 0.f;
            nx_foc[slot] = 0.f;
            if (tile < P.n_tiles) {
                int n, s, h0, w0;
                tile_coords(tile, n, s, h0, w0);
                const int hc = min(h0 + ty, ra.H - 1), wc = min(w0 + tx, ra.W - 1);
                nx_depth[slot] = __ldg(ra.depth + ((long long)n * ra.H + hc) * ra.W + wc);
                nx_foc[slot] = __ldg(ra.foc + (long long)n * ra.foc_stride + s);
            }
        };
        // layer 0 (4 -> 64, fp32 FFMA) of `tile` into A[slot]; arrives a_ready[slot]
        auto layer0 = [&](long long tile, int slot) {
            int n, s, h0, w0;
            tile_coords(tile < P.n_tiles ? tile : 0, n, s, h0, w0);
            const float x = coord_x(min(w0 + tx, ra.W - 1), ra.W, ra.step_x);
            const float y = coord_y(min(h0 + ty, ra.H - 1), ra.H, ra.step_y);
            const float z = depth_to_z(nx_depth[slot], ra.d_min, ra.d_range);
            const float fz = depth_to_z(nx_foc[slot], ra.d_min, ra.d_range);
            const uint32_t a_base = sbase + slot * TC_A_PART_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kg = hh * 4 + i;
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int f = kg * 8 + u;
                    const float4 wv = *reinterpret_cast<const float4*>(s_w0 + f * 4);
                    float a = s_w0[256 + f];
                    a = fmaf(wv.x, x, a); a = fmaf(wv.y, y, a); a = fmaf(wv.z, z, a); a = fmaf(wv.w, fz, a);
                    v[u] = fmaxf(a, 0.f);
                }
                store_split8(v, a_base + (uint32_t)kg * TC_A_LBO + a_row, 0, false);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready(slot));
        };

        long long u = blockIdx.x;
        if (my_iters > 0) {
            for (int slot = 0; slot < 2; ++slot) {
                fetch_dz(2 * u + slot, slot);
                layer0(2 * u + slot, slot);
            }
        }
        for (long long iter = 0; iter < my_iters; ++iter, u += gridDim.x) {
            int n[2], s[2], h0[2], w0[2];
            bool live[2];
            for (int slot = 0; slot < 2; ++slot) {
                const long long t = 2 * u + slot;
                live[slot] = t < P.n_tiles;
                tile_coords(live[slot] ? t : 0, n[slot], s[slot], h0[slot], w0[slot]);
            }
            const bool more = iter + 1 < my_iters;
            if (more) {                                          // depth / focus of the next pair, used by layer0 below
                fetch_dz(2 * (u + gridDim.x), 0);
                fetch_dz(2 * (u + gridDim.x) + 1, 1);
            }
            named_bar_sync(1, TC_EPI_THREADS);                   // the previous pair's gathers are done with halo / red
            const long long cstride = (long long)ra.H * ra.W;
            auto halo_fetch = [&](int slot, int idx, float4& v) -> int {
                if (!live[slot] || idx >= HH * HW) return -1;
                const int yy = idx / HW, xx = idx - yy * HW;
                const int gy = min(max(h0[slot] + yy - r, 0), ra.H - 1), gx = min(max(w0[slot] + xx - r, 0), ra.W - 1);
                const float* px = ra.img + ((long long)n[slot] * ra.Ctot + ra.c0) * cstride + (long long)gy * ra.W + gx;
                v.x = __ldg(px);
                v.y = ra.C > 1 ? __ldg(px + cstride) : 0.f;
                v.z = ra.C > 2 ? __ldg(px + 2 * cstride) : 0.f;
                v.w = ra.C > 3 ? __ldg(px + 3 * cstride) : 0.f;
                return yy * HP + xx;
            };

            // ---- hidden layers: accumulator -> ReLU -> fp16 -> A[slot] (in place), whole-layer hand-off
            for (int gi = 0; gi < P.n_hidden; ++gi) {
#pragma unroll
                for (int slot = 0; slot < 2; ++slot) {
                    float4* s_halo = reinterpret_cast<float4*>(smem + P.off_halo + slot * P.halo_bytes);
                    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int hslot = halo_fetch(slot, gi * TC_EPI_THREADS + et, hv);
                    mbar_wait(bar_accfull(slot), (afph >> slot) & 1);
                    afph ^= 1u << slot;
                    tc_fence_after_sync();
                    const uint32_t t_acc = t_lane + slot * 256;
                    const uint32_t a_base = sbase + slot * TC_A_PART_BYTES;
                    uint32_t rr[2][32];
                    tmem_ld32(t_acc + hh * 32, rr[0]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        tmem_ld_wait();
                        if (j < 3) tmem_ld32(t_acc + (j + 1) * 64 + hh * 32, rr[(j + 1) & 1]);
                        const int col = j * 64 + hh * 32;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            epi_group8(&rr[j & 1][i * 8], a_base, 0, col / 8 + i, a_row, false);
                    }
                    fence_proxy_async_smem();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar_aready(slot));
                        mbar_arrive(bar_accfree(slot));
                    }
                    if (hslot >= 0) s_halo[hslot] = hv;
                }
            }
            for (int slot = 0; slot < 2; ++slot) {
                float4* s_halo = reinterpret_cast<float4*>(smem + P.off_halo + slot * P.halo_bytes);
                for (int idx = P.n_hidden * TC_EPI_THREADS + et; idx < HH * HW; idx += TC_EPI_THREADS) {   // remainder
                    float4 hv;
                    const int hslot = halo_fetch(slot, idx, hv);
                    if (hslot >= 0) s_halo[hslot] = hv;
                }
            }
            named_bar_sync(1, TC_EPI_THREADS);                   // both halo tiles complete

            // ---- head (single block): sigmoid, gather from the slot's halo tile (render_psf.py:103-105)
            const int gi = P.n_hidden;
            const int gN = P.g[gi].N;
            const float* bias = s_bias + (P.g[gi].bias_off - P.bias_skip);
#pragma unroll 1
            for (int slot = 0; slot < 2; ++slot) {
                float ssum = 0.f, cacc[TC_MAX_C] = {0.f, 0.f, 0.f, 0.f};
                const float4* hbase = reinterpret_cast<const float4*>(smem + P.off_halo + slot * P.halo_bytes) + ty * HP + tx;
                float* s_red = reinterpret_cast<float*>(smem + P.off_red) + slot * (TC_M * 5);
                mbar_wait(bar_accfull(slot), (afph >> slot) & 1);
                afph ^= 1u << slot;
                tc_fence_after_sync();
                if (more) layer0(2 * (u + gridDim.x) + slot, slot);   // every MMA that read A[slot] has retired
#pragma unroll 1
                for (int c32 = hh * 32; c32 < gN; c32 += 64) {
                    uint32_t rr[32];
                    tmem_ld32(t_lane + slot * 256 + c32, rr);
                    int i = c32 / ks, j = c32 - i * ks;
                    int off = i * HP + j;
                    const int nvalid = kk - c32;
                    tmem_ld_wait();
                    if (nvalid >= 32) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const float sg = rcp_approx(1.0f + ex2_approx(fmaf(__uint_as_float(rr[t]), NEG_LOG2E, bias[c32 + t])));
                            const float4 px = hbase[off];
                            ssum += sg;
                            cacc[0] = fmaf(sg, px.x, cacc[0]);
                            cacc[1] = fmaf(sg, px.y, cacc[1]);
                            cacc[2] = fmaf(sg, px.z, cacc[2]);
                            cacc[3] = fmaf(sg, px.w, cacc[3]);
                            ++off;
                            if (++j == ks) { j = 0; off += HP - ks; }
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const bool ok = t < nvalid;
                            float sg = rcp_approx(1.0f + ex2_approx(fmaf(__uint_as_float(rr[t]), NEG_LOG2E, bias[c32 + t])));
                            sg = ok ? sg : 0.f;
                            const float4 px = hbase[ok ? off : 0];
                            ssum += sg;
                            cacc[0] = fmaf(sg, px.x, cacc[0]);
                            cacc[1] = fmaf(sg, px.y, cacc[1]);
                            cacc[2] = fmaf(sg, px.z, cacc[2]);
                            cacc[3] = fmaf(sg, px.w, cacc[3]);
                            ++off;
                            if (++j == ks) { j = 0; off += HP - ks; }
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accfree(slot));
                // combine the two column halves of each pixel, normalise (F.normalize p=1), store
                if (hh == 1) {
                    s_red[row * 5 + 0] = ssum;
#pragma unroll
                    for (int c = 0; c < TC_MAX_C; ++c) s_red[row * 5 + 1 + c] = cacc[c];
                }
                named_bar_sync(2 + q, 64);
                const int h = h0[slot] + ty, w = w0[slot] + tx;
                if (hh == 0 && live[slot] && h < ra.H && w < ra.W) {
                    const float inv = 1.0f / fmaxf(ssum + s_red[row * 5], 1e-12f);
                    float* o = ra.out + n[slot] * ra.os_n + ra.c0 * ra.os_c + s[slot] * ra.os_s + h * ra.os_h + w * ra.os_w;
#pragma unroll
                    for (int c = 0; c < TC_MAX_C; ++c)
                        if (c < ra.C) o[c * ra.os_c] = (cacc[c] + s_red[row * 5 + 1 + c]) * inv;
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == TC_WARP_PRODUCER) tmem_dealloc<512>(tmem_base);
}

}  // namespace aadff

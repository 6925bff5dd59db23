// Row kernel for the input gradient of the stride-2 transposed convolutions, ConvTranspose2d(64, 64, 3, stride=2)
// (models/models.py:66-78; decoder_conv.{0,3,6,9}):
//
//   dz[n, sy, sx, ci] = sum_{ky,kx,co} dy[n, 2sy+ky, 2sx+kx, co] * W[ci, co, ky, kx]          (a stride-2 GATHER over dy)
//
// The per-tap pipeline it replaces (conv_tc.cu) re-gathered every dy pixel 2.25x from L2 and converted it each time.  Here a
// CTA walks a contiguous range of output rows and streams the INPUT rows of dy through a two-slot shared-memory ring: each row
// is loaded, split to bf16 hi/lo and written ONCE, as two column-parity sub-images (even / odd columns, one 128-byte
// SWIZZLE_128B row per pixel), because a tap's gather is unit-stride inside one parity: tap kx reads parity kx&1 at pixel
// offset kx>>1 (a descriptor shifted by one row).  All nine weight taps (144 KB of bf16 hi/lo images) stay resident.
//
// One input row feeds the output rows it touches while it is in shared memory: row 2s is tap ky=0 of output row s and tap ky=2
// of output row s-1, row 2s+1 is tap ky=1 of output row s; the accumulators of the (at most two) open output rows live in
// TMEM (4 x 128 columns, so the epilogue drains finished rows while the next ones accumulate).
//
// bf16x3 in ONE MMA per K step ("quad" form): the hi and lo planes of a sub-image are adjacent, so the A operand is
// M = 128 = [hi 64 pixel rows | lo 64 pixel rows], and a tap's weight image is [W_hi 64 rows | W_lo 64 rows] = N = 128:
// D[128 x 128] holds hi*hi, hi*lo, lo*hi (and the negligible lo*lo); the epilogue adds the column halves in registers and the
// lane halves through shared memory, then applies the ReLU mask and takes the BatchNorm-backward sums exactly like the other
// dgrad kernels (EPI_MASK_BNBWD) with coalesced 128-bit loads / stores.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "bn_tail.cuh"

namespace srlz {

namespace d2 {
constexpr int NSLOT = 2;
constexpr int PLANE = 64 * 128;                         // 64 pixel rows x 128 B (one parity, hi or lo)
constexpr int SLOT_BYTES = 4 * PLANE;                   // even_hi | even_lo | odd_hi | odd_lo
constexpr int W_TAP = 2 * PLANE;                        // [W_hi 64 rows | W_lo 64 rows]
constexpr int OFF_W = NSLOT * SLOT_BYTES;               // 65536 (a shifted read past even_lo lands in odd_hi: no padding needed)
constexpr int OFF_STG = OFF_W + 9 * W_TAP;              // 212992: [64 pixels][64 channels] fp32
constexpr int OFF_BARS = OFF_STG + 64 * 64 * 4;         // 229376
constexpr int OFF_BN = OFF_BARS + 256;                  // scale | shift | mean | invstd
constexpr int SMEM_BYTES = OFF_BN + 1024 + 1024;        // 231680 <= 232448 (1 KB of alignment slack)
constexpr int THREADS = 16 * 32;                        // warps 0-7 epilogue (lane quarter x channel half) | 8-14 producers | 15 MMA
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // M = 128, N = 128
}  // namespace d2

#define D2_STAMP(idx, slot) do { if (dbg != nullptr && blockIdx.x == 0 && (idx) >= 0 && (idx) < 64) dbg[(idx) * 16 + (slot)] = clock64(); } while (0)


template <int EPI>
__global__ void __launch_bounds__(d2::THREADS, 1) dgrad_s2_rows_kernel(const float* __restrict__ in, const unsigned char* __restrict__ wbf,
                                                                       float* __restrict__ out, const float* __restrict__ e_ypre,
                                                                       const float* __restrict__ e_scale, const float* __restrict__ e_shift,
                                                                       const float* __restrict__ e_mean, const float* __restrict__ e_invstd,
                                                                       float* __restrict__ partials, int SH, int SW, int total_rows,
                                                                       long long* __restrict__ dbg, BnTail tail) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t wsm = base + d2::OFF_W, bars = base + d2::OFF_BARS;
    // mbarriers: full[2] empty[2] tfull[4] tempty[4] wfull = 13 x 8 B
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + d2::OFF_BARS + 192);
    float* s_bn = reinterpret_cast<float*>(smem + d2::OFF_BN);
    float* stg = reinterpret_cast<float*>(smem + d2::OFF_STG);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (2 + s); };
    auto tfull_bar = [&](int i) { return bars + 8u * (4 + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (8 + i); };
    const uint32_t wfull = bars + 8u * 12;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BH = 2 * SH + 1, BW = 2 * SW + 1;
    const int i0 = (int)((long long)total_rows * blockIdx.x / gridDim.x), i1 = (int)((long long)total_rows * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int s = 0; s < d2::NSLOT; ++s) { mbar_init(full_bar(s), 7); mbar_init(empty_bar(s), 1); }   // seven producer warps
        for (int i = 0; i < 4; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 8); }   // eight epilogue warps
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (EPI == EPI_MASK_BNBWD && tid < 64) {
        s_bn[tid] = e_scale[tid]; s_bn[64 + tid] = e_shift[tid]; s_bn[128 + tid] = e_mean[tid]; s_bn[192 + tid] = e_invstd[tid];
    }
    for (int e = tid; e < d2::OFF_W / 16; e += d2::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (tid == 0) {
        mbar_arrive_expect_tx(wfull, 9 * d2::W_TAP);
        for (int t = 0; t < 9; ++t) bulk_g2s(wsm + t * d2::W_TAP, wbf + (size_t)t * d2::W_TAP, d2::W_TAP, wfull);
    }

    if (warp >= 8 && warp < 15) {
        // ================================ producers (warps 8-14): one input row of dy (BW <= 112 pixels x 64 channels) per step ================================
        // thread = (16-byte chunk jc of 8 channels, pixel group pg < 28): pixels pg, pg+28, pg+56, pg+84; a warp instruction covers 4 whole
        // pixels (4 x 256 contiguous bytes of global memory).  Pixel x goes to row x>>1 of the parity-(x&1) sub-image.
        const int pidx = tid - 256, jc = pidx & 7, pg = pidx >> 3;
        const float* in_end = in + (size_t)(total_rows / SH) * BH * BW * SRLZ_C;
        auto load = [&](const float* src, float4 (&d)[8]) {
            if (pidx < 2 * BW && src + (size_t)3 * BW * SRLZ_C <= in_end) prefetch_l2(src + (size_t)2 * BW * SRLZ_C + pidx * 32);   // two rows ahead
            src += pg * SRLZ_C + jc * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (pg + 28 * q < BW) ldg8(src + q * 28 * SRLZ_C, d[2 * q], d[2 * q + 1]);
        };
        auto store = [&](int g, const float4 (&v)[8]) {
            const int slot = g & 1, ph = (g >> 1) & 1;
            if (pidx == 0) D2_STAMP(g, 0);
            mbar_wait(empty_bar(slot), ph ^ 1);
            if (pidx == 0) D2_STAMP(g, 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int x = pg + 28 * q;
                if (x < BW) {
                    uint4 hi, lo;
                    split8(v[2 * q], v[2 * q + 1], hi, lo);
                    const int e = x >> 1;
                    unsigned char* dst = smem + slot * d2::SLOT_BYTES + (x & 1) * (2 * d2::PLANE) + e * 128 + ((jc ^ (e & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = hi;
                    *reinterpret_cast<uint4*>(dst + d2::PLANE) = lo;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(slot));
            if (pidx == 0) D2_STAMP(g, 2);
        };
        // cursor over the row sequence of this CTA: for every image segment [sa, sb] of its output-row range, input rows 2sa .. 2sb+2
        int ci = i0, cn = 0, csa = 0, csb = -1, cr = 0;
        auto seg = [&]() {
            cn = ci / SH; csa = ci - cn * SH;
            csb = csa + (i1 - ci) - 1;
            if (csb > SH - 1) csb = SH - 1;
            cr = 2 * csa;
        };
        auto valid = [&]() { return ci < i1; };
        auto src_of = [&]() { return in + ((size_t)cn * BH + cr) * BW * SRLZ_C; };
        auto advance = [&]() {
            if (++cr > 2 * csb + 2) {
                ci += csb - csa + 1;
                if (ci < i1) seg();
            }
        };
        if (valid()) {
            seg();
            // two register sets: the loads of row g+1 are in flight while row g is converted and stored
            float4 va[8], vb[8];
            load(src_of(), va);
            for (int g = 0;; g += 2) {
                advance();
                const bool more1 = valid();
                if (more1) load(src_of(), vb);
                store(g, va);
                if (!more1) break;
                advance();
                const bool more2 = valid();
                if (more2) load(src_of(), va);
                store(g + 1, vb);
                if (!more2) break;
            }
        }
    } else if (warp == 15) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        mbar_wait(wfull, 0);
        int g = 0, cnt = 0;     // input rows consumed, accumulators started
        for (int i = i0; i < i1;) {
            const int n = i / SH, sa = i - n * SH;
            int sb = sa + (i1 - i) - 1;
            if (sb > SH - 1) sb = SH - 1;
            for (int r = 2 * sa; r <= 2 * sb + 2; ++r, ++g) {
                const int slot = g & 1, s = r >> 1;
                if (lane == 0) D2_STAMP(g, 3);
                mbar_wait(full_bar(slot), (g >> 1) & 1);
                fence_proxy_async_smem();   // consumer-side proxy fence (see dec12_rows_tc.cu)
                tc_fence_after();
                if (lane == 0) D2_STAMP(g, 4);
                const uint32_t a_even = base + slot * d2::SLOT_BYTES, a_odd = a_even + 2 * d2::PLANE;
                // one tap row ky of this input row into accumulator `buf`: 3 taps x 4 K steps, one M=128 x N=128 MMA each
                auto issue = [&](int ky, int buf, bool fresh) {
                    if (leader) {
                        const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint64_t ad = make_desc_sw128((kx & 1) ? a_odd : a_even + (kx >> 1) * 128);
                            const uint64_t wd = make_desc_sw128(wsm + (ky * 3 + kx) * d2::W_TAP);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                umma_bf16(d_tmem, ad + adv, wd + adv, d2::IDESC, (fresh && kx == 0 && k == 0) ? 0u : 1u);
                            }
                        }
                    }
                };
                if ((r & 1) == 0) {
                    if (s - 1 >= sa) {             // ky = 2 closes output row s-1 (the accumulator opened before the current one)
                        const int buf = (cnt - 1) & 3;
                        issue(2, buf, false);
                        if (leader) umma_commit(tfull_bar(buf));
                    }
                    if (s <= sb) {                 // ky = 0 opens output row s
                        const int buf = cnt & 3;
                        mbar_wait(tempty_bar(buf), ((cnt >> 2) & 1) ^ 1);
                        tc_fence_after();
                        issue(0, buf, true);
                        ++cnt;
                    }
                } else {
                    issue(1, (cnt - 1) & 3, false);
                }
                if (leader) umma_commit(empty_bar(slot));
                __syncwarp();
                if (lane == 0) D2_STAMP(g, 5);
            }
            i += sb - sa + 1;
        }
    } else if (warp < 8) {
        // ================================ epilogue (eight warps) ================================
        // TMEM lane p < 64: hi plane of pixel p, lane 64+p: lo plane of pixel p; columns c and 64+c: W_hi / W_lo.  Two groups of four
        // warps, group eh = warp >> 2 does channels 32 eh .. +31 (the epilogue was the critical path with one warp per TMEM lane
        // quarter: 4,100 cycles per output row against 3,400 of the producers, clock64 timeline): warp w reads lane quarter w & 3.
        // Column halves are added in registers; the lo-plane warps (quarters 2, 3) park their sums in the group's 8 KB staging
        // tile and the hi-plane warps add theirs on top; then the group's 128 threads walk the tile with thread = (channel quad
        // cq of 8, pixel lane pr of 16): 8 threads cover 128 bytes of a pixel, so global loads / stores are whole 128-byte lines.
        const int eh = warp >> 2, ew = warp & 3, et = tid & 127;
        const int cq = et & 7, pr = et >> 3;
        const int p = et & 63;                                      // pixel of this thread's TMEM lane
        const int ch0 = eh * 32 + cq * 4;
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f}, me[4] = {0.f, 0.f, 0.f, 0.f}, iv[4] = {0.f, 0.f, 0.f, 0.f};
        if (EPI == EPI_MASK_BNBWD) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { sc[e] = s_bn[ch0 + e]; sh[e] = s_bn[64 + ch0 + e]; me[e] = s_bn[128 + ch0 + e]; iv[e] = s_bn[192 + ch0 + e]; }
        }
        float* tile = stg + eh * 2048;                              // the group's tile: 64 pixels x 32 channels
        float* row = tile + p * 32;                                 // this pixel's row: 8 chunks of 16 B, XOR-swizzled by pixel
        auto group_sync = [&]() { if (eh == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory"); };
        int it = 0;
        for (int i = i0; i < i1; ++i, ++it) {
            const int buf = it & 3;
            const size_t pix0 = (size_t)i * SW;                     // (n*SH + s)*SW
            float4 yp[4];
            if (EPI == EPI_MASK_BNBWD) {                            // pre-activations fetched before the accumulator is waited for
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (pr + 16 * k < SW) yp[k] = ldg4(e_ypre + (pix0 + pr + 16 * k) * SRLZ_C + ch0);
            }
            if (tid == 0) D2_STAMP(it, 6);
            mbar_wait(tfull_bar(buf), (it >> 2) & 1);
            tc_fence_after();
            if (tid == 0) D2_STAMP(it, 7);
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 128 + eh * 32;
            float v[32];
            {
                float w[32];
                tmem_ld32(taddr, v);
                tmem_ld32(taddr + 64, w);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] += w[e];
            }
            tc_fence_before();                                      // this warp's part of the accumulator is in registers
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
            if (ew >= 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(row + ((j ^ (p & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            group_sync();                                           // (A) the lo sums are in the tile
            if (ew < 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4* q4 = reinterpret_cast<float4*>(row + ((j ^ (p & 7)) << 2));
                    const float4 l4 = *q4;
                    *q4 = make_float4(v[4 * j] + l4.x, v[4 * j + 1] + l4.y, v[4 * j + 2] + l4.z, v[4 * j + 3] + l4.w);
                }
            }
            group_sync();                                           // (B) the finished tile
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int px = pr + 16 * k;
                if (px < SW) {
                    const float4 d4 = *reinterpret_cast<const float4*>(tile + px * 32 + ((cq ^ (px & 7)) << 2));
                    float d[4] = {d4.x, d4.y, d4.z, d4.w};
                    if (EPI == EPI_MASK_BNBWD) {
                        const float ypv[4] = {yp[k].x, yp[k].y, yp[k].z, yp[k].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool on = fmaf(ypv[e], sc[e], sh[e]) > 0.f;
                            const float dz = on ? d[e] : 0.f;
                            d[e] = dz;
                            st1[e] += dz;
                            st2[e] = fmaf(dz, (ypv[e] - me[e]) * iv[e], st2[e]);
                        }
                    }
                    st4(out + (pix0 + px) * SRLZ_C + ch0, make_float4(d[0], d[1], d[2], d[3]));
                }
            }
            group_sync();                                           // (C) the tile is rewritten by the next row
            if (tid == 0) D2_STAMP(it, 8);
        }
        if (EPI == EPI_MASK_BNBWD) {                                // per-thread sums -> [pixel lane pr of 16][128], folded in a fixed order below
            asm volatile("bar.sync 4, 256;" ::: "memory");         // both groups are done with their tiles (the sums overwrite them)
#pragma unroll
            for (int e = 0; e < 4; ++e) { stg[pr * 128 + ch0 + e] = st1[e]; stg[pr * 128 + 64 + ch0 + e] = st2[e]; }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (EPI == EPI_MASK_BNBWD && tid < 128) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) v += stg[q * 128 + tid];
        partials[(size_t)blockIdx.x * 128 + tid] = v;
    }
    if (warp == 4) tmem_dealloc(tmem_base, 512);
    if (EPI == EPI_MASK_BNBWD && tail.counter != nullptr) bn_tail_run(tail, partials, reinterpret_cast<double*>(smem), tid);   // (the row ring is free)
}

bool gconv64_s2rows_supported(const GConvArgs& a) {
    const ConvGeom& g = a.g;
    return !a.transposed && a.mode == 0 && g.KH == 3 && g.KW == 3 && g.stride == 2 && g.pad == 0 && g.BH == 2 * g.SH + 1 && g.BW == 2 * g.SW + 1 &&
           g.SW >= 1 && g.SW <= 55 && a.in_scale == nullptr && a.bias == nullptr && (a.epi == EPI_PLAIN || a.epi == EPI_MASK_BNBWD);
}

// a.in = dy (B,BH,BW,64), a.out = dz (B,SH,SW,64); EPI_MASK_BNBWD: a.e_* as in the other dgrad kernels, a.partials [n_partials][128];
// wbf = the layer's dgrad image from pack_conv_w_bf16 (9 taps x {hi 8 KB | lo 8 KB}, rows = ci, K = co)
int gconv64_s2rows(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st) {
    if (!gconv64_s2rows_supported(a)) { set_error("gconv64_s2rows: unsupported geometry"); return 1; }
    const int total = a.g.B * a.g.SH;
    int gx = sm_count();
    if (gx > total) gx = total;
    if (n_partials) *n_partials = gx;
    if (a.epi == EPI_MASK_BNBWD && (a.partials == nullptr || a.e_ypre == nullptr)) { set_error("gconv64_s2rows: partials / pre-activations required"); return 1; }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dgrad_s2_rows_kernel<EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, d2::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dgrad_s2_rows_kernel<EPI_MASK_BNBWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, d2::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gconv64_s2rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    const unsigned char* w = reinterpret_cast<const unsigned char*>(wbf);
    if (a.epi == EPI_MASK_BNBWD)
        launch_k(dgrad_s2_rows_kernel<EPI_MASK_BNBWD>, gx, d2::THREADS, d2::SMEM_BYTES, st, a.in, w, a.out, a.e_ypre, a.e_scale, a.e_shift, a.e_mean, a.e_invstd,
                                                                                     a.partials, a.g.SH, a.g.SW, total, a.dbg, a.tail);
    else
        launch_k(dgrad_s2_rows_kernel<EPI_PLAIN>, gx, d2::THREADS, d2::SMEM_BYTES, st, a.in, w, a.out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                                a.g.SH, a.g.SW, total, a.dbg, BnTail{});
    return check_launch("gconv64_s2rows");
}

}  // namespace srlz

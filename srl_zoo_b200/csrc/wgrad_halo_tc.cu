// Halo-tile tcgen05 weight-gradient kernel for the 3x3 64->64 layers (sm_100a):
//   conv3x3 s1 / s2 (models/models.py:54,59) and ConvTranspose2d k3 s2 (models/models.py:66-78).
//
//   P[tap][cg][cd] = sum over small-side pixels (n,sy,sx) of big[n, sy*s-pad+ky, sx*s-pad+kx, cg] * small[n,sy,sx,cd]
//
// The per-tap kernel (wgrad_tc.cu) gathers, converts and stores one 128x64 tile per tap and pixel block: ten staged
// tiles per 128 pixels, which is what bounds it.  Here a tile is R rows of the small side on a row pitch of HW pixels:
// the small ("dense") rows and the big-side rows they touch are converted to bf16 hi/lo and written to shared memory
// ONCE, as row images (one pixel = one 128 B SWIZZLE_128B row).  The reduction dimension K of the MMA is the pixel
// index, both operands are MN-major, and tap (ky,kx) is just the big image read through a descriptor whose start
// address is shifted by whole rows.  At stride 2 the big side is staged as one image per parity class (cy,cx) of
// (ky-pad, kx-pad), each with the same pitch, so that every tap is again a pure row shift inside its class image.
// Two taps of a class are stacked on M=128 through the descriptor's leading-dimension byte offset (LBO = row distance
// of the two shifts), which keeps the nine taps in five TMEM accumulators for the CTA's whole tile range.
// bf16x3 split (lo*hi + hi*lo + hi*hi, fp32 accumulate) as everywhere else; the kernel is bound by the operand fetch of its
// MMAs (an N=64 MMA streams 6 KB of shared memory), so the 512 TMEM columns are spent on cheaper forms where they reach:
//   pair, 3 MMAs,  64 columns: A_lo x B_hi + A_hi x B_lo + A_hi x B_hi                                   (3 x N=64)
//   pair, 2 MMAs, 128 columns: A_hi x [B_hi | B_lo] (one N=128 MMA, dense planes joined by the LBO) + A_lo x B_hi into
//                              the first half; the epilogue adds the two column halves                    (N=128 + N=64)
//   single tap, 2 MMAs, 64 columns: [A_hi ; A_lo] stacked on M x B_hi, then x B_lo: rows 0-63 = hh + hl, rows 64-127 =
//                              lh + ll, written as a tenth partial image that the reduction adds to the tap  (2 x N=64)
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace srlz {

namespace wh {
constexpr int BIG_ROWS = 256;                   // rows per bf16 plane of a class-image buffer
constexpr int BIG_PLANE = BIG_ROWS * 128;       // 32 KB
constexpr int BIG_BYTES = 2 * BIG_PLANE;        // hi | lo
constexpr int DEN_ROWS = 128;
constexpr int DEN_PLANE = DEN_ROWS * 128;       // 16 KB
constexpr int DEN_BYTES = 2 * DEN_PLANE;
constexpr int NB = 2, ND = 2;                   // class-image ring, dense ring
#ifndef SRLZ_WH_PF
#define SRLZ_WH_PF 1
#endif
#ifndef SRLZ_WH_PW
#define SRLZ_WH_PW 15
#endif
constexpr int PW = SRLZ_WH_PW, PT = PW * 32;    // producer warps / threads
constexpr int THREADS = (PW + 1) * 32;          // warp 4 MMA issuer | the other warps producers (warps 0-3 run the epilogue at the end)
constexpr int SMEM_BYTES = NB * BIG_BYTES + ND * DEN_BYTES + 1024 /*align*/ + 512 /*barriers*/;
constexpr int TMEM_COLS = 512;
constexpr int NSLOT = 10;                       // partial images per CTA: nine taps + the lo-row half of the single-tap form
// f32 accumulate, bf16 x bf16, A and B MN-major, M=128, N=64 / N=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC128 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
enum { FORM_PAIR3 = 0, FORM_PAIR2 = 1, FORM_SINGLE2 = 2 };
}  // namespace wh

struct WhClass { int by0, bx0, nrows, op0, nops; };   // big pixel of image (row i, col j): by = (sy0+i)*s + by0, bx = j*s + bx0
struct WhOp { int shift, lbo_rows, col, form; };      // tap pair: rows 0-63 of the accumulator (TMEM column `col`) = shift, rows 64-127 = shift + lbo_rows
struct WhPlan {
    int ncls, HW, R, ksteps, nrb, s, nacc;
    WhClass cls[4];
    WhOp ops[5];
    int acc_tap[5][2];   // partial-image slot of each accumulator row half: tap index ky*3+kx, 9 = lo rows of the single-tap form, -1 unused
    int extra_tap;       // the tap slot 9 belongs to (-1: none)
};

__device__ __forceinline__ uint64_t wh_desc(uint32_t saddr, uint32_t lbo_bytes) {   // MN-major SWIZZLE_128B, SBO = 1024
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

#ifdef SRLZ_DEV   // clock64 timeline of CTA 0 (tools/dev_timeline.py, sites 6 / 7): [16 tiles][64 slots]
#define WH_STAMP(t, slot) do { if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0 && (t) >= 0 && (t) < 16) a.dbg[(t) * 64 + (slot)] = clock64(); } while (0)
#else
#define WH_STAMP(t, slot) do { } while (0)
#endif

// Producers: a staged item is ONE 16-byte shared-memory chunk (8 channels of one pixel: a 32-byte global load, one hi and one
// lo store).  Thread i owns chunk (i & 7) of pixels (i >> 3) + k * PT/8, k < KR, of every unit, so that every producer warp has
// the same work whatever the unit's pixel count (with half-pixel items the first warps did everything and the scheduler slots of
// the others idled), the BN scale / shift of its 8 channels live in registers, and the pixel coordinates are computed once.
// The loads of unit u+1 are issued before unit u is converted and stored: the global-load latency of a unit overlaps the
// previous unit's work.
template <bool BN_DENSE, int KR>
__global__ void __launch_bounds__(wh::THREADS, 1) gwgrad64_halo_kernel(GWgradArgs a, WhPlan p, int total_tiles) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t big_base = base, den_base = base + wh::NB * wh::BIG_BYTES;
    constexpr uint32_t MISC = wh::NB * wh::BIG_BYTES + wh::ND * wh::DEN_BYTES;
    const uint32_t bars = base + MISC;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + MISC + 256);
    auto bfull = [&](int i) { return bars + 8u * i; };
    auto bempty = [&](int i) { return bars + 8u * (wh::NB + i); };
    auto dfull = [&](int i) { return bars + 8u * (2 * wh::NB + i); };
    auto dempty = [&](int i) { return bars + 8u * (2 * wh::NB + wh::ND + i); };
    const uint32_t acc_full = bars + 8u * (2 * wh::NB + 2 * wh::ND);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvGeom g = a.g;
    // contiguous, balanced tile range (consecutive tiles of a CTA share halo rows through L2)
    const int per = total_tiles / gridDim.x, extra = total_tiles % gridDim.x;
    const int t0 = blockIdx.x * per + min((int)blockIdx.x, extra);
    const int nt = per + ((int)blockIdx.x < extra ? 1 : 0);

    if (tid == 0) {
        for (int i = 0; i < wh::NB; ++i) { mbar_init(bfull(i), wh::PW); mbar_init(bempty(i), 1); }
        for (int i = 0; i < wh::ND; ++i) { mbar_init(dfull(i), wh::PW); mbar_init(dempty(i), 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    // rows the producers never write (beyond a unit's extent) are read by the MMAs against zero dense rows: keep them finite
    for (int e = tid; e < (int)(MISC / 16); e += wh::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), wh::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp != 4) {
        // ================================ producers ================================
        const int pidx = warp < 4 ? tid : tid - 32;
        const int c8 = pidx & 7;                         // this thread's 8-channel chunk of every pixel it stages
        const int U = 1 + p.ncls, NU = nt * U;
        // Per-thread statics.  Round k stages pixel q_k = (pidx >> 3) + k * PT/8 = image (row qr, column qx) of EVERY unit:
        //   rel_s[k] / rel_b[k]: float offset of that pixel from the unit's first pixel in the small / big tensor;
        //   hmask bit (4 kind + k): the item exists in the unit; xmask: ... and its column is inside the tensor.
        // For a unit whose rows are all inside the tensor (every tile but the first and last row block of an image) issuing a
        // load is then one predicate and one add; the general path keeps the per-item bounds checks.
        int qrx[KR], rel_s[KR], rel_b[KR];               // qrx: row << 8 | column
        uint32_t hmask = 0, xmask = 0;
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            const int q = (pidx >> 3) + k * (wh::PT / 8);
            const int r = q / p.HW, x = q - r * p.HW;
            qrx[k] = (r << 8) | x;
            rel_s[k] = (r * g.SW + x) * SRLZ_C;
            rel_b[k] = (r * g.BW + x) * p.s * SRLZ_C;
            if (q < p.R * p.HW) { hmask |= 1u << k; if (x < g.SW) xmask |= 1u << k; }
            for (int c = 0; c < p.ncls; ++c) {
                const int bx = x * p.s + p.cls[c].bx0;
                if (q < p.cls[c].nrows * p.HW) { hmask |= 1u << (4 * (c + 1) + k); if (bx >= 0 && bx < g.BW) xmask |= 1u << (4 * (c + 1) + k); }
            }
        }
        float4 sc0, sc1, sh0, sh1;
        if (BN_DENSE) {
            sc0 = *reinterpret_cast<const float4*>(a.dense_scale + c8 * 8); sc1 = *reinterpret_cast<const float4*>(a.dense_scale + c8 * 8 + 4);
            sh0 = *reinterpret_cast<const float4*>(a.dense_shift + c8 * 8); sh1 = *reinterpret_cast<const float4*>(a.dense_shift + c8 * 8 + 4);
        }
        int bs = 0, bph = 0, ds = 0, dph = 0;
        // cursor of the unit whose loads are issued next: (image n, first small row sy0, kind: 0 dense, 1.. class images)
        int i_n = t0 / p.nrb, i_sy0 = (t0 - i_n * p.nrb) * p.R, i_kind = 0, i_left = nt;   // i_left: tiles not yet fully issued
        // -> bit k: the item exists in this unit (its chunk must be written), bit 8+k: it is inside the tensor (else zeros)
        auto issue = [&](float4 (&v)[KR][2]) -> uint32_t {
            int nrows, by0, bx0, s, H, W;
            const float* src0;
            if (i_kind == 0) {
                nrows = p.R; by0 = 0; bx0 = 0; s = 1; H = g.SH; W = g.SW;
                src0 = a.small + (size_t)i_n * g.SH * g.SW * SRLZ_C + c8 * 8;
            } else {
                const WhClass cl = p.cls[i_kind - 1];
                nrows = cl.nrows; by0 = cl.by0; bx0 = cl.bx0; s = p.s; H = g.BH; W = g.BW;
                src0 = a.big + (size_t)i_n * g.BH * g.BW * SRLZ_C + c8 * 8;
            }
            const uint32_t hm = (hmask >> (4 * i_kind)) & 15u, xm = (xmask >> (4 * i_kind)) & 15u;
            const int ylo = i_sy0 * s + by0, yhi = (i_sy0 + nrows - 1) * s + by0;
            uint32_t m;
            if (ylo >= 0 && yhi < H) {                     // interior unit: every row inside the tensor
                const float* tb = src0 + ((ptrdiff_t)ylo * W + bx0) * SRLZ_C;
                m = hm | (xm << 8);
#pragma unroll
                for (int k = 0; k < KR; ++k)
                    if (xm & (1u << k)) ldg8(tb + (i_kind == 0 ? rel_s[k] : rel_b[k]), v[k][0], v[k][1]);
            } else {
                m = hm;
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const int y = (i_sy0 + (qrx[k] >> 8)) * s + by0;
                    if ((xm & (1u << k)) && y >= 0 && y < H) {
                        m |= 256u << k;
                        ldg8(src0 + ((size_t)y * W + (qrx[k] & 255) * s + bx0) * SRLZ_C, v[k][0], v[k][1]);
                    }
                }
            }
#if SRLZ_WH_PF
            // the same unit of the NEXT tile goes to L2 now (one prefetch per 128-byte line): the register loads above, which
            // are all the prefetch distance the register file allows, then see L2 latency instead of HBM latency
            if ((c8 & 3) == 0 && i_left > 1) {
                int sy2 = i_sy0 + p.R;
                const float* src2 = src0;
                if (sy2 >= p.nrb * p.R) { sy2 = 0; src2 += (size_t)H * W * SRLZ_C; }
                const int ylo2 = sy2 * s + by0, yhi2 = (sy2 + nrows - 1) * s + by0;
                if (ylo2 >= 0 && yhi2 < H) {               // (first / last row blocks of an image are not prefetched)
                    const float* tb = src2 + ((ptrdiff_t)ylo2 * W + bx0) * SRLZ_C;
#pragma unroll
                    for (int k = 0; k < KR; ++k)
                        if (xm & (1u << k)) prefetch_l2(tb + (i_kind == 0 ? rel_s[k] : rel_b[k]));
                }
            }
#endif
            if (++i_kind == U) {
                i_kind = 0;
                --i_left;
                i_sy0 += p.R;
                if (i_sy0 >= p.nrb * p.R) { i_sy0 = 0; ++i_n; }
            }
            return m;
        };
        auto finish = [&](int u, float4 (&v)[KR][2], const uint32_t m) {
            const int kind = u % U;
            const int so = (warp == 0 ? 0 : (warp == wh::PW ? 32 : 99)) + 3 * kind;
            if (so < 64) WH_STAMP(u / U, so);
            unsigned char* dst;
            uint32_t plane;
            if (kind == 0) {
                mbar_wait(dempty(ds), dph ^ 1);
                dst = smem + wh::NB * wh::BIG_BYTES + ds * wh::DEN_BYTES; plane = wh::DEN_PLANE;
            } else {
                mbar_wait(bempty(bs), bph ^ 1);
                dst = smem + bs * wh::BIG_BYTES; plane = wh::BIG_PLANE;
            }
            if (so < 64) WH_STAMP(u / U, so + 1);
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                if (m & (1u << k)) {
                    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                    if (m & (256u << k)) {
                        if (BN_DENSE && kind == 0) { v[k][0] = bn_relu4(v[k][0], sc0, sh0); v[k][1] = bn_relu4(v[k][1], sc1, sh1); }
                        split8(v[k][0], v[k][1], hi, lo);
                    }
                    const int row = (pidx >> 3) + k * (wh::PT / 8);
                    unsigned char* o = dst + row * 128 + ((c8 ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(o) = hi;
                    *reinterpret_cast<uint4*>(o + plane) = lo;
                }
            }
            __syncwarp();   // (proxy fence on the consumer side: here it would drain the next unit's loads in flight)
            if (kind == 0) {
                if (lane == 0) mbar_arrive(dfull(ds));
                if (++ds == wh::ND) { ds = 0; dph ^= 1; }
            } else {
                if (lane == 0) mbar_arrive(bfull(bs));
                if (++bs == wh::NB) { bs = 0; bph ^= 1; }
            }
            if (so < 64) WH_STAMP(u / U, so + 2);
        };
        float4 va[KR][2], vb[KR][2];
        uint32_t ma = issue(va), mb = 0;
        for (int u = 0; u < NU; u += 2) {
            if (u + 1 < NU) mb = issue(vb);
            finish(u, va, ma);
            if (u + 1 < NU) {
                if (u + 2 < NU) ma = issue(va);
                finish(u + 1, vb, mb);
            }
        }
    }
    if (warp == 4) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        int bs = 0, bph = 0, ds = 0, dph = 0;
        for (int t = 0; t < nt; ++t) {
            mbar_wait(dfull(ds), dph);
            WH_STAMP(t, 16);
            const uint32_t dsb = den_base + ds * wh::DEN_BYTES;
            for (int c = 0; c < p.ncls; ++c) {
                mbar_wait(bfull(bs), bph);
                // consumer-side proxy fence: the producers' st.shared are ordered before this point by the mbarrier (release /
                // acquire); fencing here keeps MEMBAR.ALL (what fence.proxy.async lowers to) away from warps with loads in flight
                fence_proxy_async_smem();
                tc_fence_after();
                WH_STAMP(t, 17 + 2 * c);
                if (leader) {
                    const uint32_t bsb = big_base + bs * wh::BIG_BYTES;
                    for (int o = p.cls[c].op0; o < p.cls[c].op0 + p.cls[c].nops; ++o) {
                        const WhOp op = p.ops[o];
                        const uint32_t a0 = bsb + op.shift * 128;
                        const uint64_t bhi = wh_desc(dsb, 0), blo = wh_desc(dsb + wh::DEN_PLANE, 0);
                        const uint32_t d_tmem = tmem_base + op.col;
                        if (op.form == wh::FORM_PAIR3) {
                            const uint64_t ahi = wh_desc(a0, op.lbo_rows * 128), alo = wh_desc(a0 + wh::BIG_PLANE, op.lbo_rows * 128);
                            for (int k = 0; k < p.ksteps; ++k) {
                                const uint64_t adv = (uint64_t)((k * 2048) >> 4);   // 16 pixels (K rows) = two 1024 B groups
                                const uint32_t accf = (t > 0 || k > 0) ? 1u : 0u;
                                umma_bf16(d_tmem, alo + adv, bhi + adv, wh::IDESC, accf);
                                umma_bf16(d_tmem, ahi + adv, blo + adv, wh::IDESC, 1u);
                                umma_bf16(d_tmem, ahi + adv, bhi + adv, wh::IDESC, 1u);
                            }
                        } else if (op.form == wh::FORM_PAIR2) {
                            const uint64_t ahi = wh_desc(a0, op.lbo_rows * 128), alo = wh_desc(a0 + wh::BIG_PLANE, op.lbo_rows * 128);
                            const uint64_t bhl = wh_desc(dsb, wh::DEN_PLANE);        // [B_hi | B_lo] as one N=128 operand
                            for (int k = 0; k < p.ksteps; ++k) {
                                const uint64_t adv = (uint64_t)((k * 2048) >> 4);
                                const uint32_t accf = (t > 0 || k > 0) ? 1u : 0u;
                                umma_bf16(d_tmem, ahi + adv, bhl + adv, wh::IDESC128, accf);
                                umma_bf16(d_tmem, alo + adv, bhi + adv, wh::IDESC, 1u);
                            }
                        } else {
                            const uint64_t ahl = wh_desc(a0, wh::BIG_PLANE);          // [A_hi ; A_lo] stacked on M
                            for (int k = 0; k < p.ksteps; ++k) {
                                const uint64_t adv = (uint64_t)((k * 2048) >> 4);
                                const uint32_t accf = (t > 0 || k > 0) ? 1u : 0u;
                                umma_bf16(d_tmem, ahl + adv, bhi + adv, wh::IDESC, accf);
                                umma_bf16(d_tmem, ahl + adv, blo + adv, wh::IDESC, 1u);
                            }
                        }
                    }
                    umma_commit(bempty(bs));
                    if (c == p.ncls - 1) umma_commit(dempty(ds));
                }
                __syncwarp();
                WH_STAMP(t, 18 + 2 * c);
                if (++bs == wh::NB) { bs = 0; bph ^= 1; }
            }
            if (++ds == wh::ND) { ds = 0; dph ^= 1; }
        }
        if (leader) umma_commit(acc_full);
        __syncwarp();
    } else if (warp < 4) {
        // ================================ epilogue (warps 0-3) ================================
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float* dstp = a.partials + (size_t)blockIdx.x * (wh::NSLOT * SRLZ_C * SRLZ_C);
        const int row = tid;  // TMEM lane = accumulator row: rows 0-63 first tap of the pair, rows 64-127 second
#pragma unroll 1
        for (int acc = 0; acc < p.nacc; ++acc) {
            const int slot = p.acc_tap[acc][row >> 6], cg = row & 63;
            const uint32_t col = p.ops[acc].col;
            const bool two = p.ops[acc].form == wh::FORM_PAIR2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + col + h * 32, v);
                if (two) {
                    float w[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + col + 64 + h * 32, w);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += w[j];
                }
                if (slot >= 0) {
                    float* o = dstp + ((size_t)slot * SRLZ_C + cg) * SRLZ_C + h * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) st4(o + j * 4, make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) tmem_dealloc(tmem_base, wh::TMEM_COLS);
}

// producer rounds per unit: 16-byte chunk items of the largest unit over the producer threads
static int wh_rounds(const WhPlan& p) {
    int maxpx = p.R * p.HW;
    for (int c = 0; c < p.ncls; ++c) if (p.cls[c].nrows * p.HW > maxpx) maxpx = p.cls[c].nrows * p.HW;
    return (maxpx * 8 + wh::PT - 1) / wh::PT;
}

static int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Builds the class / tap-pair plan; false when the geometry does not fit the halo kernel's buffers.
static bool make_wh_plan(const ConvGeom& g, WhPlan& p) {
    if (g.KH != 3 || g.KW != 3) return false;
    const int s = g.stride;
    if (s != 1 && s != 2) return false;
    int cy[3], oy[3];
    for (int k = 0; k < 3; ++k) {
        const int t = k - g.pad;
        oy[k] = floordiv(t, s);
        cy[k] = t - oy[k] * s;
    }
    // column extent of every class: HW = SW + widest column-offset range
    int maxrange = 0;
    for (int c = 0; c < s; ++c) {
        int mn = 99, mx = -99;
        for (int k = 0; k < 3; ++k) if (cy[k] == c) { if (oy[k] < mn) mn = oy[k]; if (oy[k] > mx) mx = oy[k]; }
        if (mx >= mn && mx - mn > maxrange) maxrange = mx - mn;
    }
    p.s = s;
    p.HW = g.SW + maxrange;
    if (p.HW > 128) return false;
    p.R = 128 / p.HW;
    if (p.R > g.SH) p.R = g.SH;
    if (p.R < 1) return false;
    p.ksteps = (p.R * p.HW + 15) / 16;
    p.nrb = (g.SH + p.R - 1) / p.R;
    p.ncls = 0;
    int nops = 0;
    for (int c_y = 0; c_y < s; ++c_y)
        for (int c_x = 0; c_x < s; ++c_x) {
            int shifts[9], taps[9], n = 0, mny = 99, mxy = -99, mnx = 99;
            for (int ky = 0; ky < 3; ++ky) if (cy[ky] == c_y) { if (oy[ky] < mny) mny = oy[ky]; if (oy[ky] > mxy) mxy = oy[ky]; }
            for (int kx = 0; kx < 3; ++kx) if (cy[kx] == c_x) { if (oy[kx] < mnx) mnx = oy[kx]; }
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx)
                    if (cy[ky] == c_y && cy[kx] == c_x) { shifts[n] = (oy[ky] - mny) * p.HW + (oy[kx] - mnx); taps[n] = ky * 3 + kx; ++n; }
            if (n == 0) continue;
            for (int i = 1; i < n; ++i)   // insertion sort by shift
                for (int j = i; j > 0 && shifts[j] < shifts[j - 1]; --j) {
                    int t = shifts[j]; shifts[j] = shifts[j - 1]; shifts[j - 1] = t;
                    t = taps[j]; taps[j] = taps[j - 1]; taps[j - 1] = t;
                }
            WhClass& cl = p.cls[p.ncls++];
            cl.by0 = mny * s + c_y; cl.bx0 = mnx * s + c_x; cl.nrows = p.R + (mxy - mny); cl.op0 = nops; cl.nops = 0;
            if (cl.nrows * p.HW > wh::BIG_ROWS) return false;
            for (int i = 0; i < n; i += 2) {
                if (nops >= 5) return false;
                const bool pair = i + 1 < n;
                p.ops[nops] = WhOp{shifts[i], pair ? shifts[i + 1] - shifts[i] : 0, 0, wh::FORM_PAIR3};
                p.acc_tap[nops][0] = taps[i];
                p.acc_tap[nops][1] = pair ? taps[i + 1] : -1;
                const int max_shift = pair ? shifts[i + 1] : shifts[i];
                if (max_shift + p.ksteps * 16 > wh::BIG_ROWS) return false;
                ++nops; ++cl.nops;
            }
        }
    // forms: the first single tap takes the stacked hi/lo form, then as many pairs as the TMEM columns allow the 2-MMA form
    p.extra_tap = -1;
    int spare = wh::TMEM_COLS - 64 * nops;
    for (int o = 0; o < nops; ++o) {
        if (p.acc_tap[o][1] < 0 && p.extra_tap < 0) { p.ops[o].form = wh::FORM_SINGLE2; p.extra_tap = p.acc_tap[o][0]; p.acc_tap[o][1] = 9; }
    }
    for (int o = 0; o < nops; ++o) {
        if (p.ops[o].form == wh::FORM_PAIR3 && p.acc_tap[o][1] >= 0 && spare >= 64) { p.ops[o].form = wh::FORM_PAIR2; spare -= 64; }
    }
    for (int o = 0, col = 0; o < nops; ++o) { p.ops[o].col = col; col += p.ops[o].form == wh::FORM_PAIR2 ? 128 : 64; }
    p.nacc = nops;
    return p.R * p.HW <= wh::DEN_ROWS && p.ksteps * 16 <= wh::DEN_ROWS && wh_rounds(p) <= 4;
}

bool gwgrad64_halo_supported(const ConvGeom& g) {
    WhPlan p;
    return make_wh_plan(g, p);
}

template <bool BN, int KR>
static int launch_wh(const GWgradArgs& a, const WhPlan& p, int total, int gx, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gwgrad64_halo_kernel<BN, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, wh::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gwgrad64_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    launch_k(gwgrad64_halo_kernel<BN, KR>, gx, wh::THREADS, wh::SMEM_BYTES, st, a, p, total);
    return check_launch("gwgrad64_halo");
}

// out[(cd*64 + cg)*9 + tap] (+)= sum_cta partials[cta][tap][cg][cd] (+ partials[cta][9][cg][cd] for tap == extra_tap)   (torch OIHW / IOHW
// layout).  Eight lanes per group of four outputs: lane q adds the partials q, q+8, q+16, ... (128-bit loads, independent of
// each other: 24 MB through L2 with 19 loads per thread instead of 148), the eight lanes are folded in a fixed order.
__global__ void __launch_bounds__(256) gwgrad64_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out, int nparts, int extra_tap,
                                                              int accumulate) {
    pdl_enter();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int idx = (gid >> 3) * 4, q = gid & 7;      // idx over tap*4096 + cg*64 + cd, four cd per group of eight lanes
    constexpr int total = 9 * SRLZ_C * SRLZ_C, stride = wh::NSLOT * SRLZ_C * SRLZ_C;
    if (idx >= total) return;
    const int tap = idx / (SRLZ_C * SRLZ_C);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add = [&](const float* p) {
        const float4 v = ldg4(p);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    };
#pragma unroll 5
    for (int c = q; c < nparts; c += 8) add(partials + (size_t)c * stride + idx);
    if (tap == extra_tap) {
        const float* e = partials + 9 * SRLZ_C * SRLZ_C + (idx - tap * SRLZ_C * SRLZ_C);
#pragma unroll 5
        for (int c = q; c < nparts; c += 8) add(e + (size_t)c * stride);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
    }
    if (q == 0) {
        const int cg = (idx / SRLZ_C) % SRLZ_C, cd = idx % SRLZ_C;
        const float r[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = ((cd + e) * SRLZ_C + cg) * 9 + tap;
            out[o] = accumulate ? out[o] + r[e] : r[e];
        }
    }
}

// one [10][64][64] partial per CTA (<= #SMs CTAs)
size_t gwgrad64_partial_floats(const ConvGeom& g) { (void)g; return (size_t)sm_count() * wh::NSLOT * SRLZ_C * SRLZ_C; }

int gwgrad64_halo(const GWgradArgs& a, float* grad_out, int accumulate, cudaStream_t st) {
    WhPlan p;
    if (!make_wh_plan(a.g, p)) { set_error("gwgrad64_halo: unsupported geometry"); return 1; }
    const int total = a.g.B * p.nrb;
    int gx = sm_count();
    if (gx > total) gx = total;
    const int kr = wh_rounds(p);
    const bool bn = a.dense_scale != nullptr;
    int rc;
    if (kr <= 2) rc = bn ? launch_wh<true, 2>(a, p, total, gx, st) : launch_wh<false, 2>(a, p, total, gx, st);
    else if (kr == 3) rc = bn ? launch_wh<true, 3>(a, p, total, gx, st) : launch_wh<false, 3>(a, p, total, gx, st);
    else rc = bn ? launch_wh<true, 4>(a, p, total, gx, st) : launch_wh<false, 4>(a, p, total, gx, st);
    if (rc) return rc;
    launch_k(gwgrad64_reduce_kernel, (9 * SRLZ_C * SRLZ_C * 2 + 255) / 256, 256, 0, st, a.partials, grad_out, gx, p.extra_tap, accumulate);   // 8 lanes x 9216 output groups
    return check_launch("gwgrad64_reduce");
}

}  // namespace srlz

// Halo-tile tcgen05 implicit GEMM for the stride-1 / per-parity-class 64->64 layers (sm_100a):
//   conv3x3 s1 forward + dgrad (models/models.py:54), conv3x3 s2 dgrad (:59), ConvTranspose2d k3 s2 forward (:66-78).
//
// Every gathered pixel is loaded from HBM/L2, BN+ReLU'd, split to bf16 hi/lo and written to shared memory ONCE per
// tile, as a row image (one pixel = one 128 B SWIZZLE_128B row, row pitch HW pixels, zero border).  Each tap is then
// served to tcgen05.mma by a K-major descriptor whose start address is shifted by (dy*HW + dx) rows: output row
// m = r*HW + x reads image row m + shift, so a 128-row MMA covers R = 128/HW output rows (columns x >= width are
// discarded).  (tests/gpu_probe.py verifies that shifted, not-1024B-aligned descriptors read the absolute-address
// swizzle correctly.)  All nine 64x64 weight taps (bf16 hi/lo, 144 KB) stay resident in shared memory.
//
// Pipelining with a single image buffer: per image-row mbarriers.  Taps are issued in groups by their row offset g;
// group g needs image rows [g, g+R) and releases row g when it retires, so the producers refill the buffer for the
// next tile while the later groups are still running.  Accumulators (one per output parity class) are double
// buffered in TMEM so the epilogue overlaps the next tile.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "bn_tail.cuh"

namespace srlz {

namespace hl {
constexpr int ROWS = 248;                        // image rows per bf16 plane (conv3x3 s1 at 56x56: rows 118 .. 245 for the last tap)
constexpr int PLANE = ROWS * 128;                // 31 KB
constexpr int W_BYTES = 9 * 2 * 64 * 128;        // 9 taps x (hi + lo) x 8 KB = 144 KB
constexpr int MAXNR = 8;
constexpr int THREADS = 16 * 32;   // warpgroups: 0 epilogue | 1 MMA issuer (warp 4) + 3 register-donor warps | 2,3 producers
constexpr int SMEM_BYTES = 2 * PLANE + W_BYTES + 1024 /*align*/ + 256 /*barriers, TMEM pointer*/ + 6 * 64 * 4 + 4 * 128 * 4 + 4 * 4096 /*epilogue staging*/;
}  // namespace hl

struct HaloOp { int shift, tap, cls, group; };
struct HaloPlan {
    int ncls, out_s;                 // parity classes, output coordinate step (1 or 2)
    int cls_py[4], cls_px[4], cls_oh[4], cls_ow[4];
    int GH, GW, OH, OW;              // gathered / output tensor extents
    int min_oy, min_ox, HW, R, NR, ngroups, nrb;
    int nops;
    HaloOp ops[9];
};

#ifndef SRLZ_HL_PF
#define SRLZ_HL_PF 1
#endif
#define HL_STAMP(slot) do { if (a.dbg != nullptr && blockIdx.x == 0 && it < 64) a.dbg[it * 16 + (slot)] = clock64(); } while (0)

// S2 (single-class geometries): bf16x3 in two MMAs per K step -- a tap's weight image is [hi 64 rows | lo 64 rows], so
// one N = 128 MMA yields A_hi*W_hi (columns 0-63) and A_hi*W_lo (columns 64-127) with a single read of the A tile, A_lo*W_hi
// is an N = 64 MMA into columns 0-63, and the epilogue adds the two column halves (128 TMEM columns per accumulator).
template <bool BN_LOAD, int EPI, bool S2 = false>
__global__ void __launch_bounds__(hl::THREADS, 1) gconv64_halo_kernel(GConvArgs a, HaloPlan p, const unsigned char* __restrict__ wbf,
                                                                      int total_tiles) {
    pdl_enter();
    constexpr int NOUT = 64;                                               // MMA N of the plain form
    constexpr uint32_t TAP_BYTES = 2 * NOUT * 128, LO_OFF = NOUT * 128;   // one tap: hi plane | lo plane (64 rows x 128 B each)
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)NOUT >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)(2 * NOUT) >> 3) << 17) | ((128u >> 4) << 24);
    constexpr int ACC_COLS = S2 ? 128 : 64;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t img = base;                           // hi plane, lo plane
    const uint32_t wsm = base + 2 * hl::PLANE;           // [tap]{hi 8 KB, lo 8 KB}
    const uint32_t bars = wsm + hl::W_BYTES;             // row_full[8], row_free[8], tfull[2], tempty[2], wfull
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 2 * hl::PLANE + hl::W_BYTES + 192);
    float* s_bn = reinterpret_cast<float*>(smem + 2 * hl::PLANE + hl::W_BYTES + 256);   // [4][64] epilogue consts (bias in row 0 for fwd)
    float* s_bnl = s_bn + 4 * 64;                                                        // [2][64] load-side scale, shift
    float* s_red = s_bnl + 2 * 64;                                                       // [4][128]
    auto row_full = [&](int j) { return bars + 8u * j; };
    auto row_free = [&](int j) { return bars + 8u * (hl::MAXNR + j); };
    auto tfull_bar = [&](int i) { return bars + 8u * (2 * hl::MAXNR + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (2 * hl::MAXNR + 2 + i); };
    const uint32_t wfull = bars + 8u * (2 * hl::MAXNR + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tmem_cols = p.ncls * ACC_COLS * 2 <= 128 ? 128 : (p.ncls * ACC_COLS * 2 <= 256 ? 256 : 512);

    if (tid == 0) {
        for (int j = 0; j < hl::MAXNR; ++j) { mbar_init(row_full(j), 8); mbar_init(row_free(j), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (tid < 64) {
        if (EPI == EPI_MASK_BNBWD) {
            s_bn[tid] = a.e_scale[tid]; s_bn[64 + tid] = a.e_shift[tid]; s_bn[128 + tid] = a.e_mean[tid]; s_bn[192 + tid] = a.e_invstd[tid];
        } else {
            s_bn[tid] = a.bias != nullptr ? a.bias[tid] : 0.f;
        }
        if (BN_LOAD) { s_bnl[tid] = a.in_scale[tid]; s_bnl[64 + tid] = a.in_shift[tid]; }
    }
    // zero both image planes once: rows the producers never touch are read (into discarded output rows) by the MMAs
    for (int e = tid; e < 2 * hl::PLANE / 16; e += hl::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (tid == 0) {  // resident weights: 9 bulk copies of 16 KB
        constexpr int NCOPY = 9;
        mbar_arrive_expect_tx(wfull, NCOPY * 16384);
        for (int t = 0; t < NCOPY; ++t) bulk_g2s(wsm + t * 16384, wbf + (size_t)t * 16384, 16384, wfull);
    }

    if (warp >= 8) {
        // ================================ producers ================================
        const int pidx = tid - 256, pw = warp - 8;
        const int items_per_row = 2 * p.HW, nitems = p.NR * items_per_row;
        // Per-thread items (<= 2 half-pixel rows of 32 channels) are the same in every tile; only the source address and the
        // border test change.  The loads of item k of the NEXT tile are issued right after item k of the current tile has
        // been converted (its registers are free again), so a load has a whole tile period to land before it is used.
        float4 v[2][8];
        int irow[2], icol[2], ihalf[2];
        bool have[2], inb[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = pidx + 256 * k;
            have[k] = i < nitems;
            inb[k] = false;
            irow[k] = have[k] ? i / items_per_row : 0;
            const int rem = have[k] ? i % items_per_row : 0;
            icol[k] = rem >> 1;
            ihalf[k] = rem & 1;
        }
        auto issue_loads = [&](int tile, int k) {
            if (!have[k]) return;
            const int n = tile / p.nrb, y0 = (tile % p.nrb) * p.R;
            const int gy = y0 + p.min_oy + irow[k], gx = p.min_ox + icol[k];
            inb[k] = gy >= 0 && gy < p.GH && gx >= 0 && gx < p.GW;
            if (inb[k]) {
                const float* src = a.in + (((size_t)n * p.GH + gy) * p.GW + gx) * SRLZ_C + ihalf[k] * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j) ldg8(src + j * 8, v[k][2 * j], v[k][2 * j + 1]);
            }
        };
        // the item's 128-byte line of a later tile goes to L2 now: the register loads (one tile of prefetch distance at most)
        // then see L2 latency instead of HBM latency
        auto prefetch_tile = [&](int tile, int k) {
#if SRLZ_HL_PF
            if (!have[k] || tile >= total_tiles || p.ncls == 1) return;   // (measured: no gain for the single-class geometries)
            const int n = tile / p.nrb, y0 = (tile % p.nrb) * p.R;
            const int gy = y0 + p.min_oy + irow[k], gx = p.min_ox + icol[k];
            if (gy >= 0 && gy < p.GH && gx >= 0 && gx < p.GW) prefetch_l2(a.in + (((size_t)n * p.GH + gy) * p.GW + gx) * SRLZ_C + ihalf[k] * 32);
#endif
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            if (pidx == 0) HL_STAMP(0);
            // (with a single image buffer the producers wait on the MMAs anyway: loads at the top of the tile are as good as a
            // rolling register prefetch, measured)
            issue_loads(tile, 0);
            issue_loads(tile, 1);
            prefetch_tile(tile + (int)gridDim.x, 0);
            prefetch_tile(tile + (int)gridDim.x, 1);
            const int fph = it & 1;   // phase of the row barriers
            // Every producer warp arrives once per image row and tile.  A warp may only signal a row for THIS tile once the row's
            // previous use is over (row_free), even when it stores nothing there: otherwise its arrival could complete the
            // previous tile's phase in place of a slower warp's, and the MMAs would read a row that is still being written.
            // Rows are waited for and signalled ONE AT A TIME, in ascending order: row j+1 is released by a later tap group of
            // the previous tile than row j, and the first tap group of this tile must not wait for that.
            int arrived = 0;   // rows [0, arrived) waited for and signalled by this warp
            auto pass_rows = [&](int upto) {   // rows in which this warp stores nothing (any more)
                for (; arrived < upto; ++arrived) {
                    mbar_wait(row_free(arrived), fph ^ 1);
                    if (lane == 0) mbar_arrive(row_full(arrived));
                }
            };
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                // this warp's items of step k are i in [256k + 32pw, 256k + 32pw + 32); its next step starts 256 items later,
                // beyond the rows touched here (items_per_row <= 256 - 32)
                const int lo_i = 256 * k + 32 * pw;
                if (lo_i >= nitems) break;
                const int lo_row = lo_i / items_per_row;
                int hi_row = (lo_i + 31) / items_per_row;
                if (hi_row >= p.NR) hi_row = p.NR - 1;
                pass_rows(lo_row);
                for (int row = lo_row; row <= hi_row; ++row) {
                    mbar_wait(row_free(row), fph ^ 1);
                    if (pidx == 0 && row == lo_row) HL_STAMP(1 + 2 * k);
                    if (have[k] && irow[k] == row) {
                        const int srow = irow[k] * p.HW + icol[k];
                        unsigned char* dst = smem + srow * 128;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                            if (inb[k]) {
                                float4 x0 = v[k][2 * j], x1 = v[k][2 * j + 1];
                                if (BN_LOAD) {
                                    const int c = ihalf[k] * 32 + j * 8;
                                    x0 = bn_relu4(x0, *reinterpret_cast<const float4*>(s_bnl + c), *reinterpret_cast<const float4*>(s_bnl + 64 + c));
                                    x1 = bn_relu4(x1, *reinterpret_cast<const float4*>(s_bnl + c + 4), *reinterpret_cast<const float4*>(s_bnl + 64 + c + 4));
                                }
                                split8(x0, x1, hi, lo);
                            }
                            const int chunk = (ihalf[k] * 4 + j) ^ (srow & 7);
                            *reinterpret_cast<uint4*>(dst + chunk * 16) = hi;
                            *reinterpret_cast<uint4*>(dst + hl::PLANE + chunk * 16) = lo;
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(row_full(row));
                    arrived = row + 1;
                }
                if (pidx == 0) HL_STAMP(2 + 2 * k);
            }
            pass_rows(p.NR);
        }
    } else if (warp >= 4) {
        // ================================ MMA issuer ================================
        if (warp == 4) {
        const bool leader = elect_one();
        mbar_wait(wfull, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (lane == 0) HL_STAMP(5);
            mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) HL_STAMP(6);
            const int fph = it & 1;
            uint32_t fresh = 0xFu;  // per-class "first MMA of this tile" flags
            int rows_ready = 0;
            for (int o = 0; o < p.nops; ++o) {
                const HaloOp op = p.ops[o];
                const int need = op.group + p.R;
                for (; rows_ready < need; ++rows_ready) mbar_wait(row_full(rows_ready), fph);
                tc_fence_after();
                if (lane == 0 && (o == 0 || p.ops[o - 1].group != op.group) && op.group < 3) HL_STAMP(7 + op.group);
                if (leader) {
                    const uint32_t a_hi = img + op.shift * 128, a_lo = a_hi + hl::PLANE;
                    const uint32_t w_hi = wsm + op.tap * TAP_BYTES, w_lo = w_hi + LO_OFF;
                    const uint64_t ahi = make_desc_sw128(a_hi), alo = make_desc_sw128(a_lo);
                    const uint64_t whi = make_desc_sw128(w_hi), wlo = make_desc_sw128(w_lo);
                    const uint32_t d_tmem = tmem_base + (buf * p.ncls + op.cls) * ACC_COLS;
                    uint32_t first = (fresh >> op.cls) & 1u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        if (S2) {
                            umma_bf16(d_tmem, ahi + adv, whi + adv, IDESC2, first ? 0u : 1u);   // [W_hi | W_lo]: 2*NOUT rows from w_hi
                            first = 0;
                            umma_bf16(d_tmem, alo + adv, whi + adv, IDESC, 1u);
                        } else {
                            umma_bf16(d_tmem, alo + adv, whi + adv, IDESC, first ? 0u : 1u);
                            first = 0;
                            umma_bf16(d_tmem, ahi + adv, wlo + adv, IDESC, 1u);
                            umma_bf16(d_tmem, ahi + adv, whi + adv, IDESC, 1u);
                        }
                    }
                    const bool last_of_group = (o + 1 == p.nops) || (p.ops[o + 1].group != op.group);
                    if (last_of_group) {
                        umma_commit(row_free(op.group));
                        if (o + 1 == p.nops) {
                            for (int j = op.group + 1; j < p.NR; ++j) umma_commit(row_free(j));
                            umma_commit(tfull_bar(buf));
                        }
                    }
                }
                fresh &= ~(1u << op.cls);
                __syncwarp();
            }
            if (lane == 0) HL_STAMP(10);
        }
        }
    } else {
        // ================================ epilogue (warps 0-3) ================================
        // tcgen05.ld hands thread r of a warp accumulator row r.  Thirty-two channels at a time, the warp's 32 rows are staged
        // through shared memory (32 rows x 128 B, 16-byte chunks XOR-swizzled by row) and read back with lane l = channels
        // 4*(l&7).. of row 4i+(l>>3), so that every global access writes whole 128-byte lines of four pixels instead of 16 B
        // of 32.  BatchNorm sums stay per thread (8 channels, fixed row order) and are folded across lanes once at the end.
        // (Measured alternative: the fragment-layout load tcgen05.ld.16x256b, tools/tmem_ld_probe.cu, lets a quad store one
        // 32-byte sector per row straight from registers; the epilogue itself got 17 % faster, but twice as many, 8-byte
        // scattered stores slowed the producers and the MMA stream sharing the memory pipe: dec9.fwd 0.57 -> 0.65 ms.
        // Eight epilogue warps -- quarter x channel half, as in the dec12 dgrad kernel -- changed nothing here: ncu puts the
        // L1/shared-memory pipe at 69 % for the four-class layers (MMA operand fetch 648 KB + staging 230 KB + stores 115 KB +
        // producer stores 86 KB per tile ~ 8,400 of the 10,000 cycles at 128 B/clk): the kernel is bound by that pipe.)
        {
        unsigned char* stg = reinterpret_cast<unsigned char*>(s_red + 4 * 128) + warp * 4096;
        const int c8 = lane & 7, rsub = lane >> 3;
        float st1[8], st2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { st1[i] = 0.f; st2[i] = 0.f; }
        const int r = tid / p.HW, x = tid % p.HW;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int n = tile / p.nrb, y0 = (tile % p.nrb) * p.R;
            if (tid == 0) HL_STAMP(11);
            mbar_wait(tfull_bar(buf), (it >> 1) & 1);
            tc_fence_after();
            if (tid == 0) HL_STAMP(12);
            for (int c = 0; c < p.ncls; ++c) {
                const int yc = y0 + r;
                const bool mvalid = r < p.R && x < p.cls_ow[c] && yc < p.cls_oh[c];
                const int mypix = mvalid ? (n * p.OH + (yc * p.out_s + p.cls_py[c])) * p.OW + (x * p.out_s + p.cls_px[c]) : -1;
                int rowpix[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) rowpix[i] = __shfl_sync(0xffffffffu, mypix, 4 * i + rsub);
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (buf * p.ncls + c) * ACC_COLS;
                const bool last_acc = c == p.ncls - 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ch0 = h * 32 + c8 * 4;
                    {
                        float v[32];
                        tmem_ld32(taddr + h * 32, v);
                        if (S2) {
                            float v2[32];
                            tmem_ld32(taddr + 64 + h * 32, v2);
#pragma unroll
                            for (int e = 0; e < 32; ++e) v[e] += v2[e];
                        }
                        if (h == 1 && last_acc) {  // accumulator fully in registers: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(tempty_bar(buf));
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    __syncwarp();
                    const float4 k0 = *reinterpret_cast<const float4*>(s_bn + ch0);          // scale | bias
                    float4 k1 = k0, k2 = k0, k3 = k0;
                    if (EPI == EPI_MASK_BNBWD) {
                        k1 = *reinterpret_cast<const float4*>(s_bn + 64 + ch0);              // shift
                        k2 = *reinterpret_cast<const float4*>(s_bn + 128 + ch0);             // mean
                        k3 = *reinterpret_cast<const float4*>(s_bn + 192 + ch0);             // invstd
                    }
                    const float sc[4] = {k0.x, k0.y, k0.z, k0.w}, sh[4] = {k1.x, k1.y, k1.z, k1.w};
                    const float me[4] = {k2.x, k2.y, k2.z, k2.w}, iv[4] = {k3.x, k3.y, k3.z, k3.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = 4 * i + rsub;
                        const float4 d4 = *reinterpret_cast<const float4*>(stg + row * 128 + ((c8 ^ (row & 7)) << 4));
                        const bool valid = rowpix[i] >= 0;
                        float d[4] = {d4.x, d4.y, d4.z, d4.w};
                        if (EPI == EPI_MASK_BNBWD) {
                            const float4 yp = valid ? ldg4(a.e_ypre + (size_t)rowpix[i] * SRLZ_C + ch0) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float ypv[4] = {yp.x, yp.y, yp.z, yp.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const bool on = valid && fmaf(ypv[e], sc[e], sh[e]) > 0.f;
                                const float dz = on ? d[e] : 0.f;
                                d[e] = dz;
                                st1[h * 4 + e] += dz;
                                st2[h * 4 + e] = fmaf(dz, (ypv[e] - me[e]) * iv[e], st2[h * 4 + e]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float y = valid ? d[e] + sc[e] : 0.f;
                                d[e] = y;
                                if (EPI == EPI_STATS) {
                                    st1[h * 4 + e] += y;
                                    st2[h * 4 + e] = fmaf(y, y, st2[h * 4 + e]);
                                }
                            }
                        }
                        if (valid) st4(a.out + (size_t)rowpix[i] * SRLZ_C + ch0, make_float4(d[0], d[1], d[2], d[3]));
                    }
                    __syncwarp();   // the staging rows are rewritten by the next pass
                }
            }
            if (tid == 0) HL_STAMP(13);
        }
        if (EPI != EPI_PLAIN) {
            // lanes with equal (lane & 7) hold the same 8 channels for different rows: fold the 4 row groups in a fixed order
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                    st1[i] += __shfl_xor_sync(0xffffffffu, st1[i], o);
                    st2[i] += __shfl_xor_sync(0xffffffffu, st2[i], o);
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s_red[warp * 128 + h * 32 + c8 * 4 + e] = st1[h * 4 + e];
                        s_red[warp * 128 + 64 + h * 32 + c8 * 4 + e] = st2[h * 4 + e];
                    }
            }
        }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (EPI != EPI_PLAIN && tid < 128) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) v += s_red[w * 128 + tid];
        a.partials[(size_t)blockIdx.x * 128 + tid] = v;
    }
    if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
    if (EPI != EPI_PLAIN && a.tail.counter != nullptr) bn_tail_run(a.tail, a.partials, reinterpret_cast<double*>(smem), tid);   // (image planes are free)
}

// Builds the tap / class / shift plan; returns false when the geometry does not fit the halo kernel.
static bool make_plan(const GConvArgs& a, HaloPlan& p) {
    const ConvGeom& g = a.g;
    if (g.KH != 3 || g.KW != 3) return false;
    const int s = g.stride;
    if (!a.transposed && s != 1) return false;   // direct stride-2 gathers are not unit-stride in the gathered image
    if (s != 1 && s != 2) return false;
    const int OH = a.transposed ? g.BH : g.SH, OW = a.transposed ? g.BW : g.SW;
    p.GH = a.transposed ? g.SH : g.BH;
    p.GW = a.transposed ? g.SW : g.BW;
    p.OH = OH; p.OW = OW;
    p.out_s = a.transposed ? s : 1;
    p.ncls = p.out_s * p.out_s;
    int oy[9], ox[9], tap[9], cls[9], n = 0, maxow = 0, maxoh = 0;
    for (int c = 0; c < p.ncls; ++c) {
        const int py = c / p.out_s, px = c % p.out_s;
        p.cls_py[c] = py; p.cls_px[c] = px;
        p.cls_oh[c] = (OH - py + p.out_s - 1) / p.out_s;
        p.cls_ow[c] = (OW - px + p.out_s - 1) / p.out_s;
        if (p.cls_ow[c] > maxow) maxow = p.cls_ow[c];
        if (p.cls_oh[c] > maxoh) maxoh = p.cls_oh[c];
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                int dy, dx;
                if (a.transposed) {
                    const int ny = py + g.pad - ky, nx = px + g.pad - kx;
                    if (((ny % s) + s) % s != 0 || ((nx % s) + s) % s != 0) continue;
                    dy = ny >= 0 ? ny / s : -((-ny) / s);
                    dx = nx >= 0 ? nx / s : -((-nx) / s);
                } else {
                    dy = ky - g.pad;
                    dx = kx - g.pad;
                }
                if (n >= 9) return false;
                oy[n] = dy; ox[n] = dx; tap[n] = ky * 3 + kx; cls[n] = c;
                ++n;
            }
    }
    if (n == 0) return false;
    int mny = oy[0], mxy = oy[0], mnx = ox[0], mxx = ox[0];
    for (int i = 1; i < n; ++i) {
        if (oy[i] < mny) mny = oy[i];
        if (oy[i] > mxy) mxy = oy[i];
        if (ox[i] < mnx) mnx = ox[i];
        if (ox[i] > mxx) mxx = ox[i];
    }
    p.min_oy = mny; p.min_ox = mnx;
    p.HW = maxow + (mxx - mnx);
    p.ngroups = mxy - mny + 1;
    if (p.HW > 112) return false;   // (2 * HW half-pixel items per row; the producers need <= 224 per row)
    p.R = 128 / p.HW;
    if (p.R > maxoh) p.R = maxoh;
    p.NR = p.R + p.ngroups - 1;
    if (p.NR > hl::MAXNR) { p.R = hl::MAXNR - p.ngroups + 1; p.NR = hl::MAXNR; }
    if (p.R < 1) return false;
    const int max_shift = (p.ngroups - 1) * p.HW + (mxx - mnx);
    if (max_shift + 128 > hl::ROWS || p.NR * p.HW > hl::ROWS) return false;
    if (p.NR * 2 * p.HW > 2 * 256) return false;   // two load items per producer thread
    if (p.ncls * 64 * 2 > 512) return false;
    p.nrb = (maxoh + p.R - 1) / p.R;
    // ops ordered by row-offset group
    p.nops = 0;
    for (int grp = 0; grp < p.ngroups; ++grp)
        for (int i = 0; i < n; ++i)
            if (oy[i] - mny == grp) p.ops[p.nops++] = HaloOp{grp * p.HW + (ox[i] - mnx), tap[i], cls[i], grp};
    return true;
}

bool gconv64_halo_supported(const GConvArgs& a) {
    HaloPlan p;
    return make_plan(a, p);
}

template <bool BN, int EPI, bool S2 = false>
static int launch_halo(const GConvArgs& a, const HaloPlan& p, const unsigned char* wbf, int total, int gx, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gconv64_halo_kernel<BN, EPI, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, hl::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gconv64_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    launch_k(gconv64_halo_kernel<BN, EPI, S2>, gx, hl::THREADS, hl::SMEM_BYTES, st, a, p, wbf, total);
    return check_launch("gconv64_halo");
}

// W12[ci][co][ky][kx] -> bf16 image [shift d = dy*2+dx]{hi[16][64], lo[16][64]} (K-major SWIZZLE_128B rows of 128 B):
// row j = (py*2+px)*3 + co holds W12[ci][co][py+2dy][px+2dx] over ci; rows 12..15 are zero
__global__ void pack_dec12_fwd_bf16_kernel(const float* __restrict__ w12, unsigned char* __restrict__ dst) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // d*1024 + j*64 + ci
    if (idx >= 4 * 1024) return;
    const int d = idx >> 10, j = (idx >> 6) & 15, ci = idx & 63;
    float x = 0.f;
    if (j < 12) {
        const int pyx = j / 3, co = j % 3, ky = (pyx >> 1) + 2 * (d >> 1), kx = (pyx & 1) + 2 * (d & 1);
        x = w12[((ci * 3 + co) * 4 + ky) * 4 + kx];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const int byte = j * 128 + (((ci >> 3) ^ (j & 7)) << 4) + (ci & 7) * 2;
    unsigned char* t = dst + (size_t)d * 4096;
    *reinterpret_cast<__nv_bfloat16*>(t + byte) = hi;
    *reinterpret_cast<__nv_bfloat16*>(t + 2048 + byte) = lo;
}
int pack_dec12_fwd_bf16(const float* w12, void* dst, cudaStream_t st) {
    launch_k(pack_dec12_fwd_bf16_kernel, 16, 256, 0, st, w12, reinterpret_cast<unsigned char*>(dst));
    return check_launch("pack_dec12_fwd_bf16");
}

int gconv64_halo(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st) {
    HaloPlan p;
    if (!make_plan(a, p)) { set_error("gconv64_halo: unsupported geometry"); return 1; }
    const int total = a.g.B * p.nrb;
    int gx = sm_count();
    if (gx > total) gx = total;
    if (n_partials) *n_partials = gx;
    if (a.epi != EPI_PLAIN && a.partials == nullptr) { set_error("gconv64_halo: partials buffer required"); return 1; }
    const unsigned char* w = reinterpret_cast<const unsigned char*>(wbf);
    const bool bn = a.in_scale != nullptr;
    if (bn) {
        if (a.epi == EPI_PLAIN) return launch_halo<true, EPI_PLAIN>(a, p, w, total, gx, st);
        if (a.epi == EPI_STATS) return launch_halo<true, EPI_STATS>(a, p, w, total, gx, st);
        return launch_halo<true, EPI_MASK_BNBWD>(a, p, w, total, gx, st);
    }
    if (p.ncls == 1) {   // single-class geometries (conv3x3 s1 forward / dgrad): two-MMA form
        if (a.epi == EPI_PLAIN) return launch_halo<false, EPI_PLAIN, true>(a, p, w, total, gx, st);
        if (a.epi == EPI_STATS) return launch_halo<false, EPI_STATS, true>(a, p, w, total, gx, st);
    }
    if (a.epi == EPI_PLAIN) return launch_halo<false, EPI_PLAIN>(a, p, w, total, gx, st);
    if (a.epi == EPI_STATS) return launch_halo<false, EPI_STATS>(a, p, w, total, gx, st);
    return launch_halo<false, EPI_MASK_BNBWD>(a, p, w, total, gx, st);
}

}  // namespace srlz

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
// of the two shifts), which keeps the nine taps in five 64-column TMEM accumulators for the CTA's whole tile range.
// bf16x3 split (lo*hi + hi*lo + hi*hi, fp32 accumulate) as everywhere else.
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
constexpr int THREADS = 13 * 32;                // warp 4 MMA issuer | warps 0-3, 5-12 producers (warps 0-3 run the epilogue at the end)
constexpr int PW = 12, PT = PW * 32;            // producer warps / threads
constexpr int SMEM_BYTES = NB * BIG_BYTES + ND * DEN_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 2 * 64 * 4;
constexpr int TMEM_COLS = 512;                  // 5 accumulators x 64 columns -> next power of two
// f32 accumulate, bf16 x bf16, A and B MN-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
}  // namespace wh

struct WhClass { int by0, bx0, nrows, op0, nops; };   // big pixel of image (row i, col j): by = (sy0+i)*s + by0, bx = j*s + bx0
struct WhOp { int shift, lbo_rows, acc; };            // tap pair: rows 0-63 of the accumulator = shift, rows 64-127 = shift + lbo_rows
struct WhPlan {
    int ncls, HW, R, ksteps, nrb, s, nacc;
    WhClass cls[4];
    WhOp ops[5];
    int acc_tap[5][2];   // tap index ky*3+kx of each accumulator half (-1: unused)
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

// IPT = staged half-pixel items per producer thread and unit.  IPT == 1 (every unit fits the 384 producer threads): the
// loads of unit u+1 are issued before unit u is converted and stored, so the global-load latency of a unit overlaps the
// previous unit's work.  IPT == 2 (conv3x3 s1 at 56x56): load, wait, convert, store one unit at a time.
template <bool BN_DENSE, int IPT>
__global__ void __launch_bounds__(wh::THREADS, 1) gwgrad64_halo_kernel(GWgradArgs a, WhPlan p, int total_tiles) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t big_base = base, den_base = base + wh::NB * wh::BIG_BYTES;
    constexpr uint32_t MISC = wh::NB * wh::BIG_BYTES + wh::ND * wh::DEN_BYTES;
    const uint32_t bars = base + MISC;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + MISC + 256);
    float* s_bnl = reinterpret_cast<float*>(smem + MISC + 512);
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
    if (BN_DENSE && tid >= 64 && tid < 128) {
        s_bnl[tid - 64] = a.dense_scale[tid - 64];
        s_bnl[tid] = a.dense_shift[tid - 64];
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
        int bs = 0, bph = 0, ds = 0, dph = 0;
        // convert one half pixel row (32 channels) and write it to image row `row` of the buffer at `dst`
        auto store_item = [&](unsigned char* dst, uint32_t plane, int row, int half, const float4 (&v)[8], bool valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                if (valid) split8(v[2 * j], v[2 * j + 1], hi, lo);
                const int chunk = (half * 4 + j) ^ (row & 7);
                *reinterpret_cast<uint4*>(dst + row * 128 + chunk * 16) = hi;
                *reinterpret_cast<uint4*>(dst + plane + row * 128 + chunk * 16) = lo;
            }
        };
        if (IPT == 1) {
            const int q = pidx >> 1, half = pidx & 1;
            const int U = 1 + p.ncls, NU = nt * U;
            const int qr = q / p.HW, qx = q - qr * p.HW;     // image (row, column) of this thread's pixel in every unit
            struct Item { bool have, valid; };
            auto issue = [&](int u, float4 (&v)[8]) -> Item {
                const int t = u / U, kind = u - t * U, tile = t0 + t;
                const int n = tile / p.nrb, sy0 = (tile % p.nrb) * p.R;
                Item it{false, false};
                const float* src = nullptr;
                if (kind == 0) {
                    it.have = q < p.R * p.HW;
                    it.valid = it.have && qx < g.SW && sy0 + qr < g.SH;
                    if (it.valid) src = a.small + (((size_t)n * g.SH + sy0 + qr) * g.SW + qx) * SRLZ_C + half * 32;
                } else {
                    const WhClass cl = p.cls[kind - 1];
                    it.have = q < cl.nrows * p.HW;
                    const int by = (sy0 + qr) * p.s + cl.by0, bx = qx * p.s + cl.bx0;
                    it.valid = it.have && by >= 0 && by < g.BH && bx >= 0 && bx < g.BW;
                    if (it.valid) src = a.big + (((size_t)n * g.BH + by) * g.BW + bx) * SRLZ_C + half * 32;
                }
                if (it.valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ldg8(src + j * 8, v[2 * j], v[2 * j + 1]);
                }
                return it;
            };
            auto finish = [&](int u, float4 (&v)[8], const Item it) {
                const int kind = u % U;
                if (kind == 0) {
                    mbar_wait(dempty(ds), dph ^ 1);
                    if (it.have) {
                        if (BN_DENSE && it.valid) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                v[j] = bn_relu4(v[j], *reinterpret_cast<const float4*>(s_bnl + half * 32 + j * 4), *reinterpret_cast<const float4*>(s_bnl + 64 + half * 32 + j * 4));
                        }
                        store_item(smem + wh::NB * wh::BIG_BYTES + ds * wh::DEN_BYTES, wh::DEN_PLANE, q, half, v, it.valid);
                    }
                    __syncwarp();   // (proxy fence on the consumer side: here it would drain the next unit's loads in flight)
                    if (lane == 0) mbar_arrive(dfull(ds));
                    if (++ds == wh::ND) { ds = 0; dph ^= 1; }
                } else {
                    mbar_wait(bempty(bs), bph ^ 1);
                    if (it.have) store_item(smem + bs * wh::BIG_BYTES, wh::BIG_PLANE, q, half, v, it.valid);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bfull(bs));
                    if (++bs == wh::NB) { bs = 0; bph ^= 1; }
                }
            };
            float4 va[8], vb[8];
            Item ia = issue(0, va), ib{false, false};
            for (int u = 0; u < NU; u += 2) {
                if (u + 1 < NU) ib = issue(u + 1, vb);
                finish(u, va, ia);
                if (u + 1 < NU) {
                    if (u + 2 < NU) ia = issue(u + 2, va);
                    finish(u + 1, vb, ib);
                }
            }
        } else
        for (int t = 0; t < nt; ++t) {
            const int tile = t0 + t;
            const int n = tile / p.nrb, sy0 = (tile % p.nrb) * p.R;
            // ---- dense unit: R rows of the small side (pitch HW, columns >= SW zero) ----
            {
                const int q = pidx >> 1, half = pidx & 1;
                const bool have = q < p.R * p.HW;
                float4 v[8];
                bool valid = false;
                if (have) {
                    const int r = q / p.HW, x = q - r * p.HW;
                    valid = x < g.SW && sy0 + r < g.SH;
                    if (valid) {
                        const float* src = a.small + (((size_t)n * g.SH + sy0 + r) * g.SW + x) * SRLZ_C + half * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) ldg8(src + j * 8, v[2 * j], v[2 * j + 1]);
                    }
                }
                mbar_wait(dempty(ds), dph ^ 1);
                if (have) {
                    if (BN_DENSE && valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            v[j] = bn_relu4(v[j], *reinterpret_cast<const float4*>(s_bnl + half * 32 + j * 4), *reinterpret_cast<const float4*>(s_bnl + 64 + half * 32 + j * 4));
                    }
                    store_item(smem + wh::NB * wh::BIG_BYTES + ds * wh::DEN_BYTES, wh::DEN_PLANE, q, half, v, valid);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(dfull(ds));
                if (++ds == wh::ND) { ds = 0; dph ^= 1; }
            }
            // ---- one image per parity class of the big side ----
            for (int c = 0; c < p.ncls; ++c) {
                const WhClass cl = p.cls[c];
                const int npix = cl.nrows * p.HW;
                float4 v[2][8];
                int q[2];
                bool have[2], valid[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int i = pidx + wh::PT * k;
                    q[k] = i >> 1;
                    have[k] = q[k] < npix;
                    valid[k] = false;
                    if (have[k]) {
                        const int ir = q[k] / p.HW, j = q[k] - ir * p.HW;
                        const int by = (sy0 + ir) * p.s + cl.by0, bx = j * p.s + cl.bx0;
                        valid[k] = by >= 0 && by < g.BH && bx >= 0 && bx < g.BW;
                        if (valid[k]) {
                            const float* src = a.big + (((size_t)n * g.BH + by) * g.BW + bx) * SRLZ_C + (i & 1) * 32;
#pragma unroll
                            for (int j2 = 0; j2 < 4; ++j2) ldg8(src + j2 * 8, v[k][2 * j2], v[k][2 * j2 + 1]);
                        }
                    }
                }
                mbar_wait(bempty(bs), bph ^ 1);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (have[k]) store_item(smem + bs * wh::BIG_BYTES, wh::BIG_PLANE, q[k], (pidx + wh::PT * k) & 1, v[k], valid[k]);
                __syncwarp();
                if (lane == 0) mbar_arrive(bfull(bs));
                if (++bs == wh::NB) { bs = 0; bph ^= 1; }
            }
        }
    }
    if (warp == 4) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        int bs = 0, bph = 0, ds = 0, dph = 0;
        for (int t = 0; t < nt; ++t) {
            mbar_wait(dfull(ds), dph);
            const uint32_t dsb = den_base + ds * wh::DEN_BYTES;
            for (int c = 0; c < p.ncls; ++c) {
                mbar_wait(bfull(bs), bph);
                // consumer-side proxy fence: the producers' st.shared are ordered before this point by the mbarrier (release /
                // acquire); fencing here keeps MEMBAR.ALL (what fence.proxy.async lowers to) away from warps with loads in flight
                fence_proxy_async_smem();
                tc_fence_after();
                if (leader) {
                    const uint32_t bsb = big_base + bs * wh::BIG_BYTES;
                    for (int o = p.cls[c].op0; o < p.cls[c].op0 + p.cls[c].nops; ++o) {
                        const WhOp op = p.ops[o];
                        const uint64_t ahi = wh_desc(bsb + op.shift * 128, op.lbo_rows * 128), alo = wh_desc(bsb + wh::BIG_PLANE + op.shift * 128, op.lbo_rows * 128);
                        const uint64_t bhi = wh_desc(dsb, 0), blo = wh_desc(dsb + wh::DEN_PLANE, 0);
                        const uint32_t d_tmem = tmem_base + op.acc * 64;
                        for (int k = 0; k < p.ksteps; ++k) {
                            const uint64_t adv = (uint64_t)((k * 2048) >> 4);   // 16 pixels (K rows) = two 1024 B groups
                            const uint32_t accf = (t > 0 || k > 0) ? 1u : 0u;
                            umma_bf16(d_tmem, alo + adv, bhi + adv, wh::IDESC, accf);
                            umma_bf16(d_tmem, ahi + adv, blo + adv, wh::IDESC, 1u);
                            umma_bf16(d_tmem, ahi + adv, bhi + adv, wh::IDESC, 1u);
                        }
                    }
                    umma_commit(bempty(bs));
                    if (c == p.ncls - 1) umma_commit(dempty(ds));
                }
                __syncwarp();
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
        float* dstp = a.partials + (size_t)blockIdx.x * (9 * SRLZ_C * SRLZ_C);
        const int row = tid;  // TMEM lane = accumulator row: rows 0-63 first tap of the pair, rows 64-127 second
#pragma unroll 1
        for (int acc = 0; acc < p.nacc; ++acc) {
            const int tap = p.acc_tap[acc][row >> 6], cg = row & 63;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64 + h * 32, v);
                if (tap >= 0) {
                    float* o = dstp + ((size_t)tap * SRLZ_C + cg) * SRLZ_C + h * 32;
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
            if (cl.nrows * p.HW > wh::BIG_ROWS || cl.nrows * p.HW * 2 > 2 * wh::PT) return false;
            for (int i = 0; i < n; i += 2) {
                if (nops >= 5) return false;
                const bool pair = i + 1 < n;
                p.ops[nops] = WhOp{shifts[i], pair ? shifts[i + 1] - shifts[i] : 0, nops};
                p.acc_tap[nops][0] = taps[i];
                p.acc_tap[nops][1] = pair ? taps[i + 1] : -1;
                const int max_shift = pair ? shifts[i + 1] : shifts[i];
                if (max_shift + p.ksteps * 16 > wh::BIG_ROWS) return false;
                ++nops; ++cl.nops;
            }
        }
    p.nacc = nops;
    return p.R * p.HW * 2 <= 256 && p.ksteps * 16 <= wh::DEN_ROWS;
}

bool gwgrad64_halo_supported(const ConvGeom& g) {
    WhPlan p;
    return make_wh_plan(g, p);
}

template <bool BN, int IPT>
static int launch_wh(const GWgradArgs& a, const WhPlan& p, int total, int gx, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gwgrad64_halo_kernel<BN, IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, wh::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gwgrad64_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    gwgrad64_halo_kernel<BN, IPT><<<gx, wh::THREADS, wh::SMEM_BYTES, st>>>(a, p, total);
    return check_launch("gwgrad64_halo");
}

// out[(cd*64 + cg)*ntaps + tap] (+)= sum_cta partials[cta][tap][cg][cd]   (torch OIHW / IOHW layout), fixed order
__global__ void gwgrad64_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out, int nparts, int ntaps, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over tap*4096 + cg*64 + cd
    const int total = ntaps * SRLZ_C * SRLZ_C;
    if (idx >= total) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partials[(size_t)c * total + idx];
    const int tap = idx / (SRLZ_C * SRLZ_C);
    const int cg = (idx / SRLZ_C) % SRLZ_C, cd = idx % SRLZ_C;
    const int o = (cd * SRLZ_C + cg) * ntaps + tap;
    out[o] = accumulate ? out[o] + s : s;
}

// one [9][64][64] partial per CTA (<= #SMs CTAs)
size_t gwgrad64_partial_floats(const ConvGeom& g) { return (size_t)sm_count() * g.KH * g.KW * SRLZ_C * SRLZ_C; }

int gwgrad64_halo(const GWgradArgs& a, float* grad_out, int accumulate, cudaStream_t st) {
    WhPlan p;
    if (!make_wh_plan(a.g, p)) { set_error("gwgrad64_halo: unsupported geometry"); return 1; }
    const int total = a.g.B * p.nrb;
    int gx = sm_count();
    if (gx > total) gx = total;
    int maxitems = 0;
    for (int c = 0; c < p.ncls; ++c) if (p.cls[c].nrows * p.HW * 2 > maxitems) maxitems = p.cls[c].nrows * p.HW * 2;
    int rc;
    if (maxitems <= wh::PT) rc = a.dense_scale != nullptr ? launch_wh<true, 1>(a, p, total, gx, st) : launch_wh<false, 1>(a, p, total, gx, st);
    else rc = a.dense_scale != nullptr ? launch_wh<true, 2>(a, p, total, gx, st) : launch_wh<false, 2>(a, p, total, gx, st);
    if (rc) return rc;
    gwgrad64_reduce_kernel<<<(9 * SRLZ_C * SRLZ_C + 255) / 256, 256, 0, st>>>(a.partials, grad_out, gx, 9, accumulate);
    return check_launch("gwgrad64_reduce");
}

}  // namespace srlz

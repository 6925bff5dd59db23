// tcgen05 weight-gradient kernel for the 64->64 channel layers (sm_100a).
//
//   P[tap][cg][cd] = sum over pixels m of  gather(big)[m, tap, cg] * dense(small)[m, cd]
//
// Per block of 128 pixels the reduction dimension K is the pixel index, so both operands are MN-major in their natural
// NHWC form ([pixel][64 channels] = 128 B of bf16 per K row) -- the same SWIZZLE_128B image the forward kernel stages.
// Two taps are stacked along M (rows 0-63 = tap 2p, rows 64-127 = tap 2p+1; the second 64-wide MN block is one smem
// slot = LBO further), giving full M=128 MMAs; five such pairs cover the 9 taps and live in 5 TMEM accumulators
// (320 columns) for the whole pixel range of the CTA, so there is a single epilogue at the end.
// bf16x3 split (hi*hi + lo*hi + hi*lo) keeps fp32-level accuracy.  Roles / barriers as in conv_tc.cu.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace srlz {

namespace wg {
constexpr int TILE = 2 * 128 * 128;     // one staged 128x64 tile: bf16 hi plane + lo plane = 32 KB
constexpr int ND = 2;                   // dense-side (dy / layer input) ring
constexpr int NT = 4;                   // gathered-tap ring (pairs occupy slots (even, odd))
constexpr int SMEM_BYTES = (ND + NT) * TILE + 1024 + 512 + 2 * 64 * 4 + 2 * PATCH_MAX_FLOATS * 4;
constexpr int THREADS = 13 * 32;
constexpr int TMEM_COLS = 512;          // 5 accumulators x 64 columns -> next power of two
// f32 accumulate, bf16 x bf16, A and B MN-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
}  // namespace wg

// MN-major SWIZZLE_128B: 64 MN elements (128 B) contiguous per K row, 8-row groups SBO = 1024 B apart,
// next 64-wide MN block LBO bytes further.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <bool BN_DENSE, int MODE>
__global__ void __launch_bounds__(wg::THREADS, 1) gwgrad64_tc_kernel(GWgradArgs a, int nblocks) {
    constexpr int NPAIRS = MODE == 0 ? 5 : (MODE == 1 ? 2 : 1);   // tap pairs stacked on M=128
    constexpr int NTAPS = MODE == 0 ? 9 : (MODE == 1 ? 3 : 1);    // real gathered tiles per pixel block
    constexpr int NUNITS = 1 + 2 * NPAIRS;                          // MODE 0: dense tile + tap tiles (zero padded to pairs)
    // special modes: no zero tile -- an unpaired last tap is issued with LBO = 0 (rows 64-127 of its accumulator repeat
    // rows 0-63 and are never read); MODE 1 cycles its 3 tap tiles through 3 slots, MODE 2 its single tile through 2
    constexpr int NT = MODE == 0 ? wg::NT : (MODE == 1 ? 3 : 2);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t d_base = base;                       // ND dense tiles
    const uint32_t t_base = base + wg::ND * wg::TILE;    // NT tap tiles
    const uint32_t bars = base + (wg::ND + NT) * wg::TILE;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (wg::ND + NT) * wg::TILE + 256);
    float* s_bnl = reinterpret_cast<float*>(smem + (wg::ND + NT) * wg::TILE + 512);
    float* s_patch = s_bnl + 2 * 64;   // [2][PATCH_MAX_FLOATS] source patches of the special modes
    auto dfull = [&](int i) { return bars + 8u * i; };
    auto dempty = [&](int i) { return bars + 8u * (wg::ND + i); };
    auto tfull = [&](int i) { return bars + 8u * (2 * wg::ND + i); };
    auto tempty = [&](int i) { return bars + 8u * (2 * wg::ND + NT + i); };
    const uint32_t acc_full = bars + 8u * (2 * wg::ND + 2 * NT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvGeom g = a.g;
    const long long Ms = (long long)g.B * g.SH * g.SW;
    // contiguous, balanced range of 128-pixel blocks for this CTA (every CTA owns >= 1 block)
    const int per = nblocks / gridDim.x, extra = nblocks % gridDim.x;
    const int b0 = blockIdx.x * per + min((int)blockIdx.x, extra);
    const int nb = per + ((int)blockIdx.x < extra ? 1 : 0);

    if (tid == 0) {
        for (int i = 0; i < wg::ND; ++i) { mbar_init(dfull(i), 8); mbar_init(dempty(i), 1); }
        for (int i = 0; i < NT; ++i) { mbar_init(tfull(i), 8); mbar_init(tempty(i), 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (BN_DENSE && tid >= 64 && tid < 128) {
        s_bnl[tid - 64] = a.dense_scale[tid - 64];
        s_bnl[64 + tid - 64] = a.dense_shift[tid - 64];
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), wg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 5) {
        // ================================ producers ================================
        // unit sequence per pixel block: dense tile, tap 0..8, one zero tile (keeps tap pairs slot-aligned)
        if (MODE != 0) {
            // 8x16-pixel blocks: dense tile straight from global (prefetched one block ahead), gathered chunks from the
            // block's source patch staged in shared memory (double buffered, prefetched through registers)
            constexpr int M = MODE == 0 ? 1 : MODE;
            using PG = PatchGeom<M>;
            const int pidx = tid - 160, pix = pidx & 127, half = pidx >> 7, py = pix >> 4, px = pix & 15;
            const int GS = MODE == 1 ? 112 : 111;   // dense-side grid (dy1 / y7)
            const PatchSrc src{a.big, a.rects, a.aux0, a.aux1, a.aux2, a.coef};
            PatchIdx<M> pidx_tab;
            pidx_tab.init(pidx);
            const bool fused = MODE == 2 && a.aux0 == nullptr;
            const uint32_t pa = smem_u32(s_patch), pb = pa + 2 * PATCH_MAX_FLOATS * 4;
            if (nb > 0) {
                const int tt = b0 % 98;
                patch_prefetch<M>(pidx_tab, src, b0 / 98, (tt / 7) * 8, (tt % 7) * 16, pa, pb);
            }
            cp_async_wait_all();
            producers_bar_sync();
            int ds = 0, dph = 0, ts = 0, tph = 0;
            for (int blk = 0; blk < nb; ++blk) {
                const float* cur = s_patch + (blk & 1) * PATCH_MAX_FLOATS;
                const float* curB = cur + 2 * PATCH_MAX_FLOATS;
                const bool hn = blk + 1 < nb;
                if (hn) {
                    const int nbk = b0 + blk + 1, tt = nbk % 98;
                    const uint32_t nb_off = ((blk + 1) & 1) * PATCH_MAX_FLOATS * 4;
                    patch_prefetch<M>(pidx_tab, src, nbk / 98, (tt / 7) * 8, (tt % 7) * 16, pa + nb_off, pb + nb_off);
                }
                // dense tile (dy1 / BN+ReLU of y7) of this block
                {
                    const int bid = b0 + blk, tt = bid % 98, oy = (tt / 7) * 8 + py, ox = (tt % 7) * 16 + px;
                    float4 dv[8];
                    if (oy < GS && ox < GS) {
                        const float* p0 = a.small + (((size_t)(bid / 98) * GS + oy) * GS + ox) * SRLZ_C + half * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) ldg8(p0 + j * 8, dv[2 * j], dv[2 * j + 1]);
                        if (BN_DENSE) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                dv[j] = bn_relu4(dv[j], *reinterpret_cast<const float4*>(s_bnl + half * 32 + j * 4), *reinterpret_cast<const float4*>(s_bnl + 64 + half * 32 + j * 4));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    mbar_wait(dempty(ds), dph ^ 1);
                    unsigned char* dst = smem + ds * wg::TILE;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 hi, lo;
                        split8(dv[2 * j], dv[2 * j + 1], hi, lo);
                        const int chunk = (half * 4 + j) ^ (pix & 7);
                        *reinterpret_cast<uint4*>(dst + pix * 128 + chunk * 16) = hi;
                        *reinterpret_cast<uint4*>(dst + 128 * 128 + pix * 128 + chunk * 16) = lo;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dfull(ds));
                    if (++ds == wg::ND) { ds = 0; dph ^= 1; }
                }
#pragma unroll 1
                for (int c = 0; c < NTAPS; ++c) {
                    mbar_wait(tempty(ts), tph ^ 1);
                    unsigned char* dst = smem + (wg::ND + ts) * wg::TILE;
                    float vf[32];
                    if (half == 0) patch_gather<M, 0>(vf, cur, curB, fused, a.coef, c, py, px); else patch_gather<M, 1>(vf, cur, curB, fused, a.coef, c, py, px);
                    store_half_row(vf, dst, dst + 128 * 128, pix, half);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tfull(ts));
                    if (++ts == NT) { ts = 0; tph ^= 1; }
                }
                cp_async_wait_all();
                producers_bar_sync();
            }
        } else {
        const int pidx = tid - 160;
        const int pix = MODE == 0 ? (pidx >> 1) : (pidx & 127), half = MODE == 0 ? (pidx & 1) : (pidx >> 7);
        struct Unit { const float* src; int kind; int n, sy, sx, u; };  // kind 0 = dense, 1 = tap
        int blk = -1, u = NUNITS - 1;                  // u: 0 dense, 1..NTAPS gathered tiles, then zero tiles
        long long m = 0;
        bool mvalid = false;
        int n = 0, sy = 0, sx = 0;
        auto next_unit = [&](Unit& it) -> bool {
            if (++u >= NUNITS) {
                if (++blk >= nb) return false;
                u = 0;
                m = (long long)(b0 + blk) * 128 + pix;
                mvalid = m < Ms;
                if (mvalid) {
                    sx = (int)(m % g.SW);
                    const long long q = m / g.SW;
                    sy = (int)(q % g.SH);
                    n = (int)(q / g.SH);
                }
            }
            it.src = nullptr;
            it.kind = u == 0 ? 0 : 1;
            it.n = n; it.sy = sy; it.sx = sx; it.u = u;
            if (u == 0) {
                if (mvalid) it.src = a.small + (size_t)m * SRLZ_C + half * 32;
            } else if (MODE != 0) {
                if (u <= NTAPS && mvalid) it.src = a.big;   // marker: the special gather happens in load_unit
            } else if (u <= 9 && mvalid) {
                const int tap = u - 1, ky = tap / g.KW, kx = tap % g.KW;
                const int by = sy * g.stride - g.pad + ky, bx = sx * g.stride - g.pad + kx;
                if (by >= 0 && bx >= 0 && by < g.BH && bx < g.BW) it.src = a.big + (((size_t)n * g.BH + by) * g.BW + bx) * SRLZ_C + half * 32;
            }
            return true;
        };
        auto load_unit = [&](float4 (&v)[8], const Unit& it) {
            if (it.src != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ldg8(it.src + j * 8, v[2 * j], v[2 * j + 1]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 v0[8];
        Unit i0{nullptr, 0, 0, 0, 0, 0};
        int ds = 0, dph = 0, ts = 0, tph = 0;
        // convert + store one staged unit (registers v) and signal its full barrier
        auto process = [&](float4 (&v)[8], const Unit& it) {
            uint32_t fullb;
            unsigned char* dst;
            if (it.kind == 0) {
                mbar_wait(dempty(ds), dph ^ 1);
                dst = smem + ds * wg::TILE;
                fullb = dfull(ds);
                if (++ds == wg::ND) { ds = 0; dph ^= 1; }
                if (BN_DENSE && it.src != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 sc = *reinterpret_cast<const float4*>(s_bnl + half * 32 + j * 4);
                        const float4 sh = *reinterpret_cast<const float4*>(s_bnl + 64 + half * 32 + j * 4);
                        v[j] = bn_relu4(v[j], sc, sh);
                    }
                }
            } else {
                mbar_wait(tempty(ts), tph ^ 1);
                dst = smem + (wg::ND + ts) * wg::TILE;
                fullb = tfull(ts);
                if (++ts == NT) { ts = 0; tph ^= 1; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 hi, lo;
                split8(v[2 * j], v[2 * j + 1], hi, lo);
                const int chunk = (half * 4 + j) ^ (pix & 7);
                *reinterpret_cast<uint4*>(dst + pix * 128 + chunk * 16) = hi;
                *reinterpret_cast<uint4*>(dst + 128 * 128 + pix * 128 + chunk * 16) = lo;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(fullb);
        };
        // load -> convert -> store, one unit at a time.  (A register prefetch ring was slower: in-flight LDGs share the
        // warp's six scoreboard slots with the barrier polls and shared-memory stores that follow, which then stall
        // until the loads land; asynchronous LDGSTS staging is used where shared memory allows it, see the special modes.)
        for (;;) {
            if (!next_unit(i0)) break;
            load_unit(v0, i0);
            process(v0, i0);
        }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        int ds = 0, dph = 0, ts = 0, tph = 0;
        for (int blk = 0; blk < nb; ++blk) {
            mbar_wait(dfull(ds), dph);
            tc_fence_after();
            const uint32_t dsb = d_base + ds * wg::TILE;
            for (int p = 0; p < NPAIRS; ++p) {
                const bool paired = MODE == 0 || 2 * p + 1 < NTAPS;
                mbar_wait(tfull(ts), tph);
                if (paired) mbar_wait(tfull(ts + 1), tph);
                tc_fence_after();
                if (leader) {
                    const uint32_t tsb = t_base + ts * wg::TILE, lbo = paired ? wg::TILE : 0;
                    const uint64_t ahi = make_desc_mn_sw128(tsb, lbo), alo = make_desc_mn_sw128(tsb + 128 * 128, lbo);
                    const uint64_t bhi = make_desc_mn_sw128(dsb, 0), blo = make_desc_mn_sw128(dsb + 128 * 128, 0);
                    const uint32_t d_tmem = tmem_base + p * 64;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 2048) >> 4);  // 16 pixels (K rows) = two 1024 B groups
                        const uint32_t accf = (blk > 0 || k > 0) ? 1u : 0u;
                        umma_bf16(d_tmem, alo + adv, bhi + adv, wg::IDESC, accf);
                        umma_bf16(d_tmem, ahi + adv, blo + adv, wg::IDESC, 1u);
                        umma_bf16(d_tmem, ahi + adv, bhi + adv, wg::IDESC, 1u);
                    }
                    umma_commit(tempty(ts));
                    if (paired) umma_commit(tempty(ts + 1));
                    if (p == NPAIRS - 1) umma_commit(dempty(ds));
                }
                __syncwarp();
                ts += paired ? 2 : 1;
                if (ts == NT) { ts = 0; tph ^= 1; }
            }
            if (++ds == wg::ND) { ds = 0; dph ^= 1; }
        }
        if (leader) umma_commit(acc_full);
        __syncwarp();
    } else {
        // ================================ epilogue (warps 0-3) ================================
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float* dstp = a.partials + (size_t)blockIdx.x * (NTAPS * SRLZ_C * SRLZ_C);
        const int row = tid;  // TMEM lane = accumulator row: rows 0-63 tap 2p, rows 64-127 tap 2p+1
#pragma unroll 1
        for (int p = 0; p < NPAIRS; ++p) {
            const int tap = 2 * p + (row >> 6), cg = row & 63;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + p * 64 + h * 32, v);
                if (tap < NTAPS) {
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
    if (warp == 4) tmem_dealloc(tmem_base, wg::TMEM_COLS);
}

int gwgrad64_tc_ctas(const ConvGeom& g) {
    const long long Ms = (long long)g.B * g.SH * g.SW;
    const long long nblocks = (Ms + 127) / 128;
    int gx = sm_count();
    if (gx > nblocks) gx = (int)nblocks;
    return gx;
}

// mode 1: grad W0[co][ci][s] (+)= sum_cta P[cta][ci][s][co]     mode 2: grad W12[ci*48 + j] (+)= sum_cta P[cta][0][j][ci]
__global__ void wgrad_special_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out, int ncta, int mode, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = mode == 1 ? 64 * 147 : 64 * 48;
    if (idx >= total) return;
    size_t src;
    int stride;
    if (mode == 1) {
        const int co = idx / 147, r = idx % 147, ci = r / 49, s = r % 49;
        src = ((size_t)ci * 64 + s) * 64 + co;
        stride = 3 * 4096;
    } else {
        const int ci = idx / 48, j = idx % 48;
        src = (size_t)j * 64 + ci;
        stride = 4096;
    }
    float sum = 0.f;
    for (int c = 0; c < ncta; ++c) sum += partials[(size_t)c * stride + src];
    out[idx] = accumulate ? out[idx] + sum : sum;
}

template <bool BN, int MODE>
static int launch_wg(const GWgradArgs& a, int nblocks, int gx, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gwgrad64_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gwgrad64_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    gwgrad64_tc_kernel<BN, MODE><<<gx, wg::THREADS, wg::SMEM_BYTES, st>>>(a, nblocks);
    return check_launch("gwgrad64_tc");
}

int gwgrad64_tc(const GWgradArgs& a_in, float* grad_out, int accumulate, cudaStream_t st) {
    GWgradArgs a = a_in;
    const ConvGeom& g = a.g;
    if (a.mode == 0 && (g.KH != 3 || g.KW != 3)) { set_error("gwgrad64_tc: 3x3 taps only"); return 1; }
    const long long Ms = (long long)g.B * g.SH * g.SW;
    int nblocks = (int)((Ms + 127) / 128);
    int gx = gwgrad64_tc_ctas(g);
    if (a.mode != 0) {   // 8x16-pixel blocks, 98 per image
        nblocks = g.B * 98;
        gx = sm_count() < nblocks ? sm_count() : nblocks;
    }
    int rc;
    if (a.mode == 1) rc = launch_wg<false, 1>(a, nblocks, gx, st);
    else if (a.mode == 2) rc = a.dense_scale != nullptr ? launch_wg<true, 2>(a, nblocks, gx, st) : launch_wg<false, 2>(a, nblocks, gx, st);
    else rc = a.dense_scale != nullptr ? launch_wg<true, 0>(a, nblocks, gx, st) : launch_wg<false, 0>(a, nblocks, gx, st);
    if (rc) return rc;
    if (a.mode == 0) return gwgrad64_reduce(a.partials, grad_out, gx, 9, accumulate, st);
    const int total = a.mode == 1 ? 64 * 147 : 64 * 48;
    wgrad_special_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(a.partials, grad_out, gx, a.mode, accumulate);
    return check_launch("wgrad_special_reduce");
}

// d(decoded) summed per output channel -> decoder_conv.12.bias gradient
__global__ void __launch_bounds__(256) dec12_bias_partials_kernel(const float* __restrict__ gout, const float* __restrict__ dec,
                                                                  const float* __restrict__ tgt, float coef, long long n4,
                                                                  float* __restrict__ partials) {
    __shared__ float s_w[8][3];
    float acc[3] = {0.f, 0.f, 0.f};
    const long long plane4 = 224 * 224 / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)((i / plane4) % 3);
        float4 gg;
        if (gout != nullptr) {
            gg = ldg4(gout + i * 4);
        } else {
            const float4 d = ldg4(dec + i * 4), t = ldg4(tgt + i * 4);
            gg = make_float4(coef * (d.x - t.x), coef * (d.y - t.y), coef * (d.z - t.z), coef * (d.w - t.w));
        }
        const float sgg = (gg.x + gg.y) + (gg.z + gg.w);
        if (co == 0) acc[0] += sgg; else if (co == 1) acc[1] += sgg; else acc[2] += sgg;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = warp_sum(acc[c]);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float v = 0.f;
        for (int w2 = 0; w2 < 8; ++w2) v += s_w[w2][threadIdx.x];
        partials[(size_t)blockIdx.x * 3 + threadIdx.x] = v;
    }
}
__global__ void dec12_bias_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ grad_b, int accumulate) {
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int i = 0; i < n; ++i) v += (double)partials[(size_t)i * 3 + threadIdx.x];
        grad_b[threadIdx.x] = accumulate ? grad_b[threadIdx.x] + (float)v : (float)v;
    }
}
int dec12_bias_grad(const float* gout, const float* decoded, const float* target, float coef, int B, float* partials, float* grad_b,
                    int accumulate, cudaStream_t st) {
    const long long n4 = (long long)B * 3 * 224 * 224 / 4;
    int gx = sm_count() * 4;
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    dec12_bias_partials_kernel<<<gx, 256, 0, st>>>(gout, decoded, target, coef, n4, partials);
    int rc = check_launch("dec12_bias_partials");
    if (rc) return rc;
    dec12_bias_final_kernel<<<1, 32, 0, st>>>(partials, gx, grad_b, accumulate);
    return check_launch("dec12_bias_final");
}

}  // namespace srlz

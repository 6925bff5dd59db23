// Row-image tcgen05 kernels for the first encoder layer, Conv2d(3, 64, 7, s2, p3, bias=False) on the NCHW observation
// (models/models.py:49): forward (+ BatchNorm sums) and weight gradient.
//
// The im2col matrix of this layer has K = 3*7*7 = 147 columns per output pixel, and a per-tile software im2col converts
// every input pixel ~12 times.  Here the K axis is factored as (pair of input rows) x (channel, row-in-pair, kx):
//
//   pair image P_j[ox][k],  k = c*16 + rr*8 + kx  (c < 3, rr < 2, kx < 8; k >= 48 unused)
//             = x[n, c, 2j-3+rr, 2ox-4+kx]         (zero outside the image / inside the DAE rectangle)
//
// is one 128-byte SWIZZLE_128B row per output column ox, and output row oy is
//
//   y[oy, ox, :] = sum_{p=0..3} P_{oy+p}[ox, :] . Wp[p]       Wp[p][k][co] = W[co, c, 2p+rr, kx-1]  (0 for ky = 7 or kx = 0)
//
// i.e. four K=48 MMAs (M = 128 rows = the 112 output columns of one row, N = 64) whose A operands are four consecutive
// pair images.  Consecutive output rows share three of their four pair images, so a CTA walks a contiguous range of
// output rows with the pair images in a ring of shared-memory slots: every input pixel is loaded and converted to bf16
// hi/lo ONCE per CTA (8 contiguous floats -> one 16-byte chunk of hi and of lo), all descriptors are 1024-byte aligned.
// fp32 fidelity comes from the same bf16x3 split as the other tensor-core kernels.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "bn_tail.cuh"

namespace srlz {

namespace er {
constexpr int OW = 112, OH = 112, IW = 224;
constexpr int PPI = OH + 3;                              // pair images per input image (j = 0..114)
constexpr int NSLOT = 5;                                 // ring: 4 in use by the current output row + 1 being refilled
constexpr int SLOT_BYTES = OW * 128;                     // 14336 = 14 x 1024: every slot starts on a swizzle-atom boundary
constexpr int PLANE = NSLOT * SLOT_BYTES + 16 * 128;     // + 16 rows: the M=128 MMA on the last slot reads 16 rows past it
constexpr int W_BYTES = 4 * 16384;                       // 4 pairs x (hi 8 KB | lo 8 KB)
constexpr int THREADS = 16 * 32;                         // warps 0-7 epilogue (two groups, one per accumulator) | 8 MMA | 9-15 producers
constexpr int OFF_W = 2 * PLANE;
constexpr int OFF_BARS = OFF_W + W_BYTES;
constexpr int OFF_STG = OFF_BARS + 1024;
constexpr int SMEM_BYTES = OFF_STG + 8 * 2048 + 1024;    // 231424 <= 232448
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
__host__ __device__ inline int g0_of(int item) { return (item / OH) * PPI + item % OH; }   // first pair image of output row `item`
}  // namespace er

// the 24 values (3 channels x 8 columns) of pair image g that thread (ox, rr) owns.  The 8-column window starts at the even
// column 2ox-4 (slot kx' = kx + 1; slot 0 carries a zero weight), so it is four aligned float2 loads per channel and a
// float2 never straddles the image border.
__device__ __forceinline__ void er_load(float (&v)[3][8], const float* __restrict__ x, const int* __restrict__ rects, int g, int ox, int rr) {
    const int n = g / er::PPI, j = g - n * er::PPI;
    const int r = 2 * j - 3 + rr, c0 = 2 * ox - 4;
    const bool rok = r >= 0 && r < er::IW;
    const float* src = x + ((size_t)n * 3 * er::IW + (rok ? r : 0)) * er::IW + c0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int col = c0 + 2 * h;
        const bool ok = rok && col >= 0 && col < er::IW;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float2 t = make_float2(0.f, 0.f);
            if (ok) t = __ldg(reinterpret_cast<const float2*>(src + (size_t)c * er::IW * er::IW + 2 * h));
            v[c][2 * h] = t.x;
            v[c][2 * h + 1] = t.y;
        }
    }
    if (rects != nullptr) {   // rows [w1,w2) x cols [h1,h2) are zeroed (preprocessing/data_loader.py:55-63)
        if (r >= rects[n * 4 + 2] && r < rects[n * 4 + 3]) {
            const int h1 = rects[n * 4], h2 = rects[n * 4 + 1];
#pragma unroll
            for (int kx = 0; kx < 8; ++kx) {
                if (c0 + kx >= h1 && c0 + kx < h2) { v[0][kx] = 0.f; v[1][kx] = 0.f; v[2][kx] = 0.f; }
            }
        }
    }
}

template <int PLANE_BYTES>
__device__ __forceinline__ void er_store(const float (&v)[3][8], unsigned char* smem, int slot, int ox, int rr) {
    const int row = slot * er::OW + ox;
    unsigned char* dst = smem + row * 128;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        uint4 hi, lo;
        split8(make_float4(v[c][0], v[c][1], v[c][2], v[c][3]), make_float4(v[c][4], v[c][5], v[c][6], v[c][7]), hi, lo);
        const int chunk = (c * 2 + rr) ^ (row & 7);
        *reinterpret_cast<uint4*>(dst + chunk * 16) = hi;
        *reinterpret_cast<uint4*>(dst + PLANE_BYTES + chunk * 16) = lo;
    }
}

#define ER_STAMP(idx, slot) do { if (dbg != nullptr && blockIdx.x == 0 && (idx) >= 0 && (idx) < 64) dbg[(idx) * 16 + (slot)] = clock64(); } while (0)

// Pair-image producer loop shared by the forward and the wgrad kernel (7 warps, pidx = 0..223): pair images g_lo..g_hi go
// into ring slot (g - g_lo) % NSLOT; full / empty barriers are NSLOT consecutive mbarriers each.  Two register sets
// alternate so the loads of pair g+1 are in flight while pair g is converted (no copy = no wait in between), and the 42
// cache lines of the pair PF steps ahead are pulled into L2 (the observation is not L2 resident: first kernel of a step).
template <int NSLOT, int PLANE_BYTES>
__device__ __forceinline__ void er_produce(const float* __restrict__ x, const int* __restrict__ rects, unsigned char* smem, uint32_t full0,
                                           uint32_t empty0, int g_lo, int g_hi, int pidx, int lane, long long* dbg) {
    constexpr int PF = 6;
    const int ox = pidx % er::OW, rr = pidx / er::OW;
    auto prefetch = [&](int g) {
        if (g <= g_hi && pidx < 42) {
            const int n = g / er::PPI, j = g - n * er::PPI;
            const int r = 2 * j - 3 + pidx / 21, c = (pidx % 21) / 7, line = pidx % 7;
            if (r >= 0 && r < er::IW) prefetch_l2(x + (((size_t)n * 3 + c) * er::IW + r) * er::IW + line * 32);
        }
    };
    auto step = [&](int g, const float (&v)[3][8]) {
        const int rel = g - g_lo, slot = rel % NSLOT, ph = (rel / NSLOT) & 1;
        if (pidx == 0) ER_STAMP(rel, 0);
        mbar_wait(empty0 + 8u * slot, ph ^ 1);
        if (pidx == 0) ER_STAMP(rel, 1);
        er_store<PLANE_BYTES>(v, smem, slot, ox, rr);
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8u * slot);
        if (pidx == 0) ER_STAMP(rel, 2);
    };
    if (g_lo > g_hi) return;
    float va[3][8], vb[3][8];
    for (int d = 1; d < PF; ++d) prefetch(g_lo + d);
    er_load(va, x, rects, g_lo, ox, rr);
    for (int g = g_lo; g <= g_hi; g += 2) {
        prefetch(g + PF);
        if (g + 1 <= g_hi) er_load(vb, x, rects, g + 1, ox, rr);
        step(g, va);
        if (g + 1 <= g_hi) {
            prefetch(g + 1 + PF);
            if (g + 2 <= g_hi) er_load(va, x, rects, g + 2, ox, rr);
            step(g + 1, vb);
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(er::THREADS, 1) enc0_rows_fwd_kernel(const float* __restrict__ x, const int* __restrict__ rects,
                                                                       const unsigned char* __restrict__ wbf, const float* __restrict__ bias,
                                                                       float* __restrict__ out, float* __restrict__ partials, int total_items,
                                                                       long long* __restrict__ dbg, BnTail tail) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t img_hi = base, img_lo = base + er::PLANE, wsm = base + er::OFF_W, bars = base + er::OFF_BARS;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + er::OFF_BARS + 192);   // 19 mbarriers occupy bytes 0..151
    float* s_bias = reinterpret_cast<float*>(smem + er::OFF_BARS + 256);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (er::NSLOT + s); };
    // four 64-column accumulators: output row `it` uses accumulator it & 3; epilogue group eg = it & 1 drains accumulators eg, eg + 2
    auto tfull_bar = [&](int i) { return bars + 8u * (2 * er::NSLOT + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (2 * er::NSLOT + 4 + i); };
    const uint32_t wfull = bars + 8u * (2 * er::NSLOT + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = (int)((long long)total_items * blockIdx.x / gridDim.x), i1 = (int)((long long)total_items * (blockIdx.x + 1) / gridDim.x);
    const int g_lo = er::g0_of(i0), g_hi = i1 > i0 ? er::g0_of(i1 - 1) + 3 : g_lo - 1;

    if (tid == 0) {
        for (int s = 0; s < er::NSLOT; ++s) { mbar_init(full_bar(s), 7); mbar_init(empty_bar(s), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (tid < 64) s_bias[tid] = bias != nullptr ? bias[tid] : 0.f;
    for (int e = tid; e < 2 * er::PLANE / 16; e += er::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 8) tmem_alloc(smem_u32(tmem_ptr_smem), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (tid == 0) {
        mbar_arrive_expect_tx(wfull, er::W_BYTES);
        for (int t = 0; t < 4; ++t) bulk_g2s(wsm + t * 16384, wbf + (size_t)t * 16384, 16384, wfull);
    }

    if (warp >= 9) {
        // ================================ producers: one pair image per step ================================
        er_produce<er::NSLOT, er::PLANE>(x, rects, smem, full_bar(0), empty_bar(0), g_lo, g_hi, tid - 9 * 32, lane, dbg);
    } else if (warp == 8) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        mbar_wait(wfull, 0);
        int ready = 0, it = 0;
        for (int i = i0; i < i1; ++i, ++it) {
            const int buf = it & 3;
            if (lane == 0) ER_STAMP(it, 3);
            mbar_wait(tempty_bar(buf), ((it >> 2) & 1) ^ 1);
            if (lane == 0) ER_STAMP(it, 4);
            const int g0r = er::g0_of(i) - g_lo;
            for (; ready <= g0r + 3; ++ready) mbar_wait(full_bar(ready % er::NSLOT), (ready / er::NSLOT) & 1);
            // consumer-side proxy fence: the producers' st.shared are ordered before this point by the mbarrier (release /
            // acquire); fencing here instead of in the producers keeps MEMBAR.ALL (which a fence.proxy.async lowers to) away from
            // warps that have global loads in flight -- there it drains the prefetched loads and exposes their full latency
            fence_proxy_async_smem();
            tc_fence_after();
            if (lane == 0) ER_STAMP(it, 5);
            const int nxt = i + 1 < i1 ? er::g0_of(i + 1) - g_lo : g0r + 4;   // pair images below `nxt` are not needed again
            if (leader) {
                const uint32_t d_tmem = tmem_base + buf * 64;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int slot = (g0r + p) % er::NSLOT;
                    const uint64_t ahi = make_desc_sw128(img_hi + slot * er::SLOT_BYTES), alo = make_desc_sw128(img_lo + slot * er::SLOT_BYTES);
                    const uint64_t whi = make_desc_sw128(wsm + p * 16384), wlo = make_desc_sw128(wsm + p * 16384 + 8192);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {   // K = 48: the last 16 slots of a row are padding
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        umma_bf16(d_tmem, alo + adv, whi + adv, er::IDESC, (p | k) ? 1u : 0u);
                        umma_bf16(d_tmem, ahi + adv, wlo + adv, er::IDESC, 1u);
                        umma_bf16(d_tmem, ahi + adv, whi + adv, er::IDESC, 1u);
                    }
                    if (g0r + p < nxt) umma_commit(empty_bar(slot));
                }
                umma_commit(tfull_bar(buf));
            }
            __syncwarp();
            if (lane == 0) ER_STAMP(it, 6);
        }
    } else {
        // ================================ epilogue: group eg owns accumulator eg (every other output row) ================================
        // same register -> shared staging -> coalesced store scheme as conv_halo_tc.cu
        const int eg = warp >> 2, q = warp & 3;
        unsigned char* stg = smem + er::OFF_STG + warp * 2048;
        const int cq = lane & 3, rsub = lane >> 2;
        float st1[16], st2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { st1[i] = 0.f; st2[i] = 0.f; }
        const int ox = q * 32 + lane;
        for (int i = i0 + eg, it = eg; i < i1; i += 2, it += 2) {
            if ((tid & 127) == 0) ER_STAMP(it, 7);
            const int buf = it & 3;
            mbar_wait(tfull_bar(buf), (it >> 2) & 1);
            tc_fence_after();
            if ((tid & 127) == 0) ER_STAMP(it, 8);
            const int mypix = ox < er::OW ? i * er::OW + ox : -1;
            int rowpix[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rowpix[k] = __shfl_sync(0xffffffffu, mypix, 8 * k + rsub);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int ch0 = qq * 16 + cq * 4;
                float v[16];
                tmem_ld16(taddr + qq * 16, v);
                if (qq == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(buf));
                    if ((tid & 127) == 0) ER_STAMP(it, 9);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                const float4 b4 = *reinterpret_cast<const float4*>(s_bias + ch0);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int row = 8 * k + rsub;
                    const float4 d4 = *reinterpret_cast<const float4*>(stg + row * 64 + ((cq ^ ((row >> 1) & 3)) << 4));
                    const bool valid = rowpix[k] >= 0;
                    float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float y = valid ? d[e] + bb[e] : 0.f;
                        d[e] = y;
                        if (EPI == EPI_STATS) {
                            st1[qq * 4 + e] += y;
                            st2[qq * 4 + e] = fmaf(y, y, st2[qq * 4 + e]);
                        }
                    }
                    if (valid) st4(out + (size_t)rowpix[k] * SRLZ_C + ch0, make_float4(d[0], d[1], d[2], d[3]));
                }
                __syncwarp();
            }
            if ((tid & 127) == 0) ER_STAMP(it, 10);
        }
        if (EPI == EPI_STATS) {
            float* red = reinterpret_cast<float*>(stg);   // this warp's own staging rows: [sum 64 | sum-sq 64]
#pragma unroll
            for (int i = 0; i < 16; ++i) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    st1[i] += __shfl_xor_sync(0xffffffffu, st1[i], o);
                    st2[i] += __shfl_xor_sync(0xffffffffu, st2[i], o);
                }
            }
            if (lane < 4) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        red[qq * 16 + cq * 4 + e] = st1[qq * 4 + e];
                        red[64 + qq * 16 + cq * 4 + e] = st2[qq * 4 + e];
                    }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (EPI == EPI_STATS && tid < 128) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += reinterpret_cast<const float*>(smem + er::OFF_STG + w * 2048)[tid];
        partials[(size_t)blockIdx.x * 128 + tid] = v;
    }
    if (warp == 8) tmem_dealloc(tmem_base, 256);
    if (EPI == EPI_STATS && tail.counter != nullptr) bn_tail_run(tail, partials, reinterpret_cast<double*>(smem), tid);   // (the pair-image ring is free)
}

// a.in = observation (B,3,224,224) NCHW, a.rects = DAE rectangles or null, a.out = (B,112,112,64) NHWC pre-BN,
// a.partials = [n][128] BatchNorm sums (EPI_STATS); wbf = image written by pack_enc0_rows_bf16
int enc0_rows_fwd(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st) {
    const int total = a.g.B * er::OH;
    int gx = sm_count();
    if (gx > total) gx = total;
    if (n_partials) *n_partials = gx;
    if (a.epi != EPI_PLAIN && a.epi != EPI_STATS) { set_error("enc0_rows_fwd: unsupported epilogue"); return 1; }
    if (a.epi == EPI_STATS && a.partials == nullptr) { set_error("enc0_rows_fwd: partials buffer required"); return 1; }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(enc0_rows_fwd_kernel<EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, er::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(enc0_rows_fwd_kernel<EPI_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, er::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("enc0_rows_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    const unsigned char* w = reinterpret_cast<const unsigned char*>(wbf);
    if (a.epi == EPI_STATS)
        launch_k(enc0_rows_fwd_kernel<EPI_STATS>, gx, er::THREADS, er::SMEM_BYTES, st, a.in, a.rects, w, a.bias, a.out, a.partials, total, a.dbg, a.tail);
    else
        launch_k(enc0_rows_fwd_kernel<EPI_PLAIN>, gx, er::THREADS, er::SMEM_BYTES, st, a.in, a.rects, w, a.bias, a.out, a.partials, total, a.dbg, BnTail{});
    return check_launch("enc0_rows_fwd");
}

// W0[co][c][ky][kx] (torch layout 64,3,7,7) -> four pair images [p]{hi[64][64], lo[64][64]} (K-major SWIZZLE_128B rows of
// 128 B): row co, K slot k = c*16 + rr*8 + kx + 1 holds W0[co][c][2p+rr][kx]; ky = 7, slot kx = 0 and k >= 48 are zero
__global__ void pack_enc0_rows_bf16_kernel(const float* __restrict__ w0, unsigned char* __restrict__ dst) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // p*4096 + co*64 + k
    if (idx >= 4 * 4096) return;
    const int p = idx >> 12, co = (idx >> 6) & 63, k = idx & 63;
    float x = 0.f;
    if (k < 48) {
        const int c = k >> 4, rr = (k >> 3) & 1, kx = (k & 7) - 1, ky = 2 * p + rr;
        if (ky < 7 && kx >= 0) x = w0[((co * 3 + c) * 7 + ky) * 7 + kx];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const int byte = co * 128 + (((k >> 3) ^ (co & 7)) << 4) + (k & 7) * 2;
    unsigned char* t = dst + (size_t)p * 16384;
    *reinterpret_cast<__nv_bfloat16*>(t + byte) = hi;
    *reinterpret_cast<__nv_bfloat16*>(t + 8192 + byte) = lo;
}
int pack_enc0_rows_bf16(const float* w0, void* dst, cudaStream_t st) {
    launch_k(pack_enc0_rows_bf16_kernel, 64, 256, 0, st, w0, reinterpret_cast<unsigned char*>(dst));
    return check_launch("pack_enc0_rows_bf16");
}

// ---------------------------------------------------------------------------------------------------------------------------
// weight gradient:  dW[co, c, 2p+rr, kx] = sum_{n,oy,ox} P_{oy+p}[ox][k] * dy[n,oy,ox,co]      (k = c*16 + rr*8 + kx + 1)
// Per output row: A = pair images (MN-major: the 64 k slots of a row are M, the 112 output columns are K), B = the dy row
// (MN-major, N = 64 channels).  Two pair images in neighbouring ring slots are stacked on M = 128 through the descriptor's
// LBO; the ring has an even number of slots and stacks start on even ring positions so a stack never wraps.  An output row
// whose first pair image sits on an even position issues stacks (p0,p1) (p2,p3); on an odd position (-,p0) (p1,p2) (p3,-),
// the '-' halves land in accumulator rows nobody reads.  Five accumulators (320 TMEM columns) live for the CTA's whole
// row range and are written out once; wgrad_rows_reduce folds CTAs and stack halves into the torch layout.
// QUAD form (product path): the hi and lo planes of ONE pair image are stacked on M = 128 (LBO = plane distance, so no ring
// position ever wraps) and the hi and lo planes of the dy row on N = 128: a single M=128 x N=128 MMA per K step yields all
// four hi/lo products of a pair image (28 MMAs of 64 cycles per output row instead of 52.5 of ~57), one 128-column
// accumulator per pair (4 x 128 = all 512 TMEM columns); the epilogue adds the column halves, the reduce the row halves.
namespace ew {
constexpr int NSLOT = 6;
constexpr int RING_PLANE = NSLOT * er::SLOT_BYTES;        // 86016 = 84 x 1024
constexpr int DY_STAGE = 2 * er::SLOT_BYTES;              // hi | lo, 112 rows x 128 B each
constexpr int OFF_DY = 2 * RING_PLANE;                    // 172032
constexpr int OFF_BARS = OFF_DY + 2 * DY_STAGE;           // 229376
constexpr int SMEM_BYTES = OFF_BARS + 1024 + 1024;        // 231424
constexpr int NACC = 5;
constexpr int PART_FLOATS = NACC * 128 * 64;              // per CTA
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_Q = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int PART_FLOATS_Q = 4 * 128 * 64;               // QUAD form, per CTA
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
}  // namespace ew

template <bool QUAD>
__global__ void __launch_bounds__(er::THREADS, 1) enc0_rows_wgrad_kernel(const float* __restrict__ x, const int* __restrict__ rects,
                                                                         const float* __restrict__ dy, float* __restrict__ partials,
                                                                         int total_items, long long* __restrict__ dbg) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t ring_hi = base, ring_lo = base + ew::RING_PLANE, dyb = base + ew::OFF_DY, bars = base + ew::OFF_BARS;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + ew::OFF_BARS + 192);
    uint32_t* used_smem = reinterpret_cast<uint32_t*>(smem + ew::OFF_BARS + 196);
    auto pfull = [&](int s) { return bars + 8u * s; };
    auto pempty = [&](int s) { return bars + 8u * (ew::NSLOT + s); };
    auto dfull = [&](int i) { return bars + 8u * (2 * ew::NSLOT + i); };
    auto dempty = [&](int i) { return bars + 8u * (2 * ew::NSLOT + 2 + i); };
    const uint32_t done = bars + 8u * (2 * ew::NSLOT + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = (int)((long long)total_items * blockIdx.x / gridDim.x), i1 = (int)((long long)total_items * (blockIdx.x + 1) / gridDim.x);
    const int g_lo = er::g0_of(i0), g_hi = i1 > i0 ? er::g0_of(i1 - 1) + 3 : g_lo - 1;

    if (tid == 0) {
        for (int s = 0; s < ew::NSLOT; ++s) { mbar_init(pfull(s), 7); mbar_init(pempty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(dfull(i), 7); mbar_init(dempty(i), 1); }
        mbar_init(done, 1);
        fence_barrier_init();
        *used_smem = 0u;
    }
    for (int e = tid; e < 2 * ew::RING_PLANE / 16; e += er::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 8) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 9) {
        // ================================ pair-image producers (as in the forward kernel) ================================
        er_produce<ew::NSLOT, ew::RING_PLANE>(x, rects, smem, pfull(0), pempty(0), g_lo, g_hi, tid - 9 * 32, lane, dbg);
    } else if (warp == 8) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        int ready = 0, it = 0;
        uint32_t used = 0u;
        for (int i = i0; i < i1; ++i, ++it) {
            const int st = it & 1;
            const int g0r = er::g0_of(i) - g_lo;
            if (lane == 0) ER_STAMP(it, 3);
            for (; ready <= g0r + 3; ++ready) mbar_wait(pfull(ready % ew::NSLOT), (ready / ew::NSLOT) & 1);
            if (lane == 0) ER_STAMP(it, 4);
            mbar_wait(dfull(st), (it >> 1) & 1);
            fence_proxy_async_smem();
            tc_fence_after();
            if (lane == 0) ER_STAMP(it, 5);
            const int nxt = i + 1 < i1 ? er::g0_of(i + 1) - g_lo : g0r + 4;
            const bool odd = (g0r & 1) != 0;
            const int nstack = odd ? 3 : 2, first = odd ? g0r - 1 : g0r, acc0 = odd ? 2 : 0;
            if (QUAD) {
                if (leader) {
                    const uint32_t dsb = dyb + st * ew::DY_STAGE;
                    const uint64_t bhl = ew::desc_mn(dsb, er::SLOT_BYTES);   // N = 128: dy_hi | dy_lo
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int slot = (g0r + p) % ew::NSLOT;
                        const uint64_t ahl = ew::desc_mn(ring_hi + slot * er::SLOT_BYTES, ew::RING_PLANE);   // M = 128: P_hi ; P_lo
                        const uint32_t d_tmem = tmem_base + p * 128;
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
                            const uint64_t adv = (uint64_t)((k * 2048) >> 4);
                            umma_bf16(d_tmem, ahl + adv, bhl + adv, ew::IDESC_Q, (it | k) ? 1u : 0u);
                        }
                        if (g0r + p < nxt) umma_commit(pempty(slot));
                    }
                    umma_commit(dempty(st));
                }
                used = 0xFu;
                __syncwarp();
                if (lane == 0) ER_STAMP(it, 6);
                continue;
            }
            if (leader) {
                const uint32_t dsb = dyb + st * ew::DY_STAGE;
                const uint64_t bhi = ew::desc_mn(dsb, 0), blo = ew::desc_mn(dsb + er::SLOT_BYTES, 0);
                for (int s = 0; s < nstack; ++s) {
                    const int slot = (first + 2 * s) % ew::NSLOT;   // even: the stacked neighbour slot + 1 exists
                    const uint64_t ahi = ew::desc_mn(ring_hi + slot * er::SLOT_BYTES, er::SLOT_BYTES);
                    const uint64_t alo = ew::desc_mn(ring_lo + slot * er::SLOT_BYTES, er::SLOT_BYTES);
                    const uint32_t d_tmem = tmem_base + (acc0 + s) * 64;
                    const bool fresh = ((used >> (acc0 + s)) & 1u) == 0u;
#pragma unroll
                    for (int k = 0; k < 7; ++k) {   // K = 112 output columns = 7 x 16
                        const uint64_t adv = (uint64_t)((k * 2048) >> 4);
                        umma_bf16(d_tmem, alo + adv, bhi + adv, ew::IDESC, (fresh && k == 0) ? 0u : 1u);
                        umma_bf16(d_tmem, ahi + adv, blo + adv, ew::IDESC, 1u);
                        umma_bf16(d_tmem, ahi + adv, bhi + adv, ew::IDESC, 1u);
                    }
                }
                for (int g = g0r; g < nxt && g < g0r + 4; ++g) umma_commit(pempty(g % ew::NSLOT));
                umma_commit(dempty(st));
            }
            for (int s = 0; s < nstack; ++s) used |= 1u << (acc0 + s);
            __syncwarp();
            if (lane == 0) ER_STAMP(it, 6);
        }
        if (leader) {
            *used_smem = used;
            if (i1 > i0) umma_commit(done); else mbar_arrive(done);
        }
        __syncwarp();
    } else if (warp < 7) {
        // ================================ dy producers: one output row (112 x 64 fp32) per step ================================
        const int ox = tid >> 1, half = tid & 1;
        constexpr int PFD = 4;
        float4 va[8], vb[8];
        auto load = [&](int i, float4 (&d)[8]) {
            const float* src = dy + ((size_t)i * er::OW + ox) * SRLZ_C + half * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg8(src + j * 8, d[2 * j], d[2 * j + 1]);
        };
        auto prefetch = [&](int i) { if (i < i1) prefetch_l2(dy + (size_t)i * er::OW * SRLZ_C + tid * 32); };   // 224 lines = one row
        auto step = [&](int i, const float4 (&v)[8]) {
            const int it = i - i0, st = it & 1;
            if (tid == 0) ER_STAMP(it, 7);
            mbar_wait(dempty(st), ((it >> 1) & 1) ^ 1);
            if (tid == 0) ER_STAMP(it, 8);
            unsigned char* dst = smem + ew::OFF_DY + st * ew::DY_STAGE + ox * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 hi, lo;
                split8(v[2 * j], v[2 * j + 1], hi, lo);
                const int chunk = (half * 4 + j) ^ (ox & 7);
                *reinterpret_cast<uint4*>(dst + chunk * 16) = hi;
                *reinterpret_cast<uint4*>(dst + er::SLOT_BYTES + chunk * 16) = lo;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(dfull(st));
            if (tid == 0) ER_STAMP(it, 9);
        };
        if (i0 < i1) {
            for (int d = 1; d < PFD; ++d) prefetch(i0 + d);
            load(i0, va);
            for (int i = i0; i < i1; i += 2) {
                prefetch(i + PFD);
                if (i + 1 < i1) load(i + 1, vb);
                step(i, va);
                if (i + 1 < i1) {
                    prefetch(i + 1 + PFD);
                    if (i + 2 < i1) load(i + 2, va);
                    step(i + 1, vb);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {
        // accumulators -> partials [cta][acc][row 128][co 64]; accumulators this CTA never touched are written as zeros
        mbar_wait(done, 0);
        tc_fence_after();
        const uint32_t used = *used_smem;
        if (QUAD) {   // partials [cta][pair 4][row 128][co 64], column halves (x dy_hi, x dy_lo) added
            float* dq = partials + (size_t)blockIdx.x * ew::PART_FLOATS_Q + (size_t)(warp * 32 + lane) * 64;
            for (int p = 0; p < 4; ++p) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    float v[16], w[16];
                    if (used != 0u) {
                        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + p * 128 + qq * 16, v);
                        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + p * 128 + 64 + qq * 16, w);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) { v[e] = 0.f; w[e] = 0.f; }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st4(dq + (size_t)p * 128 * 64 + qq * 16 + j * 4,
                            make_float4(v[4 * j] + w[4 * j], v[4 * j + 1] + w[4 * j + 1], v[4 * j + 2] + w[4 * j + 2], v[4 * j + 3] + w[4 * j + 3]));
                }
            }
        }
        float* dst = partials + (size_t)blockIdx.x * ew::PART_FLOATS + (size_t)(warp * 32 + lane) * 64;
        for (int acc = 0; acc < (QUAD ? 0 : ew::NACC); ++acc) {
            const bool have = (used >> acc) & 1u;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                float v[16];
                if (have) {
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64 + qq * 16, v);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) st4(dst + (size_t)acc * 128 * 64 + qq * 16 + j * 4, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// grad[co][c][ky][kx] (+)= sum over CTAs of the two stack halves that hold pair p = ky>>1, slot k = c*16 + (ky&1)*8 + kx + 1:
//   p0: acc0 rows k      + acc2 rows 64+k      p1: acc0 rows 64+k + acc3 rows k
//   p2: acc1 rows k      + acc3 rows 64+k      p3: acc1 rows 64+k + acc4 rows k
// QUAD form: p: acc p rows k (P_hi x dy) + rows 64+k (P_lo x dy)
__global__ void __launch_bounds__(512) enc0_rows_wgrad_reduce_kernel(const float* __restrict__ partials, int nctas, float* __restrict__ grad,
                                                                    int accumulate, int quad) {
    pdl_enter();
    // one block per (c, ky, kx): 64 output channels x 8 groups of CTAs, folded in a fixed order
    __shared__ double s_part[8][64];
    const int t = blockIdx.x, co = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int kx = t % 7, ky = (t / 7) % 7, c = t / 49;
    const int p = ky >> 1, k = c * 16 + (ky & 1) * 8 + kx + 1;
    const int accA = p < 2 ? 0 : 1, rowA = (p & 1) ? 64 + k : k;
    const int accB = p == 0 ? 2 : (p == 3 ? 4 : 3), rowB = (p == 0 || p == 2) ? 64 + k : k;
    size_t offA = ((size_t)accA * 128 + rowA) * 64 + co, offB = ((size_t)accB * 128 + rowB) * 64 + co;
    if (quad) { offA = ((size_t)p * 128 + k) * 64 + co; offB = ((size_t)p * 128 + 64 + k) * 64 + co; }
    const size_t cta_stride = quad ? ew::PART_FLOATS_Q : ew::PART_FLOATS;
    double s = 0.0;
    for (int b = grp; b < nctas; b += 8) {
        const float* pb = partials + (size_t)b * cta_stride;
        s += (double)pb[offA] + (double)pb[offB];
    }
    s_part[grp][co] = s;
    __syncthreads();
    if (threadIdx.x < 64) {
        double tot = 0.0;
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) tot += s_part[g8][co];
        float* g = grad + ((co * 3 + c) * 7 + ky) * 7 + kx;
        *g = accumulate ? *g + (float)tot : (float)tot;
    }
}

size_t enc0_rows_wgrad_partial_floats() { return (size_t)sm_count() * ew::PART_FLOATS; }

// a.big = observation (B,3,224,224), a.rects = DAE rectangles or null, a.small = dy (B,112,112,64), a.partials = workspace
int enc0_rows_wgrad(const GWgradArgs& a, float* grad_out, int accumulate, cudaStream_t st) {
    const int total = a.g.B * er::OH;
    int gx = sm_count();
    if (gx > total) gx = total;
    static bool configured = false;
    const int quad = 1;   // the QUAD form (hi/lo planes stacked on M and N) is the only one built
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(enc0_rows_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ew::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("enc0_rows_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    launch_k(enc0_rows_wgrad_kernel<true>, gx, er::THREADS, ew::SMEM_BYTES, st, a.big, a.rects, a.small, a.partials, total, a.dbg);
    int rc = check_launch("enc0_rows_wgrad");
    if (rc) return rc;
    launch_k(enc0_rows_wgrad_reduce_kernel, 147, 512, 0, st, a.partials, gx, grad_out, accumulate, quad);
    return check_launch("enc0_rows_wgrad_reduce");
}

}  // namespace srlz

// Row-ring tcgen05 kernel for the forward of the last decoder layer, ConvTranspose2d(64, 3, 4, s2) + bias -> NCHW, with the
// squared error against the target fused into the epilogue (models/models.py:82, losses/losses.py:172-214).
//
//   out[2y+py, 2x+px, co] = b[co] + sum_{dy,dx in {0,1}} sum_ci a[y-dy, x-dx, ci] * W[ci, co, py+2dy, px+2dx]      a = relu(bn(y7))
//
// Same GEMM as the N = 16 variant of the halo kernel (conv_halo_tc.cu): accumulator row = x, column j = (py*2+px)*3 + co,
// the four (dy,dx) shifts are four row-shifted K-major descriptors.  The difference is the staging: there every tile
// staged its two input rows (y-1, y), i.e. every input row was loaded, BN+ReLU'd and split to bf16 hi/lo twice, and the
// producers bound the kernel.  Here a CTA walks a contiguous range of output row pairs y and keeps the input rows in a ring
// of shared-memory slots (one row image = 113 pixel rows of 128 B: pixel -1 .. 111, the border pixels stay zero), so every
// input row is staged once; two register sets keep the next row's loads in flight, and rows are pulled into L2 ahead.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace srlz {

namespace dr {
constexpr int IN = 111, OUT = 224, NY = 112;               // input edge, output edge, output row pairs per image
constexpr int RPI = NY + 1;                                // row images per input image: rows -1 .. 111 (first and last are zero)
constexpr int NSLOT = 4;                                   // 2 in use by the current y + 2 being refilled
constexpr int SLOT_BYTES = 128 * 128;                      // 128 pixel rows (113 used) x 128 B
constexpr int PLANE = NSLOT * SLOT_BYTES + 1024;           // the dx = 0 descriptor of the last slot reads one row past it
constexpr int W_BYTES = 4 * 4096;                          // 4 shifts x (hi 2 KB | lo 2 KB), 16 rows x 128 B each
constexpr int THREADS = 16 * 32;                           // warps 0-3 epilogue | 4 MMA (5-7 idle) | 8-15 producers
constexpr int OFF_W = 2 * PLANE;
constexpr int OFF_BARS = OFF_W + W_BYTES;
constexpr int SMEM_BYTES = OFF_BARS + 1024 + 1024;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);     // N = 16
constexpr uint32_t IDESC32 = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // N = 32
__host__ __device__ inline int g0_of(int item) { return (item / NY) * RPI + item % NY; }   // row image of input row y-1
}  // namespace dr

#define DR_STAMP(idx, slot) do { if (dbg != nullptr && blockIdx.x == 0 && (idx) >= 0 && (idx) < 64) dbg[(idx) * 16 + (slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(dr::THREADS, 1) dec12_rows_fwd_kernel(const float* __restrict__ in, const float* __restrict__ in_scale,
                                                                        const float* __restrict__ in_shift, const unsigned char* __restrict__ wbf,
                                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                                        const float* __restrict__ target, float* __restrict__ partials,
                                                                        int total_items, long long* __restrict__ dbg) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t img_hi = base, img_lo = base + dr::PLANE, wsm = base + dr::OFF_W, bars = base + dr::OFF_BARS;
    // mbarriers: full[4] empty[4] tfull[4] tempty[4] wfull = 17 x 8 B
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + dr::OFF_BARS + 192);
    float* s_bnl = reinterpret_cast<float*>(smem + dr::OFF_BARS + 256);   // [2][64] scale, shift
    float* s_red = s_bnl + 128;                                           // [4]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (dr::NSLOT + s); };
    auto tfull_bar = [&](int i) { return bars + 8u * (2 * dr::NSLOT + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (2 * dr::NSLOT + 4 + i); };
    const uint32_t wfull = bars + 8u * (2 * dr::NSLOT + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = (int)((long long)total_items * blockIdx.x / gridDim.x), i1 = (int)((long long)total_items * (blockIdx.x + 1) / gridDim.x);
    const int g_lo = dr::g0_of(i0), g_hi = i1 > i0 ? dr::g0_of(i1 - 1) + 1 : g_lo - 1;

    if (tid == 0) {
        for (int s = 0; s < dr::NSLOT; ++s) { mbar_init(full_bar(s), 8); mbar_init(empty_bar(s), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (tid < 64) { s_bnl[tid] = in_scale[tid]; s_bnl[64 + tid] = in_shift[tid]; }
    for (int e = tid; e < 2 * dr::PLANE / 16; e += dr::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (tid == 0) {
        mbar_arrive_expect_tx(wfull, dr::W_BYTES);
        bulk_g2s(wsm, wbf, dr::W_BYTES, wfull);
    }

    if (warp >= 8) {
        // ================================ producers: one input row (111 pixels x 64 channels) per step ================================
        // thread = (16-byte chunk jc of 8 channels, pixel group pg): pixels pg, pg+32, pg+64, pg+96 -- a warp instruction covers
        // 4 whole pixels (4 x 256 contiguous bytes of global memory, 4 whole 128-byte image rows of shared memory) and the
        // thread's 8 BatchNorm scale / shift values live in registers
        const int pidx = tid - 256, jc = pidx & 7, pg = pidx >> 3;
        const float4 sc0 = *reinterpret_cast<const float4*>(s_bnl + jc * 8), sc1 = *reinterpret_cast<const float4*>(s_bnl + jc * 8 + 4);
        const float4 sh0 = *reinterpret_cast<const float4*>(s_bnl + 64 + jc * 8), sh1 = *reinterpret_cast<const float4*>(s_bnl + 64 + jc * 8 + 4);
        constexpr int PF = 4;
        // row image g = n*RPI + (yy + 1) holds input row yy; yy = -1 and yy = 111 are zero rows
        auto row_src = [&](int g) -> const float* {
            const int n = g / dr::RPI, yy = g - n * dr::RPI - 1;
            return (yy >= 0 && yy < dr::IN) ? in + ((size_t)n * dr::IN + yy) * dr::IN * SRLZ_C : nullptr;
        };
        auto prefetch = [&](int g) {   // 111 x 256 B = 222 lines
            if (g <= g_hi && pidx < 2 * dr::IN) {
                const float* src = row_src(g);
                if (src != nullptr) prefetch_l2(src + pidx * 32);
            }
        };
        auto load = [&](int g, float4 (&d)[8]) -> bool {
            const float* src = row_src(g);
            if (src == nullptr) return false;
            src += pg * SRLZ_C + jc * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (pg + 32 * q < dr::IN) ldg8(src + q * 32 * SRLZ_C, d[2 * q], d[2 * q + 1]);
            return true;
        };
        auto step = [&](int g, const float4 (&v)[8], bool real) {
            const int rel = g - g_lo, slot = rel % dr::NSLOT, ph = (rel / dr::NSLOT) & 1;
            if (pidx == 0) DR_STAMP(rel, 0);
            mbar_wait(empty_bar(slot), ph ^ 1);
            if (pidx == 0) DR_STAMP(rel, 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (pg + 32 * q < dr::IN) {
                    const int row = slot * 128 + pg + 32 * q + 1;
                    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                    if (real) split8(bn_relu4(v[2 * q], sc0, sh0), bn_relu4(v[2 * q + 1], sc1, sh1), hi, lo);
                    unsigned char* dst = smem + row * 128 + ((jc ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = hi;
                    *reinterpret_cast<uint4*>(dst + dr::PLANE) = lo;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(slot));
            if (pidx == 0) DR_STAMP(rel, 2);
        };
        if (g_lo <= g_hi) {
            float4 va[8], vb[8];
            bool ra, rb = false;
            for (int d = 1; d < PF; ++d) prefetch(g_lo + d);
            ra = load(g_lo, va);
            for (int g = g_lo; g <= g_hi; g += 2) {
                prefetch(g + PF);
                if (g + 1 <= g_hi) rb = load(g + 1, vb);
                step(g, va, ra);
                if (g + 1 <= g_hi) {
                    prefetch(g + 1 + PF);
                    if (g + 2 <= g_hi) ra = load(g + 2, va);
                    step(g + 1, vb, rb);
                }
            }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        mbar_wait(wfull, 0);
        int ready = 0, it = 0;
        for (int i = i0; i < i1; ++i, ++it) {
            const int buf = it & 3;
            if (lane == 0) DR_STAMP(it, 3);
            mbar_wait(tempty_bar(buf), ((it >> 2) & 1) ^ 1);
            const int g0r = dr::g0_of(i) - g_lo;
            for (; ready <= g0r + 1; ++ready) mbar_wait(full_bar(ready % dr::NSLOT), (ready / dr::NSLOT) & 1);
            tc_fence_after();
            if (lane == 0) DR_STAMP(it, 4);
            const int nxt = i + 1 < i1 ? dr::g0_of(i + 1) - g_lo : g0r + 2;   // row images below `nxt` are not needed again
            if (leader) {
                const uint32_t d_tmem = tmem_base + buf * 32;
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {   // grp 0: input row y-1 (dy = 1), grp 1: input row y (dy = 0)
                    const int slot = (g0r + grp) % dr::NSLOT, dy = 1 - grp;
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const uint32_t off = slot * dr::SLOT_BYTES + (1 - dx) * 128;   // accumulator row x reads pixel x - dx = image row x - dx + 1
                        const uint64_t ahi = make_desc_sw128(img_hi + off), alo = make_desc_sw128(img_lo + off);
                        // bf16x3 in two MMAs per K step: the weight image of a shift is [hi 16 rows | lo 16 rows], so one N = 32 MMA
                        // gives A_hi*W_hi (columns 0-15) and A_hi*W_lo (columns 16-31) with a single read of the A tile; A_lo*W_hi is
                        // an N = 16 MMA into columns 0-15; the epilogue adds the two column halves
                        const uint64_t whl = make_desc_sw128(wsm + (dy * 2 + dx) * 4096);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_bf16(d_tmem, ahi + adv, whl + adv, dr::IDESC32, (grp | dx | k) ? 1u : 0u);
                            umma_bf16(d_tmem, alo + adv, whl + adv, dr::IDESC, 1u);
                        }
                    }
                    if (g0r + grp < nxt) umma_commit(empty_bar(slot));
                }
                umma_commit(tfull_bar(buf));
            }
            __syncwarp();
            if (lane == 0) DR_STAMP(it, 5);
        }
    } else if (warp < 4) {
        // ================================ epilogue ================================
        // accumulator row x = output columns 2x, 2x+1 of image rows 2y, 2y+1; a warp stores 32 consecutive float2 (256 B) per
        // (co, py); the squared error against the target is accumulated per thread and reduced once at the end
        const float bia[3] = {bias[0], bias[1], bias[2]};
        const int x = tid;
        const bool valid = x < dr::NY;
        float sse = 0.f;
        int it = 0;
        for (int i = i0; i < i1; ++i, ++it) {
            const int buf = it & 3;
            const int n = i / dr::NY, y0 = i - n * dr::NY;
            float2 tg[3][2];
            if (target != nullptr && valid) {   // target fetched before the accumulator is waited for
#pragma unroll
                for (int co = 0; co < 3; ++co)
#pragma unroll
                    for (int py = 0; py < 2; ++py)
                        tg[co][py] = __ldg(reinterpret_cast<const float2*>(target + (((size_t)n * 3 + co) * dr::OUT + 2 * y0 + py) * dr::OUT + 2 * x));
            }
            if (tid == 0) DR_STAMP(it, 6);
            mbar_wait(tfull_bar(buf), (it >> 2) & 1);
            tc_fence_after();
            if (tid == 0) DR_STAMP(it, 7);
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 32, v);
#pragma unroll
            for (int j = 0; j < 12; ++j) v[j] += v[16 + j];
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
            if (valid) {
#pragma unroll
                for (int co = 0; co < 3; ++co)
#pragma unroll
                    for (int py = 0; py < 2; ++py) {
                        const float2 o = make_float2(v[(py * 2 + 0) * 3 + co] + bia[co], v[(py * 2 + 1) * 3 + co] + bia[co]);
                        *reinterpret_cast<float2*>(out + (((size_t)n * 3 + co) * dr::OUT + 2 * y0 + py) * dr::OUT + 2 * x) = o;
                        if (target != nullptr) {
                            const float e0 = o.x - tg[co][py].x, e1 = o.y - tg[co][py].y;
                            sse = fmaf(e0, e0, sse);
                            sse = fmaf(e1, e1, sse);
                        }
                    }
            }
            if (tid == 0) DR_STAMP(it, 8);
        }
        sse = warp_sum(sse);
        if (lane == 0) s_red[warp] = sse;
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0 && partials != nullptr) partials[blockIdx.x] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    if (warp == 4) tmem_dealloc(tmem_base, 128);
}

// a.in = pre-BN input (B,111,111,64) with a.in_scale/in_shift, a.bias = (3), a.out = decoded (B,3,224,224) NCHW,
// a.aux2 = target or null, a.partials = per-CTA squared-error partials (one float per CTA) or null;
// wbf = the 16 KB image written by pack_dec12_fwd_bf16 (conv_halo_tc.cu)
int dec12_rows_fwd(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st) {
    const int total = a.g.B * dr::NY;
    int gx = sm_count();
    if (gx > total) gx = total;
    if (n_partials) *n_partials = gx;
    if (a.in_scale == nullptr || a.bias == nullptr) { set_error("dec12_rows_fwd: BN scale/shift and bias required"); return 1; }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dec12_rows_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dr::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("dec12_rows_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    dec12_rows_fwd_kernel<<<gx, dr::THREADS, dr::SMEM_BYTES, st>>>(a.in, a.in_scale, a.in_shift, reinterpret_cast<const unsigned char*>(wbf), a.bias,
                                                                  a.out, a.aux2, a.aux2 != nullptr ? a.partials : nullptr, total, a.dbg);
    return check_launch("dec12_rows_fwd");
}

}  // namespace srlz

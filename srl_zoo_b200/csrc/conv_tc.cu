// tcgen05 implicit-GEMM kernel for the 64->64 channel layers (sm_100a): forward and dgrad of
// conv3x3 (models/models.py:54,59) and ConvTranspose2d k3 s2 (models/models.py:66-78).
//
//   D[128 pixels x 64 channels] (fp32, TMEM)  +=  A[128 x 64] (gathered tap, bf16) * W_tap[64 x 64] (bf16)
//
// fp32 fidelity on bf16 tensor cores: every fp32 operand is split x = hi + lo (two bf16) and each k-step issues
// three MMAs  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (the dropped lo*lo term is ~2^-16 relative), accumulating in fp32.
//
// Warp roles (416 threads, 1 CTA per SM, persistent over tiles):
//   warps 0-3  epilogue : tcgen05.ld TMEM -> registers, +bias / BN statistics / ReLU-mask+BN-backward statistics, 128-bit stores
//   warp  4    MMA      : one lane issues tcgen05.mma (cta_group::1, kind::f16, M=128 N=64 K=16), tcgen05.commit -> mbarriers
//   warps 5-12 producers: gather the tap's 128x64 fp32 tile from the NHWC tensor (optional BN+ReLU on load), split to
//                         bf16 hi/lo and write the canonical K-major SWIZZLE_128B image; weights arrive by cp.async.bulk
// Pipelines: 4 smem stages (full/empty mbarriers) and 2 TMEM accumulators (tmem full/empty mbarriers).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace srlz {

namespace tc {
constexpr int NS = 4;
constexpr int A_BYTES = 128 * 128;               // one bf16 plane of the A tile (128 rows x 128 B)
constexpr int W_BYTES = 64 * 128;                // one bf16 plane of the weight tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
constexpr int SMEM_BYTES = NS * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 4 * 64 * 4 /*bn consts*/ + 4 * 128 * 4 /*stat red*/;
constexpr int THREADS = 13 * 32;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // f32 acc, bf16 x bf16, K-major, N=64, M=128
}  // namespace tc

// ----------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// warp-level reduce-scatter of 32 per-lane values: lane L ends with the sum over the 32 lanes of element L
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int n = 32, mask = 16; n > 1; n >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < n / 2) {
                const float keep = upper ? v[i + n / 2] : v[i];
                const float send = upper ? v[i] : v[i + n / 2];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
    }
    return v[0];
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// 8 fp32 -> 8 bf16 hi (one uint4) + 8 bf16 lo (one uint4)
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float r[8];
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
        r[2 * i] = x[2 * i] - __bfloat162float(h0);
        r[2 * i + 1] = x[2 * i + 1] - __bfloat162float(h1);
        __nv_bfloat162 hh;
        hh.x = h0;
        hh.y = h1;
        h[i] = *reinterpret_cast<uint32_t*>(&hh);
        l[i] = pack_bf16x2(r[2 * i], r[2 * i + 1]);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct TileInfo {
    int cls, tile_in_cls, py, px, OHc, OWc;
    long long Mc;
};

template <bool TRANSPOSED>
__device__ __forceinline__ bool decode_tile(const ConvGeom& g, int OH, int OW, int tile, TileInfo& t) {
    const int s = TRANSPOSED ? g.stride : 1;
    int rem = tile;
    for (int c = 0; c < s * s; ++c) {
        const int py = c / s, px = c % s;
        const int OHc = (OH - py + s - 1) / s, OWc = (OW - px + s - 1) / s;
        const long long Mc = (long long)g.B * OHc * OWc;
        const int nt = (int)((Mc + 127) / 128);
        if (rem < nt) {
            t.cls = c; t.tile_in_cls = rem; t.py = py; t.px = px; t.OHc = OHc; t.OWc = OWc; t.Mc = Mc;
            return true;
        }
        rem -= nt;
    }
    return false;
}

template <bool TRANSPOSED>
__device__ __forceinline__ bool tap_in_class(const ConvGeom& g, int py, int px, int ky, int kx) {
    if (!TRANSPOSED) return true;
    const int s = g.stride;
    return (((py + g.pad - ky) % s + s) % s == 0) && (((px + g.pad - kx) % s + s) % s == 0);
}

template <bool TRANSPOSED, bool BN_LOAD, int EPI>
__global__ void __launch_bounds__(tc::THREADS, 1) gconv64_tc_kernel(GConvArgs a, const unsigned char* __restrict__ wbf, int total_tiles) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t bars = base + tc::NS * tc::STAGE_BYTES;   // full[NS], empty[NS], tfull[2], tempty[2] (8 B each), tmem ptr
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + tc::NS * tc::STAGE_BYTES + 128);
    float* s_bn = reinterpret_cast<float*>(smem + tc::NS * tc::STAGE_BYTES + 512);  // [4][64] scale, shift, mean, invstd | bias in row 0 for fwd
    float* s_red = s_bn + 4 * 64;                                                    // [4][128]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (tc::NS + s); };
    auto tfull_bar = [&](int i) { return bars + 8u * (2 * tc::NS + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (2 * tc::NS + 2 + i); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvGeom g = a.g;
    const int s = g.stride;
    const int OH = TRANSPOSED ? g.BH : g.SH, OW = TRANSPOSED ? g.BW : g.SW;
    const int IH = TRANSPOSED ? g.SH : g.BH, IW = TRANSPOSED ? g.SW : g.BW;

    if (tid == 0) {
        for (int i = 0; i < tc::NS; ++i) {
            mbar_init(full_bar(i), 9);   // 8 producer warps + 1 expect_tx arrival (weights)
            mbar_init(empty_bar(i), 1);  // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar(i), 1);   // tcgen05.commit
            mbar_init(tempty_bar(i), 4);  // 4 epilogue warps
        }
        fence_barrier_init();
    }
    if (tid < 64) {
        if (EPI == EPI_MASK_BNBWD) {
            s_bn[tid] = a.e_scale[tid];
            s_bn[64 + tid] = a.e_shift[tid];
            s_bn[128 + tid] = a.e_mean[tid];
            s_bn[192 + tid] = a.e_invstd[tid];
        } else {
            s_bn[tid] = a.bias != nullptr ? a.bias[tid] : 0.f;
        }
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 5) {
        // ================================ producers ================================
        const int pidx = tid - 160, pix = pidx >> 1, half = pidx & 1;
        float4 lsc[8], lsh[8];
        if (BN_LOAD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                lsc[j] = ldg4(a.in_scale + half * 32 + j * 4);
                lsh[j] = ldg4(a.in_shift + half * 32 + j * 4);
            }
        }
        int stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            TileInfo t;
            decode_tile<TRANSPOSED>(g, OH, OW, tile, t);
            const long long m = (long long)t.tile_in_cls * 128 + pix;
            const bool mvalid = m < t.Mc;
            int n = 0, oy = 0, ox = 0;
            if (mvalid) {
                const int oxc = (int)(m % t.OWc);
                const long long q = m / t.OWc;
                const int oyc = (int)(q % t.OHc);
                n = (int)(q / t.OHc);
                oy = oyc * (TRANSPOSED ? s : 1) + t.py;
                ox = oxc * (TRANSPOSED ? s : 1) + t.px;
            }
            for (int ky = 0; ky < g.KH; ++ky) {
                for (int kx = 0; kx < g.KW; ++kx) {
                    if (!tap_in_class<TRANSPOSED>(g, t.py, t.px, ky, kx)) continue;
                    const int tap = ky * g.KW + kx;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    unsigned char* st_base = smem + stage * tc::STAGE_BYTES;
                    if (pidx == 0) {
                        mbar_arrive_expect_tx(full_bar(stage), 2 * tc::W_BYTES);
                        bulk_g2s(base + stage * tc::STAGE_BYTES + 2 * tc::A_BYTES, wbf + (size_t)tap * (2 * tc::W_BYTES), 2 * tc::W_BYTES,
                                 full_bar(stage));
                    }
                    int iy, ix;
                    bool ok = mvalid;
                    if (TRANSPOSED) {
                        const int ny = oy + g.pad - ky, nx = ox + g.pad - kx;
                        iy = ny / s;
                        ix = nx / s;
                        ok = ok && ny >= 0 && nx >= 0 && iy < IH && ix < IW;
                    } else {
                        iy = oy * s - g.pad + ky;
                        ix = ox * s - g.pad + kx;
                        ok = ok && iy >= 0 && ix >= 0 && iy < IH && ix < IW;
                    }
                    float4 v[8];
                    if (ok) {
                        const float* src = a.in + (((size_t)n * IH + iy) * IW + ix) * SRLZ_C + half * 32;
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = ldg4(src + j * 4);
                        if (BN_LOAD) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = bn_relu4(v[j], lsc[j], lsh[j]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 hi, lo;
                        split8(v[2 * j], v[2 * j + 1], hi, lo);
                        const int chunk = (half * 4 + j) ^ (pix & 7);
                        *reinterpret_cast<uint4*>(st_base + pix * 128 + chunk * 16) = hi;
                        *reinterpret_cast<uint4*>(st_base + tc::A_BYTES + pix * 128 + chunk * 16) = lo;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full_bar(stage));
                    if (++stage == tc::NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        int stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            TileInfo t;
            decode_tile<TRANSPOSED>(g, OH, OW, tile, t);
            const int acc = it & 1;
            mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 64;
            uint32_t first = 1;
            for (int ky = 0; ky < g.KH; ++ky) {
                for (int kx = 0; kx < g.KW; ++kx) {
                    if (!tap_in_class<TRANSPOSED>(g, t.py, t.px, ky, kx)) continue;
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sb = base + stage * tc::STAGE_BYTES;
                        const uint64_t ahi = make_desc_sw128(sb), alo = make_desc_sw128(sb + tc::A_BYTES);
                        const uint64_t whi = make_desc_sw128(sb + 2 * tc::A_BYTES), wlo = make_desc_sw128(sb + 2 * tc::A_BYTES + tc::W_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 16 bf16 = 32 B along K inside the 128 B swizzle row
                            umma_bf16(d_tmem, alo + adv, whi + adv, tc::IDESC, first ? 0u : 1u);
                            first = 0;
                            umma_bf16(d_tmem, ahi + adv, wlo + adv, tc::IDESC, 1u);
                            umma_bf16(d_tmem, ahi + adv, whi + adv, tc::IDESC, 1u);
                        }
                        umma_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == tc::NS) { stage = 0; phase ^= 1; }
                }
            }
            if (lane == 0) umma_commit(tfull_bar(acc));
            __syncwarp();
        }
    } else {
        // ================================ epilogue (warps 0-3) ================================
        float st1[2] = {0.f, 0.f}, st2[2] = {0.f, 0.f};
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            TileInfo t;
            decode_tile<TRANSPOSED>(g, OH, OW, tile, t);
            const int acc = it & 1;
            const long long m = (long long)t.tile_in_cls * 128 + tid;
            const bool mvalid = m < t.Mc;
            size_t off = 0;
            if (mvalid) {
                const int oxc = (int)(m % t.OWc);
                const long long q = m / t.OWc;
                const int oyc = (int)(q % t.OHc);
                const int n = (int)(q / t.OHc);
                const int oy = oyc * (TRANSPOSED ? s : 1) + t.py, ox = oxc * (TRANSPOSED ? s : 1) + t.px;
                off = (((size_t)n * OH + oy) * OW + ox) * SRLZ_C;
            }
            mbar_wait(tfull_bar(acc), (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64 + h * 32, v);
                if (h == 1) {
                    // both halves are in registers / consumed: release the accumulator to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                float q2[32];
                if (EPI == EPI_MASK_BNBWD) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 yp = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (mvalid) yp = ldg4(a.e_ypre + off + h * 32 + j * 4);
                        const float ypv[4] = {yp.x, yp.y, yp.z, yp.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = h * 32 + j * 4 + e;
                            const bool on = mvalid && fmaf(ypv[e], s_bn[c], s_bn[64 + c]) > 0.f;
                            const float dz = on ? v[j * 4 + e] : 0.f;
                            v[j * 4 + e] = dz;
                            q2[j * 4 + e] = dz * ((ypv[e] - s_bn[128 + c]) * s_bn[192 + c]);
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float y = mvalid ? v[i] + s_bn[h * 32 + i] : 0.f;
                        v[i] = y;
                        q2[i] = y * y;
                    }
                }
                if (mvalid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4(a.out + off + h * 32 + j * 4, make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]));
                }
                if (EPI != EPI_PLAIN) {
                    st1[h] += warp_reduce_scatter32(v, lane);
                    st2[h] += warp_reduce_scatter32(q2, lane);
                }
            }
        }
        if (EPI != EPI_PLAIN) {
            // lane L of warp w holds the sums over its rows for channels L (h=0) and 32+L (h=1)
            s_red[warp * 128 + lane] = st1[0];
            s_red[warp * 128 + 32 + lane] = st1[1];
            s_red[warp * 128 + 64 + lane] = st2[0];
            s_red[warp * 128 + 96 + lane] = st2[1];
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
    if (warp == 4) tmem_dealloc(tmem_base, 128);
}

template <bool T, bool BN, int EPI>
static int launch_tc(const GConvArgs& a, const unsigned char* wbf, int total_tiles, int gx, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gconv64_tc_kernel<T, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gconv64_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    gconv64_tc_kernel<T, BN, EPI><<<gx, tc::THREADS, tc::SMEM_BYTES, st>>>(a, wbf, total_tiles);
    return check_launch("gconv64_tc");
}

// wbf: bf16 weight image produced by pack_conv_w_bf16 ([tap]{hi 8 KB | lo 8 KB}, rows = output channel, SWIZZLE_128B)
int gconv64_tc(const GConvArgs& a_in, const void* wbf, int* n_partials, cudaStream_t st) {
    GConvArgs a = a_in;
    const ConvGeom& g = a.g;
    const int OH = a.transposed ? g.BH : g.SH, OW = a.transposed ? g.BW : g.SW;
    const int s = a.transposed ? g.stride : 1;
    int total = 0;
    for (int c = 0; c < s * s; ++c) {
        const int py = c / s, px = c % s;
        const long long Mc = (long long)g.B * ((OH - py + s - 1) / s) * ((OW - px + s - 1) / s);
        total += (int)((Mc + 127) / 128);
    }
    int gx = sm_count();
    if (gx > total) gx = total;
    if (n_partials) *n_partials = gx;
    if (a.epi != EPI_PLAIN && a.partials == nullptr) { set_error("gconv64_tc: partials buffer required"); return 1; }
    const bool bn = a.in_scale != nullptr;
    const unsigned char* w = reinterpret_cast<const unsigned char*>(wbf);
#define TC_DISPATCH(T, BN, E) return launch_tc<T, BN, E>(a, w, total, gx, st)
    if (a.transposed) {
        if (bn) { if (a.epi == EPI_PLAIN) TC_DISPATCH(true, true, EPI_PLAIN); if (a.epi == EPI_STATS) TC_DISPATCH(true, true, EPI_STATS); TC_DISPATCH(true, true, EPI_MASK_BNBWD); }
        else    { if (a.epi == EPI_PLAIN) TC_DISPATCH(true, false, EPI_PLAIN); if (a.epi == EPI_STATS) TC_DISPATCH(true, false, EPI_STATS); TC_DISPATCH(true, false, EPI_MASK_BNBWD); }
    } else {
        if (bn) { if (a.epi == EPI_PLAIN) TC_DISPATCH(false, true, EPI_PLAIN); if (a.epi == EPI_STATS) TC_DISPATCH(false, true, EPI_STATS); TC_DISPATCH(false, true, EPI_MASK_BNBWD); }
        else    { if (a.epi == EPI_PLAIN) TC_DISPATCH(false, false, EPI_PLAIN); if (a.epi == EPI_STATS) TC_DISPATCH(false, false, EPI_STATS); TC_DISPATCH(false, false, EPI_MASK_BNBWD); }
    }
#undef TC_DISPATCH
}

// fp32 pack [tap][k][n]  ->  bf16 image [tap]{hi[n][k], lo[n][k]} in the K-major SWIZZLE_128B layout (row n = 128 B)
__global__ void pack_conv_w_bf16_kernel(const float* __restrict__ src, unsigned char* __restrict__ dst, int ntaps) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over tap*4096 + n*64 + k
    if (idx >= ntaps * 4096) return;
    const int tap = idx >> 12, n = (idx >> 6) & 63, k = idx & 63;
    const float x = src[(tap * 64 + k) * 64 + n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const int byte = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    unsigned char* t = dst + (size_t)tap * (2 * tc::W_BYTES);
    *reinterpret_cast<__nv_bfloat16*>(t + byte) = hi;
    *reinterpret_cast<__nv_bfloat16*>(t + tc::W_BYTES + byte) = lo;
}

int pack_conv_w_bf16(const float* pack_f32, void* dst, int ntaps, cudaStream_t st) {
    pack_conv_w_bf16_kernel<<<(ntaps * 4096 + 255) / 256, 256, 0, st>>>(pack_f32, reinterpret_cast<unsigned char*>(dst), ntaps);
    return check_launch("pack_conv_w_bf16");
}

}  // namespace srlz

// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the sm_100a tensor-core kernels.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace srlz {

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
// One elected lane of a CONVERGED warp (elect.sync): code guarded by this predicate stays on the uniform datapath, so the
// tcgen05 instructions inside it are issued directly (an `if (lane == 0)` guard makes the compiler wrap every
// tcgen05.mma in an ELECT / R2UR / BRA.U.ANY waterfall, ~12 extra instructions per MMA on the single issuing thread).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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
// 8 fp32 -> 8 bf16 hi (one uint4) + 8 bf16 lo (one uint4).  Packed cvt.rn.bf16x2.f32 for both planes; float(hi) is
// recovered with a shift / mask (bf16 is the upper half of the fp32 pattern), lo = bf16(x - float(hi)).
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);                 // low half = element 2i, high half = element 2i+1
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = pack_bf16x2(x[2 * i] - h0, x[2 * i + 1] - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- 8x16-pixel tiles with the source patch staged in shared memory (dec12 gradient columns) ----
// mode 2 (dec12 dgrad): patch = 3 channels x 18 rows x 34 cols of d(decoded) below input tile (y0,x0), fetched as 9 aligned
// 16-byte chunks per row (the tile origin 2*x0 is a multiple of 32 columns; row stride 36 floats)
template <int MODE> struct PatchGeom;
constexpr int PATCH_THREADS = 224;   // the seven producer warps of the dec12 dgrad kernel
template <> struct PatchGeom<2> { static constexpr int PR = 18, PC = 34, PS = 36, NCH = 9, N = 3 * 18 * 9, PER = (3 * 18 * 9 + PATCH_THREADS - 1) / PATCH_THREADS, FLOATS = 3 * 18 * 36, NT = 1; };
constexpr int PATCH_MAX_FLOATS = 2560;

struct PatchSrc {            // what the patch is read from
    const float* g;          // explicit d(decoded) or null
    const float* dec;        // decoded
    const float* tgt;        // target
    float coef;
};

// per-thread chunk coordinates of the patch (independent of the tile): chunk e = pidx + PATCH_THREADS j
template <int MODE>
struct PatchIdx {
    int so[PatchGeom<MODE>::PER];     // shared-memory offset, -1 = no element
    int go[PatchGeom<MODE>::PER];     // global offset relative to the tile origin (channel plane + row*224 + col)
    short rr[PatchGeom<MODE>::PER], cc[PatchGeom<MODE>::PER];
    __device__ __forceinline__ void init(int pidx) {
        using G = PatchGeom<MODE>;
#pragma unroll
        for (int j = 0; j < G::PER; ++j) {
            const int e = pidx + PATCH_THREADS * j;
            if (e < G::N) {
                const int c = 4 * (e % G::NCH), r = (e / G::NCH) % G::PR, ci = e / (G::NCH * G::PR);
                so[j] = (ci * G::PR + r) * G::PS + c;
                go[j] = ci * (224 * 224) + r * 224 + c;
                rr[j] = (short)r; cc[j] = (short)c;
            } else {
                so[j] = -1; go[j] = 0; rr[j] = 0; cc[j] = 0;
            }
        }
    }
};

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src, bool valid) {
    const uint32_t sz = valid ? 4u : 0u;   // src-size 0: the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;  // src-size 0: 16 zero bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16_n(uint32_t dst_smem, const void* src, uint32_t src_bytes) {   // the rest of the 16 bytes: zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Asynchronous (LDGSTS) prefetch of one tile's source patch straight into shared memory: no registers and no scoreboard
// slots are held while the data is in flight (register prefetches stall the barrier polls that share a scoreboard slot).
// bufA: explicit gradient or decoded; bufB: target (fused gradient only).
template <int MODE>
__device__ __forceinline__ void patch_prefetch(const PatchIdx<MODE>& ix, const PatchSrc& s, int n, int y0, int x0, uint32_t bufA,
                                               uint32_t bufB) {
    using G = PatchGeom<MODE>;
    const int oy0 = 2 * y0, ox0 = 2 * x0;
    const long long base = (long long)n * 3 * 224 * 224 + (long long)oy0 * 224 + ox0;
#pragma unroll
    for (int j = 0; j < G::PER; ++j) {
        if (ix.so[j] >= 0) {
            const int iy = oy0 + ix.rr[j], left = 224 - (ox0 + ix.cc[j]);    // columns of the chunk inside the image
            const uint32_t bytes = iy < 224 && left > 0 ? 4u * (uint32_t)min(left, 4) : 0u;
            const long long off = bytes ? base + ix.go[j] : 0;
            if (s.g != nullptr) {
                cp_async16_n(bufA + ix.so[j] * 4, s.g + off, bytes);
            } else {
                cp_async16_n(bufA + ix.so[j] * 4, s.dec + off, bytes);
                cp_async16_n(bufB + ix.so[j] * 4, s.tgt + off, bytes);
            }
        }
    }
    cp_async_commit();
}

// the 32 K-slots [HALF*32, HALF*32+32) of pixel (py,px) of the tile
template <int MODE, int HALF>
__device__ __forceinline__ void patch_gather(float (&vf)[32], const float* buf, const float* bufB, bool fused, float coef, int c, int py,
                                             int px) {
    using G = PatchGeom<MODE>;
    {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int cky = HALF * 8 + q;   // co*4 + ky
            if (cky < 12) {
                const int o = ((cky >> 2) * G::PR + 2 * py + (cky & 3)) * G::PS + 2 * px;
                float2 g0 = *reinterpret_cast<const float2*>(buf + o), g1 = *reinterpret_cast<const float2*>(buf + o + 2);
                if (fused) {   // d(decoded) = coef * (decoded - target)
                    const float2 t0 = *reinterpret_cast<const float2*>(bufB + o), t1 = *reinterpret_cast<const float2*>(bufB + o + 2);
                    g0 = make_float2(coef * (g0.x - t0.x), coef * (g0.y - t0.y));
                    g1 = make_float2(coef * (g1.x - t1.x), coef * (g1.y - t1.y));
                }
                vf[q * 4 + 0] = g0.x; vf[q * 4 + 1] = g0.y; vf[q * 4 + 2] = g1.x; vf[q * 4 + 3] = g1.y;
            } else {
                vf[q * 4 + 0] = 0.f; vf[q * 4 + 1] = 0.f; vf[q * 4 + 2] = 0.f; vf[q * 4 + 3] = 0.f;
            }
        }
    }
}

// convert the 32 gathered values of this thread's half row and write them into the SWIZZLE_128B image (hi / lo planes);
// NJ < 4: only the first NJ 16-byte chunks (8 values each) -- the rest of the half row is known to be zero in the image already
template <int NJ = 4>
__device__ __forceinline__ void store_half_row(const float (&vf)[32], unsigned char* dst_hi, unsigned char* dst_lo, int pix, int half) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        uint4 hi, lo;
        split8(make_float4(vf[8 * j], vf[8 * j + 1], vf[8 * j + 2], vf[8 * j + 3]),
               make_float4(vf[8 * j + 4], vf[8 * j + 5], vf[8 * j + 6], vf[8 * j + 7]), hi, lo);
        const int chunk = (half * 4 + j) ^ (pix & 7);
        *reinterpret_cast<uint4*>(dst_hi + pix * 128 + chunk * 16) = hi;
        *reinterpret_cast<uint4*>(dst_lo + pix * 128 + chunk * 16) = lo;
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void producers_bar_sync() { asm volatile("bar.sync 1, 224;" ::: "memory"); }   // the PATCH_THREADS producer threads

}  // namespace srlz

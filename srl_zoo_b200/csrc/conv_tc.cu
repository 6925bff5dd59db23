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
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "bn_tail.cuh"

namespace srlz {

template <int N> struct IntC { static constexpr int value = N; };

namespace tc {
constexpr int NS = 4;
constexpr int A_BYTES = 128 * 128;               // one bf16 plane of the A tile (128 rows x 128 B)
constexpr int W_BYTES = 64 * 128;                // one bf16 plane of the weight tile
constexpr int STAGE_BYTES = 2 * A_BYTES;             // an A stage: hi plane + lo plane (32 KB)
constexpr int WSLOT_BYTES = 2 * W_BYTES;             // a weight slot: hi + lo (16 KB)
// MODE 0 : 2 A stages (64 KB) | 3 raw fp32 slots (96 KB) | 2 weight slots (32 KB) | epilogue staging (16 KB) | misc
// MODE 2  : 3 A stages (96 KB) | 3 weight slots (48 KB) | misc | 4 source-patch buffers (40 KB) | epilogue staging (16 KB)
constexpr int MISC_BYTES = 512 /*barriers*/ + 4 * 64 * 4 /*bn consts*/ + 4 * 128 * 4 /*stat red*/ + 2 * 64 * 4 /*load-side bn*/;
constexpr int SMEM_BYTES = 208 * 1024 + MISC_BYTES + 1024 /*align*/;
constexpr int THREADS = 16 * 32;   // warpgroups: 0 epilogue | 1 MMA issuer (warp 4) + 3 register-donor warps | 2,3 producers
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // f32 acc, bf16 x bf16, K-major, N=64, M=128
}  // namespace tc

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

#define TC_STAMP(itv, slot) do { if (a.dbg != nullptr && blockIdx.x == 0 && (itv) < 64) a.dbg[(itv) * 16 + (slot)] = clock64(); } while (0)

template <bool TRANSPOSED, bool BN_LOAD, int EPI, int MODE>
__global__ void __launch_bounds__(tc::THREADS, 1) gconv64_tc_kernel(GConvArgs a, const unsigned char* __restrict__ wbf, int total_tiles) {
    pdl_enter();
    constexpr int NSB = 2;                       // bf16 A stages in the MMA ring
    // Warp roles.  MODE 0: warps 0-3 epilogue | 4 MMA issuer | 5 weight loader (6-7 register donors) | 8-15 producers.
    // MODE 2 (dec12 dgrad: K = 48, one tap -- twelve MMAs per tile; the kernel is bound by its epilogue, which reads the saved
    // pre-activation and writes dz for 128 pixels x 64 channels): EIGHT epilogue warps -- warp w takes TMEM lane quarter w & 3
    // and channel half w >> 2 -- | 8-14 producers (warp 8 also stages the 32 items left over) | 15 MMA issuer with the one
    // 16 KB weight image resident (loaded once).  Measured: 4 epilogue warps = one per scheduler at an IPC of 0.13 per tile.
    constexpr bool WIDE = MODE == 2;
    constexpr int NPW = WIDE ? 7 : 8;            // producer warps
    constexpr bool YP_SMEM = EPI == EPI_MASK_BNBWD && MODE == 2;   // pre-activations of the next tile prefetched by LDGSTS
    constexpr uint32_t YP_OFF = 2 * tc::STAGE_BYTES;                // into the third A stage's space (128 rows x 256 B)
    constexpr uint32_t RAW_OFF = 2 * tc::STAGE_BYTES;                          // MODE 0: 3 raw fp32 slots filled by LDGSTS
    constexpr uint32_t W_OFF = MODE == 0 ? RAW_OFF + 3 * 32768 : 3 * tc::STAGE_BYTES;
    constexpr int NW = MODE == 0 ? 2 : 3;        // weight ring slots in use (MODE 0 gives the third slot to the epilogue staging)
    constexpr uint32_t MISC_OFF = W_OFF + 3 * tc::WSLOT_BYTES;
    constexpr uint32_t STG_OFF = MODE == 0 ? W_OFF + 2 * tc::WSLOT_BYTES : MISC_OFF + tc::MISC_BYTES + 4 * PATCH_MAX_FLOATS * 4;  // 4 x 4 KB epilogue staging
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t bars = base + MISC_OFF;   // full[3], empty[3], wfull[3], wempty[3], tfull[2], tempty[2] (8 B each), tmem ptr
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + MISC_OFF + 192);
    float* s_bn = reinterpret_cast<float*>(smem + MISC_OFF + 512);  // [4][64] scale, shift, mean, invstd | bias in row 0 for fwd
    float* s_red = s_bn + 4 * 64;                                    // [4][128]
    float* s_bnl = s_red + 4 * 128;                                  // [2][64] scale, shift applied on load
    float* s_patch = s_bnl + 2 * 64;                                 // [4][PATCH_MAX_FLOATS] (MODE 2)
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (3 + s); };
    auto wfull_bar = [&](int s) { return bars + 8u * (6 + s); };
    auto wempty_bar = [&](int s) { return bars + 8u * (9 + s); };
    auto tfull_bar = [&](int i) { return bars + 8u * (12 + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (14 + i); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvGeom g = a.g;
    const int s = g.stride;
    const int OH = TRANSPOSED ? g.BH : g.SH, OW = TRANSPOSED ? g.BW : g.SW;
    const int IH = TRANSPOSED ? g.SH : g.BH, IW = TRANSPOSED ? g.SW : g.BW;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) {
            mbar_init(full_bar(i), NPW);  // producer warps
            mbar_init(empty_bar(i), 1);   // tcgen05.commit
            mbar_init(wfull_bar(i), 1);   // expect_tx arrival of the weight loader
            mbar_init(wempty_bar(i), 1);  // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar(i), 1);   // tcgen05.commit
            mbar_init(tempty_bar(i), WIDE ? 8 : 4);  // epilogue warps
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
    if (BN_LOAD && tid >= 64 && tid < 128) {
        s_bnl[tid - 64] = a.in_scale[tid - 64];
        s_bnl[64 + tid - 64] = a.in_shift[tid - 64];
    }
    if (WIDE) {   // K slots 48..63 of the two A stages stay zero for the whole kernel (the producers never write them)
        for (int e = tid; e < NSB * tc::STAGE_BYTES / 16; e += tc::THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_ptr_smem), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (WIDE ? (warp >= 8 && warp < 15) : warp >= 8) {
        // ================================ producers ================================
        // Each thread owns half a pixel row (32 channels = 8 x LDG.128) of every stage.  Global loads run two stages
        // ahead of the convert/store work (register ring v0/v1/v2) so that the L2/HBM latency is overlapped.
        if (MODE != 0) {
            // 8x16-pixel tiles; the tile's source patch is staged once in shared memory (double buffered, prefetched
            // through registers) and every K chunk is gathered from it
            constexpr int M = 2;   // the one special producer left: dec12 gradient columns (MODE 2)
            using PG = PatchGeom<M>;
            // 224 threads, 256 (pixel, K-half) items: thread i stages item i.  K = 48 of 64 slots: a half-1 item has 16 values (two
            // chunks; the two chunks of zero padding are written once, before the loop), half the work of a half-0 item, so the
            // last half-1 warp (14) also stages the 32 items left over, 224 + i.
            const int pidx = tid - 256, pix = pidx & 127, half = pidx >> 7, py = pix >> 4, px = pix & 15;
            const int pix2 = 96 + (pidx & 31), py2 = pix2 >> 4, px2 = pix2 & 15;
            const PatchSrc src{a.aux0, a.aux1, a.aux2, a.coef};
            PatchIdx<M> pidx_tab;
            pidx_tab.init(pidx);
            const bool fused = MODE == 2 && a.aux0 == nullptr;
            // patch buffers: A[0], A[1] (observation / gradient / decoded), B[0], B[1] (target, fused dec12 gradient)
            const uint32_t pa = smem_u32(s_patch), pb = pa + 2 * PATCH_MAX_FLOATS * 4;
            int tile = blockIdx.x;
            if (tile < total_tiles) {
                const int tt = tile % 98;
                patch_prefetch<M>(pidx_tab, src, tile / 98, (tt / 7) * 8, (tt % 7) * 16, pa, pb);
            }
            cp_async_wait_all();
            producers_bar_sync();
            int stage = 0, phase = 0;
            for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
                if (pidx == 0) TC_STAMP(it, 0);
                const float* cur = s_patch + (it & 1) * PATCH_MAX_FLOATS;
                const float* curB = cur + 2 * PATCH_MAX_FLOATS;
                const int ntile = tile + gridDim.x;
                if (ntile < total_tiles) {
                    const int tt = ntile % 98;
                    const uint32_t nb_off = ((it + 1) & 1) * PATCH_MAX_FLOATS * 4;
                    patch_prefetch<M>(pidx_tab, src, ntile / 98, (tt / 7) * 8, (tt % 7) * 16, pa + nb_off, pb + nb_off);
                }
                // the patch of the tile after that goes to L2 now (3 channels x 18 rows x two 128-byte lines per tensor, one line
                // per thread), so that the LDGSTS copies issued one tile ahead see L2 latency: they are waited for at the end of
                // every tile and HBM latency did not fit into one tile period any more
                const int ptile = ntile + gridDim.x;
                if (ptile < total_tiles && pidx < 216) {
                    const int tens = pidx >= 108, l = pidx - tens * 108, ci = l / 36, r = (l % 36) >> 1, seg = l & 1;
                    const int tt = ptile % 98, iy = 2 * ((tt / 7) * 8) + r, ix = 2 * ((tt % 7) * 16) + seg * 32;
                    const float* tp = src.g != nullptr ? (tens ? nullptr : src.g) : (tens ? src.tgt : src.dec);
                    if (tp != nullptr && iy < 224 && ix < 224) prefetch_l2(tp + ((size_t)(ptile / 98) * 3 + ci) * (224 * 224) + (size_t)iy * 224 + ix);
                }
#pragma unroll 1
                for (int c = 0; c < PG::NT; ++c) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (pidx == 0 && c == 0) TC_STAMP(it, 1);
                    unsigned char* st_base = smem + stage * tc::STAGE_BYTES;
                    float vf[32];
                    if (half == 0) {
                        patch_gather<M, 0>(vf, cur, curB, fused, a.coef, c, py, px);
                        if (pidx == 0 && c == 0) TC_STAMP(it, 2);
                        store_half_row<4>(vf, st_base, st_base + tc::A_BYTES, pix, 0);
                    } else {
                        patch_gather<M, 1>(vf, cur, curB, fused, a.coef, c, py, px);
                        store_half_row<2>(vf, st_base, st_base + tc::A_BYTES, pix, 1);
                        if (warp == 14) {
                            patch_gather<M, 1>(vf, cur, curB, fused, a.coef, c, py2, px2);
                            store_half_row<2>(vf, st_base, st_base + tc::A_BYTES, pix2, 1);
                        }
                    }
                    if (pidx == 0 && c == 0) TC_STAMP(it, 3);
                    __syncwarp();   // (proxy fence on the consumer side: here it would drain the next tile's patch prefetch)
                    if (lane == 0) mbar_arrive(full_bar(stage));
                    if (pidx == 0 && c == 0) TC_STAMP(it, 4);
                    if (++stage == NSB) { stage = 0; phase ^= 1; }
                }
                if (pidx == 0) TC_STAMP(it, 5);
                cp_async_wait_all();
                if (pidx == 0) TC_STAMP(it, 6);
                producers_bar_sync();
                if (pidx == 0) TC_STAMP(it, 7);
            }
        } else {
        // MODE 0: adjacent threads share a pixel row (coalesced 256 B).  MODE 2: a warp has one `half` (no divergence in
        // the per-slot gathers) and adjacent threads are adjacent pixels.
        const int pidx = tid - 256;
        const int pix = MODE == 0 ? (pidx >> 1) : (pidx & 127), half = MODE == 0 ? (pidx & 1) : (pidx >> 7);
        struct Item { const float* src; int tap; int n, oy, ox; };
        int cur_tile = (int)blockIdx.x - (int)gridDim.x;
        int ky = g.KH, kx = 0;
        TileInfo t;
        bool mvalid = false;
        int n = 0, oy = 0, ox = 0;
        auto next_item = [&](Item& it) -> bool {
            while (true) {
                if (ky < g.KH) {
                    if (++kx == g.KW) { kx = 0; ++ky; }
                }
                if (ky >= g.KH) {
                    cur_tile += gridDim.x;
                    if (cur_tile >= total_tiles) return false;
                    decode_tile<TRANSPOSED>(g, OH, OW, cur_tile, t);
                    const long long m = (long long)t.tile_in_cls * 128 + pix;
                    mvalid = m < t.Mc;
                    if (mvalid) {
                        const int oxc = (int)(m % t.OWc);
                        const long long q = m / t.OWc;
                        const int oyc = (int)(q % t.OHc);
                        n = (int)(q / t.OHc);
                        oy = oyc * (TRANSPOSED ? s : 1) + t.py;
                        ox = oxc * (TRANSPOSED ? s : 1) + t.px;
                    }
                    ky = 0;
                    kx = 0;
                }
                if (tap_in_class<TRANSPOSED>(g, t.py, t.px, ky, kx)) break;
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
            it.tap = ky * g.KW + kx;
            if (MODE != 0) {
                it.src = mvalid ? a.in : nullptr;
                it.n = n; it.oy = oy; it.ox = ox;
                return true;
            }
            it.src = ok ? a.in + (((size_t)n * IH + iy) * IW + ix) * SRLZ_C + half * 32 : nullptr;
            return true;
        };
        // Raw fp32 half rows are staged with LDGSTS (cp.async.cg 16 B) two units ahead into a 3-slot shared-memory ring --
        // no registers or scoreboard slots are held while they are in flight -- and each thread later converts exactly
        // the bytes it requested (so cp.async.wait_group is the only synchronisation the ring needs).
        const uint32_t raw_base = base + RAW_OFF;                             // [3][128 rows][256 B], 16 B chunks XOR-swizzled
        const uint32_t my_row = pix * 256, my_sw = (2 * pix + half) & 7;
        auto issue = [&](const Item& it, int slot) {
            const uint32_t dst = raw_base + slot * 32768 + my_row;
            const float* src = it.src != nullptr ? it.src : a.in;
#pragma unroll
            for (int j = 0; j < 8; ++j) cp_async16(dst + (((half * 8 + j) ^ my_sw) << 4), src + j * 4, it.src != nullptr);
        };
        Item i0{nullptr, 0, 0, 0, 0}, i1{nullptr, 0, 0, 0, 0}, i2{nullptr, 0, 0, 0, 0};
        bool h0 = next_item(i0);
        if (h0) issue(i0, 0);
        cp_async_commit();
        bool h1 = h0 && next_item(i1);
        if (h1) issue(i1, 1);
        cp_async_commit();
        int stage = 0, phase = 0, slot = 0;
        while (h0) {
            const bool h2 = h1 && next_item(i2);
            const int slot2 = slot + 2 >= 3 ? slot - 1 : slot + 2;
            if (h2) issue(i2, slot2);
            cp_async_commit();
            cp_async_wait_group<2>();   // the unit in `slot` has landed
            mbar_wait(empty_bar(stage), phase ^ 1);
            unsigned char* st_base = smem + stage * tc::STAGE_BYTES;
            const unsigned char* rsrc = smem + RAW_OFF + slot * 32768 + my_row;
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(rsrc + (((half * 8 + j) ^ my_sw) << 4));
            if (BN_LOAD && i0.src != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_bnl + half * 32 + j * 4);
                    const float4 sh = *reinterpret_cast<const float4*>(s_bnl + 64 + half * 32 + j * 4);
                    v[j] = bn_relu4(v[j], sc, sh);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 hi, lo;
                split8(v[2 * j], v[2 * j + 1], hi, lo);
                const int chunk = (half * 4 + j) ^ (pix & 7);
                *reinterpret_cast<uint4*>(st_base + pix * 128 + chunk * 16) = hi;
                *reinterpret_cast<uint4*>(st_base + tc::A_BYTES + pix * 128 + chunk * 16) = lo;
            }
            __syncwarp();   // (proxy fence on the consumer side)
            if (lane == 0) mbar_arrive(full_bar(stage));
            if (++stage == NSB) { stage = 0; phase ^= 1; }
            if (++slot == 3) slot = 0;
            i0 = i1; h0 = h1;
            i1 = i2; h1 = h2;
        }
        }
    } else if (WIDE ? warp == 15 : warp >= 4) {
        // ================================ MMA issuer ================================
        // MODE 0: warpgroup 1 donates registers to the epilogue warpgroup (setmaxnreg moves them through the CTA pool)
        if (EPI != EPI_PLAIN && !WIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
        if (!WIDE && warp == 5) {
            const bool wleader = elect_one();
            // ---- weight loader: 16 KB cp.async.bulk per (tile, tap) into its own 3-slot ring, up to three taps ahead ----
            int ws = 0, wph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                TileInfo t;
                t.py = 0; t.px = 0;
                if (MODE == 0) decode_tile<TRANSPOSED>(g, OH, OW, tile, t);
                for (int ky = 0; ky < g.KH; ++ky) {
                    for (int kx = 0; kx < g.KW; ++kx) {
                        if (!tap_in_class<TRANSPOSED>(g, t.py, t.px, ky, kx)) continue;
                        mbar_wait(wempty_bar(ws), wph ^ 1);
                        if (wleader) {
                            mbar_arrive_expect_tx(wfull_bar(ws), tc::WSLOT_BYTES);
                            bulk_g2s(base + W_OFF + ws * tc::WSLOT_BYTES, wbf + (size_t)(ky * g.KW + kx) * tc::WSLOT_BYTES, tc::WSLOT_BYTES, wfull_bar(ws));
                        }
                        __syncwarp();
                        if (++ws == NW) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
        if (WIDE || warp == 4) {
        const bool leader = elect_one();
        int stage = 0, phase = 0, it = 0, ws = 0, wph = 0;
        if (WIDE) {   // the one weight image of the layer: resident in slot 0
            if (leader) {
                mbar_arrive_expect_tx(wfull_bar(0), tc::WSLOT_BYTES);
                bulk_g2s(base + W_OFF, wbf, tc::WSLOT_BYTES, wfull_bar(0));
            }
            __syncwarp();
            mbar_wait(wfull_bar(0), 0);
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            TileInfo t;
            t.py = 0; t.px = 0;
            if (MODE == 0) decode_tile<TRANSPOSED>(g, OH, OW, tile, t);
            const int acc = it & 1;
            if (lane == 0) TC_STAMP(it, 8);
            mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) TC_STAMP(it, 9);
            const uint32_t d_tmem = tmem_base + acc * 64;
            uint32_t first = 1;
            for (int ky = 0; ky < g.KH; ++ky) {
                for (int kx = 0; kx < g.KW; ++kx) {
                    if (!tap_in_class<TRANSPOSED>(g, t.py, t.px, ky, kx)) continue;
                    mbar_wait(full_bar(stage), phase);
                    if (!WIDE) mbar_wait(wfull_bar(ws), wph);
                    fence_proxy_async_smem();   // consumer-side proxy fence (producers: st.shared -> __syncwarp -> mbarrier.arrive)
                    tc_fence_after();
                    if (leader) {
                        const uint32_t sb = base + stage * tc::STAGE_BYTES, wb = base + W_OFF + ws * tc::WSLOT_BYTES;
                        const uint64_t ahi = make_desc_sw128(sb), alo = make_desc_sw128(sb + tc::A_BYTES);
                        const uint64_t whi = make_desc_sw128(wb), wlo = make_desc_sw128(wb + tc::W_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 16 bf16 = 32 B along K inside the 128 B swizzle row
                            umma_bf16(d_tmem, alo + adv, whi + adv, tc::IDESC, first ? 0u : 1u);
                            first = 0;
                            umma_bf16(d_tmem, ahi + adv, wlo + adv, tc::IDESC, 1u);
                            umma_bf16(d_tmem, ahi + adv, whi + adv, tc::IDESC, 1u);
                        }
                        umma_commit(empty_bar(stage));
                        if (!WIDE) umma_commit(wempty_bar(ws));
                    }
                    __syncwarp();
                    if (++stage == NSB) { stage = 0; phase ^= 1; }
                    if (!WIDE && ++ws == NW) { ws = 0; wph ^= 1; }
                }
            }
            if (leader) umma_commit(tfull_bar(acc));
            if (lane == 0) TC_STAMP(it, 10);
            __syncwarp();
        }
        }
    } else {
        // ================================ epilogue (warps 0-3; MODE 2: warps 0-7) ================================
        // tcgen05.ld hands thread r of a warp accumulator row r (32x32b).  Touching global memory in that shape means 32
        // different lines per instruction, so each half (32 channels) of the warp's 32 rows is staged through shared memory
        // (32 rows x 128 B, 16 B chunks XOR-swizzled by row) and read back with lane l = channels 4*(l&7).. of row 4i+(l>>3):
        // every global load / store then covers four full 128 B lines.  The saved pre-activations of the BN-backward
        // epilogue are fetched before the accumulator is waited for.  BatchNorm sums stay per thread (4 channels per half,
        // fixed row order) over all the CTA's tiles and are reduced across lanes once at the end.
        // MODE 0: a warp does both channel halves of its 32 rows; MODE 2: warp w does half w >> 2 of the rows of quarter w & 3.
        if (EPI != EPI_PLAIN && !WIDE) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;" ::: "memory");
        const int ew = warp & 3, eh = warp >> 2, etid = tid & 127;
        unsigned char* stg = WIDE && warp >= 4 ? smem + W_OFF + tc::WSLOT_BYTES + (warp - 4) * 4096   // (weight slots 1, 2 are unused)
                                               : smem + STG_OFF + warp * 4096;
        const int cq = lane & 7, rsub = lane >> 3;
        float st1[2][4], st2[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { st1[0][i] = 0.f; st2[0][i] = 0.f; st1[1][i] = 0.f; st2[1][i] = 0.f; }
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            TileInfo t;
            const int buf = it & 1;
            int mypix = -1;   // flat output pixel of this thread's accumulator row (-1: padding row of the tile)
            if (MODE != 0) {   // 8x16-pixel tiles (see the producers)
                const int tt = tile % 98, oy = (tt / 7) * 8 + (etid >> 4), ox = (tt % 7) * 16 + (etid & 15);
                if (oy < OH && ox < OW) mypix = ((tile / 98) * OH + oy) * OW + ox;
            } else {
                decode_tile<TRANSPOSED>(g, OH, OW, tile, t);
                const long long m = (long long)t.tile_in_cls * 128 + tid;
                if (m < t.Mc) {
                    const int oxc = (int)(m % t.OWc);
                    const long long q = m / t.OWc;
                    const int oyc = (int)(q % t.OHc);
                    const int n = (int)(q / t.OHc);
                    const int oy = oyc * (TRANSPOSED ? s : 1) + t.py, ox = oxc * (TRANSPOSED ? s : 1) + t.px;
                    mypix = (n * OH + oy) * OW + ox;
                }
            }
            int rowpix[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) rowpix[i] = __shfl_sync(0xffffffffu, mypix, 4 * i + rsub);
            float4 yp[8];   // MODE 2: the pre-activations of this warp's channel half (LDGSTS-prefetched one tile ahead)
            if (YP_SMEM) {
                // every thread copies exactly the chunks it later reads itself: no barrier, only the cp.async group wait
                const uint32_t ypb = base + YP_OFF + (uint32_t)(ew * 32) * 256u + (uint32_t)eh * 128u + (uint32_t)cq * 16u;
                auto issue = [&](const int (&rp)[8]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        cp_async16(ypb + (4 * i + rsub) * 256, a.e_ypre + (rp[i] >= 0 ? (size_t)rp[i] * SRLZ_C + eh * 32 + cq * 4 : 0), rp[i] >= 0);
                    cp_async_commit();
                };
                if (it == 0) issue(rowpix);
                cp_async_wait_all();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    yp[i] = *reinterpret_cast<const float4*>(smem + YP_OFF + (ew * 32 + 4 * i + rsub) * 256 + eh * 128 + cq * 16);
                const int ntile = tile + gridDim.x;
                if (ntile < total_tiles) {
                    const int tt = ntile % 98, oy = (tt / 7) * 8 + (etid >> 4), ox = (tt % 7) * 16 + (etid & 15);
                    const int npix = (oy < OH && ox < OW) ? ((ntile / 98) * OH + oy) * OW + ox : -1;
                    int nrow[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) nrow[i] = __shfl_sync(0xffffffffu, npix, 4 * i + rsub);
                    issue(nrow);
                }
            }
            if (tid == 0) TC_STAMP(it, 11);
            mbar_wait(tfull_bar(buf), (it >> 1) & 1);
            tc_fence_after();
            if (tid == 0) TC_STAMP(it, 12);
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 64;
            // one channel half: h = half index (runtime in MODE 2), HS = its slot in the per-thread statistics (compile time)
            auto do_half = [&](const int h, auto hs_c, const bool last) {
                constexpr int HS = decltype(hs_c)::value;
                if (EPI == EPI_MASK_BNBWD && !YP_SMEM) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        yp[i] = rowpix[i] >= 0 ? ldg4(a.e_ypre + (size_t)rowpix[i] * SRLZ_C + h * 32 + cq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float v[32];
                tmem_ld32(taddr + h * 32, v);
                if (last) {  // this warp's part of the accumulator is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(buf));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                const int ch0 = h * 32 + cq * 4;
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
                    const float4 d4 = *reinterpret_cast<const float4*>(stg + row * 128 + ((cq ^ (row & 7)) << 4));
                    const bool valid = rowpix[i] >= 0;
                    float d[4] = {d4.x, d4.y, d4.z, d4.w};
                    if (EPI == EPI_MASK_BNBWD) {
                        const float ypv[4] = {yp[i].x, yp[i].y, yp[i].z, yp[i].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool on = valid && fmaf(ypv[e], sc[e], sh[e]) > 0.f;
                            const float dz = on ? d[e] : 0.f;
                            d[e] = dz;
                            st1[HS][e] += dz;
                            st2[HS][e] = fmaf(dz, (ypv[e] - me[e]) * iv[e], st2[HS][e]);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float y = valid ? d[e] + sc[e] : 0.f;
                            d[e] = y;
                            if (EPI == EPI_STATS) {
                                st1[HS][e] += y;
                                st2[HS][e] = fmaf(y, y, st2[HS][e]);
                            }
                        }
                    }
                    if (valid) st4(a.out + (size_t)rowpix[i] * SRLZ_C + ch0, make_float4(d[0], d[1], d[2], d[3]));
                }
                __syncwarp();   // the staging rows are rewritten by the next half / tile
            };
            if (WIDE) {
                do_half(eh, IntC<0>{}, true);
            } else {
                do_half(0, IntC<0>{}, false);
                do_half(1, IntC<1>{}, true);
            }
            if (tid == 0) TC_STAMP(it, 13);
        }
        if (EPI != EPI_PLAIN) {
            // lanes with equal (lane & 7) hold the same channels for different rows: fold the 4 row groups in a fixed order
#pragma unroll
            for (int hs = 0; hs < (WIDE ? 1 : 2); ++hs)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    st1[hs][e] += __shfl_xor_sync(0xffffffffu, st1[hs][e], 8);
                    st2[hs][e] += __shfl_xor_sync(0xffffffffu, st2[hs][e], 8);
                    st1[hs][e] += __shfl_xor_sync(0xffffffffu, st1[hs][e], 16);
                    st2[hs][e] += __shfl_xor_sync(0xffffffffu, st2[hs][e], 16);
                }
            if (lane < 8) {   // s_red [4 lane quarters][sum 64 | sum-sq 64]
#pragma unroll
                for (int hs = 0; hs < (WIDE ? 1 : 2); ++hs) {
                    const int h = WIDE ? eh : hs;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s_red[ew * 128 + h * 32 + cq * 4 + e] = st1[hs][e];
                        s_red[ew * 128 + 64 + h * 32 + cq * 4 + e] = st2[hs][e];
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
    if (warp == 4) tmem_dealloc(tmem_base, 128);
    if (EPI != EPI_PLAIN && a.tail.counter != nullptr) bn_tail_run(a.tail, a.partials, reinterpret_cast<double*>(smem), tid);   // (stage buffers are free)
}

template <bool T, bool BN, int EPI, int MODE = 0>
static int launch_tc(const GConvArgs& a, const unsigned char* wbf, int total_tiles, int gx, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gconv64_tc_kernel<T, BN, EPI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gconv64_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    launch_k(gconv64_tc_kernel<T, BN, EPI, MODE>, gx, tc::THREADS, tc::SMEM_BYTES, st, a, wbf, total_tiles);
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
    if (a.mode != 0) total = g.B * 98;   // 14 x 7 tiles of 8 x 16 pixels per image (112x112 / 111x111 grids)
    int gx = sm_count();
    if (gx > total) gx = total;
    if (n_partials) *n_partials = gx;
    if (a.epi != EPI_PLAIN && a.partials == nullptr) { set_error("gconv64_tc: partials buffer required"); return 1; }
    const bool bn = a.in_scale != nullptr;
    const unsigned char* w = reinterpret_cast<const unsigned char*>(wbf);
    if (a.mode != 0 && a.mode != 2) { set_error("gconv64_tc: unknown mode %d", a.mode); return 1; }
    if (a.mode == 2) return launch_tc<false, false, EPI_MASK_BNBWD, 2>(a, w, total, gx, st);
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
    pdl_enter();
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

// W12[ci][co][ky][kx] -> fp32 pack [0][j 0..63][ci]  (j = co*16+ky*4+kx < 48, zero padded)
__global__ void pack_dec12_dgrad_kernel(const float* __restrict__ w12, float* __restrict__ pack1) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // j*64 + ci
    if (idx >= 4096) return;
    const int j = idx >> 6, ci = idx & 63;
    pack1[idx] = j < 48 ? w12[ci * 48 + j] : 0.f;
}
int pack_dec12_dgrad(const float* w12, float* pack1, cudaStream_t st) {
    launch_k(pack_dec12_dgrad_kernel, 16, 256, 0, st, w12, pack1);
    return check_launch("pack_dec12_dgrad");
}

// All six 3x3 64->64 layers in ONE launch, straight from the torch-native weights (W[a][b][tap]: Conv2d a = co, b = ci;
// ConvTranspose2d a = ci, b = co) to both bf16 hi/lo images of each layer: forward image rows n = co, K = ci; dgrad image rows
// n = ci, K = co (the same two layouts pack_conv_w + pack_conv_w_bf16 produce through the fp32 staging packs)
__global__ void pack_conv_layers_bf16_kernel(ConvPackJobs jobs) {
    pdl_enter();
    const int l = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // a*576 + b*9 + tap
    if (idx >= 9 * 4096) return;
    const int tap = idx % 9, b = (idx / 9) & 63, a = idx / (9 * 64);
    const float x = jobs.w[l][idx];
    const int ci = jobs.transposed[l] ? a : b, co = jobs.transposed[l] ? b : a;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    {   // forward image: n = co, k = ci
        const int byte = co * 128 + (((ci >> 3) ^ (co & 7)) << 4) + (ci & 7) * 2;
        unsigned char* t = jobs.fwd[l] + (size_t)tap * (2 * tc::W_BYTES);
        *reinterpret_cast<__nv_bfloat16*>(t + byte) = hi;
        *reinterpret_cast<__nv_bfloat16*>(t + tc::W_BYTES + byte) = lo;
    }
    {   // dgrad image: n = ci, k = co
        const int byte = ci * 128 + (((co >> 3) ^ (ci & 7)) << 4) + (co & 7) * 2;
        unsigned char* t = jobs.dgrad[l] + (size_t)tap * (2 * tc::W_BYTES);
        *reinterpret_cast<__nv_bfloat16*>(t + byte) = hi;
        *reinterpret_cast<__nv_bfloat16*>(t + tc::W_BYTES + byte) = lo;
    }
}

int pack_conv_layers_bf16(const ConvPackJobs& jobs, cudaStream_t st) {
    launch_k(pack_conv_layers_bf16_kernel, dim3((9 * 4096 + 255) / 256, 6), 256, 0, st, jobs);
    return check_launch("pack_conv_layers_bf16");
}

int pack_conv_w_bf16(const float* pack_f32, void* dst, int ntaps, cudaStream_t st) {
    launch_k(pack_conv_w_bf16_kernel, (ntaps * 4096 + 255) / 256, 256, 0, st, pack_f32, reinterpret_cast<unsigned char*>(dst), ntaps);
    return check_launch("pack_conv_w_bf16");
}

}  // namespace srlz

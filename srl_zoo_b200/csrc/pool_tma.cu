// TMA-fed BatchNorm + ReLU + MaxPool(3, s2) forward (models/models.py:50-52,55-57,60-62).
//
// The register-fed version of this pass (bn_pool.cu) was issue / latency bound at ~0.54 of the HBM roof: every thread fetched
// its nine window taps with nine address computations and at most three loads in flight.  Here the pre-activation tensor y
// (B, H, W, 64) fp32 NHWC is described by a 3-D tensor map (64 channels, W pixels, B*H rows) and a CTA streams the INPUT ROWS of
// its contiguous range of pooled rows through a shared-memory ring with `cp.async.bulk.tensor` (one elected thread issues, an
// mbarrier per slot counts the bytes): every input row crosses HBM once per CTA (plus one shared row per range boundary), many
// rows are in flight per SM without holding a single register, and the window taps become shared-memory loads at fixed offsets.
//
// Exact torch semantics (max_pool2d over relu(bn(y)), first maximal tap in scan order): the running maximum is taken over the
// BatchNorm outputs v = y*scale + shift with a strict '>', which visits the taps in the same order and breaks ties the same way as
// the maximum over relu(v) whenever the maximum is positive (relu leaves positive values untouched and maps everything else below
// it); when no tap is positive every relu(v) is 0, the first valid tap wins and the output is 0 -- handled after the scan.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace srlz {

namespace pt {
constexpr int CONSUMER_WARPS = 16;                      // 512 threads = 32 pooled columns x 16 channel quads per pass (14 warps = two full passes of 28 measured slower)
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;      // warp 0: TMA producer | warps 1-28: window scan
constexpr int NPW = CONSUMER_WARPS * 2;                 // pooled columns per pass
constexpr int RING_BYTES = 200 * 1024;
constexpr int MAX_SLOTS = 16;
constexpr int OFF_BARS = RING_BYTES;                    // full[16] empty[16]
constexpr int SMEM_BYTES = RING_BYTES + 512 + 128;      // + 128 B alignment slack
}  // namespace pt

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__global__ void __launch_bounds__(pt::THREADS, 1) bn_relu_pool_fwd_tma_kernel(const __grid_constant__ CUtensorMap ymap, const float* __restrict__ scale,
                                                                               const float* __restrict__ shift, float* __restrict__ out,
                                                                               unsigned char* __restrict__ argmax, int H, int W, int PH, int PW,
                                                                               int pad, int total_rows, int nslot) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t bars = base + pt::OFF_BARS;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (pt::MAX_SLOTS + s); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t row_bytes = (uint32_t)W * 256u;
    const int i0 = (int)((long long)total_rows * blockIdx.x / gridDim.x), i1 = (int)((long long)total_rows * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int s = 0; s < nslot; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), pt::CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();

    // the input-row sequence of this CTA: for every image segment [pa, pb] of its pooled-row range, rows r_lo .. r_hi with
    // r_lo = max(0, 2pa - pad), r_hi = min(H - 1, 2pb - pad + 2); sequence index g -> slot g % nslot, phase (g / nslot) & 1
    if (warp == 0) {
        if (elect_one()) {
            int g = 0;
            for (int i = i0; i < i1;) {
                const int n = i / PH, pa = i - n * PH;
                int pb = pa + (i1 - i) - 1;
                if (pb > PH - 1) pb = PH - 1;
                const int r_lo = max(0, 2 * pa - pad), r_hi = min(H - 1, 2 * pb - pad + 2);
                for (int r = r_lo; r <= r_hi; ++r, ++g) {
                    const int slot = g % nslot, ph = (g / nslot) & 1;
                    mbar_wait(empty_bar(slot), ph ^ 1);
                    mbar_arrive_expect_tx(full_bar(slot), row_bytes);
                    tma_load_3d(base + slot * row_bytes, &ymap, 0, 0, n * H + r, full_bar(slot));
                }
                i += pb - pa + 1;
            }
        }
    } else {
        // ================================ window scan: thread = (pooled column, 4 channels) ================================
        const int ct = tid - 32, c4 = ct & 15, pw0 = ct >> 4;                    // NPW pooled columns per pass
        const float4 sc = ldg4(scale + c4 * 4), sh = ldg4(shift + c4 * 4);
        int gbase = 0;                                                            // sequence index of the segment's first row
        for (int i = i0; i < i1;) {
            const int n = i / PH, pa = i - n * PH;
            int pb = pa + (i1 - i) - 1;
            if (pb > PH - 1) pb = PH - 1;
            const int r_lo = max(0, 2 * pa - pad), r_hi = min(H - 1, 2 * pb - pad + 2);
            int released = r_lo;                                                  // rows below `released` were handed back
            for (int ph = pa; ph <= pb; ++ph) {
                const int h0 = 2 * ph - pad;
                const int ha = max(h0, 0), hb = min(h0 + 2, H - 1);
                for (int r = ha; r <= hb; ++r) {                                  // rows of this window (already-waited rows return at once)
                    const int g = gbase + (r - r_lo);
                    mbar_wait(full_bar(g % nslot), (g / nslot) & 1);
                }
                const size_t orow = ((size_t)n * PH + ph) * PW;
                // the three row images of this window (hoisted out of the column loop: the slot index is a runtime modulo)
                const float* rb0 = reinterpret_cast<const float*>(smem + ((gbase + (max(h0, 0) - r_lo)) % nslot) * row_bytes) + c4 * 4;
                const float* rb1 = reinterpret_cast<const float*>(smem + ((gbase + (h0 + 1 - r_lo)) % nslot) * row_bytes) + c4 * 4;
                const float* rb2 = reinterpret_cast<const float*>(smem + ((gbase + (h0 + 2 - r_lo)) % nslot) * row_bytes) + c4 * 4;
                for (int pw = pw0; pw < PW; pw += pt::NPW) {
                    const int w0 = pw * 2 - pad;
                    const size_t po = (orow + pw) * 64 + c4 * 4;
                    float b0, b1, b2, b3;
                    int a0, a1, a2, a3, first = 0;
                    if (h0 >= 0 && w0 >= 0) {
                        // interior window (all nine taps inside; the bottom / right edges never clip with floor pooling): no
                        // predicates, taps at fixed shared-memory offsets -- 4 instructions per channel and tap
                        const float* r0 = rb0 + w0 * 64;
                        const float* r1 = rb1 + w0 * 64;
                        const float* r2 = rb2 + w0 * 64;
                        float4 y4 = *reinterpret_cast<const float4*>(r0);
                        b0 = fmaf(y4.x, sc.x, sh.x); b1 = fmaf(y4.y, sc.y, sh.y); b2 = fmaf(y4.z, sc.z, sh.z); b3 = fmaf(y4.w, sc.w, sh.w);
                        a0 = a1 = a2 = a3 = 0;
#define PT_TAP(ptr, t)                                                                                     \
    do {                                                                                                   \
        y4 = *reinterpret_cast<const float4*>(ptr);                                                        \
        const float v0 = fmaf(y4.x, sc.x, sh.x), v1 = fmaf(y4.y, sc.y, sh.y), v2 = fmaf(y4.z, sc.z, sh.z), \
                    v3 = fmaf(y4.w, sc.w, sh.w);                                                           \
        if (v0 > b0) { b0 = v0; a0 = (t); }                                                                \
        if (v1 > b1) { b1 = v1; a1 = (t); }                                                                \
        if (v2 > b2) { b2 = v2; a2 = (t); }                                                                \
        if (v3 > b3) { b3 = v3; a3 = (t); }                                                                \
    } while (0)
                        PT_TAP(r0 + 64, 1); PT_TAP(r0 + 128, 2);
                        PT_TAP(r1, 3); PT_TAP(r1 + 64, 4); PT_TAP(r1 + 128, 5);
                        PT_TAP(r2, 6); PT_TAP(r2 + 64, 7); PT_TAP(r2 + 128, 8);
                    } else {
                        // top / left border (pad = 1, first pooled row or column): taps outside the image are skipped (-inf padding)
                        b0 = b1 = b2 = b3 = -INFINITY;
                        a0 = a1 = a2 = a3 = 255;
                        first = -1;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const int h = h0 + ky;
                            if (h < 0) continue;
                            const float* yr = ky == 0 ? rb0 : (ky == 1 ? rb1 : rb2);
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                if (w0 + kx < 0) continue;
                                const int t = ky * 3 + kx;
                                if (first < 0) first = t;
                                float4 y4;
                                PT_TAP(yr + (size_t)(w0 + kx) * 64, t);
                            }
                        }
                    }
#undef PT_TAP
                    // no positive tap: every relu value is 0 and torch keeps the first tap of the window (NaN inputs aside)
                    if (!(b0 > 0.f)) { b0 = 0.f; a0 = first; }
                    if (!(b1 > 0.f)) { b1 = 0.f; a1 = first; }
                    if (!(b2 > 0.f)) { b2 = 0.f; a2 = first; }
                    if (!(b3 > 0.f)) { b3 = 0.f; a3 = first; }
                    st4(out + po, make_float4(b0, b1, b2, b3));
                    if (argmax != nullptr) *reinterpret_cast<uint32_t*>(argmax + po) = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16) | ((uint32_t)a3 << 24);
                }
                // rows below the next window's first row are done: hand their slots back (one arrival per consumer warp)
                const int keep = ph < pb ? max(2 * (ph + 1) - pad, 0) : r_hi + 1;
                __syncwarp();
                for (; released < keep; ++released) {
                    if (lane == 0) mbar_arrive(empty_bar((gbase + (released - r_lo)) % nslot));
                }
            }
            gbase += r_hi - r_lo + 1;
            i += pb - pa + 1;
        }
    }
}

// cuTensorMapEncodeTiled is a driver-API entry point: it is looked up through the runtime at first use, so that libsrlz.so itself
// only depends on libcudart (and loads on a machine without a driver, where nothing is launched anyway)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int encode_rows_map(CUtensorMap* map, const float* y, int B, int H, int W) {
    // 3-D view of the NHWC tensor: (64 channels, W pixels, B*H rows); one box = one whole row (64 x W x 1)
    const cuuint64_t gdim[3] = {64, (cuuint64_t)W, (cuuint64_t)B * H};
    const cuuint64_t gstride[2] = {256, (cuuint64_t)W * 256};      // bytes, dims 1 and 2
    const cuuint32_t box[3] = {64, (cuuint32_t)W, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    EncodeTiledFn encode = encode_tiled_fn();
    if (encode == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return 1002; }
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(y), gdim, gstride, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return 1002; }
    return 0;
}

bool bn_relu_pool_fwd_tma_supported(const float* y, int W) { return W <= 256 && (reinterpret_cast<uintptr_t>(y) & 15) == 0; }

int bn_relu_pool_fwd_tma(const float* y, const float* scale, const float* shift, float* out, unsigned char* argmax, int B, int H, int W,
                         int PH, int PW, int pad, cudaStream_t st) {
    CUtensorMap map;
    int rc = encode_rows_map(&map, y, B, H, W);
    if (rc) return rc;
    int nslot = pt::RING_BYTES / (W * 256);
    if (nslot > pt::MAX_SLOTS) nslot = pt::MAX_SLOTS;
    if (nslot < 4) { set_error("bn_relu_pool_fwd_tma: row too wide for the ring"); return 1; }
    const int total = B * PH;
    int gx = sm_count();
    if (gx > total) gx = total;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(bn_relu_pool_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pt::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("bn_relu_pool_fwd_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1002; }
        configured = true;
    }
    launch_k(bn_relu_pool_fwd_tma_kernel, gx, pt::THREADS, pt::SMEM_BYTES, st, map, scale, shift, out, argmax, H, W, PH, PW, pad, total, nslot);
    return check_launch("bn_relu_pool_fwd_tma");
}

}  // namespace srlz

// First encoder layer: Conv2d(3, 64, k7, s2, p3, bias=False) (models/models.py:49) on the NCHW fp32
// observation, writing the pre-BN activation in NHWC, with BatchNorm statistics in the epilogue and
// the DAE zero-pixel rectangle (preprocessing/data_loader.py:55-63) applied while staging the input.
#include "common.cuh"
#include "kernels.h"

namespace srlz {

#define E0_IMG 224
#define E0_OUT 112
#define E0_TH 8
#define E0_TW 16
#define E0_PR 21   // patch rows  = 2*(TH-1)+7
#define E0_PC 37   // patch cols  = 2*(TW-1)+7
#define E0_PS 40   // padded patch row stride
#define E0_K 147

__device__ __forceinline__ void e0_load_patch(float* patch, const float* __restrict__ x, const int* __restrict__ rects,
                                              int n, int oy0, int ox0, int tid) {
    int h1 = 0, h2 = 0, w1 = 0, w2 = 0;
    if (rects != nullptr) {
        h1 = rects[n * 4 + 0]; h2 = rects[n * 4 + 1]; w1 = rects[n * 4 + 2]; w2 = rects[n * 4 + 3];
    }
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int e = tid; e < 3 * E0_PR * E0_PC; e += 256) {
        const int c = e % E0_PC;
        const int t = e / E0_PC;
        const int r = t % E0_PR;
        const int ci = t / E0_PR;
        const int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;
        if (iy >= 0 && iy < E0_IMG && ix >= 0 && ix < E0_IMG) {
            // tensor (C, dim2, dim3): the occluded block is [:, w1:w2, h1:h2]
            const bool masked = (iy >= w1) && (iy < w2) && (ix >= h1) && (ix < h2);
            if (!masked) v = __ldg(x + (((size_t)n * 3 + ci) * E0_IMG + iy) * E0_IMG + ix);
        }
        patch[(ci * E0_PR + r) * E0_PS + c] = v;
    }
}

__global__ void __launch_bounds__(256, 2) enc0_fwd_kernel(Enc0Args a, int ntiles, int want_stats) {
    extern __shared__ __align__(16) float smem[];
    float* Wsm = smem;                       // [147][64]
    float* patch = smem + E0_K * 64;         // [3][21][40]
    float* s_red = patch + 3 * E0_PR * E0_PS;  // [8][128]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < E0_K * 64 / 4; i += 256) st4(Wsm + i * 4, ldg4(a.wpack + i * 4));
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
    const int tiles_per_img = (E0_OUT / E0_TH) * (E0_OUT / E0_TW);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int tt = tile % tiles_per_img;
        const int oy0 = (tt / (E0_OUT / E0_TW)) * E0_TH, ox0 = (tt % (E0_OUT / E0_TW)) * E0_TW;
        __syncthreads();
        e0_load_patch(patch, a.x, a.rects, n, oy0, ox0, tid);
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        // thread: output column ty of the tile, rows 0..7, channels tx*4..+3
        for (int ci = 0; ci < 3; ++ci) {
            for (int ky = 0; ky < 7; ++ky) {
                const float* prow = patch + (ci * E0_PR + ky) * E0_PS + 2 * ty;
                const float* wrow = Wsm + ((ci * 7 + ky) * 7) * 64 + tx * 4;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float4 w = *reinterpret_cast<const float4*>(wrow + kx * 64);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float xv = prow[(2 * i) * E0_PS + kx];
                        acc[i][0] = fmaf(xv, w.x, acc[i][0]);
                        acc[i][1] = fmaf(xv, w.y, acc[i][1]);
                        acc[i][2] = fmaf(xv, w.z, acc[i][2]);
                        acc[i][3] = fmaf(xv, w.w, acc[i][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const size_t off = (((size_t)n * E0_OUT + oy0 + i) * E0_OUT + ox0 + ty) * 64 + tx * 4;
            st4(a.y + off, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                st1[j] += acc[i][j];
                st2[j] = fmaf(acc[i][j], acc[i][j], st2[j]);
            }
        }
    }
    if (want_stats) {
        const int w = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 16);
            st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 16);
        }
        __syncthreads();
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_red[w * 128 + lane * 4 + j] = st1[j];
                s_red[w * 128 + 64 + lane * 4 + j] = st2[j];
            }
        }
        __syncthreads();
        if (tid < 128) {
            float v = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) v += s_red[ww * 128 + tid];
            a.partials[(size_t)blockIdx.x * 128 + tid] = v;
        }
    }
}

int enc0_fwd(const Enc0Args& a, int* n_partials, cudaStream_t st) {
    const int ntiles = a.B * (E0_OUT / E0_TH) * (E0_OUT / E0_TW);
    const int smem = (E0_K * 64 + 3 * E0_PR * E0_PS + 8 * 128) * (int)sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(enc0_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    int gx = 2 * sm_count() * 2;
    if (gx > ntiles) gx = ntiles;
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    if (n_partials) *n_partials = gx;
    enc0_fwd_kernel<<<gx, 256, smem, st>>>(a, ntiles, a.partials != nullptr ? 1 : 0);
    return check_launch("enc0_fwd");
}

// wgrad: dW[co][k] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, ci, 2oy-3+ky, 2ox-3+kx]
// thread: 4 output channels (tx) x 10 taps k = kg + 16*j ; accumulators persist over the CTA's tiles.
__global__ void __launch_bounds__(256, 2) enc0_wgrad_kernel(Enc0WgradArgs a, int ntiles) {
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                          // [3][21][40]
    float* dys = smem + 3 * E0_PR * E0_PS;        // [128][64]
    const int tid = threadIdx.x, tx = tid & 15, kg = tid >> 4;
    int koff[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        int k = kg + 16 * j;
        if (k >= E0_K) k = 0;  // dummy (result discarded)
        const int ci = k / 49, ky = (k / 7) % 7, kx = k % 7;
        koff[j] = (ci * E0_PR + ky) * E0_PS + kx;
    }
    float acc[10][4];
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
    const int tiles_per_img = (E0_OUT / E0_TH) * (E0_OUT / E0_TW);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int tt = tile % tiles_per_img;
        const int oy0 = (tt / (E0_OUT / E0_TW)) * E0_TH, ox0 = (tt % (E0_OUT / E0_TW)) * E0_TW;
        __syncthreads();
        e0_load_patch(patch, a.x, a.rects, n, oy0, ox0, tid);
        for (int e = tid; e < 128 * 16; e += 256) {
            const int p = e >> 4, c4 = e & 15;
            const size_t off = (((size_t)n * E0_OUT + oy0 + (p >> 4)) * E0_OUT + ox0 + (p & 15)) * 64 + c4 * 4;
            st4(dys + p * 64 + c4 * 4, ldg4(a.dy + off));
        }
        __syncthreads();
#pragma unroll 2
        for (int p = 0; p < 128; ++p) {
            const float4 d = *reinterpret_cast<const float4*>(dys + p * 64 + tx * 4);
            const float* pb = patch + (2 * (p >> 4)) * E0_PS + 2 * (p & 15);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const float xv = pb[koff[j]];
                acc[j][0] = fmaf(xv, d.x, acc[j][0]);
                acc[j][1] = fmaf(xv, d.y, acc[j][1]);
                acc[j][2] = fmaf(xv, d.z, acc[j][2]);
                acc[j][3] = fmaf(xv, d.w, acc[j][3]);
            }
        }
    }
    float* dst = a.partials + (size_t)blockIdx.x * (E0_K * 64);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const int k = kg + 16 * j;
        if (k < E0_K) st4(dst + k * 64 + tx * 4, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
    }
}

// grad[co*147 + k] (+)= sum_cta partials[cta][k*64 + co]
__global__ void enc0_wgrad_reduce_kernel(const float* __restrict__ partials, float* __restrict__ grad, int ncta,
                                         int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= E0_K * 64) return;
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += partials[(size_t)c * (E0_K * 64) + idx];
    const int k = idx >> 6, co = idx & 63;
    const int o = co * E0_K + k;
    grad[o] = accumulate ? grad[o] + s : s;
}

static int enc0_wgrad_ctas(int ntiles) {
    int gx = 2 * sm_count();
    if (gx > ntiles) gx = ntiles;
    return gx;
}

size_t enc0_wgrad_partial_floats() { return (size_t)(2 * sm_count()) * E0_K * 64; }

int enc0_wgrad(const Enc0WgradArgs& a, cudaStream_t st) {
    const int ntiles = a.B * (E0_OUT / E0_TH) * (E0_OUT / E0_TW);
    const int smem = (3 * E0_PR * E0_PS + 128 * 64) * (int)sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(enc0_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    const int gx = enc0_wgrad_ctas(ntiles);
    enc0_wgrad_kernel<<<gx, 256, smem, st>>>(a, ntiles);
    int rc = check_launch("enc0_wgrad");
    if (rc) return rc;
    enc0_wgrad_reduce_kernel<<<(E0_K * 64 + 255) / 256, 256, 0, st>>>(a.partials, a.grad, gx, a.accumulate);
    return check_launch("enc0_wgrad_reduce");
}

}  // namespace srlz

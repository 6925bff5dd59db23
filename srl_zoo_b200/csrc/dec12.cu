// Last decoder layer: ConvTranspose2d(64, 3, k4, s2) + bias (models/models.py:82), 111x111x64 (NHWC, pre-BN of
// the previous layer; BN+ReLU applied while staging) -> (B,3,224,224) NCHW, with the pixel-wise squared error
// against the target (losses/losses.py:172-181,199-214) reduced in the epilogue.  Backward: dgrad (+ReLU mask
// and BatchNorm-backward statistics of the previous layer) and wgrad/bias-grad.
#include "common.cuh"
#include "kernels.h"

namespace srlz {

#define D12_IN 111
#define D12_OUT 224
#define D12_T 16        // tile edge in input-pixel units
#define D12_PS 20       // floats per staged pixel (16 channels + 4 pad)

// ------------------------------ forward ------------------------------
// thread (ti,tj) owns the 2x2 output block (2i+py, 2j+px), i,j in [0,112), fed by inputs (i-dy, j-dx).
__global__ void __launch_bounds__(256, 2) dec12_fwd_kernel(Dec12FwdArgs a, int ntiles) {
    __shared__ __align__(16) float Wsm[64 * 4 * 12];             // [ci][dy*2+dx][(py*2+px)*3+co]
    __shared__ __align__(16) float patch[17 * 17 * D12_PS];      // 16-channel chunk of the 17x17 input patch
    __shared__ float s_red[8];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    for (int e = tid; e < 64 * 48; e += 256) {
        const int ci = e / 48, r = e % 48;
        const int d = r / 12, q = r % 12;
        const int dy = d >> 1, dx = d & 1, pyx = q / 3, co = q % 3;
        const int ky = (pyx >> 1) + 2 * dy, kx = (pyx & 1) + 2 * dx;
        Wsm[e] = __ldg(a.w + ((ci * 3 + co) * 4 + ky) * 4 + kx);
    }
    const float b0 = __ldg(a.bias + 0), b1 = __ldg(a.bias + 1), b2 = __ldg(a.bias + 2);
    float sse = 0.f;
    const int tiles_per_img = 7 * 7;  // 112 / 16
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int tt = tile % tiles_per_img;
        const int i0 = (tt / 7) * D12_T, j0 = (tt % 7) * D12_T;
        float acc[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[q][0] = b0; acc[q][1] = b1; acc[q][2] = b2; }
        for (int chunk = 0; chunk < 4; ++chunk) {
            __syncthreads();
            for (int e = tid; e < 17 * 17 * 4; e += 256) {
                const int pix = e >> 2, c4 = e & 3;
                const int iy = i0 - 1 + pix / 17, ix = j0 - 1 + pix % 17;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (iy >= 0 && iy < D12_IN && ix >= 0 && ix < D12_IN) {
                    const int c = chunk * 16 + c4 * 4;
                    v = ldg4(a.ypre + (((size_t)n * D12_IN + iy) * D12_IN + ix) * 64 + c);
                    v = bn_relu4(v, ldg4(a.scale + c), ldg4(a.shift + c));
                }
                st4(patch + pix * D12_PS + c4 * 4, v);
            }
            __syncthreads();
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int dy = d >> 1, dx = d & 1;
                const float* pp = patch + ((ti + 1 - dy) * 17 + (tj + 1 - dx)) * D12_PS;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 av = *reinterpret_cast<const float4*>(pp + c4 * 4);
                    const float avs[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float* wp = Wsm + ((chunk * 16 + c4 * 4 + cc) * 4 + d) * 12;
                        const float4 w0 = *reinterpret_cast<const float4*>(wp);
                        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
                        const float4 w2 = *reinterpret_cast<const float4*>(wp + 8);
                        const float x = avs[cc];
                        acc[0][0] = fmaf(x, w0.x, acc[0][0]); acc[0][1] = fmaf(x, w0.y, acc[0][1]); acc[0][2] = fmaf(x, w0.z, acc[0][2]);
                        acc[1][0] = fmaf(x, w0.w, acc[1][0]); acc[1][1] = fmaf(x, w1.x, acc[1][1]); acc[1][2] = fmaf(x, w1.y, acc[1][2]);
                        acc[2][0] = fmaf(x, w1.z, acc[2][0]); acc[2][1] = fmaf(x, w1.w, acc[2][1]); acc[2][2] = fmaf(x, w2.x, acc[2][2]);
                        acc[3][0] = fmaf(x, w2.y, acc[3][0]); acc[3][1] = fmaf(x, w2.z, acc[3][1]); acc[3][2] = fmaf(x, w2.w, acc[3][2]);
                    }
                }
            }
        }
        const int oy = 2 * (i0 + ti), ox = 2 * (j0 + tj);
#pragma unroll
        for (int co = 0; co < 3; ++co) {
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const size_t off = (((size_t)n * 3 + co) * D12_OUT + oy + py) * D12_OUT + ox;
                const float2 v = make_float2(acc[py * 2 + 0][co], acc[py * 2 + 1][co]);
                *reinterpret_cast<float2*>(a.out + off) = v;
                if (a.target != nullptr) {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(a.target + off));
                    const float e0 = v.x - t.x, e1 = v.y - t.y;
                    sse = fmaf(e0, e0, sse);
                    sse = fmaf(e1, e1, sse);
                }
            }
        }
    }
    if (a.sse_partials != nullptr) {
        sse = warp_sum(sse);
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = sse;
        __syncthreads();
        if (tid == 0) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += s_red[w];
            a.sse_partials[blockIdx.x] = v;
        }
    }
}

int dec12_fwd(const Dec12FwdArgs& a, int* n_partials, cudaStream_t st) {
    const int ntiles = a.B * 49;
    int gx = 2 * sm_count() * 2;
    if (gx > ntiles) gx = ntiles;
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    if (n_partials) *n_partials = gx;
    dec12_fwd_kernel<<<gx, 256, 0, st>>>(a, ntiles);
    return check_launch("dec12_fwd");
}

// ------------------------------ backward ------------------------------
#define D12_GS 34   // g patch edge: 2*(16-1)+4

// stage grad-wrt-decoded for input tile (iy0..+16, ix0..+16): rows 2*iy0 + r, cols 2*ix0 + c  (r,c in [0,34)).
// WITH_BIAS: per-thread sums of the 32x32 block this tile owns (each output pixel is owned by exactly one
// tile) are left in gp[3*34*34 + tid*3 + co] for the bias gradient.
template <bool WITH_BIAS>
__device__ __forceinline__ void d12_load_gpatch(float* gp, const Dec12BwdArgs& a, int n, int iy0, int ix0, int tid) {
    float own0 = 0.f, own1 = 0.f, own2 = 0.f;
    for (int e = tid; e < 3 * D12_GS * D12_GS; e += 256) {
        const int c = e % D12_GS;
        const int t = e / D12_GS;
        const int r = t % D12_GS;
        const int co = t / D12_GS;
        const int oy = 2 * iy0 + r, ox = 2 * ix0 + c;
        float v = 0.f;
        if (oy < D12_OUT && ox < D12_OUT) {
            const size_t off = (((size_t)n * 3 + co) * D12_OUT + oy) * D12_OUT + ox;
            v = (a.gout != nullptr) ? __ldg(a.gout + off) : a.coef * (__ldg(a.decoded + off) - __ldg(a.target + off));
        }
        gp[e] = v;
        if (WITH_BIAS && r < 32 && c < 32) {
            if (co == 0) own0 += v;
            else if (co == 1) own1 += v;
            else own2 += v;
        }
    }
    if (WITH_BIAS) {
        gp[3 * D12_GS * D12_GS + tid * 3 + 0] = own0;
        gp[3 * D12_GS * D12_GS + tid * 3 + 1] = own1;
        gp[3 * D12_GS * D12_GS + tid * 3 + 2] = own2;
    }
}

// dgrad + ReLU mask + BN-backward statistics.  thread: 4 consecutive pixels (same row) x 16 channels.
__global__ void __launch_bounds__(256, 2) dec12_dgrad_kernel(Dec12BwdArgs a, int ntiles) {
    __shared__ __align__(16) float Wd[48 * 64];                                   // [(co*4+ky)*4+kx][ci]
    __shared__ __align__(16) float gp[3 * D12_GS * D12_GS + 256 * 3];
    __shared__ float s_stat[8][128];
    __shared__ __align__(16) float s_bn[4][64];
    const int tid = threadIdx.x;
    const int cq = tid & 3, pg = tid >> 2;
    const int ti = pg >> 2, tjg = pg & 3;
    for (int e = tid; e < 48 * 64; e += 256) {
        const int ci = e & 63, t = e >> 6;  // t = (co*4+ky)*4+kx
        Wd[e] = __ldg(a.w + ci * 48 + t);
    }
    for (int e = tid; e < 8 * 128; e += 256) (&s_stat[0][0])[e] = 0.f;
    if (tid < 64) {
        s_bn[0][tid] = __ldg(a.scale + tid);
        s_bn[1][tid] = __ldg(a.shift + tid);
        s_bn[2][tid] = __ldg(a.mean + tid);
        s_bn[3][tid] = __ldg(a.invstd + tid);
    }
    const int tiles_per_img = 7 * 7;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int tt = tile % tiles_per_img;
        const int iy0 = (tt / 7) * D12_T, ix0 = (tt % 7) * D12_T;
        __syncthreads();
        d12_load_gpatch<false>(gp, a, n, iy0, ix0, tid);
        __syncthreads();
        float acc[4][16];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[p][c] = 0.f;
#pragma unroll 1
        for (int co = 0; co < 3; ++co) {
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const float* grow = gp + (co * D12_GS + 2 * ti + ky) * D12_GS + 8 * tjg;
                float gv[10];
#pragma unroll
                for (int h = 0; h < 5; ++h) {
                    const float2 t2 = *reinterpret_cast<const float2*>(grow + 2 * h);
                    gv[2 * h] = t2.x;
                    gv[2 * h + 1] = t2.y;
                }
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const float* wrow = Wd + ((co * 4 + ky) * 4 + kx) * 64 + cq * 16;
                    float wv[16];
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wrow + f * 4);
                        wv[f * 4 + 0] = w4.x; wv[f * 4 + 1] = w4.y; wv[f * 4 + 2] = w4.z; wv[f * 4 + 3] = w4.w;
                    }
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float gval = gv[2 * p + kx];
#pragma unroll
                        for (int c = 0; c < 16; ++c) acc[p][c] = fmaf(gval, wv[c], acc[p][c]);
                    }
                }
            }
        }
        // epilogue: mask by relu(bn(ypre)) > 0, accumulate sum dz and sum dz*xhat, store dz
        float s1[16], s2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
        const int iy = iy0 + ti;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ix = ix0 + tjg * 4 + p;
            if (iy < D12_IN && ix < D12_IN) {
                const size_t off = (((size_t)n * D12_IN + iy) * D12_IN + ix) * 64 + cq * 16;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const float4 yp = ldg4(a.ypre + off + f * 4);
                    const float4 scf = *reinterpret_cast<const float4*>(&s_bn[0][cq * 16 + f * 4]);
                    const float4 shf = *reinterpret_cast<const float4*>(&s_bn[1][cq * 16 + f * 4]);
                    const float4 mef = *reinterpret_cast<const float4*>(&s_bn[2][cq * 16 + f * 4]);
                    const float4 inf = *reinterpret_cast<const float4*>(&s_bn[3][cq * 16 + f * 4]);
                    float4 v = make_float4(acc[p][f * 4 + 0], acc[p][f * 4 + 1], acc[p][f * 4 + 2], acc[p][f * 4 + 3]);
                    v.x = fmaf(yp.x, scf.x, shf.x) > 0.f ? v.x : 0.f;
                    v.y = fmaf(yp.y, scf.y, shf.y) > 0.f ? v.y : 0.f;
                    v.z = fmaf(yp.z, scf.z, shf.z) > 0.f ? v.z : 0.f;
                    v.w = fmaf(yp.w, scf.w, shf.w) > 0.f ? v.w : 0.f;
                    s1[f * 4 + 0] += v.x; s1[f * 4 + 1] += v.y; s1[f * 4 + 2] += v.z; s1[f * 4 + 3] += v.w;
                    s2[f * 4 + 0] = fmaf(v.x, (yp.x - mef.x) * inf.x, s2[f * 4 + 0]);
                    s2[f * 4 + 1] = fmaf(v.y, (yp.y - mef.y) * inf.y, s2[f * 4 + 1]);
                    s2[f * 4 + 2] = fmaf(v.z, (yp.z - mef.z) * inf.z, s2[f * 4 + 2]);
                    s2[f * 4 + 3] = fmaf(v.w, (yp.w - mef.w) * inf.w, s2[f * 4 + 3]);
                    st4(a.dz + off + f * 4, v);
                }
            }
        }
        // lanes with equal (lane & 3) share the channel quarter: reduce over lane bits 2..4
#pragma unroll
        for (int c = 0; c < 16; ++c) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
                s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o);
            }
        }
        const int lane = tid & 31, w = tid >> 5;
        if (lane < 4) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                s_stat[w][lane * 16 + c] += s1[c];
                s_stat[w][64 + lane * 16 + c] += s2[c];
            }
        }
    }
    __syncthreads();
    if (tid < 128) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_stat[w][tid];
        a.stat_partials[(size_t)blockIdx.x * 128 + tid] = v;
    }
}

// wgrad + bias grad.  thread: input channel ci = tid & 63, 12 of the 48 (co,ky,kx) taps.
#define D12_WP (3072 + 4)
__global__ void __launch_bounds__(256, 2) dec12_wgrad_kernel(Dec12BwdArgs a, int ntiles) {
    __shared__ __align__(16) float gp[3 * D12_GS * D12_GS + 256 * 3];
    const int tid = threadIdx.x;
    const int ci = tid & 63, tq = tid >> 6;
    const float sc = __ldg(a.scale + ci), sh = __ldg(a.shift + ci);
    float acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.f;
    float bsum = 0.f;  // threads 0..2: bias gradient of output channel tid
    const int tiles_per_img = 7 * 7;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int tt = tile % tiles_per_img;
        const int iy0 = (tt / 7) * D12_T, ix0 = (tt % 7) * D12_T;
        __syncthreads();
        d12_load_gpatch<true>(gp, a, n, iy0, ix0, tid);
        __syncthreads();
        if (tid < 3) {  // fixed-order reduction of the per-thread bias partials of this tile
            float v = 0.f;
            for (int t = 0; t < 256; ++t) v += gp[3 * D12_GS * D12_GS + t * 3 + tid];
            bsum += v;
        }
        const int nrow = (D12_IN - iy0) < D12_T ? (D12_IN - iy0) : D12_T;
        const int ncol = (D12_IN - ix0) < D12_T ? (D12_IN - ix0) : D12_T;
        for (int r = 0; r < nrow; ++r) {
            const float* abase = a.ypre + (((size_t)n * D12_IN + iy0 + r) * D12_IN + ix0) * 64 + ci;
#pragma unroll 4
            for (int c = 0; c < ncol; ++c) {
                const float av = fmaxf(fmaf(__ldg(abase + (size_t)c * 64), sc, sh), 0.f);
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const int cky = tq * 3 + rr;  // = co*4 + ky
                    const float* grow = gp + ((cky >> 2) * D12_GS + 2 * r + (cky & 3)) * D12_GS + 2 * c;
                    const float2 g01 = *reinterpret_cast<const float2*>(grow);
                    const float2 g23 = *reinterpret_cast<const float2*>(grow + 2);
                    acc[rr * 4 + 0] = fmaf(av, g01.x, acc[rr * 4 + 0]);
                    acc[rr * 4 + 1] = fmaf(av, g01.y, acc[rr * 4 + 1]);
                    acc[rr * 4 + 2] = fmaf(av, g23.x, acc[rr * 4 + 2]);
                    acc[rr * 4 + 3] = fmaf(av, g23.y, acc[rr * 4 + 3]);
                }
            }
        }
    }
    float* dst = a.w_partials + (size_t)blockIdx.x * D12_WP;
#pragma unroll
    for (int j = 0; j < 12; ++j) dst[ci * 48 + tq * 12 + j] = acc[j];
    if (tid < 3) dst[3072 + tid] = bsum;
}

__global__ void dec12_wgrad_reduce_kernel(const float* __restrict__ partials, float* __restrict__ gw,
                                          float* __restrict__ gb, int ncta, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3075) return;
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += partials[(size_t)c * D12_WP + idx];
    float* o = idx < 3072 ? gw + idx : gb + (idx - 3072);
    *o = accumulate ? *o + s : s;
}

size_t dec12_wgrad_partial_floats() { return (size_t)(2 * sm_count() * 2) * D12_WP; }

int dec12_bwd(const Dec12BwdArgs& a, int* n_stat_partials, cudaStream_t st) {
    const int ntiles = a.B * 49;
    int gx = 2 * sm_count() * 2;
    if (gx > ntiles) gx = ntiles;
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    if (n_stat_partials) *n_stat_partials = gx;
    dec12_wgrad_kernel<<<gx, 256, 0, st>>>(a, ntiles);
    int rc = check_launch("dec12_wgrad");
    if (rc) return rc;
    dec12_wgrad_reduce_kernel<<<(3075 + 255) / 256, 256, 0, st>>>(a.w_partials, a.grad_w, a.grad_b, gx, a.accumulate);
    rc = check_launch("dec12_wgrad_reduce");
    if (rc) return rc;
    if (a.skip_dgrad) return 0;
    dec12_dgrad_kernel<<<gx, 256, 0, st>>>(a, ntiles);
    return check_launch("dec12_dgrad");
}

}  // namespace srlz

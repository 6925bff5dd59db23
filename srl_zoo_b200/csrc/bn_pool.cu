// BatchNorm2d (train / eval; torch defaults eps=1e-5, momentum=0.1; models/models.py:50,55,60,67,71,75,79),
// fused BN+ReLU+MaxPool2d(3, s2) forward (models/models.py:51-52,56-57,61-62) and the matching backward pieces.
// All tensors NHWC with C = 64.
#include "common.cuh"
#include "kernels.h"
#include "bn_tail.cuh"

namespace srlz {

// partials [n][128] -> S[0:64] = column sums of the first half, S[64:128] = of the second (double, fixed order)
__device__ __forceinline__ void reduce_partials_128(const float* __restrict__ partials, int n, double* s_buf /*[8][128]*/,
                                                    int tid) {
    const int grp = tid >> 7, j = tid & 127;
    double acc = 0.0;
    for (int r = grp; r < n; r += 8) acc += (double)partials[(size_t)r * 128 + j];
    s_buf[grp * 128 + j] = acc;
    __syncthreads();
    if (tid < 128) {
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) v += s_buf[g * 128 + tid];
        s_buf[tid] = v;  // group 0 row reused as the result (each thread only rewrites its own column)
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partials, int n, double count,
                                                           BnParams bn, int training, float* __restrict__ scale,
                                                           float* __restrict__ shift, float* __restrict__ mean_out,
                                                           float* __restrict__ invstd_out) {
    pdl_enter();
    __shared__ double s_buf[8 * 128];
    const int tid = threadIdx.x;
    if (training) reduce_partials_128(partials, n, s_buf, tid);
    if (tid < 64) {
        if (training) {
            bn_forward_finish(s_buf, count, bn, scale, tid);   // scale = bnsave
        } else {
            const float mean = bn.running_mean[tid], invstd = 1.0f / sqrtf(bn.running_var[tid] + (float)BN_EPS);
            const float sc = bn.gamma[tid] * invstd;
            scale[tid] = sc;
            shift[tid] = bn.beta[tid] - mean * sc;
            mean_out[tid] = mean;
            invstd_out[tid] = invstd;
        }
    }
}

// bnsave layout (5 x 64 floats): scale | shift | mean | invstd | biased batch variance
int bn_finalize(const float* partials, int n_partials, long long count, const BnParams& bn, int training, float* bnsave,
                cudaStream_t st) {
    launch_k(bn_finalize_kernel, 1, 1024, 0, st, partials, n_partials, (double)count, bn, training, bnsave, bnsave + 64,
                                           bnsave + 128, bnsave + 192);
    return check_launch("bn_finalize");
}

__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ var, double count,
                                         BnParams bn) {
    pdl_enter();
    const int tid = threadIdx.x;
    if (tid < 64) {
        const double m = mean[tid], v = var[tid];
        const double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
        bn.running_mean[tid] = (float)((1.0 - BN_MOM) * (double)bn.running_mean[tid] + BN_MOM * m);
        bn.running_var[tid] = (float)((1.0 - BN_MOM) * (double)bn.running_var[tid] + BN_MOM * unbiased);
    }
    if (tid == 0 && bn.num_batches_tracked != nullptr) *bn.num_batches_tracked += 1;
}

int bn_running_update(const float* bnsave, long long count, const BnParams& bn, cudaStream_t st) {
    launch_k(bn_running_update_kernel, 1, 64, 0, st, bnsave + 128, bnsave + 256, (double)count, bn);
    return check_launch("bn_running_update");
}

template <int N> struct IntC { static constexpr int value = N; };

static int ew_grid(long long total_threads) {
    long long b = (total_threads + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// out[n,ph,pw,c] = max over the 3x3 window (stride 2, `pad`, -inf padding) of relu(y*scale+shift); argmax = first maximal tap in
// scan order (torch max_pool2d semantics).  The kernel is the TMA-fed row-ring kernel of pool_tma.cu.
int bn_relu_pool_fwd(const float* y, const float* scale, const float* shift, float* out, unsigned char* argmax, int B,
                     int H, int W, int PH, int PW, int pad, cudaStream_t st) {
    if (!bn_relu_pool_fwd_tma_supported(y, W)) { set_error("bn_relu_pool_fwd: rows wider than 256 pixels or a misaligned tensor"); return 1; }
    return bn_relu_pool_fwd_tma(y, scale, shift, out, argmax, B, H, W, PH, PW, pad, st);
}

// dy[n,h,w,c] = gamma*invstd*(dz - c1 - xhat*c2)  with  dz = relu_mask * (sum of dpool over the windows whose argmax is (h,w)):
// MaxPool(3, s2) + ReLU + BatchNorm backward of a pooled stage in ONE full-size pass (c1, c2 from pool_bwd_stats).
// One CTA walks whole input rows (n, h).  Window q covers columns 2q-pad .. 2q-pad+2, so the columns pair up as
// (m, s) = (2q-pad+1, 2q-pad+2): m lies in window column q only (kx = 1), s in q (kx = 2) and q+1 (kx = 0).  A thread
// takes one column pair x 4 channels: the two window columns (x at most two window rows) are loaded once for both
// pixels -- 10 loads per 2 pixels instead of 18.  Accumulation order per pixel: (ph, pw) ascending, as in the reference.
__global__ void __launch_bounds__(256, 3) pool_bwd_bn_apply_kernel(const float* __restrict__ dpool, const unsigned char* __restrict__ argmax,
                                                                const float* __restrict__ y, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ coef, float* __restrict__ dy, int B, int H, int W,
                                                                int PH, int PW, int pad) {
    pdl_enter();
    const int tid = threadIdx.x, c4 = tid & 15, q00 = tid >> 4;
    const float4 sc = ldg4(scale + c4 * 4), sh = ldg4(shift + c4 * 4), me = ldg4(mean + c4 * 4);
    float4 ka, kb, kc;   // dy = ka*dz + kb*(y - mean) + kc  ==  gamma*invstd*(dz - c1 - (y - mean)*invstd*c2), folded per channel
    {
        const float4 iv = ldg4(invstd + c4 * 4), ga = ldg4(gamma + c4 * 4), c1 = ldg4(coef + c4 * 4), c2 = ldg4(coef + 64 + c4 * 4);
        ka = make_float4(ga.x * iv.x, ga.y * iv.y, ga.z * iv.z, ga.w * iv.w);
        kb = make_float4(-ka.x * iv.x * c2.x, -ka.y * iv.y * c2.y, -ka.z * iv.z * c2.z, -ka.w * iv.w * c2.w);
        kc = make_float4(-ka.x * c1.x, -ka.y * c1.y, -ka.z * c1.z, -ka.w * c1.w);
    }
    const int nrows = B * H;
    const int nq = (W + pad) / 2 + 1;   // q = -1 .. nq-2 : every column of the row is the m or the s of exactly one q
    auto finish = [&](float4 g, float4 yp) {
        float4 r;
        r.x = fmaf(yp.x, sc.x, sh.x) > 0.f ? g.x : 0.f;
        r.y = fmaf(yp.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
        r.z = fmaf(yp.z, sc.z, sh.z) > 0.f ? g.z : 0.f;
        r.w = fmaf(yp.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
        r.x = fmaf(ka.x, r.x, fmaf(kb.x, yp.x - me.x, kc.x));
        r.y = fmaf(ka.y, r.y, fmaf(kb.y, yp.y - me.y, kc.y));
        r.z = fmaf(ka.z, r.z, fmaf(kb.z, yp.z - me.z, kc.z));
        r.w = fmaf(ka.w, r.w, fmaf(kb.w, yp.w - me.w, kc.w));
        return r;
    };
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int n = row / H, h = row - n * H;
        // window rows covering input row h: ph in [ceil((h+pad-2)/2), floor((h+pad)/2)] clipped to [0, PH)
        const int th = h + pad - 2;
        const int ph_a = th <= 0 ? 0 : (th + 1) >> 1;
        const int ph_b = min((h + pad) >> 1, PH - 1);
        const float* yrow = y + (size_t)row * W * 64 + c4 * 4;
        float* drow = dy + (size_t)row * W * 64 + c4 * 4;
        auto columns = [&](auto NWIN) {   // NWIN = window rows covering this input row (compile-time per code path)
        constexpr int nwin = decltype(NWIN)::value;
        for (int qi = q00; qi < nq; qi += 16) {
            const int q = qi - 1;
            const int wm = 2 * q - pad + 1, ws = wm + 1;
            const bool m_ok = wm >= 0 && wm < W, s_ok = ws >= 0 && ws < W;
            const bool qa_ok = q >= 0 && q < PW, qb_ok = q + 1 < PW;   // window columns q, q+1
            float4 ym = make_float4(0.f, 0.f, 0.f, 0.f), ys = ym;
            if (m_ok) ym = ldg4(yrow + (size_t)wm * 64);
            if (s_ok) ys = ldg4(yrow + (size_t)ws * 64);
            float4 d[nwin > 0 ? nwin : 1][2];
            uchar4 am[nwin > 0 ? nwin : 1][2];
            bool ok[nwin > 0 ? nwin : 1][2];
#pragma unroll
            for (int i = 0; i < nwin; ++i) {
                const int ph = ph_a + i;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const bool wok = j == 0 ? qa_ok : qb_ok;
                    ok[i][j] = wok;
                    const size_t po = (((size_t)n * PH + ph) * PW + (wok ? q + j : 0)) * 64 + c4 * 4;
                    am[i][j] = *reinterpret_cast<const uchar4*>(argmax + po);
                    d[i][j] = ldg4(dpool + po);
                }
            }
            float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), gs = gm;
#pragma unroll
            for (int i = 0; i < nwin; ++i) {
                const int ky3 = (h - ((ph_a + i) * 2 - pad)) * 3;
                const unsigned char t_m = (unsigned char)(ky3 + 1), t_s0 = (unsigned char)(ky3 + 2), t_s1 = (unsigned char)ky3;
                if (ok[i][0]) {   // window column q: m is its kx = 1, s its kx = 2
                    if (am[i][0].x == t_m) gm.x += d[i][0].x;
                    if (am[i][0].y == t_m) gm.y += d[i][0].y;
                    if (am[i][0].z == t_m) gm.z += d[i][0].z;
                    if (am[i][0].w == t_m) gm.w += d[i][0].w;
                    if (am[i][0].x == t_s0) gs.x += d[i][0].x;
                    if (am[i][0].y == t_s0) gs.y += d[i][0].y;
                    if (am[i][0].z == t_s0) gs.z += d[i][0].z;
                    if (am[i][0].w == t_s0) gs.w += d[i][0].w;
                }
                if (ok[i][1]) {   // window column q+1: s is its kx = 0
                    if (am[i][1].x == t_s1) gs.x += d[i][1].x;
                    if (am[i][1].y == t_s1) gs.y += d[i][1].y;
                    if (am[i][1].z == t_s1) gs.z += d[i][1].z;
                    if (am[i][1].w == t_s1) gs.w += d[i][1].w;
                }
            }
            if (m_ok) st4(drow + (size_t)wm * 64, finish(gm, ym));
            if (s_ok) st4(drow + (size_t)ws * 64, finish(gs, ys));
        }
        };
        // rows with (h + pad) odd lie in ONE window row (its ky = 1), the others in two (ky = 2 of the upper window, ky = 0 of the
        // lower one), rows below the last window in none: one specialised code path each (the choice is uniform per row)
        const int nw = ph_b - ph_a + 1;
        if (nw >= 2) columns(IntC<2>{}); else if (nw == 1) columns(IntC<1>{}); else columns(IntC<0>{});
    }
}

// BatchNorm-backward statistics of a pooled stage taken on the POOLED side.  Max-pool backward routes every dpool element to
// exactly one input position (its argmax), so  sum_pos dz = sum_p m*dpool[p]  and  sum_pos dz*xhat = sum_p m*dpool[p]*xhat[argmax p]
// with m = (a[p] > 0) (the ReLU mask of the winning element) and, where m holds, a[p] = gamma*xhat + beta, i.e.
// xhat = (a[p] - beta)/gamma.  This reads two quarter-size tensors instead of two full-size ones and lets the single
// full-size pass (pool_bwd_bn_apply) write the finished dy.  gamma == 0 (xhat not recoverable from a) gathers y instead.
__global__ void __launch_bounds__(256) pool_bwd_stats_kernel(const float* __restrict__ dpool, const float* __restrict__ a,
                                                             const unsigned char* __restrict__ argmax, const float* __restrict__ y,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             float* __restrict__ partials, long long nquads, int H, int W,
                                                             int PH, int PW, int pad) {
    pdl_enter();
    __shared__ float s_red[8][128];
    const int tid = threadIdx.x, c4 = tid & 15;
    const float4 ga4 = ldg4(gamma + c4 * 4), be4 = ldg4(beta + c4 * 4);
    const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long idx = (long long)blockIdx.x * blockDim.x + tid; idx < nquads; idx += (long long)gridDim.x * blockDim.x) {
        const float4 d4 = ldg4(dpool + idx * 4), a4 = ldg4(a + idx * 4);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!(av[j] > 0.f)) continue;
            float xh;
            if (ga[j] != 0.f) {
                xh = (av[j] - be[j]) / ga[j];
            } else {
                const long long p = idx >> 4;
                const int pw = (int)(p % PW), ph = (int)((p / PW) % PH);
                const long long n = p / ((long long)PW * PH);
                const int tap = argmax[idx * 4 + j], c = c4 * 4 + j;
                const int h = ph * 2 - pad + tap / 3, w = pw * 2 - pad + tap % 3;
                xh = (y[((n * H + h) * W + w) * 64 + c] - mean[c]) * invstd[c];
            }
            st1[j] += d[j];
            st2[j] = fmaf(d[j], xh, st2[j]);
        }
    }
    const int wp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 16);
        st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 16);
    }
    if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s_red[wp][lane * 4 + j] = st1[j];
            s_red[wp][64 + lane * 4 + j] = st2[j];
        }
    }
    __syncthreads();
    if (tid < 128) {
        float v = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) v += s_red[ww][tid];
        partials[(size_t)blockIdx.x * 128 + tid] = v;
    }
}

int pool_bwd_stats(const float* dpool, const float* a, const unsigned char* argmax, const float* y, const float* gamma,
                   const float* beta, const float* mean, const float* invstd, float* partials, int* n_partials, int B, int H,
                   int W, int PH, int PW, int pad, cudaStream_t st) {
    const long long nquads = (long long)B * PH * PW * 16;
    int gx = ew_grid(nquads);
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    if (n_partials) *n_partials = gx;
    launch_k(pool_bwd_stats_kernel, gx, 256, 0, st, dpool, a, argmax, y, gamma, beta, mean, invstd, partials, nquads, H, W, PH, PW, pad);
    return check_launch("pool_bwd_stats");
}

int pool_bwd_bn_apply(const float* dpool, const unsigned char* argmax, const float* y, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const float* gamma, const float* coef, float* dy, int B, int H,
                      int W, int PH, int PW, int pad, cudaStream_t st) {
    int gx = B * H;
    if (gx > sm_count() * 8) gx = sm_count() * 8;
    launch_k(pool_bwd_bn_apply_kernel, gx, 256, 0, st, dpool, argmax, y, scale, shift, mean, invstd, gamma, coef, dy, B, H, W, PH, PW, pad);
    return check_launch("pool_bwd_bn_apply");
}

__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partials, int n, double count,
                                                               float* __restrict__ coef, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, int accumulate) {
    pdl_enter();
    __shared__ double s_buf[8 * 128];
    const int tid = threadIdx.x;
    reduce_partials_128(partials, n, s_buf, tid);
    if (tid < 64) bn_backward_finish(s_buf, count, coef, dgamma, dbeta, accumulate, tid);
}

int bn_bwd_finalize(const float* partials, int n_partials, long long count, float* coef, float* dgamma, float* dbeta,
                    int accumulate, cudaStream_t st) {
    launch_k(bn_bwd_finalize_kernel, 1, 1024, 0, st, partials, n_partials, (double)count, coef, dgamma, dbeta, accumulate);
    return check_launch("bn_bwd_finalize");
}

// dy = gamma*invstd*(dz - c1 - xhat*c2), in place; optional per-CTA column sums of dy (conv bias gradient)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dz, const float* __restrict__ y,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ coef,
                                                           long long npix, float* __restrict__ partials) {
    pdl_enter();
    __shared__ float s_red[8][64];
    const int tid = threadIdx.x, c4 = tid & 15;
    const float4 ga = ldg4(gamma + c4 * 4), me = ldg4(mean + c4 * 4), iv = ldg4(invstd + c4 * 4);
    const float4 c1 = ldg4(coef + c4 * 4), c2 = ldg4(coef + 64 + c4 * 4);
    const long long total = npix * 16;
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long idx = (long long)blockIdx.x * blockDim.x + tid; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const size_t off = (size_t)idx * 4;
        const float4 d = *reinterpret_cast<const float4*>(dz + off);
        const float4 yp = ldg4(y + off);
        float4 r;
        r.x = ga.x * iv.x * (d.x - c1.x - (yp.x - me.x) * iv.x * c2.x);
        r.y = ga.y * iv.y * (d.y - c1.y - (yp.y - me.y) * iv.y * c2.y);
        r.z = ga.z * iv.z * (d.z - c1.z - (yp.z - me.z) * iv.z * c2.z);
        r.w = ga.w * iv.w * (d.w - c1.w - (yp.w - me.w) * iv.w * c2.w);
        bs[0] += r.x; bs[1] += r.y; bs[2] += r.z; bs[3] += r.w;
        st4(dz + off, r);
    }
    if (partials != nullptr) {
        const int wp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < 4; ++j) bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 16);
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s_red[wp][lane * 4 + j] = bs[j];
        }
        __syncthreads();
        if (tid < 64) {
            float v = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) v += s_red[ww][tid];
            partials[(size_t)blockIdx.x * 64 + tid] = v;
        }
    }
}

__global__ void __launch_bounds__(1024) rows_sum64_kernel(const float* __restrict__ partials, int n,
                                                          float* __restrict__ out, int accumulate) {
    pdl_enter();
    __shared__ double s_buf[16][64];
    const int tid = threadIdx.x, grp = tid >> 6, j = tid & 63;
    double v = 0.0;
    for (int r = grp; r < n; r += 16) v += (double)partials[(size_t)r * 64 + j];
    s_buf[grp][j] = v;
    __syncthreads();
    if (tid < 64) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 16; ++g) t += s_buf[g][tid];
        out[tid] = accumulate ? out[tid] + (float)t : (float)t;
    }
}

int bn_bwd_apply(float* dz, const float* y, const float* gamma, const float* mean, const float* invstd, const float* coef,
                 long long npix, float* dbias, float* partials, int accumulate, cudaStream_t st) {
    int gx = ew_grid(npix * 16);
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    launch_k(bn_bwd_apply_kernel, gx, 256, 0, st, dz, y, gamma, mean, invstd, coef, npix, dbias != nullptr ? partials : nullptr);
    int rc = check_launch("bn_bwd_apply");
    if (rc || dbias == nullptr) return rc;
    launch_k(rows_sum64_kernel, 1, 1024, 0, st, partials, gx, dbias, accumulate);
    return check_launch("rows_sum64");
}

}  // namespace srlz

// Small dense / elementwise kernels of the train step: the bottleneck Linear layers
// (models/autoencoders.py:95,99 ; models/vae.py:52-56), VAE reparameterise + KL (models/models.py:147-165,
// losses/losses.py:239-256), squared-error reductions (losses/losses.py:172-181,199-214), Adam
// (models/learner.py:199) and weight (un)packing between torch-native and kernel layouts.
#include "common.cuh"
#include "kernels.h"

namespace srlz {

// ---------------------------------------------------------------------------------------------
// generic strided SGEMM: C[i,j] (+)= sum_k A(i,k) B(k,j) + bias[j].  64x64 tile, BK=16, 4x4 per thread.
// ---------------------------------------------------------------------------------------------
// One K tile (16) of both operands: each thread fetches 4 A and 4 B elements into registers (orientation picked so that the
// unit-stride dimension is the fastest-varying one across threads) ...
struct SgemmFrag { float a[4], b[4]; };
__device__ __forceinline__ void sgemm_fetch(SgemmFrag& f, const float* __restrict__ A, long long sai, long long sak, const float* __restrict__ B,
                                            long long sbk, long long sbj, int i0, int j0, int k0, int kend, int M, int N, int tid) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = tid + 256 * r;  // 0..1023
        int ai, ak;
        if (sak == 1) { ak = e & 15; ai = e >> 4; } else { ai = e & 63; ak = e >> 6; }
        const int gi = i0 + ai, gk = k0 + ak;
        f.a[r] = (gi < M && gk < kend) ? __ldg(A + gi * sai + gk * sak) : 0.f;
        int bj, bk;
        if (sbk == 1) { bk = e & 15; bj = e >> 4; } else { bj = e & 63; bk = e >> 6; }
        const int gj = j0 + bj, gk2 = k0 + bk;
        f.b[r] = (gj < N && gk2 < kend) ? __ldg(B + gk2 * sbk + gj * sbj) : 0.f;
    }
}
// ... and parks them in one of the two shared-memory buffers
__device__ __forceinline__ void sgemm_park(const SgemmFrag& f, float (*As)[64 + 4], float (*Bs)[64 + 4], long long sak, long long sbk, int tid) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = tid + 256 * r;
        int ai, ak;
        if (sak == 1) { ak = e & 15; ai = e >> 4; } else { ai = e & 63; ak = e >> 6; }
        As[ak][ai] = f.a[r];
        int bj, bk;
        if (sbk == 1) { bk = e & 15; bj = e >> 4; } else { bj = e & 63; bk = e >> 6; }
        Bs[bk][bj] = f.b[r];
    }
}
__device__ __forceinline__ void sgemm_tile_fma(float (&acc)[4][4], const float (*As)[64 + 4], const float (*Bs)[64 + 4], int tx, int ty) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}
// K loop over [kbeg, kend): double-buffered shared memory, the next tile's global loads are in flight while this one is multiplied
// (these products run on a handful of CTAs per SM, so the exposed load latency of a single-buffered loop was most of their time)
__device__ __forceinline__ void sgemm_mainloop(float (&acc)[4][4], const float* __restrict__ A, long long sai, long long sak,
                                               const float* __restrict__ B, long long sbk, long long sbj, int i0, int j0, int kbeg, int kend,
                                               int M, int N, float (*As)[16][64 + 4], float (*Bs)[16][64 + 4]) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    SgemmFrag f;
    sgemm_fetch(f, A, sai, sak, B, sbk, sbj, i0, j0, kbeg, kend, M, N, tid);
    sgemm_park(f, As[0], Bs[0], sak, sbk, tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += 16, buf ^= 1) {
        const bool more = k0 + 16 < kend;
        if (more) sgemm_fetch(f, A, sai, sak, B, sbk, sbj, i0, j0, k0 + 16, kend, M, N, tid);
        sgemm_tile_fma(acc, As[buf], Bs[buf], tx, ty);
        if (more) sgemm_park(f, As[buf ^ 1], Bs[buf ^ 1], sak, sbk, tid);   // the other buffer was last read one iteration ago
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long long sai, long long sak,
                                                    const float* __restrict__ B, long long sbk, long long sbj,
                                                    float* __restrict__ C, long long sci, long long scj,
                                                    const float* __restrict__ bias, int M, int N, int K, int accumulate, int perm) {
    pdl_enter();
    __shared__ float As[2][16][64 + 4];
    __shared__ float Bs[2][16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    sgemm_mainloop(acc, A, sai, sak, B, sbk, sbj, i0, j0, 0, K, M, N, As, Bs);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty * 4 + i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tx * 4 + j;
            if (gj >= N) continue;
            float v = acc[i][j];
            if (bias != nullptr) v += __ldg(bias + gj);
            // perm: the row (1) / column (2) index runs over the 2304 bottleneck features in the kernels' NHWC order hw*64+c and
            // is written at its torch position c*36+hw (models/autoencoders.py:108) -- the layout conversion of the FC weight
            // gradients, formerly a separate pass
            const long long oi = perm == 1 ? (gi & 63) * 36 + (gi >> 6) : gi, oj = perm == 2 ? (gj & 63) * 36 + (gj >> 6) : gj;
            float* c = C + oi * sci + oj * scj;
            *c = accumulate ? *c + v : v;
        }
    }
}

int sgemm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
          long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, cudaStream_t st);
int sgemm_perm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
               long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, int perm, cudaStream_t st);

// split-K variant for the K = 2304 bottleneck products (few output tiles): grid.z slices of K write partial tiles into `ws`
// ([splits][M][N]), a second kernel adds them in slice order (+bias) -- deterministic, no atomics.
__global__ void __launch_bounds__(256) sgemm_splitk_kernel(const float* __restrict__ A, long long sai, long long sak,
                                                           const float* __restrict__ B, long long sbk, long long sbj,
                                                           float* __restrict__ ws, int M, int N, int K, int kchunk) {
    pdl_enter();
    __shared__ float As[2][16][64 + 4];
    __shared__ float Bs[2][16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (kbeg < kend) sgemm_mainloop(acc, A, sai, sak, B, sbk, sbj, i0, j0, kbeg, kend, M, N, As, Bs);
    float* out = ws + (size_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty * 4 + i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tx * 4 + j;
            if (gj < N) out[(size_t)gi * N + gj] = acc[i][j];
        }
    }
}

__global__ void sgemm_splitk_reduce_kernel(const float* __restrict__ ws, int splits, float* __restrict__ C, long long sci, long long scj,
                                           const float* __restrict__ bias, int M, int N, int accumulate) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int i = idx / N, j = idx % N;
    float v = bias != nullptr ? __ldg(bias + j) : 0.f;
    for (int z = 0; z < splits; ++z) v += ws[(size_t)z * M * N + idx];
    float* c = C + i * sci + j * scj;
    *c = accumulate ? *c + v : v;
}

int sgemm_splitk(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C, long long sci,
                 long long scj, const float* bias, int M, int N, int K, int accumulate, float* ws, size_t ws_floats, cudaStream_t st) {
    int splits = 16;
    while (splits > 1 && (size_t)splits * M * N > ws_floats) splits >>= 1;
    int kchunk = ((K + splits - 1) / splits + 15) / 16 * 16;
    splits = (K + kchunk - 1) / kchunk;
    if (splits <= 1) return sgemm(A, sai, sak, B, sbk, sbj, C, sci, scj, bias, M, N, K, accumulate, st);
    dim3 grid((N + 63) / 64, (M + 63) / 64, splits);
    launch_k(sgemm_splitk_kernel, grid, 256, 0, st, A, sai, sak, B, sbk, sbj, ws, M, N, K, kchunk);
    int rc = check_launch("sgemm_splitk");
    if (rc) return rc;
    launch_k(sgemm_splitk_reduce_kernel, (M * N + 255) / 256, 256, 0, st, ws, splits, C, sci, scj, bias, M, N, accumulate);
    return check_launch("sgemm_splitk_reduce");
}

int sgemm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
          long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, cudaStream_t st) {
    return sgemm_perm(A, sai, sak, B, sbk, sbj, C, sci, scj, bias, M, N, K, accumulate, 0, st);
}

int sgemm_perm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
               long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, int perm, cudaStream_t st) {
    if ((perm == 1 && M != 2304) || (perm == 2 && N != 2304)) { set_error("sgemm_perm: the permuted axis must have 2304 entries"); return 1; }
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    launch_k(sgemm_kernel, grid, 256, 0, st, A, sai, sak, B, sbk, sbj, C, sci, scj, bias, M, N, K, accumulate, perm);
    return check_launch("sgemm");
}

// out[j] (+)= sum_i A[i,j]: 64 columns x 4 row groups per CTA (row group g adds rows g, g+4, ...; the groups are folded in a
// fixed order).  perm: j runs over the 2304 bottleneck features in NHWC order and is written at its torch position.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int M, int N, float* __restrict__ out, int accumulate, int perm) {
    pdl_enter();
    __shared__ float s_part[4][64];
    const int tid = threadIdx.x, g = tid >> 6, j = blockIdx.x * 64 + (tid & 63);
    float s = 0.f;
    if (j < N) {
#pragma unroll 4
        for (int i = g; i < M; i += 4) s += A[(size_t)i * N + j];
    }
    s_part[g][tid & 63] = s;
    __syncthreads();
    if (g == 0 && j < N) {
        const float v = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
        const int o = perm ? (j & 63) * 36 + (j >> 6) : j;
        out[o] = accumulate ? out[o] + v : v;
    }
}

int colsum(const float* A, int M, int N, float* out, int accumulate, cudaStream_t st) { return colsum_perm(A, M, N, out, accumulate, 0, st); }

int colsum_perm(const float* A, int M, int N, float* out, int accumulate, int perm, cudaStream_t st) {
    if (perm && N != 2304) { set_error("colsum_perm: the permuted axis must have 2304 entries"); return 1; }
    launch_k(colsum_kernel, (N + 63) / 64, 256, 0, st, A, M, N, out, accumulate, perm);
    return check_launch("colsum");
}

// ---------------------------------------------------------------------------------------------
// block-level scalar reduction helper: per-CTA partial (fixed order) into partials[blockIdx.x]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_partial(float v, float* partials) {
    __shared__ float s_w[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_w[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) t += s_w[i];
        partials[blockIdx.x] = t;
    }
}

static int red_grid(long long n_threads) {
    long long b = (n_threads + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// z = eps*exp(0.5*logvar)+mu (train) | mu (eval) ; partial of sum(1 + logvar - mu^2 - exp(logvar))
__global__ void __launch_bounds__(256) vae_reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                              const float* __restrict__ eps, float* __restrict__ z,
                                                              float* __restrict__ partials, int n, int training) {
    pdl_enter();
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float m = mu[i], lv = logvar[i];
        z[i] = training ? fmaf(eps[i], expf(0.5f * lv), m) : m;
        s += 1.f + lv - m * m - expf(lv);
    }
    block_partial(s, partials);
}

int vae_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, float* kl_partials, int n,
                    int training, int* n_partials, cudaStream_t st) {
    const int gx = red_grid(n);
    if (n_partials) *n_partials = gx;
    launch_k(vae_reparam_fwd_kernel, gx, 256, 0, st, mu, logvar, eps, z, kl_partials, n, training);
    return check_launch("vae_reparam_fwd");
}

// dmu = dz + kl_coef*mu (+ extra) ; dlogvar = dz*eps*0.5*exp(0.5 lv) + kl_coef*0.5*(exp(lv)-1) (+ extra)
__global__ void vae_reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ logvar,
                                       const float* __restrict__ eps, const float* __restrict__ gmu_extra,
                                       const float* __restrict__ glv_extra, float kl_coef, const float* __restrict__ mu,
                                       float* __restrict__ dmu, float* __restrict__ dlogvar, int n, int training) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = dz[i], lv = logvar[i];
    float gm = g + kl_coef * mu[i];
    float gl = kl_coef * 0.5f * (expf(lv) - 1.f);
    if (training) gl = fmaf(g * eps[i], 0.5f * expf(0.5f * lv), gl);
    if (gmu_extra != nullptr) gm += gmu_extra[i];
    if (glv_extra != nullptr) gl += glv_extra[i];
    dmu[i] = gm;
    dlogvar[i] = gl;
}

int vae_reparam_bwd(const float* dz, const float* logvar, const float* eps, const float* gmu_extra,
                    const float* glogvar_extra, float kl_coef, const float* mu, float* dmu, float* dlogvar, int n,
                    int training, cudaStream_t st) {
    launch_k(vae_reparam_bwd_kernel, (n + 255) / 256, 256, 0, st, dz, logvar, eps, gmu_extra, glogvar_extra, kl_coef, mu, dmu,
                                                            dlogvar, n, training);
    return check_launch("vae_reparam_bwd");
}

// partial sums of (a-b)^2, 128-bit loads (n must be a multiple of 4 and pointers 16B aligned)
__global__ void __launch_bounds__(256) sse_partials_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           long long n4, float* __restrict__ partials) {
    pdl_enter();
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = ldg4(a + i * 4), y = ldg4(b + i * 4);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
    }
    block_partial(s, partials);
}

__global__ void sse_partials_scalar_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                           float* __restrict__ partials) {
    pdl_enter();
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        s = fmaf(d, d, s);
    }
    block_partial(s, partials);
}

int sse_partials(const float* a, const float* b, long long n, float* partials, int* n_partials, cudaStream_t st) {
    const bool vec = (n % 4 == 0) && (((uintptr_t)a | (uintptr_t)b) % 16 == 0);
    const int gx = red_grid(vec ? n / 4 : n);
    if (n_partials) *n_partials = gx;
    if (vec)
        launch_k(sse_partials_kernel, gx, 256, 0, st, a, b, n / 4, partials);
    else
        launch_k(sse_partials_scalar_kernel, gx, 256, 0, st, a, b, n, partials);
    return check_launch("sse_partials");
}

// out (+)= scale * sum(partials[0:n])  in double, fixed order
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partials, int n, float scale,
                                                           float* __restrict__ out, int accumulate) {
    pdl_enter();
    __shared__ double s_d[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v += (double)partials[i];
    s_d[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float r = (float)(s_d[0] * (double)scale);
        *out = accumulate ? *out + r : r;
    }
}

int sum_partials(const float* partials, int n, float scale, float* out, int accumulate, cudaStream_t st) {
    launch_k(sum_partials_kernel, 1, 256, 0, st, partials, n, scale, out, accumulate);
    return check_launch("sum_partials");
}

__global__ void mse_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4, float coef,
                                float* __restrict__ g) {
    pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = ldg4(a + i * 4), y = ldg4(b + i * 4);
        st4(g + i * 4, make_float4(coef * (x.x - y.x), coef * (x.y - y.y), coef * (x.z - y.z), coef * (x.w - y.w)));
    }
}

int mse_grad(const float* a, const float* b, long long n, float coef, float* g, cudaStream_t st) {
    if (n % 4 != 0) { set_error("mse_grad: n must be a multiple of 4"); return 1; }
    launch_k(mse_grad_kernel, red_grid(n / 4), 256, 0, st, a, b, n / 4, coef, g);
    return check_launch("mse_grad");
}

// torch.optim.Adam step (models/learner.py:199,495): denom = sqrt(v)/sqrt(bc2) + eps ; p -= lr/bc1 * m/denom
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt) {
    pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              float bc1, float bc2, cudaStream_t st) {
    launch_k(adam_kernel, red_grid(n), 256, 0, st, p, g, m, v, n, lr, b1, b2, eps, bc1, sqrtf(bc2));
    return check_launch("adam");
}

// ---------------------------------------------------------------------------------------------
// layout conversions.  NCHW flatten index c*36+hw (models/autoencoders.py:108)  <->  NHWC index hw*64+c
//   row_mode 0: matrix (rows, 2304): permute columns       (encoder fc weight)
//   row_mode 1: matrix (2304, rows): permute rows           (decoder fc weight; bias with rows = 1)
// ---------------------------------------------------------------------------------------------
__global__ void permute_fc_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int to_packed,
                                  int row_mode, int accumulate) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = rows * 2304;
    if (idx >= total) return;
    int r, q;  // q = index along the 2304 axis in the DESTINATION layout
    if (row_mode == 0) { r = idx / 2304; q = idx % 2304; } else { q = idx / rows; r = idx % rows; }
    int qs;    // matching index in the source layout
    if (to_packed) { const int hw = q >> 6, c = q & 63; qs = c * 36 + hw; }   // dst packed (hw*64+c) <- src torch
    else           { const int c = q / 36, hw = q % 36; qs = hw * 64 + c; }   // dst torch (c*36+hw) <- src packed
    const int sidx = (row_mode == 0) ? r * 2304 + qs : qs * rows + r;
    const float v = src[sidx];
    dst[idx] = accumulate ? dst[idx] + v : v;
}

int permute_fc(const float* src, float* dst, int rows, int to_packed, int row_mode, int accumulate, cudaStream_t st) {
    const int total = rows * 2304;
    launch_k(permute_fc_kernel, (total + 255) / 256, 256, 0, st, src, dst, rows, to_packed, row_mode, accumulate);
    return check_launch("permute_fc");
}

// torch W[a][b][tap] (a,b in 64).  Conv2d: a=co,b=ci ; ConvTranspose2d: a=ci,b=co.
//   fwd pack   [tap][ci][co]   (gathered = layer input)
//   dgrad pack [tap][co][ci]   (gathered = dy)
__global__ void pack_conv_w_kernel(const float* __restrict__ w, float* __restrict__ fwd, float* __restrict__ dgr, int ntaps,
                                   int transposed_conv) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 4096 * ntaps) return;
    const int tap = idx % ntaps, b = (idx / ntaps) & 63, a = idx / (ntaps * 64);
    const float v = w[idx];
    const int ci = transposed_conv ? a : b, co = transposed_conv ? b : a;
    fwd[(tap * 64 + ci) * 64 + co] = v;
    dgr[(tap * 64 + co) * 64 + ci] = v;
}

// Eval-mode BatchNorm folded into the conv before it (inference path, models/learner.py:67-88): w'[co][...] = w[co][...] * s[co],
// b'[co] = beta[co] - running_mean[co] * s[co], s = gamma / sqrt(running_var + eps); double arithmetic, one rounding to fp32
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float* __restrict__ w_out,
                               float* __restrict__ b_out, int per_co) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 64 * per_co) return;
    const int co = idx / per_co;
    const double s = (double)gamma[co] / sqrt((double)var[co] + 1e-5);
    w_out[idx] = (float)((double)w[idx] * s);
    if (idx % per_co == 0) b_out[co] = (float)((double)beta[co] - (double)mean[co] * s);
}

int fold_bn(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float* w_out, float* b_out,
            int per_co, cudaStream_t st) {
    launch_k(fold_bn_kernel, (64 * per_co + 255) / 256, 256, 0, st, w, gamma, beta, mean, var, w_out, b_out, per_co);
    return check_launch("fold_bn");
}

__global__ void fill_kernel(float* __restrict__ p, float v, int n) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
int fill(float* p, float v, int n, cudaStream_t st) {
    launch_k(fill_kernel, (n + 255) / 256, 256, 0, st, p, v, n);
    return check_launch("fill");
}

int pack_conv_w(const float* w, float* fwd_pack, float* dgrad_pack, int ntaps, int transposed_conv, cudaStream_t st) {
    launch_k(pack_conv_w_kernel, (4096 * ntaps + 255) / 256, 256, 0, st, w, fwd_pack, dgrad_pack, ntaps, transposed_conv);
    return check_launch("pack_conv_w");
}

}  // namespace srlz

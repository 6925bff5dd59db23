// Generic 64->64 channel "gather convolution" kernels (fp32 SIMT scaffold path).
//
// One kernel family covers, for the hidden layers of the reference's conv stacks
// (models/models.py:54,59 conv3x3 ; :66,70,74,78 ConvTranspose2d k3 s2):
//   forward of Conv2d / ConvTranspose2d, dgrad of both, and wgrad of both.
// Tensors are NHWC with C = 64; weights are pre-packed as [tap][c_gathered][c_out].
#include "common.cuh"
#include "kernels.h"

namespace srlz {

// ------------------------------------------------------------------------------------------
// forward / dgrad : out[m, co] = sum_{tap, cg} gather(in)[m, tap, cg] * W[tap][cg][co]
// CTA tile 128 output pixels x 64 channels, 256 threads, thread tile 8 pixels x 4 channels.
// ------------------------------------------------------------------------------------------
#define GC_AS 68  // padded row stride (floats) of the A tile in shared memory

template <bool TRANSPOSED, bool BN_LOAD, int EPI>
__global__ void __launch_bounds__(256, 2) gconv64_kernel(GConvArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                 // [128][GC_AS]
    float* Ws = smem + 128 * GC_AS;   // [64][64]
    __shared__ int s_n[128], s_oy[128], s_ox[128];
    __shared__ int s_taps[64];
    __shared__ int s_ntaps;
    __shared__ float s_red[8][128];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const ConvGeom g = a.g;
    const int s = g.stride;
    const int OH = TRANSPOSED ? g.BH : g.SH, OW = TRANSPOSED ? g.BW : g.SW;
    const int IH = TRANSPOSED ? g.SH : g.BH, IW = TRANSPOSED ? g.SW : g.BW;
    int py = 0, px = 0, cs = 1, OHc = OH, OWc = OW;
    if (TRANSPOSED) {
        cs = s;
        py = blockIdx.y / s;
        px = blockIdx.y % s;
        OHc = (OH - py + s - 1) / s;
        OWc = (OW - px + s - 1) / s;
    }
    const long long Mc = (long long)g.B * OHc * OWc;
    const int ntiles = (int)((Mc + 127) / 128);

    if (tid == 0) {
        int nt = 0;
        for (int ky = 0; ky < g.KH; ++ky)
            for (int kx = 0; kx < g.KW; ++kx) {
                if (TRANSPOSED) {
                    // need (oy + pad - ky) % s == 0 for every oy of this parity class
                    if (((py + g.pad - ky) % s + s) % s != 0) continue;
                    if (((px + g.pad - kx) % s + s) % s != 0) continue;
                }
                s_taps[nt++] = ky * g.KW + kx;
            }
        s_ntaps = nt;
    }

    float4 lsc = make_float4(1.f, 1.f, 1.f, 1.f), lsh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BN_LOAD) {
        lsc = ldg4(a.in_scale + tx * 4);
        lsh = ldg4(a.in_shift + tx * 4);
    }
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
    float4 esc, esh, emean, einv;
    if (EPI == EPI_MASK_BNBWD) {
        esc = ldg4(a.e_scale + tx * 4);
        esh = ldg4(a.e_shift + tx * 4);
        emean = ldg4(a.e_mean + tx * 4);
        einv = ldg4(a.e_invstd + tx * 4);
    }
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias != nullptr) bias4 = ldg4(a.bias + tx * 4);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        if (tid < 128) {
            long long m = (long long)tile * 128 + tid;
            if (m < Mc) {
                int oxc = (int)(m % OWc);
                long long t = m / OWc;
                int oyc = (int)(t % OHc);
                s_n[tid] = (int)(t / OHc);
                s_oy[tid] = oyc * cs + py;
                s_ox[tid] = oxc * cs + px;
            } else {
                s_n[tid] = -1;
                s_oy[tid] = 0;
                s_ox[tid] = 0;
            }
        }
        __syncthreads();
        const int ntaps = s_ntaps;

        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        float4 av[8], wv[4];
        auto load_tap = [&](int tap) {
            const int ky = tap / g.KW, kx = tap % g.KW;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = ty + 16 * j;
                const int n = s_n[p];
                int iy, ix;
                bool ok = n >= 0;
                if (TRANSPOSED) {
                    // parity already guaranteed by the tap list
                    const int ny = s_oy[p] + g.pad - ky, nx = s_ox[p] + g.pad - kx;
                    iy = ny / s;
                    ix = nx / s;
                    ok = ok && ny >= 0 && nx >= 0 && iy < IH && ix < IW;
                } else {
                    iy = s_oy[p] * s - g.pad + ky;
                    ix = s_ox[p] * s - g.pad + kx;
                    ok = ok && iy >= 0 && ix >= 0 && iy < IH && ix < IW;
                }
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    v = ldg4(a.in + (((size_t)n * IH + iy) * IW + ix) * SRLZ_C + tx * 4);
                    if (BN_LOAD) v = bn_relu4(v, lsc, lsh);
                }
                av[j] = v;
            }
            const float* wp = a.wpack + (size_t)tap * (SRLZ_C * SRLZ_C);
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = ldg4(wp + (tid + 256 * j) * 4);
        };

        if (ntaps > 0) load_tap(s_taps[0]);
        for (int ti = 0; ti < ntaps; ++ti) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) st4(As + (ty + 16 * j) * GC_AS + tx * 4, av[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) st4(Ws + (tid + 256 * j) * 4, wv[j]);
            __syncthreads();
            if (ti + 1 < ntaps) load_tap(s_taps[ti + 1]);
#pragma unroll 4
            for (int k4 = 0; k4 < 16; ++k4) {
                float4 af[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * GC_AS + k4 * 4);
                float4 b0 = *reinterpret_cast<const float4*>(Ws + (k4 * 4 + 0) * SRLZ_C + tx * 4);
                float4 b1 = *reinterpret_cast<const float4*>(Ws + (k4 * 4 + 1) * SRLZ_C + tx * 4);
                float4 b2 = *reinterpret_cast<const float4*>(Ws + (k4 * 4 + 2) * SRLZ_C + tx * 4);
                float4 b3 = *reinterpret_cast<const float4*>(Ws + (k4 * 4 + 3) * SRLZ_C + tx * 4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][0] = fmaf(af[i].x, b0.x, acc[i][0]);
                    acc[i][1] = fmaf(af[i].x, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(af[i].x, b0.z, acc[i][2]);
                    acc[i][3] = fmaf(af[i].x, b0.w, acc[i][3]);
                    acc[i][0] = fmaf(af[i].y, b1.x, acc[i][0]);
                    acc[i][1] = fmaf(af[i].y, b1.y, acc[i][1]);
                    acc[i][2] = fmaf(af[i].y, b1.z, acc[i][2]);
                    acc[i][3] = fmaf(af[i].y, b1.w, acc[i][3]);
                    acc[i][0] = fmaf(af[i].z, b2.x, acc[i][0]);
                    acc[i][1] = fmaf(af[i].z, b2.y, acc[i][1]);
                    acc[i][2] = fmaf(af[i].z, b2.z, acc[i][2]);
                    acc[i][3] = fmaf(af[i].z, b2.w, acc[i][3]);
                    acc[i][0] = fmaf(af[i].w, b3.x, acc[i][0]);
                    acc[i][1] = fmaf(af[i].w, b3.y, acc[i][1]);
                    acc[i][2] = fmaf(af[i].w, b3.z, acc[i][2]);
                    acc[i][3] = fmaf(af[i].w, b3.w, acc[i][3]);
                }
            }
        }

        // epilogue
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = ty + 16 * i;
            const int n = s_n[p];
            if (n < 0) continue;
            const size_t off = (((size_t)n * OH + s_oy[p]) * OW + s_ox[p]) * SRLZ_C + tx * 4;
            float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
            if (EPI == EPI_STATS) {
                st1[0] += v.x; st1[1] += v.y; st1[2] += v.z; st1[3] += v.w;
                st2[0] = fmaf(v.x, v.x, st2[0]); st2[1] = fmaf(v.y, v.y, st2[1]);
                st2[2] = fmaf(v.z, v.z, st2[2]); st2[3] = fmaf(v.w, v.w, st2[3]);
            } else if (EPI == EPI_MASK_BNBWD) {
                const float4 yp = ldg4(a.e_ypre + off);
                v.x = fmaf(yp.x, esc.x, esh.x) > 0.f ? v.x : 0.f;
                v.y = fmaf(yp.y, esc.y, esh.y) > 0.f ? v.y : 0.f;
                v.z = fmaf(yp.z, esc.z, esh.z) > 0.f ? v.z : 0.f;
                v.w = fmaf(yp.w, esc.w, esh.w) > 0.f ? v.w : 0.f;
                st1[0] += v.x; st1[1] += v.y; st1[2] += v.z; st1[3] += v.w;
                st2[0] = fmaf(v.x, (yp.x - emean.x) * einv.x, st2[0]);
                st2[1] = fmaf(v.y, (yp.y - emean.y) * einv.y, st2[1]);
                st2[2] = fmaf(v.z, (yp.z - emean.z) * einv.z, st2[2]);
                st2[3] = fmaf(v.w, (yp.w - emean.w) * einv.w, st2[3]);
            }
            st4(a.out + off, v);
        }
    }

    if (EPI != EPI_PLAIN) {
        const int w = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], 16);
            st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], 16);
        }
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_red[w][lane * 4 + j] = st1[j];
                s_red[w][64 + lane * 4 + j] = st2[j];
            }
        }
        __syncthreads();
        if (tid < 128) {
            float v = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) v += s_red[ww][tid];
            a.partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 128 + tid] = v;
        }
    }
}

template <bool T, bool BN, int EPI>
static int launch_gconv(const GConvArgs& a, int gx, int gy, cudaStream_t st) {
    const int smem = (128 * GC_AS + 64 * 64) * (int)sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(gconv64_kernel<T, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    gconv64_kernel<T, BN, EPI><<<dim3(gx, gy), 256, smem, st>>>(a);
    return check_launch("gconv64");
}

int gconv64(const GConvArgs& a_in, int* n_partials, cudaStream_t st) {
    GConvArgs a = a_in;
    const ConvGeom& g = a.g;
    const int nclass = a.transposed ? g.stride * g.stride : 1;
    const int OH = a.transposed ? g.BH : g.SH, OW = a.transposed ? g.BW : g.SW;
    const int s = a.transposed ? g.stride : 1;
    const long long Mc = (long long)g.B * ((OH + s - 1) / s) * ((OW + s - 1) / s);  // largest class
    const int ntiles = (int)((Mc + 127) / 128);
    int cap = (2 * sm_count() * 2) / nclass;
    if (cap < 1) cap = 1;
    const int gx = ntiles < cap ? ntiles : cap;
    if (n_partials) *n_partials = gx * nclass;
    if (a.epi != EPI_PLAIN && a.partials == nullptr) { set_error("gconv64: partials buffer required"); return 1; }
    const bool bn = a.in_scale != nullptr;
#define GC_DISPATCH(T, BN, E) return launch_gconv<T, BN, E>(a, gx, nclass, st)
    if (a.transposed) {
        if (bn) { if (a.epi == EPI_PLAIN) GC_DISPATCH(true, true, EPI_PLAIN); if (a.epi == EPI_STATS) GC_DISPATCH(true, true, EPI_STATS); GC_DISPATCH(true, true, EPI_MASK_BNBWD); }
        else    { if (a.epi == EPI_PLAIN) GC_DISPATCH(true, false, EPI_PLAIN); if (a.epi == EPI_STATS) GC_DISPATCH(true, false, EPI_STATS); GC_DISPATCH(true, false, EPI_MASK_BNBWD); }
    } else {
        if (bn) { if (a.epi == EPI_PLAIN) GC_DISPATCH(false, true, EPI_PLAIN); if (a.epi == EPI_STATS) GC_DISPATCH(false, true, EPI_STATS); GC_DISPATCH(false, true, EPI_MASK_BNBWD); }
        else    { if (a.epi == EPI_PLAIN) GC_DISPATCH(false, false, EPI_PLAIN); if (a.epi == EPI_STATS) GC_DISPATCH(false, false, EPI_STATS); GC_DISPATCH(false, false, EPI_MASK_BNBWD); }
    }
#undef GC_DISPATCH
}

// ------------------------------------------------------------------------------------------
// wgrad : P[tap][cg][cd] = sum_m gather(big)[m, tap, cg] * dense(small)[m, cd]
// grid (chunks, taps). 256 threads = 4 pixel sub-groups x 64 threads with 8x8 micro tiles.
// ------------------------------------------------------------------------------------------
template <bool BN_DENSE>
__global__ void __launch_bounds__(256, 2) gwgrad64_kernel(GWgradArgs a) {
    __shared__ __align__(16) float Gs[32][SRLZ_C];
    __shared__ __align__(16) float Ds[32][SRLZ_C];
    __shared__ __align__(16) float Rs[SRLZ_C * SRLZ_C];
    const int tid = threadIdx.x;
    const int q = tid >> 6, r = tid & 63;
    const int gq = r >> 3, dq = r & 7;  // gathered-channel group, dense-channel group
    const ConvGeom g = a.g;
    const int tap = blockIdx.y;
    const int ky = tap / g.KW, kx = tap % g.KW;
    const long long Ms = (long long)g.B * g.SH * g.SW;
    const long long m0 = (long long)blockIdx.x * a.chunk_len;
    long long m1 = m0 + a.chunk_len;
    if (m1 > Ms) m1 = Ms;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // load mapping: 32 pixels x 16 float4 = 512 float4 per operand, 2 per thread
    const int lp0 = tid >> 4, lc = (tid & 15) * 4;  // pixels lp0 and lp0+16
    float4 dsc = make_float4(1.f, 1.f, 1.f, 1.f), dsh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BN_DENSE) {
        dsc = ldg4(a.dense_scale + lc);
        dsh = ldg4(a.dense_shift + lc);
    }

    for (long long mb = m0; mb < m1; mb += 32) {
        float4 gv[2], dv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const long long m = mb + lp0 + 16 * j;
            gv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < m1) {
                const int sx = (int)(m % g.SW);
                const long long t = m / g.SW;
                const int sy = (int)(t % g.SH);
                const int n = (int)(t / g.SH);
                const int by = sy * g.stride - g.pad + ky, bx = sx * g.stride - g.pad + kx;
                if (by >= 0 && bx >= 0 && by < g.BH && bx < g.BW) {
                    gv[j] = ldg4(a.big + (((size_t)n * g.BH + by) * g.BW + bx) * SRLZ_C + lc);
                    float4 d = ldg4(a.small + (size_t)m * SRLZ_C + lc);
                    if (BN_DENSE) d = bn_relu4(d, dsc, dsh);
                    dv[j] = d;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            st4(&Gs[lp0 + 16 * j][lc], gv[j]);
            st4(&Ds[lp0 + 16 * j][lc], dv[j]);
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
            const int p = q * 8 + pp;
            const float4 g0 = *reinterpret_cast<const float4*>(&Gs[p][gq * 4]);
            const float4 g1 = *reinterpret_cast<const float4*>(&Gs[p][32 + gq * 4]);
            const float4 d0 = *reinterpret_cast<const float4*>(&Ds[p][dq * 4]);
            const float4 d1 = *reinterpret_cast<const float4*>(&Ds[p][32 + dq * 4]);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gg[i], dd[j], acc[i][j]);
        }
    }

    // reduce the 4 pixel sub-groups in fixed order, then write this CTA's partial
    for (int round = 0; round < 4; ++round) {
        __syncthreads();
        if (q == round) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cg = (i < 4) ? gq * 4 + i : 32 + gq * 4 + (i - 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cd = (j < 4) ? dq * 4 + j : 32 + dq * 4 + (j - 4);
                    float v = acc[i][j];
                    if (round > 0) v += Rs[cg * SRLZ_C + cd];
                    Rs[cg * SRLZ_C + cd] = v;
                }
            }
        }
    }
    __syncthreads();
    float* dst = a.partials + ((size_t)blockIdx.x * gridDim.y + tap) * (SRLZ_C * SRLZ_C);
    for (int i = tid; i < SRLZ_C * SRLZ_C / 4; i += 256) st4(dst + i * 4, *reinterpret_cast<const float4*>(Rs + i * 4));
}

// out[(cd*64 + cg)*ntaps + tap] (+)= sum_chunk partials[chunk][tap][cg][cd]   (torch OIHW / IOHW layout)
__global__ void gwgrad64_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out, int nchunks,
                                       int ntaps, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over tap*4096 + cg*64 + cd
    const int total = ntaps * SRLZ_C * SRLZ_C;
    if (idx >= total) return;
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += partials[(size_t)c * total + idx];
    const int tap = idx / (SRLZ_C * SRLZ_C);
    const int cg = (idx / SRLZ_C) % SRLZ_C, cd = idx % SRLZ_C;
    const int o = (cd * SRLZ_C + cg) * ntaps + tap;
    out[o] = accumulate ? out[o] + s : s;
}

int gwgrad64_reduce(const float* partials, float* grad_out, int nchunks, int ntaps, int accumulate, cudaStream_t st) {
    const int total = ntaps * SRLZ_C * SRLZ_C;
    gwgrad64_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partials, grad_out, nchunks, ntaps, accumulate);
    return check_launch("gwgrad64_reduce");
}

int gwgrad64_chunks(const ConvGeom& g) {
    const long long Ms = (long long)g.B * g.SH * g.SW;
    const int ntaps = g.KH * g.KW;
    int want = (2 * sm_count() * 2 + ntaps - 1) / ntaps;  // ~4 CTAs per SM overall
    long long maxc = (Ms + 255) / 256;                      // at least 256 pixels per chunk
    if (maxc < 1) maxc = 1;
    if (want > maxc) want = (int)maxc;
    if (want < 1) want = 1;
    return want;
}

int gwgrad64(const GWgradArgs& a_in, float* grad_out, int accumulate, cudaStream_t st) {
    GWgradArgs a = a_in;
    const ConvGeom& g = a.g;
    const long long Ms = (long long)g.B * g.SH * g.SW;
    const int ntaps = g.KH * g.KW;
    const int nch = gwgrad64_chunks(g);
    long long len = (Ms + nch - 1) / nch;
    len = (len + 31) / 32 * 32;
    a.chunk_len = (int)len;
    const int chunks = (int)((Ms + len - 1) / len);
    if (a.dense_scale)
        gwgrad64_kernel<true><<<dim3(chunks, ntaps), 256, 0, st>>>(a);
    else
        gwgrad64_kernel<false><<<dim3(chunks, ntaps), 256, 0, st>>>(a);
    int rc = check_launch("gwgrad64");
    if (rc) return rc;
    return gwgrad64_reduce(a.partials, grad_out, chunks, ntaps, accumulate, st);
}

size_t gwgrad64_partial_floats(const ConvGeom& g) {
    int chunks = gwgrad64_chunks(g) + 1;
    if (chunks < sm_count()) chunks = sm_count();  // the tcgen05 version writes one partial per CTA (<= #SMs)
    return (size_t)chunks * g.KH * g.KW * SRLZ_C * SRLZ_C;
}

}  // namespace srlz

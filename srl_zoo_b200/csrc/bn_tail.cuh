// "The last CTA finalizes": the tensor-core kernels that leave per-CTA partial sums of a BatchNorm reduction (forward statistics,
// backward sums) used to be followed by a one-CTA finalize launch each.  With a BnTail in its arguments the producer does the
// finalize itself: every CTA writes its partial row, fences, and takes a ticket from a global counter; the CTA that draws the last
// ticket re-reads all rows through L2 (fixed order, double precision -- the same arithmetic as the stand-alone kernels, which
// remain for callers without a counter) and writes the finished quantities.  One integer atomic per CTA, no floating-point
// atomics: results do not depend on the order in which CTAs finish.  The counter is reset by the last CTA.
// Measured (profiles/r2_fused_finalize.md): 22 launches fewer per train step, -0.23 ms of 11.6 (bs = 128: -0.14 of 6.63); the tail
// itself (fence, ticket, 148 rows through L2 by all 512 threads, finish) costs a few us while the rest of the GPU idles, against
// ~12 us for the separate launch.
// The elementwise producers (up to 1184 rows per launch) keep their finalize launches: a one-CTA tail over 600 KB takes 20 us
// and a two-level tail (groups of 8 rows first) 6-8 us -- no better than the launch it replaces (both measured).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace srlz {

#define BN_EPS 1e-5
#define BN_MOM 0.1

// column sums of rows [n][128] -> S[128] in s_buf[0..128) (double, fixed order).  s_buf: 16 * 128 doubles.
// 128-bit loads: 32 threads per row, 16 row groups (512 threads; fewer threads -> fewer groups take part, same order of the fold).
__device__ __forceinline__ void tail_column_sums128(const float* partials, int n, double* s_buf, int tid) {
    const int grp = tid >> 5, j4 = tid & 31;
    if (grp < 16) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
        for (int r = grp; r < n; r += 16) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(partials + (size_t)r * 128) + j4);   // L2: rows written by other CTAs
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
        double* o = s_buf + grp * 128 + j4 * 4;
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
    }
    __syncthreads();
    if (tid < 128) {
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < 16; ++g) v += s_buf[g * 128 + tid];
        s_buf[tid] = v;   // (each thread only rewrites its own column of group 0)
    }
    __syncthreads();
}

// forward statistics: S[0:64] = sum, S[64:128] = sum of squares over `count` elements per channel -> bnsave, running statistics
__device__ __forceinline__ void bn_forward_finish(const double* S, double count, const BnParams& bn, float* bnsave, int c) {
    const double m = S[c] / count;
    double var = S[64 + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m, invstd = (float)(1.0 / sqrt(var + BN_EPS));
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    bn.running_mean[c] = (float)((1.0 - BN_MOM) * (double)bn.running_mean[c] + BN_MOM * m);
    bn.running_var[c] = (float)((1.0 - BN_MOM) * (double)bn.running_var[c] + BN_MOM * unbiased);
    const float sc = bn.gamma[c] * invstd;
    bnsave[BNS_SCALE + c] = sc;
    bnsave[BNS_SHIFT + c] = bn.beta[c] - mean * sc;
    bnsave[BNS_MEAN + c] = mean;
    bnsave[BNS_INVSTD + c] = invstd;
    bnsave[BNS_VAR + c] = (float)var;   // biased batch variance, for bn_running_update replays
    if (c == 0 && bn.num_batches_tracked != nullptr) *bn.num_batches_tracked += 1;
}

// backward sums: S[0:64] = sum dz, S[64:128] = sum dz * xhat -> coef (c1 | c2), dbeta, dgamma (+=)
__device__ __forceinline__ void bn_backward_finish(const double* S, double count, float* coef, float* dgamma, float* dbeta, int accumulate, int c) {
    const double s1 = S[c], s2 = S[64 + c];
    coef[c] = (float)(s1 / count);
    coef[64 + c] = (float)(s2 / count);
    dbeta[c] = accumulate ? dbeta[c] + (float)s1 : (float)s1;
    dgamma[c] = accumulate ? dgamma[c] + (float)s2 : (float)s2;
}

// Call at the very end of the kernel, by ALL threads of the CTA (>= 128), after this CTA's partial row [128] has been written
// (by any of its threads).  s_buf: 16 * 128 doubles of shared memory that nothing else uses any more (8-byte aligned).
// (noinline: called once at the end of a kernel; keeps its registers out of the producer's allocation)
static __device__ __noinline__ void bn_tail_run(const BnTail& t, const float* partials, double* s_buf, int tid) {
    __shared__ int s_last;
    __threadfence();                 // this thread's part of the CTA's row is visible device-wide ...
    __syncthreads();
    if (tid == 0) {                  // ... before the ticket is drawn
        const unsigned int ticket = atomicAdd(t.counter, 1u);
        s_last = ticket == gridDim.x - 1;
        if (s_last) *t.counter = 0u; // every other CTA has drawn its ticket: ready for the next launch on the stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    tail_column_sums128(partials, (int)gridDim.x, s_buf, tid);
    if (tid < 64) {
        if (t.kind == BnTail::FORWARD) bn_forward_finish(s_buf, t.count, t.bn, t.out0, tid);
        else bn_backward_finish(s_buf, t.count, t.out0, t.out1, t.out2, t.accumulate, tid);
    }
}

}  // namespace srlz

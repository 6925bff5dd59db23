// Forward / inverse model heads and their losses, forward + backward in one call
// (models/forward_inverse.py:21-31,62-70 ; models/models.py:229-237 ; losses/losses.py:102-129).
//   forward model : pred = s + Linear(S+A -> S)([s, onehot(a)])        loss = mean((pred - s')^2)
//   inverse model : logits = Linear(2S -> A)([s, s'])                   loss = CrossEntropy(logits, a)
#include "../../include/srlz.h"
#include "common.cuh"
#include "kernels.h"

namespace srlz {

__device__ __forceinline__ void block_partial_h(float v, float* partials) {
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

// tmp holds s.Wf[:, :S]^T + bf ; diff = tmp + Wf[j][S+a_b] + s - s' ; gp = cf*diff ; gns = -gp
__global__ void __launch_bounds__(256) fwd_loss_kernel(const float* __restrict__ tmp, const float* __restrict__ s,
                                                       const float* __restrict__ ns, const long long* __restrict__ actions,
                                                       const float* __restrict__ wf, int B, int S, int A, float cf,
                                                       float* __restrict__ gp, float* __restrict__ gns,
                                                       float* __restrict__ partials) {
    pdl_enter();
    float acc = 0.f;
    const int n = B * S;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = i / S, j = i % S;
        const long long a = actions[b];
        // an action outside [0, A) would index past the weight row: torch's scatter_ raises there (models/models.py:229-237);
        // on the device the term turns into NaN, which the learner reports (models/learner.py:520-522), and nothing is read
        const float wa = (a >= 0 && a < A) ? wf[(size_t)j * (S + A) + S + a] : __int_as_float(0x7fc00000);
        const float d = tmp[i] + wa + s[i] - ns[i];
        acc = fmaf(d, d, acc);
        const float g = cf * d;
        gp[i] = g;
        gns[i] = -g;
    }
    block_partial_h(acc, partials);
}

// gWf[j][S + a] (+)= sum_{b : a_b == a} gp[b][j]
__global__ void onehot_wgrad_kernel(const float* __restrict__ gp, const long long* __restrict__ actions, int B, int S, int A,
                                    float* __restrict__ gwf, int accumulate) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * A) return;
    const int j = idx / A, a = idx % A;
    float v = 0.f;
    for (int b = 0; b < B; ++b)
        if ((int)actions[b] == a) v += gp[(size_t)b * S + j];
    float* o = gwf + (size_t)j * (S + A) + S + a;
    *o = accumulate ? *o + v : v;
}

// one thread per sample: loss_b = logsumexp(logits) - logits[a] ; glogit = ci * (softmax - onehot)
__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ logits, const long long* __restrict__ actions, int B,
                                                 int A, float ci, float* __restrict__ glogit, float* __restrict__ partials) {
    pdl_enter();
    float acc = 0.f;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float* l = logits + (size_t)b * A;
        float m = l[0];
        for (int a = 1; a < A; ++a) m = fmaxf(m, l[a]);
        float se = 0.f;
        for (int a = 0; a < A; ++a) se += expf(l[a] - m);
        const float lse = m + logf(se);
        const long long t = actions[b];
        acc += (t >= 0 && t < A) ? lse - l[t] : __int_as_float(0x7fc00000);   // out-of-range target: NaN loss, no out-of-bounds read
        for (int a = 0; a < A; ++a) glogit[(size_t)b * A + a] = ci * (expf(l[a] - lse) - (a == t ? 1.f : 0.f));
    }
    block_partial_h(acc, partials);
}

// dmu = coef*mu ; dlogvar = coef*0.5*(exp(logvar)-1)      (gradient of coef * KL, losses/losses.py:253)
__global__ void kl_grad_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int n, float coef,
                               float* __restrict__ dmu, float* __restrict__ dlv) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        dmu[i] = coef * mu[i];
        dlv[i] = coef * 0.5f * (expf(logvar[i]) - 1.f);
    }
}

__global__ void __launch_bounds__(256) kl_sum_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int n,
                                                     float* __restrict__ partials) {
    pdl_enter();
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float m = mu[i], lv = logvar[i];
        s += 1.f + lv - m * m - expf(lv);
    }
    block_partial_h(s, partials);
}

// ---- small elementwise pieces of the reference's module API outside the fused step (mlp heads, split models) ----
__global__ void relu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    pdl_enter();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaxf(x[i], 0.f);
}
__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx, long long n) {
    pdl_enter();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gx[i] = y[i] > 0.f ? gy[i] : 0.f;
}
__global__ void colmask_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y, int rows, int cols) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) y[i] = x[i] * mask[i % cols];
}
// out[r] = [a[r, 0:ca] | b[r, 0:cb]]  or, with idx, [a[r] | onehot(idx[r], cb)]
__global__ void cat_cols_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb,
                                const long long* __restrict__ idx, float* __restrict__ out, int rows) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, w = ca + cb;
    if (i >= rows * w) return;
    const int r = i / w, c = i % w;
    float v;
    if (c < ca) v = a[(size_t)r * ca + c];
    else if (idx != nullptr) {
        const long long t = idx[r];
        v = (t >= 0 && t < cb) ? (t == c - ca ? 1.f : 0.f) : __int_as_float(0x7fc00000);   // scatter_ raises on a bad index
    } else v = b[(size_t)r * cb + (c - ca)];
    out[i] = v;
}
__global__ void reparam_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                               float* __restrict__ z, int n) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = fmaf(eps[i], expf(0.5f * lv[i]), mu[i]);
}
__global__ void reparam_bwd_kernel(const float* __restrict__ gz, const float* __restrict__ lv, const float* __restrict__ eps,
                                   float* __restrict__ gmu, float* __restrict__ glv, int n) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        gmu[i] = gz[i];
        glv[i] = gz[i] * eps[i] * 0.5f * expf(0.5f * lv[i]);
    }
}

static size_t align64f(size_t n) { return (n + 63) / 64 * 64; }

}  // namespace srlz

using namespace srlz;

#define RC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

extern "C" {

size_t srlz_heads_workspace_bytes(int B, int state_dim, int action_dim) {
    return (align64f((size_t)B * state_dim) * 2 + align64f((size_t)B * action_dim) * 2 + SRLZ_MAX_PART) * sizeof(float);
}

/* out[0] = -0.5 * sum(1 + logvar - mu^2 - exp(logvar))   (losses/losses.py:253); workspace >= 1184 floats */
int srlz_kl(const float* mu, const float* logvar, int n, float* out, void* workspace, void* stream) {
    if (mu == nullptr || logvar == nullptr || out == nullptr || workspace == nullptr || n <= 0) { set_error("srlz_kl: bad argument"); return SRLZ_E_ARG; }
    int gx = (n + 255) / 256;
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    float* part = reinterpret_cast<float*>(workspace);
    launch_k(kl_sum_kernel, gx, 256, 0, (cudaStream_t)stream, mu, logvar, n, part);
    RC(check_launch("kl_sum"));
    return sum_partials(part, gx, -0.5f, out, 0, (cudaStream_t)stream);
}

int srlz_kl_grad(const float* mu, const float* logvar, int n, float coef, float* dmu, float* dlogvar, void* stream) {
    if (mu == nullptr || logvar == nullptr || dmu == nullptr || dlogvar == nullptr || n <= 0) { set_error("srlz_kl_grad: bad argument"); return SRLZ_E_ARG; }
    launch_k(kl_grad_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, mu, logvar, n, coef, dmu, dlogvar);
    return check_launch("kl_grad");
}

/* out[0] = mean_b CrossEntropy(logits[b], actions[b]) ; glogit = (softmax - onehot)/B  (losses/losses.py:126-127) */
int srlz_cross_entropy(const float* logits, const int64_t* actions, int B, int A, float* out, float* glogit, void* workspace,
                       void* stream) {
    if (logits == nullptr || actions == nullptr || out == nullptr || glogit == nullptr || workspace == nullptr || B <= 0 || A <= 0) {
        set_error("srlz_cross_entropy: bad argument");
        return SRLZ_E_ARG;
    }
    int gx = (B + 255) / 256;
    if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
    float* part = reinterpret_cast<float*>(workspace);
    launch_k(ce_kernel, gx, 256, 0, (cudaStream_t)stream, logits, reinterpret_cast<const long long*>(actions), B, A, 1.f / (float)B, glogit, part);
    RC(check_launch("ce"));
    return sum_partials(part, gx, 1.f / (float)B, out, 0, (cudaStream_t)stream);
}

/* nn.ReLU of the mlp heads (models/forward_inverse.py:50-56,79-83) and its backward (from the OUTPUT y) */
int srlz_relu(const float* x, float* y, int64_t n, void* stream) {
    if (x == nullptr || y == nullptr || n <= 0) { set_error("srlz_relu: bad argument"); return SRLZ_E_ARG; }
    launch_k(relu_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, x, y, n);
    return check_launch("relu");
}
int srlz_relu_bwd(const float* y, const float* gy, float* gx, int64_t n, void* stream) {
    if (y == nullptr || gy == nullptr || gx == nullptr || n <= 0) { set_error("srlz_relu_bwd: bad argument"); return SRLZ_E_ARG; }
    launch_k(relu_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, y, gy, gx, n);
    return check_launch("relu_bwd");
}
/* y[r, c] = x[r, c] * mask[c]: SRLModulesSplit.detachSplit (models/modules.py:189-234) is a 0/1 column mask; its backward is the same op */
int srlz_colmask(const float* x, const float* mask, float* y, int rows, int cols, void* stream) {
    if (x == nullptr || mask == nullptr || y == nullptr || rows <= 0 || cols <= 0) { set_error("srlz_colmask: bad argument"); return SRLZ_E_ARG; }
    launch_k(colmask_kernel, (rows * cols + 255) / 256, 256, 0, (cudaStream_t)stream, x, mask, y, rows, cols);
    return check_launch("colmask");
}
/* th.cat((a, b), dim=1), or th.cat((a, encodeOneHot(idx, cb)), dim=1) when idx is given (models/models.py:229-237) */
int srlz_cat_cols(const float* a, int ca, const float* b, int cb, const int64_t* idx, float* out, int rows, void* stream) {
    if (a == nullptr || out == nullptr || (b == nullptr && idx == nullptr) || rows <= 0 || ca <= 0 || cb <= 0) { set_error("srlz_cat_cols: bad argument"); return SRLZ_E_ARG; }
    launch_k(cat_cols_kernel, (rows * (ca + cb) + 255) / 256, 256, 0, (cudaStream_t)stream, a, ca, b, cb, reinterpret_cast<const long long*>(idx), out, rows);
    return check_launch("cat_cols");
}
/* z = eps * exp(0.5 * logvar) + mu (models/models.py:155-163) and its backward */
int srlz_reparam(const float* mu, const float* logvar, const float* eps, float* z, int n, void* stream) {
    if (mu == nullptr || logvar == nullptr || eps == nullptr || z == nullptr || n <= 0) { set_error("srlz_reparam: bad argument"); return SRLZ_E_ARG; }
    launch_k(reparam_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, mu, logvar, eps, z, n);
    return check_launch("reparam");
}
int srlz_reparam_bwd(const float* gz, const float* logvar, const float* eps, float* gmu, float* glogvar, int n, void* stream) {
    if (gz == nullptr || logvar == nullptr || eps == nullptr || gmu == nullptr || glogvar == nullptr || n <= 0) { set_error("srlz_reparam_bwd: bad argument"); return SRLZ_E_ARG; }
    launch_k(reparam_bwd_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, gz, logvar, eps, gmu, glogvar, n);
    return check_launch("reparam_bwd");
}

int srlz_heads(const float* s, const float* ns, const int64_t* actions, int B, int norm_batch, int state_dim, int action_dim,
               const float* fwd_w, const float* fwd_b, const float* inv_w, const float* inv_b, float w_fwd, float w_inv,
               float* loss_out, float* gs, float* gns, float* g_fwd_w, float* g_fwd_b, float* g_inv_w, float* g_inv_b,
               int accumulate, void* workspace, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int S = state_dim, A = action_dim;
    if (s == nullptr || ns == nullptr || actions == nullptr || gs == nullptr || gns == nullptr || workspace == nullptr || B <= 0) {
        set_error("srlz_heads: null argument or B <= 0");
        return SRLZ_E_ARG;
    }
    if (norm_batch <= 0) norm_batch = B;
    float* tmp = reinterpret_cast<float*>(workspace);
    float* gp = tmp + align64f((size_t)B * S);
    float* logits = gp + align64f((size_t)B * S);
    float* glogit = logits + align64f((size_t)B * A);
    float* partials = glogit + align64f((size_t)B * A);
    const long long* act = reinterpret_cast<const long long*>(actions);
    const size_t bytesBS = (size_t)B * S * sizeof(float);
    cudaMemsetAsync(gs, 0, bytesBS, st);
    cudaMemsetAsync(gns, 0, bytesBS, st);
    if (loss_out != nullptr) cudaMemsetAsync(loss_out, 0, 2 * sizeof(float), st);
    if (w_fwd != 0.f) {
        if (fwd_w == nullptr || fwd_b == nullptr || g_fwd_w == nullptr || g_fwd_b == nullptr) { set_error("srlz_heads: forward head pointers missing"); return SRLZ_E_ARG; }
        RC(sgemm(s, S, 1, fwd_w, 1, S + A, tmp, S, 1, fwd_b, B, S, S, 0, st));
        const float cf = w_fwd * 2.f / ((float)norm_batch * (float)S);
        int gx = (B * S + 255) / 256;
        if (gx > SRLZ_MAX_PART) gx = SRLZ_MAX_PART;
        launch_k(fwd_loss_kernel, gx, 256, 0, st, tmp, s, ns, act, fwd_w, B, S, A, cf, gp, gns, partials);
        RC(check_launch("fwd_loss"));
        if (loss_out != nullptr) RC(sum_partials(partials, gx, 1.f / ((float)norm_batch * (float)S), loss_out, 0, st));
        cudaMemcpyAsync(gs, gp, bytesBS, cudaMemcpyDeviceToDevice, st);
        RC(sgemm(gp, S, 1, fwd_w, S + A, 1, gs, S, 1, nullptr, B, S, S, 1, st));
        RC(sgemm(gp, 1, S, s, S, 1, g_fwd_w, S + A, 1, nullptr, S, S, B, accumulate, st));
        launch_k(onehot_wgrad_kernel, (S * A + 127) / 128, 128, 0, st, gp, act, B, S, A, g_fwd_w, accumulate);
        RC(check_launch("onehot_wgrad"));
        RC(colsum(gp, B, S, g_fwd_b, accumulate, st));
    }
    if (w_inv != 0.f) {
        if (inv_w == nullptr || inv_b == nullptr || g_inv_w == nullptr || g_inv_b == nullptr) { set_error("srlz_heads: inverse head pointers missing"); return SRLZ_E_ARG; }
        RC(sgemm(s, S, 1, inv_w, 1, 2 * S, logits, A, 1, inv_b, B, A, S, 0, st));
        RC(sgemm(ns, S, 1, inv_w + S, 1, 2 * S, logits, A, 1, nullptr, B, A, S, 1, st));
        int gx = (B + 255) / 256;
        launch_k(ce_kernel, gx, 256, 0, st, logits, act, B, A, w_inv / (float)norm_batch, glogit, partials);
        RC(check_launch("ce"));
        if (loss_out != nullptr) RC(sum_partials(partials, gx, 1.f / (float)norm_batch, loss_out + 1, 0, st));
        RC(sgemm(glogit, A, 1, inv_w, 2 * S, 1, gs, S, 1, nullptr, B, S, A, 1, st));
        RC(sgemm(glogit, A, 1, inv_w + S, 2 * S, 1, gns, S, 1, nullptr, B, S, A, 1, st));
        RC(sgemm(glogit, 1, A, s, S, 1, g_inv_w, 2 * S, 1, nullptr, A, S, B, accumulate, st));
        RC(sgemm(glogit, 1, A, ns, S, 1, g_inv_w + S, 2 * S, 1, nullptr, A, S, B, accumulate, st));
        RC(colsum(glogit, B, A, g_inv_b, accumulate, st));
    }
    return 0;
}

}  // extern "C"

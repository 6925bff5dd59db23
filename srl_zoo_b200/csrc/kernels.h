// Internal launcher prototypes of libsrlz (host side).  All launchers enqueue on the given stream,
// never synchronise, never allocate; they return 0 or a non-zero error code (message via set_error).
#pragma once
#include "common.cuh"

namespace srlz {

enum { EPI_PLAIN = 0, EPI_STATS = 1, EPI_MASK_BNBWD = 2, EPI_DEC12 = 3 };

struct BnParams {
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* num_batches_tracked;
};
// partials [n][128] (sum, sumsq) -> scale/shift (+ mean/invstd saved), running stats updated when training
// bnsave (5 x 64 floats): scale | shift | mean | invstd | biased batch variance
#define BNS_SCALE 0
#define BNS_SHIFT 64
#define BNS_MEAN 128
#define BNS_INVSTD 192
#define BNS_VAR 256
#define BNS_FLOATS 320

// Fused finalize of a per-CTA partial-sum reduction by the last CTA of the producing kernel (bn_tail.cuh).  counter == nullptr:
// the kernel only writes its partial rows and the caller launches bn_finalize / bn_bwd_finalize itself.
struct BnTail {
    enum { FORWARD = 1, BACKWARD = 2 };
    unsigned int* counter;   // zero before the launch (the last CTA resets it)
    int kind;
    int accumulate;          // BACKWARD: add to dgamma / dbeta
    double count;            // elements per channel
    BnParams bn;             // FORWARD
    float* out0;             // FORWARD: bnsave (BNS_FLOATS) | BACKWARD: coef (128)
    float* out1;             // BACKWARD: dgamma
    float* out2;             // BACKWARD: dbeta
};

struct GConvArgs {
    const float* in;        // gathered tensor, NHWC C=64
    const float* bias;      // [64] or null
    const float* in_scale;  // BN+ReLU applied to `in` on load (null = plain)
    const float* in_shift;
    float* out;             // NHWC C=64
    const float* e_ypre;    // EPI_MASK_BNBWD: pre-BN activation at the output location
    const float* e_scale;
    const float* e_shift;
    const float* e_mean;
    const float* e_invstd;
    float* partials;        // [n_partials][128]  (sum / sum-sq   or   sum dz / sum dz*xhat)
    BnTail tail;            // optional fused finalize of those partials
    ConvGeom g;
    int transposed;
    int epi;
    long long* dbg;         // optional timeline buffer (clock64 stamps of CTA 0), tests only
    // special producer of the per-tap tcgen05 kernel (conv_tc.cu): mode 2 = dec12 dgrad columns (48 = 3x4x4 values of
    // d(decoded) per pixel); mode 1 marks the first layer's NCHW observation input for enc0_rows_fwd
    int mode;
    const int* rects;       // mode 1: (B,4) occlusion rectangles or null
    const float* aux0;      // mode 2: explicit d(decoded) (B,3,224,224) or null
    const float* aux1;      // mode 2: decoded
    const float* aux2;      // mode 2: target
    float coef;             // mode 2: d(decoded) = coef * (decoded - target) when aux0 is null
};
// per-tap tcgen05 pipeline (conv_tc.cu): weights as the bf16 hi/lo image written by pack_conv_w_bf16
int gconv64_tc(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st);
// halo-tile tcgen05 version (conv_halo_tc.cu) for stride-1 / per-parity-class geometries
bool gconv64_halo_supported(const GConvArgs& a);
int gconv64_halo(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st);
int pack_conv_w_bf16(const float* pack_f32, void* dst, int ntaps, cudaStream_t st);
// the six 3x3 64->64 layers (encoder_conv.{4,8}, decoder_conv.{0,3,6,9}) in one launch: torch weights -> forward + dgrad images
struct ConvPackJobs {
    const float* w[6];
    unsigned char* fwd[6];
    unsigned char* dgrad[6];
    int transposed[6];
};
int pack_conv_layers_bf16(const ConvPackJobs& jobs, cudaStream_t st);
// row kernel for the stride-2 gathers (dgrad of the transposed convolutions; dgrad_s2_rows_tc.cu): every dy row staged once
bool gconv64_s2rows_supported(const GConvArgs& a);
int gconv64_s2rows(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st);
// decoder_conv.12 forward (dec12_rows_tc.cu): a.in = pre-BN input (B,111,111,64) with a.in_scale/in_shift, a.bias = (3),
// a.out = decoded (B,3,224,224) NCHW, a.aux2 = target or null, a.partials = per-CTA squared-error partials or null;
// weights image (16 KB) from pack_dec12_fwd_bf16; every input row staged once per CTA
int pack_dec12_fwd_bf16(const float* w12, void* dst, cudaStream_t st);
int dec12_rows_fwd(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st);
struct GWgradArgs;
// row-staged wgrad of the same layer (a.small = pre-BN input + dense_scale/shift, aux0 / aux1, aux2, coef = the gradient source)
int dec12_rows_wgrad(const GWgradArgs& a, float* grad_out, float* grad_bias, int accumulate, cudaStream_t st);   // + bias gradient
size_t dec12_rows_wgrad_partial_floats();
// fp32 [tap][k][n] staging pack for the dec12 dgrad producer (then pack_conv_w_bf16): 1 chunk
int pack_dec12_dgrad(const float* w12, float* pack1, cudaStream_t st);
// row-image tcgen05 kernels of the first encoder layer (enc0_rows_tc.cu): a.in = NCHW observation, a.rects = DAE rectangles
int enc0_rows_fwd(const GConvArgs& a, const void* wbf, int* n_partials, cudaStream_t st);
int pack_enc0_rows_bf16(const float* w0, void* dst, cudaStream_t st);   // 64 KB image for enc0_rows_fwd
struct GWgradArgs;
int enc0_rows_wgrad(const GWgradArgs& a, float* grad_out, int accumulate, cudaStream_t st);   // big = observation, small = dy
size_t enc0_rows_wgrad_partial_floats();
#define SRLZ_WBF_FLOATS (9 * 4096)  // bytes of one 9-tap bf16 hi/lo image = 9 * 16 KB = 36864 floats

struct GWgradArgs {
    const float* big;          // gathered side, NHWC C=64
    const float* small;        // dense side, NHWC C=64
    const float* dense_scale;  // BN+ReLU on load of the dense side (null = plain)
    const float* dense_shift;
    float* partials;           // workspace: gwgrad64_partial_floats()
    ConvGeom g;
    int chunk_len;
    // special gathered sides of the tcgen05 kernel (wgrad_tc.cu): mode 1 = enc0 im2col chunks of the NCHW observation
    // (`big` = observation), mode 2 = dec12 columns of d(decoded) (`small` = pre-BN input of the layer)
    int mode;
    const int* rects;
    const float* aux0;      // mode 2: explicit d(decoded) or null
    const float* aux1;      // mode 2: decoded
    const float* aux2;      // mode 2: target
    float coef;
    long long* dbg;         // optional timeline buffer (tests only)
};
size_t gwgrad64_partial_floats(const ConvGeom& g);
// halo-tile tcgen05 wgrad of the 3x3 layers (wgrad_halo_tc.cu): every pixel converted once per tile, taps = row-shifted descriptors
bool gwgrad64_halo_supported(const ConvGeom& g);
int gwgrad64_halo(const GWgradArgs& a, float* grad_out, int accumulate, cudaStream_t st);

// ---- input pipeline (preprocess.cu) ----
// uint8 RGB frames (B,224,224,3) HWC -> normalised fp32 (B,3,224(W),224(H)), bit-exact with the reference's host arithmetic
int preprocess_u8(const unsigned char* frames, float* out, int B, cudaStream_t st);

// ---- BatchNorm / pooling ----   (BnParams, BnTail, bnsave layout: top of this file)
int bn_finalize(const float* partials, int n_partials, long long count, const BnParams& bn, int training,
                float* bnsave, cudaStream_t st);
// replays the running-stat update of a finished training forward (VAE getStates passes, learner.py:402)
int bn_running_update(const float* bnsave, long long count, const BnParams& bn, cudaStream_t st);
int bn_relu_pool_fwd(const float* y, const float* scale, const float* shift, float* out, unsigned char* argmax,
                     int B, int H, int W, int PH, int PW, int pad, cudaStream_t st);
// TMA-fed version (pool_tma.cu): input rows streamed through a shared-memory ring by cp.async.bulk.tensor
bool bn_relu_pool_fwd_tma_supported(const float* y, int W);
int bn_relu_pool_fwd_tma(const float* y, const float* scale, const float* shift, float* out, unsigned char* argmax, int B, int H, int W,
                         int PH, int PW, int pad, cudaStream_t st);
// BN-backward statistics of a pooled stage from the pooled side (dpool and the pooled activation a); with
// pool_bwd_bn_apply this is the product path: one full-size pass instead of two
int pool_bwd_stats(const float* dpool, const float* a, const unsigned char* argmax, const float* y, const float* gamma,
                   const float* beta, const float* mean, const float* invstd, float* partials, int* n_partials, int B, int H,
                   int W, int PH, int PW, int pad, cudaStream_t st);
// MaxPool + ReLU + BatchNorm backward of a pooled stage in one full-size pass (coef from pool_bwd_stats + bn_bwd_finalize)
int pool_bwd_bn_apply(const float* dpool, const unsigned char* argmax, const float* y, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const float* gamma, const float* coef, float* dy, int B, int H,
                      int W, int PH, int PW, int pad, cudaStream_t st);
// partials [n][128] (sum dz, sum dz*xhat) -> coef[0:64]=c1, coef[64:128]=c2 ; dgamma, dbeta (+=)
int bn_bwd_finalize(const float* partials, int n_partials, long long count, float* coef, float* dgamma,
                    float* dbeta, int accumulate, cudaStream_t st);
// dy = gamma*invstd*(dz - c1 - xhat*c2) in place over dz ; optional per-channel sum of dy -> dbias (+=)
int bn_bwd_apply(float* dz, const float* y, const float* gamma, const float* mean, const float* invstd,
                 const float* coef, long long npix, float* dbias, float* partials, int accumulate, cudaStream_t st);

// ---- dense / elementwise ----
// C[i,j] = sum_k A(i,k)*B(k,j) (+bias[j]) with arbitrary element strides; beta: C = acc ? C + .. : ..
int sgemm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
          long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, cudaStream_t st);
int sgemm_splitk(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
                 long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, float* ws,
                 size_t ws_floats, cudaStream_t st);
// perm: 1 = C's row index, 2 = C's column index runs over the 2304 bottleneck features in NHWC order and is stored at the torch position
int sgemm_perm(const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
               long long sci, long long scj, const float* bias, int M, int N, int K, int accumulate, int perm, cudaStream_t st);
int colsum(const float* A, int M, int N, float* out, int accumulate, cudaStream_t st);  // out[j] = sum_i A[i,j]
int colsum_perm(const float* A, int M, int N, float* out, int accumulate, int perm, cudaStream_t st);
int vae_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, float* kl_partials, int n,
                    int training, int* n_partials, cudaStream_t st);
int vae_reparam_bwd(const float* dz, const float* logvar, const float* eps, const float* gmu_extra,
                    const float* glogvar_extra, float kl_coef, const float* mu, float* dmu, float* dlogvar, int n,
                    int training, cudaStream_t st);
int sse_partials(const float* a, const float* b, long long n, float* partials, int* n_partials, cudaStream_t st);
int sum_partials(const float* partials, int n, float scale, float* out, int accumulate, cudaStream_t st);
int mse_grad(const float* a, const float* b, long long n, float coef, float* g, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              float bc1, float bc2, cudaStream_t st);
int permute_fc(const float* src, float* dst, int rows, int to_packed, int row_mode, int accumulate, cudaStream_t st);
int pack_conv_w(const float* w, float* fwd_pack, float* dgrad_pack, int ntaps, int transposed_conv, cudaStream_t st);
// eval-mode BatchNorm folded into the conv before it: w' = w * s[co], b' = beta - mean * s, s = gamma / sqrt(var + eps)
int fold_bn(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float* w_out, float* b_out,
            int per_co, cudaStream_t st);
int fill(float* p, float v, int n, cudaStream_t st);

}  // namespace srlz

/*
 * libsrlz -- C ABI of the B200-native train-step path for araffin/srl-zoo's conv autoencoder / beta-VAE /
 * denoising autoencoder (the drop-in boundary).
 *
 * Conventions
 *   - plain C, no C++ or torch types.  Every pointer is a CUDA DEVICE pointer unless stated otherwise;
 *     tensors are fp32 and contiguous in the torch-native layout of the reference module they replace.
 *   - the library never allocates, frees or retains device memory: outputs, the `saved` block (activations
 *     kept for backward) and the `workspace` block are caller-allocated (sizes from srlz_*_bytes()).
 *   - every call only ENQUEUES work on the stream passed in (cudaStream_t as void*); no hidden sync.
 *   - return value 0 = ok; otherwise an error code, message via srlz_last_error() (thread-local).
 *
 * Reference interfaces replaced (araffin/srl-zoo, file:line relative to the reference root):
 *   srlz_forward   <- SRLModules.forward / getStates       models/modules.py:75-85
 *                     BaseModelAutoEncoder.forward          models/models.py:106-114  (encoder_conv :47-63,
 *                     decoder_conv :65-83), CNNAutoEncoder.encode/decode models/autoencoders.py:102-118,
 *                     BaseModelVAE.forward/reparameterize   models/models.py:147-176, CNNVAE models/vae.py:59-75,
 *                     DAE mask preprocessing/data_loader.py:55-63, fused reconstruction / generation /
 *                     KL sums losses/losses.py:172-214,239-256
 *   srlz_backward  <- loss.backward() through the same modules   models/learner.py:489
 *   srlz_heads_*   <- forwardModel / inverseModel + their losses  models/forward_inverse.py:21-31,62-70,
 *                     losses/losses.py:102-129
 *   srlz_adam_step <- th.optim.Adam(...).step()                  models/learner.py:199,495
 */
#ifndef SRLZ_H_
#define SRLZ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRLZ_VERSION 1
#define SRLZ_E_ARG 1001   /* bad argument / shape */
#define SRLZ_E_CUDA 1002  /* CUDA launch or runtime error */
#define SRLZ_MAX_BATCH 2048 /* images per model call: the largest pre-BN tensor (B x 112 x 112 x 64) keeps 32-bit element indices;
                               larger minibatches are split by the caller (the reference's largest per-GPU minibatch is 256) */

typedef struct srlz_bn {  /* nn.BatchNorm2d(64) */
    const float* weight;
    const float* bias;
    float* running_mean;
    float* running_var;
    int64_t* num_batches_tracked;
} srlz_bn;

/* Parameters + buffers of SRLModules.model (CNNAutoEncoder or CNNVAE), torch-native layouts:
 * Conv2d OIHW, ConvTranspose2d IOHW, Linear (out,in). */
typedef struct srlz_net {
    int32_t is_vae;
    int32_t state_dim;
    const float* enc_w[3];      /* encoder_conv.{0,4,8}.weight : (64,3,7,7) (64,64,3,3) (64,64,3,3) */
    srlz_bn enc_bn[3];          /* encoder_conv.{1,5,9} */
    const float* dec_w[5];      /* decoder_conv.{0,3,6,9,12}.weight : 4 x (64,64,3,3), (64,3,4,4) */
    const float* dec_b[5];      /* decoder_conv.{0,3,6,9,12}.bias */
    srlz_bn dec_bn[4];          /* decoder_conv.{1,4,7,10} */
    const float* fc_enc_w[2];   /* AE: encoder_fc.0 ; VAE: encoder_fc1 (mu), encoder_fc2 (logvar) : (S,2304) */
    const float* fc_enc_b[2];
    const float* fc_dec_w;      /* decoder_fc.0.weight (2304,S) */
    const float* fc_dec_b;
} srlz_net;

/* Gradient destinations, same tensors / layouts as srlz_net. */
typedef struct srlz_net_grads {
    float* enc_w[3];
    float* enc_bn_w[3];
    float* enc_bn_b[3];
    float* dec_w[5];
    float* dec_b[5];
    float* dec_bn_w[4];
    float* dec_bn_b[4];
    float* fc_enc_w[2];
    float* fc_enc_b[2];
    float* fc_dec_w;
    float* fc_dec_b;
} srlz_net_grads;

int srlz_version(void);
const char* srlz_last_error(void);

/* sizes (bytes / floats) of the caller-allocated blocks for a per-call batch of B images */
size_t srlz_pack_floats(int is_vae, int state_dim);
size_t srlz_saved_bytes(int B, int state_dim, int is_vae);
size_t srlz_workspace_bytes(int B, int state_dim, int is_vae);
/* byte offsets of the tensors inside `saved` (for tests / introspection); returns the number written.
 * order: y1,a1,am1,y2,a2,am2,y3,a3,am3,lat,z,d0,y4,y5,y6,y7,bnsave  (see srlz_saved_names) */
int srlz_saved_layout(int B, int state_dim, int is_vae, size_t* offsets, int max_entries);
const char* srlz_saved_names(void);

/* torch-native weights -> kernel layouts ([tap][cin][cout] packs, NHWC-permuted fc).  Once per step. */
int srlz_pack_weights(const srlz_net* net, float* wpack, void* stream);

/* One model call (models/modules.py:82-85).
 *  x        (B,3,224,224) observation; rects (B,4) int32 (h1,h2,w1,w2) or NULL: DAE mask applied on load.
 *  eps      (B,S) N(0,1) draw of models/models.py:161 (VAE, training) or NULL.
 *  training 1: batch statistics + running-stat update ; 0: running statistics (model.eval()).
 *  lat      out: AE (B,S) encoded states ; VAE (B,2S)? no -> mu in `lat`, logvar in `logvar`.
 *  decoded  out (B,3,224,224) or NULL for an encoder-only pass (getStates).
 *  target   optional (B,3,224,224): loss_out[0] = sum (decoded-target)^2 (fused in the last decoder tile).
 *  loss_out [2] floats: [0] SSE (if target), [1] VAE: sum(1 + logvar - mu^2 - exp(logvar)) (KL = -0.5 * it).
 */
int srlz_forward(const srlz_net* net, const float* wpack, const float* x, const int32_t* rects, const float* eps,
                 int B, int training, float* lat, float* logvar, float* decoded, const float* target, float* loss_out,
                 void* saved, void* workspace, void* stream);

/* Replays the BatchNorm running-stat update of the forward whose `saved` block is given
 * (the reference's extra train-mode getStates() passes, models/learner.py:402). */
int srlz_replay_running_stats(const srlz_net* net, int B, void* saved, void* stream);

/* Backward of one model call.  Gradient w.r.t. decoded is either explicit (g_decoded) or, when g_decoded is
 * NULL and decoded/target are given, mse_coef*(decoded-target).  g_lat / g_logvar: optional upstream gradients
 * w.r.t. the encoded states (AE) / mu and logvar (VAE); kl_coef: d(total)/d(KL) (VAE).
 * accumulate != 0: parameter gradients are added to the destination instead of overwriting it.
 * has_decoder = 0: encoder-only call (getStates).
 * Alignment: x 8 bytes; g_decoded / decoded / target 16 bytes (they are fetched in 16-byte chunks).
 * All kernels are launched with programmatic stream serialization and wait for the preceding kernel of `stream`
 * before their first global-memory access: stream ordering is exactly that of plain launches. */
int srlz_backward(const srlz_net* net, const float* wpack, const srlz_net_grads* grads, int accumulate, const float* x,
                  const int32_t* rects, const float* eps, int B, int training, int has_decoder, const float* g_decoded,
                  const float* decoded, const float* target, float mse_coef, const float* g_lat, const float* g_logvar,
                  float kl_coef, void* saved, void* workspace, void* stream);

/* Inference path (BaseLearner._predFn / predStatesWithDataLoader, models/learner.py:67-88,570-577; evaluation/predict_dataset.py:
 * 36-47): eval-mode getStates with every encoder BatchNorm (running statistics) FOLDED into the conv before it, so a call is
 * conv -> (+bias) ReLU MaxPool three times and one dense layer: no statistics, no normalisation constants, nothing kept for a
 * backward.  srlz_eval_pack folds + packs the weights ONCE per weight version into `epack` (srlz_eval_pack_floats floats);
 * srlz_encode_eval then needs 7 launches per batch.  states (B,S): AE encoded states / VAE mu.  rects: optional DAE mask. */
size_t srlz_eval_pack_floats(int state_dim);
size_t srlz_eval_workspace_bytes(int B, int state_dim);
int srlz_eval_pack(const srlz_net* net, float* epack, void* stream);
int srlz_encode_eval(const srlz_net* net, const float* epack, const float* x, const int32_t* rects, int B, float* states,
                     void* workspace, void* stream);

/* Decoder-only call and its backward: BaseModelAutoEncoder.decode / BaseModelVAE.decode on a given latent (models/
 * autoencoders.py:111-118, models/vae.py:68-75; called directly by evaluation/enjoy_latent.py:35,136 and, with a masked latent, by
 * SRLModulesSplit.forwardAutoencoder / forwardVAE, models/modules.py:236-258).  z (B,S) -> decoded (B,3,224,224); the backward
 * takes d(decoded) and produces the decoder's parameter gradients and g_z (B,S).  Same `saved` / `workspace` blocks as srlz_forward. */
int srlz_decode(const srlz_net* net, const float* wpack, const float* z, int B, int training, float* decoded, void* saved,
                void* workspace, void* stream);
int srlz_decode_backward(const srlz_net* net, const float* wpack, const srlz_net_grads* grads, int accumulate, int B, int training,
                         const float* g_decoded, float* g_z, void* saved, void* workspace, void* stream);

/* small elementwise pieces of the reference's module API outside the fused step: nn.ReLU of the mlp inverse / reward heads
 * (models/forward_inverse.py:50-56,79-83) and its backward from the output; the 0/1 column mask SRLModulesSplit.detachSplit
 * amounts to (models/modules.py:189-234; backward = the same op); th.cat((a, b), 1) / th.cat((a, encodeOneHot(idx, cb)), 1)
 * (models/models.py:229-237, models/forward_inverse.py:30,70); the VAE reparameterisation z = eps*exp(logvar/2) + mu and its
 * backward (models/models.py:155-163) */
int srlz_relu(const float* x, float* y, int64_t n, void* stream);
int srlz_relu_bwd(const float* y, const float* gy, float* gx, int64_t n, void* stream);
int srlz_colmask(const float* x, const float* mask, float* y, int rows, int cols, void* stream);
int srlz_cat_cols(const float* a, int ca, const float* b, int cb, const int64_t* idx, float* out, int rows, void* stream);
int srlz_reparam(const float* mu, const float* logvar, const float* eps, float* z, int n, void* stream);
int srlz_reparam_bwd(const float* gz, const float* logvar, const float* eps, float* gmu, float* glogvar, int n, void* stream);

/* forward / inverse model heads + their losses (models/forward_inverse.py:21-31,62-70; losses/losses.py:102-129).
 *  loss_out[0] = mean((s + W_f [s, onehot(a)] + b_f - s')^2), loss_out[1] = CrossEntropy(W_i [s, s'] + b_i, a).
 *  w_fwd / w_inv: loss weights (0 disables the term); gradients w.r.t. s, s' and the head parameters are produced
 *  in the same call (gs, gns: (B,S) overwritten). */
size_t srlz_heads_workspace_bytes(int B, int state_dim, int action_dim);
/* norm_batch: the batch size the mean-type losses are normalised by (global batch under data parallelism; <= 0: B) */
int srlz_heads(const float* s, const float* ns, const int64_t* actions, int B, int norm_batch, int state_dim, int action_dim,
               const float* fwd_w, const float* fwd_b, const float* inv_w, const float* inv_b, float w_fwd, float w_inv,
               float* loss_out, float* gs, float* gns, float* g_fwd_w, float* g_fwd_b, float* g_inv_w, float* g_inv_b,
               int accumulate, void* workspace, void* stream);

/* stand-alone loss kernels for the reference's loss-function API (losses/losses.py:126-127,253); workspace >= 1184 floats */
int srlz_kl(const float* mu, const float* logvar, int n, float* out, void* workspace, void* stream);
int srlz_kl_grad(const float* mu, const float* logvar, int n, float coef, float* dmu, float* dlogvar, void* stream);
int srlz_cross_entropy(const float* logits, const int64_t* actions, int B, int A, float* out, float* glogit, void* workspace,
                       void* stream);

/* Input pipeline hand-over in uint8 (preprocessing/data_loader.py:38-65,247-256 and preprocessing/utils.py:20-32): `frames` =
 * B RGB uint8 images (B,224,224,3) in the loader's native HWC order (after its cv2.resize / cvtColor); `out` = the
 * (B,3,224,224) float32 tensor the reference's loader delivers ((x/255 - mean)/std in fp32, dims (C, W, H)), bit-exact.
 * The DAE rectangle is not applied here: srlz_forward / srlz_backward zero it on load from `rects`. */
int srlz_preprocess_u8(const uint8_t* frames, float* out, int B, void* stream);

/* sum (a-b)^2 -> out[0] (out_scale applied) ; grad: g = coef*(a-b) */
int srlz_sse(const float* a, const float* b, int64_t n, float out_scale, float* out, void* workspace, void* stream);
int srlz_mse_grad(const float* a, const float* b, int64_t n, float coef, float* g, void* stream);

/* fused Adam over a flat buffer (torch.optim.Adam semantics, models/learner.py:199): step >= 1 */
int srlz_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int step, void* stream);

/* per-call-site device timing with CUDA events on the launching stream (bench.py roofline line).
 * srlz_prof_report synchronises, writes "tag launches total_ms" lines and resets the recorder. */
long long srlz_launch_count(void);  /* kernels launched by this library in this process */
void srlz_prof_enable(int on);
int srlz_prof_report(char* buf, int buf_len);

/* op-level entry points (unit tests of the product kernels at layer granularity)
 *  srlz_op_conv64: forward / dgrad of a 64->64 3x3 layer site through the SAME dispatch the model entry points use (halo-tile
 *    tcgen05 kernel, stride-2 row kernel or per-tap pipeline, chosen by geometry alone): replaces Conv2d(64,64,3,s,p) /
 *    ConvTranspose2d(64,64,3,2) forward and their input gradients (models/models.py:54,59,66-78).  `in` / `out` are NHWC C=64;
 *    geometry: big side (BH,BW), small side (SH,SW), by = sy*stride - pad + ky; transposed = 1 when the output is the big side.
 *    wbf = bf16 hi/lo SWIZZLE_128B weight image from srlz_op_pack_conv_w_bf16 (9 taps x 16 KB) of the fp32 [tap][k][n] pack
 *    srlz_op_pack_conv_w writes (fwd_pack for forward, dgrad_pack for the input gradient).  in_scale / in_shift: BatchNorm +
 *    ReLU applied to `in` on load (or NULL); stats_partials: [n_partials][128] per-CTA sum / sum-of-squares (or NULL).
 *  srlz_op_wgrad64: weight gradient of the same sites (halo-tile tcgen05 kernel) in torch layout (64,64,3,3). */
int srlz_op_pack_conv_w(const float* w, float* fwd_pack, float* dgrad_pack, int ntaps, int transposed_conv, void* stream);
int srlz_op_pack_conv_w_bf16(const float* pack_f32, void* dst, int ntaps, void* stream);
int srlz_op_conv64(const float* in, const void* wbf, const float* bias, const float* in_scale, const float* in_shift,
                   float* out, int B, int BH, int BW, int SH, int SW, int K, int stride, int pad, int transposed,
                   float* stats_partials, int* n_partials, void* stream);
int srlz_op_wgrad64(const float* big, const float* small, const float* dense_scale, const float* dense_shift,
                    float* grad_out, int B, int BH, int BW, int SH, int SW, int K, int stride, int pad, void* workspace,
                    void* stream);
size_t srlz_op_wgrad64_workspace_bytes(int B, int BH, int BW, int SH, int SW, int K, int stride, int pad);
/* first / last layer kernels at layer granularity (the row-image / row-ring tcgen05 kernels of Conv2d(3,64,7,2,3) and
 * ConvTranspose2d(64,3,4,2), models/models.py:49,82).  workspace: srlz_op_layer_workspace_bytes() bytes.
 *  enc0_fwd : x (B,3,224,224) NCHW [+ rects (B,4) int32, DAE mask on load] -> y (B,112,112,64) NHWC pre-BN, optional per-CTA
 *             BatchNorm partials [n_partials][128] (sum | sum of squares per channel)
 *  enc0_wgrad: grad_w (64,3,7,7) for dy (B,112,112,64) NHWC
 *  dec12_fwd: decoded (B,3,224,224) NCHW = ConvTranspose2d(relu(y7*scale+shift)) + bias, y7 (B,111,111,64) NHWC; with target,
 *             sse_out[0] = sum (decoded-target)^2
 *  dec12_bwd: gradient source g_decoded, or coef*(decoded-target) when g_decoded is NULL -> grad_w (64,3,4,4), grad_b (3),
 *             dz (B,111,111,64) = ReLU-masked gradient w.r.t. the BatchNorm output, bn_partials [n_partials][128] =
 *             per-CTA (sum dz | sum dz*xhat), xhat = (y7-mean)*invstd */
size_t srlz_op_layer_workspace_bytes(void);
int srlz_op_enc0_fwd(const float* x, const int32_t* rects, const float* w, float* y, float* stats_partials, int* n_partials,
                     int B, void* workspace, void* stream);
int srlz_op_enc0_wgrad(const float* x, const int32_t* rects, const float* dy, float* grad_w, int B, void* workspace,
                       void* stream);
int srlz_op_dec12_fwd(const float* y7, const float* scale, const float* shift, const float* w, const float* bias,
                      float* decoded, const float* target, float* sse_out, int B, void* workspace, void* stream);
int srlz_op_dec12_bwd(const float* y7, const float* scale, const float* shift, const float* mean, const float* invstd,
                      const float* w, const float* g_decoded, const float* decoded, const float* target, float coef,
                      float* grad_w, float* grad_b, float* dz, float* bn_partials, int* n_partials, int B, void* workspace,
                      void* stream);
/* BatchNorm (per-channel scale / shift) + ReLU + MaxPool2d(3, 2, pad) forward of a pooled encoder stage (models/models.py:50-52,
 * 55-57,60-62): y (B,H,W,64) NHWC -> out (B,PH,PW,64) and argmax (uint8, first maximal tap ky*3+kx in scan order; may be NULL) */
int srlz_op_bn_relu_pool(const float* y, const float* scale, const float* shift, float* out, uint8_t* argmax, int B, int H, int W,
                         int PH, int PW, int pad, void* stream);
/* C[i,j] (+)= sum_k A(i,k) B(k,j) + bias[j] with element strides (sa_i, sa_k), (sb_k, sb_j), (sc_i, sc_j) */
int srlz_op_sgemm(const float* A, int64_t sa_i, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_j, float* C,
                  int64_t sc_i, int64_t sc_j, const float* bias, int M, int N, int K, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRLZ_H_ */

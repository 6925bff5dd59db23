// extern "C" entry points of libsrlz (include/srlz.h): orchestration of one model call forward / backward.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>

#include "../../include/srlz.h"
#include "common.cuh"
#include "kernels.h"

namespace srlz {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);  // every kernel launch in the library is followed by exactly one check_launch
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SRLZ_E_CUDA;
    }
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        if (n * 8 > SRLZ_MAX_PART) n = SRLZ_MAX_PART / 8;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------
// optional per-call-site timing (CUDA events on the launching stream); used by bench.py for the roofline line
// ---------------------------------------------------------------------------------------------------
enum ProfTag {
    T_PACK = 0, T_ENC0_FWD, T_ENC4_FWD, T_ENC8_FWD, T_BN_FIN, T_POOL_FWD, T_FC_FWD, T_VAE, T_DEC0_FWD, T_DEC3_FWD,
    T_DEC6_FWD, T_DEC9_FWD, T_DEC12_FWD, T_DEC12_BWD, T_BN_BWD, T_DEC9_WGRAD, T_DEC9_DGRAD, T_DEC6_WGRAD, T_DEC6_DGRAD,
    T_DEC3_WGRAD, T_DEC3_DGRAD, T_DEC0_WGRAD, T_DEC0_DGRAD, T_FC_BWD, T_POOL_BWD, T_ENC8_WGRAD, T_ENC8_DGRAD, T_ENC4_WGRAD,
    T_ENC4_DGRAD, T_ENC0_WGRAD, T_HEADS, T_ADAM, T_DEC12_WGRAD, T_DEC12_DGRAD, T_POOL_BWD_STATS, T_BN_BWD_FIN, T_PREPROC, T_COUNT
};
static const char* kTagNames[T_COUNT] = {
    "pack_weights", "enc0.fwd", "enc4.fwd", "enc8.fwd", "bn.finalize", "bn_relu_pool.fwd", "fc.fwd", "vae.reparam_kl",
    "dec0.fwd", "dec3.fwd", "dec6.fwd", "dec9.fwd", "dec12.fwd", "dec12.bwd", "bn.bwd", "dec9.wgrad", "dec9.dgrad",
    "dec6.wgrad", "dec6.dgrad", "dec3.wgrad", "dec3.dgrad", "dec0.wgrad", "dec0.dgrad", "fc.bwd", "pool.bwd", "enc8.wgrad",
    "enc8.dgrad", "enc4.wgrad", "enc4.dgrad", "enc0.wgrad", "heads", "adam", "dec12.wgrad", "dec12.dgrad", "pool.bwd_stats", "bn.bwd_finalize", "preprocess_u8"};
#define PROF_MAX 8192
struct ProfRec { cudaEvent_t e0, e1; int tag; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;            // the recorder is process-wide: begin/end pairs of one caller thread are kept together
static ProfRec g_prof[PROF_MAX];
static int g_prof_n = 0, g_prof_created = 0;
static thread_local bool t_prof_open = false;

static void prof_begin(int tag, cudaStream_t st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    g_prof_mu.lock();
    if (g_prof_n >= PROF_MAX) { g_prof_mu.unlock(); return; }
    t_prof_open = true;
    if (g_prof_n >= g_prof_created) {
        cudaEventCreate(&g_prof[g_prof_n].e0);
        cudaEventCreate(&g_prof[g_prof_n].e1);
        g_prof_created = g_prof_n + 1;
    }
    g_prof[g_prof_n].tag = tag;
    cudaEventRecord(g_prof[g_prof_n].e0, st);
}
static void prof_end(cudaStream_t st) {
    if (!t_prof_open) return;
    cudaEventRecord(g_prof[g_prof_n].e1, st);
    ++g_prof_n;
    t_prof_open = false;
    g_prof_mu.unlock();
}
#define PROF(tag, call) do { prof_begin(tag, st); int rc_ = (call); prof_end(st); if (rc_) return rc_; } while (0)

// ---------------------------------------------------------------------------------------------------
// geometry of the network (models/models.py:47-83) and the layout of the caller-allocated blocks
// ---------------------------------------------------------------------------------------------------
static const int kDecIn[4] = {6, 13, 27, 55};     // input edge of decoder_conv.{0,3,6,9}
static const int kDecOut[4] = {13, 27, 55, 111};

struct Saved {  // byte offsets inside `saved`
    size_t y1, a1, am1, y2, a2, am2, y3, a3, am3, lat, z, d0, y4, y5, y6, y7, bnsave, total;
};
static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static Saved saved_layout(int B, int S, int is_vae) {
    Saved s;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
    const size_t F = sizeof(float), b = (size_t)B;
    s.y1 = take(b * 112 * 112 * 64 * F);
    s.a1 = take(b * 56 * 56 * 64 * F);
    s.am1 = take(b * 56 * 56 * 64);
    s.y2 = take(b * 56 * 56 * 64 * F);
    s.a2 = take(b * 27 * 27 * 64 * F);
    s.am2 = take(b * 27 * 27 * 64);
    s.y3 = take(b * 14 * 14 * 64 * F);
    s.a3 = take(b * 6 * 6 * 64 * F);
    s.am3 = take(b * 6 * 6 * 64);
    s.lat = take(b * S * (is_vae ? 2 : 1) * F);  // AE: states ; VAE: mu | logvar
    s.z = take(b * S * F);
    s.d0 = take(b * 2304 * F);
    s.y4 = take(b * 13 * 13 * 64 * F);
    s.y5 = take(b * 27 * 27 * 64 * F);
    s.y6 = take(b * 55 * 55 * 64 * F);
    s.y7 = take(b * 111 * 111 * 64 * F);
    s.bnsave = take(7 * BNS_FLOATS * F);
    s.total = o;
    return s;
}

struct Pack {  // float offsets inside `wpack`
    size_t fc_enc, fc_dec_w, fc_dec_b, total;
    size_t enc_fb[2], enc_db[2], dec_fb[4], dec_db[4];  // bf16 hi/lo images for the tcgen05 kernels
    size_t dec12_d, dec12_db;                           // dec12 dgrad columns (fp32 staging + bf16 image)
    size_t dec12_fb;                                    // dec12 forward: 4 shifts x (hi|lo) 16-row bf16 images (16 KB)
    size_t enc0_rb;                                     // enc0 row-image kernels: 4 row pairs x (hi|lo) 64-row bf16 images (64 KB)
};
static Pack pack_layout(int S, int is_vae) {
    Pack p;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    p.fc_enc = take((size_t)(is_vae ? 2 : 1) * S * 2304);
    p.fc_dec_w = take((size_t)2304 * S);
    p.fc_dec_b = take(2304);
    for (int i = 0; i < 2; ++i) { p.enc_fb[i] = take(SRLZ_WBF_FLOATS); p.enc_db[i] = take(SRLZ_WBF_FLOATS); }
    for (int i = 0; i < 4; ++i) { p.dec_fb[i] = take(SRLZ_WBF_FLOATS); p.dec_db[i] = take(SRLZ_WBF_FLOATS); }
    p.dec12_d = take(4096); p.dec12_db = take(4096);
    p.dec12_fb = take(4096);
    p.enc0_rb = take(4 * 4096);
    p.total = o;
    return p;
}

struct Work {  // byte offsets inside `workspace`
    size_t bufA, bufB, partials, sse, wpart, coef, tmpw, tmpv, glat, gmu, glv, da3, tailc, total;
};
static size_t wgrad_partial_floats_max(int B) {
    size_t m = enc0_rows_wgrad_partial_floats();
    if (dec12_rows_wgrad_partial_floats() > m) m = dec12_rows_wgrad_partial_floats();
    const int big[6] = {56, 27, 13, 27, 55, 111}, small[6] = {56, 14, 6, 13, 27, 55};
    const int stride[6] = {1, 2, 2, 2, 2, 2}, pad[6] = {1, 1, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) {
        ConvGeom g{B, big[i], big[i], small[i], small[i], 3, 3, stride[i], pad[i]};
        size_t f = gwgrad64_partial_floats(g);
        if (f > m) m = f;
    }
    return m;
}
static Work work_layout(int B, int S, int is_vae) {
    Work w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
    const size_t F = sizeof(float), b = (size_t)B;
    w.bufA = take(b * 112 * 112 * 64 * F);
    w.bufB = take(b * 56 * 56 * 64 * F);
    w.partials = take((size_t)SRLZ_MAX_PART * 128 * F);
    w.sse = take((size_t)SRLZ_MAX_PART * F);
    w.wpart = take(wgrad_partial_floats_max(B) * F);
    w.coef = take(128 * F);
    w.tmpw = take((size_t)2304 * S * (is_vae ? 2 : 1) * F);
    w.tmpv = take(2304 * F);
    w.glat = take(b * S * F);
    w.gmu = take(b * S * F);
    w.glv = take(b * S * F);
    w.da3 = take(b * 2304 * F);
    w.tailc = take(256);   // ticket counter of the fused finalizes (bn_tail.cuh): zeroed at the top of every forward / backward
    w.total = o;
    return w;
}

#ifdef SRLZ_DEV
static long long* g_dbg = nullptr;   // development builds only: clock64 timeline buffer
static int g_dbg_site = -1;          // which call site stamps it (0 enc0.fwd, 1 dec12.dgrad, 2 dec9.dgrad, 3 dec12.fwd, 4 enc0.wgrad, 5 dec12.wgrad, 6 dec9.wgrad, 7 enc4.wgrad, 8 enc4.fwd, 9 dec9.fwd)
#define DBG_AT(site) (g_dbg_site == (site) ? g_dbg : nullptr)
#else
#define DBG_AT(site) nullptr
#endif

// The ONE path of every 64->64 layer: the halo-tile kernel wherever the gathered image is unit-stride for the taps
// (conv3x3 s1 forward / dgrad, every transposed-conv forward, dgrad of the stride-2 conv), the row kernel for the stride-2
// gathers of the transposed-conv dgrads, the per-tap pipeline for the one remaining stride-2 gather (encoder_conv.8 forward).
static int conv64(const GConvArgs& a, const float* wpack, size_t bf_off, int* np, cudaStream_t st) {
    if (gconv64_halo_supported(a)) return gconv64_halo(a, wpack + bf_off, np, st);
    if (gconv64_s2rows_supported(a)) return gconv64_s2rows(a, wpack + bf_off, np, st);
    return gconv64_tc(a, wpack + bf_off, np, st);
}

static int wgrad64(const GWgradArgs& a, float* grad, int acc, cudaStream_t st) { return gwgrad64_halo(a, grad, acc, st); }

static BnParams to_bn(const srlz_bn& b) {
    BnParams p;
    p.gamma = b.weight; p.beta = b.bias; p.running_mean = b.running_mean; p.running_var = b.running_var;
    p.num_batches_tracked = reinterpret_cast<long long*>(b.num_batches_tracked);
    return p;
}

#define RC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

__global__ void add_or_copy_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, int n) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b != nullptr ? b[i] : 0.f);
}

// z_in != nullptr: decoder-only call (srlz_decode): the encoder and the bottleneck are skipped and z_in feeds decoder_fc
static int forward_impl(const srlz_net* net, const float* wpack, const float* x, const int* rects, const float* eps, int B,
                        int training, float* lat_out, float* logvar_out, float* decoded, const float* target,
                        float* loss_out, char* saved, char* ws, cudaStream_t st, const float* z_in = nullptr) {
    const int S = net->state_dim, vae = net->is_vae;
    const Saved sv = saved_layout(B, S, vae);
    const Pack pk = pack_layout(S, vae);
    const Work wk = work_layout(B, S, vae);
    float* partials = reinterpret_cast<float*>(ws + wk.partials);
    float* bns = reinterpret_cast<float*>(saved + sv.bnsave);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(saved + off); };
    auto U = [&](size_t off) { return reinterpret_cast<unsigned char*>(saved + off); };
    int np = 0;
    // BatchNorm statistics are finalized by the last CTA of the kernel that produces them (training mode; bn_tail.cuh)
    unsigned int* tailc = reinterpret_cast<unsigned int*>(ws + wk.tailc);
    if (training) cudaMemsetAsync(tailc, 0, sizeof(unsigned int), st);
    auto fwd_tail = [&](const srlz_bn& b, long long count, int bn_idx) {
        BnTail t{};
        if (training) { t.counter = tailc; t.kind = BnTail::FORWARD; t.count = (double)count; t.bn = to_bn(b); t.out0 = bns + bn_idx * BNS_FLOATS; }
        return t;
    };

    float* z = F(sv.z);
    if (z_in != nullptr) {
        cudaMemcpyAsync(z, z_in, (size_t)B * S * sizeof(float), cudaMemcpyDeviceToDevice, st);   // kept for decoder_fc's wgrad
    } else {
    // ---- encoder (models/models.py:47-63) ----
    {
        GConvArgs e{};
        e.in = x; e.out = F(sv.y1); e.partials = partials; e.g = ConvGeom{B, 224, 224, 112, 112, 1, 3, 2, 3}; e.transposed = 0;
        e.epi = training ? EPI_STATS : EPI_PLAIN; e.mode = 1; e.rects = rects; e.dbg = DBG_AT(0);
        e.tail = fwd_tail(net->enc_bn[0], (long long)B * 112 * 112, 0);
        PROF(T_ENC0_FWD, enc0_rows_fwd(e, wpack + pk.enc0_rb, &np, st));
    }
    if (!training) PROF(T_BN_FIN, bn_finalize(partials, np, (long long)B * 112 * 112, to_bn(net->enc_bn[0]), training, bns + 0 * BNS_FLOATS, st));
    PROF(T_POOL_FWD, bn_relu_pool_fwd(F(sv.y1), bns + BNS_SCALE, bns + BNS_SHIFT, F(sv.a1), U(sv.am1), B, 112, 112, 56, 56, 1, st));

    GConvArgs c{};
    c.in = F(sv.a1); c.out = F(sv.y2); c.partials = partials;
    c.g = ConvGeom{B, 56, 56, 56, 56, 3, 3, 1, 1}; c.transposed = 0; c.epi = training ? EPI_STATS : EPI_PLAIN; c.dbg = DBG_AT(8);
    c.tail = fwd_tail(net->enc_bn[1], (long long)B * 56 * 56, 1);
    PROF(T_ENC4_FWD, conv64(c, wpack, pk.enc_fb[0], &np, st));
    c.dbg = nullptr;
    if (!training) PROF(T_BN_FIN, bn_finalize(partials, np, (long long)B * 56 * 56, to_bn(net->enc_bn[1]), training, bns + 1 * BNS_FLOATS, st));
    PROF(T_POOL_FWD, bn_relu_pool_fwd(F(sv.y2), bns + BNS_FLOATS + BNS_SCALE, bns + BNS_FLOATS + BNS_SHIFT, F(sv.a2), U(sv.am2), B, 56, 56, 27, 27, 0, st));

    c.in = F(sv.a2); c.out = F(sv.y3);
    c.g = ConvGeom{B, 27, 27, 14, 14, 3, 3, 2, 1};
    c.tail = fwd_tail(net->enc_bn[2], (long long)B * 14 * 14, 2);
    PROF(T_ENC8_FWD, conv64(c, wpack, pk.enc_fb[1], &np, st));
    if (!training) PROF(T_BN_FIN, bn_finalize(partials, np, (long long)B * 14 * 14, to_bn(net->enc_bn[2]), training, bns + 2 * BNS_FLOATS, st));
    PROF(T_POOL_FWD, bn_relu_pool_fwd(F(sv.y3), bns + 2 * BNS_FLOATS + BNS_SCALE, bns + 2 * BNS_FLOATS + BNS_SHIFT, F(sv.a3), U(sv.am3), B, 14, 14, 6, 6, 0, st));

    // ---- bottleneck (models/autoencoders.py:102-118 ; models/vae.py:59-75 ; models/models.py:147-165) ----
    float* lat = F(sv.lat);  // AE: states (B,S) ; VAE: mu (B,S) then logvar (B,S)
    const float* fce = wpack + pk.fc_enc;
    float* tmpw_f = reinterpret_cast<float*>(ws + wk.tmpw);
    const size_t tmpw_n = (size_t)2304 * S * (vae ? 2 : 1);
    PROF(T_FC_FWD, sgemm_splitk(F(sv.a3), 2304, 1, fce, 1, 2304, lat, S, 1, net->fc_enc_b[0], B, S, 2304, 0, tmpw_f, tmpw_n, st));
    if (vae) {
        float* lv = lat + (size_t)B * S;
        PROF(T_FC_FWD, sgemm_splitk(F(sv.a3), 2304, 1, fce + (size_t)S * 2304, 1, 2304, lv, S, 1, net->fc_enc_b[1], B, S, 2304, 0, tmpw_f, tmpw_n, st));
        // encoder-only passes (getStates) never sample: z is unused there
        const int sample = training && decoded != nullptr;
        if (sample && eps == nullptr) { set_error("srlz_forward: VAE training forward needs eps"); return SRLZ_E_ARG; }
        float* ssep = reinterpret_cast<float*>(ws + wk.sse);
        PROF(T_VAE, vae_reparam_fwd(lat, lv, eps, z, ssep, B * S, sample, &np, st));
        if (loss_out != nullptr) RC(sum_partials(ssep, np, 1.f, loss_out + 1, 0, st));
        if (lat_out != nullptr) cudaMemcpyAsync(lat_out, lat, (size_t)B * S * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (logvar_out != nullptr) cudaMemcpyAsync(logvar_out, lv, (size_t)B * S * sizeof(float), cudaMemcpyDeviceToDevice, st);
    } else {
        if (lat_out != nullptr) cudaMemcpyAsync(lat_out, lat, (size_t)B * S * sizeof(float), cudaMemcpyDeviceToDevice, st);
        z = lat;
    }
    }   // encoder + bottleneck
    if (decoded == nullptr) return 0;

    // ---- decoder (models/models.py:65-83) ----
    PROF(T_FC_FWD, sgemm(z, S, 1, wpack + pk.fc_dec_w, 1, S, F(sv.d0), 2304, 1, wpack + pk.fc_dec_b, B, 2304, S, 0, st));
    const size_t yoff[5] = {sv.d0, sv.y4, sv.y5, sv.y6, sv.y7};
    for (int l = 0; l < 4; ++l) {
        GConvArgs d{};
        d.in = F(yoff[l]); d.bias = net->dec_b[l]; d.out = F(yoff[l + 1]);
        if (l > 0) { d.in_scale = bns + (2 + l) * BNS_FLOATS + BNS_SCALE; d.in_shift = bns + (2 + l) * BNS_FLOATS + BNS_SHIFT; }
        d.partials = partials;
        d.g = ConvGeom{B, kDecOut[l], kDecOut[l], kDecIn[l], kDecIn[l], 3, 3, 2, 0};
        d.transposed = 1; d.epi = training ? EPI_STATS : EPI_PLAIN;
        if (l == 3) d.dbg = DBG_AT(9);
        d.tail = fwd_tail(net->dec_bn[l], (long long)B * kDecOut[l] * kDecOut[l], 3 + l);
        PROF(T_DEC0_FWD + l, conv64(d, wpack, pk.dec_fb[l], &np, st));
        if (!training) PROF(T_BN_FIN, bn_finalize(partials, np, (long long)B * kDecOut[l] * kDecOut[l], to_bn(net->dec_bn[l]), training, bns + (3 + l) * BNS_FLOATS, st));
    }
    float* ssep = reinterpret_cast<float*>(ws + wk.sse);
    {
        GConvArgs d{};
        d.in = F(sv.y7); d.in_scale = bns + 6 * BNS_FLOATS + BNS_SCALE; d.in_shift = bns + 6 * BNS_FLOATS + BNS_SHIFT;
        d.bias = net->dec_b[4]; d.out = decoded; d.aux2 = target; d.partials = target != nullptr ? ssep : nullptr;
        d.g = ConvGeom{B, 224, 224, 111, 111, 4, 4, 2, 0}; d.transposed = 1; d.epi = EPI_DEC12;
        d.dbg = DBG_AT(3);
        PROF(T_DEC12_FWD, dec12_rows_fwd(d, wpack + pk.dec12_fb, &np, st));
    }
    if (target != nullptr && loss_out != nullptr) RC(sum_partials(ssep, np, 1.f, loss_out, 0, st));
    return 0;
}

static int backward_impl(const srlz_net* net, const float* wpack, const srlz_net_grads* gr, int acc, const float* x,
                         const int* rects, const float* eps, int B, int training, int has_decoder, const float* g_decoded,
                         const float* decoded, const float* target, float mse_coef, const float* g_lat, const float* g_logvar,
                         float kl_coef, char* saved, char* ws, cudaStream_t st, float* g_z_out = nullptr) {
    // g_z_out != nullptr: decoder-only call (srlz_decode_backward): stops after decoder_fc and hands out d/dz
    const int S = net->state_dim, vae = net->is_vae;
    const Saved sv = saved_layout(B, S, vae);
    const Pack pk = pack_layout(S, vae);
    const Work wk = work_layout(B, S, vae);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(saved + off); };
    auto U = [&](size_t off) { return reinterpret_cast<unsigned char*>(saved + off); };
    auto W = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    float* bufA = W(wk.bufA);
    float* bufB = W(wk.bufB);
    float* partials = W(wk.partials);
    float* wpart = W(wk.wpart);
    float* coef = W(wk.coef);
    float* bns = F(sv.bnsave);
    float* glat = W(wk.glat);
    int np = 0;
    // The BatchNorm-backward sums of the tensor-core dgrad kernels are finalized by the last CTA of the kernel that produces them
    // (bn_tail.cuh); eval-mode BN (a fixed affine map: coef zeroed after the finalize) keeps the separate launch.
    unsigned int* tailc = reinterpret_cast<unsigned int*>(ws + wk.tailc);
    cudaMemsetAsync(tailc, 0, sizeof(unsigned int), st);
    auto bwd_tail = [&](long long npix, float* dgamma, float* dbeta) {
        BnTail t{};
        if (training) { t.counter = tailc; t.kind = BnTail::BACKWARD; t.accumulate = acc; t.count = (double)npix; t.out0 = coef; t.out1 = dgamma; t.out2 = dbeta; }
        return t;
    };
    auto bn_bwd = [&](float* dz, const float* y, const srlz_bn& bn, int bn_idx, long long npix, float* dgamma, float* dbeta,
                      float* dbias) -> int {   // (the producer of `partials` ran with bwd_tail(npix, dgamma, dbeta))
        const float* b = bns + bn_idx * BNS_FLOATS;
        if (!training) PROF(T_BN_BWD_FIN, bn_bwd_finalize(partials, np, npix, coef, dgamma, dbeta, acc, st));
        if (!training) cudaMemsetAsync(coef, 0, 128 * sizeof(float), st);  // eval-mode BN is a fixed affine map
        PROF(T_BN_BWD, bn_bwd_apply(dz, y, bn.weight, b + BNS_MEAN, b + BNS_INVSTD, coef, npix, dbias, partials, acc, st));
        return 0;
    };

    // MaxPool + ReLU + BatchNorm backward of one encoder stage: the BN-backward sums come from the pooled side (dpool and the
    // pooled activation `a`, quarter-size tensors), so one full-size pass recomputes the masked pooling gradient and writes
    // the finished dy (instead of masked gradient written -> re-read with y -> rewritten).
    auto pool_bn_bwd = [&](const float* dpool, const float* a, const unsigned char* am, const float* y, const srlz_bn& bn, int bn_idx,
                           float* dy, int H, int PH, int pad, float* dgamma, float* dbeta) -> int {
        const float* b = bns + bn_idx * BNS_FLOATS;
        PROF(T_POOL_BWD_STATS, pool_bwd_stats(dpool, a, am, y, bn.weight, bn.bias, b + BNS_MEAN, b + BNS_INVSTD, partials, &np, B, H, H, PH, PH, pad, st));
        PROF(T_BN_BWD_FIN, bn_bwd_finalize(partials, np, (long long)B * H * H, coef, dgamma, dbeta, acc, st));
        if (!training) cudaMemsetAsync(coef, 0, 128 * sizeof(float), st);  // eval-mode BN is a fixed affine map
        PROF(T_POOL_BWD, pool_bwd_bn_apply(dpool, am, y, b + BNS_SCALE, b + BNS_SHIFT, b + BNS_MEAN, b + BNS_INVSTD, bn.weight, coef, dy, B, H, H, PH, PH, pad, st));
        return 0;
    };
    const float* z = (vae || g_z_out != nullptr) ? F(sv.z) : F(sv.lat);
    if (has_decoder) {
        // ---- decoder_conv.12 (ConvTranspose2d 64->3) ----
        const float* b6 = bns + 6 * BNS_FLOATS;
        if (g_decoded == nullptr && (decoded == nullptr || target == nullptr)) { set_error("srlz_backward: need g_decoded or decoded+target"); return SRLZ_E_ARG; }
        {
            // wgrad + bias grad in one pass, then dgrad (+ReLU mask + BN-backward sums), all on the tensor-core kernels
            GWgradArgs wg{};
            wg.big = g_decoded != nullptr ? g_decoded : decoded; wg.small = F(sv.y7); wg.dense_scale = b6 + BNS_SCALE; wg.dense_shift = b6 + BNS_SHIFT;
            wg.partials = wpart; wg.g = ConvGeom{B, 224, 224, 111, 111, 4, 4, 2, 0}; wg.mode = 2;
            wg.aux0 = g_decoded; wg.aux1 = decoded; wg.aux2 = target; wg.coef = mse_coef; wg.dbg = DBG_AT(5);
            PROF(T_DEC12_WGRAD, dec12_rows_wgrad(wg, gr->dec_w[4], gr->dec_b[4], acc, st));
            GConvArgs dg{};
            dg.in = g_decoded != nullptr ? g_decoded : decoded; dg.out = bufA; dg.partials = partials;
            dg.g = ConvGeom{B, 224, 224, 111, 111, 1, 1, 2, 0}; dg.transposed = 0; dg.epi = EPI_MASK_BNBWD; dg.mode = 2;
            dg.e_ypre = F(sv.y7); dg.e_scale = b6 + BNS_SCALE; dg.e_shift = b6 + BNS_SHIFT; dg.e_mean = b6 + BNS_MEAN; dg.e_invstd = b6 + BNS_INVSTD;
            dg.aux0 = g_decoded; dg.aux1 = decoded; dg.aux2 = target; dg.coef = mse_coef;
            dg.dbg = DBG_AT(1);
            dg.tail = bwd_tail((long long)B * 111 * 111, gr->dec_bn_w[3], gr->dec_bn_b[3]);
            PROF(T_DEC12_DGRAD, gconv64_tc(dg, wpack + pk.dec12_db, &np, st));
        }
        RC(bn_bwd(bufA, F(sv.y7), net->dec_bn[3], 6, (long long)B * 111 * 111, gr->dec_bn_w[3], gr->dec_bn_b[3], gr->dec_b[3]));
        // ---- decoder_conv.{9,6,3,0} ----
        const size_t yoff[5] = {sv.d0, sv.y4, sv.y5, sv.y6, sv.y7};
        float* cur = bufA;   // dy of layer l's output
        float* nxt = bufB;
        for (int l = 3; l >= 0; --l) {
            const ConvGeom g{B, kDecOut[l], kDecOut[l], kDecIn[l], kDecIn[l], 3, 3, 2, 0};
            GWgradArgs wg{};
            wg.big = cur; wg.small = F(yoff[l]); wg.partials = wpart; wg.g = g;
            if (l > 0) { wg.dense_scale = bns + (2 + l) * BNS_FLOATS + BNS_SCALE; wg.dense_shift = bns + (2 + l) * BNS_FLOATS + BNS_SHIFT; }
            if (l == 3) wg.dbg = DBG_AT(6);
            PROF(T_DEC0_WGRAD - 2 * l, wgrad64(wg, gr->dec_w[l], acc, st));
            GConvArgs dg{};
            dg.in = cur; dg.out = nxt; dg.g = g; dg.transposed = 0; dg.partials = partials;
            if (l > 0) {
                const float* bl = bns + (2 + l) * BNS_FLOATS;
                dg.epi = EPI_MASK_BNBWD; dg.e_ypre = F(yoff[l]); dg.e_scale = bl + BNS_SCALE; dg.e_shift = bl + BNS_SHIFT;
                dg.e_mean = bl + BNS_MEAN; dg.e_invstd = bl + BNS_INVSTD;
                dg.tail = bwd_tail((long long)B * kDecIn[l] * kDecIn[l], gr->dec_bn_w[l - 1], gr->dec_bn_b[l - 1]);
            } else {
                dg.epi = EPI_PLAIN;
            }
            if (l == 3) dg.dbg = DBG_AT(2);
            PROF(T_DEC0_DGRAD - 2 * l, conv64(dg, wpack, pk.dec_db[l], &np, st));
            if (l > 0)
                RC(bn_bwd(nxt, F(yoff[l]), net->dec_bn[l - 1], 2 + l, (long long)B * kDecIn[l] * kDecIn[l], gr->dec_bn_w[l - 1],
                          gr->dec_bn_b[l - 1], gr->dec_b[l - 1]));
            float* t = cur; cur = nxt; nxt = t;
        }
        float* dd0 = cur;  // (B,6,6,64) = (B,2304) in NHWC order
        // ---- decoder_fc ----
        float* tmpw = W(wk.tmpw);
        // weight / bias gradients written straight at their torch positions (rows of the 2304 axis permuted in the epilogue)
        PROF(T_FC_BWD, sgemm_perm(dd0, 1, 2304, z, S, 1, gr->fc_dec_w, S, 1, nullptr, 2304, S, B, acc, 1, st));
        PROF(T_FC_BWD, colsum_perm(dd0, B, 2304, gr->fc_dec_b, acc, 1, st));
        PROF(T_FC_BWD, sgemm_splitk(dd0, 2304, 1, wpack + pk.fc_dec_w, S, 1, glat, S, 1, nullptr, B, S, 2304, 0, tmpw, (size_t)2304 * S * (vae ? 2 : 1), st));
    } else {
        cudaMemsetAsync(glat, 0, (size_t)B * S * sizeof(float), st);
    }
    if (g_z_out != nullptr) {
        cudaMemcpyAsync(g_z_out, glat, (size_t)B * S * sizeof(float), cudaMemcpyDeviceToDevice, st);
        return 0;
    }

    // ---- bottleneck ----
    float* da3 = W(wk.da3);
    float* tmpw = W(wk.tmpw);
    const float* fce = wpack + pk.fc_enc;
    if (vae) {
        float* gmu = W(wk.gmu);
        float* glv = W(wk.glv);
        const float* mu = F(sv.lat);
        const float* lv = mu + (size_t)B * S;
        PROF(T_VAE, vae_reparam_bwd(glat, lv, eps, g_lat, g_logvar, kl_coef, mu, gmu, glv, B * S, training && has_decoder, st));
        const float* gs[2] = {gmu, glv};
        for (int h = 0; h < 2; ++h) {
            PROF(T_FC_BWD, sgemm_perm(gs[h], 1, S, F(sv.a3), 2304, 1, gr->fc_enc_w[h], 2304, 1, nullptr, S, 2304, B, acc, 2, st));
            PROF(T_FC_BWD, colsum(gs[h], B, S, gr->fc_enc_b[h], acc, st));
            PROF(T_FC_BWD, sgemm(gs[h], S, 1, fce + (size_t)h * S * 2304, 2304, 1, da3, 2304, 1, nullptr, B, 2304, S, h, st));
        }
    } else {
        float* gst = W(wk.gmu);
        launch_k(add_or_copy_kernel, (B * S + 255) / 256, 256, 0, st, gst, glat, g_lat, B * S);
        PROF(T_FC_BWD, sgemm_perm(gst, 1, S, F(sv.a3), 2304, 1, gr->fc_enc_w[0], 2304, 1, nullptr, S, 2304, B, acc, 2, st));
        PROF(T_FC_BWD, colsum(gst, B, S, gr->fc_enc_b[0], acc, st));
        PROF(T_FC_BWD, sgemm(gst, S, 1, fce, 2304, 1, da3, 2304, 1, nullptr, B, 2304, S, 0, st));
    }

    // ---- encoder ----
    RC(pool_bn_bwd(da3, F(sv.a3), U(sv.am3), F(sv.y3), net->enc_bn[2], 2, bufA, 14, 6, 0, gr->enc_bn_w[2], gr->enc_bn_b[2]));
    {
        const ConvGeom g{B, 27, 27, 14, 14, 3, 3, 2, 1};
        GWgradArgs wg{}; wg.big = F(sv.a2); wg.small = bufA; wg.partials = wpart; wg.g = g;
        PROF(T_ENC8_WGRAD, wgrad64(wg, gr->enc_w[2], acc, st));
        GConvArgs dg{}; dg.in = bufA; dg.out = bufB; dg.g = g; dg.transposed = 1; dg.epi = EPI_PLAIN;
        PROF(T_ENC8_DGRAD, conv64(dg, wpack, pk.enc_db[1], &np, st));
    }
    RC(pool_bn_bwd(bufB, F(sv.a2), U(sv.am2), F(sv.y2), net->enc_bn[1], 1, bufA, 56, 27, 0, gr->enc_bn_w[1], gr->enc_bn_b[1]));
    {
        const ConvGeom g{B, 56, 56, 56, 56, 3, 3, 1, 1};
        GWgradArgs wg{}; wg.big = F(sv.a1); wg.small = bufA; wg.partials = wpart; wg.g = g; wg.dbg = DBG_AT(7);
        PROF(T_ENC4_WGRAD, wgrad64(wg, gr->enc_w[1], acc, st));
        GConvArgs dg{}; dg.in = bufA; dg.out = bufB; dg.g = g; dg.transposed = 1; dg.epi = EPI_PLAIN;
        PROF(T_ENC4_DGRAD, conv64(dg, wpack, pk.enc_db[0], &np, st));
    }
    RC(pool_bn_bwd(bufB, F(sv.a1), U(sv.am1), F(sv.y1), net->enc_bn[0], 0, bufA, 112, 56, 1, gr->enc_bn_w[0], gr->enc_bn_b[0]));
    {
        GWgradArgs wg{};
        wg.big = x; wg.small = bufA; wg.partials = wpart; wg.g = ConvGeom{B, 224, 224, 112, 112, 7, 7, 2, 3}; wg.mode = 1; wg.rects = rects; wg.dbg = DBG_AT(4);
        PROF(T_ENC0_WGRAD, enc0_rows_wgrad(wg, gr->enc_w[0], acc, st));
    }
    return 0;
}

}  // namespace srlz

using namespace srlz;

extern "C" {

int srlz_version(void) { return SRLZ_VERSION; }

/* number of CUDA kernels this library has launched in this process */
long long srlz_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void srlz_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on.store(on != 0);
    g_prof_n = 0;
}

/* synchronises the device, then writes "tag count total_ms\n" lines for every call site seen since enable */
int srlz_prof_report(char* buf, int buf_len) {
    if (buf == nullptr || buf_len <= 0) return SRLZ_E_ARG;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms[T_COUNT] = {0};
    int cnt[T_COUNT] = {0};
    for (int i = 0; i < g_prof_n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1) == cudaSuccess) {
            ms[g_prof[i].tag] += t;
            cnt[g_prof[i].tag] += 1;
        }
    }
    int o = 0;
    buf[0] = 0;
    for (int t = 0; t < T_COUNT && o < buf_len - 64; ++t)
        if (cnt[t] > 0) o += snprintf(buf + o, buf_len - o, "%s %d %.6f\n", kTagNames[t], cnt[t], ms[t]);
    g_prof_n = 0;
    return 0;
}
const char* srlz_last_error(void) { return g_err; }

size_t srlz_pack_floats(int is_vae, int state_dim) { return pack_layout(state_dim, is_vae).total; }
size_t srlz_saved_bytes(int B, int state_dim, int is_vae) { return saved_layout(B, state_dim, is_vae).total; }
size_t srlz_workspace_bytes(int B, int state_dim, int is_vae) { return work_layout(B, state_dim, is_vae).total; }

const char* srlz_saved_names(void) { return "y1,a1,am1,y2,a2,am2,y3,a3,am3,lat,z,d0,y4,y5,y6,y7,bnsave"; }

int srlz_saved_layout(int B, int state_dim, int is_vae, size_t* offsets, int max_entries) {
    const Saved s = saved_layout(B, state_dim, is_vae);
    const size_t v[17] = {s.y1, s.a1, s.am1, s.y2, s.a2, s.am2, s.y3, s.a3, s.am3, s.lat, s.z, s.d0, s.y4, s.y5, s.y6, s.y7, s.bnsave};
    int n = 0;
    for (; n < 17 && n < max_entries; ++n) offsets[n] = v[n];
    return n;
}

int srlz_pack_weights(const srlz_net* net, float* wpack, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (net == nullptr || wpack == nullptr) { set_error("srlz_pack_weights: null argument"); return SRLZ_E_ARG; }
    const int S = net->state_dim, vae = net->is_vae;
    const Pack pk = pack_layout(S, vae);
    auto pack_all = [&]() -> int {
        ConvPackJobs jobs{};
        for (int i = 0; i < 2; ++i) {
            jobs.w[i] = net->enc_w[1 + i]; jobs.transposed[i] = 0;
            jobs.fwd[i] = reinterpret_cast<unsigned char*>(wpack + pk.enc_fb[i]); jobs.dgrad[i] = reinterpret_cast<unsigned char*>(wpack + pk.enc_db[i]);
        }
        for (int i = 0; i < 4; ++i) {
            jobs.w[2 + i] = net->dec_w[i]; jobs.transposed[2 + i] = 1;
            jobs.fwd[2 + i] = reinterpret_cast<unsigned char*>(wpack + pk.dec_fb[i]); jobs.dgrad[2 + i] = reinterpret_cast<unsigned char*>(wpack + pk.dec_db[i]);
        }
        RC(pack_conv_layers_bf16(jobs, st));
        RC(pack_dec12_dgrad(net->dec_w[4], wpack + pk.dec12_d, st));
        RC(pack_conv_w_bf16(wpack + pk.dec12_d, wpack + pk.dec12_db, 1, st));
        RC(pack_dec12_fwd_bf16(net->dec_w[4], wpack + pk.dec12_fb, st));
        RC(pack_enc0_rows_bf16(net->enc_w[0], wpack + pk.enc0_rb, st));
        for (int h = 0; h < (vae ? 2 : 1); ++h) RC(permute_fc(net->fc_enc_w[h], wpack + pk.fc_enc + (size_t)h * S * 2304, S, 1, 0, 0, st));
        RC(permute_fc(net->fc_dec_w, wpack + pk.fc_dec_w, S, 1, 1, 0, st));
        RC(permute_fc(net->fc_dec_b, wpack + pk.fc_dec_b, 1, 1, 1, 0, st));
        return 0;
    };
    PROF(T_PACK, pack_all());
    return 0;
}

int srlz_forward(const srlz_net* net, const float* wpack, const float* x, const int32_t* rects, const float* eps, int B,
                 int training, float* lat, float* logvar, float* decoded, const float* target, float* loss_out, void* saved,
                 void* workspace, void* stream) {
    if (net == nullptr || wpack == nullptr || x == nullptr || saved == nullptr || workspace == nullptr || B <= 0) {
        set_error("srlz_forward: null argument or B <= 0");
        return SRLZ_E_ARG;
    }
    if (B > SRLZ_MAX_BATCH) { set_error("srlz_forward: B exceeds SRLZ_MAX_BATCH (2048 images per call)"); return SRLZ_E_ARG; }
    if (net->state_dim <= 0 || net->state_dim % 4 != 0) { set_error("srlz_forward: state_dim must be a positive multiple of 4"); return SRLZ_E_ARG; }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(decoded) | reinterpret_cast<uintptr_t>(target)) & 7) {   // float2 accesses in the first / last-layer kernels
        set_error("srlz_forward: x / decoded / target must be 8-byte aligned");
        return SRLZ_E_ARG;
    }
    return forward_impl(net, wpack, x, rects, eps, B, training, lat, logvar, decoded, target, loss_out, (char*)saved,
                        (char*)workspace, (cudaStream_t)stream);
}

int srlz_replay_running_stats(const srlz_net* net, int B, void* saved, void* stream) {
    if (net == nullptr || saved == nullptr) { set_error("srlz_replay_running_stats: null argument"); return SRLZ_E_ARG; }
    const Saved sv = saved_layout(B, net->state_dim, net->is_vae);
    const float* bns = reinterpret_cast<const float*>((char*)saved + sv.bnsave);
    const long long cnt[3] = {(long long)B * 112 * 112, (long long)B * 56 * 56, (long long)B * 14 * 14};
    for (int i = 0; i < 3; ++i) RC(bn_running_update(bns + i * BNS_FLOATS, cnt[i], to_bn(net->enc_bn[i]), (cudaStream_t)stream));
    return 0;
}

int srlz_backward(const srlz_net* net, const float* wpack, const srlz_net_grads* grads, int accumulate, const float* x,
                  const int32_t* rects, const float* eps, int B, int training, int has_decoder, const float* g_decoded,
                  const float* decoded, const float* target, float mse_coef, const float* g_lat, const float* g_logvar,
                  float kl_coef, void* saved, void* workspace, void* stream) {
    if (net == nullptr || wpack == nullptr || grads == nullptr || x == nullptr || saved == nullptr || workspace == nullptr || B <= 0) {
        set_error("srlz_backward: null argument or B <= 0");
        return SRLZ_E_ARG;
    }
    if (B > SRLZ_MAX_BATCH) { set_error("srlz_backward: B exceeds SRLZ_MAX_BATCH (2048 images per call)"); return SRLZ_E_ARG; }
    if ((reinterpret_cast<uintptr_t>(x) & 7) ||
        ((reinterpret_cast<uintptr_t>(g_decoded) | reinterpret_cast<uintptr_t>(decoded) | reinterpret_cast<uintptr_t>(target)) & 15)) {
        set_error("srlz_backward: x must be 8-byte, g_decoded / decoded / target 16-byte aligned");
        return SRLZ_E_ARG;
    }
    return backward_impl(net, wpack, grads, accumulate, x, rects, eps, B, training, has_decoder, g_decoded, decoded, target,
                         mse_coef, g_lat, g_logvar, kl_coef, (char*)saved, (char*)workspace, (cudaStream_t)stream);
}

int srlz_preprocess_u8(const uint8_t* frames, float* out, int B, void* stream) {
    if (frames == nullptr || out == nullptr || B <= 0) { set_error("srlz_preprocess_u8: null argument or B <= 0"); return SRLZ_E_ARG; }
    if ((reinterpret_cast<uintptr_t>(frames) & 3) || (reinterpret_cast<uintptr_t>(out) & 3)) { set_error("srlz_preprocess_u8: pointers must be 4-byte aligned"); return SRLZ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    PROF(T_PREPROC, preprocess_u8(frames, out, B, st));
    return 0;
}

// ---- inference path: eval-mode encoder with BatchNorm folded into the conv weights ----
struct EvalPack {  // float offsets inside `epack`
    size_t enc0_rb, enc_fb[2], fc_enc, bias[3], ones, wtmp, ftmp, dtmp, total;
};
static EvalPack eval_pack_layout(int S) {
    EvalPack p;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    p.enc0_rb = take(4 * 4096);
    for (int i = 0; i < 2; ++i) p.enc_fb[i] = take(SRLZ_WBF_FLOATS);
    p.fc_enc = take((size_t)S * 2304);
    for (int i = 0; i < 3; ++i) p.bias[i] = take(64);
    p.ones = take(64);
    p.wtmp = take(9 * 4096);   // folded weights in torch layout (largest: 64 x 64 x 3 x 3)
    p.ftmp = take(9 * 4096);   // fp32 [tap][ci][co] staging pack
    p.dtmp = take(9 * 4096);   // (dgrad pack written by the same kernel; unused)
    p.total = o;
    return p;
}
struct EvalWork { size_t y1, a1, y2, a2, y3, a3, tmpw, total; };
static EvalWork eval_work_layout(int B, int S) {
    EvalWork w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
    const size_t F = sizeof(float), b = (size_t)B;
    w.y1 = take(b * 112 * 112 * 64 * F); w.a1 = take(b * 56 * 56 * 64 * F);
    w.y2 = take(b * 56 * 56 * 64 * F);   w.a2 = take(b * 27 * 27 * 64 * F);
    w.y3 = take(b * 14 * 14 * 64 * F);   w.a3 = take(b * 6 * 6 * 64 * F);
    w.tmpw = take((size_t)2304 * S * F);
    w.total = o;
    return w;
}

size_t srlz_eval_pack_floats(int state_dim) { return eval_pack_layout(state_dim).total; }
size_t srlz_eval_workspace_bytes(int B, int state_dim) { return eval_work_layout(B, state_dim).total; }

int srlz_eval_pack(const srlz_net* net, float* epack, void* stream) {
    if (net == nullptr || epack == nullptr) { set_error("srlz_eval_pack: null argument"); return SRLZ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const EvalPack ep = eval_pack_layout(net->state_dim);
    auto fold = [&](int i, int per_co) {
        const srlz_bn& bn = net->enc_bn[i];
        return fold_bn(net->enc_w[i], bn.weight, bn.bias, bn.running_mean, bn.running_var, epack + ep.wtmp, epack + ep.bias[i], per_co, st);
    };
    RC(fold(0, 147));
    RC(pack_enc0_rows_bf16(epack + ep.wtmp, epack + ep.enc0_rb, st));
    for (int i = 0; i < 2; ++i) {
        RC(fold(1 + i, 576));
        RC(pack_conv_w(epack + ep.wtmp, epack + ep.ftmp, epack + ep.dtmp, 9, 0, st));
        RC(pack_conv_w_bf16(epack + ep.ftmp, epack + ep.enc_fb[i], 9, st));
    }
    RC(permute_fc(net->fc_enc_w[0], epack + ep.fc_enc, net->state_dim, 1, 0, 0, st));   // AE: states ; VAE: mu (getStates, models/models.py:126-131)
    RC(fill(epack + ep.ones, 1.f, 64, st));
    return 0;
}

int srlz_encode_eval(const srlz_net* net, const float* epack, const float* x, const int32_t* rects, int B, float* states, void* workspace,
                     void* stream) {
    if (net == nullptr || epack == nullptr || x == nullptr || states == nullptr || workspace == nullptr || B <= 0) {
        set_error("srlz_encode_eval: null argument or B <= 0");
        return SRLZ_E_ARG;
    }
    if (B > SRLZ_MAX_BATCH) { set_error("srlz_encode_eval: B exceeds SRLZ_MAX_BATCH (2048 images per call)"); return SRLZ_E_ARG; }
    if (net->state_dim <= 0 || net->state_dim % 4 != 0) { set_error("srlz_encode_eval: state_dim must be a positive multiple of 4"); return SRLZ_E_ARG; }
    if (reinterpret_cast<uintptr_t>(x) & 7) { set_error("srlz_encode_eval: x must be 8-byte aligned"); return SRLZ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int S = net->state_dim;
    const EvalPack ep = eval_pack_layout(S);
    const EvalWork wk = eval_work_layout(B, S);
    char* ws = (char*)workspace;
    auto W = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const float* ones = epack + ep.ones;
    int np = 0;
    GConvArgs e{};
    e.in = x; e.out = W(wk.y1); e.g = ConvGeom{B, 224, 224, 112, 112, 1, 3, 2, 3}; e.epi = EPI_PLAIN; e.mode = 1; e.rects = rects;
    PROF(T_ENC0_FWD, enc0_rows_fwd(e, epack + ep.enc0_rb, &np, st));
    PROF(T_POOL_FWD, bn_relu_pool_fwd(W(wk.y1), ones, epack + ep.bias[0], W(wk.a1), nullptr, B, 112, 112, 56, 56, 1, st));
    GConvArgs c{};
    c.in = W(wk.a1); c.out = W(wk.y2); c.g = ConvGeom{B, 56, 56, 56, 56, 3, 3, 1, 1}; c.epi = EPI_PLAIN;
    PROF(T_ENC4_FWD, conv64(c, epack, ep.enc_fb[0], &np, st));
    PROF(T_POOL_FWD, bn_relu_pool_fwd(W(wk.y2), ones, epack + ep.bias[1], W(wk.a2), nullptr, B, 56, 56, 27, 27, 0, st));
    c.in = W(wk.a2); c.out = W(wk.y3); c.g = ConvGeom{B, 27, 27, 14, 14, 3, 3, 2, 1};
    PROF(T_ENC8_FWD, conv64(c, epack, ep.enc_fb[1], &np, st));
    PROF(T_POOL_FWD, bn_relu_pool_fwd(W(wk.y3), ones, epack + ep.bias[2], W(wk.a3), nullptr, B, 14, 14, 6, 6, 0, st));
    PROF(T_FC_FWD, sgemm_splitk(W(wk.a3), 2304, 1, epack + ep.fc_enc, 1, 2304, states, S, 1, net->fc_enc_b[0], B, S, 2304, 0, W(wk.tmpw), (size_t)2304 * S, st));
    return 0;
}

int srlz_decode(const srlz_net* net, const float* wpack, const float* z, int B, int training, float* decoded, void* saved,
                void* workspace, void* stream) {
    if (net == nullptr || wpack == nullptr || z == nullptr || decoded == nullptr || saved == nullptr || workspace == nullptr || B <= 0) {
        set_error("srlz_decode: null argument or B <= 0");
        return SRLZ_E_ARG;
    }
    if (B > SRLZ_MAX_BATCH) { set_error("srlz_decode: B exceeds SRLZ_MAX_BATCH (2048 images per call)"); return SRLZ_E_ARG; }
    if (net->state_dim <= 0 || net->state_dim % 4 != 0) { set_error("srlz_decode: state_dim must be a positive multiple of 4"); return SRLZ_E_ARG; }
    if (reinterpret_cast<uintptr_t>(decoded) & 7) { set_error("srlz_decode: decoded must be 8-byte aligned"); return SRLZ_E_ARG; }
    return forward_impl(net, wpack, nullptr, nullptr, nullptr, B, training, nullptr, nullptr, decoded, nullptr, nullptr, (char*)saved,
                        (char*)workspace, (cudaStream_t)stream, z);
}

int srlz_decode_backward(const srlz_net* net, const float* wpack, const srlz_net_grads* grads, int accumulate, int B, int training,
                         const float* g_decoded, float* g_z, void* saved, void* workspace, void* stream) {
    if (net == nullptr || wpack == nullptr || grads == nullptr || g_decoded == nullptr || g_z == nullptr || saved == nullptr ||
        workspace == nullptr || B <= 0) {
        set_error("srlz_decode_backward: null argument or B <= 0");
        return SRLZ_E_ARG;
    }
    if (B > SRLZ_MAX_BATCH) { set_error("srlz_decode_backward: B exceeds SRLZ_MAX_BATCH (2048 images per call)"); return SRLZ_E_ARG; }
    if (reinterpret_cast<uintptr_t>(g_decoded) & 15) { set_error("srlz_decode_backward: g_decoded must be 16-byte aligned"); return SRLZ_E_ARG; }
    return backward_impl(net, wpack, grads, accumulate, nullptr, nullptr, nullptr, B, training, 1, g_decoded, nullptr, nullptr, 0.f,
                         nullptr, nullptr, 0.f, (char*)saved, (char*)workspace, (cudaStream_t)stream, g_z);
}

int srlz_sse(const float* a, const float* b, int64_t n, float out_scale, float* out, void* workspace, void* stream) {
    int np = 0;
    float* part = reinterpret_cast<float*>(workspace);
    RC(sse_partials(a, b, n, part, &np, (cudaStream_t)stream));
    return sum_partials(part, np, out_scale, out, 0, (cudaStream_t)stream);
}

int srlz_mse_grad(const float* a, const float* b, int64_t n, float coef, float* g, void* stream) {
    return mse_grad(a, b, n, coef, g, (cudaStream_t)stream);
}

int srlz_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   int step, void* stream) {
    if (step < 1) { set_error("srlz_adam_step: step must be >= 1"); return SRLZ_E_ARG; }
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    cudaStream_t st = (cudaStream_t)stream;
    PROF(T_ADAM, adam_step(p, g, m, v, n, lr, beta1, beta2, eps, (float)bc1, (float)bc2, st));
    return 0;
}

size_t srlz_op_wgrad64_workspace_bytes(int B, int BH, int BW, int SH, int SW, int K, int stride, int pad) {
    return gwgrad64_partial_floats(ConvGeom{B, BH, BW, SH, SW, K, K, stride, pad}) * sizeof(float);
}

#ifdef SRLZ_DEV
/* development builds only (build.py --dev): clock64 timeline of CTA 0 of one call site (64x16 int64 device buffer) */
void srlz_dev_set_debug_buffer(void* p, int site) {
    g_dbg = reinterpret_cast<long long*>(p);
    g_dbg_site = site;
}
#endif

int srlz_op_pack_conv_w_bf16(const float* pack_f32, void* dst, int ntaps, void* stream) {
    return pack_conv_w_bf16(pack_f32, dst, ntaps, (cudaStream_t)stream);
}

/* the product dispatch of a 64->64 layer site (conv64 above): forward / dgrad of Conv2d(64,64,3) and ConvTranspose2d(64,64,3,2) */
int srlz_op_conv64(const float* in, const void* wbf, const float* bias, const float* in_scale, const float* in_shift, float* out,
                   int B, int BH, int BW, int SH, int SW, int K, int stride, int pad, int transposed, float* stats_partials,
                   int* n_partials, void* stream) {
    if (in == nullptr || wbf == nullptr || out == nullptr || B <= 0 || K != 3) { set_error("srlz_op_conv64: bad argument"); return SRLZ_E_ARG; }
    GConvArgs a{};
    a.in = in; a.bias = bias; a.in_scale = in_scale; a.in_shift = in_shift; a.out = out;
    a.partials = stats_partials; a.g = ConvGeom{B, BH, BW, SH, SW, K, K, stride, pad}; a.transposed = transposed;
    a.epi = stats_partials != nullptr ? EPI_STATS : EPI_PLAIN;
    return conv64(a, reinterpret_cast<const float*>(wbf), 0, n_partials, (cudaStream_t)stream);
}

int srlz_op_sgemm(const float* A, int64_t sa_i, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_j, float* C, int64_t sc_i,
                  int64_t sc_j, const float* bias, int M, int N, int K, int accumulate, void* stream) {
    if (A == nullptr || B == nullptr || C == nullptr || M <= 0 || N <= 0 || K <= 0) { set_error("srlz_op_sgemm: bad argument"); return SRLZ_E_ARG; }
    return sgemm(A, sa_i, sa_k, B, sb_k, sb_j, C, sc_i, sc_j, bias, M, N, K, accumulate, (cudaStream_t)stream);
}

/* weight gradient of a 64->64 3x3 layer site (the halo-tile tcgen05 kernel) */
int srlz_op_wgrad64(const float* big, const float* small, const float* dense_scale, const float* dense_shift, float* grad_out,
                    int B, int BH, int BW, int SH, int SW, int K, int stride, int pad, void* workspace, void* stream) {
    if (big == nullptr || small == nullptr || grad_out == nullptr || workspace == nullptr || B <= 0 || K != 3) { set_error("srlz_op_wgrad64: bad argument"); return SRLZ_E_ARG; }
    GWgradArgs a{};
    a.big = big; a.small = small; a.dense_scale = dense_scale; a.dense_shift = dense_shift;
    a.partials = reinterpret_cast<float*>(workspace); a.g = ConvGeom{B, BH, BW, SH, SW, K, K, stride, pad};
    return wgrad64(a, grad_out, 0, (cudaStream_t)stream);
}

/* ---- op-level entries of the first / last layer kernels (unit tests at layer granularity).  `workspace`: srlz_op_layer_workspace_bytes() ---- */
size_t srlz_op_layer_workspace_bytes(void) {
    size_t m = enc0_rows_wgrad_partial_floats();
    if (dec12_rows_wgrad_partial_floats() > m) m = dec12_rows_wgrad_partial_floats();
    return (m + (size_t)SRLZ_MAX_PART * 128 + 8 * 4096 + 4096) * sizeof(float);
}

/* Conv2d(3,64,7,2,3) forward on the NCHW observation (models/models.py:49): y (B,112,112,64) NHWC pre-BN, optional DAE rectangles
 * applied on load, optional per-CTA BatchNorm sum / sum-of-squares partials [n_partials][128] */
int srlz_op_enc0_fwd(const float* x, const int32_t* rects, const float* w, float* y, float* stats_partials, int* n_partials, int B,
                     void* workspace, void* stream) {
    if (x == nullptr || w == nullptr || y == nullptr || workspace == nullptr || B <= 0 || (reinterpret_cast<uintptr_t>(x) & 7)) { set_error("srlz_op_enc0_fwd: bad argument"); return SRLZ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    float* img = reinterpret_cast<float*>(workspace);
    RC(pack_enc0_rows_bf16(w, img, st));
    GConvArgs e{};
    e.in = x; e.out = y; e.partials = stats_partials; e.g = ConvGeom{B, 224, 224, 112, 112, 1, 3, 2, 3};
    e.epi = stats_partials != nullptr ? EPI_STATS : EPI_PLAIN; e.mode = 1; e.rects = rects;
    return enc0_rows_fwd(e, img, n_partials, st);
}

/* weight gradient of that layer: grad_w (64,3,7,7) = d/dW sum(y * dy), dy (B,112,112,64) NHWC */
int srlz_op_enc0_wgrad(const float* x, const int32_t* rects, const float* dy, float* grad_w, int B, void* workspace, void* stream) {
    if (x == nullptr || dy == nullptr || grad_w == nullptr || workspace == nullptr || B <= 0 || (reinterpret_cast<uintptr_t>(x) & 7)) { set_error("srlz_op_enc0_wgrad: bad argument"); return SRLZ_E_ARG; }
    GWgradArgs wg{};
    wg.big = x; wg.small = dy; wg.partials = reinterpret_cast<float*>(workspace); wg.g = ConvGeom{B, 224, 224, 112, 112, 7, 7, 2, 3}; wg.mode = 1; wg.rects = rects;
    return enc0_rows_wgrad(wg, grad_w, 0, (cudaStream_t)stream);
}

/* ConvTranspose2d(64,3,4,2) + bias on relu(y7*scale+shift) -> decoded (B,3,224,224) NCHW (models/models.py:82); with `target`,
 * sse_out[0] = sum (decoded-target)^2 (losses/losses.py:172-181,210-211) */
int srlz_op_dec12_fwd(const float* y7, const float* scale, const float* shift, const float* w, const float* bias, float* decoded,
                      const float* target, float* sse_out, int B, void* workspace, void* stream) {
    if (y7 == nullptr || scale == nullptr || shift == nullptr || w == nullptr || bias == nullptr || decoded == nullptr || workspace == nullptr || B <= 0 ||
        ((reinterpret_cast<uintptr_t>(decoded) | reinterpret_cast<uintptr_t>(target)) & 7)) { set_error("srlz_op_dec12_fwd: bad argument"); return SRLZ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    float* img = reinterpret_cast<float*>(workspace);
    float* ssep = img + 4096;
    RC(pack_dec12_fwd_bf16(w, img, st));
    GConvArgs d{};
    d.in = y7; d.in_scale = scale; d.in_shift = shift; d.bias = bias; d.out = decoded; d.aux2 = target;
    d.partials = target != nullptr ? ssep : nullptr; d.g = ConvGeom{B, 224, 224, 111, 111, 4, 4, 2, 0}; d.transposed = 1; d.epi = EPI_DEC12;
    int np = 0;
    RC(dec12_rows_fwd(d, img, &np, st));
    if (target != nullptr && sse_out != nullptr) RC(sum_partials(ssep, np, 1.f, sse_out, 0, st));
    return 0;
}

/* backward of that layer: gradient source = g_decoded, or coef*(decoded-target) when g_decoded is NULL.  grad_w (64,3,4,4), grad_b (3);
 * dz (B,111,111,64) = ReLU-masked gradient w.r.t. the BatchNorm output; bn_partials [n_partials][128] = per-CTA sum dz / sum dz*xhat */
int srlz_op_dec12_bwd(const float* y7, const float* scale, const float* shift, const float* mean, const float* invstd, const float* w,
                      const float* g_decoded, const float* decoded, const float* target, float coef, float* grad_w, float* grad_b,
                      float* dz, float* bn_partials, int* n_partials, int B, void* workspace, void* stream) {
    if (y7 == nullptr || scale == nullptr || shift == nullptr || mean == nullptr || invstd == nullptr || w == nullptr || grad_w == nullptr ||
        grad_b == nullptr || dz == nullptr || bn_partials == nullptr || workspace == nullptr || B <= 0 ||
        (g_decoded == nullptr && (decoded == nullptr || target == nullptr)) ||
        ((reinterpret_cast<uintptr_t>(g_decoded) | reinterpret_cast<uintptr_t>(decoded) | reinterpret_cast<uintptr_t>(target)) & 7)) { set_error("srlz_op_dec12_bwd: bad argument"); return SRLZ_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    float* pk = reinterpret_cast<float*>(workspace);
    float* img = pk + 4096;
    float* wpart = img + 4096;
    RC(pack_dec12_dgrad(w, pk, st));
    RC(pack_conv_w_bf16(pk, img, 1, st));
    GWgradArgs wg{};
    wg.big = g_decoded != nullptr ? g_decoded : decoded; wg.small = y7; wg.dense_scale = scale; wg.dense_shift = shift;
    wg.partials = wpart; wg.g = ConvGeom{B, 224, 224, 111, 111, 4, 4, 2, 0}; wg.mode = 2;
    wg.aux0 = g_decoded; wg.aux1 = decoded; wg.aux2 = target; wg.coef = coef;
    RC(dec12_rows_wgrad(wg, grad_w, grad_b, 0, st));
    GConvArgs dg{};
    dg.in = g_decoded != nullptr ? g_decoded : decoded; dg.out = dz; dg.partials = bn_partials;
    dg.g = ConvGeom{B, 224, 224, 111, 111, 1, 1, 2, 0}; dg.transposed = 0; dg.epi = EPI_MASK_BNBWD; dg.mode = 2;
    dg.e_ypre = y7; dg.e_scale = scale; dg.e_shift = shift; dg.e_mean = mean; dg.e_invstd = invstd;
    dg.aux0 = g_decoded; dg.aux1 = decoded; dg.aux2 = target; dg.coef = coef;
    return gconv64_tc(dg, img, n_partials, st);
}

/* BatchNorm (scale / shift) + ReLU + MaxPool2d(3, 2, pad) forward of one pooled encoder stage (models/models.py:50-52,55-57,60-62):
 * y (B,H,W,64) NHWC -> out (B,PH,PW,64), argmax (same shape, uint8: first maximal tap ky*3+kx in scan order) or NULL */
int srlz_op_bn_relu_pool(const float* y, const float* scale, const float* shift, float* out, uint8_t* argmax, int B, int H, int W,
                         int PH, int PW, int pad, void* stream) {
    if (y == nullptr || scale == nullptr || shift == nullptr || out == nullptr || B <= 0) { set_error("srlz_op_bn_relu_pool: bad argument"); return SRLZ_E_ARG; }
    return bn_relu_pool_fwd(y, scale, shift, out, argmax, B, H, W, PH, PW, pad, (cudaStream_t)stream);
}

int srlz_op_pack_conv_w(const float* w, float* fwd_pack, float* dgrad_pack, int ntaps, int transposed_conv, void* stream) {
    return pack_conv_w(w, fwd_pack, dgrad_pack, ntaps, transposed_conv, (cudaStream_t)stream);
}

}  // extern "C"

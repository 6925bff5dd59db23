"""Thin op-level wrappers over libsrlz entry points (used by the heads' stand-alone API and by the tests)."""
import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr


def sgemm(a, b, out, bias=None, trans_a=False, trans_b=False, accumulate=False):
    """out (+)= op(a) @ op(b) + bias   (row-major 2-D float32 CUDA tensors; op = transpose when trans_*)."""
    for t in (a, b, out):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise RuntimeError("sgemm expects 2-D float32 CUDA tensors")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    Kb, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != Kb or tuple(out.shape) != (M, N):
        raise RuntimeError("sgemm shape mismatch: %s x %s -> %s" % (tuple(a.shape), tuple(b.shape), tuple(out.shape)))
    sa = (a.stride(1), a.stride(0)) if trans_a else (a.stride(0), a.stride(1))
    sb = (b.stride(1), b.stride(0)) if trans_b else (b.stride(0), b.stride(1))
    check(lib.srlz_op_sgemm(ptr(a), sa[0], sa[1], ptr(b), sb[0], sb[1], ptr(out), out.stride(0), out.stride(1), ptr(bias),
                            M, N, K, int(accumulate), stream_ptr()), "sgemm")
    return out


def sse(a, b, scale=1.0):
    """scale * sum((a-b)^2) as a 1-element CUDA tensor (losses/losses.py:172-181,210-211)."""
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    ws = torch.empty(2048, dtype=torch.float32, device=a.device)
    check(lib.srlz_sse(ptr(a), ptr(b), a.numel(), float(scale), ptr(out), ptr(ws), stream_ptr()), "sse")
    return out


def mse_grad(a, b, coef):
    """coef * (a - b)"""
    a, b = a.contiguous(), b.contiguous()
    g = torch.empty_like(a)
    check(lib.srlz_mse_grad(ptr(a), ptr(b), a.numel(), float(coef), ptr(g), stream_ptr()), "mse_grad")
    return g


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib.srlz_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, int(step), stream_ptr()), "adam")


def pack_conv_w(w, transposed_conv):
    ntaps = w.shape[2] * w.shape[3]
    f = torch.empty(ntaps, 64, 64, dtype=torch.float32, device=w.device)
    d = torch.empty(ntaps, 64, 64, dtype=torch.float32, device=w.device)
    check(lib.srlz_op_pack_conv_w(ptr(w.contiguous()), ptr(f), ptr(d), ntaps, int(transposed_conv), stream_ptr()), "pack_conv_w")
    return f, d


def pack_conv_w_bf16(pack_f32):
    """fp32 [tap][k][n] pack -> bf16 hi/lo SWIZZLE_128B image (uint8 tensor, 16 KB per tap) for the tcgen05 kernels"""
    ntaps = pack_f32.shape[0]
    dst = torch.empty(ntaps * 16384, dtype=torch.uint8, device=pack_f32.device)
    check(lib.srlz_op_pack_conv_w_bf16(ptr(pack_f32), ptr(dst), ntaps, stream_ptr()), "pack_conv_w_bf16")
    return dst


def conv64(x_nhwc, wbf, out_nhwc, big_hw, small_hw, k, stride, pad, transposed, bias=None, in_scale=None, in_shift=None,
           want_stats=False):
    """Forward / dgrad of a 64->64 3x3 layer site through the product dispatch (csrc/api.cu conv64: halo-tile kernel,
    stride-2 row kernel or per-tap pipeline by geometry).  Returns (out, stats[128] or None)."""
    B = x_nhwc.shape[0]
    part = torch.zeros(1184, 128, dtype=torch.float32, device=x_nhwc.device) if want_stats else None
    n = C.c_int(0)
    check(lib.srlz_op_conv64(ptr(x_nhwc), ptr(wbf), ptr(bias), ptr(in_scale), ptr(in_shift), ptr(out_nhwc), B, big_hw[0],
                             big_hw[1], small_hw[0], small_hw[1], k, stride, pad, int(transposed), ptr(part), C.byref(n),
                             stream_ptr()), "conv64")
    stats = part[:n.value].double().sum(0).float() if want_stats else None
    return out_nhwc, stats


def wgrad64(big, small, big_hw, small_hw, k, stride, pad, dense_scale=None, dense_shift=None):
    """-> weight gradient of a 64->64 3x3 layer site in torch layout (64,64,k,k) indexed [c_dense][c_gathered][ky][kx]."""
    B = big.shape[0]
    nbytes = lib.srlz_op_wgrad64_workspace_bytes(B, big_hw[0], big_hw[1], small_hw[0], small_hw[1], k, stride, pad)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=big.device)
    out = torch.empty(64, 64, k, k, dtype=torch.float32, device=big.device)
    check(lib.srlz_op_wgrad64(ptr(big), ptr(small), ptr(dense_scale), ptr(dense_shift), ptr(out), B, big_hw[0], big_hw[1],
                              small_hw[0], small_hw[1], k, stride, pad, ptr(ws), stream_ptr()), "wgrad64")
    return out


def _layer_ws(dev):
    return torch.empty(lib.srlz_op_layer_workspace_bytes(), dtype=torch.uint8, device=dev)


def enc0_fwd(x, w, rects=None, want_stats=False):
    """Conv2d(3,64,7,2,3) on the NCHW observation (models/models.py:49) -> (y (B,112,112,64) NHWC, stats[128] or None)"""
    B = x.shape[0]
    y = torch.empty(B, 112, 112, 64, dtype=torch.float32, device=x.device)
    part = torch.zeros(1184, 128, dtype=torch.float32, device=x.device) if want_stats else None
    n = C.c_int(0)
    check(lib.srlz_op_enc0_fwd(ptr(x), ptr(rects), ptr(w.contiguous()), ptr(y), ptr(part), C.byref(n), B, ptr(_layer_ws(x.device)),
                               stream_ptr()), "enc0_fwd")
    return y, (part[:n.value].double().sum(0).float() if want_stats else None)


def enc0_wgrad(x, dy_nhwc, rects=None):
    B = x.shape[0]
    g = torch.empty(64, 3, 7, 7, dtype=torch.float32, device=x.device)
    check(lib.srlz_op_enc0_wgrad(ptr(x), ptr(rects), ptr(dy_nhwc), ptr(g), B, ptr(_layer_ws(x.device)), stream_ptr()), "enc0_wgrad")
    return g


def dec12_fwd(y7_nhwc, scale, shift, w, bias, target=None):
    """ConvTranspose2d(64,3,4,2) + bias on relu(y7*scale+shift) (models/models.py:82) -> (decoded NCHW, sse or None)"""
    B = y7_nhwc.shape[0]
    out = torch.empty(B, 3, 224, 224, dtype=torch.float32, device=y7_nhwc.device)
    sse_out = torch.zeros(1, dtype=torch.float32, device=y7_nhwc.device) if target is not None else None
    check(lib.srlz_op_dec12_fwd(ptr(y7_nhwc), ptr(scale), ptr(shift), ptr(w.contiguous()), ptr(bias), ptr(out), ptr(target), ptr(sse_out),
                                B, ptr(_layer_ws(y7_nhwc.device)), stream_ptr()), "dec12_fwd")
    return out, sse_out


def dec12_bwd(y7_nhwc, scale, shift, mean, invstd, w, g_decoded=None, decoded=None, target=None, coef=0.0):
    """-> (grad_w (64,3,4,4), grad_b (3), dz (B,111,111,64), bn sums [128] = sum dz | sum dz*xhat)"""
    B, dev = y7_nhwc.shape[0], y7_nhwc.device
    gw = torch.empty(64, 3, 4, 4, dtype=torch.float32, device=dev)
    gb = torch.empty(3, dtype=torch.float32, device=dev)
    dz = torch.empty(B, 111, 111, 64, dtype=torch.float32, device=dev)
    part = torch.zeros(1184, 128, dtype=torch.float32, device=dev)
    n = C.c_int(0)
    check(lib.srlz_op_dec12_bwd(ptr(y7_nhwc), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), ptr(w.contiguous()), ptr(g_decoded),
                                ptr(decoded), ptr(target), float(coef), ptr(gw), ptr(gb), ptr(dz), ptr(part), C.byref(n), B,
                                ptr(_layer_ws(dev)), stream_ptr()), "dec12_bwd")
    return gw, gb, dz, part[:n.value].double().sum(0).float()


def preprocess_u8(frames, out=None):
    """uint8 RGB frames (B,224,224,3) HWC -> the normalised (B,3,224,224) float32 tensor the reference's loader delivers
    (preprocessing/utils.py:20-32, preprocessing/data_loader.py:255), bit-exact, on the device."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous() and tuple(frames.shape[1:]) == (224, 224, 3)):
        raise RuntimeError("preprocess_u8 expects contiguous uint8 CUDA frames of shape (B,224,224,3)")
    B = frames.shape[0]
    if out is None:
        out = torch.empty(B, 3, 224, 224, dtype=torch.float32, device=frames.device)
    check(lib.srlz_preprocess_u8(ptr(frames), ptr(out), B, stream_ptr()), "preprocess_u8")
    return out


def bn_relu_pool(y_nhwc, scale, shift, pad, want_argmax=True):
    """BatchNorm scale / shift + ReLU + MaxPool2d(3, 2, pad) of a pooled encoder stage (models/models.py:50-52) -> (out, argmax)"""
    B, H, W, _ = y_nhwc.shape
    PH, PW = (H + 2 * pad - 3) // 2 + 1, (W + 2 * pad - 3) // 2 + 1
    out = torch.empty(B, PH, PW, 64, dtype=torch.float32, device=y_nhwc.device)
    am = torch.empty(B, PH, PW, 64, dtype=torch.uint8, device=y_nhwc.device) if want_argmax else None
    check(lib.srlz_op_bn_relu_pool(ptr(y_nhwc), ptr(scale), ptr(shift), ptr(out), ptr(am), B, H, W, PH, PW, pad, stream_ptr()), "bn_relu_pool")
    return out, am

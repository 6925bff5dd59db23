"""
Host-side mirror of the reference's model interface for the hot path: `B200SRLModules` is a drop-in for
`SRLModules` (models/modules.py:17-100) restricted to model_type="custom_cnn" with "autoencoder", "dae" or
"vae" in `losses` (+ optional "forward" / "inverse" heads).  Same constructor signature, same methods and
tuple orders (AE: (states, decoded) models/models.py:114 ; VAE: (decoded, mu, logvar) models/models.py:176),
same state_dict key set and the same parameter-initialisation RNG order (heads -> conv stacks -> FCs,
models/modules.py:37-49), so `srl_model.pth` files and seeds are interchangeable.

The torch.nn layers created here are parameter CONTAINERS only: every forward / backward runs in libsrlz
(hand-written sm_100a CUDA, include/srlz.h).  There is no PyTorch fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import MAX_BATCH, SrlzBn, SrlzNet, SrlzNetGrads, check, lib, ptr, stream_ptr

IMG = 224          # preprocessing/preprocess.py:7-8
EVAL_CHUNK = 256   # models/learner.py:33 (BATCH_SIZE): call size of a split prediction batch
FLAT = 64 * 6 * 6  # models/autoencoders.py:95

_ENC = ((0, 3, 7, 2, 3), (4, 64, 3, 1, 1), (8, 64, 3, 2, 1))   # (index, cin, k, stride, pad)  models/models.py:49,54,59
_DEC = ((0, 64, 3), (3, 64, 3), (6, 64, 3), (9, 64, 3), (12, 3, 4))  # (index, cout, k)        models/models.py:66-82


def _conv_stack_containers():
    """encoder_conv / decoder_conv with the reference's Sequential indices (ReLU / MaxPool slots kept so that
    state_dict keys line up: encoder_conv.{0,1,4,5,8,9}, decoder_conv.{0,1,3,4,6,7,9,10,12})."""
    enc, dec = [], []
    for (_, cin, k, s, p), pool_pad in zip(_ENC, (1, 0, 0)):
        enc += [nn.Conv2d(cin, 64, k, s, p, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                nn.MaxPool2d(3, 2, pool_pad)]
    for idx, cout, k in _DEC:
        dec.append(nn.ConvTranspose2d(64, cout, k, 2))
        if idx != 12:
            dec += [nn.BatchNorm2d(64), nn.ReLU(True)]
    return nn.Sequential(*enc), nn.Sequential(*dec)


class _ConvNet(nn.Module):
    """Parameter container mirroring CNNAutoEncoder (models/autoencoders.py:84-118) / CNNVAE (models/vae.py:43-75)."""

    def __init__(self, state_dim, is_vae):
        super().__init__()
        self.is_vae = is_vae
        self.state_dim = state_dim
        self.encoder_conv, self.decoder_conv = _conv_stack_containers()
        self._scratch = _Scratch()
        if is_vae:
            self.encoder_fc1 = nn.Linear(FLAT, state_dim)
            self.encoder_fc2 = nn.Linear(FLAT, state_dim)
            self.decoder_fc = nn.Sequential(nn.Linear(state_dim, FLAT))
        else:
            self.encoder_fc = nn.Sequential(nn.Linear(FLAT, state_dim))
            self.decoder_fc = nn.Sequential(nn.Linear(state_dim, FLAT))

    def forward(self, x):  # pragma: no cover - never used: the owner module dispatches to libsrlz
        raise RuntimeError("container only; call the owning B200SRLModules")

    def decode(self, z):
        """BaseModelAutoEncoder.decode / BaseModelVAE.decode (models/autoencoders.py:111-118, models/vae.py:68-75): called on
        the inner model by evaluation/enjoy_latent.py:35,136 and by the split forward passes.  z (B,S) -> (B,3,224,224)."""
        params = [p for _, _, p in self.slots()]
        return _Decode.apply(self, z.contiguous(), *params)

    # ordered list of the learnable tensors in the slot order of srlz_net / srlz_net_grads
    def slots(self):
        e, d = self.encoder_conv, self.decoder_conv
        out = [("enc_w", i, e[idx].weight) for i, idx in enumerate((0, 4, 8))]
        out += [("enc_bn_w", i, e[idx].weight) for i, idx in enumerate((1, 5, 9))]
        out += [("enc_bn_b", i, e[idx].bias) for i, idx in enumerate((1, 5, 9))]
        out += [("dec_w", i, d[idx].weight) for i, idx in enumerate((0, 3, 6, 9, 12))]
        out += [("dec_b", i, d[idx].bias) for i, idx in enumerate((0, 3, 6, 9, 12))]
        out += [("dec_bn_w", i, d[idx].weight) for i, idx in enumerate((1, 4, 7, 10))]
        out += [("dec_bn_b", i, d[idx].bias) for i, idx in enumerate((1, 4, 7, 10))]
        if self.is_vae:
            out += [("fc_enc_w", 0, self.encoder_fc1.weight), ("fc_enc_w", 1, self.encoder_fc2.weight),
                    ("fc_enc_b", 0, self.encoder_fc1.bias), ("fc_enc_b", 1, self.encoder_fc2.bias)]
        else:
            out += [("fc_enc_w", 0, self.encoder_fc[0].weight), ("fc_enc_b", 0, self.encoder_fc[0].bias)]
        out += [("fc_dec_w", None, self.decoder_fc[0].weight), ("fc_dec_b", None, self.decoder_fc[0].bias)]
        return out

    def net_struct(self):
        """srlz_net filled with the CURRENT device pointers (cheap; rebuilt per call so .to()/load_state_dict are safe)."""
        n = SrlzNet()
        n.is_vae, n.state_dim = int(self.is_vae), int(self.state_dim)
        e, d = self.encoder_conv, self.decoder_conv

        def bn(dst, m):
            dst.weight, dst.bias = m.weight.data_ptr(), m.bias.data_ptr()
            dst.running_mean, dst.running_var = m.running_mean.data_ptr(), m.running_var.data_ptr()
            dst.num_batches_tracked = m.num_batches_tracked.data_ptr()

        for i, idx in enumerate((0, 4, 8)):
            n.enc_w[i] = e[idx].weight.data_ptr()
            bn(n.enc_bn[i], e[idx + 1])
        for i, idx in enumerate((0, 3, 6, 9, 12)):
            n.dec_w[i], n.dec_b[i] = d[idx].weight.data_ptr(), d[idx].bias.data_ptr()
            if idx != 12:
                bn(n.dec_bn[i], d[idx + 1])
        if self.is_vae:
            fcs = (self.encoder_fc1, self.encoder_fc2)
        else:
            fcs = (self.encoder_fc[0],)
        for i, fc in enumerate(fcs):
            n.fc_enc_w[i], n.fc_enc_b[i] = fc.weight.data_ptr(), fc.bias.data_ptr()
        n.fc_dec_w, n.fc_dec_b = self.decoder_fc[0].weight.data_ptr(), self.decoder_fc[0].bias.data_ptr()
        return n


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: the B200 path has no CPU fallback" % what)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError("%s must be contiguous float32" % what)


class _Scratch:
    """Caller-side ownership of the blocks libsrlz needs (it never allocates): packed weights + workspace."""

    def __init__(self):
        self.wpack = None
        self.pack_key = None
        self.ws = None
        self.ws_key = None

    def get_pack(self, net, device):
        n = lib.srlz_pack_floats(int(net.is_vae), int(net.state_dim))
        if self.wpack is None or self.wpack.numel() != n or self.wpack.device != device:
            self.wpack = torch.empty(n, dtype=torch.float32, device=device)
            self.pack_key = None
        return self.wpack

    def packed(self, cn, net_struct, device):
        """the kernel-layout weight pack, rebuilt (srlz_pack_weights, ~35 launches) only when a parameter changed since the
        last pack: the key tracks every in-place update torch knows of (optimizer.step, load_state_dict) and the data pointers"""
        wpack = self.get_pack(cn, device)
        key = (getattr(cn, "_weights_version", 0),) + tuple((p.data_ptr(), p._version) for _, _, p in cn.slots())
        if self.pack_key != key:
            check(lib.srlz_pack_weights(C.byref(net_struct), ptr(wpack), stream_ptr()), "pack_weights")
            self.pack_key = key
        return wpack

    def get_ws(self, B, net, device):
        key = (B, int(net.state_dim), int(net.is_vae), device)
        if self.ws_key != key:
            self.ws = torch.empty(lib.srlz_workspace_bytes(B, int(net.state_dim), int(net.is_vae)), dtype=torch.uint8, device=device)
            self.ws_key = key
        return self.ws


def _grad_views(cn, dev):
    """a zeroed flat gradient buffer + per-parameter views in slot order + the srlz_net_grads struct pointing at them"""
    slots = cn.slots()
    flat = torch.zeros(sum(p.numel() for _, _, p in slots), dtype=torch.float32, device=dev)
    grads = SrlzNetGrads()
    views, off = [], 0
    for name, idx, p in slots:
        v = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
        views.append(v)
        if idx is None:
            setattr(grads, name, v.data_ptr())
        else:
            getattr(grads, name)[idx] = v.data_ptr()
    return slots, views, grads


_DEC_NAMES = ("dec_w", "dec_b", "dec_bn_w", "dec_bn_b", "fc_dec_w", "fc_dec_b")


class _Decode(torch.autograd.Function):
    """decoder-only call (srlz_decode / srlz_decode_backward)"""

    @staticmethod
    def forward(ctx, cn, z, *params):
        _require_cuda(z, "latent states")
        B, S, dev = z.shape[0], cn.state_dim, z.device
        if tuple(z.shape) != (B, S):
            raise RuntimeError("expected latent states of shape (B,%d), got %s" % (S, tuple(z.shape)))
        net = cn.net_struct()
        wpack = cn._scratch.packed(cn, net, dev)
        ws = cn._scratch.get_ws(B, cn, dev)
        saved = torch.empty(lib.srlz_saved_bytes(B, S, int(cn.is_vae)), dtype=torch.uint8, device=dev)
        decoded = torch.empty(B, 3, IMG, IMG, dtype=torch.float32, device=dev)
        check(lib.srlz_decode(C.byref(net), ptr(wpack), ptr(z), B, int(cn.training), ptr(decoded), ptr(saved), ptr(ws), stream_ptr()), "decode")
        ctx.cn, ctx.B, ctx.training, ctx.saved_block = cn, B, cn.training, saved
        return decoded

    @staticmethod
    def backward(ctx, g_dec):
        cn = ctx.cn
        dev = g_dec.device
        slots, views, grads = _grad_views(cn, dev)
        net = cn.net_struct()
        wpack = cn._scratch.get_pack(cn, dev)
        ws = cn._scratch.get_ws(ctx.B, cn, dev)
        gz = torch.empty(ctx.B, cn.state_dim, dtype=torch.float32, device=dev)
        check(lib.srlz_decode_backward(C.byref(net), ptr(wpack), C.byref(grads), 0, ctx.B, int(ctx.training), ptr(g_dec.contiguous()),
                                       ptr(gz), ptr(ctx.saved_block), ptr(ws), stream_ptr()), "decode_backward")
        ctx.saved_block = None
        views = [v if name in _DEC_NAMES else None for (name, _, _), v in zip(slots, views)]
        return (None, gz) + tuple(views)


class _ModelCall(torch.autograd.Function):
    """One `model(x)` of the reference (models/modules.py:82-85) through srlz_forward / srlz_backward."""

    @staticmethod
    def forward(ctx, owner, x, rects, eps, want_decoder, *params):
        cn = owner.model
        _require_cuda(x, "observations")
        B = x.shape[0]
        if tuple(x.shape[1:]) != (3, IMG, IMG):
            raise RuntimeError("expected observations of shape (B,3,%d,%d), got %s" % (IMG, IMG, tuple(x.shape)))
        dev = x.device
        S = cn.state_dim
        training = owner.training
        net = cn.net_struct()
        wpack = cn._scratch.packed(cn, net, dev)
        ws = cn._scratch.get_ws(B, cn, dev)
        saved = torch.empty(lib.srlz_saved_bytes(B, S, int(cn.is_vae)), dtype=torch.uint8, device=dev)
        lat = torch.empty(B, S, dtype=torch.float32, device=dev)
        logvar = torch.empty(B, S, dtype=torch.float32, device=dev) if cn.is_vae else None
        decoded = torch.empty(B, 3, IMG, IMG, dtype=torch.float32, device=dev) if want_decoder else None
        if cn.is_vae and training and want_decoder and eps is None:
            # models/models.py:161 -- drawn by torch on the tensor's device with the same call
            eps = torch.empty(B, S, dtype=torch.float32, device=dev).normal_()
        # the squared error against the INPUT rides along for free in the last decoder tile (the AE / VAE losses compare decoded
        # with the very tensor the model was called on, learner.py:393,400,452-468): the loss functions pick it up when they are
        # handed that pair (srl_zoo_b200.losses._SSE) and hand back a coefficient instead of a (B,3,224,224) gradient tensor
        fused = want_decoder and rects is None
        loss_raw = torch.zeros(2, dtype=torch.float32, device=dev) if fused else None
        check(lib.srlz_forward(C.byref(net), ptr(wpack), ptr(x), ptr(rects), ptr(eps), B, int(training), ptr(lat),
                               ptr(logvar), ptr(decoded), ptr(x) if fused else None, ptr(loss_raw), ptr(saved), ptr(ws), stream_ptr()),
              "forward")
        ctx.owner, ctx.B, ctx.training, ctx.want_decoder = owner, B, training, want_decoder
        ctx.saved_block, ctx.x, ctx.rects, ctx.eps = saved, x, rects, eps
        ctx.fused = None
        if fused:
            ctx.fused = {"x_ptr": x.data_ptr(), "x_version": x._version, "shape": tuple(x.shape), "sse": loss_raw, "coef": None,
                         "dummy": torch.zeros(1, dtype=torch.float32, device=dev).expand(B, 3, IMG, IMG), "decoded": None}
            decoded._srlz_fused = ctx.fused
        ctx.set_materialize_grads(False)
        outs = [lat]
        if cn.is_vae:
            outs.append(logvar)
        if want_decoder:
            outs.append(decoded)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        owner = ctx.owner
        cn = owner.model
        if ctx.needs_input_grad[1]:
            # the first layer's input gradient is not part of the path (learner.py never differentiates w.r.t. the
            # observations except through the perceptual loss, which install() keeps on the reference class)
            raise RuntimeError("B200SRLModules does not compute a gradient w.r.t. its input observations")
        g_lat = gouts[0]
        g_logvar = gouts[1] if cn.is_vae else None
        g_dec = gouts[-1] if ctx.want_decoder else None
        dev = ctx.x.device
        n_par = len(cn.slots())
        if g_lat is None and g_logvar is None and g_dec is None:
            return (None,) * (5 + n_par)
        has_decoder = g_dec is not None
        net = cn.net_struct()
        wpack = cn._scratch.get_pack(cn, dev)   # packed in forward; weights are unchanged until optimizer.step()
        ws = cn._scratch.get_ws(ctx.B, cn, dev)
        slots, views, grads = _grad_views(cn, dev)
        cg = lambda t: None if t is None else t.contiguous()
        dec_t, tgt_t, mse_coef = None, None, 0.0
        rec = ctx.fused
        if rec is not None and rec["coef"] is not None and has_decoder:
            # the fused squared-error term: d(total)/d(decoded) = coef * (decoded - x), recomputed inside the last layer's backward
            coef = 2.0 * float(rec["coef"])
            if g_dec.data_ptr() == rec["dummy"].data_ptr() and all(st == 0 for st in g_dec.stride()):
                g_dec, dec_t, tgt_t, mse_coef = None, rec["decoded"], ctx.x, coef      # the only consumer of decoded: no gradient tensor at all
            else:   # decoded has other consumers too: their (materialised) gradient + the fused term, explicitly
                from . import ops
                g_dec = g_dec + ops.mse_grad(rec["decoded"], ctx.x, coef)
            rec["coef"], rec["decoded"] = None, None
        g_lat, g_logvar, g_dec = cg(g_lat), cg(g_logvar), cg(g_dec)
        check(lib.srlz_backward(C.byref(net), ptr(wpack), C.byref(grads), 0, ptr(ctx.x), ptr(ctx.rects), ptr(ctx.eps), ctx.B,
                                int(ctx.training), int(has_decoder), ptr(g_dec), ptr(dec_t), ptr(tgt_t), mse_coef, ptr(g_lat),
                                ptr(g_logvar), 0.0, ptr(ctx.saved_block), ptr(ws), stream_ptr()), "backward")
        ctx.saved_block = None
        if not has_decoder:  # decoder tensors received no gradient in this call
            views = [None if name in _DEC_NAMES else v for (name, _, _), v in zip(slots, views)]
        return (None, None, None, None, None) + tuple(views)


class _Linear(torch.autograd.Function):
    """y = x W^T + b through libsrlz's strided SGEMM (heads: models/forward_inverse.py:30-31,70)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return _sgemm_nt(x, w, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        return _sgemm_nn(g, w), _sgemm_tn(g, x), g.sum(0)


def _sgemm_nt(x, w, b):
    # the heads are tiny (B x 200); they run through torch matmul-free path: srlz_heads covers the fused
    # training step, this helper only serves the reference's stand-alone forwardModel()/inverseModel() calls.
    out = torch.empty(x.shape[0], w.shape[0], dtype=torch.float32, device=x.device)
    from . import ops
    ops.sgemm(x, w, out, bias=b, trans_b=True)
    return out


def _sgemm_nn(g, w):
    from . import ops
    out = torch.empty(g.shape[0], w.shape[1], dtype=torch.float32, device=g.device)
    ops.sgemm(g, w, out)
    return out


def _sgemm_tn(g, x):
    from . import ops
    out = torch.empty(g.shape[1], x.shape[1], dtype=torch.float32, device=g.device)
    ops.sgemm(g, x, out, trans_a=True)
    return out


class _ReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.srlz_relu(ptr(x), ptr(y), x.numel(), stream_ptr()), "relu")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        gx = torch.empty_like(y)
        check(lib.srlz_relu_bwd(ptr(y), ptr(g.contiguous()), ptr(gx), y.numel(), stream_ptr()), "relu_bwd")
        return gx


class _ColMask(torch.autograd.Function):
    """x * mask[None, :] with a 0/1 column mask: what SRLModulesSplit.detachSplit computes (models/modules.py:189-234)"""

    @staticmethod
    def forward(ctx, x, mask):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.srlz_colmask(ptr(x), ptr(mask), ptr(y), x.shape[0], x.shape[1], stream_ptr()), "colmask")
        ctx.mask = mask
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gx = torch.empty_like(g)
        check(lib.srlz_colmask(ptr(g), ptr(ctx.mask), ptr(gx), g.shape[0], g.shape[1], stream_ptr()), "colmask")
        return gx, None


class _CatCols(torch.autograd.Function):
    """th.cat((a, b), dim=1) -- or th.cat((a, encodeOneHot(idx, n)), dim=1) when b is an int64 index column (models/models.py:229-237)"""

    @staticmethod
    def forward(ctx, a, b, n_onehot):
        a = a.contiguous()
        rows, ca = a.shape
        onehot = b.dtype == torch.int64
        if onehot and not (b.is_cuda and b.is_contiguous() and tuple(b.shape) == (rows, 1)):
            raise RuntimeError("actions must be a contiguous int64 CUDA tensor of shape (%d,1)" % rows)
        cb = n_onehot if onehot else b.shape[1]
        out = torch.empty(rows, ca + cb, dtype=torch.float32, device=a.device)
        check(lib.srlz_cat_cols(ptr(a), ca, None if onehot else ptr(b.contiguous()), cb, ptr(b) if onehot else None, ptr(out), rows,
                                stream_ptr()), "cat_cols")
        ctx.ca, ctx.onehot = ca, onehot
        return out

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.ca], (None if ctx.onehot else g[:, ctx.ca:]), None


class _Reparam(torch.autograd.Function):
    """z = eps * exp(0.5 * logvar) + mu  (models/models.py:155-163)"""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        z = torch.empty_like(mu)
        check(lib.srlz_reparam(ptr(mu), ptr(logvar), ptr(eps), ptr(z), mu.numel(), stream_ptr()), "reparam")
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, gz):
        logvar, eps = ctx.saved_tensors
        gz = gz.contiguous()
        gmu, glv = torch.empty_like(gz), torch.empty_like(gz)
        check(lib.srlz_reparam_bwd(ptr(gz), ptr(logvar), ptr(eps), ptr(gmu), ptr(glv), gz.numel(), stream_ptr()), "reparam_bwd")
        return gmu, glv, None


def _mlp(seq, x):
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) of the mlp inverse / reward heads through libsrlz's SGEMM + ReLU kernels"""
    for layer in seq:
        x = _ReLU.apply(x) if isinstance(layer, nn.ReLU) else _Linear.apply(x, layer.weight, layer.bias)
    return x


class B200SRLModules(nn.Module):
    """Drop-in for models.modules.SRLModules (models/modules.py:17-100) on the B200 hot path."""

    def __init__(self, state_dim=2, action_dim=6, cuda=False, model_type="custom_cnn", losses=None,
                 inverse_model_type="linear"):
        super().__init__()
        losses = list(losses) if losses is not None else []
        if model_type != "custom_cnn" or not any(k in losses for k in ("autoencoder", "dae", "vae")):
            raise ValueError("B200SRLModules covers model_type='custom_cnn' with 'autoencoder', 'dae' or 'vae' in losses "
                             "(got model_type=%r, losses=%r); other configurations stay on the reference class" % (model_type, losses))
        if state_dim % 4 != 0:
            raise ValueError("state_dim must be a multiple of 4 for the 128-bit kernels (got %d)" % state_dim)
        self.model_type, self.losses, self.cuda_flag = model_type, losses, cuda
        self.state_dim, self.action_dim = state_dim, action_dim
        # creation order == models/modules.py:37-39 then :42-49 (identical RNG consumption => identical init)
        self.forward_net = nn.Linear(state_dim + action_dim, state_dim)            # forward_inverse.py:16
        if inverse_model_type == "linear":                                         # forward_inverse.py:47-56
            self.inverse_net = nn.Linear(state_dim * 2, action_dim)
        elif inverse_model_type == "mlp":
            self.inverse_net = nn.Sequential(nn.Linear(state_dim * 2, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(),
                                             nn.Linear(128, action_dim))
        else:
            raise ValueError("Unknown model_type for inverse model: {}".format(inverse_model_type))
        self.inverse_model_type = inverse_model_type
        self.reward_net = nn.Sequential(nn.Linear(2 * state_dim, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(),
                                        nn.Linear(16, 2))                          # forward_inverse.py:79-83
        is_vae = not ("autoencoder" in losses or "dae" in losses)                  # modules.py:43-46
        self.model = _ConvNet(state_dim, is_vae)

    # ---- reference API ----
    def forward(self, x):
        """AE/DAE: (states, decoded) ; VAE: (decoded, mu, logvar)   (models/models.py:114,176)"""
        return self._call(x, None, None)

    def forward_masked(self, x, rects, eps=None):
        """DAE fast path: the zero-pixel rectangles (preprocessing/data_loader.py:55-63) are applied inside the
        first encoder tile instead of materialising `noisy_obs`.  rects: (B,4) int32 (h1,h2,w1,w2)."""
        return self._call(x, rects, eps)

    def _call(self, x, rects, eps):
        params = [p for _, _, p in self.model.slots()]
        outs = _ModelCall.apply(self, x.contiguous(), rects, eps, True, *params)
        if self.model.is_vae:
            mu, logvar, decoded = outs
            return decoded, mu, logvar
        states, decoded = outs
        return states, decoded

    def getStates(self, observations):
        """models/models.py:85-90 (AE: encode) / :126-131 (VAE: mu).  Encoder-only pass.  In eval mode without autograd (the
        prediction path, models/learner.py:67-88,570-577) this is the folded inference path: srlz_encode_eval."""
        if not self.training and not torch.is_grad_enabled():
            return self._eval_states(observations)
        params = [p for _, _, p in self.model.slots()]
        outs = _ModelCall.apply(self, observations.contiguous(), None, None, False, *params)
        return outs[0]

    # ---- inference path ----
    def _eval_key(self):
        cn = self.model
        ts = [p for _, _, p in cn.slots()] + list(cn.encoder_conv.buffers())
        return (getattr(self, "_weights_version", 0),) + tuple((t.data_ptr(), t._version) for t in ts)

    def _eval_states(self, x):
        """eval-mode getStates with BatchNorm folded into the conv weights: the fold + pack runs once per weight version (the
        key tracks in-place updates of every tensor and the fused engine's own counter), each batch is then 7 launches"""
        cn = self.model
        _require_cuda(x, "observations")
        x = x.contiguous()
        B, S, dev = x.shape[0], cn.state_dim, x.device
        if tuple(x.shape[1:]) != (3, IMG, IMG):
            raise RuntimeError("expected observations of shape (B,3,%d,%d), got %s" % (IMG, IMG, tuple(x.shape)))
        if B > MAX_BATCH:   # eval mode has no batch statistics: a prediction batch above the library's limit is split into
            # calls of the reference's own minibatch size (models/learner.py:33), each row's result independent of the others
            return torch.cat([self._eval_states(x[i:i + EVAL_CHUNK]) for i in range(0, B, EVAL_CHUNK)])
        net = cn.net_struct()
        key = self._eval_key()
        cache = getattr(self, "_eval_cache", None)
        if cache is None or cache["key"] != key:
            epack = torch.empty(lib.srlz_eval_pack_floats(S), dtype=torch.float32, device=dev)
            check(lib.srlz_eval_pack(C.byref(net), ptr(epack), stream_ptr()), "eval_pack")
            cache = self._eval_cache = {"key": key, "epack": epack, "ws": None, "B": -1}
        if cache["B"] != B:
            cache["ws"] = torch.empty(lib.srlz_eval_workspace_bytes(B, S), dtype=torch.uint8, device=dev)
            cache["B"] = B
        states = torch.empty(B, S, dtype=torch.float32, device=dev)
        check(lib.srlz_encode_eval(C.byref(net), ptr(cache["epack"]), ptr(x), None, B, ptr(states), ptr(cache["ws"]), stream_ptr()), "encode_eval")
        return states

    def forwardModel(self, state, action):
        """models/forward_inverse.py:21-31"""
        cat = _CatCols.apply(state, action, self.action_dim)
        return state + _Linear.apply(cat, self.forward_net.weight, self.forward_net.bias)

    def inverseModel(self, state, next_state):
        """models/forward_inverse.py:62-70 ('linear' and 'mlp' heads, :47-56)"""
        cat = _CatCols.apply(state, next_state, 0)
        if self.inverse_model_type == "linear":
            return _Linear.apply(cat, self.inverse_net.weight, self.inverse_net.bias)
        return _mlp(self.inverse_net, cat)

    def rewardModel(self, state, next_state):
        """models/forward_inverse.py:86-95"""
        return _mlp(self.reward_net, _CatCols.apply(state, next_state, 0))


def split_masks(split_dimensions, state_dim):
    """SRLModulesSplit.detachSplit (models/modules.py:189-234) as one 0/1 column mask per loss name: every split owns the next
    n_dim state columns; a split with n_dim == -1 shares the columns of the split before it; detachSplit(t, index) keeps the
    columns of `index` and zeroes the rest (columns keep their order), so it equals t * mask[index]."""
    masks, start, prev = {}, 0, (0, 0)
    for key, n_dim in split_dimensions.items():
        n_dim = int(n_dim)
        if n_dim == -1 and start > 0:
            rng = prev
        else:
            rng = (start, start + n_dim)
            start += n_dim
            prev = rng
        m = torch.zeros(state_dim, dtype=torch.float32)
        m[rng[0]:rng[1]] = 1.0
        masks[key] = m
    return masks


class B200SRLModulesSplit(B200SRLModules):
    """Drop-in for models.modules.SRLModulesSplit (models/modules.py:103-288): the AE / VAE reconstructs from its own split of
    the state only, the heads see theirs; the encoder and the decoder run as two libsrlz calls with the column mask between."""

    def __init__(self, state_dim=2, action_dim=6, cuda=False, model_type="custom_cnn", losses=None, split_dimensions=None,
                 n_hidden_reward=16, inverse_model_type="linear"):
        assert len(split_dimensions) == len(losses), "Please specify as many split dimensions {} as losses {} !".format(
            len(split_dimensions), len(losses))
        n_dims = sum(split_dimensions.values()) + list(split_dimensions.values()).count(-1)   # modules.py:122-127
        assert n_dims == state_dim, "The sum of all splits' dimensions {} must be equal to the state dimension {}".format(
            sum(split_dimensions.values()), str(state_dim))
        if "triplet" in losses:
            raise ValueError("triplet not supported when splitting representation")
        if n_hidden_reward != 16:
            raise ValueError("n_hidden_reward other than the reference's default 16 is outside the B200 path")
        super().__init__(state_dim, action_dim, cuda, model_type, losses, inverse_model_type)
        self.split_dimensions = split_dimensions
        for key, m in split_masks(split_dimensions, state_dim).items():
            self.register_buffer("_mask_" + key.replace("-", "_"), m, persistent=False)   # not part of srl_model.pth

    def detachSplit(self, tensor, index):
        return _ColMask.apply(tensor, getattr(self, "_mask_" + index.replace("-", "_")))

    def forward(self, x):
        """modules.py:181-187: forwardAutoencoder (:249-258) / forwardVAE (:236-247)"""
        cn = self.model
        params = [p for _, _, p in cn.slots()]
        outs = _ModelCall.apply(self, x.contiguous(), None, None, False, *params)     # encoder-only pass
        if cn.is_vae:
            mu, logvar = self.detachSplit(outs[0], "vae"), self.detachSplit(outs[1], "vae")
            z = mu
            if self.training:   # models/models.py:155-165
                z = _Reparam.apply(mu, logvar, torch.empty_like(mu).normal_())
            return cn.decode(z), mu, logvar
        index = "autoencoder" if "autoencoder" in self.losses else "dae"
        encoded = outs[0]
        return encoded, cn.decode(self.detachSplit(encoded, index))

    def forward_masked(self, x, rects, eps=None):
        raise NotImplementedError("the split model takes pre-masked observations (learner.py:395-397)")

    def inverseModel(self, state, next_state):
        return super().inverseModel(self.detachSplit(state, "inverse"), self.detachSplit(next_state, "inverse"))

    def forwardModel(self, state, action):
        return super().forwardModel(self.detachSplit(state, "forward"), action)

    def rewardModel(self, state, next_state):
        return super().rewardModel(self.detachSplit(state, "reward"), self.detachSplit(next_state, "reward"))

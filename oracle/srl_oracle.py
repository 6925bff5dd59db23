"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing on the product path may import this file.

A CPU / plain-PyTorch fp32 restatement of the ONE hot path of araffin/srl-zoo that this
repository accelerates: the conv autoencoder / beta-VAE / denoising-AE train step.  The
reference's arithmetic lives in a third-party dependency (PyTorch: ATen/MKLDNN; pinned by the
reference at pytorch=0.4.1, environment.yml:58) -- this file restates the reference's *call
sites* with torch.nn.functional in a table-driven, functional style (no nn.Module graph), each
function citing the reference file:line it follows.

Parity pinning: the reference's own tests hold no golden vectors for this path (all of them are
return-code smoke tests, tests/common.py:17-18) => "parity unpinned" by the reference's tests.
The oracle is instead pinned against the LIVE reference modules imported from /root/reference in
the build container (oracle/validate_against_reference.py; generated fixtures are committed under
tests/golden/ by oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# preprocessing/preprocess.py:7-10
IMG = 224
N_CH = 3
FLAT = 64 * 6 * 6  # models/autoencoders.py:95

# (key prefix, kind, args) in the exact creation order of the reference => identical RNG use.
# heads first (models/modules.py:37-39), conv stacks (models/models.py:47-83), FCs last
# (models/autoencoders.py:94-100 / models/vae.py:52-58).
ENC_CONVS = [  # (state_dict index, cin, cout, k, stride, pad)     models/models.py:49,54,59
    (0, N_CH, 64, 7, 2, 3),
    (4, 64, 64, 3, 1, 1),
    (8, 64, 64, 3, 2, 1),
]
ENC_BNS = [1, 5, 9]  # models/models.py:50,55,60
ENC_POOLS = [(3, 2, 1), (3, 2, 0), (3, 2, 0)]  # (k, s, p)  models/models.py:52,57,62
DEC_CONVTS = [  # (index, cin, cout, k, stride)              models/models.py:66,70,74,78,82
    (0, 64, 64, 3, 2),
    (3, 64, 64, 3, 2),
    (6, 64, 64, 3, 2),
    (9, 64, 64, 3, 2),
    (12, 64, N_CH, 4, 2),
]
DEC_BNS = [1, 4, 7, 10]  # models/models.py:67,71,75,79
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# models/learner.py:204-207
DEFAULT_WEIGHTS = {"forward": 1.0, "inverse": 2.0, "autoencoder": 1.0, "vae": 0.5e-6, "dae": 1.0}


def build_state(kind, state_dim=200, action_dim=6, seed=1, inverse_model_type="linear"):
    """Fresh parameters + buffers, keyed exactly like SRLModules.state_dict().

    Creation order follows models/modules.py:37-49 -> models/models.py:47-83 ->
    models/autoencoders.py:94-100 (AE/DAE) or models/vae.py:52-58 (VAE), so that
    torch.manual_seed(seed) (models/learner.py:59) yields the reference's initial weights.
    kind in {"ae", "vae"} ("dae" uses the AE network, models/modules.py:43).
    """
    assert kind in ("ae", "vae")
    torch.manual_seed(seed)
    sd = OrderedDict()

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[prefix + "." + k] = v.detach().clone()

    put("forward_net", nn.Linear(state_dim + action_dim, state_dim))   # forward_inverse.py:16
    if inverse_model_type == "linear":
        put("inverse_net", nn.Linear(2 * state_dim, action_dim))       # forward_inverse.py:48
    else:                                                              # forward_inverse.py:50-56 ("mlp", n_hidden=128)
        for i, (a, b) in zip((0, 2, 4), ((2 * state_dim, 128), (128, 128), (128, action_dim))):
            put("inverse_net.%d" % i, nn.Linear(a, b))
    for i, (a, b) in zip((0, 2, 4), ((2 * state_dim, 16), (16, 16), (16, 2))):  # forward_inverse.py:79-83
        put("reward_net.%d" % i, nn.Linear(a, b))
    for (idx, cin, cout, k, s, p), bn in zip(ENC_CONVS, ENC_BNS):
        put("model.encoder_conv.%d" % idx, nn.Conv2d(cin, cout, k, s, p, bias=False))
        put("model.encoder_conv.%d" % bn, nn.BatchNorm2d(cout))
    for j, (idx, cin, cout, k, s) in enumerate(DEC_CONVTS):
        put("model.decoder_conv.%d" % idx, nn.ConvTranspose2d(cin, cout, k, s))
        if j < len(DEC_BNS):
            put("model.decoder_conv.%d" % DEC_BNS[j], nn.BatchNorm2d(cout))
    if kind == "ae":
        put("model.encoder_fc.0", nn.Linear(FLAT, state_dim))
        put("model.decoder_fc.0", nn.Linear(state_dim, FLAT))
    else:
        put("model.encoder_fc1", nn.Linear(FLAT, state_dim))
        put("model.encoder_fc2", nn.Linear(FLAT, state_dim))
        put("model.decoder_fc.0", nn.Linear(state_dim, FLAT))
    return sd


def is_buffer(key):
    return key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked")


def split_state(sd):
    """-> (params requiring grad, buffers).  Params are leaf tensors with requires_grad."""
    params, bufs = OrderedDict(), OrderedDict()
    for k, v in sd.items():
        if is_buffer(k):
            bufs[k] = v.clone()
        else:
            params[k] = v.clone().requires_grad_(True)
    return params, bufs


def _bn(x, P, B, prefix, training):
    # nn.BatchNorm2d forward (torch defaults eps=1e-5, momentum=0.1); updates running stats in place
    if training:
        B[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, B[prefix + ".running_mean"], B[prefix + ".running_var"],
                        P[prefix + ".weight"], P[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)


def encoder_conv(P, B, x, training):
    """models/models.py:47-63 : 3 x [conv -> BN -> ReLU -> MaxPool]"""
    for (idx, _, _, _, s, p), bn, (pk, ps, pp) in zip(ENC_CONVS, ENC_BNS, ENC_POOLS):
        x = F.conv2d(x, P["model.encoder_conv.%d.weight" % idx], None, s, p)
        x = _bn(x, P, B, "model.encoder_conv.%d" % bn, training)
        x = F.relu(x)
        x = F.max_pool2d(x, pk, ps, pp)
    return x


def decoder_conv(P, B, x, training):
    """models/models.py:65-83 : 4 x [convT3 s2 -> BN -> ReLU] -> convT4 s2"""
    for j, (idx, _, _, _, s) in enumerate(DEC_CONVTS):
        pre = "model.decoder_conv.%d" % idx
        x = F.conv_transpose2d(x, P[pre + ".weight"], P[pre + ".bias"], s)
        if j < len(DEC_BNS):
            x = _bn(x, P, B, "model.decoder_conv.%d" % DEC_BNS[j], training)
            x = F.relu(x)
    return x


def ae_encode(P, B, x, training):
    """models/autoencoders.py:102-109 (flatten is NCHW order)"""
    h = encoder_conv(P, B, x, training)
    h = h.reshape(h.size(0), -1)
    return F.linear(h, P["model.encoder_fc.0.weight"], P["model.encoder_fc.0.bias"])


def decode(P, B, z, training):
    """models/autoencoders.py:111-118 / models/vae.py:68-75"""
    h = F.linear(z, P["model.decoder_fc.0.weight"], P["model.decoder_fc.0.bias"])
    h = h.view(z.size(0), 64, 6, 6)
    return decoder_conv(P, B, h, training)


def ae_forward(P, B, x, training):
    """models/models.py:106-114 -> (encoded, decoded)"""
    enc = ae_encode(P, B, x, training)
    return enc, decode(P, B, enc, training).view(x.size())


def vae_encode(P, B, x, training):
    """models/vae.py:59-66 -> (mu, logvar)"""
    h = encoder_conv(P, B, x, training)
    h = h.reshape(h.size(0), -1)
    return (F.linear(h, P["model.encoder_fc1.weight"], P["model.encoder_fc1.bias"]),
            F.linear(h, P["model.encoder_fc2.weight"], P["model.encoder_fc2.bias"]))


def vae_forward(P, B, x, training, eps=None):
    """models/models.py:147-176 -> (decoded, mu, logvar).  eps: the N(0,1) draw of
    models/models.py:161 made explicit (if None it is drawn with the same call)."""
    mu, logvar = vae_encode(P, B, x, training)
    if training:
        std = logvar.mul(0.5).exp()
        if eps is None:
            eps = std.new(std.size()).normal_()
        z = eps.mul(std).add(mu)
    else:
        z = mu
    return decode(P, B, z, training).view(x.size()), mu, logvar


def get_states(kind, P, B, x, training):
    """models/models.py:85-90 (AE: encode) / :126-131 (VAE: encode(x)[0])"""
    if kind == "vae":
        return vae_encode(P, B, x, training)[0]
    return ae_encode(P, B, x, training)


def one_hot(actions, n):
    """models/models.py:229-237"""
    out = torch.zeros(actions.shape[0], n, device=actions.device)
    return out.scatter_(1, actions, 1.0)


def forward_model(P, s, actions, action_dim):
    """models/forward_inverse.py:21-31"""
    cat = torch.cat((s, one_hot(actions, action_dim).to(s.dtype)), dim=1)
    return s + F.linear(cat, P["forward_net.weight"], P["forward_net.bias"])


def _mlp3(P, prefix, x):
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) with state_dict keys prefix.{0,2,4}"""
    for i in (0, 2, 4):
        x = F.linear(x, P["%s.%d.weight" % (prefix, i)], P["%s.%d.bias" % (prefix, i)])
        if i < 4:
            x = F.relu(x)
    return x


def inverse_model(P, s, ns):
    """models/forward_inverse.py:62-70 with the 'linear' (:47-48) or the 'mlp' (:50-56) head, whichever the state holds"""
    x = torch.cat((s, ns), dim=1)
    if "inverse_net.weight" in P:
        return F.linear(x, P["inverse_net.weight"], P["inverse_net.bias"])
    return _mlp3(P, "inverse_net", x)


def reward_model(P, s, ns):
    """models/forward_inverse.py:78-95"""
    return _mlp3(P, "reward_net", torch.cat((s, ns), dim=1))


def detach_split(t, split_dimensions, index):
    """SRLModulesSplit.detachSplit (models/modules.py:189-234): walk the splits in order; a split owns the next n_dim columns
    (n_dim == -1: the same columns as the split before it); columns of `index` are kept, all others replaced by zeros."""
    pieces, start, prev = [], 0, 0
    for key, n_dim in split_dimensions.items():
        n_dim = int(n_dim)
        shared = n_dim == -1 and start > 0
        if shared:
            if key != index:
                continue                      # modules.py:207-211: nothing appended, the previous piece covers these columns
            pieces[-1] = t[:, start - prev:start]   # modules.py:220-222: re-attach the columns shared with the previous split
            continue
        pieces.append(t[:, start:start + n_dim] if key == index else torch.zeros_like(t[:, start:start + n_dim]))
        prev = n_dim
        start += n_dim
    return torch.cat(pieces, dim=1)


def reconstruction_loss(a, b):
    """losses/losses.py:172-181"""
    return torch.sum((a - b) ** 2) / a.nelement()


def apply_occlusion(x, rects):
    """preprocessing/data_loader.py:55-63 after the loader transpose (data_loader.py:255):
    the zeroed block of tensor (C, W, H) is [:, w1:w2, h1:h2]; rects[i] = (h1, h2, w1, w2)."""
    out = x.clone()
    for i in range(x.size(0)):
        h1, h2, w1, w2 = [int(v) for v in rects[i]]
        out[i, :, w1:w2, h1:h2] = 0.0
    return out


def preprocess_u8(image_u8, rect=None):
    """preprocessing/data_loader.py:49-65 + :255 and preprocessing/utils.py:20-32 from the point where the loader holds an
    RGB uint8 (H, W, 3) image (after cv2.resize / cvtColor): float32, /255, -mean, /std (in that order, fp32, in place), the
    optional DAE rectangle [h1:h2, w1:w2] zeroed in normalised space, then `reshape((1,)+shape).transpose(0, 3, 2, 1)`, i.e. a
    (1, 3, W, H) tensor.  The parity target of SURVEY.md 8f row N1 (uint8 hand-over fused into the first layer's load stage)."""
    x = np.asarray(image_u8).astype(np.float32)
    assert x.ndim == 3 and x.shape[-1] == 3
    x /= 255.
    for c, (m, sd) in enumerate(((0.485, 0.229), (0.456, 0.224), (0.406, 0.225))):
        x[..., c] -= m
    for c, (m, sd) in enumerate(((0.485, 0.229), (0.456, 0.224), (0.406, 0.225))):
        x[..., c] /= sd
    if rect is not None:
        h1, h2, w1, w2 = [int(v) for v in rect]
        x[h1:h2, w1:w2, :] = 0.
    return torch.tensor(x.reshape((1,) + x.shape).transpose(0, 3, 2, 1))


def sample_rects(n, occlusion_percentage=0.5, rng=None):
    """preprocessing/data_loader.py:23-35,56-59 run single-threaded -> (n,4) int32 (h1,h2,w1,w2)."""
    rng = rng or np.random
    out = np.zeros((n, 4), dtype=np.int32)
    for i in range(n):
        vals = []
        for _ in range(2):
            c1 = rng.randint(IMG)
            lo = max(0, c1 - IMG * occlusion_percentage)
            hi = min(c1 + IMG * occlusion_percentage, IMG)
            c2 = rng.randint(low=int(lo), high=int(hi))
            vals += [min(c1, c2), max(c1, c2)]
        out[i] = vals
    return out


class Adam:
    """th.optim.Adam(params, lr) restated (models/learner.py:199): beta=(0.9,0.999), eps=1e-8, wd=0;
    params with grad None are skipped (Appendix A.9)."""

    def __init__(self, params, lr=0.005):
        self.lr, self.b1, self.b2, self.eps = lr, 0.9, 0.999, 1e-8
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}

    @torch.no_grad()
    def step(self, params):
        for k, p in params.items():
            if p.grad is None:
                continue
            self.t[k] += 1
            t = self.t[k]
            self.m[k].mul_(self.b1).add_(p.grad, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(p.grad, p.grad, value=1 - self.b2)
            bc1, bc2 = 1 - self.b1 ** t, 1 - self.b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def train_step(kind, P, B, obs, next_obs, actions=None, eps=None, next_eps=None, rects=None,
               next_rects=None, use_forward=False, use_inverse=False, beta=1.0, weights=None,
               action_dim=6, training=True, optimizer=None, use_reward=False, rewards=None, split_dimensions=None):
    """One minibatch of SRL4robotics.learn (models/learner.py:373-497).

    kind: "ae" | "dae" | "vae".  use_reward / rewards: the reward head (learner.py:443-450, weight 1.0 at learner.py:204);
    split_dimensions (OrderedDict): SRLModulesSplit semantics (models/modules.py:103-288).  Returns dict(losses={name: unweighted scalar}, total, states,
    next_states, decoded, next_decoded[, mu, logvar, ...]).  Gradients are left in P[*].grad;
    optimizer.step() is applied when given and training (validation minibatches run eval mode,
    still call backward, never step: learner.py:362-366,487-497).
    """
    w = dict(DEFAULT_WEIGHTS)
    w["reward"] = 1.0   # learner.py:204
    if weights:
        w.update(weights)
    for p in P.values():  # optimizer.zero_grad()  learner.py:373
        p.grad = None
    out = {}
    terms = []  # (name, weight, value)   LossManager.addToLosses  losses.py:35-44
    ds = (lambda t, index: detach_split(t, split_dimensions, index)) if split_dimensions is not None else (lambda t, index: t)

    def ae_call(x):   # models/models.py:106-114 ; split: models/modules.py:249-258 (decode sees only the autoencoder's split)
        enc = ae_encode(P, B, x, training)
        return enc, decode(P, B, ds(enc, "dae" if kind == "dae" else "autoencoder"), training).view(x.size())

    def vae_call(x, e):   # models/models.py:147-176 ; split: models/modules.py:236-247
        mu_, lv_ = vae_encode(P, B, x, training)
        mu_, lv_ = ds(mu_, "vae"), ds(lv_, "vae")
        if training:
            std = lv_.mul(0.5).exp()
            z = (e if e is not None else std.new(std.size()).normal_()).mul(std).add(mu_)
        else:
            z = mu_
        return decode(P, B, z, training).view(x.size()), mu_, lv_

    if kind in ("ae", "dae"):
        x, nx = obs, next_obs
        if kind == "dae":  # learner.py:395-397: the model sees the noisy tensors
            x, nx = apply_occlusion(obs, rects), apply_occlusion(next_obs, next_rects)
        states, dec = ae_call(x)            # learner.py:393 (two separate calls)
        nstates, ndec = ae_call(nx)
    else:
        dec, mu, logvar = vae_call(obs, eps)            # learner.py:400
        ndec, nmu, nlogvar = vae_call(next_obs, next_eps)
        states = get_states("vae", P, B, obs, training)                    # learner.py:402 (extra passes)
        nstates = get_states("vae", P, B, next_obs, training)
        out.update(mu=mu, logvar=logvar, next_mu=nmu, next_logvar=nlogvar)
    if use_forward:  # learner.py:432-436, losses.py:102-114 ; split: models/modules.py:270-279
        pred = forward_model(P, ds(states, "forward"), actions, action_dim)
        terms.append(("forward_loss", w["forward"], reconstruction_loss(pred, nstates)))
    if use_inverse:  # learner.py:438-441, losses.py:117-129 ; split: models/modules.py:260-268
        logits = inverse_model(P, ds(states, "inverse"), ds(nstates, "inverse"))
        terms.append(("inverse_loss", w["inverse"], F.cross_entropy(logits, actions.squeeze(1))))
    if use_reward:   # learner.py:443-450, losses.py:158-170 ; split: models/modules.py:281-288
        rlogits = reward_model(P, ds(states, "reward"), ds(nstates, "reward"))
        terms.append(("reward_loss", w["reward"], F.cross_entropy(rlogits, rewards)))
    if kind in ("ae", "dae"):  # learner.py:452-455, losses.py:184-196 (target = clean obs)
        val = reconstruction_loss(obs, dec) + reconstruction_loss(next_obs, ndec)
        terms.append(("reconstruction_loss", w["dae" if kind == "dae" else "autoencoder"], val))
    else:  # learner.py:457-468, losses.py:239-256 and 199-214
        kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
        kl = kl + -0.5 * torch.sum(1 + nlogvar - nmu.pow(2) - nlogvar.exp())
        terms.append(("kl_loss", beta, kl))
        gen = F.mse_loss(dec, obs, reduction="sum") + F.mse_loss(ndec, next_obs, reduction="sum")
        terms.append(("generation_loss", w["vae"], gen))
    total = sum(wt * v for _, wt, v in terms)  # losses.py:55-56
    total.backward()                            # learner.py:489
    if optimizer is not None and training:
        optimizer.step(P)                       # learner.py:495
    out.update(losses={n: float(v.detach()) for n, _, v in terms},
               weights={n: wt for n, wt, _ in terms}, total=float(total.detach()),
               states=states.detach(), next_states=nstates.detach(),
               decoded=dec.detach(), next_decoded=ndec.detach())
    return out


def synthetic_batch(bs, seed=1234, n_actions=6):
    """uint8 U{0..255} -> /255, ImageNet mean/std (preprocessing/utils.py:20-32),
    laid out (B,3,224,224) like the loader delivers (data_loader.py:255)."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)

    def one():
        u8 = torch.randint(0, 256, (bs, 3, IMG, IMG), generator=g, dtype=torch.uint8)
        return ((u8.float() / 255.0) - mean) / std

    obs, nobs = one(), one()
    actions = torch.randint(0, n_actions, (bs, 1), generator=g, dtype=torch.int64)
    return obs, nobs, actions

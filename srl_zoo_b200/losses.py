"""
Host-side mirror of the reference's loss API for the hot path (losses/losses.py): the same free functions with
the same names, argument order, `loss_manager.addToLosses(name, weight, value)` protocol and loss names
('reconstruction_loss', 'generation_loss', 'kl_loss', 'forward_loss', 'inverse_loss'), backed by libsrlz kernels.
Values are 0-dim CUDA tensors that support .item() and take part in loss.backward() (models/learner.py:484-489).
"""
import torch

from ._lib import check, lib, ptr, stream_ptr
from . import ops


class LossManager:
    """losses/losses.py:19-59"""

    def __init__(self, model, loss_history=None):
        self.reg_params = [p for n, p in model.named_parameters() if ".bias" not in n and p.requires_grad]
        self.loss_history = loss_history
        self.names, self.weights, self.losses = [], [], []

    def addToLosses(self, name, weight, loss_value):
        self.names.append(name)
        self.weights.append(weight)
        self.losses.append(loss_value)

    def updateLossHistory(self):
        if self.loss_history is None:
            return
        for name, w, loss in zip(self.names, self.weights, self.losses):
            if w > 0:
                if len(self.loss_history[name]) > 0:
                    self.loss_history[name][-1] += w * loss.item()
                else:
                    self.loss_history[name].append(w * loss.item())

    def computeTotalLoss(self):
        return sum(self.weights[i] * self.losses[i] for i in range(len(self.losses)))

    def resetLosses(self):
        self.names, self.weights, self.losses = [], [], []


class _SSE(torch.autograd.Function):
    """sum((a-b)^2) with gradient 2*g*(a-b) to either side."""

    @staticmethod
    def forward(ctx, a, b):
        # the kernels walk both tensors with a.numel(): a broadcast or dtype mismatch the reference would accept (or raise on)
        # must never become an out-of-bounds read here; install() routes such calls to the reference implementation
        if tuple(a.shape) != tuple(b.shape) or a.dtype != torch.float32 or b.dtype != torch.float32 or not (a.is_cuda and b.is_cuda):
            raise RuntimeError("squared-error kernels need two CUDA float32 tensors of the same shape, got %s %s / %s %s"
                               % (tuple(a.shape), a.dtype, tuple(b.shape), b.dtype))
        ctx.save_for_backward(a, b)
        return ops.sse(a, b).reshape(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = ops.mse_grad(a, b, 1.0) * (2.0 * g) if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else None
        return (ga if ctx.needs_input_grad[0] else None), (-ga if ctx.needs_input_grad[1] else None)


class _FusedSSE(torch.autograd.Function):
    """sum((decoded - x)^2) already reduced inside the model call that produced `decoded` from `x` (modules._ModelCall): the value
    is picked up, and backward hands the model call a COEFFICIENT (its last layer recomputes coef*(decoded - x) on the fly)
    instead of materialising a (B,3,224,224) gradient; the zero-stride placeholder keeps autograd's bookkeeping intact."""

    @staticmethod
    def forward(ctx, decoded, rec):
        ctx.rec = rec
        ctx.save_for_backward(decoded)
        return rec["sse"][0].clone().reshape(())

    @staticmethod
    def backward(ctx, g):
        (decoded,) = ctx.saved_tensors
        rec = ctx.rec
        rec["coef"] = g if rec["coef"] is None else rec["coef"] + g
        rec["decoded"] = decoded
        return rec["dummy"], None


def _sse(a, b):
    """sum((a-b)^2): the fused value when (a, b) is (model input, decoded output of that model call), else the stand-alone kernels"""
    for x, dec in ((a, b), (b, a)):
        rec = getattr(dec, "_srlz_fused", None)
        if rec is not None and torch.is_tensor(x) and x.data_ptr() == rec["x_ptr"] and x._version == rec["x_version"] \
                and tuple(x.shape) == rec["shape"] and not x.requires_grad and dec.requires_grad:
            return _FusedSSE.apply(dec, rec)
    return _SSE.apply(a, b)


class _KL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        mu, logvar = mu.contiguous(), logvar.contiguous()
        ctx.save_for_backward(mu, logvar)
        out = torch.empty(1, dtype=torch.float32, device=mu.device)
        ws = torch.empty(2048, dtype=torch.float32, device=mu.device)
        check(lib.srlz_kl(ptr(mu), ptr(logvar), mu.numel(), ptr(out), ptr(ws), stream_ptr()), "kl")
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        mu, logvar = ctx.saved_tensors
        dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
        check(lib.srlz_kl_grad(ptr(mu), ptr(logvar), mu.numel(), 1.0, ptr(dmu), ptr(dlv), stream_ptr()), "kl_grad")
        return dmu * g, dlv * g


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, actions):
        logits = logits.contiguous()
        B, A = logits.shape
        if not (actions.is_cuda and actions.dtype == torch.int64 and tuple(actions.shape) == (B,)):
            raise RuntimeError("cross-entropy targets must be an int64 CUDA tensor of shape (%d,), got %s %s" % (B, tuple(actions.shape), actions.dtype))
        out = torch.empty(1, dtype=torch.float32, device=logits.device)
        gl = torch.empty_like(logits)
        ws = torch.empty(2048, dtype=torch.float32, device=logits.device)
        check(lib.srlz_cross_entropy(ptr(logits), ptr(actions.contiguous()), B, A, ptr(out), ptr(gl), ptr(ws), stream_ptr()), "ce")
        ctx.save_for_backward(gl)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        (gl,) = ctx.saved_tensors
        return gl * g, None


def reconstructionLoss(input_image, target_image):
    """losses/losses.py:172-181"""
    return _sse(input_image, target_image) / input_image.numel()


def autoEncoderLoss(obs, decoded_obs, next_obs, decoded_next_obs, weight, loss_manager):
    """losses/losses.py:184-196"""
    ae_loss = reconstructionLoss(obs, decoded_obs) + reconstructionLoss(next_obs, decoded_next_obs)
    loss_manager.addToLosses('reconstruction_loss', weight, ae_loss)
    return weight * ae_loss


def generationLoss(decoded, next_decoded, obs, next_obs, weight, loss_manager):
    """losses/losses.py:199-214"""
    generation_loss = _sse(decoded, obs) + _sse(next_decoded, next_obs)
    loss_manager.addToLosses('generation_loss', weight, generation_loss)
    return weight * generation_loss


def kullbackLeiblerLoss(mu, next_mu, logvar, next_logvar, loss_manager, beta=1):
    """losses/losses.py:239-256"""
    kl_divergence = _KL.apply(mu, logvar) + _KL.apply(next_mu, next_logvar)
    loss_manager.addToLosses('kl_loss', beta, kl_divergence)
    return beta * kl_divergence


def forwardModelLoss(next_states_pred, next_states, weight, loss_manager):
    """losses/losses.py:102-114"""
    forward_loss = reconstructionLoss(next_states_pred, next_states)
    loss_manager.addToLosses('forward_loss', weight, forward_loss)
    return weight * forward_loss


def inverseModelLoss(actions_pred, actions_st, weight, loss_manager):
    """losses/losses.py:117-129"""
    inverse_loss = _CrossEntropy.apply(actions_pred, actions_st.squeeze(1))
    loss_manager.addToLosses('inverse_loss', weight, inverse_loss)
    return weight * inverse_loss


def rewardModelLoss(rewards_pred, rewards_st, weight, loss_manager):
    """losses/losses.py:158-170 (categorical reward prediction: cross-entropy, mean over the batch)"""
    reward_loss = _CrossEntropy.apply(rewards_pred, rewards_st)
    loss_manager.addToLosses('reward_loss', weight, reward_loss)
    return weight * reward_loss


def supports(name, args):
    """True when the libsrlz kernels cover this call of loss function `name` (CUDA float32 tensors, equal shapes where two
    tensors are compared, int64 targets): install() routes everything else to the reference's own implementation."""
    ts = [a for a in args if torch.is_tensor(a)]
    if not ts or not all(t.is_cuda for t in ts):
        return False
    fl = [t for t in ts if t.is_floating_point()]
    if not all(t.dtype == torch.float32 for t in fl):
        return False
    if name == "autoEncoderLoss":       # (obs, decoded_obs, next_obs, decoded_next_obs, ...)
        return len(ts) >= 4 and ts[0].shape == ts[1].shape and ts[2].shape == ts[3].shape
    if name == "generationLoss":        # (decoded, next_decoded, obs, next_obs, ...)
        return len(ts) >= 4 and ts[0].shape == ts[2].shape and ts[1].shape == ts[3].shape
    if name == "forwardModelLoss":      # (next_states_pred, next_states, ...)
        return len(ts) >= 2 and ts[0].shape == ts[1].shape
    if name == "kullbackLeiblerLoss":   # (mu, next_mu, logvar, next_logvar, ...)
        return len(ts) >= 4 and ts[0].shape == ts[2].shape and ts[1].shape == ts[3].shape
    if name == "rewardModelLoss":       # (rewards_pred (B,2), rewards_st (B,), ...)
        return len(ts) >= 2 and ts[1].dtype == torch.int64 and ts[0].dim() == 2 and tuple(ts[1].shape) == (ts[0].shape[0],)
    if name == "inverseModelLoss":      # (actions_pred (B,A), actions_st (B,1), ...)
        return len(ts) >= 2 and ts[1].dtype == torch.int64 and ts[0].dim() == 2 and tuple(ts[1].shape) == (ts[0].shape[0], 1)
    return True

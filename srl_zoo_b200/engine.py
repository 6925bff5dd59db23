"""
Fused train step for the hot path: one minibatch body of SRL4robotics.learn (models/learner.py:373-497)
executed as a fixed sequence of libsrlz launches on one stream:

    pack weights -> forward(obs) -> forward(next_obs) [-> heads] -> backward(next_obs) -> backward(obs)
    -> ONE all-reduce over the flat gradient buffer (data parallel) -> fused Adam -> per-loss scalars

Differences from driving the drop-in module through autograd (modules.py) -- all result-preserving:
  * reconstruction / generation squared error is reduced inside the last decoder tile and its gradient is
    recomputed on the fly in backward (no (B,3,224,224) gradient tensor);
  * the VAE's two extra train-mode getStates() passes (learner.py:402) reuse mu (bit-equal in the reference) and
    only replay the BatchNorm running-stat updates, in the reference's order obs, next_obs, obs, next_obs;
  * parameters, gradients and Adam moments live in flat buffers; nn.Parameters are views (state_dict unchanged);
  * validation minibatches (eval mode) skip backward: the reference computes and discards those gradients
    (learner.py:487-492).
"""
import ctypes as C

import torch

from ._lib import SrlzNetGrads, check, lib, ptr, stream_ptr
from .modules import IMG, B200SRLModules
from . import parallel

N_PIX = 3 * IMG * IMG
LOSS_SLOTS = 8  # tail of the flat gradient buffer: per-loss scalars ride in the same all-reduce
# models/learner.py:204-207
DEFAULT_WEIGHTS = {"forward": 1.0, "inverse": 2.0, "autoencoder": 1.0, "vae": 0.5e-6, "dae": 1.0}


class TrainStep:
    def __init__(self, module, batch_size, lr=0.005, beta=1.0, losses_weights=None, world_size=1, process_group=None):
        if not isinstance(module, B200SRLModules):
            raise TypeError("TrainStep drives a B200SRLModules")
        if hasattr(module, "split_dimensions") or "reward" in module.losses:
            raise NotImplementedError("the fused step covers the BASELINE configs (AE / DAE / VAE + linear forward / inverse heads); "
                                      "split models and the reward head run through the drop-in module + loss functions (autograd)")
        self.module = module
        self.B = int(batch_size)                 # pairs per rank and per call
        self.world = int(world_size)
        self.pg = process_group
        self.global_B = self.B * self.world
        self.lr, self.beta = float(lr), float(beta)
        self.w = dict(DEFAULT_WEIGHTS)
        if losses_weights:
            self.w.update(losses_weights)
        losses = module.losses
        self.kind = "vae" if module.model.is_vae else ("dae" if "dae" in losses else "ae")
        self.use_forward = "forward" in losses
        self.use_inverse = "inverse" in losses
        if self.use_inverse and module.inverse_model_type != "linear":
            raise NotImplementedError("the fused step has the linear inverse head; the mlp head runs through the drop-in module (autograd)")
        self.step_count = 0
        self._flatten()
        self._alloc()

    # ---- flat parameter / gradient / moment buffers ----
    def _flatten(self):
        params = [p for p in self.module.parameters() if p.requires_grad]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs the module on a CUDA device (no CPU fallback)")
        self.device = dev
        n = sum(p.numel() for p in params)
        self.n_params = n
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n + LOSS_SLOTS, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self._gview = {}
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            self._gview[id(p)] = p.grad
            off += k
        cn = self.module.model
        self.grads = SrlzNetGrads()
        for name, idx, p in cn.slots():
            gp = self._gview[id(p)].data_ptr()
            if idx is None:
                setattr(self.grads, name, gp)
            else:
                getattr(self.grads, name)[idx] = gp
        self.loss_tail = self.flat_g[n:]

    def _alloc(self):
        cn, dev, B = self.module.model, self.device, self.B
        S, vae = cn.state_dim, int(cn.is_vae)
        u8 = lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev)
        f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        self.wpack = f32(lib.srlz_pack_floats(vae, S))
        self.ws = u8(lib.srlz_workspace_bytes(B, S, vae))
        self.saved = [u8(lib.srlz_saved_bytes(B, S, vae)) for _ in range(2)]
        self.lat = [f32(B, S) for _ in range(2)]
        self.logvar = [f32(B, S) if vae else None for _ in range(2)]
        self.decoded = [f32(B, 3, IMG, IMG) for _ in range(2)]
        self.loss_raw = [torch.zeros(2, dtype=torch.float32, device=dev) for _ in range(2)]
        self.heads_loss = torch.zeros(2, dtype=torch.float32, device=dev)
        self.gs = [f32(B, S) for _ in range(2)]
        self.heads_ws = u8(lib.srlz_heads_workspace_bytes(B, S, self.module.action_dim))

    # ---- one minibatch ----
    def step(self, obs, next_obs, actions=None, eps=None, next_eps=None, rects=None, next_rects=None, training=True,
             ready_events=None):
        """obs / next_obs: (B,3,224,224) float32 CUDA; actions (B,1) int64; eps: (B,S) draws for the VAE (drawn with
        torch's generator when None, models/models.py:161); rects: (B,4) int32 DAE rectangles.
        Returns a CUDA tensor of LOSS_SLOTS floats (see loss_names()) holding the UNWEIGHTED per-loss values of the
        global batch (after the all-reduce)."""
        mod, cn, B = self.module, self.module.model, self.B
        S = cn.state_dim
        st = stream_ptr()
        xs = (obs, next_obs)
        for x in xs:
            if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, 3, IMG, IMG)):
                raise RuntimeError("observations must be contiguous float32 CUDA tensors of shape (%d,3,%d,%d)" % (B, IMG, IMG))
        if self.kind == "dae" and (rects is None or next_rects is None):
            raise RuntimeError("dae step needs rects / next_rects (int32 (B,4): h1,h2,w1,w2)")
        rc = (rects, next_rects) if self.kind == "dae" else (None, None)
        for r in rc:
            if r is not None and not (r.is_cuda and r.dtype == torch.int32 and r.is_contiguous() and tuple(r.shape) == (B, 4)):
                raise RuntimeError("rects must be contiguous int32 CUDA tensors of shape (%d,4): (h1,h2,w1,w2) per image" % B)
        ep = [None, None]
        if self.kind == "vae" and training:
            ep = [eps if eps is not None else torch.empty(B, S, dtype=torch.float32, device=self.device).normal_(),
                  next_eps if next_eps is not None else torch.empty(B, S, dtype=torch.float32, device=self.device).normal_()]
        mod.train(training)
        net = cn.net_struct()
        check(lib.srlz_pack_weights(C.byref(net), ptr(self.wpack), st), "pack_weights")
        for i in range(2):
            if ready_events is not None:   # step_host: xs[i] is being filled by the copy stream
                torch.cuda.current_stream().wait_event(ready_events[i])
            check(lib.srlz_forward(C.byref(net), ptr(self.wpack), ptr(xs[i]), ptr(rc[i]), ptr(ep[i]), B, int(training),
                                   ptr(self.lat[i]), ptr(self.logvar[i]), ptr(self.decoded[i]), ptr(xs[i]),
                                   ptr(self.loss_raw[i]), ptr(self.saved[i]), ptr(self.ws), st), "forward")
        if self.kind == "vae" and training:  # learner.py:402: states = getStates(obs), getStates(next_obs)
            for i in range(2):
                check(lib.srlz_replay_running_stats(C.byref(net), B, ptr(self.saved[i]), st), "replay_running_stats")
        heads = self.use_forward or self.use_inverse
        if heads:
            if actions is None:
                raise RuntimeError("forward / inverse losses need actions")
            # the kernels index the head weights with the raw action value: torch's scatter_ / CrossEntropyLoss would raise
            # on a bad tensor, so this does too (the value range is checked on the device inside srlz_heads: out-of-range
            # actions poison the loss with NaN, which the learner turns into exit code 11, models/learner.py:520-522)
            if not (actions.is_cuda and actions.dtype == torch.int64 and actions.is_contiguous() and tuple(actions.shape) == (B, 1)):
                raise RuntimeError("actions must be a contiguous int64 CUDA tensor of shape (%d,1)" % B)
            wf = self.w["forward"] if self.use_forward else 0.0
            wi = self.w["inverse"] if self.use_inverse else 0.0
            fw = mod.forward_net
            iw = mod.inverse_net if self.use_inverse else None   # (an mlp inverse head is a Sequential: only touched when in use)
            check(lib.srlz_heads(ptr(self.lat[0]), ptr(self.lat[1]), ptr(actions), B, self.global_B, S, mod.action_dim,
                                 ptr(fw.weight), ptr(fw.bias), ptr(iw.weight) if iw is not None else None,
                                 ptr(iw.bias) if iw is not None else None, wf, wi, ptr(self.heads_loss),
                                 ptr(self.gs[0]), ptr(self.gs[1]), ptr(fw.weight.grad), ptr(fw.bias.grad),
                                 ptr(iw.weight.grad) if iw is not None else None, ptr(iw.bias.grad) if iw is not None else None,
                                 0, ptr(self.heads_ws), st), "heads")
        if training:
            wkey = "vae" if self.kind == "vae" else ("dae" if self.kind == "dae" else "autoencoder")
            mse_coef = parallel.mse_coef(self.kind, self.w[wkey], self.global_B)
            kl_coef = self.beta if self.kind == "vae" else 0.0
            for j, i in enumerate((1, 0)):
                check(lib.srlz_backward(C.byref(net), ptr(self.wpack), C.byref(self.grads), int(j > 0), ptr(xs[i]), ptr(rc[i]),
                                        ptr(ep[i]), B, 1, 1, None, ptr(self.decoded[i]), ptr(xs[i]), mse_coef,
                                        ptr(self.gs[i]) if heads else None, None, kl_coef, ptr(self.saved[i]), ptr(self.ws),
                                        st), "backward")
        # per-loss scalars (unweighted, this rank's share of the global batch) -> tail of the flat gradient buffer
        t = self.loss_tail
        t.zero_()
        sse = self.loss_raw[0][0] + self.loss_raw[1][0]
        if self.kind == "vae":
            t[0] = sse                                                   # generation_loss (sum)   losses.py:210-211
            t[1] = -0.5 * (self.loss_raw[0][1] + self.loss_raw[1][1])    # kl_loss (sum)           losses.py:253-254
        else:
            t[0] = sse * parallel.recon_scale(self.kind, self.global_B)  # reconstruction_loss     losses.py:181,194
        if heads:
            t[2:4] = self.heads_loss
        parallel.allreduce_flat(self.flat_g if training else t, self.world, self.pg)
        if training:
            self.step_count += 1
            check(lib.srlz_adam_step(ptr(self.flat_p), ptr(self.flat_g), ptr(self.m), ptr(self.v), self.n_params, self.lr,
                                     0.9, 0.999, 1e-8, self.step_count, st), "adam")
            mod._weights_version = getattr(mod, "_weights_version", 0) + 1   # the kernel wrote the parameters behind torch's back
            mod.model._weights_version = mod._weights_version
        return t

    def preprocess(self, frames, out=None):
        """uint8 RGB frames (B,224,224,3), the loader's native HWC order -> the normalised (B,3,224,224) float32 tensor the
        reference's loader delivers (preprocessing/utils.py:20-32, data_loader.py:255), bit-exact, on the device."""
        from . import ops
        return ops.preprocess_u8(frames, out)

    def _issue_h2d(self, slot, obs_host, next_obs_host, actions_host):
        """Host -> device copies of one minibatch into staging set `slot`, on the copy stream; one event per tensor.
        uint8 frames land in the uint8 staging pair (4x fewer bytes over PCIe / C2C) and are normalised on the device."""
        st = self._stage[slot]
        u8 = obs_host.dtype == torch.uint8
        with torch.cuda.stream(self._copy_stream):
            st["obs_u8" if u8 else "obs"].copy_(obs_host, non_blocking=True)
            st["ev"][0].record(self._copy_stream)
            st["nobs_u8" if u8 else "nobs"].copy_(next_obs_host, non_blocking=True)
            if actions_host is not None:
                st["act"].copy_(actions_host, non_blocking=True)
            st["ev"][1].record(self._copy_stream)
        st["key"] = (obs_host.data_ptr(), next_obs_host.data_ptr(), None if actions_host is None else actions_host.data_ptr())

    def step_host(self, obs_host, next_obs_host, actions_host=None, prefetch=None, **kw):
        """Reference-facing entry with HOST buffers (models/learner.py:368-371 does the same .to(device) per minibatch):
        pinned host tensors are copied to resident device staging buffers, the fused step runs, and the per-loss scalars
        come back to the host.  Returns a CPU tensor of LOSS_SLOTS floats.  Two input formats: (B,3,224,224) float32 (what
        the reference's loader delivers) or (B,224,224,3) uint8 RGB frames (what its loader holds before normalising,
        preprocessing/data_loader.py:38-47: a quarter of the bytes; /255, mean / std and the (C,W,H) transpose then run on
        the device, bit-exactly: SURVEY.md 8f N1).

        The copies run on their own stream: forward(obs) starts as soon as obs has landed while next_obs is still in
        flight.  `prefetch=(obs_host, next_obs_host[, actions_host])` names the NEXT minibatch (what the reference's loader
        queue already holds, preprocessing/data_loader.py:129-193): its copies are issued into the second staging set and
        overlap this step's kernels; the next call finds them there (matched by host address)."""
        if not hasattr(self, "_stage"):
            f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
            u8 = lambda: torch.empty(self.B, IMG, IMG, 3, dtype=torch.uint8, device=self.device)
            self._stage = [dict(obs=f32(self.B, 3, IMG, IMG), nobs=f32(self.B, 3, IMG, IMG), obs_u8=u8(), nobs_u8=u8(),
                                act=torch.empty(self.B, 1, dtype=torch.int64, device=self.device),
                                ev=[torch.cuda.Event(), torch.cuda.Event()], key=None) for _ in range(2)]
            self._stage_cur = 0
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._loss_host = torch.empty(LOSS_SLOTS, dtype=torch.float32).pin_memory()
        key = (obs_host.data_ptr(), next_obs_host.data_ptr(), None if actions_host is None else actions_host.data_ptr())
        cur = self._stage_cur
        if self._stage[cur]["key"] != key:   # not prefetched by the previous call: copy now
            self._copy_stream.wait_stream(torch.cuda.current_stream())
            self._issue_h2d(cur, obs_host, next_obs_host, actions_host)
        st = self._stage[cur]
        ready = st["ev"]
        if obs_host.dtype == torch.uint8:
            for name, ev in zip(("obs", "nobs"), st["ev"]):
                torch.cuda.current_stream().wait_event(ev)
                self.preprocess(st[name + "_u8"], st[name])
            ready = None
        t = self.step(st["obs"], st["nobs"], st["act"] if actions_host is not None else None, ready_events=ready, **kw)
        st["key"] = None
        if prefetch is not None:
            # the other staging set was last read by the previous call, which ended with a stream synchronize
            self._issue_h2d(1 - cur, prefetch[0], prefetch[1], prefetch[2] if len(prefetch) > 2 else None)
            self._stage_cur = 1 - cur
        self._loss_host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._loss_host

    # ---- resume: Adam state in torch.optim.Adam's own format ----
    def optimizer_state_dict(self):
        """The engine's Adam state as `torch.optim.Adam(learnable_params, lr).state_dict()` would hold it
        (models/learner.py:194-199): save it next to `srl_model.pth`; either side can resume from it."""
        from .checkpoint import adam_state_to_torch
        params = [p for p in self.module.parameters() if p.requires_grad]
        sd = adam_state_to_torch(params, self.m, self.v, self.step_count, lr=self.lr)
        for st in sd["state"].values():
            st["exp_avg"], st["exp_avg_sq"] = st["exp_avg"].cpu(), st["exp_avg_sq"].cpu()
        return sd

    def load_optimizer_state_dict(self, sd):
        from .checkpoint import adam_state_from_torch
        params = [p for p in self.module.parameters() if p.requires_grad]
        self.step_count = adam_state_from_torch(sd, params, self.m, self.v)
        self.lr = float(sd["param_groups"][0].get("lr", self.lr))

    def h2d_bytes_per_step(self, with_actions=False, uint8=True):
        return 2 * self.B * N_PIX * (1 if uint8 else 4) + (self.B * 8 if with_actions else 0)

    def d2h_bytes_per_step(self):
        return LOSS_SLOTS * 4

    def loss_names(self):
        names = ["generation_loss", "kl_loss"] if self.kind == "vae" else ["reconstruction_loss", None]
        names += ["forward_loss" if self.use_forward else None, "inverse_loss" if self.use_inverse else None]
        return names

    def loss_weights(self):
        if self.kind == "vae":
            w = [self.w["vae"], self.beta]
        else:
            w = [self.w["dae" if self.kind == "dae" else "autoencoder"], 0.0]
        return w + [self.w["forward"] if self.use_forward else 0.0, self.w["inverse"] if self.use_inverse else 0.0]

    def total_loss(self, t):
        """sum_i weight_i * loss_i  (LossManager.computeTotalLoss, losses/losses.py:55-56)"""
        w = torch.tensor(self.loss_weights(), dtype=torch.float32, device=t.device)
        return (t[:4] * w).sum()

    @torch.no_grad()
    def predict_states(self, obs):
        """eval-mode getStates (BaseLearner._predFn, models/learner.py:67-75) on the folded inference path; `obs` is the
        normalised (B,3,224,224) float32 tensor or uint8 RGB frames (B,224,224,3) (normalised on the device)"""
        if obs.dtype == torch.uint8:
            obs = self.preprocess(obs)
        was = self.module.training
        self.module.eval()
        try:
            return self.module.getStates(obs)
        finally:
            self.module.train(was)

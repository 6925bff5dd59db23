"""Shared helpers for the GPU parity tests: the CUDA path (srl_zoo_b200) against the oracle (oracle/srl_oracle.py)."""
import os

import numpy as np
import torch

from oracle import srl_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
S, A = 200, 6


def rel_err(a, b):
    """max |a-b| / max|b|"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def norm_rel(a, b):
    """max over batch rows of ||a-b|| / ||b||   (the headline states metric)"""
    a, b = a.detach().double().cpu().reshape(a.shape[0], -1), b.detach().double().cpu().reshape(b.shape[0], -1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-30)).max().item()


def cosine(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def make_pair(kind, losses, seed=1, device="cuda", inverse_model_type="linear", split_dimensions=None):
    """(B200 module on device, oracle params P, oracle buffers B) with identical weights; the oracle's tensors live on
    `device` too, so its torch.nn.functional calls run on the GPU (cuDNN / cuBLAS, TF32 off) and the tests stay lean."""
    import srl_zoo_b200
    torch.manual_seed(seed)
    if split_dimensions is None:
        mod = srl_zoo_b200.B200SRLModules(S, A, True, "custom_cnn", losses, inverse_model_type).to(device)
    else:
        mod = srl_zoo_b200.B200SRLModulesSplit(S, A, True, "custom_cnn", losses, split_dimensions, 16, inverse_model_type).to(device)
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=seed, inverse_model_type=inverse_model_type)
    msd = mod.state_dict()
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(msd[k].cpu(), sd[k]), k  # same RNG order as the reference (models/modules.py:37-49)
    P, B = O.split_state({k: v.to(device) for k, v in sd.items()})
    return mod, P, B


def inputs(bs, seed=1234, device="cuda"):
    obs, nobs, actions = O.synthetic_batch(bs, seed=seed)
    g = torch.Generator().manual_seed(7)
    eps = (torch.randn(bs, S, generator=g), torch.randn(bs, S, generator=g))
    rng = np.random.RandomState(1)
    rects = (O.sample_rects(bs, rng=rng), O.sample_rects(bs, rng=rng))
    cpu = dict(obs=obs, nobs=nobs, actions=actions, eps=eps, rects=rects)
    dev = dict(obs=obs.to(device), nobs=nobs.to(device), actions=actions.to(device),
               eps=tuple(e.to(device) for e in eps),
               rects=tuple(torch.from_numpy(r).to(device) for r in rects))
    return cpu, dev


def tf32_off():
    """stock torch lets cuDNN / cuBLAS use TF32 (Appendix A.12): the GPU-side oracle must be true fp32"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def oracle_state(kind, dtype=torch.float32, device="cuda", seed=1):
    """oracle parameters / buffers with the reference's initial weights (same seed as make_pair) on `device` in `dtype`:
    the oracle's functions are plain torch.nn.functional calls, so they run unchanged on the GPU (cuDNN, TF32 off)"""
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=seed)
    sd = {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device)) for k, v in sd.items()}
    return O.split_state(sd)


def oracle_step(kind, losses, P, B, x, dtype=torch.float32, optimizer=None, training=True):
    """O.train_step on the inputs dict `x` (cpu or device tensors), cast to `dtype`"""
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t
    rects = [r.cpu().numpy() if torch.is_tensor(r) else r for r in x["rects"]]
    return O.train_step(kind, P, B, cast(x["obs"]), cast(x["nobs"]), x["actions"], cast(x["eps"][0]), cast(x["eps"][1]),
                        rects[0], rects[1], use_forward="forward" in losses, use_inverse="inverse" in losses,
                        optimizer=optimizer, training=training)

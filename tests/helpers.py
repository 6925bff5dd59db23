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


def make_pair(kind, losses, seed=1, device="cuda"):
    """(B200 module on device, oracle params P, oracle buffers B) with identical weights."""
    import srl_zoo_b200
    torch.manual_seed(seed)
    mod = srl_zoo_b200.B200SRLModules(S, A, True, "custom_cnn", losses).to(device)
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=seed)
    msd = mod.state_dict()
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(msd[k].cpu(), sd[k]), k  # same RNG order as the reference (models/modules.py:37-49)
    P, B = O.split_state(sd)
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

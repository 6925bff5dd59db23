"""Eval-mode BatchNorm folding for the encoder-only inference path.

`predStatesWithDataLoader` / `getStates` (models/learner.py:77-88,570-577) run the encoder in eval mode: every
`Conv2d -> BatchNorm2d` pair (models/models.py:49-50,54-55,59-60) is then one affine map per output channel,

    bn(conv(x, W)) = conv(x, W * s[:, None, None, None]) + t,    s = gamma / sqrt(running_var + eps),  t = beta - running_mean * s

so conv + ReLU + MaxPool fuse into one kernel with no batch statistics (no HBM round trip between them).  This module is
the host-side weight transformation of that path (pure tensor arithmetic, unit-tested on CPU against the oracle)."""
import torch

BN_EPS = 1e-5
ENCODER_PAIRS = ((0, 1), (4, 5), (8, 9))   # (conv index, BatchNorm index) inside model.encoder_conv


def fold_encoder_bn(state_dict, prefix="model.encoder_conv."):
    """-> [(weight (64,Cin,k,k), bias (64,))] for the three encoder stages, BatchNorm (running statistics) folded in."""
    out = []
    for conv, bn in ENCODER_PAIRS:
        w = state_dict["%s%d.weight" % (prefix, conv)].detach()
        gamma, beta = state_dict["%s%d.weight" % (prefix, bn)].detach(), state_dict["%s%d.bias" % (prefix, bn)].detach()
        mean, var = state_dict["%s%d.running_mean" % (prefix, bn)].detach(), state_dict["%s%d.running_var" % (prefix, bn)].detach()
        s = (gamma.double() / torch.sqrt(var.double() + BN_EPS))
        wf = (w.double() * s[:, None, None, None]).to(w.dtype)
        bf = (beta.double() - mean.double() * s).to(w.dtype)
        out.append((wf, bf))
    return out

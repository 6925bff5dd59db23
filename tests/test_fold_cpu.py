"""Eval-mode BatchNorm folding (srl_zoo_b200/fold.py) against the oracle's eval-mode encoder."""
import torch
import torch.nn.functional as F

from oracle import srl_oracle as O
from srl_zoo_b200.fold import fold_encoder_bn


def test_folded_encoder_equals_eval_mode_oracle():
    sd = O.build_state("ae", 200, 6, seed=1)
    g = torch.Generator().manual_seed(2)
    for bn in (1, 5, 9):   # non-trivial running statistics and affine parameters
        sd["model.encoder_conv.%d.running_mean" % bn] = 0.3 * torch.randn(64, generator=g)
        sd["model.encoder_conv.%d.running_var" % bn] = 0.5 + torch.rand(64, generator=g)
        sd["model.encoder_conv.%d.weight" % bn] = 0.5 + torch.rand(64, generator=g)
        sd["model.encoder_conv.%d.bias" % bn] = 0.2 * torch.randn(64, generator=g)
    P, B = O.split_state(sd)
    x, _, _ = O.synthetic_batch(2, seed=5)
    with torch.no_grad():
        ref = O.encoder_conv(P, B, x, training=False)
        h = x
        for (w, b), (_, _, _, _, s, p), (pk, ps, pp) in zip(fold_encoder_bn(sd), O.ENC_CONVS, O.ENC_POOLS):
            h = F.max_pool2d(F.relu(F.conv2d(h, w, b, s, p)), pk, ps, pp)
        states_ref = O.get_states("ae", P, B, x, training=False)
        states = F.linear(h.reshape(2, -1), P["model.encoder_fc.0.weight"], P["model.encoder_fc.0.bias"])
    assert h.shape == ref.shape == (2, 64, 6, 6)
    assert (h - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert ((states - states_ref).norm(dim=1) / states_ref.norm(dim=1)).max().item() < 1e-4

"""GPU parity at the BASELINE.json per-GPU batch sizes (cfg2 AE 256, cfg3 VAE 1024/8 = 128, cfg4 AE+forward+inverse
512/4 = 128, cfg5 DAE 2048/8 = 256): one fused train step through the C ABI against the oracle run ON THE GPU (the oracle's
functions are torch.nn.functional calls: cuDNN / cuBLAS with TF32 off), and against an fp64 run of the same oracle as the
yardstick for the gradients.  This is the regime where the wgrad kernels differ most from the toy sizes: CTAs walk long row
ranges with TMEM accumulators live for the whole range, and the cross-CTA partial sums are folded in a fixed order.

Gates (stated per tensor class, against fp64):
  states / mu / logvar   <= 1e-4  max over batch of ||d|| / ||ref||       (north-star tolerance)
  decoded                <= 1e-4  of max|ref|
  per-loss scalars       <= 1e-5  relative
  BatchNorm buffers      <= 1e-5  of max|ref|
  gradients              cosine >= 0.9999 for every tensor, and max|d| / max|ref| <= GATE[group] against fp64, one
                         gate per position in the backward chain (the error of the bf16x3 tensor-core products, 2^-17 per
                         operand, compounds layer by layer and is amplified by every train-mode BatchNorm backward: measured
                         on B200 at these sizes it grows from 6e-5 at the last decoder layer to 1e-2 at the first encoder
                         layer, 4-10x the fp32 cuDNN oracle's own distance from fp64 at the same tensor; each gate is the
                         largest value measured over the four configs with 2-2.5x headroom):
                           decoder_conv.12 / .10   2e-4      decoder_conv.9 / .7   5e-4     decoder_conv.6 / .4   1.5e-3
                           decoder_conv.3 / .1     2.5e-3    decoder_conv.0        2e-2     decoder_fc, encoder_fc*  1e-2
                           encoder_conv.8 / .9     1.5e-2    encoder_conv.{0,1,4,5}  2.5e-2  forward_net, inverse_net  1e-4
  pre-BatchNorm biases   exact gradient 0: |g| <= 1e-5 * max|g of the matching weight|
"""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import srl_oracle as O

pytestmark = pytest.mark.gpu

CFGS = {
    "cfg2_ae_256": ("ae", ["autoencoder"], 256),
    "cfg3_vae_128": ("vae", ["vae"], 128),
    "cfg4_ae_fwd_inv_128": ("ae", ["autoencoder", "forward", "inverse"], 128),
    "cfg5_dae_256": ("dae", ["dae"], 256),
}
NOISE_BIAS = {"model.decoder_conv.%d.bias" % i: "model.decoder_conv.%d.weight" % i for i in (0, 3, 6, 9)}
GATE = {"dec12": 2e-4, "dec9": 5e-4, "dec6": 1.5e-3, "dec3": 2.5e-3, "dec0": 2e-2, "fc": 1e-2, "enc8": 1.5e-2, "enc04": 2.5e-2,
        "heads": 1e-4}


def tensor_class(k):
    """position of a parameter in the backward chain: a decoder BatchNorm is grouped with the transposed conv that consumes its
    output (their gradients are taken from the same dy), an encoder BatchNorm with the conv that feeds it"""
    if k.startswith(("forward_net", "inverse_net")):
        return "heads"
    if "_fc" in k:
        return "fc"
    idx = int(k.split(".")[2])
    if k.startswith("model.encoder_conv"):
        return "enc8" if idx >= 8 else "enc04"
    return {12: "dec12", 10: "dec12", 9: "dec9", 7: "dec9", 6: "dec6", 4: "dec6", 3: "dec3", 1: "dec3", 0: "dec0"}[idx]


@pytest.mark.parametrize("name", list(CFGS))
def test_train_step_at_baseline_batch_size(name):
    import srl_zoo_b200
    H.tf32_off()
    kind, losses, bs = CFGS[name]
    mod, _, _ = H.make_pair(kind, losses)
    cpu, dev = H.inputs(bs)
    del cpu
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.005)
    t = eng.step(dev["obs"], dev["nobs"], dev["actions"], dev["eps"][0], dev["eps"][1], dev["rects"][0], dev["rects"][1])
    torch.cuda.synchronize()
    got_losses = {n: t[i].item() for i, n in enumerate(eng.loss_names()) if n}
    grads = {n: p.grad.detach().clone() for n, p in mod.named_parameters()}
    lat = [x.clone() for x in eng.lat]
    logvar = [x.clone() if x is not None else None for x in eng.logvar]
    dec_stats = []
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}

    # fp32 oracle on the GPU: forward values, losses, BN buffers
    P, B = H.oracle_state(kind)
    r = H.oracle_step(kind, losses, P, B, dev)
    for n, v in got_losses.items():
        assert abs(v - r["losses"][n]) <= 1e-5 * abs(r["losses"][n]), (n, v, r["losses"][n])
    if kind == "vae":
        assert H.norm_rel(lat[0], r["mu"]) < 1e-4 and H.norm_rel(lat[1], r["next_mu"]) < 1e-4
        assert H.norm_rel(logvar[0], r["logvar"]) < 1e-4 and H.norm_rel(logvar[1], r["next_logvar"]) < 1e-4
    else:
        assert H.norm_rel(lat[0], r["states"]) < 1e-4 and H.norm_rel(lat[1], r["next_states"]) < 1e-4
    assert H.rel_err(eng.decoded[0], r["decoded"]) < 1e-4
    assert H.rel_err(eng.decoded[1], r["next_decoded"]) < 1e-4
    for k in B:
        # after the step the engine has applied Adam to the parameters, the buffers are the forward's
        assert H.rel_err(sd[k].float(), B[k].float()) < 1e-5, k
    g32 = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in P.items()}
    del r, P, B, eng
    torch.cuda.empty_cache()

    # fp64 oracle on the GPU: the yardstick for every gradient
    P64, B64 = H.oracle_state(kind, torch.float64)
    H.oracle_step(kind, losses, P64, B64, dev, torch.float64)
    report = []
    for k, p in P64.items():
        if p.grad is None:   # unused heads (Appendix A.9): exactly zero in the flat buffer
            assert grads[k].abs().max().item() == 0.0, k
            continue
        g64 = p.grad
        if k in NOISE_BIAS:  # bias followed by train-mode BatchNorm: exact gradient 0
            wmax = P64[NOISE_BIAS[k]].grad.abs().max().item()
            assert grads[k].abs().max().item() <= 1e-5 * wmax, (k, grads[k].abs().max().item(), wmax)
            continue
        err, noise, cos = H.rel_err(grads[k], g64), H.rel_err(g32[k], g64), H.cosine(grads[k], g64)
        cls = tensor_class(k)
        report.append((k, cls, err, noise, cos))
    for k, cls, err, noise, cos in report:
        print("%-34s %-10s err %.2e  oracle-fp32 %.2e  cos-1 %.1e" % (k, cls, err, noise, cos - 1.0))
    for k, cls, err, noise, cos in report:
        assert cos > 0.9999, (k, cos)
        assert err <= GATE[cls], (k, cls, err, noise)

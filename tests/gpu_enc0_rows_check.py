"""Development check (not a pytest): row-image tcgen05 kernels of the first encoder layer (enc0_rows_tc.cu) against the
im2col tcgen05 kernels they replace, inside full train steps (lr = 0 so both runs see the same weights), then timings.
usage: python tests/gpu_enc0_rows_check.py [B_timing]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import srl_zoo_b200
from srl_zoo_b200 import _lib
from srl_zoo_b200._lib import lib
from srl_zoo_b200.occlusion import sample_rects


def run(kind, bs, mode, rects=None):
    torch.manual_seed(1)
    mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", [kind]).cuda()
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.0)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(bs, 3, 224, 224, generator=g).cuda()
    nobs = torch.randn(bs, 3, 224, 224, generator=g).cuda()
    kw = {}
    if rects is not None:
        kw = dict(rects=rects[0], next_rects=rects[1])
    lib.srlz_set_tensor_cores(mode)
    eng.step(obs, nobs, **kw)
    torch.cuda.synchronize()
    lib.srlz_set_tensor_cores(1)
    n = bs * 112 * 112 * 64
    y1 = [eng.saved[i][:n * 4].view(torch.float32).clone() for i in range(2)]
    gw = dict(mod.named_parameters())["model.encoder_conv.0.weight"].grad.clone()
    return y1, gw, eng.lat[0].clone()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


ok = True
for kind in ("autoencoder", "dae"):
    for bs in (1, 3) if kind == "dae" else (1, 3, 48):
        rects = None
        if kind == "dae":
            rng = np.random.RandomState(3)
            rects = tuple(torch.from_numpy(sample_rects(bs, rng=rng)).cuda() for _ in range(2))
        y_old, _, s_old = run(kind, bs, 2, rects)
        y_new, g_new, s_new = run(kind, bs, 1, rects)
        _, g_old, _ = run(kind, bs, 3, rects)   # same (row-image) forward, im2col wgrad: identical dy on both sides
        e = [rel(y_new[i], y_old[i]) for i in range(2)]
        eg, es = rel(g_new, g_old), rel(s_new, s_old)
        good = max(e) < 2e-5 and eg < 2e-4 and es < 1e-4
        ok &= good
        print("%-11s B=%d  y1 rel %.2e %.2e  enc0 wgrad rel %.2e  states rel %.2e  %s" % (kind, bs, e[0], e[1], eg, es, "OK" if good else "FAIL"))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for mode, name in ((2, "im2col kernels"), (1, "row-image kernels")):
    torch.manual_seed(1)
    mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
    eng = srl_zoo_b200.TrainStep(mod, B, lr=0.0)
    obs = torch.randn(B, 3, 224, 224, device="cuda")
    nobs = torch.randn(B, 3, 224, 224, device="cuda")
    lib.srlz_set_tensor_cores(mode)
    for _ in range(2):
        eng.step(obs, nobs)
    _lib.prof_enable(True)
    for _ in range(3):
        eng.step(obs, nobs)
    prof = _lib.prof_report()
    _lib.prof_enable(False)
    lib.srlz_set_tensor_cores(1)
    print("%-18s B=%d  enc0.fwd %.3f ms/call  enc0.wgrad %.3f ms/call" % (name, B, prof["enc0.fwd"][1] / prof["enc0.fwd"][0], prof["enc0.wgrad"][1] / prof["enc0.wgrad"][0]))
    del eng, mod
# clock64 timelines of CTA 0 (SRLZ_DBG_SITE=0: forward kernel, 4: wgrad kernel)
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
eng = srl_zoo_b200.TrainStep(mod, B, lr=0.0)
obs = torch.randn(B, 3, 224, 224, device="cuda")
eng.step(obs, obs)
for site, names, col in ((0, ["P:top", "P:free", "P:done", "M:top", "M:tmem", "M:pairs", "M:issued", "E:top", "E:tfull", "E:ld", "E:done"], 6),
                         (4, ["P:top", "P:free", "P:done", "M:top", "M:pairs", "M:dy", "M:issued", "D:top", "D:free", "D:done"], 6)):
    os.environ["SRLZ_DBG_SITE"] = str(site)
    dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
    lib.srlz_set_debug_buffer(_lib.ptr(dbg))
    eng.step(obs, obs)
    torch.cuda.synchronize()
    lib.srlz_set_debug_buffer(None)
    d = dbg.cpu()
    t0 = int(d[0, 0])
    print("== site %d" % site)
    print("it/rel " + " ".join("%8s" % n for n in names))
    for it in list(range(0, 4)) + list(range(30, 40)):
        print("%5d  " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(len(names))))
    print("cycles per output row (M:issued, rows 20..60): %.0f" % ((int(d[60, col]) - int(d[20, col])) / 40.0))
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

"""Development check (not a pytest): row-ring tcgen05 forward of the last decoder layer (dec12_rows_tc.cu) against the
halo-tile kernel it replaces, inside full train steps (lr = 0), then timings and a clock64 timeline of CTA 0.
usage: python tests/gpu_dec12_rows_check.py [B_timing]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import srl_zoo_b200
from srl_zoo_b200 import _lib
from srl_zoo_b200._lib import lib


def run(bs, mode):
    torch.manual_seed(1)
    mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.0)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(bs, 3, 224, 224, generator=g).cuda()
    nobs = torch.randn(bs, 3, 224, 224, generator=g).cuda()
    lib.srlz_set_tensor_cores(mode)
    t = eng.step(obs, nobs)
    torch.cuda.synchronize()
    lib.srlz_set_tensor_cores(1)
    gw = torch.cat([dict(mod.named_parameters())["model.decoder_conv.12.weight"].grad.reshape(-1), dict(mod.named_parameters())["model.decoder_conv.12.bias"].grad.reshape(-1)]).clone()
    return [d.clone() for d in eng.decoded], t.clone(), gw


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


ok = True
for bs in (1, 2, 5, 48):
    d_old, t_old, _ = run(bs, 5)
    d_new, t_new, g_new = run(bs, 1)
    _, _, g_old = run(bs, 6)   # same forward, per-tap wgrad: identical operands on both sides
    e = [rel(d_new[i], d_old[i]) for i in range(2)]
    el = abs(t_new[0].item() - t_old[0].item()) / abs(t_old[0].item())
    eg = rel(g_new, g_old)
    good = max(e) < 1e-5 and el < 1e-6 and eg < 1e-4
    ok &= good
    print("B=%d  decoded rel %.2e %.2e  loss rel %.2e  dec12 wgrad rel %.2e  %s" % (bs, e[0], e[1], el, eg, "OK" if good else "FAIL"))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for mode, name in ((2, "halo-tile / per-tap kernels"), (1, "row kernels")):
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
    print("%-28s B=%d  dec12.fwd %.3f ms/call  dec12.wgrad (+bias) %.3f ms/call" % (name, B, prof["dec12.fwd"][1] / prof["dec12.fwd"][0], prof["dec12.wgrad"][1] / 6.0))
    if mode == 1:
        for site, names, col in ((3, ["P:top", "P:free", "P:done", "M:top", "M:rows", "M:issued", "E:top", "E:tfull", "E:done"], 5),
                                 (5, ["A:top", "A:free", "A:done", "M:top", "M:ready", "M:issued", "G:top", "G:free", "G:done"], 5)):
            os.environ["SRLZ_DBG_SITE"] = str(site)
            dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
            lib.srlz_set_debug_buffer(_lib.ptr(dbg))
            eng.step(obs, nobs)
            torch.cuda.synchronize()
            lib.srlz_set_debug_buffer(None)
            d = dbg.cpu()
            t0 = int(d[0, 0])
            print("== site %d" % site)
            print("it/rel " + " ".join("%8s" % n for n in names))
            for it in list(range(0, 3)) + list(range(30, 36)):
                print("%5d  " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(len(names))))
            print("cycles per row (M:issued, 20..60): %.0f" % ((int(d[60, col]) - int(d[20, col])) / 40.0))
    lib.srlz_set_tensor_cores(1)
    del eng, mod
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

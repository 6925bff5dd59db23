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
    return [d.clone() for d in eng.decoded], t.clone()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


ok = True
for bs in (1, 2, 5):
    d_old, t_old = run(bs, 5)
    d_new, t_new = run(bs, 1)
    e = [rel(d_new[i], d_old[i]) for i in range(2)]
    el = abs(t_new[0].item() - t_old[0].item()) / abs(t_old[0].item())
    good = max(e) < 1e-5 and el < 1e-6
    ok &= good
    print("B=%d  decoded rel %.2e %.2e  loss rel %.2e  %s" % (bs, e[0], e[1], el, "OK" if good else "FAIL"))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for mode, name in ((5, "halo-tile kernel"), (1, "row-ring kernel")):
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
    print("%-18s B=%d  dec12.fwd %.3f ms/call" % (name, B, prof["dec12.fwd"][1] / prof["dec12.fwd"][0]))
    if mode == 1:
        os.environ["SRLZ_DBG_SITE"] = "3"
        dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
        lib.srlz_set_debug_buffer(_lib.ptr(dbg))
        eng.step(obs, nobs)
        torch.cuda.synchronize()
        lib.srlz_set_debug_buffer(None)
        d = dbg.cpu()
        names = ["P:top", "P:free", "P:done", "M:top", "M:rows", "M:issued", "E:top", "E:tfull", "E:done"]
        t0 = int(d[0, 0])
        print("it/rel " + " ".join("%8s" % n for n in names))
        for it in list(range(0, 3)) + list(range(30, 38)):
            print("%5d  " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(len(names))))
        print("cycles per output row pair (M:issued, 20..60): %.0f" % ((int(d[60, 5]) - int(d[20, 5])) / 40.0))
    lib.srlz_set_tensor_cores(1)
    del eng, mod
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

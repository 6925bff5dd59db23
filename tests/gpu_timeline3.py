"""clock64 timeline of CTA 0 of one call site inside a full train step (SRLZ_DBG_SITE: 1 dec12.dgrad, 2 dec9.dgrad)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import srl_zoo_b200
from srl_zoo_b200._lib import lib, ptr

bs = 128
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
eng = srl_zoo_b200.TrainStep(mod, bs)
obs = torch.randn(bs, 3, 224, 224, device="cuda")
nobs = torch.randn(bs, 3, 224, 224, device="cuda")
eng.step(obs, nobs)
for site in (1, 2):
    os.environ["SRLZ_DBG_SITE"] = str(site)
    dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
    lib.srlz_set_debug_buffer(ptr(dbg))
    eng.step(obs, nobs)
    torch.cuda.synchronize()
    lib.srlz_set_debug_buffer(None)
    d = dbg.cpu()
    names = ["P:top", "P:emp0", "P:gath", "P:stor", "P:arr0", "P:chunks", "P:pst", "P:bar", "M:top", "M:tempty", "M:commit", "E:top", "E:tfull", "E:end"]
    base = [int(v) for v in d[2] if int(v)]
    t0 = min(base) if base else 0
    print("site %d" % site)
    print("it " + " ".join("%8s" % n for n in names))
    for it in range(2, 10):
        print("%2d " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(14)))
    k = 11
    print("cycles per tile (epilogue top to top): %.0f" % ((int(d[40, k]) - int(d[10, k])) / 30.0))

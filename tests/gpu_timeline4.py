"""clock64 timeline of CTA 0 of the dec12 forward (halo kernel, N=16) inside a full train step (SRLZ_DBG_SITE=3)."""
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
os.environ["SRLZ_DBG_SITE"] = "3"
dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
lib.srlz_set_debug_buffer(ptr(dbg))
eng.step(obs, nobs)
torch.cuda.synchronize()
lib.srlz_set_debug_buffer(None)
d = dbg.cpu()
names = ["P:top", "P:free0", "P:st0", "P:free1", "P:st1", "M:top", "M:tempty", "M:g0rdy", "M:g1rdy", "M:g2rdy", "M:issued", "E:top", "E:tfull", "E:done", "P:landed", "P:issued"]
t0 = int(d[2, 0])
print("it " + " ".join("%8s" % n for n in names))
for it in range(2, 14):
    print("%2d " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(16)))
print("cycles per tile (producer top to top): %.0f" % ((int(d[40, 0]) - int(d[10, 0])) / 30.0))

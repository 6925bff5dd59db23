"""clock64 timeline of CTA 0 of the enc0 row-image wgrad kernel inside a full train step (SRLZ_DBG_SITE=4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import srl_zoo_b200
from srl_zoo_b200._lib import lib, ptr

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.0)
obs = torch.randn(bs, 3, 224, 224, device="cuda")
eng.step(obs, obs)
os.environ["SRLZ_DBG_SITE"] = "4"
dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
lib.srlz_set_debug_buffer(ptr(dbg))
eng.step(obs, obs)
torch.cuda.synchronize()
lib.srlz_set_debug_buffer(None)
d = dbg.cpu()
names = ["P:top", "P:free", "P:done", "M:top", "M:pairs", "M:dy", "M:issued", "D:top", "D:free", "D:done"]
t0 = int(d[0, 0])
print("it/rel " + " ".join("%8s" % n for n in names))
for it in list(range(0, 3)) + list(range(30, 40)):
    print("%5d  " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(len(names))))
print("cycles per output row (M:issued, rows 20..60): %.0f" % ((int(d[60, 6]) - int(d[20, 6])) / 40.0))

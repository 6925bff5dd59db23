"""clock64 timeline of CTA 0 of the enc0 forward (tcgen05, smem-patch producers), B=128."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import srl_zoo_b200
from srl_zoo_b200._lib import lib, ptr

bs = 128
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
mod.train()
obs = torch.randn(bs, 3, 224, 224, device="cuda")
with torch.no_grad():
    mod.getStates(obs)
dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
lib.srlz_set_debug_buffer(ptr(dbg))
with torch.no_grad():
    mod.getStates(obs)
torch.cuda.synchronize()
lib.srlz_set_debug_buffer(None)
d = dbg.cpu()
names = ["P:top", "P:emp0", "P:gath", "P:stor", "P:arr0", "P:chunks", "P:pst", "P:bar", "M:top", "M:tempty", "M:commit", "E:top", "E:tfull"]
t0 = int(d[0, 0])
print("it " + " ".join("%8s" % n for n in names))
for it in range(2, 12):
    print("%2d " % it + " ".join("%8d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%8s" % "-" for k in range(13)))
print("cycles per tile (producer top to top): %.0f" % ((int(d[40, 0]) - int(d[10, 0])) / 30.0))

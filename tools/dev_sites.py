"""Per-call-site ms per step of a full train step for a given build of the library (A/B of development builds).

    SRLZ_DEV_DEFS="-DSRLZ_WH_PW=16" SRLZ_DEV_OUT=libsrlz_v16.so python srl_zoo_b200/build.py --dev
    python tools/dev_sites.py srl_zoo_b200/csrc/libsrlz_v16.so [B] [filter,filter...]

Prints the call sites whose name contains one of the filters (default: all), the sum over all sites, and the un-profiled
step time (CUDA events over 10 steps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from srl_zoo_b200 import _lib

_lib.LIB_PATH = os.path.abspath(sys.argv[1])   # before the first use of the lazy handle
import srl_zoo_b200

bs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
filt = sys.argv[3].split(",") if len(sys.argv) > 3 else [""]
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
eng = srl_zoo_b200.TrainStep(mod, bs, lr=1e-4)
obs = torch.randn(bs, 3, 224, 224, device="cuda")
nobs = torch.randn(bs, 3, 224, 224, device="cuda")
for _ in range(4):
    eng.step(obs, nobs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = eng.step(obs, nobs)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 10
K = 6
_lib.prof_enable(True)
for _ in range(K):
    eng.step(obs, nobs)
torch.cuda.synchronize()
prof = _lib.prof_report()
_lib.prof_enable(False)
tot = sum(v[1] for v in prof.values()) / K
print("%s  step %.3f ms  (sum of sites %.3f)  loss0 %.6f" % (os.path.basename(sys.argv[1]), step_ms, tot, float(loss[0])))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    if any(f in k for f in filt):
        print("   %-16s %7.3f ms/step" % (k, v[1] / K))

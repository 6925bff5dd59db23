"""clock64 timeline of CTA 0 of one call site inside a full train step (development build only).

    python srl_zoo_b200/build.py --dev            # -> srl_zoo_b200/csrc/libsrlz_dev.so (compiled with -DSRLZ_DEV)
    python tools/dev_timeline.py SITE [B]         # SITE: 0 enc0.fwd, 2 dec9.dgrad, 3 dec12.fwd, 4 enc0.wgrad, 5 dec12.wgrad, 6 dec9.wgrad, 7 enc4.wgrad, 8 enc4.fwd, 9 dec9.fwd

Prints the 16 stamp slots (cycles relative to the first stamp) for a few iterations of CTA 0's loop."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from srl_zoo_b200 import _lib

_lib.LIB_PATH = os.environ.get("SRLZ_DEV_LIB", os.path.join(ROOT, "srl_zoo_b200", "csrc", "libsrlz_dev.so"))   # before the first use of the lazy handle
import srl_zoo_b200

site = int(sys.argv[1])
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.0)
obs = torch.randn(bs, 3, 224, 224, device="cuda")
for _ in range(2):
    eng.step(obs, obs)
dbg = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
real = _lib.lib._real if hasattr(_lib.lib, "_real") else None
fn = _lib._LazyLib._real.srlz_dev_set_debug_buffer
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = None
fn(C.c_void_p(dbg.data_ptr()), site)
eng.step(obs, obs)
torch.cuda.synchronize()
fn(None, -1)
d = dbg.cpu()
nz = d[d > 0]
t0 = int(nz.min()) if nz.numel() else 0
if site in (6, 7):   # halo wgrad: [16 tiles][64 slots]; producer warp 0: 3*kind + {before empty wait, after it, after arrive}; warp 12 at +32;
    d = d.reshape(16, 64)      # MMA warp: 16 dense full, 17+2c class c full, 18+2c class c issued
    for t in range(16):
        row = lambda lo, hi: " ".join(("%7d" % (int(d[t, k]) - t0)) if int(d[t, k]) else "%7s" % "-" for k in range(lo, hi))
        print("tile %2d  P0: %s" % (t, row(0, 15)))
        print("        P12: %s" % row(32, 47))
        print("        MMA: %s" % row(16, 26))
    sys.exit(0)
print("it    " + " ".join("%8d" % k for k in range(16)))
for it in list(range(0, 6)) + list(range(30, 44)):
    print("%5d " % it + " ".join(("%8d" % (int(d[it, k]) - t0)) if int(d[it, k]) else "%8s" % "-" for k in range(16)))

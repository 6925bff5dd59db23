"""clock64 timeline of CTA 0 of the halo kernel (enc4 fwd geometry, B=256)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from srl_zoo_b200 import ops
from srl_zoo_b200._lib import lib, ptr

dev = "cuda"
g = torch.Generator().manual_seed(3)
w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
names = ["P:top", "P:free0", "P:st0", "P:free1", "P:st1", "M:top", "M:tempty", "M:g0rdy", "M:g1rdy", "M:g2rdy", "M:issued", "E:top", "E:tfull", "E:done"]
for label, tconv, big, small, s, p in (("enc4 fwd", False, 56, 56, 1, 1), ("dec9 fwd", True, 111, 55, 2, 0)):
    fpk, dpk = ops.pack_conv_w(w.to(dev), tconv)
    fbf = ops.pack_conv_w_bf16(fpk)
    Bn = 256
    hin, hout = (small, big) if tconv else (big, small)
    x = torch.randn(Bn, hin, hin, 64, device=dev)
    out = torch.empty(Bn, hout, hout, 64, device=dev)
    dbg = torch.zeros(64, 16, dtype=torch.int64, device=dev)
    ops.conv64_tc(x, fbf, out, (big, big), (small, small), 3, s, p, tconv, want_stats=True, halo=True)
    lib.srlz_set_debug_buffer(ptr(dbg))
    ops.conv64_tc(x, fbf, out, (big, big), (small, small), 3, s, p, tconv, want_stats=True, halo=True)
    torch.cuda.synchronize()
    lib.srlz_set_debug_buffer(None)
    d = dbg.cpu()
    t0 = int(d[0, 0])
    print("== %s (cycles relative to first producer stamp)" % label)
    print("it " + " ".join("%9s" % n for n in names))
    for it in range(2, 14):
        print("%2d " % it + " ".join("%9d" % (int(d[it, k]) - t0) if int(d[it, k]) else "%9s" % "-" for k in range(14)))
    per_tile = (int(d[40, 10]) - int(d[10, 10])) / 30.0
    print("steady-state cycles per tile (MMA issue to issue): %.0f" % per_tile)

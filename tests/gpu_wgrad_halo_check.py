"""Halo-tile tcgen05 wgrad (csrc/wgrad_halo_tc.cu) vs fp64 torch at every layer geometry, and timing against the
per-tap tcgen05 kernel at BASELINE config-2 sizes.  Run under `timeout` on a GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F

import helpers as H
from srl_zoo_b200 import ops
from srl_zoo_b200._lib import lib


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    cases = [("conv s1 p1 56", False, 2, 56, 56, 1, 1), ("conv s1 p1 56 B=7", False, 7, 56, 56, 1, 1), ("conv s2 p1 27->14", False, 3, 27, 14, 2, 1),
             ("convT 6->13", True, 3, 13, 6, 2, 0), ("convT 13->27", True, 4, 27, 13, 2, 0), ("convT 27->55", True, 5, 55, 27, 2, 0),
             ("convT 55->111", True, 3, 111, 55, 2, 0)]
    for name, tconv, Bn, big, small, s, p in cases:
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        if tconv:
            x = torch.randn(Bn, 64, small, small, generator=g)
            dy = torch.randn(Bn, 64, big, big, generator=g)
            sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
            wr = w.double().clone().requires_grad_(True)
            act = F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double())
            (F.conv_transpose2d(act, wr, None, s) * dy.double()).sum().backward()
            args = (nhwc(dy).to(dev), nhwc(x).to(dev), (big, big), (small, small), 3, s, p)
            kw = dict(dense_scale=sc.to(dev), dense_shift=sh.to(dev))
        else:
            x = torch.randn(Bn, 64, big, big, generator=g)
            dy = torch.randn(Bn, 64, small, small, generator=g)
            wr = w.double().clone().requires_grad_(True)
            (F.conv2d(x.double(), wr, None, s, p) * dy.double()).sum().backward()
            args = (nhwc(x).to(dev), nhwc(dy).to(dev), (big, big), (small, small), 3, s, p)
            kw = {}
        lib.srlz_set_tensor_cores(1)
        gh = ops.wgrad64(*args, tensor_cores=True, **kw)
        lib.srlz_set_tensor_cores(2)   # per-tap tcgen05 kernel
        gt = ops.wgrad64(*args, tensor_cores=True, **kw)
        lib.srlz_set_tensor_cores(1)
        torch.cuda.synchronize()
        eh = H.rel_err(gh, wr.grad)
        per_tap = [(gh[:, :, ky, kx].double().cpu() - wr.grad[:, :, ky, kx]).abs().max().item() / wr.grad.abs().max().item() for ky in range(3) for kx in range(3)]
        print("%-22s halo rel %.3e  per-tap kernel rel %.3e  %s" % (name, eh, H.rel_err(gt, wr.grad), "OK" if eh < 2e-5 else "FAIL " + " ".join("%.1e" % e for e in per_tap)), flush=True)
    for name, Bn, big, small, s, p in (("enc4 wgrad B=256", 256, 56, 56, 1, 1), ("dec9 wgrad B=256", 256, 111, 55, 2, 0), ("dec6 wgrad B=256", 256, 55, 27, 2, 0)):
        xb = torch.randn(Bn, big, big, 64, device=dev)
        xs = torch.randn(Bn, small, small, 64, device=dev)
        for mode, label in ((1, "halo   "), (2, "per-tap")):
            lib.srlz_set_tensor_cores(mode)
            for _ in range(2):
                ops.wgrad64(xb, xs, (big, big), (small, small), 3, s, p, tensor_cores=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.wgrad64(xb, xs, (big, big), (small, small), 3, s, p, tensor_cores=True)
            e1.record()
            torch.cuda.synchronize()
            print("%s %s %.3f ms" % (name, label, e0.elapsed_time(e1) / 5), flush=True)
        lib.srlz_set_tensor_cores(1)


if __name__ == "__main__":
    main()

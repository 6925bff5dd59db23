"""halo-tile tcgen05 kernel bring-up (run under `timeout`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F

import helpers as H
from srl_zoo_b200 import ops


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    print(torch.cuda.get_device_name(0), flush=True)
    # direct stride-1 conv fwd + its dgrad (transposed s1), and stride-2 conv dgrad (transposed s2 pad 1)
    for name, Bn, big, small, s, p in (("conv s1 p1 56", 3, 56, 56, 1, 1), ("conv s1 p1 9", 2, 9, 9, 1, 1), ("conv s2 p1 27->14", 3, 27, 14, 2, 1)):
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        fpk, dpk = ops.pack_conv_w(w.to(dev), False)
        fbf, dbf = ops.pack_conv_w_bf16(fpk), ops.pack_conv_w_bf16(dpk)
        x = torch.randn(Bn, 64, big, big, generator=g)
        dy = torch.randn(Bn, 64, small, small, generator=g)
        xr = x.double().clone().requires_grad_(True)
        ref = F.conv2d(xr, w.double(), None, s, p)
        (ref * dy.double()).sum().backward()
        if s == 1:
            out = torch.full((Bn, small, small, 64), float("nan"), device=dev)
            _, stats = ops.conv64_tc(nhwc(x).to(dev), fbf, out, (big, big), (small, small), 3, s, p, False, want_stats=True, halo=True)
            torch.cuda.synchronize()
            print("%-20s fwd   halo rel %.3e  stats %.3e / %.3e" % (name, H.rel_err(nchw(out), ref), H.rel_err(stats[:64], ref.sum((0, 2, 3))),
                  H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3)))), flush=True)
        outd = torch.full((Bn, big, big, 64), float("nan"), device=dev)
        ops.conv64_tc(nhwc(dy).to(dev), dbf, outd, (big, big), (small, small), 3, s, p, True, halo=True)
        torch.cuda.synchronize()
        print("%-20s dgrad halo rel %.3e" % (name, H.rel_err(nchw(outd), xr.grad)), flush=True)
    for name, Bn, small in (("convT 6->13", 3, 6), ("convT 13->27", 2, 13), ("convT 27->55", 2, 27), ("convT 55->111", 2, 55)):
        big = 2 * small + 1
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        bias = torch.randn(64, generator=g)
        fpk, dpk = ops.pack_conv_w(w.to(dev), True)
        fbf = ops.pack_conv_w_bf16(fpk)
        x = torch.randn(Bn, 64, small, small, generator=g)
        sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
        ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), 2)
        out = torch.full((Bn, big, big, 64), float("nan"), device=dev)
        _, stats = ops.conv64_tc(nhwc(x).to(dev), fbf, out, (big, big), (small, small), 3, 2, 0, True, bias=bias.to(dev), want_stats=True, halo=True)
        torch.cuda.synchronize()
        print("%-20s fwd   halo rel %.3e  stats %.3e / %.3e" % (name, H.rel_err(nchw(out), ref), H.rel_err(stats[:64], ref.sum((0, 2, 3))),
              H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3)))), flush=True)
        refb = F.conv_transpose2d(F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double()), w.double(), bias.double(), 2)
        out.fill_(float("nan"))
        ops.conv64_tc(nhwc(x).to(dev), fbf, out, (big, big), (small, small), 3, 2, 0, True, bias=bias.to(dev), in_scale=sc.to(dev), in_shift=sh.to(dev), halo=True)
        torch.cuda.synchronize()
        print("%-20s fwd+bn halo rel %.3e" % (name, H.rel_err(nchw(out), refb)), flush=True)
    for name, tconv, Bn, big, small, s, p in (("enc4 fwd B=256", False, 256, 56, 56, 1, 1), ("dec9 fwd B=256", True, 256, 111, 55, 2, 0)):
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        fpk, dpk = ops.pack_conv_w(w.to(dev), tconv)
        fbf = ops.pack_conv_w_bf16(fpk)
        hin = small if tconv else big
        hout = big if tconv else small
        x = torch.randn(Bn, hin, hin, 64, device=dev)
        out = torch.empty(Bn, hout, hout, 64, device=dev)
        for halo, label in ((True, "halo   "), (False, "per-tap")):
            for _ in range(2):
                ops.conv64_tc(x, fbf, out, (big, big), (small, small), 3, s, p, tconv, want_stats=True, halo=halo)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv64_tc(x, fbf, out, (big, big), (small, small), 3, s, p, tconv, want_stats=True, halo=halo)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            macs = Bn * hout * hout * 64 * 64 * (9 if not tconv else 2.25)
            gb = (x.numel() + out.numel()) * 4 / 1e9
            print("%s tcgen05 %s %.3f ms  %.1f TFLOP/s (algorithmic)  %.0f GB/s (in+out)" % (name, label, ms, 2 * macs / ms / 1e9, gb / ms * 1e3), flush=True)


if __name__ == "__main__":
    main()

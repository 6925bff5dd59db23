"""tcgen05 kernel bring-up check (run under `timeout`): conv64_tc vs fp64 torch and vs the fp32 SIMT kernel."""
import os
import sys
import time

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
    cases = [("conv s1 p1 56", False, 2, 56, 56, 1, 1), ("conv s2 p1 27->14", False, 3, 27, 14, 2, 1),
             ("convT 6->13", True, 3, 13, 6, 2, 0), ("convT 55->111", True, 1, 111, 55, 2, 0),
             ("conv s1 p1 56 B=64", False, 64, 56, 56, 1, 1)]
    for name, tconv, Bn, big, small, s, p in cases:
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        bias = torch.randn(64, generator=g) if tconv else None
        fpk, dpk = ops.pack_conv_w(w.to(dev), tconv)
        fbf, dbf = ops.pack_conv_w_bf16(fpk), ops.pack_conv_w_bf16(dpk)
        torch.cuda.synchronize()
        if tconv:
            x = torch.randn(Bn, 64, small, small, generator=g)
            ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), s)
            out = torch.full((Bn, big, big, 64), float("nan"), device=dev)
            t0 = time.time()
            _, stats = ops.conv64_tc(nhwc(x).to(dev), fbf, out, (big, big), (small, small), 3, s, p, True, bias=bias.to(dev), want_stats=True)
            torch.cuda.synchronize()
            print("%-22s fwd   tc rel %.3e  stats %.3e / %.3e  (%.1f ms)" % (name, H.rel_err(nchw(out), ref), H.rel_err(stats[:64], ref.sum((0, 2, 3))),
                  H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3))), (time.time() - t0) * 1e3), flush=True)
            sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
            refb = F.conv_transpose2d(F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double()), w.double(), bias.double(), s)
            ops.conv64_tc(nhwc(x).to(dev), fbf, out, (big, big), (small, small), 3, s, p, True, bias=bias.to(dev), in_scale=sc.to(dev), in_shift=sh.to(dev))
            print("%-22s fwd+bn tc rel %.3e" % (name, H.rel_err(nchw(out), refb)), flush=True)
            dy = torch.randn(Bn, 64, big, big, generator=g)
            refd = F.conv2d(dy.double(), w.double(), None, s)
            outd = torch.full((Bn, small, small, 64), float("nan"), device=dev)
            ops.conv64_tc(nhwc(dy).to(dev), dbf, outd, (big, big), (small, small), 3, s, p, False)
            print("%-22s dgrad tc rel %.3e" % (name, H.rel_err(nchw(outd), refd)), flush=True)
        else:
            x = torch.randn(Bn, 64, big, big, generator=g)
            ref = F.conv2d(x.double(), w.double(), None, s, p)
            out = torch.full((Bn, small, small, 64), float("nan"), device=dev)
            t0 = time.time()
            _, stats = ops.conv64_tc(nhwc(x).to(dev), fbf, out, (big, big), (small, small), 3, s, p, False, want_stats=True)
            torch.cuda.synchronize()
            print("%-22s fwd   tc rel %.3e  stats %.3e / %.3e  (%.1f ms)" % (name, H.rel_err(nchw(out), ref), H.rel_err(stats[:64], ref.sum((0, 2, 3))),
                  H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3))), (time.time() - t0) * 1e3), flush=True)
            dy = torch.randn(Bn, 64, small, small, generator=g)
            xr = x.double().clone().requires_grad_(True)
            (F.conv2d(xr, w.double(), None, s, p) * dy.double()).sum().backward()
            outd = torch.full((Bn, big, big, 64), float("nan"), device=dev)
            ops.conv64_tc(nhwc(dy).to(dev), dbf, outd, (big, big), (small, small), 3, s, p, True)
            print("%-22s dgrad tc rel %.3e" % (name, H.rel_err(nchw(outd), xr.grad)), flush=True)
    # wgrad (tcgen05, MN-major operands, tap pairs stacked along M)
    for name, tconv, Bn, big, small, s, p in (("wgrad conv s1 56", False, 2, 56, 56, 1, 1), ("wgrad conv s2 27->14", False, 3, 27, 14, 2, 1),
                                              ("wgrad convT 6->13", True, 3, 13, 6, 2, 0), ("wgrad convT 27->55", True, 5, 55, 27, 2, 0)):
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        if tconv:
            x = torch.randn(Bn, 64, small, small, generator=g)
            dy = torch.randn(Bn, 64, big, big, generator=g)
            sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
            wr = w.double().clone().requires_grad_(True)
            act = F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double())
            (F.conv_transpose2d(act, wr, None, s) * dy.double()).sum().backward()
            gw = ops.wgrad64(nhwc(dy).to(dev), nhwc(x).to(dev), (big, big), (small, small), 3, s, p, dense_scale=sc.to(dev), dense_shift=sh.to(dev), tensor_cores=True)
        else:
            x = torch.randn(Bn, 64, big, big, generator=g)
            dy = torch.randn(Bn, 64, small, small, generator=g)
            wr = w.double().clone().requires_grad_(True)
            (F.conv2d(x.double(), wr, None, s, p) * dy.double()).sum().backward()
            gw = ops.wgrad64(nhwc(x).to(dev), nhwc(dy).to(dev), (big, big), (small, small), 3, s, p, tensor_cores=True)
        torch.cuda.synchronize()
        print("%-22s tc rel %.3e" % (name, H.rel_err(gw, wr.grad)), flush=True)
    for name, Bn, big, small, s, p in (("enc4 wgrad B=256", 256, 56, 56, 1, 1), ("dec9 wgrad B=256", 256, 111, 55, 2, 0)):
        xb = torch.randn(Bn, big, big, 64, device=dev)
        xs = torch.randn(Bn, small, small, 64, device=dev)
        for tcf, label in ((True, "tcgen05"), (False, "simt   ")):
            for _ in range(2):
                ops.wgrad64(xb, xs, (big, big), (small, small), 3, s, p, tensor_cores=tcf)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.wgrad64(xb, xs, (big, big), (small, small), 3, s, p, tensor_cores=tcf)
            e1.record()
            torch.cuda.synchronize()
            print("%s %s %.3f ms" % (name, label, e0.elapsed_time(e1) / 5), flush=True)
    # timing at the enc4 / dec9 sizes of BASELINE config 2
    for name, tconv, Bn, big, small, s, p in (("enc4 fwd B=256", False, 256, 56, 56, 1, 1), ("dec9 fwd B=256", True, 256, 111, 55, 2, 0)):
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
        fpk, dpk = ops.pack_conv_w(w.to(dev), tconv)
        fbf = ops.pack_conv_w_bf16(fpk)
        hin = small if tconv else big
        hout = big if tconv else small
        x = torch.randn(Bn, hin, hin, 64, device=dev)
        out = torch.empty(Bn, hout, hout, 64, device=dev)
        for fn, label in ((ops.conv64_tc, "tcgen05"), (ops.conv64, "simt   ")):
            wgt = fbf if fn is ops.conv64_tc else fpk
            for _ in range(2):
                fn(x, wgt, out, (big, big), (small, small), 3, s, p, tconv, want_stats=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn(x, wgt, out, (big, big), (small, small), 3, s, p, tconv, want_stats=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            macs = Bn * hout * hout * 64 * 64 * (9 if not tconv else 2.25)
            print("%s %s %.3f ms  %.1f TFLOP/s (algorithmic)" % (name, label, ms, 2 * macs / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()

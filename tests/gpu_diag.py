"""Diagnostic sweep (not a test): prints error tables for every stage of the CUDA path against torch / the oracle.
    python tests/gpu_diag.py > gpurun_out/diag.txt
"""
import ctypes as C
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from oracle import srl_oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H
import srl_zoo_b200
from srl_zoo_b200 import ops
from srl_zoo_b200._lib import lib, ptr, stream_ptr, check


def section(name):
    print("\n==== %s" % name, flush=True)


def guard(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc(file=sys.stdout)
    torch.cuda.synchronize()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def diag_ops():
    section("op-level: conv64 / wgrad64 / sgemm vs torch (fp64 reference)")
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    for name, Bn, big, small, k, s, p, tconv in (("enc4 conv3x3 s1 p1", 3, 56, 56, 3, 1, 1, False),
                                                  ("enc8 conv3x3 s2 p1", 3, 27, 14, 3, 2, 1, False),
                                                  ("dec convT3 s2 6->13", 3, 13, 6, 3, 2, 0, True),
                                                  ("dec convT3 s2 13->27", 2, 27, 13, 3, 2, 0, True),
                                                  ("dec convT3 s2 55->111", 1, 111, 55, 3, 2, 0, True)):
        if tconv:
            w = torch.randn(64, 64, k, k, generator=g) * 0.05          # IOHW
            x = torch.randn(Bn, 64, small, small, generator=g)
            bias = torch.randn(64, generator=g)
            ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), s)
            fpk, dpk = ops.pack_conv_w(w.to(dev), True)
            out = torch.empty(Bn, big, big, 64, device=dev)
            _, stats = ops.conv64(nhwc(x).to(dev), fpk, out, (big, big), (small, small), k, s, p, True, bias=bias.to(dev), want_stats=True)
            print("%-24s fwd   rel %.3e   stats rel %.3e" % (name, H.rel_err(nchw(out), ref),
                  H.rel_err(stats[:64], ref.sum((0, 2, 3)))), flush=True)
            # BN-on-load variant
            sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
            refb = F.conv_transpose2d(F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double()), w.double(), bias.double(), s)
            ops.conv64(nhwc(x).to(dev), fpk, out, (big, big), (small, small), k, s, p, True, bias=bias.to(dev), in_scale=sc.to(dev), in_shift=sh.to(dev))
            print("%-24s fwd+bnload rel %.3e" % (name, H.rel_err(nchw(out), refb)), flush=True)
            # dgrad of the transposed conv = direct conv of dy
            dy = torch.randn(Bn, 64, big, big, generator=g)
            refd = F.conv2d(dy.double(), w.double(), None, s)
            outd = torch.empty(Bn, small, small, 64, device=dev)
            ops.conv64(nhwc(dy).to(dev), dpk, outd, (big, big), (small, small), k, s, p, False)
            print("%-24s dgrad rel %.3e" % (name, H.rel_err(nchw(outd), refd)), flush=True)
            # wgrad: dW[ci][co] = sum in[ci] * dy[co]
            xr = x.double().requires_grad_(False)
            wr = w.double().clone().requires_grad_(True)
            (F.conv_transpose2d(xr, wr, None, s) * dy.double()).sum().backward()
            gw = ops.wgrad64(nhwc(dy).to(dev), nhwc(x).to(dev), (big, big), (small, small), k, s, p)
            print("%-24s wgrad rel %.3e" % (name, H.rel_err(gw, wr.grad)), flush=True)
        else:
            w = torch.randn(64, 64, k, k, generator=g) * 0.05          # OIHW
            x = torch.randn(Bn, 64, big, big, generator=g)
            ref = F.conv2d(x.double(), w.double(), None, s, p)
            fpk, dpk = ops.pack_conv_w(w.to(dev), False)
            out = torch.empty(Bn, small, small, 64, device=dev)
            _, stats = ops.conv64(nhwc(x).to(dev), fpk, out, (big, big), (small, small), k, s, p, False, want_stats=True)
            print("%-24s fwd   rel %.3e   stats rel %.3e / %.3e" % (name, H.rel_err(nchw(out), ref), H.rel_err(stats[:64], ref.sum((0, 2, 3))),
                  H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3)))), flush=True)
            dy = torch.randn(Bn, 64, small, small, generator=g)
            xr = x.double().clone().requires_grad_(True)
            wr = w.double().clone().requires_grad_(True)
            (F.conv2d(xr, wr, None, s, p) * dy.double()).sum().backward()
            outd = torch.empty(Bn, big, big, 64, device=dev)
            ops.conv64(nhwc(dy).to(dev), dpk, outd, (big, big), (small, small), k, s, p, True)
            print("%-24s dgrad rel %.3e" % (name, H.rel_err(nchw(outd), xr.grad)), flush=True)
            gw = ops.wgrad64(nhwc(x).to(dev), nhwc(dy).to(dev), (big, big), (small, small), k, s, p)
            print("%-24s wgrad rel %.3e" % (name, H.rel_err(gw, wr.grad)), flush=True)
    a = torch.randn(37, 201, generator=g)
    b = torch.randn(53, 201, generator=g)
    out = torch.empty(37, 53, device=dev)
    ops.sgemm(a.to(dev), b.to(dev), out, bias=None, trans_b=True)
    print("sgemm nt rel %.3e" % H.rel_err(out, a.double() @ b.double().t()))
    out2 = torch.empty(201, 53, device=dev)
    ops.sgemm(a.to(dev), torch.randn(37, 53, generator=torch.Generator().manual_seed(1)).to(dev), out2, trans_a=True)
    print("sgemm tn rel %.3e" % H.rel_err(out2, a.double().t() @ torch.randn(37, 53, generator=torch.Generator().manual_seed(1)).double()))


def saved_tensors(mod, saved, Bn):
    names = lib.srlz_saved_names().decode().split(",")
    offs = (C.c_size_t * 32)()
    n = lib.srlz_saved_layout(Bn, H.S, int(mod.model.is_vae), offs, 32)
    shapes = {"y1": (Bn, 112, 112, 64), "a1": (Bn, 56, 56, 64), "y2": (Bn, 56, 56, 64), "a2": (Bn, 27, 27, 64), "y3": (Bn, 14, 14, 64),
              "a3": (Bn, 6, 6, 64), "lat": (Bn * (2 if mod.model.is_vae else 1), H.S), "z": (Bn, H.S), "d0": (Bn, 6, 6, 64),
              "y4": (Bn, 13, 13, 64), "y5": (Bn, 27, 27, 64), "y6": (Bn, 55, 55, 64), "y7": (Bn, 111, 111, 64), "bnsave": (7, 5, 64)}
    out = {}
    for i in range(n):
        nm = names[i]
        if nm in shapes:
            cnt = int(np.prod(shapes[nm]))
            out[nm] = saved[offs[i]:offs[i] + 4 * cnt].view(torch.float32).view(shapes[nm])
    return out


def diag_forward(kind):
    section("forward intermediates, kind=%s (train mode, B=3) vs torch CPU" % kind)
    losses = {"ae": ["autoencoder"], "vae": ["vae"]}[kind]
    mod, P, Bf = H.make_pair(kind, losses)
    cpu, dev = H.inputs(3)
    x = dev["obs"]
    eps = dev["eps"][0][:3] if kind == "vae" else None
    mod.train()
    cn = mod.model
    net = cn.net_struct()
    wpack = mod._scratch.get_pack(cn, x.device)
    check(lib.srlz_pack_weights(C.byref(net), ptr(wpack), stream_ptr()))
    ws = mod._scratch.get_ws(3, cn, x.device)
    saved = torch.zeros(lib.srlz_saved_bytes(3, H.S, int(cn.is_vae)), dtype=torch.uint8, device=x.device)
    lat = torch.empty(3, H.S, device=x.device)
    lv = torch.empty(3, H.S, device=x.device) if kind == "vae" else None
    dec = torch.empty(3, 3, 224, 224, device=x.device)
    lo = torch.zeros(2, device=x.device)
    check(lib.srlz_forward(C.byref(net), ptr(wpack), ptr(x), None, ptr(eps), 3, 1, ptr(lat), ptr(lv), ptr(dec), ptr(x), ptr(lo),
                           ptr(saved), ptr(ws), stream_ptr()), "forward")
    torch.cuda.synchronize()
    sv = saved_tensors(mod, saved, 3)
    # torch reference, layer by layer
    with torch.no_grad():
        xr = cpu["obs"][:3]
        p = lambda k: P[k].detach()
        t = F.conv2d(xr, p("model.encoder_conv.0.weight"), None, 2, 3)
        print("y1 rel %.3e" % H.rel_err(nchw(sv["y1"]), t))
        def bnrelu(t, pre):
            m = t.mean((0, 2, 3)); v = t.var((0, 2, 3), unbiased=False)
            return F.relu((t - m.view(1, -1, 1, 1)) / torch.sqrt(v.view(1, -1, 1, 1) + 1e-5) * p(pre + ".weight").view(1, -1, 1, 1) + p(pre + ".bias").view(1, -1, 1, 1)), m, v
        a, m, v = bnrelu(t, "model.encoder_conv.1")
        print("bn1 mean rel %.3e  invstd rel %.3e" % (H.rel_err(sv["bnsave"][0, 2], m), H.rel_err(sv["bnsave"][0, 3], 1 / torch.sqrt(v + 1e-5))))
        a = F.max_pool2d(a, 3, 2, 1)
        print("a1 rel %.3e" % H.rel_err(nchw(sv["a1"]), a))
        t = F.conv2d(a, p("model.encoder_conv.4.weight"), None, 1, 1)
        print("y2 rel %.3e" % H.rel_err(nchw(sv["y2"]), t))
        a, m, v = bnrelu(t, "model.encoder_conv.5")
        a = F.max_pool2d(a, 3, 2)
        print("a2 rel %.3e" % H.rel_err(nchw(sv["a2"]), a))
        t = F.conv2d(a, p("model.encoder_conv.8.weight"), None, 2, 1)
        print("y3 rel %.3e" % H.rel_err(nchw(sv["y3"]), t))
        a, m, v = bnrelu(t, "model.encoder_conv.9")
        a = F.max_pool2d(a, 3, 2)
        print("a3 rel %.3e" % H.rel_err(nchw(sv["a3"]), a))
        flat = a.reshape(3, -1)
        if kind == "vae":
            mu = F.linear(flat, p("model.encoder_fc1.weight"), p("model.encoder_fc1.bias"))
            lvr = F.linear(flat, p("model.encoder_fc2.weight"), p("model.encoder_fc2.bias"))
            print("mu rel %.3e  logvar rel %.3e" % (H.rel_err(lat, mu), H.rel_err(lv, lvr)))
            z = cpu["eps"][0][:3] * torch.exp(0.5 * lvr) + mu
            print("z rel %.3e  klsum rel %.3e" % (H.rel_err(sv["z"], z), H.rel_err(lo[1], (1 + lvr - mu * mu - lvr.exp()).sum())))
        else:
            z = F.linear(flat, p("model.encoder_fc.0.weight"), p("model.encoder_fc.0.bias"))
            print("states rel %.3e  norm-rel %.3e" % (H.rel_err(lat, z), H.norm_rel(lat, z)))
        d = F.linear(z, p("model.decoder_fc.0.weight"), p("model.decoder_fc.0.bias")).view(3, 64, 6, 6)
        print("d0 rel %.3e" % H.rel_err(nchw(sv["d0"]), d))
        for j, (idx, bnidx, nm) in enumerate(((0, 1, "y4"), (3, 4, "y5"), (6, 7, "y6"), (9, 10, "y7"))):
            t = F.conv_transpose2d(d, p("model.decoder_conv.%d.weight" % idx), p("model.decoder_conv.%d.bias" % idx), 2)
            print("%s rel %.3e" % (nm, H.rel_err(nchw(sv[nm]), t)))
            d, m, v = bnrelu(t, "model.decoder_conv.%d" % bnidx)
        t = F.conv_transpose2d(d, p("model.decoder_conv.12.weight"), p("model.decoder_conv.12.bias"), 2)
        print("decoded rel %.3e   sse rel %.3e" % (H.rel_err(dec, t), H.rel_err(lo[0], ((t - xr) ** 2).sum())))


def diag_step(kind, losses, use_fwd=False, use_inv=False, bs=2, steps=2):
    section("engine step kind=%s losses=%s vs oracle (B=%d)" % (kind, losses, bs))
    mod, P, Bf = H.make_pair(kind, losses)
    cpu, dev = H.inputs(bs)
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.005)
    opt = O.Adam(P, lr=0.005)
    for step in range(steps):
        t = eng.step(dev["obs"], dev["nobs"], dev["actions"], dev["eps"][0], dev["eps"][1], dev["rects"][0], dev["rects"][1])
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone().cpu() for n, p in mod.named_parameters()}
        r = O.train_step(kind, P, Bf, cpu["obs"], cpu["nobs"], cpu["actions"], cpu["eps"][0], cpu["eps"][1], cpu["rects"][0],
                         cpu["rects"][1], use_forward=use_fwd, use_inverse=use_inv, optimizer=None)
        names = eng.loss_names()
        for i, n in enumerate(names):
            if n:
                print("step%d loss %-20s cuda %.8e  oracle %.8e  rel %.2e" % (step, n, t[i].item(), r["losses"][n],
                      abs(t[i].item() - r["losses"][n]) / max(abs(r["losses"][n]), 1e-30)))
        print("step%d states norm-rel %.3e   decoded rel %.3e" % (step, H.norm_rel(eng.lat[0], r["states"] if kind != "vae" else r["mu"]),
              H.rel_err(eng.decoded[0], r["decoded"])))
        for k, pp in P.items():
            if pp.grad is None:
                continue
            print("   grad %-40s rel %.3e cos %.6f  |ref| %.3e" % (k, H.rel_err(grads[k], pp.grad), H.cosine(grads[k], pp.grad), pp.grad.abs().max().item()))
        opt.step(P)
        sd = mod.state_dict()
        worst = ("", 0.0)
        for k in sd:
            ref = Bf[k] if O.is_buffer(k) else P[k].detach()
            e = H.rel_err(sd[k].float(), ref.float())
            if e > worst[1]:
                worst = (k, e)
        print("step%d worst post-Adam param/buffer: %s rel %.3e" % (step, worst[0], worst[1]))


def diag_autograd(kind, losses):
    section("drop-in module + loss functions through autograd, kind=%s" % kind)
    from srl_zoo_b200 import losses as L
    mod, P, Bf = H.make_pair(kind, losses)
    cpu, dev = H.inputs(2)
    mod.train()
    lm = L.LossManager(mod, None)
    if kind == "vae":
        torch.manual_seed(5)
        (d, mu, lv), (nd, nmu, nlv) = mod(dev["obs"]), mod(dev["nobs"])
        s, ns = mod.getStates(dev["obs"]), mod.getStates(dev["nobs"])
        L.kullbackLeiblerLoss(mu, nmu, lv, nlv, loss_manager=lm, beta=1.0)
        L.generationLoss(d, nd, dev["obs"], dev["nobs"], weight=0.5e-6, loss_manager=lm)
    else:
        (s, d), (ns, nd) = mod(dev["obs"]), mod(dev["nobs"])
        L.autoEncoderLoss(dev["obs"], d, dev["nobs"], nd, weight=1.0, loss_manager=lm)
    if "forward" in losses:
        L.forwardModelLoss(mod.forwardModel(s, dev["actions"]), ns, weight=1.0, loss_manager=lm)
    if "inverse" in losses:
        L.inverseModelLoss(mod.inverseModel(s, ns), dev["actions"], weight=2.0, loss_manager=lm)
    loss = lm.computeTotalLoss()
    loss.backward()
    torch.cuda.synchronize()
    if kind == "vae":
        torch.manual_seed(5)
        e0 = torch.empty(2, H.S, device="cuda").normal_().cpu()
        e1 = torch.empty(2, H.S, device="cuda").normal_().cpu()
    else:
        e0 = e1 = None
    r = O.train_step(kind, P, Bf, cpu["obs"], cpu["nobs"], cpu["actions"], e0, e1, use_forward="forward" in losses,
                     use_inverse="inverse" in losses)
    print("total cuda %.8e oracle %.8e" % (loss.item(), r["total"]))
    for (k, pp) in P.items():
        g = dict(mod.named_parameters())[k].grad
        if pp.grad is None:
            print("   grad %-40s oracle None, cuda %s" % (k, "None" if g is None else "tensor"))
            continue
        print("   grad %-40s rel %.3e cos %.6f" % (k, H.rel_err(g, pp.grad), H.cosine(g, pp.grad)))


def diag_speed():
    section("speed (B=256, AE engine step)")
    mod, P, Bf = H.make_pair("ae", ["autoencoder"])
    bs = 256
    obs = torch.randn(bs, 3, 224, 224, device="cuda")
    nobs = torch.randn(bs, 3, 224, 224, device="cuda")
    eng = srl_zoo_b200.TrainStep(mod, bs)
    for _ in range(2):
        eng.step(obs, nobs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        t = eng.step(obs, nobs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("ms/step %.2f  images/s %.0f  loss %.5f" % (ms, 2 * bs / ms * 1e3, t[0].item()))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    guard(diag_ops)
    guard(lambda: diag_forward("ae"))
    guard(lambda: diag_forward("vae"))
    guard(lambda: diag_step("ae", ["autoencoder"]))
    guard(lambda: diag_step("dae", ["dae"]))
    guard(lambda: diag_step("vae", ["vae"]))
    guard(lambda: diag_step("ae", ["autoencoder", "forward", "inverse"], True, True))
    guard(lambda: diag_step("vae", ["vae", "forward", "inverse"], True, True))
    guard(lambda: diag_autograd("ae", ["autoencoder", "forward", "inverse"]))
    guard(lambda: diag_autograd("vae", ["vae"]))
    guard(diag_speed)

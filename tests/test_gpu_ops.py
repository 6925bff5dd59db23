"""GPU: op-level kernels of libsrlz against torch (fp64 reference of the same op)."""
import pytest
import torch
import torch.nn.functional as F

import helpers as H

pytestmark = pytest.mark.gpu
TOL = 2e-5  # fp32 accumulation over K <= 576*... terms against an fp64 reference


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("Bn,big,small,s,p", [(3, 56, 56, 1, 1), (7, 56, 56, 1, 1), (3, 27, 14, 2, 1), (5, 9, 5, 2, 1)])
def test_conv3x3_family(Bn, big, small, s, p):
    """conv3x3 (models/models.py:217-226) on the product tcgen05 kernels: forward + BN statistics (halo-tile kernel at stride 1,
    per-tap pipeline at stride 2), dgrad (halo-tile kernel), wgrad (halo-tile kernel) against fp64 torch"""
    from srl_zoo_b200 import ops
    g = torch.Generator().manual_seed(3)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    x = torch.randn(Bn, 64, big, big, generator=g)
    dy = torch.randn(Bn, 64, small, small, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, s, p)
    (ref * dy.double()).sum().backward()
    fpk, dpk = ops.pack_conv_w(w.cuda(), False)
    fbf, dbf = ops.pack_conv_w_bf16(fpk), ops.pack_conv_w_bf16(dpk)
    out = torch.full((Bn, small, small, 64), float("nan"), device="cuda")
    _, stats = ops.conv64(nhwc(x).cuda(), fbf, out, (big, big), (small, small), 3, s, p, False, want_stats=True)
    assert H.rel_err(nchw(out), ref) < TOL
    assert H.rel_err(stats[:64], ref.sum((0, 2, 3))) < TOL and H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3))) < TOL
    outd = torch.full((Bn, big, big, 64), float("nan"), device="cuda")
    ops.conv64(nhwc(dy).cuda(), dbf, outd, (big, big), (small, small), 3, s, p, True)
    assert H.rel_err(nchw(outd), xr.grad) < TOL
    gw = ops.wgrad64(nhwc(x).cuda(), nhwc(dy).cuda(), (big, big), (small, small), 3, s, p)
    assert H.rel_err(gw, wr.grad) < TOL


@pytest.mark.parametrize("Bn,small", [(3, 6), (4, 13), (5, 27), (3, 55), (4, 1)])
def test_conv_transpose3x3_family(Bn, small):
    """ConvTranspose2d(64,64,3,stride=2) (models/models.py:66-78) on the product tcgen05 kernels: forward (+bias, +BN/ReLU on
    load, +BN statistics; halo-tile kernel), dgrad (stride-2 gather), wgrad (halo-tile kernel) against fp64 torch"""
    from srl_zoo_b200 import ops
    big = 2 * small + 1
    g = torch.Generator().manual_seed(4)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    x = torch.randn(Bn, 64, small, small, generator=g)
    bias = torch.randn(64, generator=g)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    dy = torch.randn(Bn, 64, big, big, generator=g)
    fpk, dpk = ops.pack_conv_w(w.cuda(), True)
    fbf, dbf = ops.pack_conv_w_bf16(fpk), ops.pack_conv_w_bf16(dpk)
    out = torch.full((Bn, big, big, 64), float("nan"), device="cuda")
    ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), 2)
    _, stats = ops.conv64(nhwc(x).cuda(), fbf, out, (big, big), (small, small), 3, 2, 0, True, bias=bias.cuda(), want_stats=True)
    assert H.rel_err(nchw(out), ref) < TOL
    assert H.rel_err(stats[:64], ref.sum((0, 2, 3))) < TOL and H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3))) < TOL
    act = F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double())
    wr = w.double().requires_grad_(True)
    refb = F.conv_transpose2d(act, wr, bias.double(), 2)
    out.fill_(float("nan"))
    ops.conv64(nhwc(x).cuda(), fbf, out, (big, big), (small, small), 3, 2, 0, True, bias=bias.cuda(), in_scale=sc.cuda(), in_shift=sh.cuda())
    assert H.rel_err(nchw(out), refb) < TOL
    (refb * dy.double()).sum().backward()
    gw = ops.wgrad64(nhwc(dy).cuda(), nhwc(x).cuda(), (big, big), (small, small), 3, 2, 0, dense_scale=sc.cuda(), dense_shift=sh.cuda())
    assert H.rel_err(gw, wr.grad) < TOL
    outd = torch.full((Bn, small, small, 64), float("nan"), device="cuda")
    ops.conv64(nhwc(dy).cuda(), dbf, outd, (big, big), (small, small), 3, 2, 0, False)
    assert H.rel_err(nchw(outd), F.conv2d(dy.double(), w.double(), None, 2)) < TOL


@pytest.mark.parametrize("Bn,masked", [(1, False), (3, True), (48, False)])
def test_first_layer_row_kernels(Bn, masked):
    """Conv2d(3,64,7,2,3) (models/models.py:49) forward + BN statistics and weight gradient on the row-image tcgen05 kernels
    (csrc/enc0_rows_tc.cu), with the DAE rectangle zeroed on load (preprocessing/data_loader.py:55-63); B=48 makes every CTA
    walk a multi-image row range.  fp64 torch (on the GPU) is the reference."""
    import numpy as np
    from oracle import srl_oracle as O
    from srl_zoo_b200 import ops
    g = torch.Generator().manual_seed(6)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).cuda()
    x = torch.randn(Bn, 3, 224, 224, generator=g).cuda()
    dy = torch.randn(Bn, 64, 112, 112, generator=g).cuda()
    rects, xin = None, x
    if masked:
        r = O.sample_rects(Bn, rng=np.random.RandomState(3))
        rects, xin = torch.from_numpy(r).cuda(), O.apply_occlusion(x, r)
    wr = w.double().requires_grad_(True)
    ref = F.conv2d(xin.double(), wr, None, 2, 3)
    (ref * dy.double()).sum().backward()
    y, stats = ops.enc0_fwd(x, w, rects=rects, want_stats=True)
    assert H.rel_err(nchw(y), ref) < TOL
    assert H.rel_err(stats[:64], ref.sum((0, 2, 3))) < TOL and H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3))) < TOL
    gw = ops.enc0_wgrad(x, nhwc(dy), rects=rects)
    assert H.rel_err(gw, wr.grad) < TOL


@pytest.mark.parametrize("Bn,explicit", [(1, True), (2, False), (5, True), (48, False)])
def test_last_layer_row_kernels(Bn, explicit):
    """ConvTranspose2d(64,3,4,2) (models/models.py:82) on relu(bn(y7)): forward + fused squared error (row-ring kernel), weight +
    bias gradient (row-staged kernel) and input gradient with ReLU mask + BatchNorm-backward sums (per-tap kernel, special
    producer), with the gradient source explicit or recomputed as coef*(decoded-target).  fp64 torch (on the GPU) is the reference."""
    from srl_zoo_b200 import ops
    g = torch.Generator().manual_seed(8)
    dev = "cuda"
    w = (torch.randn(64, 3, 4, 4, generator=g) * 0.05).to(dev)
    bias = torch.randn(3, generator=g).to(dev)
    y7 = torch.randn(Bn, 111, 111, 64, generator=g).to(dev)
    gamma, beta = (torch.rand(64, generator=g) + 0.5).to(dev), (torch.randn(64, generator=g) * 0.1).to(dev)
    mean, invstd = (torch.randn(64, generator=g) * 0.1).to(dev), (torch.rand(64, generator=g) + 0.7).to(dev)
    scale = gamma * invstd
    shift = beta - mean * scale
    target = torch.randn(Bn, 3, 224, 224, generator=g).to(dev)
    coef = 0.37
    a = F.relu(y7.double() * scale.double() + shift.double()).permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = w.double().requires_grad_(True), bias.double().requires_grad_(True)
    ref = F.conv_transpose2d(a, wr, br, 2)
    decoded, sse = ops.dec12_fwd(y7, scale, shift, w, bias, target=target)
    assert H.rel_err(decoded, ref) < TOL
    assert H.rel_err(sse, ((ref - target.double()) ** 2).sum()) < 1e-5
    if explicit:
        gdec = torch.randn(Bn, 3, 224, 224, generator=g).to(dev)
        gref = gdec.double()
        out = ops.dec12_bwd(y7, scale, shift, mean, invstd, w, g_decoded=gdec)
    else:
        gref = coef * (decoded.double() - target.double())
        out = ops.dec12_bwd(y7, scale, shift, mean, invstd, w, decoded=decoded, target=target, coef=coef)
    (ref * gref).sum().backward()
    gw, gb, dz, sums = out
    assert H.rel_err(gw, wr.grad) < TOL and H.rel_err(gb, br.grad) < TOL
    dz_ref = (a.grad * (a.detach() > 0)).permute(0, 2, 3, 1)
    assert H.rel_err(dz, dz_ref) < TOL
    xhat = (y7.double() - mean.double()) * invstd.double()
    assert H.rel_err(sums[:64], dz_ref.sum((0, 1, 2))) < 1e-4 and H.rel_err(sums[64:], (dz_ref * xhat).sum((0, 1, 2))) < 1e-4


def test_sgemm_sse_adam():
    from srl_zoo_b200 import ops
    g = torch.Generator().manual_seed(5)
    a, b, c = torch.randn(37, 201, generator=g), torch.randn(53, 201, generator=g), torch.randn(37, 53, generator=g)
    bias = torch.randn(53, generator=g)
    out = torch.empty(37, 53, device="cuda")
    ops.sgemm(a.cuda(), b.cuda(), out, bias=bias.cuda(), trans_b=True)
    assert H.rel_err(out, a.double() @ b.double().t() + bias.double()) < TOL
    out2 = torch.empty(201, 53, device="cuda")
    ops.sgemm(a.cuda(), c.cuda(), out2, trans_a=True)
    assert H.rel_err(out2, a.double().t() @ c.double()) < TOL
    ops.sgemm(a.cuda(), c.cuda(), out2, trans_a=True, accumulate=True)
    assert H.rel_err(out2, 2 * (a.double().t() @ c.double())) < TOL
    x, y = torch.randn(3, 3, 224, 224, generator=g), torch.randn(3, 3, 224, 224, generator=g)
    assert H.rel_err(ops.sse(x.cuda(), y.cuda()), ((x.double() - y.double()) ** 2).sum()) < 1e-6
    assert H.rel_err(ops.sse(x.cuda()[:, :, :7, :5].contiguous(), y.cuda()[:, :, :7, :5].contiguous()),
                     ((x[:, :, :7, :5].double() - y[:, :, :7, :5].double()) ** 2).sum()) < 1e-6
    assert torch.equal(ops.mse_grad(x.cuda(), y.cuda(), 0.25).cpu(), 0.25 * (x - y))
    # Adam: identical gradients => identical trajectory as th.optim.Adam (models/learner.py:199)
    p = torch.randn(10007, generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=0.005)
    pc, m, v = p.cuda(), torch.zeros(10007, device="cuda"), torch.zeros(10007, device="cuda")
    for step in range(1, 6):
        gr = torch.randn(10007, generator=g) * (10.0 ** (step - 3))
        ref.grad = gr.clone()
        opt.step()
        ops.adam_step(pc, gr.cuda(), m, v, 0.005, step)
        assert H.rel_err(pc, ref.detach()) < 1e-6, step


def test_preprocess_u8_bit_exact():
    """uint8 RGB HWC frames -> normalised (B,3,W,H) fp32 on the device, BIT-EXACT against the reference's
    truth table (tests/golden/preprocess_u8.npz `lut`, recorded from preprocessing/utils.py preprocessInput) and against the
    oracle's host restatement; then the DAE rectangle applied by the first-layer kernel on load equals the reference's
    zeroing of the normalised image (preprocessing/data_loader.py:55-63)."""
    import os
    import numpy as np
    from oracle import srl_oracle as O
    from srl_zoo_b200 import ops
    fx = np.load(os.path.join(H.GOLD, "preprocess_u8.npz"))
    rng = np.random.RandomState(5)
    frames = rng.randint(0, 256, (5, 224, 224, 3)).astype(np.uint8)
    frames[0, :, :, :] = (np.arange(224 * 224 * 3) % 256).reshape(224, 224, 3).astype(np.uint8)   # every byte value in every channel
    got = ops.preprocess_u8(torch.from_numpy(frames).cuda()).cpu().numpy()
    want = fx["lut"][frames, np.arange(3)[None, None, None, :]].transpose(0, 3, 2, 1)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert np.array_equal(got[1:2], O.preprocess_u8(frames[1]).numpy())
    with pytest.raises(RuntimeError):
        ops.preprocess_u8(torch.from_numpy(frames).cuda().float())


def test_step_host_uint8_matches_float_path():
    """TrainStep.step_host on pinned uint8 HWC frames == step() on the tensors the reference's loader would have delivered
    (normalised on the host by the oracle's restatement of the loader), bit for bit: losses and parameters after the step."""
    import numpy as np
    import srl_zoo_b200
    from oracle import srl_oracle as O
    bs = 2
    rng = np.random.RandomState(9)
    f0, f1 = [rng.randint(0, 256, (bs, 224, 224, 3)).astype(np.uint8) for _ in range(2)]
    host = lambda f: torch.cat([O.preprocess_u8(f[i]) for i in range(bs)]).contiguous()
    act = torch.from_numpy(rng.randint(0, 6, (bs, 1))).long()
    out = {}
    for mode in ("float_device", "uint8_host"):
        mod, _, _ = H.make_pair("ae", ["autoencoder", "forward", "inverse"])
        eng = srl_zoo_b200.TrainStep(mod, bs, lr=1e-3)
        if mode == "float_device":
            t = eng.step(host(f0).cuda(), host(f1).cuda(), act.cuda()).cpu().clone()
        else:
            t = eng.step_host(torch.from_numpy(f0).pin_memory(), torch.from_numpy(f1).pin_memory(), act.pin_memory()).clone()
        torch.cuda.synchronize()
        out[mode] = (t, eng.flat_p.detach().cpu().clone())
        assert eng.h2d_bytes_per_step(True) == 2 * bs * 224 * 224 * 3 + bs * 8
    assert torch.equal(out["float_device"][0], out["uint8_host"][0]) and torch.equal(out["float_device"][1], out["uint8_host"][1])


def test_heads_reject_bad_actions():
    """the heads index their weights with the action value: shape / dtype / device are checked on the host (RuntimeError), the
    value range on the device (NaN loss, no out-of-bounds access) -- torch's scatter_ / CrossEntropyLoss raise in the reference"""
    import srl_zoo_b200
    mod, _, _ = H.make_pair("ae", ["autoencoder", "forward", "inverse"])
    cpu, dev = H.inputs(2)
    eng = srl_zoo_b200.TrainStep(mod, 2, lr=1e-3)
    for bad in (dev["actions"].int(), dev["actions"].reshape(-1), cpu["actions"]):
        with pytest.raises(RuntimeError):
            eng.step(dev["obs"], dev["nobs"], bad)
    oob = dev["actions"].clone()
    oob[0, 0] = 6
    t = eng.step(dev["obs"], dev["nobs"], oob, training=False)
    assert torch.isnan(t[2]) and torch.isnan(t[3]) and torch.isfinite(t[0])


@pytest.mark.parametrize("Bn,H,pad", [(3, 112, 1), (5, 56, 0), (7, 14, 0), (150, 14, 0), (2, 112, 1)])
def test_bn_relu_pool_forward(Bn, H, pad):
    """BatchNorm + ReLU + MaxPool2d(3, 2, pad) of the three pooled encoder stages (models/models.py:50-52,55-57,60-62) on the TMA-fed
    kernel (csrc/pool_tma.cu): values against fp64 torch; argmax = FIRST maximal tap in scan order (torch semantics), incl. windows
    with no positive tap (every relu value 0: the first valid tap wins) and negative BatchNorm scales."""
    from srl_zoo_b200 import ops
    g = torch.Generator().manual_seed(21)
    y = torch.randn(Bn, H, H, 64, generator=g).cuda()
    y[0, : H // 2] -= 6.0                                            # a region where no tap is positive
    sc = (torch.rand(64, generator=g) + 0.5).cuda()
    sc[::5] *= -1.0
    sh = (torch.randn(64, generator=g) * 0.3).cuda()
    out, am = ops.bn_relu_pool(y, sc, sh, pad)
    v = torch.relu(y.double() * sc.double() + sh.double()).permute(0, 3, 1, 2)
    ref, idx = F.max_pool2d(v, 3, 2, pad, return_indices=True)
    assert (nchw(out).double() - ref).abs().max().item() <= 1e-6 * max(ref.abs().max().item(), 1.0)
    # the tap the kernel recorded -> input position; torch's flat index of the first maximum must be the same position wherever the
    # fp32 and fp64 orderings agree (they differ only at ties closer than fp32 rounding)
    PH = ref.shape[2]
    ph = torch.arange(PH, device="cuda").view(1, PH, 1, 1)
    pw = torch.arange(PH, device="cuda").view(1, 1, PH, 1)
    amn = am.long()
    pos = (2 * ph - pad + amn // 3) * H + (2 * pw - pad + amn % 3)
    same = (pos.permute(0, 3, 1, 2) == idx)
    assert same.float().mean().item() > 0.9999
    assert bool(same[0, :, : PH // 2 - 1].all())                     # the all-zero region: first valid tap, exactly as torch
    out2, none = ops.bn_relu_pool(y, sc, sh, pad, want_argmax=False)
    assert none is None and torch.equal(out, out2)

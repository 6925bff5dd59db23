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


@pytest.mark.parametrize("Bn,big,small,s,p", [(3, 56, 56, 1, 1), (3, 27, 14, 2, 1), (5, 9, 5, 2, 1)])
def test_conv3x3_family(Bn, big, small, s, p):
    """conv3x3 (models/models.py:217-226): forward + BN statistics, dgrad, wgrad"""
    from srl_zoo_b200 import ops
    g = torch.Generator().manual_seed(3)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    x = torch.randn(Bn, 64, big, big, generator=g)
    dy = torch.randn(Bn, 64, small, small, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, s, p)
    (ref * dy.double()).sum().backward()
    fpk, dpk = ops.pack_conv_w(w.cuda(), False)
    out = torch.empty(Bn, small, small, 64, device="cuda")
    _, stats = ops.conv64(nhwc(x).cuda(), fpk, out, (big, big), (small, small), 3, s, p, False, want_stats=True)
    assert H.rel_err(nchw(out), ref) < TOL
    assert H.rel_err(stats[:64], ref.sum((0, 2, 3))) < TOL and H.rel_err(stats[64:], (ref * ref).sum((0, 2, 3))) < TOL
    outd = torch.empty(Bn, big, big, 64, device="cuda")
    ops.conv64(nhwc(dy).cuda(), dpk, outd, (big, big), (small, small), 3, s, p, True)
    assert H.rel_err(nchw(outd), xr.grad) < TOL
    gw = ops.wgrad64(nhwc(x).cuda(), nhwc(dy).cuda(), (big, big), (small, small), 3, s, p)
    assert H.rel_err(gw, wr.grad) < TOL


@pytest.mark.parametrize("Bn,small", [(3, 6), (2, 13), (1, 55), (4, 1)])
def test_conv_transpose3x3_family(Bn, small):
    """ConvTranspose2d(64,64,3,stride=2) (models/models.py:66-78): forward (+bias, +BN/ReLU on load), dgrad, wgrad"""
    from srl_zoo_b200 import ops
    big = 2 * small + 1
    g = torch.Generator().manual_seed(4)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    x = torch.randn(Bn, 64, small, small, generator=g)
    bias = torch.randn(64, generator=g)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    dy = torch.randn(Bn, 64, big, big, generator=g)
    fpk, dpk = ops.pack_conv_w(w.cuda(), True)
    out = torch.empty(Bn, big, big, 64, device="cuda")
    ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), 2)
    ops.conv64(nhwc(x).cuda(), fpk, out, (big, big), (small, small), 3, 2, 0, True, bias=bias.cuda())
    assert H.rel_err(nchw(out), ref) < TOL
    act = F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double())
    wr = w.double().requires_grad_(True)
    refb = F.conv_transpose2d(act, wr, bias.double(), 2)
    ops.conv64(nhwc(x).cuda(), fpk, out, (big, big), (small, small), 3, 2, 0, True, bias=bias.cuda(), in_scale=sc.cuda(), in_shift=sh.cuda())
    assert H.rel_err(nchw(out), refb) < TOL
    (refb * dy.double()).sum().backward()
    gw = ops.wgrad64(nhwc(dy).cuda(), nhwc(x).cuda(), (big, big), (small, small), 3, 2, 0, dense_scale=sc.cuda(), dense_shift=sh.cuda())
    assert H.rel_err(gw, wr.grad) < TOL
    outd = torch.empty(Bn, small, small, 64, device="cuda")
    ops.conv64(nhwc(dy).cuda(), dpk, outd, (big, big), (small, small), 3, 2, 0, False)
    assert H.rel_err(nchw(outd), F.conv2d(dy.double(), w.double(), None, 2)) < TOL


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


@pytest.mark.parametrize("tconv,Bn,big,small,s,p", [(False, 2, 56, 56, 1, 1), (False, 7, 56, 56, 1, 1), (False, 3, 27, 14, 2, 1),
                                                    (True, 3, 13, 6, 2, 0), (True, 4, 27, 13, 2, 0), (True, 5, 55, 27, 2, 0),
                                                    (True, 3, 111, 55, 2, 0)])
def test_wgrad_tensor_core_kernels(tconv, Bn, big, small, s, p):
    """tcgen05 weight gradients of every 3x3 layer geometry (models/models.py:54,59,66-78): the halo-tile kernel
    (csrc/wgrad_halo_tc.cu) and the per-tap kernel (csrc/wgrad_tc.cu) against an fp64 torch reference."""
    from srl_zoo_b200 import ops
    from srl_zoo_b200._lib import lib
    g = torch.Generator().manual_seed(3)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    wr = w.double().clone().requires_grad_(True)
    if tconv:
        x = torch.randn(Bn, 64, small, small, generator=g)
        dy = torch.randn(Bn, 64, big, big, generator=g)
        sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
        act = F.relu(x.double() * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double())
        (F.conv_transpose2d(act, wr, None, s) * dy.double()).sum().backward()
        args, kw = (nhwc(dy).cuda(), nhwc(x).cuda(), (big, big), (small, small), 3, s, p), dict(dense_scale=sc.cuda(), dense_shift=sh.cuda())
    else:
        x = torch.randn(Bn, 64, big, big, generator=g)
        dy = torch.randn(Bn, 64, small, small, generator=g)
        (F.conv2d(x.double(), wr, None, s, p) * dy.double()).sum().backward()
        args, kw = (nhwc(x).cuda(), nhwc(dy).cuda(), (big, big), (small, small), 3, s, p), {}
    try:
        for mode in (1, 2):   # 1: halo-tile kernel where the geometry fits, 2: per-tap kernel
            lib.srlz_set_tensor_cores(mode)
            gw = ops.wgrad64(*args, tensor_cores=True, **kw)
            assert H.rel_err(gw, wr.grad) < TOL, mode
    finally:
        lib.srlz_set_tensor_cores(1)

"""CPU restatements of the index / operand mappings the tcgen05 row kernels and the pooling-backward kernel use
(srl_zoo_b200/csrc/enc0_rows_tc.cu, dec12_rows_tc.cu, bn_pool.cu), checked against torch's own convolution operators
(the oracle's arithmetic) and against the definition of MaxPool2d backward.  These are host-side models of the device
code's addressing: they pin the documented decompositions, not the CUDA implementation (the GPU parity tests do that)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import srl_oracle as O


# ------------------------------------------------------------------------------------------------ bf16 hi/lo split
def split_bf16(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def test_bf16x3_in_two_products():
    """A*W ~ A_hi*[W_hi | W_lo] + A_lo*W_hi (one N=2n product + one N=n product, column halves added): the identity behind
    the two-MMA form; the dropped terms (A_hi*W_lo is kept, A_lo*W_lo is not) leave a 2^-16 relative error per product."""
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 64, generator=g)
    W = torch.randn(64, 64, generator=g)
    Ah, Al = split_bf16(A)
    Wh, Wl = split_bf16(W)
    wide = Ah.double() @ torch.cat([Wh, Wl], 0).double().t()          # (128, 128): columns 0-63 hi*hi, 64-127 hi*lo
    acc = wide.clone()
    acc[:, :64] += Al.double() @ Wh.double().t()
    got = acc[:, :64] + acc[:, 64:]
    ref = A.double() @ W.double().t()
    three = Al.double() @ Wh.double().t() + Ah.double() @ Wl.double().t() + Ah.double() @ Wh.double().t()
    assert torch.equal(got, three)
    bound = (A.abs().double() @ W.abs().double().t()) * 2.0 ** -15
    assert ((got - ref).abs() <= bound).all()


# ------------------------------------------------------------------------------------------------ enc0 row images
PPI = 115  # pair images per input image (enc0_rows_tc.cu: er::PPI)


def enc0_pair_image(x, n, j, rect=None):
    """P_j[ox][k], k = c*16 + rr*8 + kx' : x[n, c, 2j-3+rr, 2ox-4+kx'] (zero outside the image / inside the DAE rectangle)."""
    img = x[n].clone()
    if rect is not None:
        h1, h2, w1, w2 = rect
        img[:, w1:w2, h1:h2] = 0.0
    P = torch.zeros(112, 64)
    for rr in range(2):
        r = 2 * j - 3 + rr
        if not 0 <= r < 224:
            continue
        for c in range(3):
            row = F.pad(img[c, r], (4, 8))        # index col + 4
            for ox in range(112):
                c0 = 2 * ox - 4
                P[ox, c * 16 + rr * 8:c * 16 + rr * 8 + 8] = row[c0 + 4:c0 + 12]
    return P


def enc0_pair_weights(w0):
    """Wp[p][co][k] = W0[co, c, 2p+rr, kx'-1] (zero for ky = 7, kx' = 0, k >= 48)  -- pack_enc0_rows_bf16_kernel"""
    Wp = torch.zeros(4, 64, 64)
    for p in range(4):
        for c in range(3):
            for rr in range(2):
                ky = 2 * p + rr
                if ky >= 7:
                    continue
                for kxp in range(1, 8):
                    Wp[p, :, c * 16 + rr * 8 + kxp] = w0[:, c, ky, kxp - 1]
    return Wp


def test_enc0_row_image_forward_decomposition():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 224, 224, generator=g)
    w0 = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    rect = (50, 130, 20, 100)
    ref = F.conv2d(O.apply_occlusion(x, np.array([rect], dtype=np.int32)), w0, stride=2, padding=3)[0]     # (64,112,112)
    Wp = enc0_pair_weights(w0)
    images = {}
    for oy in (0, 1, 55, 110, 111):
        acc = torch.zeros(112, 64, dtype=torch.float64)
        for p in range(4):
            j = oy + p
            if j not in images:
                images[j] = enc0_pair_image(x, 0, j, rect)
            acc += images[j].double() @ Wp[p].double().t()
        assert torch.allclose(acc.t().float(), ref[:, oy, :], atol=2e-4), oy


def test_enc0_row_image_wgrad_stacks_and_reduce():
    """Even / odd ring positions: (p0,p1)(p2,p3) -> accumulators 0,1 ; (-,p0)(p1,p2)(p3,-) -> accumulators 2,3,4; the reduce
    kernel's (accumulator, row) pairs must reassemble dW for every CTA range start parity."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 3, 224, 224, generator=g)
    w0 = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).requires_grad_()
    y = F.conv2d(x, w0, stride=2, padding=3)
    dy = torch.randn(y.shape, generator=g)
    rows = (3, 4, 5, 6, 7, 110, 111)                 # the output rows this "CTA" owns
    mask = torch.zeros_like(dy)
    mask[:, :, list(rows), :] = 1.0
    (ref,) = torch.autograd.grad(y, w0, dy * mask)
    for i0 in (3, 4):                                # range start parity decides which rows are "odd"
        acc = torch.zeros(5, 128, 64, dtype=torch.float64)
        g_lo = i0                                    # g0_of(i0) for image 0
        for oy in rows:
            if oy < i0:
                continue
            g0r = oy - g_lo
            d = dy[0, :, oy, :].t().double()         # (112 ox, 64 co)
            if g0r % 2 == 0:
                stacks = [(oy, 0), (oy + 2, 1)]
            else:
                stacks = [(oy - 1, 2), (oy + 1, 3), (oy + 3, 4)]
            for j, a in stacks:
                top = enc0_pair_image(x, 0, j).double()          # rows 0-63 of the stack
                bot = enc0_pair_image(x, 0, j + 1).double()      # rows 64-127 (LBO = one slot)
                acc[a, :64] += top.t() @ d
                acc[a, 64:] += bot.t() @ d
        got = torch.zeros(64, 3, 7, 7, dtype=torch.float64)
        for c in range(3):
            for ky in range(7):
                for kx in range(7):
                    p, k = ky >> 1, c * 16 + (ky & 1) * 8 + kx + 1
                    accA, rowA = (0 if p < 2 else 1), (64 + k if p & 1 else k)
                    accB = 2 if p == 0 else (4 if p == 3 else 3)
                    rowB = 64 + k if p in (0, 2) else k
                    got[:, c, ky, kx] = acc[accA, rowA] + acc[accB, rowB]
        sel = [r for r in rows if r >= i0]
        m2 = torch.zeros_like(dy)
        m2[:, :, sel, :] = 1.0
        (ref2,) = torch.autograd.grad(F.conv2d(x, w0, stride=2, padding=3), w0, dy * m2)
        assert torch.allclose(got.float(), ref2, atol=2e-3, rtol=1e-4), i0
    assert ref.shape == (64, 3, 7, 7)


def test_enc0_row_image_wgrad_quad_form():
    """QUAD form: [P_hi ; P_lo] (M = 128) x [dy_hi | dy_lo] (N = 128) per pair image: the epilogue adds the column halves,
    the reduce adds accumulator rows k and 64+k; the result is dW up to the bf16 split error (lo*lo included)."""
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 3, 224, 224, generator=g)
    w0 = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).requires_grad_()
    y = F.conv2d(x, w0, stride=2, padding=3)
    dy = torch.randn(y.shape, generator=g)
    rows = (0, 1, 2, 56, 111)
    mask = torch.zeros_like(dy)
    mask[:, :, list(rows), :] = 1.0
    (ref,) = torch.autograd.grad(y, w0, dy * mask)
    acc = torch.zeros(4, 128, 128, dtype=torch.float64)
    for oy in rows:
        dh, dl = split_bf16(dy[0, :, oy, :].t().contiguous())                    # (112, 64) each
        B = torch.cat([dh, dl], 1).double()                                      # N = 128
        for p in range(4):
            ph, pl = split_bf16(enc0_pair_image(x, 0, oy + p))
            A = torch.cat([ph, pl], 1).double()                                  # (112, 128): M = 128 as columns here
            acc[p] += A.t() @ B
    part = acc[:, :, :64] + acc[:, :, 64:]                                       # epilogue: column halves
    got = torch.zeros(64, 3, 7, 7, dtype=torch.float64)
    for c in range(3):
        for ky in range(7):
            for kx in range(7):
                p, k = ky >> 1, c * 16 + (ky & 1) * 8 + kx + 1
                got[:, c, ky, kx] = part[p, k] + part[p, 64 + k]                 # reduce: row halves
    assert torch.allclose(got.float(), ref, atol=2e-3, rtol=1e-4)


# ------------------------------------------------------------------------------------------------ dec12 rows
def test_dec12_row_ring_forward_columns_and_shifts():
    """out[2y+py, 2x+px, co] = b + sum_{dy,dx} a[y-dy, x-dx] . W[:, co, py+2dy, px+2dx]; accumulator column (py*2+px)*3 + co,
    row image = pixel p at image row p+1 (rows 0 and 112 zero), shift (dy,dx) reads image row x + 1 - dx of input row y - dy."""
    g = torch.Generator().manual_seed(3)
    a = torch.relu(torch.randn(1, 64, 111, 111, generator=g))
    w = torch.randn(64, 3, 4, 4, generator=g) * 0.1
    b = torch.randn(3, generator=g)
    ref = F.conv_transpose2d(a, w, b, stride=2)[0]                      # (3,224,224)
    # weight image of shift d = dy*2+dx: row j = (py*2+px)*3 + co, K = ci   (pack_dec12_fwd_bf16_kernel)
    Wimg = torch.zeros(4, 16, 64)
    for d in range(4):
        for j in range(12):
            pyx, co = j // 3, j % 3
            Wimg[d, j] = w[:, co, (pyx >> 1) + 2 * (d >> 1), (pyx & 1) + 2 * (d & 1)]

    def row_image(yy):
        img = torch.zeros(130, 64)
        if 0 <= yy < 111:
            img[1:112] = a[0, :, yy, :].t()
        return img

    for y in (0, 1, 57, 110, 111):
        acc = torch.zeros(128, 16, dtype=torch.float64)
        for grp in range(2):
            dy = 1 - grp
            img = row_image(y - dy).double()
            for dx in range(2):
                acc += img[1 - dx:1 - dx + 128] @ Wimg[dy * 2 + dx].double().t()
        for x in (0, 1, 60, 110, 111):
            for co in range(3):
                for py in range(2):
                    for px in range(2):
                        got = acc[x, (py * 2 + px) * 3 + co].item() + b[co].item()
                        assert abs(got - ref[co, 2 * y + py, 2 * x + px].item()) < 2e-4, (y, x, co, py, px)


def test_dec12_wgrad_gradient_columns_and_bias_rule():
    """G[x][k], k = co*16 + ky*4 + kx = g[co, 2y+ky, 2x+kx] written by column-pair tasks (m, co, ky): float2 g[.., 2m..2m+1]
    is kx = 0,1 of pixel m (row m+1) and kx = 2,3 of pixel m-1 (row m); the bias gradient counts an image row through its
    ky = 0,1 tasks, and the last two rows through the ky = 2,3 tasks of y = 110."""
    g = torch.Generator().manual_seed(4)
    a = torch.relu(torch.randn(1, 64, 111, 111, generator=g))
    w = (torch.randn(64, 3, 4, 4, generator=g) * 0.1).requires_grad_()
    b = torch.zeros(3, requires_grad=True)
    out = F.conv_transpose2d(a, w, b, stride=2)
    gout = torch.randn(out.shape, generator=g)
    ref_w, ref_b = torch.autograd.grad(out, (w, b), gout)
    acc = torch.zeros(64, 64, dtype=torch.float64)       # [k][ci]
    bsum = torch.zeros(3, dtype=torch.float64)
    for y in range(111):
        G = torch.zeros(128, 64, dtype=torch.float64)
        for task in range(112 * 12):
            m, cky = task % 112, task // 112
            co, ky = cky >> 2, cky & 3
            v = gout[0, co, 2 * y + ky, 2 * m:2 * m + 2].double()
            if ky < 2 or y == 110:
                bsum[co] += v.sum()
            if m < 111:
                G[m + 1, cky * 4 + 0:cky * 4 + 2] = v
            if m >= 1:
                G[m, cky * 4 + 2:cky * 4 + 4] = v
        A = torch.zeros(128, 64, dtype=torch.float64)
        A[1:112] = a[0, :, y, :].t().double()
        acc += G[:112].t() @ A[:112]
    got = acc[:48].t().reshape(64, 3, 4, 4)              # grad[ci*48 + k]
    assert torch.allclose(got.float(), ref_w, atol=2e-3, rtol=1e-4)
    assert torch.allclose(bsum.float(), ref_b, atol=2e-2, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ pooling backward
def _pool_definition(H, W, PH, PW, pad):
    s = set()
    for ph in range(PH):
        for pw in range(PW):
            for ky in range(3):
                for kx in range(3):
                    h, w = 2 * ph - pad + ky, 2 * pw - pad + kx
                    if 0 <= h < H and 0 <= w < W:
                        s.add((h, w, ph, pw, ky * 3 + kx))
    return s


def _pool_kernel_visits(H, W, PH, PW, pad):
    """loop structure of pool_bwd_bn_apply_kernel: rows with their (<= 2) window rows, columns as (m, s) pairs"""
    out, stores = [], []
    nq = (W + pad) // 2 + 1
    for h in range(H):
        th = h + pad - 2
        ph_a = 0 if th <= 0 else (th + 1) >> 1
        ph_b = min((h + pad) >> 1, PH - 1)
        for qi in range(nq):
            q = qi - 1
            wm = 2 * q - pad + 1
            ws = wm + 1
            m_ok, s_ok = 0 <= wm < W, 0 <= ws < W
            qa_ok, qb_ok = 0 <= q < PW, q + 1 < PW
            for i in range(2):
                ph = ph_a + i
                ky3 = (h - (ph * 2 - pad)) * 3
                if ph <= ph_b and qa_ok:
                    if m_ok:
                        out.append((h, wm, ph, q, ky3 + 1))
                    if s_ok:
                        out.append((h, ws, ph, q, ky3 + 2))
                if ph <= ph_b and qb_ok and s_ok:
                    out.append((h, ws, ph, q + 1, ky3))
            if m_ok:
                stores.append((h, wm))
            if s_ok:
                stores.append((h, ws))
    return out, stores


@pytest.mark.parametrize("geo", [(112, 112, 56, 56, 1), (56, 56, 27, 27, 0), (14, 14, 6, 6, 0), (9, 9, 4, 4, 0), (8, 8, 4, 4, 1)])
def test_pool_backward_column_pairs_visit_every_window_tap_once(geo):
    H, W, PH, PW, pad = geo
    visits, stores = _pool_kernel_visits(*geo)
    for h, w, ph, pw, tap in visits:
        assert 0 <= tap < 9 and (2 * ph - pad + tap // 3, 2 * pw - pad + tap % 3) == (h, w)
    assert len(visits) == len(set(visits))
    assert set(visits) == _pool_definition(*geo)
    assert sorted(stores) == sorted((h, w) for h in range(H) for w in range(W))


def test_pooled_side_bn_backward_sums():
    """sum dz and sum dz*xhat taken over the pooled tensor (pool_bwd_stats_kernel): m = (a > 0), xhat = (a - beta)/gamma."""
    g = torch.Generator().manual_seed(5)
    y = torch.randn(2, 8, 14, 14, generator=g)
    gamma = 0.5 + torch.rand(8, generator=g)
    beta = 0.2 * torch.randn(8, generator=g)
    mean, var = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    xhat = (y - mean[None, :, None, None]) * invstd[None, :, None, None]
    z = (xhat * gamma[None, :, None, None] + beta[None, :, None, None]).requires_grad_()
    a = F.max_pool2d(torch.relu(z), 3, 2)
    dpool = torch.randn(a.shape, generator=g)
    (dz,) = torch.autograd.grad(a, z, dpool)
    ref1, ref2 = dz.sum((0, 2, 3)), (dz * xhat).sum((0, 2, 3))
    m = (a > 0).float()
    xh = (a.detach() - beta[None, :, None, None]) / gamma[None, :, None, None]
    got1, got2 = (m * dpool).sum((0, 2, 3)), (m * dpool * xh).sum((0, 2, 3))
    assert torch.allclose(got1, ref1, atol=1e-4)
    assert torch.allclose(got2, ref2, atol=1e-4)


def test_stride2_dgrad_row_kernel_dataflow():
    """Host model of csrc/dgrad_s2_rows_tc.cu (dgrad of ConvTranspose2d(64,64,3,2), models/models.py:66-78): a CTA's output-row range
    is cut into image segments; input rows 2sa..2sb+2 stream through the ring as two column-parity sub-images; tap kx reads parity
    kx&1 at pixel offset kx>>1; row 2s is ky=0 of output row s and ky=2 of output row s-1, row 2s+1 is ky=1 of output row s; at
    most two accumulators are open and they are opened / closed in output-row order (buffer = count & 3).  Against F.conv2d."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    for (B, SH, SW, n_cta) in ((3, 5, 5, 4), (2, 6, 13, 5), (1, 3, 2, 3)):
        BH, BW, C = 2 * SH + 1, 2 * SW + 1, 4
        dy = torch.randn(B, BH, BW, C, generator=g, dtype=torch.float64)          # NHWC
        W = torch.randn(C, C, 3, 3, generator=g, dtype=torch.float64)             # ConvTranspose2d weight [ci][co][ky][kx]
        ref = F.conv2d(dy.permute(0, 3, 1, 2), W, None, 2).permute(0, 2, 3, 1)    # (B, SH, SW, ci)
        wd = W.permute(2, 3, 1, 0).reshape(9, C, C)                               # dgrad pack [tap][co][ci]
        out = torch.full((B * SH, SW, C), float("nan"), dtype=torch.float64)
        total = B * SH
        for cta in range(n_cta):
            i0, i1 = total * cta // n_cta, total * (cta + 1) // n_cta
            acc, cnt, open_rows, closed = {}, 0, [], []
            i = i0
            while i < i1:
                n, sa = divmod(i, SH)
                sb = min(SH - 1, sa + (i1 - i) - 1)
                for r in range(2 * sa, 2 * sb + 3):
                    row = dy[n, r]                                                # (BW, C)
                    sub = {0: torch.zeros(64, C, dtype=torch.float64), 1: torch.zeros(64, C, dtype=torch.float64)}
                    for x in range(BW):
                        sub[x & 1][x >> 1] = row[x]                               # producer: pixel x -> row x>>1 of parity x&1

                    def issue(ky, buf):
                        for kx in range(3):
                            a = sub[kx & 1][(kx >> 1):(kx >> 1) + SW]             # descriptor shifted by kx>>1 rows; lanes < SW are valid
                            acc[buf] = acc[buf] + a @ wd[ky * 3 + kx]
                    s = r >> 1
                    if r % 2 == 0:
                        if s - 1 >= sa:
                            buf = (cnt - 1) & 3
                            issue(2, buf)
                            closed.append(((n * SH + s - 1), buf))
                            out[n * SH + s - 1] = acc[buf]
                            open_rows.remove(buf)
                        if s <= sb:
                            buf = cnt & 3
                            assert buf not in open_rows
                            acc[buf] = torch.zeros(SW, C, dtype=torch.float64)
                            open_rows.append(buf)
                            issue(0, buf)
                            cnt += 1
                    else:
                        issue(1, (cnt - 1) & 3)
                    assert len(open_rows) <= 2
                i += sb - sa + 1
            assert [c[0] for c in closed] == list(range(i0, i1))                  # the epilogue's order: it-th finished row = i0 + it
            assert [c[1] for c in closed] == [k & 3 for k in range(i1 - i0)]
        assert torch.allclose(out.reshape(B, SH, SW, C), ref, atol=1e-12)


# ------------------------------------------------------------------------------------------------ halo wgrad (wgrad_halo_tc.cu)
def _wh_plan(BH, SH, stride, pad, PT=480, TMEM=512):
    """Host model of make_wh_plan(): parity-class images of the big side, tap pairs stacked on M through a row shift, and the
    MMA form of every op (0: pair, three N=64 MMAs, 64 columns | 1: pair, N=128 + N=64, 128 columns | 2: single tap with
    [A_hi ; A_lo] stacked on M, 64 columns, its lo rows = partial image 9)."""
    s = stride
    oy = [(k - pad) // s for k in range(3)]                 # floor division
    cy = [(k - pad) - oy[k] * s for k in range(3)]
    maxrange = max(max(oy[k] for k in range(3) if cy[k] == c) - min(oy[k] for k in range(3) if cy[k] == c)
                   for c in range(s) if any(cy[k] == c for k in range(3)))
    HW = SH + maxrange
    R = min(128 // HW, SH)
    classes, ops = [], []
    for c_y in range(s):
        for c_x in range(s):
            kys = [k for k in range(3) if cy[k] == c_y]
            kxs = [k for k in range(3) if cy[k] == c_x]
            if not kys or not kxs:
                continue
            mny, mxy, mnx = min(oy[k] for k in kys), max(oy[k] for k in kys), min(oy[k] for k in kxs)
            taps = sorted(((oy[ky] - mny) * HW + (oy[kx] - mnx), ky * 3 + kx) for ky in kys for kx in kxs)
            classes.append(dict(by0=mny * s + c_y, bx0=mnx * s + c_x, nrows=R + (mxy - mny), taps=taps))
            for i in range(0, len(taps), 2):
                pair = taps[i:i + 2]
                ops.append(dict(cls=len(classes) - 1, taps=[t for _, t in pair], form=0))
    extra = None
    for op in ops:
        if len(op["taps"]) == 1 and extra is None:
            op["form"], extra = 2, op["taps"][0]
    spare = TMEM - 64 * len(ops)
    for op in ops:
        if op["form"] == 0 and len(op["taps"]) == 2 and spare >= 64:
            op["form"], spare = 1, spare - 64
    maxpx = max([R * HW] + [c["nrows"] * HW for c in classes])
    return dict(HW=HW, R=R, classes=classes, ops=ops, extra=extra, rounds=-(-maxpx * 8 // PT), PT=PT,
                cols=sum(128 if op["form"] == 1 else 64 for op in ops))


WGRAD_LAYERS = [  # (big, small, stride, pad): encoder_conv.4 / .8, decoder_conv.0 / .3 / .6 / .9 (models/models.py:54,59,66-78)
    (56, 56, 1, 1), (27, 14, 2, 1), (13, 6, 2, 0), (27, 13, 2, 0), (55, 27, 2, 0), (111, 55, 2, 0)]


@pytest.mark.parametrize("geo", WGRAD_LAYERS)
def test_halo_wgrad_plan_covers_every_tap_once_within_tmem(geo):
    big, small, s, pad = geo
    p = _wh_plan(big, small, s, pad)
    taps = [t for op in p["ops"] for t in op["taps"]]
    assert sorted(taps) == list(range(9))                       # nine taps, each in exactly one accumulator half
    assert len(p["ops"]) <= 5 and p["cols"] <= 512              # TMEM columns
    assert sum(op["form"] == 2 for op in p["ops"]) <= 1 and p["rounds"] <= 4
    assert p["R"] * p["HW"] <= 128                               # one dense tile = at most 128 K rows


@pytest.mark.parametrize("geo", WGRAD_LAYERS)
def test_halo_wgrad_class_images_reproduce_the_weight_gradient(geo):
    """P[tap][cg][cd] = sum over small pixels of big[.. sy*s - pad + ky, sx*s - pad + kx, cg] * small[sy, sx, cd], evaluated
    the way the kernel does -- tile by tile, every tap as a row shift of its parity-class image on the pitch HW, columns
    beyond the tensors zero -- equals torch's weight gradient of the layer."""
    big, small, s, pad = geo
    p = _wh_plan(big, small, s, pad)
    HW, R = p["HW"], p["R"]
    g = torch.Generator().manual_seed(3)
    C = 4                                                        # channels are independent: a small C keeps the model cheap
    xb = torch.randn(big, big, C, generator=g, dtype=torch.float64)
    dy = torch.randn(small, small, C, generator=g, dtype=torch.float64)
    P = torch.zeros(9, C, C, dtype=torch.float64)
    for sy0 in range(0, small, R):
        dense = torch.zeros(R * HW + 2 * HW + 4, C, dtype=torch.float64)       # (rows beyond R*HW stay zero)
        for r in range(R):
            for x in range(small):
                if sy0 + r < small:
                    dense[r * HW + x] = dy[sy0 + r, x]
        for cl in p["classes"]:
            img = torch.zeros(cl["nrows"] * HW + 2 * HW + 4, C, dtype=torch.float64)
            for i in range(cl["nrows"]):
                for j in range(HW):
                    by, bx = (sy0 + i) * s + cl["by0"], j * s + cl["bx0"]
                    if 0 <= by < big and 0 <= bx < big:
                        img[i * HW + j] = xb[by, bx]
            for shift, tap in cl["taps"]:
                K = R * HW
                P[tap] += img[shift:shift + K].t() @ dense[:K]
    # torch: conv2d(x (1,C,big,big), w (C,C,3,3), stride s, padding pad) -> grad of w[cd, cg, ky, kx]
    x4 = xb.permute(2, 0, 1)[None].clone().requires_grad_(False)
    w = torch.zeros(C, C, 3, 3, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(x4, w, stride=s, padding=pad)
    assert out.shape[-1] >= small
    out[..., :small, :small].backward(dy.permute(2, 0, 1)[None])
    ref = w.grad                                                 # [cd, cg, ky, kx]
    got = P.reshape(3, 3, C, C).permute(3, 2, 0, 1)              # P[tap][cg][cd] -> [cd, cg, ky, kx]
    assert torch.allclose(got, ref, atol=1e-9)


@pytest.mark.parametrize("geo", WGRAD_LAYERS)
def test_halo_wgrad_chunk_items_cover_every_unit_once(geo):
    """Producer mapping: thread i owns 16-byte chunk i & 7 of pixels (i >> 3) + k * PT/8, k < rounds, of every unit (dense tile,
    class images): every (pixel, chunk) of every unit is written by exactly one (thread, round)."""
    big, small, s, pad = geo
    p = _wh_plan(big, small, s, pad)
    PT, KR = p["PT"], p["rounds"]
    for npx in [p["R"] * p["HW"]] + [c["nrows"] * p["HW"] for c in p["classes"]]:
        seen = np.zeros((npx, 8), dtype=np.int32)
        for i in range(PT):
            for k in range(KR):
                q = (i >> 3) + k * (PT // 8)
                if q < npx:
                    seen[q, i & 7] += 1
        assert (seen == 1).all()


# ------------------------------------------------------------------------------------------------ seven-warp producer mappings
def test_seven_warp_producer_item_coverage():
    """The kernels that gave their idle warps to the epilogue stage with 224 producer threads (warps 8-14):
    * dec12 dgrad (conv_tc.cu, MODE 2): 256 (pixel, K-half) items -- thread i item i, warp 14 also items 224 + (i & 31); a half-1
      item writes K chunks 4, 5 only (chunks 6, 7 = K slots 48..63 are the zero padding written once);
    * stride-2 dgrad row kernel (dgrad_s2_rows_tc.cu): chunk jc = i & 7 of pixels (i >> 3) + 28 q, q < 4, of a row of <= 112 pixels."""
    seen = np.zeros((128, 8), dtype=np.int32)               # (pixel of the tile, 16-byte K chunk)
    for i in range(224):
        pix, half = i & 127, i >> 7
        for j in range(4 if half == 0 else 2):
            seen[pix, half * 4 + j] += 1
        if i >> 5 == 6:                                     # warp 14: the 32 items left over, all of half 1
            for j in range(2):
                seen[96 + (i & 31), 4 + j] += 1
    assert (seen[:, :6] == 1).all() and (seen[:, 6:] == 0).all()
    for BW in (111, 55, 27, 13):
        cover = np.zeros((BW, 8), dtype=np.int32)
        for i in range(224):
            for q in range(4):
                x = (i >> 3) + 28 * q
                if x < BW:
                    cover[x, i & 7] += 1
        assert (cover == 1).all()

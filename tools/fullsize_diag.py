"""Localises forward errors at BASELINE config-2 size: decoded vs the oracle's modules on the same GPU (cuDNN, TF32 off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import helpers as H
import srl_zoo_b200
from oracle import srl_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for rep in range(3):
    mod, P, B = H.make_pair("ae", ["autoencoder"])
    g = torch.Generator().manual_seed(11)
    obs = torch.randn(bs, 3, 224, 224, generator=g).cuda()
    nobs = torch.randn(bs, 3, 224, 224, generator=g).cuda()
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=0.005)
    Pg = {k: v.detach().cuda() for k, v in P.items()}
    Bg = {k: v.cuda() for k, v in B.items()}
    with torch.no_grad():
        ref_states, ref_dec = O.ae_forward(Pg, Bg, obs, True)
    eng.step(obs, nobs)
    torch.cuda.synchronize()
    d = (eng.decoded[0] - ref_dec).abs()
    scale = ref_dec.abs().max().item()
    per_img = d.amax(dim=(1, 2, 3)) / scale
    per_row = d.amax(dim=(0, 1, 3)) / scale
    bad = (per_img > 1e-4).nonzero().flatten().tolist()
    print("rep %d: states rel %.2e decoded rel %.2e  images over 1e-4: %d %s  worst rows %s" % (
        rep, H.norm_rel(eng.lat[0], ref_states), (d.max() / scale).item(), len(bad), bad[:8],
        torch.topk(per_row, 4).indices.tolist()), flush=True)
    if bad:
        i = bad[0]
        di = d[i].amax(dim=0)
        ys, xs = (di > 1e-4 * scale).nonzero(as_tuple=True)
        print("   image %d: %d bad pixels, y range %d..%d, x range %d..%d" % (i, len(ys), ys.min().item(), ys.max().item(), xs.min().item(), xs.max().item()))

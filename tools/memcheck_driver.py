"""One small pass over every product entry point, meant to run under compute-sanitizer (bs = 2, so the whole pass stays
within a couple of minutes under the tool):

    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/memcheck_driver.py
    PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool initcheck --error-exitcode 3 python tools/memcheck_driver.py

Covers: the fused train step for AE / DAE (rectangles) / VAE + forward + inverse, the pinned uint8 hand-over (step_host), the
drop-in module through autograd (training kernels with an mlp inverse head and the reward head), the folded eval-mode encoder
and the decoder-only call.  No oracle here: the parity tests are in tests/; this is only about addresses.
SRLZ_LIB=<path> runs a build variant; SRLZ_MEMCHECK_KEEP_GOING=1 reports a failing section and carries on (survey mode)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import srl_zoo_b200  # noqa: E402
from srl_zoo_b200 import _lib  # noqa: E402
from srl_zoo_b200.occlusion import sample_rects  # noqa: E402

if os.environ.get("SRLZ_LIB"):   # before the first use of the lazy handle
    _lib.LIB_PATH = os.path.abspath(os.environ["SRLZ_LIB"])
KEEP_GOING = os.environ.get("SRLZ_MEMCHECK_KEEP_GOING") == "1"

S, A = 200, 6
bs = int(os.environ.get("SRLZ_MEMCHECK_BS", "2"))
dev = "cuda:0"
torch.manual_seed(0)
g = torch.Generator().manual_seed(3)
obs, nobs = [torch.randn(bs, 3, 224, 224, generator=g).to(dev) for _ in range(2)]
act = torch.randint(0, A, (bs, 1), generator=g).to(dev)
eps = [torch.randn(bs, S, generator=g).to(dev) for _ in range(2)]
rng = np.random.RandomState(1)
rects = [torch.from_numpy(sample_rects(bs, rng=rng)).to(dev) for _ in range(2)]


def fused(losses):
    mod = srl_zoo_b200.B200SRLModules(S, A, True, "custom_cnn", losses).to(dev)
    eng = srl_zoo_b200.TrainStep(mod, bs, lr=1e-3)
    for _ in range(2):
        if "dae" in losses:
            t = eng.step(obs, nobs, act, eps[0], eps[1], rects[0], rects[1])
        else:
            t = eng.step(obs, nobs, act, eps[0], eps[1])
    torch.cuda.synchronize()
    print("step", losses, [round(float(v), 5) for v in t.cpu()])
    f0, f1 = [torch.from_numpy(rng.randint(0, 256, (bs, 224, 224, 3)).astype(np.uint8)).pin_memory() for _ in range(2)]
    if "dae" not in losses:
        t = eng.step_host(f0, f1, act.cpu().pin_memory())
        torch.cuda.synchronize()
        print("step_host", losses, [round(float(v), 5) for v in t.cpu()])
    mod.eval()
    with torch.no_grad():
        s = mod.getStates(obs)
        s2 = eng.predict_states(nobs)
        d = mod.model.decode(s)
    torch.cuda.synchronize()
    print("eval", losses, float(s.abs().mean()), float(s2.abs().mean()), float(d.abs().mean()))


def dropin():
    """drop-in module through autograd (the install()ed learner body's kernels), mlp inverse head + reward head"""
    mod = srl_zoo_b200.B200SRLModules(S, A, True, "custom_cnn", ["autoencoder", "inverse", "reward", "forward"], "mlp").to(dev)
    opt = torch.optim.Adam(mod.parameters(), lr=1e-3)
    (st, dec), (nst, ndec) = mod(obs), mod(nobs)
    loss = ((dec - obs) ** 2).mean() + ((ndec - nobs) ** 2).mean()
    loss = loss + ((mod.forwardModel(st, act) - nst) ** 2).mean()
    loss = loss + torch.nn.functional.cross_entropy(mod.inverseModel(st, nst), act.squeeze(1))
    loss = loss + mod.rewardModel(st, nst).square().mean()
    opt.zero_grad()
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    print("dropin autograd", float(loss.detach()))


for name, fn in [("ae", lambda: fused(["autoencoder"])), ("dae", lambda: fused(["dae"])),
                 ("vae_fwd_inv", lambda: fused(["vae", "forward", "inverse"])), ("dropin", dropin)]:
    try:
        fn()
    except RuntimeError as e:
        if not KEEP_GOING:
            raise
        print("SECTION FAILED", name, str(e)[:200])
print("MEMCHECK_DRIVER_DONE launches", srl_zoo_b200.lib.srlz_launch_count())

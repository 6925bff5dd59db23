"""Oracle (oracle/srl_oracle.py) against the committed fixtures that oracle/make_golden.py produced from
the LIVE reference (models/learner.py:373-497 replayed with the reference's own modules).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import srl_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"ae": ("ae", False, False), "dae": ("dae", False, False), "vae": ("vae", False, False),
         "ae_fwd_inv": ("ae", True, True)}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    kind, use_fwd, use_inv = CASES[name]
    fx = np.load(os.path.join(GOLD, "step_%s.npz" % name))
    if str(fx["meta_torch"]) != torch.__version__:
        pytest.skip("fixtures generated with torch %s" % fx["meta_torch"])
    bs, S, A = int(fx["meta_bs"]), int(fx["meta_state_dim"]), int(fx["meta_action_dim"])
    obs, nobs, actions = O.synthetic_batch(bs, seed=int(fx["meta_input_seed"]))
    assert rel([obs.double().sum().item(), nobs.double().sum().item()], fx["obs_checksum"]) < 1e-12
    assert np.array_equal(actions.numpy(), fx["actions"])
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=int(fx["meta_seed"]))
    for k, v in sd.items():  # same init as the reference (RNG order, models/modules.py:37-49)
        assert rel([v.double().sum().item(), v.double().abs().sum().item()], fx["w0sum/" + k]) < 1e-12, k
    P, B = O.split_state(sd)
    with torch.no_grad():
        ev = O.get_states("vae" if kind == "vae" else "ae", P, {k: v.clone() for k, v in B.items()}, obs, False)
    assert rel(ev.numpy(), fx["eval_states"]) < 1e-6
    opt = O.Adam(P, lr=0.005)
    r = O.train_step(kind, P, B, obs, nobs, actions, torch.from_numpy(fx["eps"]), torch.from_numpy(fx["next_eps"]),
                     fx["rects"], fx["next_rects"], use_forward=use_fwd, use_inverse=use_inv, optimizer=opt)
    for n, v in r["losses"].items():
        assert rel(v, fx["loss/" + n]) < 1e-6, n
    assert rel(r["total"], fx["total"]) < 1e-6
    assert rel(r["states"].numpy(), fx["states"]) < 1e-6
    assert rel(r["next_states"].numpy(), fx["next_states"]) < 1e-6
    assert rel(r["decoded"][:, :, ::8, ::8].numpy(), fx["decoded_sub"]) < 1e-6
    for k, p in P.items():
        if p.grad is None:
            assert "gsum/" + k not in fx.files, k
            continue
        assert abs(p.grad.double().norm().item() - fx["gsum/" + k][1]) <= 1e-5 * fx["gsum/" + k][1] + 1e-12, k
        if "g/" + k in fx.files:
            assert rel(p.grad.numpy(), fx["g/" + k]) < 1e-5, k
    for k in sd:
        mine = B[k] if O.is_buffer(k) else P[k].detach()
        if "w1/" + k in fx.files:
            assert rel(mine.numpy(), fx["w1/" + k]) < 1e-5, k


def test_occlusion_semantics():
    """preprocessing/data_loader.py:55-63 + :255 -- zero block is tensor[:, w1:w2, h1:h2], all channels,
    value 0.0 in normalised space; empty rectangles allowed."""
    x = torch.ones(2, 3, O.IMG, O.IMG)
    rects = np.array([[10, 20, 30, 50], [5, 5, 0, 224]], dtype=np.int32)
    y = O.apply_occlusion(x, rects)
    assert y[0, :, 30:50, 10:20].abs().sum() == 0
    assert y[0].sum() == 3 * (224 * 224 - 20 * 10)
    assert torch.equal(y[1], x[1])  # empty rectangle
    r = O.sample_rects(64, rng=np.random.RandomState(0))
    assert (r[:, 0] <= r[:, 1]).all() and (r[:, 2] <= r[:, 3]).all() and r.min() >= 0 and r.max() <= 224
    assert ((r[:, 1] - r[:, 0]) <= 112).all() and ((r[:, 3] - r[:, 2]) <= 112).all()


def test_preprocess_u8_matches_reference_fixture():
    """SURVEY.md 8f N1 parity target: uint8 RGB HWC -> normalised (1,3,W,H) fp32, bit-exact against the vectors recorded from the
    reference's preprocessInput + loader transpose (oracle/make_golden_preprocess.py)."""
    import numpy as np
    fx = np.load(os.path.join(GOLD, "preprocess_u8.npz"))
    plain = O.preprocess_u8(fx["image"])
    masked = O.preprocess_u8(fx["image"], fx["rect"])
    assert tuple(plain.shape) == (1, 3, 10, 12)                       # (1, C, W, H): the loader swaps H and W
    assert np.array_equal(plain.numpy(), fx["plain"]) and np.array_equal(masked.numpy(), fx["masked"])
    h1, h2, w1, w2 = [int(v) for v in fx["rect"]]
    assert float(masked[0, :, w1:w2, h1:h2].abs().max()) == 0.0       # the block apply_occlusion zeroes after the transpose
    assert np.array_equal(O.apply_occlusion(plain, fx["rect"][None]).numpy(), fx["masked"])


def test_preprocess_u8_truth_table_pins_full_size_frames():
    """the normalisation is elementwise in (byte value, channel): the reference's 256 x 3 truth table (recorded by
    oracle/make_golden_preprocess.py from preprocessInput) determines a full 224 x 224 frame; the oracle reproduces it bit-exactly"""
    import numpy as np
    fx = np.load(os.path.join(GOLD, "preprocess_u8.npz"))
    im = np.random.RandomState(3).randint(0, 256, (224, 224, 3)).astype(np.uint8)
    want = fx["lut"][im, np.arange(3)[None, None, :]].transpose(2, 1, 0)[None]      # (1, C, W, H): data_loader.py:255
    assert np.array_equal(O.preprocess_u8(im).numpy(), want)


@pytest.mark.parametrize("name", ["ae_mlp_reward", "ae_split", "vae_split_mlp_reward"])
def test_oracle_matches_reference_golden_cheap_heads_and_split(name):
    """SURVEY.md 8a A9 (mlp inverse head, models/forward_inverse.py:50-56) and 8f N4 (reward head :78-95, SRLModulesSplit
    models/modules.py:103-288): the oracle against fixtures recorded from the reference's own modules + loss functions
    (oracle/make_golden_heads.py), so that these rows are pinned on the GPU box too, not only in the build container."""
    from collections import OrderedDict
    fx = np.load(os.path.join(GOLD, "heads_%s.npz" % name))
    if str(fx["meta_torch"]) != torch.__version__:
        pytest.skip("fixtures generated with torch %s" % fx["meta_torch"])
    kind, losses, inv_type = str(fx["meta_kind"]), str(fx["meta_losses"]).split(","), str(fx["meta_inverse_model_type"])
    split = None
    if str(fx["meta_split"]):
        split = OrderedDict((k, int(v)) for k, v in (kv.split(":") for kv in str(fx["meta_split"]).split(",")))
    S, A, bs = 200, 6, 2
    obs, nobs, actions = O.synthetic_batch(bs, seed=1234)
    assert rel([obs.double().sum().item(), nobs.double().sum().item()], fx["obs_checksum"]) < 1e-12
    assert np.array_equal(actions.numpy(), fx["actions"])
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=1, inverse_model_type=inv_type)
    assert set("w0sum/" + k for k in sd) == set(f for f in fx.files if f.startswith("w0sum/"))
    for k, v in sd.items():
        assert rel([v.double().sum().item(), v.double().abs().sum().item()], fx["w0sum/" + k]) < 1e-12, k
    P, B = O.split_state(sd)
    r = O.train_step(kind, P, B, obs, nobs, actions, torch.from_numpy(fx["eps"]), torch.from_numpy(fx["next_eps"]),
                     use_forward="forward" in losses, use_inverse="inverse" in losses, use_reward="reward" in losses,
                     rewards=torch.from_numpy(fx["rewards"]), split_dimensions=split)
    assert set("loss/" + n for n in r["losses"]) == set(f for f in fx.files if f.startswith("loss/"))
    for n, v in r["losses"].items():
        assert rel(v, fx["loss/" + n]) < 1e-6, n
    assert rel(r["states"].numpy(), fx["states"]) < 1e-6
    assert rel(r["decoded"][:, :, ::8, ::8].numpy(), fx["decoded_sub"]) < 1e-6
    assert rel([r["decoded"].double().sum().item(), r["decoded"].double().pow(2).sum().item()], fx["decoded_checksum"]) < 1e-6
    nograd = set(k for k in str(fx["nograd"]).split(",") if k)
    for k, p in P.items():
        if k in nograd:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        assert abs(p.grad.double().norm().item() - fx["gsum/" + k][1]) <= 1e-5 * fx["gsum/" + k][1] + 1e-12, k
        if "g/" + k in fx.files:
            assert rel(p.grad.numpy(), fx["g/" + k]) < 2e-5, k

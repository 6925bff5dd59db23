"""CPU tests of the drop-in boundary's host logic (srl_zoo_b200/install.py): which configurations are routed to
the B200 module, which fall through to the reference class unchanged, and how the loss functions are routed.  The second
half runs BASELINE.json configs[0] (autoencoder loss, mlp, 64x64x3, bs=32, CPU) through the installed names with the LIVE
reference when /root/reference is present (build container only)."""
import os
import sys
import types

import pytest
import torch

import srl_zoo_b200
from srl_zoo_b200 import losses as b200_losses

inst = sys.modules["srl_zoo_b200.install"]   # the module (the package attribute `install` is the function)

LOSS_NAMES = ("autoEncoderLoss", "generationLoss", "kullbackLeiblerLoss", "forwardModelLoss", "inverseModelLoss")


class FakeRefModules:
    def __init__(self, **kw):
        self.kw = kw


def fake_learner():
    calls = []
    ns = types.SimpleNamespace(SRLModules=FakeRefModules, LossManager=object, calls=calls)
    for name in LOSS_NAMES:
        setattr(ns, name, (lambda n: (lambda *a, **k: calls.append((n, a, k)) or "ref:" + n))(name))
    return ns


def test_dispatch_routes_only_hot_path_configurations():
    ns, mods = fake_learner(), types.SimpleNamespace()
    replaced = srl_zoo_b200.install(ns, mods)
    assert replaced["SRLModules"] is FakeRefModules and mods.B200SRLModules is srl_zoo_b200.B200SRLModules
    hot = ns.SRLModules(200, 6, True, "custom_cnn", ["autoencoder"])
    assert isinstance(hot, srl_zoo_b200.B200SRLModules)
    assert isinstance(ns.SRLModules(state_dim=200, action_dim=6, cuda=True, model_type="custom_cnn", losses=["vae", "forward", "inverse"]),
                      srl_zoo_b200.B200SRLModules)
    # everything else goes to the reference class with the reference's own keyword arguments
    # the cheap heads ride along: mlp inverse head, reward head
    for kw in (dict(losses=["autoencoder", "inverse"], inverse_model_type="mlp"), dict(losses=["dae", "reward"])):
        assert isinstance(ns.SRLModules(state_dim=200, cuda=True, model_type="custom_cnn", **kw), srl_zoo_b200.B200SRLModules), kw
    cold = [dict(model_type="mlp", losses=["autoencoder"], cuda=True, state_dim=200),                 # configs[0]: the CPU plumbing config
            dict(model_type="custom_cnn", losses=["autoencoder"], cuda=False, state_dim=200),         # no CPU path in libsrlz
            dict(model_type="custom_cnn", losses=["inverse", "forward"], cuda=True, state_dim=200),   # no autoencoder family loss
            dict(model_type="custom_cnn", losses=["autoencoder", "triplet"], cuda=True, state_dim=200),
            dict(model_type="custom_cnn", losses=["autoencoder", "priors"], cuda=True, state_dim=200),
            dict(model_type="custom_cnn", losses=["autoencoder"], cuda=True, state_dim=3),
            dict(model_type="resnet", losses=["autoencoder"], cuda=True, state_dim=200),
            dict(model_type="custom_cnn", losses=None, cuda=True, state_dim=200)]
    for kw in cold:
        m = ns.SRLModules(**kw)
        assert isinstance(m, FakeRefModules), kw
        for k, v in kw.items():
            assert m.kw[k] == v
        assert set(m.kw) == {"state_dim", "action_dim", "cuda", "model_type", "losses", "inverse_model_type"}


def test_perceptual_configs_and_their_denoiser_stay_on_the_reference():
    """the perceptual loss differentiates a frozen DAE w.r.t. its INPUT (learner.py:404-412), which the B200 module
    does not compute: the VAE and the denoiser constructed after it (learner.py:319, losses=["dae"]) must both be reference modules"""
    ns = fake_learner()
    srl_zoo_b200.install(ns)
    assert isinstance(ns.SRLModules(state_dim=200, cuda=True, model_type="custom_cnn", losses=["vae", "perceptual"]), FakeRefModules)
    assert isinstance(ns.SRLModules(state_dim=200, action_dim=6, model_type="custom_cnn", cuda=True, losses=["dae"]), FakeRefModules)
    ns2 = fake_learner()
    srl_zoo_b200.install(ns2)   # a fresh install without a perceptual model: the same DAE construction is hot
    assert isinstance(ns2.SRLModules(state_dim=200, action_dim=6, model_type="custom_cnn", cuda=True, losses=["dae"]), srl_zoo_b200.B200SRLModules)


def test_split_model_dispatch_and_masks():
    """SRLModulesSplit (models/modules.py:103-288): dispatch through the installed name, and detachSplit as a column mask
    (incl. a split that shares the dimensions of the one before it, n_dim = -1)"""
    from collections import OrderedDict
    from srl_zoo_b200.modules import split_masks

    class FakeSplit(FakeRefModules):
        pass
    ns = fake_learner()
    ns.SRLModulesSplit = FakeSplit
    srl_zoo_b200.install(ns)
    sd = OrderedDict([("autoencoder", 150), ("forward", 50), ("inverse", -1)])
    m = ns.SRLModulesSplit(state_dim=200, action_dim=6, cuda=True, model_type="custom_cnn", losses=["autoencoder", "forward", "inverse"],
                           split_dimensions=sd)
    assert isinstance(m, srl_zoo_b200.B200SRLModulesSplit)
    assert "_mask_forward" not in m.state_dict()                       # masks are not part of srl_model.pth
    cold = ns.SRLModulesSplit(state_dim=200, action_dim=6, cuda=False, model_type="custom_cnn", losses=["autoencoder", "forward", "inverse"],
                              split_dimensions=sd)
    assert isinstance(cold, FakeSplit) and cold.kw["split_dimensions"] is sd
    masks = split_masks(sd, 200)
    assert masks["autoencoder"].sum() == 150 and masks["autoencoder"][:150].all()
    assert masks["forward"][150:].all() and masks["forward"].sum() == 50
    assert torch.equal(masks["inverse"], masks["forward"])             # shared dimensions


REF_ROOT = os.environ.get("SRL_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "models")), reason="live reference only in the build container")
def test_split_masks_equal_the_reference_detach_split():
    """split_masks() against the LIVE reference's SRLModulesSplit.detachSplit for several split layouts (CPU, mlp model)"""
    from collections import OrderedDict
    from oracle import ref_loader
    from srl_zoo_b200.modules import split_masks
    ref = ref_loader.load()
    for sd, losses in ((OrderedDict([("autoencoder", 6), ("forward", 2), ("inverse", -1)]), ["autoencoder", "forward", "inverse"]),
                       (OrderedDict([("vae", 3), ("reward", 4), ("inverse", 1)]), ["vae", "reward", "inverse"]),
                       (OrderedDict([("autoencoder", 5), ("inverse", 3)]), ["autoencoder", "inverse"])):
        S = sum(v for v in sd.values() if v > 0)
        model = ref.modules.SRLModulesSplit(state_dim=S, action_dim=4, cuda=False, model_type="mlp", losses=losses, split_dimensions=sd)
        t = torch.arange(1, 3 * S + 1, dtype=torch.float32).reshape(3, S)
        masks = split_masks(sd, S)
        for key in sd:
            assert torch.equal(model.detachSplit(t, key), t * masks[key][None, :]), (dict(sd), key)


def test_dispatch_without_a_reference_class_fails_loudly():
    factory = inst._make_dispatch(None)
    assert isinstance(factory(200, 6, True, "custom_cnn", ["dae"]), srl_zoo_b200.B200SRLModules)
    with pytest.raises(ValueError):
        factory(200, 6, False, "mlp", ["autoencoder"])


def test_loss_functions_route_cpu_tensors_to_the_reference():
    ns = fake_learner()
    srl_zoo_b200.install(ns)
    assert ns.LossManager is b200_losses.LossManager          # pure host bookkeeping, same protocol
    x = torch.zeros(2, 3)
    for name in LOSS_NAMES:
        assert getattr(ns, name)(x, x, x, x, 1.0, None) == "ref:" + name
    assert [c[0] for c in ns.calls] == list(LOSS_NAMES)
    # LossManager protocol (losses/losses.py:19-59): weighted sum, history only for positive weights
    lin = torch.nn.Sequential(torch.nn.Linear(2, 2))            # parameter names '0.weight', '0.bias'
    hist = {"a": [], "b": []}
    lm = b200_losses.LossManager(lin, hist)
    lm.addToLosses("a", 2.0, torch.tensor(1.5))
    lm.addToLosses("b", 0.0, torch.tensor(7.0))
    assert float(lm.computeTotalLoss()) == 3.0
    lm.updateLossHistory()
    assert hist == {"a": [3.0], "b": []}
    lm.resetLosses()
    assert lm.names == [] and lm.losses == []
    assert len(lm.reg_params) == 1                              # '.bias' names excluded, the reference's own filter


REF = os.environ.get("SRL_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="live reference only in the build container")
def test_config0_mlp_on_cpu_runs_through_the_installed_names(monkeypatch):
    """BASELINE.json configs[0]: autoencoder loss, --model-type mlp, state-dim 200, 64x64x3 synthetic obs, bs=32 on CPU --
    one minibatch body of models/learner.py:373-497 with the names install() rebinds (they must fall through to the
    reference's own classes and functions and train)."""
    sys.dont_write_bytecode = True
    monkeypatch.syspath_prepend(REF)
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.colored = lambda s, *a, **k: s
        monkeypatch.setitem(sys.modules, "termcolor", m)
    import preprocessing.preprocess as PP
    from models.modules import SRLModules as RefSRLModules
    import losses.losses as RL
    monkeypatch.setattr(PP, "IMAGE_WIDTH", 64)
    monkeypatch.setattr(PP, "IMAGE_HEIGHT", 64)
    ns = types.SimpleNamespace(SRLModules=RefSRLModules, LossManager=RL.LossManager)
    for name in LOSS_NAMES:
        setattr(ns, name, getattr(RL, name))
    srl_zoo_b200.install(ns)
    torch.manual_seed(1)
    model = ns.SRLModules(state_dim=200, action_dim=6, cuda=False, model_type="mlp", losses=["autoencoder"])
    assert isinstance(model, RefSRLModules)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=0.005)
    g = torch.Generator().manual_seed(1234)
    obs, nobs = torch.randn(32, 3, 64, 64, generator=g), torch.randn(32, 3, 64, 64, generator=g)
    before = [p.detach().clone() for p in model.parameters()]
    lm = ns.LossManager(model, {"reconstruction_loss": []})
    losses = []
    for _ in range(3):
        opt.zero_grad()
        lm.resetLosses()
        (s, d), (ns_, nd) = model(obs), model(nobs)
        assert s.shape == (32, 200) and d.shape[0] == 32
        ns.autoEncoderLoss(obs, d, nobs, nd, 1.0, lm)
        lm.updateLossHistory()
        loss = lm.computeTotalLoss()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0]
    assert any(not torch.equal(a, b) for a, b in zip(before, [p.detach() for p in model.parameters()]))
    assert lm.names == ["reconstruction_loss"]

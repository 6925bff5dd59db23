"""
Pins oracle/srl_oracle.py against the LIVE reference (araffin/srl-zoo imported from /root/reference).
Runs only in the build container (the GPU box has no /root/reference).  TEST INFRASTRUCTURE.

    PYTHONDONTWRITEBYTECODE=1 python oracle/validate_against_reference.py

For every kind (ae, dae, vae, ae+forward+inverse) it checks, on identical seeds / inputs:
  * initial state_dict: same key set, bit-equal tensors (same RNG consumption order)
  * one train step replayed exactly as models/learner.py:373-497 with the reference's own
    SRLModules + LossManager + loss functions + th.optim.Adam:
    per-loss scalars, states, decoded, every gradient, BN buffers, parameters after Adam.
Exit code 0 = oracle pinned.
"""
import os
import sys
import types

REF = os.environ.get("SRL_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_stub("termcolor", colored=lambda s, *a, **k: s)  # utils.py:10 (not installed here)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.modules import SRLModules  # noqa: E402  (reference)
import losses.losses as RL  # noqa: E402  (reference)
from oracle import srl_oracle as O  # noqa: E402


def ref_step(kind, model, opt, obs, nobs, actions, eps_pair, rects_pair, use_fwd, use_inv, beta=1.0):
    """models/learner.py:373-497 with the reference's own objects."""
    W = {"forward": 1.0, "inverse": 2.0, "autoencoder": 1.0, "vae": 0.5e-6, "dae": 1.0}
    lm = RL.LossManager(model, None)
    model.train()
    opt.zero_grad()
    lm.resetLosses()
    out = {}
    if kind == "ae":
        (s, d), (ns, nd) = model(obs), model(nobs)
    elif kind == "dae":
        x, nx = O.apply_occlusion(obs, rects_pair[0]), O.apply_occlusion(nobs, rects_pair[1])
        (s, d), (ns, nd) = model(x), model(nx)
    else:
        # make the reference's internal normal_() draw equal to the oracle's explicit eps:
        # temporarily replace Tensor.normal_ by a feeder.
        feed = list(eps_pair)
        orig = torch.Tensor.normal_

        def fake_normal_(self, *a, **k):
            return self.copy_(feed.pop(0))

        torch.Tensor.normal_ = fake_normal_
        try:
            (d, mu, lv), (nd, nmu, nlv) = model(obs), model(nobs)
        finally:
            torch.Tensor.normal_ = orig
        s, ns = model.getStates(obs), model.getStates(nobs)
        out.update(mu=mu, logvar=lv)
    if use_fwd:
        RL.forwardModelLoss(model.forwardModel(s, actions), ns, weight=W["forward"], loss_manager=lm)
    if use_inv:
        RL.inverseModelLoss(model.inverseModel(s, ns), actions, weight=W["inverse"], loss_manager=lm)
    if kind in ("ae", "dae"):
        RL.autoEncoderLoss(obs, d, nobs, nd, weight=W["dae" if kind == "dae" else "autoencoder"], loss_manager=lm)
    else:
        RL.kullbackLeiblerLoss(mu, nmu, lv, nlv, loss_manager=lm, beta=beta)
        RL.generationLoss(d, nd, obs, nobs, weight=W["vae"], loss_manager=lm)
    loss = lm.computeTotalLoss()
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    opt.step()
    out.update(losses={n: float(v) for n, v in zip(lm.names, lm.losses)}, total=float(loss), states=s.detach(),
               next_states=ns.detach(), decoded=d.detach(), grads=grads)
    return out


def close(a, b, tol, what):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    ref = max(b.abs().max().item(), 1e-30)
    ok = err <= tol * ref
    print("   %-58s max|d|=%.3e  rel=%.3e %s" % (what, err, err / ref, "ok" if ok else "FAIL"))
    return ok


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ok = True
    bs, S, A = 2, 200, 6
    obs, nobs, actions = O.synthetic_batch(bs, seed=1234)
    g = torch.Generator().manual_seed(7)
    eps_pair = (torch.randn(bs, S, generator=g), torch.randn(bs, S, generator=g))
    rng = np.random.RandomState(1)
    rects_pair = (O.sample_rects(bs, rng=rng), O.sample_rects(bs, rng=rng))
    for kind, losses, use_fwd, use_inv in (("ae", ["autoencoder"], False, False),
                                           ("dae", ["dae"], False, False),
                                           ("vae", ["vae"], False, False),
                                           ("ae", ["autoencoder", "forward", "inverse"], True, True),
                                           ("vae", ["vae", "forward", "inverse"], True, True)):
        print("== kind=%s losses=%s" % (kind, losses))
        torch.manual_seed(1)
        ref = SRLModules(state_dim=S, action_dim=A, model_type="custom_cnn", losses=losses)
        net = "vae" if kind == "vae" else "ae"
        sd = O.build_state(net, S, A, seed=1)
        rsd = ref.state_dict()
        same_keys = list(rsd.keys()) == list(sd.keys())
        print("   state_dict keys identical and ordered: %s (%d keys)" % (same_keys, len(sd)))
        ok &= same_keys
        init_equal = all(torch.equal(rsd[k], sd[k]) for k in sd)
        print("   initial tensors bit-equal: %s" % init_equal)
        ok &= init_equal
        P, B = O.split_state(sd)
        ropt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=0.005)
        oopt = O.Adam(P, lr=0.005)
        for step in range(2):
            r = ref_step(kind, ref, ropt, obs, nobs, actions, eps_pair, rects_pair, use_fwd, use_inv)
            o = O.train_step(kind, P, B, obs, nobs, actions, eps_pair[0], eps_pair[1], rects_pair[0], rects_pair[1],
                             use_forward=use_fwd, use_inverse=use_inv, optimizer=None)
            for n in r["losses"]:
                ok &= close(torch.tensor(o["losses"][n]), torch.tensor(r["losses"][n]), 1e-6, "step%d loss %s" % (step, n))
            ok &= close(o["states"], r["states"], 1e-6, "step%d states" % step)
            ok &= close(o["decoded"], r["decoded"], 1e-6, "step%d decoded" % step)
            worst = 0.0
            for k, gr in r["grads"].items():
                go = P[k].grad
                if gr is None:
                    ok &= go is None
                    continue
                e = (go - gr).abs().max().item() / max(gr.abs().max().item(), 1e-30)
                worst = max(worst, e)
            print("   step%d worst grad rel err over %d tensors: %.3e" % (step, len(r["grads"]), worst))
            ok &= worst < 1e-5
            oopt.step(P)
            rsd = ref.state_dict()
            worst = 0.0
            for k in rsd:
                mine = B[k] if O.is_buffer(k) else P[k].detach()
                e = (mine.double() - rsd[k].double()).abs().max().item() / max(rsd[k].double().abs().max().item(), 1e-30)
                worst = max(worst, e)
            print("   step%d worst post-Adam param/buffer rel err: %.3e" % (step, worst))
            ok &= worst < 1e-5
    ok &= validate_cheap_heads_and_split()
    print("ORACLE PINNED" if ok else "ORACLE MISMATCH")
    return 0 if ok else 1



def head_cases():
    """the cheap-head / split configurations pinned beside the BASELINE ones: (name, kind, losses, inverse_model_type, split_dimensions)"""
    from collections import OrderedDict
    return [("ae_mlp_reward", "ae", ["autoencoder", "inverse", "reward"], "mlp", None),
            ("ae_split", "ae", ["autoencoder", "forward", "inverse"], "linear", OrderedDict([("autoencoder", 150), ("forward", 50), ("inverse", -1)])),
            ("vae_split_mlp_reward", "vae", ["vae", "reward", "inverse"], "mlp", OrderedDict([("vae", 120), ("reward", 40), ("inverse", 40)]))]


def head_inputs():
    bs, S = 2, 200
    obs, nobs, actions = O.synthetic_batch(bs, seed=1234)
    rewards = torch.tensor([1, 0])
    g = torch.Generator().manual_seed(7)
    eps_pair = (torch.randn(bs, S, generator=g), torch.randn(bs, S, generator=g))
    return obs, nobs, actions, rewards, eps_pair


def ref_heads_step(kind, losses, inv_type, split, obs, nobs, actions, rewards, eps_pair, S=200, A=6):
    """one forward + backward of the reference's OWN modules and loss functions for a cheap-head / split configuration
    (models/forward_inverse.py:50-56,78-95, models/modules.py:103-288, losses/losses.py:102-170) -> (module, LossManager, states, decoded)"""
    from models.modules import SRLModulesSplit
    torch.manual_seed(1)
    if split is None:
        ref = SRLModules(state_dim=S, action_dim=A, model_type="custom_cnn", losses=losses, inverse_model_type=inv_type)
    else:
        ref = SRLModulesSplit(state_dim=S, action_dim=A, model_type="custom_cnn", losses=losses, split_dimensions=split, inverse_model_type=inv_type)
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    lm = RL.LossManager(ref, None)
    ref.train()
    if kind == "vae":
        feed = list(eps_pair)
        orig = torch.Tensor.normal_
        torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(feed.pop(0))
        try:
            (d, mu, lv), (nd, nmu, nlv) = ref(obs), ref(nobs)
        finally:
            torch.Tensor.normal_ = orig
        s, ns = ref.getStates(obs), ref.getStates(nobs)
    else:
        (s, d), (ns, nd) = ref(obs), ref(nobs)
    if "forward" in losses:
        RL.forwardModelLoss(ref.forwardModel(s, actions), ns, weight=1.0, loss_manager=lm)
    if "inverse" in losses:
        RL.inverseModelLoss(ref.inverseModel(s, ns), actions, weight=2.0, loss_manager=lm)
    if "reward" in losses:
        RL.rewardModelLoss(ref.rewardModel(s, ns), rewards, weight=1.0, loss_manager=lm)
    if kind == "ae":
        RL.autoEncoderLoss(obs, d, nobs, nd, weight=1.0, loss_manager=lm)
    else:
        RL.kullbackLeiblerLoss(mu, nmu, lv, nlv, loss_manager=lm, beta=1.0)
        RL.generationLoss(d, nd, obs, nobs, weight=0.5e-6, loss_manager=lm)
    lm.computeTotalLoss().backward()
    return ref, init, lm, s.detach(), d.detach()


def oracle_heads_step(kind, losses, inv_type, split, obs, nobs, actions, rewards, eps_pair, S=200, A=6):
    sd = O.build_state("vae" if kind == "vae" else "ae", S, A, seed=1, inverse_model_type=inv_type)
    P, B = O.split_state(sd)
    o = O.train_step(kind, P, B, obs, nobs, actions, eps_pair[0], eps_pair[1], use_forward="forward" in losses,
                     use_inverse="inverse" in losses, use_reward="reward" in losses, rewards=rewards, split_dimensions=split)
    return sd, P, B, o


def validate_cheap_heads_and_split():
    """mlp inverse head (models/forward_inverse.py:50-56), reward head (:78-95, losses/losses.py:158-170) and SRLModulesSplit
    (models/modules.py:103-288): one backward pass of the reference's own modules + loss functions against the oracle."""
    ok = True
    obs, nobs, actions, rewards, eps_pair = head_inputs()
    for name, kind, losses, inv_type, split in head_cases():
        print("== kind=%s losses=%s inverse=%s split=%s" % (kind, losses, inv_type, None if split is None else dict(split)))
        ref, init, lm, s, d = ref_heads_step(kind, losses, inv_type, split, obs, nobs, actions, rewards, eps_pair)
        sd, P, B, o = oracle_heads_step(kind, losses, inv_type, split, obs, nobs, actions, rewards, eps_pair)
        same = list(init.keys()) == list(sd.keys()) and all(torch.equal(init[k], sd[k]) for k in sd)
        print("   state_dict keys + initial tensors identical: %s" % same)
        ok &= same
        for n, v in zip(lm.names, lm.losses):
            ok &= close(torch.tensor(o["losses"][n]), v.detach(), 1e-6, "loss %s" % n)
        ok &= close(o["decoded"], d, 1e-6, "decoded")
        worst = 0.0
        for k, p in ref.named_parameters():
            if p.grad is None:
                ok &= P[k].grad is None
                continue
            worst = max(worst, (P[k].grad - p.grad).abs().max().item() / max(p.grad.abs().max().item(), 1e-30))
        print("   worst grad rel err: %.3e" % worst)
        ok &= worst < 1e-5
    return ok


if __name__ == "__main__":
    sys.exit(main())

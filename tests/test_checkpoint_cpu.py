"""Resume bookkeeping (srl_zoo_b200/checkpoint.py) on CPU: the flat Adam moments of the fused engine and
torch.optim.Adam's own state_dict (what the reference's optimizer holds, models/learner.py:199) convert into each other
losslessly, and a run resumed through the conversion continues bit-identically."""
import copy

import pytest
import torch

from srl_zoo_b200.checkpoint import adam_state_from_torch, adam_state_to_torch


def make():
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    x, y = torch.randn(16, 5), torch.randn(16, 3)
    return net, x, y


def steps(net, opt, x, y, n):
    for _ in range(n):
        opt.zero_grad()
        ((net(x) - y) ** 2).mean().backward()
        opt.step()


def test_round_trip_and_bit_identical_resume():
    net, x, y = make()
    opt = torch.optim.Adam(net.parameters(), lr=0.005)
    steps(net, opt, x, y, 3)
    params = [p for p in net.parameters() if p.requires_grad]
    n = sum(p.numel() for p in params)
    m, v = torch.empty(n), torch.empty(n)
    step = adam_state_from_torch(opt.state_dict(), params, m, v)
    assert step == 3
    off = 0
    for p in params:   # flat layout = parameter order, as the engine's buffers
        assert torch.equal(m[off:off + p.numel()].view_as(p), opt.state[p]["exp_avg"])
        off += p.numel()
    sd = adam_state_to_torch(params, m, v, step, lr=0.005)
    # resume in a fresh optimizer on a copy of the model, continue both
    net2 = copy.deepcopy(net)
    opt2 = torch.optim.Adam(net2.parameters(), lr=0.005)
    opt2.load_state_dict(sd)
    steps(net, opt, x, y, 4)
    steps(net2, opt2, x, y, 4)
    for a, b in zip(net.parameters(), net2.parameters()):
        assert torch.equal(a, b)
    m2, v2 = torch.empty(n), torch.empty(n)
    assert adam_state_from_torch(opt2.state_dict(), [p for p in net2.parameters()], m2, v2) == 7


def test_fresh_optimizer_and_rejections():
    net, x, y = make()
    params = list(net.parameters())
    n = sum(p.numel() for p in params)
    m, v = torch.ones(n), torch.ones(n)
    opt = torch.optim.Adam(params, lr=0.005)
    assert adam_state_from_torch(opt.state_dict(), params, m, v) == 0 and m.abs().sum() == 0 and v.abs().sum() == 0
    assert adam_state_to_torch(params, m, v, 0)["state"] == {}
    with pytest.raises(ValueError):
        adam_state_to_torch(params, torch.zeros(n + 1), torch.zeros(n + 1), 1)
    bad = torch.optim.Adam(params, lr=0.005, weight_decay=0.1).state_dict()
    with pytest.raises(ValueError):
        adam_state_from_torch(bad, params, m, v)
    with pytest.raises(ValueError):
        adam_state_from_torch(opt.state_dict(), params[:-1], m, v)
    steps(net, opt, x, y, 1)
    sd = opt.state_dict()
    sd["state"][0]["step"] = torch.tensor(5.0)   # parameters at different steps: not representable
    with pytest.raises(ValueError):
        adam_state_from_torch(sd, params, m, v)


def test_parameter_that_never_gets_a_gradient():
    """reward_net (always) and forward_net / inverse_net (when those losses are off) never receive a gradient on the hot
    path: torch.optim.Adam keeps no state entry for them (Appendix A.9).  Such a checkpoint must load, and the flat
    moments must write the same sparse state back."""
    torch.manual_seed(5)
    used, unused = torch.nn.Linear(5, 3), torch.nn.Linear(4, 2)
    params = list(unused.parameters()) + list(used.parameters())     # heads first, as models/modules.py:37-49
    x, y = torch.randn(8, 5), torch.randn(8, 3)
    opt = torch.optim.Adam(params, lr=0.005)
    for _ in range(4):
        opt.zero_grad()
        ((used(x) - y) ** 2).mean().backward()
        opt.step()
    sd = opt.state_dict()
    assert sorted(sd["state"].keys()) == [2, 3]                      # no entries for the unused layer
    n = sum(p.numel() for p in params)
    m, v = torch.ones(n), torch.ones(n)
    assert adam_state_from_torch(sd, params, m, v) == 4
    k = sum(p.numel() for p in unused.parameters())
    assert m[:k].abs().sum() == 0 and v[:k].abs().sum() == 0 and m[k:].abs().sum() > 0
    back = adam_state_to_torch(params, m, v, 4)
    assert sorted(back["state"].keys()) == [2, 3]
    opt2 = torch.optim.Adam(params, lr=0.005)
    opt2.load_state_dict(back)
    for i in (2, 3):
        assert torch.equal(opt2.state_dict()["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])

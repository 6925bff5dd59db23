"""Optimizer-state interchange for resume.

The reference saves only `srl_model.pth` (`models/learner.py:352,518`) and restarts Adam from zero moments; its optimizer is
`th.optim.Adam(learnable_params, lr)` over `[p for p in model.parameters() if p.requires_grad]` (`models/learner.py:194-199`).
The fused engine keeps Adam's first / second moments as two flat fp32 buffers in that same parameter order, so its state maps
one-to-one onto `torch.optim.Adam.state_dict()`:  a checkpoint written by either side resumes on the other.

Pure tensor bookkeeping (no CUDA, no libsrlz): unit-tested on CPU against torch.optim.Adam itself."""
import torch


def adam_state_to_torch(params, m_flat, v_flat, step, lr=0.005, betas=(0.9, 0.999), eps=1e-8):
    """flat moments -> a dict `torch.optim.Adam(params, lr).load_state_dict()` accepts.
    params: the learnable parameters in optimizer order; m_flat / v_flat: flat buffers of their total size."""
    state, off = {}, 0
    for i, p in enumerate(params):
        k = p.numel()
        # a parameter whose moments are exactly zero never received a gradient (unused heads, Appendix A.9):
        # torch.optim.Adam holds no state entry for it, so none is written
        if step > 0 and (bool(m_flat[off:off + k].any()) or bool(v_flat[off:off + k].any())):
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": m_flat[off:off + k].detach().reshape(p.shape).clone(),
                        "exp_avg_sq": v_flat[off:off + k].detach().reshape(p.shape).clone()}
        off += k
    if off != m_flat.numel() or off != v_flat.numel():
        raise ValueError("moment buffers hold %d / %d values, the parameters %d" % (m_flat.numel(), v_flat.numel(), off))
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None,
             "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
             "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def adam_state_from_torch(sd, params, m_flat, v_flat):
    """`torch.optim.Adam.state_dict()` -> flat moments (in place); returns the step count (0 for a fresh optimizer).
    Only what the fused Adam implements is accepted: one parameter group, no weight decay, no amsgrad / maximize."""
    groups = sd["param_groups"]
    if len(groups) != 1:
        raise ValueError("one parameter group expected, got %d" % len(groups))
    g = groups[0]
    if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
        raise ValueError("weight_decay / amsgrad / maximize are not part of the reference's optimizer (models/learner.py:199)")
    ids = list(g["params"])
    if len(ids) != len(params):
        raise ValueError("checkpoint has %d parameters, the model %d" % (len(ids), len(params)))
    steps, off = set(), 0
    for pid, p in zip(ids, params):
        k = p.numel()
        st = sd["state"].get(pid)
        if st is None:
            # torch.optim.Adam creates no state for a parameter that never received a gradient (reward_net always,
            # forward_net / inverse_net when those losses are off: Appendix A.9): zero moments, no vote on the step
            m_flat[off:off + k].zero_()
            v_flat[off:off + k].zero_()
        else:
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("parameter %d: moment shape %s vs parameter shape %s" % (pid, tuple(st["exp_avg"].shape), tuple(p.shape)))
            m_flat[off:off + k].copy_(st["exp_avg"].reshape(-1))
            v_flat[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(round(float(st["step"]))))
        off += k
    if off != m_flat.numel():
        raise ValueError("moment buffer holds %d values, the parameters %d" % (m_flat.numel(), off))
    if len(steps) > 1:
        raise ValueError("parameters are at different Adam steps %s: the fused Adam keeps one step count" % sorted(steps))
    return steps.pop() if steps else 0

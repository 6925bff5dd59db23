"""
Generates tests/golden/heads_*.npz from the LIVE reference (araffin/srl-zoo imported from /root/reference): the cheap heads and the
split model (SURVEY.md 8a A9 `mlp` inverse head, 8f N4 reward head + SRLModulesSplit).  Build-container only; the fixtures it writes
are committed.  TEST INFRASTRUCTURE.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_heads.py

Each fixture holds one forward + backward of the reference's own modules and loss functions (models/forward_inverse.py:50-56,78-95,
models/modules.py:103-288, losses/losses.py:102-170,184-256) at bs=2 / state_dim=200 / seed=1 on the inputs of
`validate_against_reference.head_inputs()`: initial-weight checksums, per-loss scalars, states, decoded (subsampled ::8 + checksums),
per-parameter gradient (sum, norm) + the small gradients in full, and which parameters received no gradient.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle.validate_against_reference as V  # noqa: E402  (installs stubs, imports the reference)

SMALL = 4096


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    obs, nobs, actions, rewards, eps_pair = V.head_inputs()
    for name, kind, losses, inv_type, split in V.head_cases():
        ref, init, lm, s, d = V.ref_heads_step(kind, losses, inv_type, split, obs, nobs, actions, rewards, eps_pair)
        fx = {"meta_torch": np.array(torch.__version__), "meta_kind": np.array(kind), "meta_losses": np.array(",".join(losses)),
              "meta_inverse_model_type": np.array(inv_type),
              "meta_split": np.array("" if split is None else ",".join("%s:%d" % kv for kv in split.items())),
              "rewards": rewards.numpy(), "eps": eps_pair[0].numpy(), "next_eps": eps_pair[1].numpy(), "actions": actions.numpy(),
              "obs_checksum": np.array([obs.double().sum().item(), nobs.double().sum().item()])}
        for k, v in init.items():
            fx["w0sum/" + k] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        for n, v in zip(lm.names, lm.losses):
            fx["loss/" + n] = np.array(float(v.detach()))
        fx["states"] = s.numpy()
        fx["decoded_sub"] = d[:, :, ::8, ::8].numpy()
        fx["decoded_checksum"] = np.array([d.double().sum().item(), d.double().pow(2).sum().item()])
        nograd = []
        for k, p in ref.named_parameters():
            if p.grad is None:
                nograd.append(k)
                continue
            fx["gsum/" + k] = np.array([p.grad.double().sum().item(), p.grad.double().norm().item()])
            if p.grad.numel() <= SMALL:
                fx["g/" + k] = p.grad.numpy()
        fx["nograd"] = np.array(",".join(nograd))
        path = os.path.join(out_dir, "heads_%s.npz" % name)
        np.savez_compressed(path, **fx)
        print("wrote %s (%d arrays, %.1f KB)" % (path, len(fx), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

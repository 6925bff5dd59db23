"""
Generates tests/golden/*.npz from the LIVE reference (araffin/srl-zoo imported from /root/reference).
Build-container only; the fixtures it writes are committed.  TEST INFRASTRUCTURE.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Each fixture holds, for one loss configuration at bs=2 / state_dim=200 / seed=1:
  inputs     : regenerated from seeds by oracle.srl_oracle.synthetic_batch (seed stored) + eps, rects stored
  weights    : regenerated from torch.manual_seed(1) init; per-tensor checksums stored (sum, abs-sum)
  outputs of the reference's own train step (models/learner.py:373-497): per-loss scalars, states,
  decoded (subsampled ::8 + full-tensor checksums), per-parameter gradient (norm, sum) + the small
  gradients in full, BN buffers after the step, parameters after one Adam step (small ones in full,
  checksums for all), eval-mode getStates.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle.validate_against_reference as V  # noqa: E402  (installs stubs, imports the reference)
from oracle import srl_oracle as O  # noqa: E402

CONFIGS = {
    "ae": ("ae", ["autoencoder"], False, False),
    "dae": ("dae", ["dae"], False, False),
    "vae": ("vae", ["vae"], False, False),
    "ae_fwd_inv": ("ae", ["autoencoder", "forward", "inverse"], True, True),
}
SMALL = 4096  # tensors up to this many elements are stored in full


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    bs, S, A, seed = 2, 200, 6, 1
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    obs, nobs, actions = O.synthetic_batch(bs, seed=1234)
    g = torch.Generator().manual_seed(7)
    eps_pair = (torch.randn(bs, S, generator=g), torch.randn(bs, S, generator=g))
    rng = np.random.RandomState(1)
    rects_pair = (O.sample_rects(bs, rng=rng), O.sample_rects(bs, rng=rng))
    for name, (kind, losses, use_fwd, use_inv) in CONFIGS.items():
        torch.manual_seed(seed)
        ref = V.SRLModules(state_dim=S, action_dim=A, model_type="custom_cnn", losses=losses)
        fx = {"meta_bs": bs, "meta_state_dim": S, "meta_action_dim": A, "meta_seed": seed, "meta_input_seed": 1234,
              "meta_torch": np.array(torch.__version__), "eps": eps_pair[0].numpy(), "next_eps": eps_pair[1].numpy(),
              "rects": rects_pair[0], "next_rects": rects_pair[1], "actions": actions.numpy(),
              "obs_checksum": np.array([obs.double().sum().item(), nobs.double().sum().item()])}
        for k, v in ref.state_dict().items():
            fx["w0sum/" + k] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        # eval-mode states before any training (headline tolerance)
        ref.eval()
        with torch.no_grad():
            fx["eval_states"] = ref.getStates(obs).numpy()
        opt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=0.005)
        r = V.ref_step(kind, ref, opt, obs, nobs, actions, eps_pair, rects_pair, use_fwd, use_inv)
        for n, v in r["losses"].items():
            fx["loss/" + n] = np.array(v)
        fx["total"] = np.array(r["total"])
        fx["states"] = r["states"].numpy()
        fx["next_states"] = r["next_states"].numpy()
        fx["decoded_sub"] = r["decoded"][:, :, ::8, ::8].numpy()
        fx["decoded_checksum"] = np.array([r["decoded"].double().sum().item(), r["decoded"].double().pow(2).sum().item()])
        if "mu" in r:
            fx["mu"] = r["mu"].detach().numpy()
            fx["logvar"] = r["logvar"].detach().numpy()
        for k, gr in r["grads"].items():
            if gr is None:
                continue
            fx["gsum/" + k] = np.array([gr.double().sum().item(), gr.double().norm().item()])
            if gr.numel() <= SMALL:
                fx["g/" + k] = gr.numpy()
        for k, v in ref.state_dict().items():
            fx["w1sum/" + k] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
            if O.is_buffer(k) or v.numel() <= SMALL:
                fx["w1/" + k] = v.numpy()
        path = os.path.join(out_dir, "step_%s.npz" % name)
        np.savez_compressed(path, **fx)
        print("wrote %s (%d arrays, %.1f KB)" % (path, len(fx), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

"""The reference's unchanged `train.py` + `models/learner.py` on the GPU under `srl_zoo_b200.install()`:
learn() runs end to end on a synthetic JPEG dataset, writes every artefact the reference's tools read (`srl_model.pth`,
`exp_config.json`, `states_rewards.npz`, `image_to_state.json`, `loss_history.npz`: learner.py:97-118,516-518, train.py:195-201),
and its loss history and learned states equal those of the STOCK reference run with the same seed and arguments (the reference's
own modules on the same GPU through cuDNN with TF32 off: the comparison run stays off the host CPU so that the suite is lean).
Needs the vendored reference copy (oracle/_ref, written by build()); both runs happen in their own processes."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode, work, extra=()):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "learner_driver.py"), "--mode", mode, "--work", str(work)] + list(extra),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "DRIVER_OK" in r.stdout, r.stdout[-3000:]
    return json.load(open(os.path.join(work, "logs", mode, "driver_info.json")))


def test_learner_with_l1_l2_regularisation_and_mlp_inverse_head(tmp_path):
    """learner.py:421-425: with --l1-reg / --l2-reg the step body adds the reference's OWN l1Loss / l2Loss (losses/losses.py:132-155,
    plain torch over `loss_manager.reg_params`) to the routed losses; their gradients meet the library's in the same `.grad` tensors.
    Together with `--inverse-model-type mlp` (forward_inverse.py:50-56) through the unchanged learner: loss history == the stock run's."""
    from oracle import ref_loader
    if ref_loader.find_root() is None:
        pytest.skip("no reference copy (oracle/_ref is written by __graft_entry__.build() in the build container)")
    extra = ["--losses", "autoencoder", "inverse", "--train-args=--l1-reg 1e-7 --l2-reg 1e-3 --inverse-model-type mlp"]
    b200 = _run("b200", tmp_path, extra)
    ref = _run("ref_gpu", tmp_path, extra)
    assert b200["model_class"] == "B200SRLModules" and ref["model_class"] == "SRLModules" and b200["launches"] > 1000
    want = {"train_loss", "val_loss", "reconstruction_loss", "inverse_loss", "l1_loss", "l2_loss"}
    assert set(b200["loss_history"]) == set(ref["loss_history"]) == want
    for k, v in ref["loss_history"].items():
        for a, b in zip(b200["loss_history"][k], v):
            assert abs(a - b) <= 2e-4 * abs(b), (k, a, b)
    sr = np.load(os.path.join(tmp_path, "logs", "b200", "states_rewards.npz"))
    sr_ref = np.load(os.path.join(tmp_path, "logs", "ref_gpu", "states_rewards.npz"))
    rel = np.linalg.norm(sr["states"] - sr_ref["states"], axis=1) / np.linalg.norm(sr_ref["states"], axis=1)
    assert rel.max() < 5e-4, rel.max()


@pytest.mark.parametrize("losses", [["autoencoder"], ["vae", "forward", "inverse"]])
def test_unchanged_learner_under_install_matches_stock_run(tmp_path, losses):
    from oracle import ref_loader
    if ref_loader.find_root() is None:
        pytest.skip("no reference copy (oracle/_ref is written by __graft_entry__.build() in the build container)")
    extra = ["--losses"] + losses
    b200 = _run("b200", tmp_path, extra)
    cpu = _run("ref_gpu", tmp_path, extra)          # the stock reference (named `cpu` below for brevity: the comparison run)
    assert b200["model_class"] == "B200SRLModules" and b200["device"].startswith("cuda") and b200["launches"] > 1000
    assert cpu["model_class"] == "SRLModules" and cpu["device"].startswith("cuda")
    # loss history: same keys (loss_history.npz keys, train.py:201), same values (lr = 1e-5: the trajectories stay together)
    assert sorted(b200["loss_history"]) == sorted(cpu["loss_history"])
    want = {"train_loss", "val_loss"} | ({"reconstruction_loss"} if "autoencoder" in losses else {"generation_loss", "kl_loss", "forward_loss", "inverse_loss"})
    assert set(b200["loss_history"]) == want
    # (the VAE draws eps with torch's CUDA generator in both runs, models/models.py:161: same seed, same call order, same draws)
    vae = "vae" in losses
    for k, v in cpu["loss_history"].items():
        assert len(v) == len(b200["loss_history"][k]) == 2
        tol = 2e-4
        for a, b in zip(b200["loss_history"][k], v):
            assert abs(a - b) <= tol * abs(b), (k, a, b)
    # artefacts of the drop-in run
    log = os.path.join(tmp_path, "logs", "b200")
    for f in ("srl_model.pth", "exp_config.json", "states_rewards.npz", "image_to_state.json", "loss_history.npz"):
        assert os.path.isfile(os.path.join(log, f)), f
    cfg = json.load(open(os.path.join(log, "exp_config.json")))
    assert cfg["state-dim"] == 200 and cfg["model-type"] == "custom_cnn" and sorted(cfg["losses"]) == sorted(losses)
    hist = np.load(os.path.join(log, "loss_history.npz"))
    assert set(hist.files) == want
    sr, sr_cpu = np.load(os.path.join(log, "states_rewards.npz")), np.load(os.path.join(tmp_path, "logs", "ref_gpu", "states_rewards.npz"))
    assert sr["states"].shape == (41, 200) and sr["rewards"].shape == (41,)
    rel = np.linalg.norm(sr["states"] - sr_cpu["states"], axis=1) / np.linalg.norm(sr_cpu["states"], axis=1)
    assert rel.max() < 5e-4, rel.max()   # learned states of the two runs (after 2 epochs of lr = 1e-5 training each)
    # srl_model.pth written under the drop-in loads into the REFERENCE's own class (loadSavedModel path, learner.py:217-257) ...
    ref = ref_loader.load()
    sd = torch.load(os.path.join(log, "srl_model.pth"), map_location="cpu")
    ref_model = ref.modules.SRLModules(state_dim=200, action_dim=6, model_type="custom_cnn", cuda=False, losses=losses)
    ref_model.load_state_dict(sd)
    # ... and predicts the saved states from the saved images with the reference's own loader preprocessing
    import cv2
    from preprocessing.data_loader import preprocessImage
    ims = []
    for i in (0, 7, 40):
        im = cv2.imread(os.path.join(tmp_path, "data", "synth", "record_000", "frame%06d.jpg" % i))
        im = preprocessImage(im)
        ims.append(torch.tensor(im.reshape((1,) + im.shape).transpose(0, 3, 2, 1)))
    ref_model.eval()
    with torch.no_grad():
        st = ref_model.getStates(torch.cat(ims))
    assert H.norm_rel(torch.from_numpy(sr["states"][[0, 7, 40]]), st) < 1e-4

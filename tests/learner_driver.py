"""Runs the reference's UNCHANGED train.py (train.py:23-212 -> SRL4robotics.learn, models/learner.py:259-579) on a synthetic JPEG
dataset, in its own process (the reference's loader forks a worker): stock on the CPU (`--mode cpu`), stock on the GPU (`--mode
ref_gpu`: the reference's own modules through cuDNN / cuBLAS with TF32 off) or on the GPU with srl_zoo_b200.install() applied to
`models.learner` (`--mode b200`).  Not a pytest: tests/test_gpu_learner.py launches it.

    python tests/learner_driver.py --mode b200 --work DIR [--losses autoencoder] [--epochs 2] [-bs 8] [-lr 1e-5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["cpu", "ref_gpu", "b200"], required=True)
    ap.add_argument("--work", required=True)
    ap.add_argument("--losses", nargs="+", default=["autoencoder"])
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("-bs", type=int, default=8)
    ap.add_argument("-lr", type=float, default=1e-5)
    ap.add_argument("--frames", type=int, default=41)
    ap.add_argument("--train-args", default="", help="further train.py arguments, space separated (e.g. '--l1-reg 1e-6 --inverse-model-type mlp')")
    a = ap.parse_args()
    from oracle import ref_loader, synth_dataset
    if not os.path.isdir(os.path.join(a.work, "data", "synth")):
        synth_dataset.make_dataset(a.work, n_frames=a.frames)
    log = os.path.join(a.work, "logs", a.mode)
    info = {}

    def before(learner):
        if a.mode == "ref_gpu":   # stock torch lets cuDNN / cuBLAS use TF32 (Appendix A.12): the comparison run must be true fp32
            import torch
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        if a.mode == "b200":
            import srl_zoo_b200
            import models.modules
            srl_zoo_b200.install(learner, models.modules)
        real = learner.SRL4robotics.learn

        def timed(self, *args, **kw):
            info["model_class"] = type(self.model).__name__
            info["device"] = str(self.device)
            t0 = time.perf_counter()
            out = real(self, *args, **kw)
            info["learn_seconds"] = time.perf_counter() - t0
            return out
        learner.SRL4robotics.learn = timed   # stopwatch + introspection around the unchanged learn()

    argv = ["--no-display-plots", "--epochs", str(a.epochs), "--losses"] + a.losses + ["--model-type", "custom_cnn", "--state-dim", "200",
            "-bs", str(a.bs), "-lr", str(a.lr), "--data-folder", "synth", "--log-folder", log]
    argv += a.train_args.split()
    if a.mode == "cpu":
        argv.insert(0, "--no-cuda")
    g = ref_loader.run_train_py(a.work, argv, before_main=before)
    info["loss_history"] = {k: [float(x) for x in v] for k, v in g["loss_history"].items()}
    if a.mode == "b200":
        import srl_zoo_b200
        info["launches"] = int(srl_zoo_b200.lib.srlz_launch_count())
    with open(os.path.join(log, "driver_info.json"), "w") as f:
        json.dump(info, f)
    print("DRIVER_OK", json.dumps(info))


if __name__ == "__main__":
    main()

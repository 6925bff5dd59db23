"""
ORACLE -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Copies the reference files the hot path needs from the read-only reference checkout into the
git-ignored `oracle/_ref/`, so that the UNMODIFIED reference modules travel to the GPU box with the repository snapshot
(`/root/reference` does not exist there).  Nothing under `oracle/_ref/` is ever committed, and no product file reads it:
its consumers are `bench.py --impl reference` / `cpu_baseline` (the reference's own modules timed on the host cores) and the
tests that run the reference's unchanged `learn()` under `srl_zoo_b200.install()`.

    python oracle/vendor_reference.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/models.py", "models/autoencoders.py", "models/vae.py", "models/modules.py",
         "models/forward_inverse.py", "models/priors.py", "models/triplet.py", "models/supervised.py", "models/custom_layers.py",
         "models/learner.py", "losses/__init__.py", "losses/losses.py", "losses/utils.py", "preprocessing/__init__.py",
         "preprocessing/preprocess.py", "preprocessing/utils.py", "preprocessing/data_loader.py", "utils.py", "pipeline.py",
         "train.py", "LICENSE"]


def vendor(src="/root/reference", dst=DST):
    """-> dst when the copy exists afterwards (copied now or earlier), None when there is no reference checkout to copy from"""
    if not os.path.isdir(os.path.join(src, "models")):
        return dst if os.path.isfile(os.path.join(dst, "models", "learner.py")) else None
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.isfile(d) or open(s, "rb").read() != open(d, "rb").read():
            shutil.copyfile(s, d)
    return dst


if __name__ == "__main__":
    print(vendor(*(sys.argv[1:2])))

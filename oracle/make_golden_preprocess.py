"""
Generates tests/golden/preprocess_u8.npz from the LIVE reference's own normalisation (preprocessing/utils.py:6-35) and the
loader's layout expression (preprocessing/data_loader.py:255).  Build-container only; the fixture is committed.
TEST INFRASTRUCTURE.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_preprocess.py
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SRL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

from preprocessing.utils import preprocessInput  # noqa: E402  (reference)


def reference_tensor(im_u8, rect=None):
    im = preprocessInput(im_u8.astype(np.float32), mode="image_net")       # data_loader.py:53
    if rect is not None:
        h1, h2, w1, w2 = rect
        im[h1:h2, w1:w2, :] = 0.                                           # data_loader.py:63
    return im.reshape((1,) + im.shape).transpose(0, 3, 2, 1).copy()        # data_loader.py:255


def main():
    rng = np.random.RandomState(7)
    im = rng.randint(0, 256, size=(12, 10, 3)).astype(np.uint8)            # H=12, W=10: the W/H swap is visible
    rect = (2, 9, 1, 6)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "preprocess_u8.npz")
    # the normalisation is elementwise in (byte value, channel): its complete truth table from the reference, 256 x 3 floats,
    # pins the arithmetic of a full-size (224 x 224) frame without committing one
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2)          # (256, 1, 3): value v in every channel
    lut = preprocessInput(ramp.astype(np.float32), mode="image_net").reshape(256, 3).copy()
    np.savez_compressed(out, image=im, rect=np.array(rect, dtype=np.int32), plain=reference_tensor(im), masked=reference_tensor(im, rect), lut=lut)
    print("wrote", out)


if __name__ == "__main__":
    main()

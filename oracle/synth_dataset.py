"""
ORACLE -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Writes a small synthetic dataset folder in the reference's on-disk format (docs/guide/config.rst:40-55, utils.py:95-123,
preprocessing/data_loader.py:195-256): `data/<name>/record_000/frame%06d.jpg` (224x224 JPEGs written with cv2),
`preprocessed_data.npz` {actions, rewards, episode_starts}, `ground_truth.npz` {images_path, ground_truth_states,
target_positions} and `dataset_config.json`, so that the reference's literal train.py / learn() can run with no downloads.
"""
import json
import os

import numpy as np


def make_dataset(workdir, name="synth", n_frames=200, n_actions=6, seed=0, size=224):
    import cv2
    rng = np.random.RandomState(seed)
    folder = os.path.join(workdir, "data", name)
    rec = os.path.join(folder, "record_000")
    os.makedirs(rec, exist_ok=True)
    paths = []
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(n_frames):
        # a moving blob on a textured background: compressible, non-constant images
        cx, cy = 112 + 80 * np.cos(0.07 * i), 112 + 80 * np.sin(0.05 * i)
        blob = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 18.0 ** 2))
        img = np.stack([(80 + 120 * blob), (60 + 40 * np.sin(xx / 9.0) + 100 * blob), (90 + 50 * np.cos(yy / 11.0))], axis=-1)
        img = np.clip(img + rng.randint(0, 12, img.shape), 0, 255).astype(np.uint8)
        rel = "%s/record_000/frame%06d" % (name, i)
        cv2.imwrite(os.path.join(workdir, "data", rel + ".jpg"), img)
        paths.append(rel + ".jpg")
    actions = rng.randint(0, n_actions, n_frames).astype(np.int64)
    actions[:n_actions] = np.arange(n_actions)                       # every action present (n_actions = max + 1, train.py:140)
    rewards = (rng.rand(n_frames) < 0.1).astype(np.int64)
    episode_starts = np.zeros(n_frames, dtype=bool)
    episode_starts[0] = True
    np.savez(os.path.join(folder, "preprocessed_data.npz"), actions=actions, rewards=rewards, episode_starts=episode_starts)
    np.savez(os.path.join(folder, "ground_truth.npz"), images_path=np.array(paths), ground_truth_states=rng.randn(n_frames, 3),
             target_positions=rng.randn(1, 3))
    with open(os.path.join(folder, "dataset_config.json"), "w") as f:
        json.dump({"relative_pos": False}, f)
    return name

"""Host-side mirror of the DAE occlusion sampler (preprocessing/data_loader.py:23-35,55-59): draws one rectangle
per image; the zeroing itself happens inside the first encoder tile (csrc/enc0.cu) from these coordinates."""
import numpy as np

IMG = 224  # preprocessing/preprocess.py:7-8


def sample_coordinates(coord_1, max_distance, percentage, rng=np.random):
    """second coordinate within +-max_distance*percentage of the first, clipped to [0, max_distance] (data_loader.py:23-35)"""
    lo = max(0, coord_1 - max_distance * percentage)
    hi = min(coord_1 + max_distance * percentage, max_distance)
    coord_2 = rng.randint(low=int(lo), high=int(hi))
    return min(coord_1, coord_2), max(coord_1, coord_2)


def sample_rects(n, occlusion_percentage=0.5, rng=np.random):
    """-> (n,4) int32 rows (h1,h2,w1,w2); after the loader's transpose (data_loader.py:255) the zeroed block of the
    (C, W, H) tensor is [:, w1:w2, h1:h2]."""
    out = np.zeros((n, 4), dtype=np.int32)
    for i in range(n):
        h1, h2 = sample_coordinates(rng.randint(IMG), IMG, occlusion_percentage, rng)
        w1, w2 = sample_coordinates(rng.randint(IMG), IMG, occlusion_percentage, rng)
        out[i] = (h1, h2, w1, w2)
    return out

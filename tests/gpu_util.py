"""helpers for the -m gpu tests"""
import numpy as np


def random_csr(rng, n, max_len=1100, pad_front=None):
    """n random messages of random length packed at arbitrary byte offsets (with gaps)."""
    lens = rng.integers(0, max_len + 1, n)
    # make the interesting boundaries likely
    special = np.array([0, 1, 7, 8, 9, 135, 136, 137, 271, 272, 273, 543, 544, 545, 559, 560, 561, 1087, 1088, 1089])
    pick = rng.random(n) < 0.3
    lens = np.where(pick, special[rng.integers(0, len(special), n)], lens)
    off = np.zeros(n + 1, np.uint64)
    off[0] = pad_front if pad_front is not None else int(rng.integers(0, 16))
    off[1:] = off[0] + np.cumsum(lens).astype(np.uint64)
    data = rng.integers(0, 256, int(off[-1]) + 32, dtype=np.uint8)
    return data, off


def bit(bitmap, i):
    return (int(bitmap[i // 64]) >> (i % 64)) & 1

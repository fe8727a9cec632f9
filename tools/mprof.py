import sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from phant_b200 import gpu
ctx = gpu.Context(0)
rng = np.random.default_rng(1)
n = 2_000_000
k = np.unique(rng.integers(0, 256, (n, 32), dtype=np.uint8), axis=0)
n = len(k)
keys = np.ascontiguousarray(k.reshape(-1)); koff = (np.arange(n + 1) * 32).astype(np.uint32)
vals = rng.integers(0, 256, n * 80, dtype=np.uint8); voff = (np.arange(n + 1) * 80).astype(np.uint64)
for _ in range(2):
    ctx.mpt_root(keys, koff, vals, voff, n)

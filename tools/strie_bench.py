#!/usr/bin/env python3
"""U kind 1 (sparse resident trie) update latency, development tool: build N random keys, then value upserts of M keys.
  python tools/strie_bench.py --keys 4000000 --dirty 100000 --steps 5
Run under `ncu --metrics gpu__time_duration.sum` to see which kernels an update spends its device time in."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phant_b200 import gpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", type=int, default=4_000_000)
    ap.add_argument("--dirty", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    ctx = gpu.Context(0)
    rng = np.random.default_rng(6)
    keys = rng.integers(0, 256, (a.keys, 32), dtype=np.uint8)
    vals = rng.integers(0, 256, a.keys * 78, dtype=np.uint8)
    voff = (np.arange(a.keys + 1, dtype=np.uint64) * 78).astype(np.uint32)
    t = ctx.trie_open(0, kind=1)
    t0 = time.perf_counter()
    t.update(np.ascontiguousarray(keys.reshape(-1)), vals, voff, a.keys)
    build = time.perf_counter() - t0
    pick = rng.choice(a.keys, size=a.dirty, replace=False)
    k = np.ascontiguousarray(keys[pick].reshape(-1))
    v = rng.integers(0, 256, a.dirty * 78, dtype=np.uint8)
    o = (np.arange(a.dirty + 1) * 78).astype(np.uint32)
    t.update(k, v, o, a.dirty)
    ctx.reset_stats()
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        t.update(k, v, o, a.dirty)
        ts.append(1e3 * (time.perf_counter() - t0))
    st = ctx.stats()
    print(json.dumps({"keys": a.keys, "dirty": a.dirty, "build_s": build, "update_ms": sorted(ts), "launches_per_update": st["launches"] / a.steps,
                      "keccak_ms_per_update": st["keccak_ms"] / a.steps, "keccak_msgs_per_update": st["keccak_msgs"] / a.steps}))
    t.close()
    ctx.close()


if __name__ == "__main__":
    main()

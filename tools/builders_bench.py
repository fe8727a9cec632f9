#!/usr/bin/env python3
"""Timings of the builder entry points M / S / U on one GPU next to the CPU oracle (development tool).
One JSON object per line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from phant_b200 import gpu  # noqa: E402


def timeit(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    o = oracle_lib.get()
    o.use_reference_keccak(True)
    ctx = gpu.Context(0)
    rng = np.random.default_rng(1)
    # ---- M: index trie of 400 withdrawals (fixture scale) and secure tries of growing size ----
    from helpers import index_trie_items
    wd = [rng.integers(0, 256, 48, dtype=np.uint8).tobytes() for _ in range(400)]
    kv = index_trie_items(wd)
    keys, koff = oracle_lib.csr([k for k, _ in kv], np.uint32)
    vals, voff = oracle_lib.csr([v for _, v in kv], np.uint64)
    g = timeit(lambda: ctx.mpt_root(keys, koff, vals, voff, len(kv)), 20)
    c = timeit(lambda: o.mptize(kv), 20)
    print(json.dumps({"what": "M 400-item index trie", "gpu_ms": g * 1e3, "cpu_ms": c * 1e3}), flush=True)
    for n in (10_000, 200_000, 2_000_000):
        k = np.unique(rng.integers(0, 256, (n, 32), dtype=np.uint8), axis=0)
        n = len(k)
        keys = np.ascontiguousarray(k.reshape(-1))
        koff = (np.arange(n + 1) * 32).astype(np.uint32)
        vals = rng.integers(0, 256, n * 80, dtype=np.uint8)
        voff = (np.arange(n + 1) * 80).astype(np.uint64)
        ctx.reset_stats()
        g = timeit(lambda: ctx.mpt_root(keys, koff, vals, voff, n), 3)
        st = ctx.stats()
        t0 = time.perf_counter()
        out = np.zeros(32, np.uint8)
        rc = o.lib.oracle_mptize(keys.ctypes.data_as(oracle_lib.u8p), koff.ctypes.data_as(oracle_lib.u32p), vals.ctypes.data_as(oracle_lib.u8p),
                                 voff.ctypes.data_as(oracle_lib.u64p), n, out.ctypes.data_as(oracle_lib.u8p))
        c = time.perf_counter() - t0
        assert rc == 0 and ctx.mpt_root(keys, koff, vals, voff, n) == out.tobytes()
        print(json.dumps({"what": f"M secure trie {n} keys x 80 B", "gpu_ms": g * 1e3, "cpu_1thread_ms": c * 1e3, "launches_per_call": st["launches"] // 4,
                          "keys_per_s_gpu": n / g}), flush=True)
    # ---- M batched: 3,000 tries of 100 withdrawals in one forest build vs one call each ----
    lists = []
    for t in range(3000):
        wd = [rng.integers(0, 256, 48, dtype=np.uint8).tobytes() for _ in range(100)]
        lists.append(index_trie_items(wd))
    flat = [kv for l in lists for kv in l]
    keys, koff = oracle_lib.csr([k for k, _ in flat], np.uint32)
    vals, voff = oracle_lib.csr([v for _, v in flat], np.uint64)
    seg = np.zeros(len(lists) + 1, np.uint32)
    seg[1:] = np.cumsum([len(l) for l in lists])
    g = timeit(lambda: ctx.mpt_roots(keys, koff, vals, voff, seg, len(lists)), 5)
    t0 = time.perf_counter()
    want = [o.mptize(l) for l in lists[:300]]
    c = (time.perf_counter() - t0) * 10
    assert ctx.mpt_roots(keys, koff, vals, voff, seg, len(lists))[:300] == want
    print(json.dumps({"what": "M batched: 3000 index tries x 100 items, one phant_gpu_mpt_roots call", "gpu_ms": g * 1e3, "cpu_1thread_ms": c * 1e3,
                      "tries_per_s_gpu": len(lists) / g}), flush=True)
    # ---- S: state root of 500k accounts, 10% with 8 storage slots ----
    na = 500_000
    addr = rng.integers(0, 256, na * 20, dtype=np.uint8)
    nonce = rng.integers(0, 1000, na).astype(np.uint64)
    bal = np.zeros((na, 32), np.uint8); bal[:, 20:] = rng.integers(0, 256, (na, 12), dtype=np.uint8)
    code = np.zeros(1, np.uint8); coff = np.zeros(na + 1, np.uint64)
    has = rng.random(na) < 0.1
    soff = np.zeros(na + 1, np.uint64); soff[1:] = np.cumsum(np.where(has, 8, 0))
    ns = int(soff[-1])
    skeys = rng.integers(0, 256, ns * 32, dtype=np.uint8); svals = rng.integers(1, 256, ns * 32, dtype=np.uint8)
    g = timeit(lambda: ctx.state_root(na, addr, nonce, np.ascontiguousarray(bal.reshape(-1)), code, coff, skeys, svals, soff), 3)
    print(json.dumps({"what": f"S state root: {na} accounts, {ns} storage slots (host pointers, pageable)", "gpu_ms": g * 1e3,
                      "accounts_per_s": na / g}), flush=True)
    # ---- U: config C4 ----
    t = ctx.trie_open(6)
    n = 100_000
    pos = rng.choice(16 ** 6, size=n, replace=False).astype(np.uint32)
    keys = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    keys[:, 0] = (pos >> 16) & 0xff
    keys[:, 1] = (pos >> 8) & 0xff
    keys[:, 2] = pos & 0xff
    flat = np.ascontiguousarray(keys.reshape(-1))
    v = rng.integers(0, 256, n * 78, dtype=np.uint8)
    voff = (np.arange(n + 1) * 78).astype(np.uint32)
    import torch
    pk, pv, po = (torch.from_numpy(x).pin_memory() for x in (flat, v, voff))  # pinned host buffers, as a real caller would hold
    ctx.reset_stats()
    g = timeit(lambda: t.update(pk, pv, po, n), 20)
    st = ctx.stats()
    g_pageable = timeit(lambda: t.update(flat, v, voff, n), 5)
    oc = o.ctrie(6)
    vals_list = [v[78 * i:78 * i + 78].tobytes() for i in range(n)]
    t0 = time.perf_counter()
    r = oc.update(flat, vals_list)
    c = time.perf_counter() - t0
    assert r == t.update(flat, v, voff, n)
    print(json.dumps({"what": "U C4: 100k dirty leaves into 16^6-leaf trie (host pointers)", "gpu_ms": g * 1e3, "cpu_1thread_ms": c * 1e3,
                      "gpu_ms_pageable_host": g_pageable * 1e3, "launches_per_call": st["launches"] // 21}), flush=True)
    t.close()
    ctx.close()


if __name__ == "__main__":
    main()

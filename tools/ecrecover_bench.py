#!/usr/bin/env python3
"""Throughput of phant_gpu_ecrecover_batch (row N4) against the CPU oracle; development tool, one JSON line.
  python tools/ecrecover_bench.py --n 200000
Inputs: random hashes, r drawn until it is the abscissa of a curve point (so every signature recovers a key, the expensive
path), random s, recid 0/1.  Device pointers, CUDA events around the call; the oracle is timed on a small sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import oracle_lib
    from phant_b200 import gpu
    o = oracle_lib.get()
    rng = np.random.default_rng(1)
    P = 2 ** 256 - 0x1000003D1
    base = []
    while len(base) < 256:  # 256 distinct valid (r, s, recid); the batch repeats them with different hashes
        r = int.from_bytes(rng.bytes(32), "big") % (P - 1) + 1
        if pow((pow(r, 3, P) + 7) % P, (P - 1) // 2, P) == 1 and r < 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141:
            base.append(r.to_bytes(32, "big") + rng.bytes(32)[:31].rjust(32, b"\x01") + bytes([int(rng.integers(0, 2))]))
    sigs = np.frombuffer(b"".join(base[i % 256] for i in range(a.n)), np.uint8).copy()
    hashes = rng.integers(0, 256, 32 * a.n, dtype=np.uint8)
    dev = torch.device("cuda", 0)
    d_h, d_s = torch.from_numpy(hashes).to(dev), torch.from_numpy(sigs).to(dev)
    d_addr = torch.zeros(20 * a.n, dtype=torch.uint8, device=dev)
    d_ok = torch.zeros(a.n, dtype=torch.uint8, device=dev)
    ctx = gpu.Context(0, gpu.FLAG_DEVICE_PTRS)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    for _ in range(2):
        ctx.ecrecover_batch(d_h, d_s, a.n, None, d_addr, d_ok)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(side)
    for _ in range(a.iters):
        ctx.ecrecover_batch(d_h, d_s, a.n, None, d_addr, d_ok)
    e1.record(side)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    sample = 300
    t0 = time.perf_counter()
    for i in range(sample):
        o.ecrecover(hashes[32 * i:32 * i + 32].tobytes(), sigs[65 * i:65 * i + 65].tobytes())
    cpu = sample / (time.perf_counter() - t0)
    addr = d_addr.cpu().numpy().reshape(-1, 20)
    want = o.ecrecover(hashes[:32].tobytes(), sigs[:65].tobytes())
    print(json.dumps({"what": "ecrecover + address, one signature per thread", "n": a.n, "gpu_ms": ms, "sigs_per_s_gpu": a.n / ms * 1e3,
                      "all_ok": bool(d_ok.all().item()), "first_address_ok": addr[0].tobytes() == o.keccak256(want[1:])[12:],
                      "oracle_sigs_per_s_1thread": cpu}), flush=True)


if __name__ == "__main__":
    main()

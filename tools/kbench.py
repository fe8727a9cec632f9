#!/usr/bin/env python3
"""Micro-benchmark of the kernels on one GPU (development tool; bench.py is the contract benchmark).

Generates C2 / C3 witnesses in HBM, then times hash-only and verify with CUDA events through the
library's own stats, for each Keccak kernel variant.  Prints one JSON object per line.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phant_b200 import gpu  # noqa: E402


def gen(ctx, which, n, depth=8):
    n_nodes, n_bytes = ctx.synth_sizes(which, n, depth=depth)
    t = dict(
        nodes=torch.empty(n_bytes + 64, dtype=torch.uint8, device="cuda"),
        off=torch.empty(n_nodes + 1, dtype=torch.int64, device="cuda"),
        first=torch.empty(n + 1, dtype=torch.int64, device="cuda"),
        keys=torch.empty(n * 32, dtype=torch.uint8, device="cuda"),
        roots=torch.empty(n * 32, dtype=torch.uint8, device="cuda"),
        n=n, n_nodes=n_nodes, n_bytes=n_bytes)
    ctx.synth(which, n, t["nodes"], t["off"], t["first"], t["keys"], t["roots"], depth=depth)
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--which", type=int, default=2)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--variants", default="staged,staged_nobin,direct,direct_nobin,warp")
    ap.add_argument("--uniform", type=int, default=0, help="hash-only line over N random messages of this many bytes (SURVEY 8d: 532 and 112)")
    a = ap.parse_args()
    ctx = gpu.Context(0)
    if a.uniform:
        size, n = a.uniform, a.n
        g = torch.Generator(device="cuda").manual_seed(1)
        data = torch.randint(0, 256, (n * size + 64,), dtype=torch.uint8, device="cuda", generator=g)
        off = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * size
        out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
        ctx.set_flags(gpu.FLAG_DEVICE_PTRS)
        for it in range(a.iters + 1):
            if it == 1:
                ctx.reset_stats()
            ctx.keccak256_batch(data, off, n, out)
            ctx.synchronize()
        s = ctx.stats()
        ms = s["keccak_ms"] / a.iters
        print(json.dumps({"uniform_bytes": size, "n": n, "keccak_ms": round(ms, 4), "mh_s": round(n / ms / 1e3, 1),
                          "gperm_s": round(s["keccak_perms"] / a.iters / ms / 1e6, 3), "algorithmic_gb_s": round(n * (size + 32) / ms / 1e6, 1)}), flush=True)
        return
    t = gen(ctx, a.which, a.n)
    digests = torch.empty((t["n_nodes"], 32), dtype=torch.uint8, device="cuda")
    status = torch.empty(a.n, dtype=torch.uint8, device="cuda")
    bitmap = torch.zeros((a.n + 63) // 64, dtype=torch.int64, device="cuda")
    flagmap = {"staged": 0, "staged_nobin": 1 << 6, "direct": 1 << 4, "direct_nobin": (1 << 4) | (1 << 6), "warp": 1 << 5}
    for name in a.variants.split(","):
        fl = gpu.FLAG_DEVICE_PTRS | flagmap[name]
        ctx.set_flags(fl)
        iters = a.iters if name != "warp" else max(1, a.iters // 3)
        for phase in ("hash", "verify"):
            for it in range(iters + 1):
                if it == 1:
                    ctx.reset_stats()
                if phase == "hash":
                    ctx.keccak256_batch(t["nodes"], t["off"], t["n_nodes"], digests)
                else:
                    ctx.verify_proofs(a.n, t["nodes"], t["off"], t["first"], t["keys"], t["roots"], a.n, bitmap, status, None, None)
                ctx.synchronize()
            s = ctx.stats()
            k_ms = s["keccak_ms"] / iters
            w_ms = s["walk_ms"] / iters
            perms = s["keccak_perms"] / iters
            rec = dict(which=a.which, n=a.n, variant=name, phase=phase, keccak_ms=round(k_ms, 4), walk_ms=round(w_ms, 4),
                       gperm_s=round(perms / k_ms / 1e6, 3), mh_s=round(t["n_nodes"] / k_ms / 1e3, 2),
                       gb_s=round((t["n_bytes"] + 32 * t["n_nodes"]) / k_ms / 1e6, 2), launches=s["launches"] // iters)
            if phase == "verify":
                rec["proofs_s"] = round(a.n / (k_ms + w_ms) * 1e3)
                ok = int((status.cpu().numpy() != 0).sum())
                rec["accepted"] = ok
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()

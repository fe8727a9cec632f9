#!/usr/bin/env python3
"""S sharded over GPUs by the top nibble of keccak(address): every rank holds the same flat account table,
builds the subtrees of its own root-branch slots (phant_gpu_state_subtree_roots), one all-reduce of 16 x 32 bytes + flags,
every rank hashes the root branch.  Checks the result against the unsharded phant_gpu_state_root on rank 0 and prints one
JSON line.  Development tool.
  python tools/state_sharded_bench.py --accounts 500000
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/state_sharded_bench.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tables(rng, n, slots_per=0.8):
    addr = rng.integers(0, 256, (n, 20), dtype=np.uint8)
    nonce = rng.integers(0, 1000, n).astype(np.uint64)
    bal = np.zeros((n, 32), np.uint8)
    bal[:, 24:] = rng.integers(0, 256, (n, 8), dtype=np.uint8)
    code_len = np.where(rng.random(n) < 0.1, 64, 0).astype(np.uint64)
    coff = np.zeros(n + 1, np.uint64)
    coff[1:] = np.cumsum(code_len)
    code = rng.integers(0, 256, max(int(coff[-1]), 1), dtype=np.uint8)
    ns = rng.poisson(slots_per, n).astype(np.uint64)
    soff = np.zeros(n + 1, np.uint64)
    soff[1:] = np.cumsum(ns)
    m = int(soff[-1])
    skeys = rng.integers(0, 256, (max(m, 1), 32), dtype=np.uint8)
    svals = np.zeros((max(m, 1), 32), np.uint8)
    svals[:, 20:] = rng.integers(0, 256, (max(m, 1), 12), dtype=np.uint8)
    return addr, nonce, bal, code, coff, skeys, svals, soff


def select(t, keep):
    """rows of the flat table where keep is set (CSR parts re-based)"""
    addr, nonce, bal, code, coff, skeys, svals, soff = t
    idx = np.nonzero(keep)[0]
    clen = (coff[1:] - coff[:-1])[idx]
    slen = (soff[1:] - soff[:-1])[idx]
    ncoff = np.zeros(len(idx) + 1, np.uint64)
    ncoff[1:] = np.cumsum(clen)
    nsoff = np.zeros(len(idx) + 1, np.uint64)
    nsoff[1:] = np.cumsum(slen)

    def gather(rows, off, lens, width):
        total = int(lens.sum())
        if total == 0:
            return np.zeros((1, width) if width > 1 else 1, np.uint8)
        src = np.repeat(off[:-1][idx].astype(np.int64), lens.astype(np.int64)) + (np.arange(total) - np.repeat(np.cumsum(lens) - lens, lens.astype(np.int64))).astype(np.int64)
        return rows[src]
    ncode = gather(code, coff, clen, 1)
    nsk = gather(skeys, soff, slen, 32)
    nsv = gather(svals, soff, slen, 32)
    c = np.ascontiguousarray
    return len(idx), c(addr[idx]).reshape(-1), c(nonce[idx]), c(bal[idx]).reshape(-1), c(ncode), ncoff, c(nsk).reshape(-1), c(nsv).reshape(-1), nsoff


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--accounts", type=int, default=500_000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from phant_b200 import gpu, host, shard
    ctx = gpu.Context(local)
    n = a.accounts
    t = tables(np.random.default_rng(2024), n)  # the same StateDB on every rank (phant's state lives on the host)
    # which root-branch slot each account sits under: one K call over the addresses
    h = np.zeros((n, 32), np.uint8)
    ctx.keccak256_batch(np.ascontiguousarray(t[0]).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * 20), n, h)
    slot = h[:, 0] >> 4
    owner = np.array([shard.nibble_owner(v, world) for v in range(16)])[slot]
    mine = select(t, owner == rank)

    def sharded():
        refs, mask = ctx.state_subtree_roots(*mine)
        refs_all, mask_all = shard.allgather_subtree_roots(refs, mask)
        return host.keccak256(ctx, shard.root_branch_rlp(refs_all, mask_all))

    def timed(fn):
        for _ in range(3):
            r = fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            r = fn()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / a.steps], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return r, float(dt.item())

    root_sh, t_sh = timed(sharded)
    out = {"what": f"S sharded by top nibble: {n} accounts, {int(t[7][-1])} storage slots, host pointers", "n_gpus": world,
           "ms_per_root_sharded": t_sh * 1e3, "accounts_on_rank0": int(mine[0]), "root": root_sh.hex()}
    if rank == 0:
        full = (n, np.ascontiguousarray(t[0]).reshape(-1), t[1], np.ascontiguousarray(t[2]).reshape(-1), t[3], t[4],
                np.ascontiguousarray(t[5]).reshape(-1), np.ascontiguousarray(t[6]).reshape(-1), t[7])
        root_full, t_full = None, None
        for _ in range(3):
            root_full = ctx.state_root(*full)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            root_full = ctx.state_root(*full)
        t_full = (time.perf_counter() - t0) / a.steps
        out.update({"ms_per_root_one_gpu": t_full * 1e3, "roots_equal": root_full == root_sh})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)
        assert out["roots_equal"]


if __name__ == "__main__":
    main()

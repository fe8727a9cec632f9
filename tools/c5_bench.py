#!/usr/bin/env python3
"""Config C5 (BASELINE.json): N synthetic blocks x 300 tx, deduplicated witness, blocks sharded over the GPUs,
one NCCL all-reduce over the per-block reject counts.  Development tool: prints one JSON line on rank 0.
  python tools/c5_bench.py --blocks 1000            (1 GPU)
  python -m torch.distributed.run --nproc-per-node 8 ... tools/c5_bench.py --blocks 1000
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
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1000)
    ap.add_argument("--txs", type=int, default=300)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import oracle_lib
    from phant_b200 import gpu, shard
    per = (a.blocks + world - 1) // world
    b0, b1 = min(rank * per, a.blocks), min((rank + 1) * per, a.blocks)
    o = oracle_lib.get()
    t0 = time.time()
    w = o.synth_blocks(b1 - b0, txs=a.txs, first=b0, threads=max(1, 16 // world))  # setup: witness built on the host (oracle generator)
    gen_s = time.time() - t0
    n = w["n_proofs"]
    d = {k: torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else (v.astype(np.int64) if v.dtype == np.uint32 else v)).to(dev)
         for k, v in w.items() if isinstance(v, np.ndarray)}
    status = torch.zeros(n, dtype=torch.uint8, device=dev)
    bitmap = torch.zeros((n + 63) // 64, dtype=torch.int64, device=dev)
    ctx = gpu.Context(local, gpu.FLAG_DEVICE_PTRS)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    blocks_global = d["block_of_proof"] + b0

    def step():
        ctx.verify_proofs(n, d["nodes"], d["node_off"], d["proof_first"], d["keys32"], d["roots32"], n, bitmap, status, None, None,
                          n_nodes=w["n_nodes"], nodes_bytes=w["n_bytes"], node_index=d["node_index"])
        return shard.block_reject_counts(status, blocks_global, a.blocks)  # one all-reduce per batch

    for _ in range(3):
        rej = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        rej = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    tot = torch.tensor([n, w["n_nodes"], w["n_bytes"], w["n_refs"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    st = ctx.stats()
    if rank == 0:
        dt = float(t.item())
        rej = rej.cpu().numpy()
        bad = np.nonzero(rej)[0]
        expect = np.array([b for b in range(a.blocks) if b % 100 == 37])
        print(json.dumps({"config": f"C5: {a.blocks} blocks x {a.txs} tx, deduplicated witness", "n_gpus": world, "proofs": int(tot[0]),
                          "unique_nodes": int(tot[1]), "node_bytes": int(tot[2]), "node_refs": int(tot[3]), "ms_per_batch": 1e3 * dt / a.steps,
                          "proofs_per_s": int(tot[0]) * a.steps / dt, "blocks_per_s": a.blocks * a.steps / dt,
                          "kernel_ms_rank0": {"keccak": st["keccak_ms"] / a.steps, "walk": st["walk_ms"] / a.steps},
                          "rejected_blocks": bad.tolist(), "rejected_blocks_ok": bool((bad == expect).all()) if len(bad) == len(expect) else False,
                          "host_witness_build_s": gen_s}))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

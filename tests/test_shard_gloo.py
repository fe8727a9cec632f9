"""N > 1 host path on CPU: world_size 2 over gloo.  Each rank takes its proof shard, produces the shard's
verdict words (here from the oracle -- the GPU is not involved in this test of the host logic) and one
all-reduce must assemble the bitmap a single process computes over the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from phant_b200 import shard
    o = oracle_lib.get()
    lo, hi = shard.shard_range(n, rank, world)
    assert lo % 64 == 0
    # the shard's witness is regenerated from (seed, index): no data moves between ranks
    nodes, node_off, first, keys, roots = o.synth_c2(hi - lo, depth=8, first=lo, threads=2)
    bitmap, status, _, _ = o.verify_proofs(nodes, node_off, first, keys, roots, threads=2)
    local = torch.from_numpy(bitmap.view(np.int64).copy())
    g = shard.allreduce_accept_bitmap(local, lo, n)
    blocks = torch.from_numpy(((np.arange(lo, hi)) // 300).astype(np.int64))
    rej = shard.block_reject_counts(torch.from_numpy(status.copy()), blocks, (n + 299) // 300)
    np.save(os.path.join(out_dir, f"bitmap_{rank}.npy"), g.numpy())
    np.save(os.path.join(out_dir, f"rej_{rank}.npy"), rej.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bitmap_allreduce(tmp_path, oracle):
    from phant_b200 import shard
    n, world = 10_000, 2
    assert shard.shard_range(n, 0, world) == (0, 5056) and shard.shard_range(n, 1, world) == (5056, 10_000)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    nodes, node_off, first, keys, roots = oracle.synth_c2(n, depth=8, threads=4)
    want, status, _, _ = oracle.verify_proofs(nodes, node_off, first, keys, roots, threads=4)
    for r in range(world):
        got = np.load(tmp_path / f"bitmap_{r}.npy").view(np.uint64)
        assert (got == want).all()
        rej = np.load(tmp_path / f"rej_{r}.npy")
        expect = np.bincount(np.arange(n)[status == 0] // 300, minlength=(n + 299) // 300)
        assert (rej == expect).all()


def test_shard_ranges_cover_and_align():
    from phant_b200 import shard
    for n in (0, 1, 63, 64, 65, 1000, 1_000_000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                assert lo == prev and lo <= hi <= n and (lo % 64 == 0 or lo == n)
                prev = hi
            assert prev == n

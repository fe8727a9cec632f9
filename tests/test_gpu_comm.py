"""Multi-GPU behind the C ABI (comm.cu): needs two B200s (`gpurun --gpus 2`); on a one-GPU box everything but the
single-rank degenerate case is skipped.  Host logic of the same paths runs on CPU over gloo in tests/test_shard_gloo.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_shard_range_of_the_abi_equals_the_host_logic():
    """phant_gpu_shard_range / phant_gpu_sharded_bitmap_words are pure functions: callable without a device"""
    from phant_b200 import gpu, shard
    for n in (0, 1, 63, 64, 65, 1000, 1_000_000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            per = None
            for r in range(world):
                assert gpu.shard_range(n, r, world) == shard.shard_range(n, r, world)
                lo, hi = gpu.shard_range(n, r, world)
                per = max(per or 0, hi - lo)
            words = gpu.sharded_bitmap_words(n, world)
            assert words % world == 0 and words * 64 >= n and words // world * 64 >= per
    assert [gpu.nibble_owner(v, 2) for v in range(16)] == [shard.nibble_owner(v, 2) for v in range(16)]
    assert [gpu.nibble_owner(v, 8) for v in range(16)] == [shard.nibble_owner(v, 8) for v in range(16)]


@pytest.mark.gpu
def test_world_size_one_sharded_call_is_the_plain_call(oracle):
    """no communicator: phant_gpu_verify_proofs_sharded degenerates to phant_gpu_verify_proofs (host and device pointers)"""
    import torch
    from phant_b200 import gpu
    n = 5000
    o = oracle.synth_c2(n, depth=8, first=3)
    want = oracle.verify_proofs(*o, threads=4)
    ctx = gpu.Context(0)
    bitmap = np.zeros(gpu.sharded_bitmap_words(n, 1), np.uint64)
    status = np.zeros(n, np.uint8)
    ctx.verify_proofs_sharded(n, n, o[0], o[1], o[2], o[3], o[4], n, bitmap, status)
    assert (status == want[1]).all() and (bitmap[:len(want[0])] == want[0]).all()
    ctx.set_flags(gpu.FLAG_DEVICE_PTRS)
    d = [torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda() for a in o]
    d_nodes = torch.zeros(d[0].numel() + 64, dtype=torch.uint8, device="cuda")
    d_nodes[:d[0].numel()] = d[0]
    d_bitmap = torch.zeros(len(bitmap), dtype=torch.int64, device="cuda")
    d_status = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # torch's stream and the context's stream are not ordered against each other
    ctx.verify_proofs_sharded(n, n, d_nodes, d[1], d[2], d[3], d[4], n, d_bitmap, d_status)
    ctx.comm_fence()
    ctx.synchronize()
    assert (d_status.cpu().numpy() == want[1]).all()
    assert (d_bitmap.cpu().numpy().view(np.uint64)[:len(want[0])] == want[0]).all()
    counts = np.zeros(4, np.uint32)
    ctx.set_flags(0)
    ctx.block_reject_counts(status, (np.arange(n) * 4 // n).astype(np.uint32), n, 4, counts)
    assert (counts == np.bincount((np.arange(n) * 4 // n)[status == 0], minlength=4)).all()
    ctx.close()


@pytest.mark.gpu
def test_world_size_two_through_the_abi_from_cpp():
    """host/comm_test.cpp: one process, two contexts, two host threads, NCCL inside libphantgpu.so"""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    lib = os.path.join(ROOT, "phant_b200", "lib")
    exe = os.path.join(ROOT, "host", "comm_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "host", "comm_test.cpp"), f"-L{lib}", "-lphantgpu",
                    f"-Wl,-rpath,{lib}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def _worker(rank, world, n, tmp, peer=False):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import time
    import torch
    import oracle_lib
    from phant_b200 import gpu
    torch.cuda.set_device(rank)
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(gpu.comm_unique_id())
        os.rename(idf + ".tmp", idf)
    while not os.path.exists(idf):
        time.sleep(0.05)
    ctx = gpu.Context(rank, gpu.FLAG_DEVICE_PTRS)
    ctx.comm_init(open(idf, "rb").read(), rank, world)
    assert ctx.comm_info()[:2] == (rank, world)
    if peer:
        ctx.comm_enable_peer(n)   # cudaIpc mappings between the two processes; raises if the box cannot do it
        assert ctx.comm_peer_status()["enabled"]
    o = oracle_lib.get()
    lo, hi = gpu.shard_range(n, rank, world)
    w = o.synth_c2(hi - lo, depth=8, first=lo, threads=2)
    want_local = o.verify_proofs(*w, threads=2)
    d = [torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda() for a in w]
    d_nodes = torch.zeros(d[0].numel() + 64, dtype=torch.uint8, device="cuda")
    d_nodes[:d[0].numel()] = d[0]
    words = gpu.sharded_bitmap_words(n, world)
    gb = [torch.zeros(words, dtype=torch.int64, device="cuda") for _ in range(2)]
    d_status = torch.zeros(hi - lo, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # torch's stream and the context's stream are not ordered against each other
    for k in range(7 if peer else 5):  # alternate the two buffers; steps overlap on the comm stream
        ctx.verify_proofs_sharded(hi - lo, n, d_nodes, d[1], d[2], d[3], d[4], hi - lo, gb[k & 1], d_status)
    ctx.comm_fence()
    ctx.synchronize()
    assert (d_status.cpu().numpy() == want_local[1]).all()
    if peer:
        st = ctx.comm_peer_status()
        assert st["steps"] == 7 and not st["timed_out"], st
    np.save(os.path.join(tmp, f"bitmap_{rank}.npy"), torch.stack(gb).cpu().numpy())
    # host-pointer form
    ctx.set_flags(0)
    hb = np.zeros(words, np.uint64)
    hs = np.zeros(hi - lo, np.uint8)
    ctx.verify_proofs_sharded(hi - lo, n, w[0], w[1], w[2], w[3], w[4], hi - lo, hb, hs)
    assert (hs == want_local[1]).all()
    np.save(os.path.join(tmp, f"hbitmap_{rank}.npy"), hb)
    ctx.close()


@pytest.mark.gpu
def test_two_processes_gather_the_accept_bitmap(tmp_path, oracle):
    """one process per GPU (the bench.py layout): id from rank 0 through a file, device-pointer calls overlapping on the comm
    stream, then the host-pointer form; every rank must hold the bitmap a single process computes over the whole batch"""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    n, world = 100_000, 2
    _check_two_process_gather(tmp_path, oracle, n, world, peer=False)


@pytest.mark.gpu
def test_two_processes_gather_through_peer_memory(tmp_path, oracle):
    """the same, with the peer transport: the walk kernel's epilogue stores the ballot words into the other process's bitmap
    through a cudaIpc mapping and publishes the step; no collective launch.  n is chosen so that both shards are equal and
    64-aligned (the condition for the peer path); the host-pointer form still goes through NCCL."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _check_two_process_gather(tmp_path, oracle, 128 * 1024, 2, peer=True)


def _check_two_process_gather(tmp_path, oracle, n, world, peer):
    import torch.multiprocessing as mp
    from phant_b200 import gpu
    mp.spawn(_worker, args=(world, n, str(tmp_path), peer), nprocs=world, join=True)
    o = oracle.synth_c2(n, depth=8, threads=4)
    want = np.unpackbits(oracle.verify_proofs(*o, threads=4)[0].view(np.uint8), bitorder="little")[:n]
    per = gpu.sharded_bitmap_words(n, world) // world * 64
    for r in range(world):
        for name in (f"bitmap_{r}.npy", f"hbitmap_{r}.npy"):
            arr = np.load(tmp_path / name)
            for row in arr.reshape(-1, arr.shape[-1]):
                bits = np.unpackbits(row.view(np.uint8), bitorder="little")
                got = np.concatenate([bits[q * per: q * per + (gpu.shard_range(n, q, world)[1] - gpu.shard_range(n, q, world)[0])] for q in range(world)])
                assert (got == want).all(), (r, name)


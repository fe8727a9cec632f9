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


# ---------------------------------------------------------------------------------------------------------------
# state root sharded by top nibble: the host logic over gloo, the per-rank device work stood in for by the oracle
# ---------------------------------------------------------------------------------------------------------------
class OracleCtx:
    """Stands where phant_b200.gpu.Context does in StateDB.root_sharded, computing with the CPU oracle (this is the test of
    the HOST logic: slot ownership, the one all-reduce, the root branch, the lone-slot fallback)."""

    def __init__(self, o):
        self.o = o

    def keccak256_batch(self, data, off, n, out):
        for i in range(n):
            out[i] = np.frombuffer(self.o.keccak256(data[int(off[i]):int(off[i + 1])].tobytes()), np.uint8)

    @staticmethod
    def _dicts(n, addr, nonce, bal, code, coff, skeys, svals, soff):
        out = []
        for i in range(n):
            st = {skeys[32 * j:32 * j + 32].tobytes().hex(): svals[32 * j:32 * j + 32].tobytes().hex() for j in range(int(soff[i]), int(soff[i + 1]))}
            out.append({"address": addr[20 * i:20 * i + 20].tobytes().hex(), "nonce": int(nonce[i]), "balance": bal[32 * i:32 * i + 32].tobytes().hex(),
                        "code": code[int(coff[i]):int(coff[i + 1])].tobytes().hex(), "storage": st})
        return out

    def state_root(self, n, *tables):
        return self.o.state_root(self._dicts(n, *tables))

    def state_subtree_roots(self, n, *tables):
        from helpers import secure_account_items, _rlp_item
        items = secure_account_items(self.o.keccak256, self.o.mptize, self._dicts(n, *tables))
        refs, mask = np.zeros((16, 32), np.uint8), 0
        for v in range(16):
            sub = [(k, x) for k, x in items if k[0] >> 4 == v]
            if not sub:
                continue
            # two dummy keys under other slots force a root BRANCH whose slot v is exactly the subtree's reference
            dummies = [(bytes([(((v + d) % 16) << 4)]) + bytes(31), b"\x01" * 40) for d in (1, 2)]
            t = self.o.trie(sorted(sub + dummies))
            rootnode = t.prove(sub[0][0])[0]
            pos = _rlp_item(rootnode, 0, len(rootnode))[1]  # payload start of the list
            for slot in range(16):
                it = _rlp_item(rootnode, pos, len(rootnode))
                if slot == v:
                    assert it[2] - it[1] == 32
                    refs[v] = np.frombuffer(rootnode[it[1]:it[2]], np.uint8)
                pos = it[2]
            mask |= 1 << v
        return refs, mask


def _random_statedb(rng, n, lone_nibble_of=None):
    from phant_b200 import host
    db = host.StateDB()
    o = lone_nibble_of
    while len(db.db) < n:
        addr = rng.integers(0, 256, 20, dtype=np.uint8).tobytes()
        if o is not None and o.keccak256(addr)[0] >> 4 != 7:
            continue
        st = {int(rng.integers(1, 1 << 62)): int(rng.choice([0, 1, 255, 1 << 200])) for _ in range(int(rng.choice([0, 0, 1, 3])))}
        db.db[addr] = host.AccountState(int(rng.choice([0, 1, 300])), int(rng.choice([0, 5, 10 ** 18])),
                                        rng.integers(0, 256, int(rng.choice([0, 0, 40])), dtype=np.uint8).tobytes(), st)
    return db


def _state_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    o = oracle_lib.get()
    ctx = OracleCtx(o)
    roots = []
    for case, (n, lone) in enumerate([(0, False), (1, False), (2, False), (40, False), (300, False), (3, True)]):
        db = _random_statedb(np.random.default_rng(100 + case), n, o if lone else None)  # every rank: the same StateDB
        roots.append(db.root_sharded(ctx, rank, world))
    with open(os.path.join(out_dir, f"roots_{rank}.txt"), "w") as f:
        f.write("\n".join(r.hex() for r in roots))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_state_root(tmp_path, oracle):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_state_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ctx = OracleCtx(oracle)
    want = []
    for case, (n, lone) in enumerate([(0, False), (1, False), (2, False), (40, False), (300, False), (3, True)]):
        db = _random_statedb(np.random.default_rng(100 + case), n, oracle if lone else None)
        want.append(db.root(ctx).hex())
    assert want[0] == "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"
    for r in range(world):
        assert open(tmp_path / f"roots_{r}.txt").read().split("\n") == want, r


def test_nibble_ownership_and_root_branch():
    from phant_b200 import shard
    for world in (1, 2, 3, 4, 8, 16, 32):
        owners = [shard.nibble_owner(v, world) for v in range(16)]
        assert owners == sorted(owners) and owners[0] == 0 and max(owners) == min(world, 16) - 1
        assert all(owners.count(r) >= 16 // min(world, 16) for r in range(min(world, 16)))
    refs = np.arange(512, dtype=np.uint8).reshape(16, 32)
    full = shard.root_branch_rlp(refs, 0xffff)
    assert len(full) == 532 and full[:3] == bytes([0xf9, 0x02, 0x11]) and full[3] == 0xa0 and full[-1] == 0x80
    two = shard.root_branch_rlp(refs, 0b101)
    assert len(two) == 2 + 2 * 33 + 15 and two[0] == 0xf8 and two[1] == 81

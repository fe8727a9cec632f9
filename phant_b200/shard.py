"""Multi-GPU host logic: proofs shard by contiguous index range, one process per GPU,
no data-path exchange; ONE all-reduce per batch assembles the global accept bitmap on every rank.

Range boundaries are multiples of 64 so that accept-bitmap words are disjoint between ranks, which makes
SUM over int64 words equal to bitwise OR (NCCL has no bitwise reduction).  Pure host logic: works with
torch.distributed over NCCL (GPU) or gloo (CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_proofs, rank, world):
    """[lo, hi) of proofs for `rank`; every boundary except the last is a multiple of 64."""
    per = ((n_proofs + world - 1) // world + 63) // 64 * 64
    lo = min(rank * per, n_proofs)
    hi = min(lo + per, n_proofs)
    return lo, hi


def bitmap_words(n_proofs):
    return (n_proofs + 63) // 64


def allreduce_accept_bitmap(local_words, lo, n_proofs, group=None):
    """local_words: int64 tensor with the accept words of proofs [lo, lo + 64*len); returns the global bitmap
    (int64 tensor of bitmap_words(n_proofs)) on every rank after a single all-reduce."""
    assert lo % 64 == 0
    g = torch.zeros(bitmap_words(n_proofs), dtype=torch.int64, device=local_words.device)
    g[lo // 64: lo // 64 + local_words.numel()] = local_words
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)  # disjoint words: SUM == OR
    return g


def block_reject_counts(status, block_of_proof, n_blocks, group=None):
    """per-block reject counts (config C5: per-block accept = AND over its proofs), one all-reduce."""
    rej = torch.zeros(n_blocks, dtype=torch.int64, device=status.device)
    rej.index_add_(0, block_of_proof.to(torch.int64), (status == 0).to(torch.int64))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rej, op=dist.ReduceOp.SUM, group=group)
    return rej


# ---- state root sharded by the top nibble of keccak(address): 16 independent subtrees under the root branch ----
def nibble_owner(v, world):
    """rank that builds the subtree under root-branch slot v: contiguous slot ranges, ranks beyond 16 stay idle"""
    return v * min(world, 16) // 16


def _sum_over_ranks(t, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
    return t


def allgather_subtree_roots(refs, mask, group=None):
    """refs: (16, 32) uint8 with zero rows for slots this rank does not own; mask: its populated slots.  One all-reduce
    (SUM over disjoint slots == gather) of 16 x 32 bytes + 16 presence flags; returns (refs_all, mask_all) on every rank."""
    t = torch.zeros(16 * 33, dtype=torch.int32)
    t[:512] = torch.from_numpy(np.ascontiguousarray(refs, np.uint8).reshape(-1).astype(np.int32))
    t[512:] = torch.tensor([(mask >> v) & 1 for v in range(16)], dtype=torch.int32)
    t = _sum_over_ranks(t, group)
    flags = t[512:].tolist()
    assert all(f in (0, 1) for f in flags), "two ranks claimed the same root-branch slot"
    return t[:512].numpy().astype(np.uint8).reshape(16, 32), sum(1 << v for v in range(16) if flags[v])


def sum_bytes(b, group=None):
    """all-reduce of a byte string that is non-zero on one rank only"""
    t = torch.from_numpy(np.ascontiguousarray(b, np.uint8).astype(np.int32))
    return _sum_over_ranks(t, group).numpy().astype(np.uint8)


def root_branch_rlp(refs_all, mask_all):
    """rlp([ref_0 .. ref_15, ""]) of the account trie's root branch (src/mpt/mpt.zig:218-247): populated slots hold the
    32-byte subtree hash, the others and the value slot the empty string"""
    body = b"".join((b"\xa0" + bytes(refs_all[v])) if (mask_all >> v) & 1 else b"\x80" for v in range(16)) + b"\x80"
    n = len(body)
    if n < 56:
        return bytes([0xc0 + n]) + body
    ln = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([0xf7 + len(ln)]) + ln + body

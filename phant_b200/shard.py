"""Multi-GPU host logic: proofs shard by contiguous index range, one process per GPU,
no data-path exchange; ONE all-reduce per batch assembles the global accept bitmap on every rank.

Range boundaries are multiples of 64 so that accept-bitmap words are disjoint between ranks, which makes
SUM over int64 words equal to bitwise OR (NCCL has no bitwise reduction).  Pure host logic: works with
torch.distributed over NCCL (GPU) or gloo (CPU tests).
"""
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

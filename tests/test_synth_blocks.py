"""The device-side C5 generator (phant_b200/synth_blocks.py) checked on the CPU: torch lays the bytes out on the CPU and a
stand-in context hashes with the oracle; the resulting deduplicated witness must verify under the oracle walk, exactly the
corrupted block's first proof must be rejected, and the bytes of a block must not depend on how the block range is cut."""
import numpy as np
import torch

from phant_b200 import synth_blocks


class OracleHashCtx:
    """stands where phant_b200.gpu.Context does: only keccak256_batch / flags are used by the generator"""
    flags = 0

    def __init__(self, o):
        self.o = o

    def set_flags(self, f):
        self.flags = f

    def keccak256_batch(self, msgs, off, n, out):
        m = msgs.numpy()
        o = off.numpy().astype(np.uint64)
        out.copy_(torch.from_numpy(self.o.keccak256_batch(m, o, threads=4)))


def _np(w):
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in w.items()}


def test_generated_blocks_verify_and_only_the_bad_block_is_refused(oracle):
    txs, first, nb = 24, 36, 3  # blocks 36, 37 (corrupted), 38
    w = _np(synth_blocks.synth_blocks(OracleHashCtx(oracle), "cpu", first, nb, txs=txs))
    n = w["n_proofs"]
    assert n == 4 * txs * nb and w["node_off"][-1] == w["n_bytes"] and len(w["node_off"]) == w["n_nodes"] + 1
    assert w["proof_first"][-1] == w["n_refs"] == len(w["node_index"])
    bitmap, status, voff, vlen = oracle.verify_proofs(w["nodes"], w["node_off"].astype(np.uint64), w["proof_first"].astype(np.uint64), w["keys32"],
                                                      w["roots32"], threads=4, node_index=w["node_index"].astype(np.uint64))
    expect = np.ones(n, np.uint8)
    expect[4 * txs * 1] = 0  # first sender's account proof of block 37
    assert (status == expect).all()
    assert (w["block_of_proof"] == np.repeat(np.arange(first, first + nb), 4 * txs)).all()
    # account proofs have 8 nodes, storage proofs 6; values: the 78-byte account body / the 33-byte slot value
    lens = np.diff(w["proof_first"])
    assert (lens.reshape(-1, 4) == [8, 8, 6, 6]).all()
    assert set(vlen[status == 1].tolist()) == {78, 33}
    # shared top levels are stored once: far fewer distinct nodes than references
    assert w["n_nodes"] < w["n_refs"]


def test_block_bytes_do_not_depend_on_the_sharding(oracle):
    ctx = OracleHashCtx(oracle)
    whole = _np(synth_blocks.synth_blocks(ctx, "cpu", 0, 4, txs=10))
    halves = [_np(synth_blocks.synth_blocks(ctx, "cpu", b, 2, txs=10)) for b in (0, 2)]
    assert (whole["keys32"] == np.concatenate([h["keys32"] for h in halves])).all()
    assert (whole["roots32"] == np.concatenate([h["roots32"] for h in halves])).all()

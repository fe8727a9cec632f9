"""Synthetic block witnesses of BASELINE.json config C5 ("1000 synthetic blocks x 300 tx each, full witness verify"),
built ON THE DEVICE: torch lays the bytes out, the library's own batched Keccak (entry point K) hashes every level.

Shape: per block a VIRTUAL state trie -- only the paths of the touched accounts exist, every
sibling off those paths is a counter-based PRF hash of (seed, trie, path) -- so the nodes shared between proofs (the root,
the top levels) are byte-identical and the witness is DEDUPLICATED: `nodes` holds each distinct node once and proof p is
the list node_index[proof_first[p] .. proof_first[p+1]).  Per transaction: two account proofs (sender A, contract B:
7 full 532-byte branches + the 112-byte account leaf, as src/mpt/mpt.zig:218-281 encodes them) under the block's state
root, and two storage proofs (5 branches + a 67-byte leaf) inside B's storage trie, whose root is the storageRoot field
of B's account leaf.  Blocks with index % 100 == 37 carry one corrupted node (one flipped bit in the first sender's
account leaf): that block must be refused.  A block's bytes depend on (seed, block index) only, so any sharding of the
block range over ranks yields the same witnesses.

This is workload generation (setup, untimed); the verifier never sees anything but the CSR arrays it returns.
"""
import torch

from . import gpu

ACC_DEPTH, STO_DEPTH = 7, 5  # branch levels; a proof has depth + 1 nodes
_M64 = (1 << 64) - 1


def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >> 63 else x


_C1, _C2, _C3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)
_K1, _K2 = _s64(0xA24BAED4963EE407), _s64(0xD1342543DE82EF95)


def _lsr(z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _prf_words(seed, ids, stream, n_words):
    """splitmix64 streams keyed by (seed, id, stream): [n] int64 -> [n, n_words] int64 (all arithmetic wraps mod 2^64)"""
    s = (ids * _K1) ^ _s64(seed ^ ((stream * 0xD1342543DE82EF95) & _M64))
    out = []
    for _ in range(n_words + 1):
        s = s + _C1
        z = s
        z = (z ^ _lsr(z, 30)) * _C2
        z = (z ^ _lsr(z, 27)) * _C3
        out.append(z ^ _lsr(z, 31))
    return torch.stack(out[1:], dim=1)  # the first output only decorrelates neighbouring ids


def _prf_bytes(seed, ids, stream, n_bytes):
    w = _prf_words(seed, ids, stream, (n_bytes + 7) // 8).contiguous()
    return w.view(torch.uint8)[:, :n_bytes]


def _hash_rows(ctx, rows):
    """keccak256 of every row of a [n, L] uint8 tensor through the library (device pointers)"""
    n, L = rows.shape
    dev = rows.device
    flat = torch.zeros(n * L + 64, dtype=torch.uint8, device=dev)
    flat[: n * L] = rows.reshape(-1)
    off = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    out = torch.empty(n, 32, dtype=torch.uint8, device=dev)
    # the library works on the CONTEXT's stream (a non-blocking one unless the caller set another): torch's kernels that
    # filled `flat` / `off` must have finished before it reads them, and it must have finished before torch reads `out`
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    ctx.keccak256_batch(flat, off, n, out)
    if dev.type == "cuda":
        ctx.synchronize()
    return out


def _prefix(keys, n_nibbles):
    """first n nibbles of each 32-byte key as int64"""
    p = torch.zeros(keys.shape[0], dtype=torch.int64, device=keys.device)
    for i in range(n_nibbles):
        b = keys[:, i >> 1].to(torch.int64)
        p = p * 16 + ((b & 15) if (i & 1) else (b >> 4))
    return p


def _build(ctx, seed, trie_id, keys, leaves, depth):
    """Union tries over (trie_id, key): returns (levels, chain, trie_ids_sorted, roots) where levels[l] is the [n_l, 532]
    byte tensor of the distinct branch nodes of level l (l < depth), levels[depth] the leaves (one per key), chain[:, l] the
    index of key k's node inside levels[l]."""
    dev = keys.device
    n = keys.shape[0]
    child_uid = trie_id * (16 ** depth) + _prefix(keys, depth)
    child_hash = _hash_rows(ctx, leaves)
    anc = torch.arange(n, dtype=torch.int64, device=dev)
    levels = [None] * (depth + 1)
    chain = torch.empty(n, depth + 1, dtype=torch.int64, device=dev)
    levels[depth] = leaves
    chain[:, depth] = anc
    for lvl in range(depth - 1, -1, -1):
        uid, inv = torch.unique(child_uid >> 4, return_inverse=True)
        slot = child_uid & 15
        m = uid.shape[0]
        node = torch.empty(m, 532, dtype=torch.uint8, device=dev)
        node[:, 0], node[:, 1], node[:, 2], node[:, 531] = 0xF9, 0x02, 0x11, 0x80
        sib = _prf_bytes(seed, (uid.unsqueeze(1) * 16 + torch.arange(16, device=dev)).reshape(-1), 0x51B + lvl, 32).reshape(m, 16, 32)
        body = node[:, 3:531].view(m, 16, 33)
        body[:, :, 0] = 0xA0
        body[:, :, 1:] = sib
        body[inv, slot, 1:] = child_hash  # the touched children replace their PRF placeholders
        levels[lvl] = node
        anc = inv[anc]
        chain[:, lvl] = anc
        child_uid, child_hash = uid, _hash_rows(ctx, node)
    return levels, chain, child_uid, child_hash


def synth_blocks(ctx, device, first_block, n_blocks, txs=300, seed=0x5048414E54):
    """-> dict of device tensors: nodes (uint8, +64 bytes of padding), node_off, node_index, proof_first (int64), keys32, roots32
    (uint8), block_of_proof (int32, GLOBAL block index), and the python ints n_proofs / n_nodes / n_bytes / n_refs"""
    dev = torch.device(device)
    saved = ctx.flags
    ctx.set_flags(gpu.FLAG_DEVICE_PTRS)
    try:
        nb = n_blocks
        blk = torch.arange(first_block, first_block + nb, dtype=torch.int64, device=dev)
        # ---- storage tries: one per (block, tx), two slots each ----
        bt = (blk.unsqueeze(1) * txs + torch.arange(txs, device=dev)).reshape(-1)               # [nb*txs] trie id
        sid = (bt.unsqueeze(1) * 2 + torch.arange(2, device=dev)).reshape(-1)                    # [nb*txs*2] slot id
        skeys = _prf_bytes(seed, sid, 1, 32).clone()
        pre = _prf_words(seed, bt, 2, 2)                                                         # 5 random nibbles per slot ...
        p0, p1 = _lsr(pre[:, 0], 11) & 0xFFFFF, _lsr(pre[:, 1], 11) & 0xFFFFF
        p1 = torch.where(p1 == p0, p0 ^ 1, p1)                                                   # ... distinct inside the pair
        pref = torch.stack([p0, p1], dim=1).reshape(-1)
        skeys[:, 0] = (pref >> 12).to(torch.uint8)
        skeys[:, 1] = ((pref >> 4) & 0xFF).to(torch.uint8)
        skeys[:, 2] = (((pref & 15) << 4) | (skeys[:, 2].to(torch.int64) & 15)).to(torch.uint8)
        ns = skeys.shape[0]
        sleaf = torch.empty(ns, 67, dtype=torch.uint8, device=dev)
        sleaf[:, 0], sleaf[:, 1], sleaf[:, 2] = 0xF8, 65, 0x80 + 30
        sleaf[:, 3] = 0x30 | (skeys[:, 2] & 15)                                                  # 59 path nibbles: odd -> 0x3n first
        sleaf[:, 4:33] = skeys[:, 3:32]
        sleaf[:, 33], sleaf[:, 34] = 0xA1, 0xA0
        sleaf[:, 35:67] = _prf_bytes(seed, sid, 3, 32)
        sleaf[:, 35] |= 0x80
        s_levels, s_chain, s_tries, s_roots = _build(ctx, seed, bt.repeat_interleave(2), skeys, sleaf, STO_DEPTH)
        assert s_tries.shape[0] == nb * txs and bool((s_tries == bt).all())                     # sorted unique == construction order
        # ---- accounts: A_t = 2t (storage root = PRF), B_t = 2t + 1 (storage root = its trie) ----
        na = 2 * txs
        a_loc = torch.arange(na, dtype=torch.int64, device=dev)
        aid = (blk.unsqueeze(1) * na + a_loc).reshape(-1)
        akeys = _prf_bytes(seed, aid, 4, 32).clone()
        off28 = (blk * 0x632BE5AB) & 0xFFFFFFF
        apre = ((a_loc.unsqueeze(0) * 0x9E3779B1 + off28.unsqueeze(1)) & 0xFFFFFFF).reshape(-1)  # 7 nibbles, distinct per block
        akeys[:, 0] = (apre >> 20).to(torch.uint8)
        akeys[:, 1] = ((apre >> 12) & 0xFF).to(torch.uint8)
        akeys[:, 2] = ((apre >> 4) & 0xFF).to(torch.uint8)
        akeys[:, 3] = (((apre & 15) << 4) | (akeys[:, 3].to(torch.int64) & 15)).to(torch.uint8)
        n_acc = akeys.shape[0]
        aleaf = torch.empty(n_acc, 112, dtype=torch.uint8, device=dev)
        aleaf[:, 0], aleaf[:, 1], aleaf[:, 2] = 0xF8, 110, 0x80 + 29
        aleaf[:, 3] = 0x30 | (akeys[:, 3] & 15)                                                  # 57 path nibbles
        aleaf[:, 4:32] = akeys[:, 4:32]
        aleaf[:, 32], aleaf[:, 33], aleaf[:, 34], aleaf[:, 35] = 0xB8, 78, 0xF8, 76
        misc = _prf_words(seed, aid, 5, 2)
        aleaf[:, 36] = (1 + (_lsr(misc[:, 0], 8) % 127)).to(torch.uint8)                         # nonce 1..127
        aleaf[:, 37] = 0x88
        aleaf[:, 38:46] = misc[:, 1:2].contiguous().view(torch.uint8)
        aleaf[:, 38] |= 0x80                                                                     # balance: 8 bytes, top byte non-zero
        aleaf[:, 46] = 0xA0
        sroot = _prf_bytes(seed, aid, 6, 32).clone().view(nb, na, 32)
        sroot[:, 1::2, :] = s_roots.view(nb, txs, 32)
        aleaf[:, 47:79] = sroot.view(n_acc, 32)
        aleaf[:, 79] = 0xA0
        aleaf[:, 80:112] = _prf_bytes(seed, aid, 7, 32)
        a_levels, a_chain, a_tries, a_roots = _build(ctx, seed, blk.repeat_interleave(na), akeys, aleaf, ACC_DEPTH)
        assert bool((a_tries == blk).all())
        # ---- one arena: [storage levels 0..4, storage leaves, account levels 0..6, account leaves] ----
        parts = s_levels + a_levels
        base, node_off, nodes = [], [], []
        n_nodes = n_bytes = 0
        for p in parts:
            m, L = p.shape
            base.append(n_nodes)
            node_off.append(n_bytes + torch.arange(m, dtype=torch.int64, device=dev) * L)
            n_nodes += m
            n_bytes += m * L
        arena = torch.zeros(n_bytes + 64, dtype=torch.uint8, device=dev)
        o = 0
        for p in parts:
            arena[o:o + p.numel()] = p.reshape(-1)
            o += p.numel()
        # corruption AFTER the tries are built: one bit of the first sender's account leaf of every bad block
        bad = (blk % 100 == 37).nonzero().flatten() * na
        arena[node_off[-1][bad] + 50] ^= 1
        node_off.append(torch.tensor([n_bytes], dtype=torch.int64, device=dev))
        node_off = torch.cat(node_off)
        s_idx = s_chain + torch.tensor(base[:STO_DEPTH + 1], dtype=torch.int64, device=dev)
        a_idx = a_chain + torch.tensor(base[STO_DEPTH + 1:], dtype=torch.int64, device=dev)
        # ---- proofs in transaction order: acct(A), acct(B), slot(B, 0), slot(B, 1) ----
        a_idx = a_idx.view(nb * txs, 2 * (ACC_DEPTH + 1))
        s_idx = s_idx.view(nb * txs, 2 * (STO_DEPTH + 1))
        node_index = torch.cat([a_idx, s_idx], dim=1).reshape(-1).contiguous()
        per_tx = 2 * (ACC_DEPTH + 1) + 2 * (STO_DEPTH + 1)
        starts = torch.tensor([0, ACC_DEPTH + 1, 2 * (ACC_DEPTH + 1), 2 * (ACC_DEPTH + 1) + STO_DEPTH + 1], dtype=torch.int64, device=dev)
        n_tx = nb * txs
        proof_first = (torch.arange(n_tx, dtype=torch.int64, device=dev).unsqueeze(1) * per_tx + starts).reshape(-1)
        proof_first = torch.cat([proof_first, torch.tensor([n_tx * per_tx], dtype=torch.int64, device=dev)])
        keys32 = torch.cat([akeys.view(n_tx, 2, 32), skeys.view(n_tx, 2, 32)], dim=1).reshape(-1).contiguous()
        state_root = a_roots.repeat_interleave(txs, dim=0)                                       # [n_tx, 32]
        tx_sroot = s_roots                                                                       # [n_tx, 32]
        roots32 = torch.stack([state_root, state_root, tx_sroot, tx_sroot], dim=1).reshape(-1).contiguous()
        block_of_proof = blk.repeat_interleave(4 * txs).to(torch.int32)
        return {"nodes": arena, "node_off": node_off, "node_index": node_index, "proof_first": proof_first, "keys32": keys32,
                "roots32": roots32, "block_of_proof": block_of_proof, "n_proofs": 4 * n_tx, "n_nodes": n_nodes, "n_bytes": n_bytes,
                "n_refs": n_tx * per_tx}
    finally:
        ctx.set_flags(saved)

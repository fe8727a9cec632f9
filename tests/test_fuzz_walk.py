"""Fuzzing the proof walk: thousands of structurally damaged proofs, three implementations that must agree.

  CPU   oracle (C) vs the independent Python statement (a few hundred cases)
  GPU   CUDA walk vs oracle (tens of thousands of cases, chains and node sets)

Damage is aimed at what the walk parses: RLP headers, lengths, hex-prefix flags, item counts, child references,
plus plain bit flips, truncation, duplication and reordering of nodes.
"""
import numpy as np
import pytest

import oracle_lib
from helpers import py_verify
from test_oracle_proofs import batch_of


def base_proofs(oracle, rng, n_keys=96):
    """genuine proofs from tries with hashed and embedded nodes, extensions and short/long values"""
    out = []
    # secure-style trie: 32-byte random keys, mixed value sizes
    keys = sorted(rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(n_keys))
    kv = [(k, rng.integers(0, 256, int(rng.choice([1, 20, 33, 60, 120])), dtype=np.uint8).tobytes()) for k in keys]
    t = oracle.trie(kv)
    out += [(t.prove(k), k, t.root()) for k in keys[:40]]
    out += [(t.prove(k), k, t.root()) for k in (rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(10))]
    # long shared prefixes: extensions + embedded leaves
    base = rng.integers(0, 256, 30, dtype=np.uint8).tobytes()
    kv2 = sorted({base + bytes([a, b]): bytes([v]) for a, b, v in rng.integers(0, 256, (40, 3))}.items())
    t2 = oracle.trie(kv2)
    out += [(t2.prove(k), k, t2.root()) for k, _ in kv2[:25]]
    out += [(t2.prove(base + b"\x00\x00"), base + b"\x00\x00", t2.root())]
    return out


def damage(proof, rng):
    nl, key, root = proof
    nl = [bytearray(n) for n in nl]
    kind = int(rng.integers(0, 12))
    if not nl:
        return ([], key, bytes(rng.integers(0, 256, 32, dtype=np.uint8)))
    j = int(rng.integers(0, len(nl)))
    n = nl[j]
    if kind == 0:      # random bit flips
        for _ in range(int(rng.integers(1, 4))):
            b = int(rng.integers(0, 8 * len(n)))
            n[b >> 3] ^= 1 << (b & 7)
    elif kind == 1:    # overwrite the list header
        n[0] = int(rng.choice([0x00, 0x7f, 0x80, 0xb7, 0xb8, 0xbf, 0xc0, 0xc1, 0xf7, 0xf8, 0xf9, 0xfa, 0xff]))
    elif kind == 2:    # tweak a length byte
        if len(n) > 2:
            n[1 + int(rng.integers(0, min(3, len(n) - 1)))] ^= int(rng.integers(1, 256))
    elif kind == 3:    # truncate / extend
        if rng.random() < 0.5 and len(n) > 1:
            del n[int(rng.integers(1, len(n))):]
        else:
            n += bytes(rng.integers(0, 256, int(rng.integers(1, 5)), dtype=np.uint8))
    elif kind == 4:    # overwrite some item marker (0x80 / 0xa0 positions are frequent in branches)
        pos = int(rng.integers(0, len(n)))
        n[pos] = int(rng.choice([0x80, 0xa0, 0xa1, 0x9f, 0xc0, 0xc1, 0x81, 0x00]))
    elif kind == 5:    # delete a byte / insert a byte
        pos = int(rng.integers(0, len(n)))
        if rng.random() < 0.5 and len(n) > 1:
            del n[pos]
        else:
            n.insert(pos, int(rng.integers(0, 256)))
    elif kind == 6:    # drop a node
        del nl[j]
    elif kind == 7:    # duplicate a node
        nl.insert(j, bytearray(nl[j]))
    elif kind == 8:    # swap two nodes
        k2 = int(rng.integers(0, len(nl)))
        nl[j], nl[k2] = nl[k2], nl[j]
    elif kind == 9:    # another key
        key = bytes(rng.integers(0, 256, 32, dtype=np.uint8)) if rng.random() < 0.5 else bytes([key[0] ^ (1 << int(rng.integers(0, 8)))]) + key[1:]
    elif kind == 10:   # another root
        root = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
    else:              # leave it intact
        pass
    return ([bytes(x) for x in nl], key, root)


def test_fuzz_oracle_vs_python(oracle):
    rng = np.random.default_rng(2024)
    base = base_proofs(oracle, rng)
    cases = [damage(base[int(rng.integers(0, len(base)))], rng) for _ in range(2500)]
    nodes, node_off, first, keys, roots = batch_of(cases)
    bitmap, status, voff, vlen = oracle.verify_proofs(nodes, node_off, first, keys, roots)
    seen = set()
    for i, (nl, key, root) in enumerate(cases):
        st, val = py_verify(oracle.keccak256, nl, key, root)
        assert st == status[i], (i, st, int(status[i]))
        if st == 1:
            assert nodes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() == val
        seen.add(st)
    assert seen == {0, 1, 2}


@pytest.mark.gpu
def test_fuzz_gpu_vs_oracle(oracle):
    from phant_b200 import gpu
    ctx = gpu.Context(0)
    rng = np.random.default_rng(77)
    base = base_proofs(oracle, rng)
    cases = [damage(base[int(rng.integers(0, len(base)))], rng) for _ in range(30000)]
    nodes, node_off, first, keys, roots = batch_of(cases)
    want = oracle.verify_proofs(nodes, node_off, first, keys, roots, threads=8)
    n = len(cases)
    for flags in (0, gpu.FLAG_KECCAK_DIRECT):
        bitmap = np.zeros((n + 63) // 64, np.uint64)
        status = np.full(n, 77, np.uint8)
        voff = np.zeros(n, np.uint64)
        vlen = np.zeros(n, np.uint32)
        ctx.set_flags(flags)
        ctx.verify_proofs(n, nodes, node_off, first, np.ascontiguousarray(keys), np.ascontiguousarray(roots), n, bitmap, status, voff, vlen)
        bad = np.nonzero(status != want[1])[0]
        assert bad.size == 0, (flags, bad[:10], status[bad[:10]], want[1][bad[:10]])
        assert (bitmap == want[0]).all()
        ok = status == 1
        assert (voff[ok] == want[2][ok]).all() and (vlen[ok] == want[3][ok]).all()
    assert set(np.unique(want[1]).tolist()) == {0, 1, 2}
    # the same damaged nodes as an unordered set: GPU vs oracle in bag mode (every node of every case in one bag per root
    # would mix tries, so take the cases of the first trie only and their common root)
    root0 = base[0][2]
    sel = [c for c in cases if c[2] == root0][:8000]
    bag = list({nd: 1 for c in sel for nd in c[0]})
    bnodes, boff = oracle_lib.csr(bag, np.uint64)
    bkeys = np.frombuffer(b"".join(c[1] for c in sel), np.uint8)
    broot = np.frombuffer(root0, np.uint8)
    want_bag = oracle.verify_bag(bnodes, boff, bkeys, broot, threads=8)
    st = np.zeros(len(sel), np.uint8)
    ctx.set_flags(0)
    ctx.verify_witness(len(bag), bnodes, boff, len(sel), bkeys, broot, 1, None, st, None, None)
    assert (st == want_bag[0]).all(), np.nonzero(st != want_bag[0])[0][:10]
    ctx.close()

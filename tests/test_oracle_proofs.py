"""Proof walk: oracle (C) vs an independent Python statement, on proofs cut from tries whose roots the
fixtures pin, plus deterministic mutations and the synthetic C2/C3 witnesses.  CPU-only.

Parity for accept/reject is UNPINNED by the reference (it has no verifier); this is the strongest
anchor available: the roots are reference-pinned, the walk is stated twice.
"""
import numpy as np
import pytest

import oracle_lib
from helpers import py_verify, secure_account_items


def batch_of(proofs):
    """[(nodes list, key32, root)] -> arrays for oracle.verify_proofs"""
    flat = [n for p in proofs for n in p[0]]
    nodes, node_off = oracle_lib.csr(flat, np.uint64)
    first = np.zeros(len(proofs) + 1, np.uint64)
    first[1:] = np.cumsum([len(p[0]) for p in proofs])
    keys = np.frombuffer(b"".join(p[1] for p in proofs), np.uint8)
    roots = np.frombuffer(b"".join(p[2] for p in proofs), np.uint8)
    return nodes, node_off, first, keys, roots


def check_same(oracle, proofs, expect_status=None):
    nodes, node_off, first, keys, roots = batch_of(proofs)
    bitmap, status, voff, vlen = oracle.verify_proofs(nodes, node_off, first, keys, roots)
    for i, (nl, key, root) in enumerate(proofs):
        st, val = py_verify(oracle.keccak256, nl, key, root)
        assert st == status[i], (i, st, status[i])
        assert bool((int(bitmap[i // 64]) >> (i % 64)) & 1) == (st != 0)
        if st == 1:
            assert nodes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() == val
        if expect_status is not None:
            assert st == expect_status[i], (i, st, expect_status[i])
    return status


def mutations(nl, key, root, rng):
    """deterministic corruptions of one proof -> list of (nodes, key, root)"""
    out = []
    j = int(rng.integers(0, len(nl)))
    n = bytearray(nl[j])
    bit = int(rng.integers(0, 8 * len(n)))
    n[bit >> 3] ^= 1 << (bit & 7)
    out.append((nl[:j] + [bytes(n)] + nl[j + 1:], key, root))                      # bit flip
    out.append((nl[:j] + [nl[j][:-1]] + nl[j + 1:], key, root))                    # truncated node
    out.append((nl[:j] + [nl[j] + b"\x00"] + nl[j + 1:], key, root))               # trailing byte
    if len(nl) > 1:
        out.append((nl[:-1], key, root))                                           # truncated chain
        out.append((nl[1:], key, root))                                            # missing root node
        out.append(([nl[1], nl[0]] + nl[2:], key, root))                           # swapped nodes
    out.append((nl + [nl[-1]], key, root))                                         # trailing node
    out.append((nl, key, bytes([root[0] ^ 1]) + root[1:]))                         # wrong root
    return out


def test_fixture_account_proofs(oracle, golden):
    g = golden("fixture_states.json.gz")
    rng = np.random.default_rng(3)
    done = 0
    for tkey, accounts in sorted(g["tables"].items()):
        if not accounts or (len(accounts) > 40 and done > 6):
            continue
        items = secure_account_items(oracle.keccak256, oracle.mptize, accounts)
        trie = oracle.trie(items)
        root = trie.root()
        proofs, expect = [], []
        for k, v in items:                                   # inclusion: every account
            proofs.append((trie.prove(k), k, root))
            expect.append(1)
        for _ in range(8):                                   # exclusion: random absent keys
            k = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
            proofs.append((trie.prove(k), k, root))
            expect.append(2)
        status = check_same(oracle, proofs, expect)
        # values are the account RLP
        nodes, node_off, first, keys, roots = batch_of(proofs)
        _, _, voff, vlen = oracle.verify_proofs(nodes, node_off, first, keys, roots)
        for i, (k, v) in enumerate(items):
            assert nodes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() == v
        # mutations of a sample must be rejected -- except a wrong key, which flips presence
        bad = []
        for i in rng.choice(len(items), size=min(6, len(items)), replace=False):
            bad += mutations(*proofs[int(i)], rng)
        st = check_same(oracle, bad)
        assert (st == 0).all()
        done += 1
    assert done >= 8


def test_fixture_storage_proofs(oracle, golden):
    from helpers import rlp_int_be
    g = golden("fixture_states.json.gz")
    n = 0
    for tkey, accounts in sorted(g["tables"].items()):
        for a in accounts:
            st = sorted((oracle.keccak256(bytes.fromhex(k)), rlp_int_be(bytes.fromhex(v))) for k, v in a["storage"].items()
                        if int(v, 16) != 0)
            if not st or n > 40:
                continue
            trie = oracle.trie(st)
            root = trie.root()
            proofs = [(trie.prove(k), k, root) for k, _ in st]
            absent = oracle.keccak256(b"absent" + bytes([n]))
            proofs.append((trie.prove(absent), absent, root))
            check_same(oracle, proofs, [1] * len(st) + [2])
            n += 1
    assert n > 5


def test_embedded_nodes(oracle):
    """32-byte keys sharing 62-63 nibbles with 1-byte values: leaves and branches embed (mpt.zig:104,112)."""
    base = bytes(range(31))
    kv = sorted((base + bytes([b]), bytes([v])) for b, v in [(0x10, 1), (0x11, 2), (0x1f, 3), (0x20, 4), (0x77, 5)])
    trie = oracle.trie(kv)
    n_nodes, n_hashed, _ = trie.stats()
    assert n_hashed < n_nodes                                   # something is embedded
    root = trie.root()
    proofs = [(trie.prove(k), k, root) for k, _ in kv]
    absent = [base + bytes([0x12]), base + bytes([0x30]), bytes([0xff]) + base, base[:30] + b"\xee\x10"]
    proofs += [(trie.prove(k), k, root) for k in absent]
    st = check_same(oracle, proofs, [1] * len(kv) + [2] * len(absent))
    rng = np.random.default_rng(5)
    bad = []
    for p in proofs[:5]:
        bad += mutations(*p, rng)
    assert (check_same(oracle, bad) == 0).all()


def test_empty_trie(oracle):
    empty = bytes.fromhex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421")
    k = bytes(32)
    check_same(oracle, [([], k, empty), ([], k, bytes(32))], [2, 0])


def test_wrong_key_is_absent_not_reject(oracle):
    """A valid chain walked with another key diverges: absent (or reject when the chain has unused nodes)."""
    rng = np.random.default_rng(11)
    keys = sorted(rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(64))
    kv = [(k, b"v" * 40) for k in keys]
    trie = oracle.trie(kv)
    root = trie.root()
    proofs = []
    for k in keys[:16]:
        other = bytes([k[0] ^ 0x80]) + k[1:]
        proofs.append((trie.prove(k), other, root))
    st = check_same(oracle, proofs)
    assert set(st.tolist()) <= {0, 2}


@pytest.mark.parametrize("which", ["c2", "c3"])
def test_synthetic_witnesses(oracle, which):
    n = 600
    if which == "c2":
        nodes, node_off, first, keys, roots = oracle.synth_c2(n, depth=8)
        assert int(node_off[-1]) == n * 3836 and (np.diff(first) == 8).all()
    else:
        nodes, node_off, first, keys, roots = oracle.synth_c3(n)
        d = np.diff(first)
        assert d.min() >= 4 and d.max() <= 12
    bitmap, status, voff, vlen = oracle.verify_proofs(nodes, node_off, first, keys, roots, threads=4)
    expect = np.where(np.arange(n) % 97 == 0, 0, 1)
    assert (status == expect).all()
    # python statement agrees on a sample
    for i in list(range(0, n, 97)) + list(range(1, n, 53)):
        nl = [nodes[int(node_off[j]):int(node_off[j + 1])].tobytes() for j in range(int(first[i]), int(first[i + 1]))]
        st, _ = py_verify(oracle.keccak256, nl, keys[32 * i:32 * i + 32].tobytes(), roots[32 * i:32 * i + 32].tobytes())
        assert st == status[i]
    # generation is index-addressable: a later window reproduces the same proofs
    if which == "c2":
        n2 = oracle.synth_c2(100, depth=8, first=200)
        assert (n2[0][:100 * 3836] == nodes[200 * 3836:300 * 3836]).all()
        assert (n2[4] == roots[200 * 32:300 * 32]).all()


def test_ctrie_update_matches_full_rebuild(oracle):
    depth = 3
    t = oracle.ctrie(depth)
    r0 = t.root()
    rng = np.random.default_rng(2)
    keys = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    vals = [rng.integers(0, 256, 70, dtype=np.uint8).tobytes() for _ in range(50)]
    r1 = t.update(keys.reshape(-1), vals)
    assert r1 != r0 and t.root() == r1
    # order independence and idempotence
    t2 = oracle.ctrie(depth)
    perm = rng.permutation(50)
    # distinct leaf positions only (contract of the update); drop collisions
    pos = [(k[0] << 4) | (k[1] >> 4) for k in keys]
    if len(set(pos)) == len(pos):
        assert t2.update(keys[perm].reshape(-1), [vals[i] for i in perm]) == r1
    assert t.update(keys.reshape(-1), vals) == r1


def shuffled_bag(w, rng, drop=None, extra=0):
    """the distinct nodes of a block witness as an unordered bag (optionally without node `drop`, plus junk nodes)"""
    n = w["n_nodes"]
    order = [int(i) for i in rng.permutation(n) if drop is None or int(i) != drop]
    items = [w["nodes"][int(w["node_off"][i]):int(w["node_off"][i + 1])].tobytes() for i in order]
    items += [rng.integers(0, 256, int(rng.integers(1, 600)), dtype=np.uint8).tobytes() for _ in range(extra)]
    return oracle_lib.csr(items, np.uint64)


def test_bag_witness_matches_chain_verdicts(oracle):
    """the same proofs verified as chains and as an unordered set of nodes give the same present/absent verdicts"""
    rng = np.random.default_rng(6)
    w = oracle.synth_blocks(2, txs=40, first=0, threads=4)
    chain = oracle.verify_proofs(w["nodes"], w["node_off"], w["proof_first"], w["keys32"], w["roots32"], threads=4, node_index=w["node_index"])
    nodes, node_off = shuffled_bag(w, rng, extra=25)
    st, voff, vlen = oracle.verify_bag(nodes, node_off, w["keys32"], w["roots32"], threads=4)
    assert (st == chain[1]).all() and (st == 1).all()
    vals_chain = [w["nodes"][int(o):int(o) + int(l)].tobytes() for o, l in zip(chain[2], chain[3])]
    vals_bag = [nodes[int(o):int(o) + int(l)].tobytes() for o, l in zip(voff, vlen)]
    assert vals_chain == vals_bag
    # drop one node: every key whose path crosses it reports "missing" (3), the others are untouched
    victim = int(w["node_index"][int(w["proof_first"][5]) + 2])
    nodes2, node_off2 = shuffled_bag(w, rng, drop=victim)
    st2, _, _ = oracle.verify_bag(nodes2, node_off2, w["keys32"], w["roots32"], threads=4)
    uses = np.array([victim in w["node_index"][int(a):int(b)] for a, b in zip(w["proof_first"][:-1], w["proof_first"][1:])])
    assert (st2[uses] == 3).all() and (st2[~uses] == 1).all() and uses.any()
    # absent keys and the empty root
    k = rng.integers(0, 256, 32, dtype=np.uint8)
    st3, _, _ = oracle.verify_bag(nodes, node_off, k, w["roots32"][:32].copy())
    assert st3[0] in (2, 3)  # a random key leaves the materialised paths: absent if it dies in a known node, missing otherwise
    empty = np.frombuffer(bytes.fromhex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"), np.uint8)
    assert oracle.verify_bag(nodes, node_off, k, empty)[0][0] == 2


def kat_batch(g):
    return [([bytes.fromhex(n) for n in c["nodes"]], bytes.fromhex(c["key"]), bytes.fromhex(c["root"])) for c in g["cases"]]


def test_proof_kat(oracle, golden):
    """the committed walk vectors (tests/golden/proof_kat.json.gz, made by make_proof_kat.py): oracle and Python statement
    both reproduce every recorded status and value"""
    g = golden("proof_kat.json.gz")
    proofs = kat_batch(g)
    nodes, node_off, first, keys, roots = batch_of(proofs)
    bitmap, status, voff, vlen = oracle.verify_proofs(nodes, node_off, first, keys, roots)
    for i, c in enumerate(g["cases"]):
        assert int(status[i]) == c["status"], c["name"]
        st, val = py_verify(oracle.keccak256, *proofs[i])
        assert st == c["status"], c["name"]
        if c["status"] == 1:
            assert val.hex() == c["value"] == nodes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes().hex(), c["name"]
    assert {c["status"] for c in g["cases"]} == {0, 1, 2} and len(g["cases"]) >= 45

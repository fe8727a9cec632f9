#!/usr/bin/env python3
"""Known-answer vectors for the proof walk (P0) -> tests/golden/proof_kat.json.gz.

The reference has no verifier, so these vectors cannot come from it (parity for accept / reject is UNPINNED).
They freeze the semantics this repository implements so that every decision can be audited by hand and so that the CUDA
walk can be checked against committed answers without the oracle:

  * genuine inclusion / exclusion proofs cut from a trie whose ROOT the reference's fixtures pin (the largest fixture
    state, key = keccak(address), value = account RLP) and from a trie with embedded (< 32 byte) nodes and extensions;
  * named structural damage of those proofs (what each one does is in the case's `name`);
  * single malformed nodes presented as their own root (root = keccak(node)), so the hash check passes and only the
    structural rules R2-R4 can refuse them.

An answer is written only when the two independent statements of the walk -- the C oracle (oracle/verify.c) and the
Python one (tests/helpers.py::py_verify) -- agree on it; the script stops otherwise.

  python tests/golden/make_proof_kat.py        (needs the oracle built: python -c "import __graft_entry__ as g; g.build()")
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
from helpers import py_verify, rlp_list, rlp_str, secure_account_items  # noqa: E402
from test_oracle_proofs import batch_of  # noqa: E402


def main():
    o = oracle_lib.get()
    k = o.keccak256
    cases = []

    def add(name, nodes, key, root):
        cases.append({"name": name, "nodes": [bytes(n) for n in nodes], "key": bytes(key), "root": bytes(root)})

    # ---- 1. account trie of the largest fixture state (root pinned by the fixture's header field) ----
    g = json.loads(gzip.open(os.path.join(HERE, "fixture_states.json.gz")).read())
    name, accounts = max(g["tables"].items(), key=lambda kv: len(kv[1]))
    pinned = {t["pre"]: t["pre_root"] for t in g["tests"]} | {t["post"]: t["post_root"] for t in g["tests"]}
    items = secure_account_items(k, o.mptize, accounts)
    trie = o.trie(items)
    assert trie.root().hex() == pinned[name], "the fixture pins this root"
    root = trie.root()
    keys = [kk for kk, _ in items]
    for j in (0, 7, len(keys) // 2, len(keys) - 1):
        add(f"fixture account {j}: inclusion", trie.prove(keys[j]), keys[j], root)
    for label in (b"nobody", b"absent-1", b"absent-2"):
        ak = k(label)
        add(f"fixture trie: exclusion of keccak({label.decode()})", trie.prove(ak), ak, root)
    nl, key = trie.prove(keys[7]), keys[7]
    assert len(nl) >= 2
    leaf = nl[-1]
    add("leaf value: one bit flipped (hash mismatch)", nl[:-1] + [leaf[:-1] + bytes([leaf[-1] ^ 1])], key, root)
    add("last node dropped (chain ends on a hash reference)", nl[:-1], key, root)
    add("root node dropped", nl[1:], key, root)
    add("first two nodes swapped", [nl[1], nl[0]] + nl[2:], key, root)
    add("one node too many (leaf repeated)", nl + [nl[-1]], key, root)
    add("root hash differs in one bit", nl, key, bytes([root[0] ^ 0x80]) + root[1:])
    add("node with a trailing zero byte", nl[:-1] + [leaf + b"\x00"], key, root)
    add("node cut by one byte", nl[:-1] + [leaf[:-1]], key, root)
    other = keys[8]
    add("genuine proof of another key presented for this key", trie.prove(other), key, root)
    add("empty chain against a non-empty root", [], key, root)
    add("empty chain against the empty root (absent)", [], key, bytes.fromhex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"))
    add("exclusion proof with its last node dropped", trie.prove(k(b"nobody"))[:-1], k(b"nobody"), root)

    # ---- 2. embedded nodes and extensions: 30 shared bytes, 2 free, 1-byte values ----
    base = bytes(range(1, 31))
    kv2 = sorted({base + bytes([a, b]): bytes([v | 1]) for a, b, v in np.random.default_rng(11).integers(0, 256, (24, 3))}.items())
    t2 = o.trie(kv2)
    for j in (0, 5, 23 if len(kv2) > 23 else len(kv2) - 1):
        add(f"embedded-leaf trie key {j}: inclusion", t2.prove(kv2[j][0]), kv2[j][0], t2.root())
    miss = base + bytes([kv2[0][0][30], kv2[0][0][31] ^ 0xff])
    if miss not in dict(kv2):
        add("embedded-leaf trie: sibling key absent", t2.prove(miss), miss, t2.root())
    far = bytes(32)
    add("embedded-leaf trie: key diverging inside the root extension (absent)", t2.prove(far), far, t2.root())
    p2 = t2.prove(kv2[5][0])
    add("embedded-leaf trie: extension node dropped", p2[1:], kv2[5][0], t2.root())

    # ---- 3. single malformed nodes as their own root: only the structural rules can refuse them ----
    key0 = bytes(32)
    hp_leaf_all = b"\x20" + key0                     # even leaf flag + 64 zero nibbles
    good_leaf = rlp_list([rlp_str(hp_leaf_all), rlp_str(b"value")])
    def self_rooted(name, node, key=key0):
        add(name, [node], key, k(node))
    self_rooted("single leaf covering the whole key: inclusion", good_leaf)
    self_rooted("single leaf, other key: absent", good_leaf, bytes([0x10]) + bytes(31))
    self_rooted("leaf whose path is one nibble short (63 nibbles): cannot match, absent", rlp_list([rlp_str(b"\x30" + bytes(31)), rlp_str(b"v")]))
    self_rooted("leaf with hex-prefix flag 4", rlp_list([rlp_str(b"\x40" + key0), rlp_str(b"v")]))
    self_rooted("leaf with even flag and a non-zero padding nibble", rlp_list([rlp_str(b"\x21" + key0), rlp_str(b"v")]))
    self_rooted("leaf with an empty hex-prefix string", rlp_list([rlp_str(b""), rlp_str(b"v")]))
    self_rooted("leaf path longer than the key (66 nibbles)", rlp_list([rlp_str(b"\x20" + bytes(33)), rlp_str(b"v")]))
    self_rooted("list header in long form for a short payload (non-canonical)", b"\xf8" + bytes([len(good_leaf) - 1]) + good_leaf[1:])
    self_rooted("list shorter than the node (trailing byte inside the hash)", good_leaf + b"\x00")
    self_rooted("value string with a non-canonical single byte (0x81 0x05)", rlp_list([rlp_str(hp_leaf_all), b"\x81\x05"]))
    self_rooted("three-item list", rlp_list([rlp_str(hp_leaf_all), rlp_str(b"v"), rlp_str(b"w")]))
    self_rooted("leaf whose value is a list", rlp_list([rlp_str(hp_leaf_all), rlp_list([rlp_str(b"v")])]))
    self_rooted("node that is a string, not a list", rlp_str(b"x" * 40))
    empty16 = [rlp_str(b"")] * 16
    self_rooted("branch with no children and a value, key not exhausted: empty slot, absent", rlp_list(empty16 + [rlp_str(b"v")]))
    self_rooted("branch with 16 items", rlp_list(empty16))
    self_rooted("branch with 18 items", rlp_list(empty16 + [rlp_str(b""), rlp_str(b"")]))
    self_rooted("branch whose slot 0 holds a 31-byte string", rlp_list([rlp_str(b"\x01" * 31)] + empty16[1:] + [rlp_str(b"")]))
    self_rooted("branch whose slot 0 holds a 33-byte string", rlp_list([rlp_str(b"\x01" * 33)] + empty16[1:] + [rlp_str(b"")]))
    self_rooted("branch whose slot 0 holds a hash but the chain ends", rlp_list([rlp_str(b"\x01" * 32)] + empty16[1:] + [rlp_str(b"")]))
    emb_leaf = rlp_list([rlp_str(b"\x3f" + bytes(31)), rlp_str(b"v")])  # 63 remaining nibbles, odd flag: 36 bytes -> too big to embed
    self_rooted("branch with an embedded child of 32+ bytes", rlp_list([emb_leaf] + empty16[1:] + [rlp_str(b"")]))
    self_rooted("extension with an empty path", rlp_list([rlp_str(b"\x00"), rlp_str(b"\x01" * 32)]))
    self_rooted("extension whose child is the empty string", rlp_list([rlp_str(b"\x00\x00"), rlp_str(b"")]))
    self_rooted("extension that matches, hash child, chain ends", rlp_list([rlp_str(b"\x00\x00"), rlp_str(b"\x01" * 32)]))
    self_rooted("extension that diverges, last node: absent", rlp_list([rlp_str(b"\x00\x10"), rlp_str(b"\x01" * 32)]))

    # ---- answers: both statements of the walk must agree ----
    batch = [(c["nodes"], c["key"], c["root"]) for c in cases]
    nodes, node_off, first, keys32, roots = batch_of(batch)
    _, status, voff, vlen = o.verify_proofs(nodes, node_off, first, keys32, roots)
    out = []
    for i, c in enumerate(cases):
        st, val = py_verify(k, c["nodes"], c["key"], c["root"])
        assert st == int(status[i]), (c["name"], st, int(status[i]))
        if st == 1:
            assert nodes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() == val
        out.append({"name": c["name"], "key": c["key"].hex(), "root": c["root"].hex(), "nodes": [n.hex() for n in c["nodes"]],
                    "status": st, "value": val.hex() if st == 1 else None})
        print(f"{st}  {c['name']}")
    blob = json.dumps({"source": "tests/golden/make_proof_kat.py (oracle C walk == Python walk; fixture root " + pinned[name] + ")",
                       "status_legend": {"0": "reject", "1": "present", "2": "proven absent"}, "cases": out},
                      separators=(",", ":"), sort_keys=True).encode()
    path = os.path.join(HERE, "proof_kat.json.gz")
    with open(path, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as z:
            z.write(blob)
    print(f"wrote {path}: {len(out)} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()

"""The device sources of the proof walk and of the node summary (phant_b200/csrc/walk_one.cuh, node_summary.cuh) compiled
as HOST code (tests/hostcheck/walk_host.cpp) and fuzzed against the oracle without a GPU: the committed vectors, genuine and
structurally damaged proofs as chains (with and without the summary fast path, plain and deduplicated) and as node sets.
The product never runs this way; the -m gpu tests check the same things through the C ABI on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from test_fuzz_walk import base_proofs, damage
from test_oracle_proofs import batch_of, kat_batch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostwalk(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostwalk") / "libwalkhost.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so,
                    os.path.join(HERE, "hostcheck", "walk_host.cpp")], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def digests_of(oracle, nodes, node_off):
    n = len(node_off) - 1
    out = np.zeros((max(n, 1), 32), np.uint8)
    for j in range(n):
        out[j] = np.frombuffer(oracle.keccak256(nodes[int(node_off[j]):int(node_off[j + 1])].tobytes()), np.uint8)
    return out


def run_chain(lib, oracle, nodes, node_off, first, keys, roots, use_summary, node_index=None):
    n, nn = len(first) - 1, len(node_off) - 1
    dg = digests_of(oracle, nodes, node_off)
    status, voff, vlen = np.full(n, 9, np.uint8), np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    padded = np.concatenate([nodes, np.zeros(64, np.uint8)])  # the device buffers carry the same slack
    lib.hostwalk_chain(_p(padded), _p(node_off), C.c_uint64(nn), _p(node_index), _p(first), C.c_uint64(n), _p(np.ascontiguousarray(keys)),
                       _p(np.ascontiguousarray(roots)), C.c_uint64(n), _p(dg), C.c_int(use_summary), _p(status), _p(voff), _p(vlen))
    return status, voff, vlen


def test_committed_vectors(hostwalk, oracle, golden):
    g = golden("proof_kat.json.gz")
    nodes, node_off, first, keys, roots = batch_of(kat_batch(g))
    for use_summary in (0, 1):
        status, voff, vlen = run_chain(hostwalk, oracle, nodes, node_off, first, keys, roots, use_summary)
        for i, c in enumerate(g["cases"]):
            assert int(status[i]) == c["status"], (use_summary, c["name"])
            if c["status"] == 1:
                assert nodes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes().hex() == c["value"], c["name"]


def test_fuzz_chains_and_sets(hostwalk, oracle):
    rng = np.random.default_rng(515)
    base = base_proofs(oracle, rng)
    cases = [damage(base[int(rng.integers(0, len(base)))], rng) for _ in range(6000)]
    nodes, node_off, first, keys, roots = batch_of(cases)
    want = oracle.verify_proofs(nodes, node_off, first, keys, roots, threads=4)
    for use_summary in (0, 1):
        status, voff, vlen = run_chain(hostwalk, oracle, nodes, node_off, first, keys, roots, use_summary)
        bad = np.nonzero(status != want[1])[0]
        assert bad.size == 0, (use_summary, bad[:10], status[bad[:10]], want[1][bad[:10]])
        ok = status == 1
        assert (voff[ok] == want[2][ok]).all() and (vlen[ok] == want[3][ok]).all()
    assert set(np.unique(want[1]).tolist()) == {0, 1, 2}
    # the same proofs as a deduplicated witness: distinct nodes once, chains of indices
    flat = [nd for c in cases for nd in c[0]]
    uniq = list({nd: 1 for nd in flat})
    where = {nd: j for j, nd in enumerate(uniq)}
    unodes, uoff = oracle_lib.csr(uniq, np.uint64)
    index = np.array([where[nd] for nd in flat] or [0], np.uint64)
    status, voff, vlen = run_chain(hostwalk, oracle, unodes, uoff, first, keys, roots, 1, node_index=index)
    want_d = oracle.verify_proofs(unodes, uoff, first, keys, roots, threads=4, node_index=index)
    assert (status == want_d[1]).all() and (status == want[1]).all()
    # node sets: every case of the first trie in one bag
    root0 = base[0][2]
    sel = [c for c in cases if c[2] == root0][:3000]
    bag = list({nd: 1 for c in sel for nd in c[0]})
    bnodes, boff = oracle_lib.csr(bag, np.uint64)
    bkeys = np.frombuffer(b"".join(c[1] for c in sel), np.uint8)
    broot = np.frombuffer(root0, np.uint8)
    want_bag = oracle.verify_bag(bnodes, boff, bkeys, broot, threads=4)
    dg = digests_of(oracle, bnodes, boff)
    for use_summary in (0, 1):
        st, vo, vl = np.full(len(sel), 9, np.uint8), np.zeros(len(sel), np.uint64), np.zeros(len(sel), np.uint32)
        padded = np.concatenate([bnodes, np.zeros(64, np.uint8)])
        hostwalk.hostwalk_bag(_p(padded), _p(boff), C.c_uint64(len(bag)), C.c_uint64(len(sel)), _p(bkeys), _p(broot), C.c_uint64(1), _p(dg),
                              C.c_int(use_summary), _p(st), _p(vo), _p(vl))
        assert (st == want_bag[0]).all(), (use_summary, np.nonzero(st != want_bag[0])[0][:10])
    assert 3 in set(want_bag[0].tolist())


def test_synthetic_c2_c3(hostwalk, oracle):
    for which in (2, 3):
        w = oracle.synth_c2(3000, depth=8, threads=4) if which == 2 else oracle.synth_c3(3000, threads=4)
        nodes, node_off, first, keys, roots = w
        want = oracle.verify_proofs(nodes, node_off, first, keys, roots, threads=4)
        status, _, _ = run_chain(hostwalk, oracle, nodes, node_off, first, keys, roots, 1)
        assert (status == want[1]).all(), which
        assert (np.nonzero(status == 0)[0] % 97 == 0).all() and (status[::97] == 0).all(), which

"""small invocations of every entry point, for compute-sanitizer (memcheck / racecheck / initcheck)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib
from phant_b200 import gpu
from gpu_util import random_csr

o = oracle_lib.get()
ctx = gpu.Context(0)
rng = np.random.default_rng(0)
data, off = random_csr(rng, 5000)      # >= 4096 messages: the two-launch regrouping (keccak_class / keccak_regroup) runs
out = np.zeros((5000, 32), np.uint8)
for flags in (0, 1 << 4, 1 << 5, 1 << 6):
    ctx.set_flags(flags)
    ctx.keccak256_batch(data, off, 5000, out)
assert (out == o.keccak256_batch(data, off, threads=4)).all()
ctx.set_flags(0)
w = o.synth_c3(3000)
n = 3000
bm = np.zeros((n + 63) // 64, np.uint64); st = np.zeros(n, np.uint8)
ctx.verify_proofs(n, w[0], w[1], w[2], w[3], w[4], n, bm, st, None, None)
assert (st == o.verify_proofs(*w, threads=4)[1]).all()
b = o.synth_blocks(6, txs=20)
n = b["n_proofs"]
bm = np.zeros((n + 63) // 64, np.uint64); st = np.zeros(n, np.uint8)
ctx.verify_proofs(n, b["nodes"], b["node_off"], b["proof_first"], b["keys32"], b["roots32"], n, bm, st, None, None, n_nodes=b["n_nodes"],
                  nodes_bytes=b["n_bytes"], node_index=b["node_index"])
kv = sorted((rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), rng.integers(0, 256, 60, dtype=np.uint8).tobytes()) for _ in range(700))
keys, koff = oracle_lib.csr([k for k, _ in kv], np.uint32); vals, voff = oracle_lib.csr([v for _, v in kv], np.uint64)
assert ctx.mpt_root(keys, koff, vals, voff, len(kv)) == o.mptize(kv)
t = ctx.trie_open(3)
pos = rng.choice(4096, 200, replace=False)
keys = rng.integers(0, 256, (200, 32), dtype=np.uint8); keys[:, 0] = pos >> 4; keys[:, 1] = ((pos & 15) << 4) | (keys[:, 1] & 15)
v, vo = oracle_lib.csr([rng.integers(0, 256, 70, dtype=np.uint8).tobytes() for _ in range(200)], np.uint32)
t.update(np.ascontiguousarray(keys.reshape(-1)), v, vo, 200)
t.close()
# U kind 1: sparse resident trie -- build past the first dense level, value updates, deletes
t = ctx.trie_open(0, kind=1)
state = {rng.integers(0, 256, 32, dtype=np.uint8).tobytes(): rng.integers(0, 256, 50, dtype=np.uint8).tobytes() for _ in range(900)}
def apply(ch):
    ks = list(ch)
    v, vo = oracle_lib.csr([ch[k] for k in ks], np.uint32)
    return t.update(np.frombuffer(b"".join(ks), np.uint8), v, vo, len(ks))
assert apply(state) == o.mptize(sorted(state.items()))
ch = {k: b"x" * 40 for k in list(state)[:50]}
ch.update({k: b"" for k in list(state)[50:90]})
for k, v in ch.items():
    if v: state[k] = v
    else: del state[k]
assert apply(ch) == o.mptize(sorted(state.items()))
t.close()
# W: node-set witness, B: blooms, R: sender recovery
nodes, node_off = oracle_lib.csr([b["nodes"][int(b["node_off"][j]):int(b["node_off"][j + 1])].tobytes() for j in range(b["n_nodes"])], np.uint64)
st = np.zeros(n, np.uint8)
ctx.verify_witness(len(node_off) - 1, nodes, node_off, n, b["keys32"], b["roots32"], n, None, st, None, None)
items, ioff = random_csr(rng, 300, max_len=40, pad_front=0)
blooms = np.zeros((4, 256), np.uint8)
ctx.logs_bloom(items, ioff, (np.arange(300) % 4).astype(np.uint32), 300, 4, blooms)
h = rng.integers(0, 256, 32 * 64, dtype=np.uint8); sg = rng.integers(0, 256, 65 * 64, dtype=np.uint8); sg[64::65] &= 1
ok = np.zeros(64, np.uint8); addr = np.zeros((64, 20), np.uint8)
ctx.ecrecover_batch(h, sg, 64, None, addr, ok)
ctx.close()
print("sanitize workload ok")

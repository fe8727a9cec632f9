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
data, off = random_csr(rng, 5000)
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
ctx.close()
print("sanitize workload ok")

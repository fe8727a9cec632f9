"""R on the GPU: phant_gpu_ecrecover_batch through the C ABI against the reference's vectors and the oracle."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PHANT_GPU_ECRECOVER", "1") != "1", reason="PHANT_GPU_ECRECOVER=0")]

N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


@pytest.fixture(scope="module")
def ctx():
    from phant_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


def run(ctx, hashes, sigs):
    n = len(hashes)
    h = np.frombuffer(b"".join(hashes), np.uint8)
    s = np.frombuffer(b"".join(sigs), np.uint8)
    pub, addr, ok = np.zeros((n, 65), np.uint8), np.zeros((n, 20), np.uint8), np.full(n, 7, np.uint8)
    ctx.ecrecover_batch(h, s, n, pub, addr, ok)
    return pub, addr, ok


def test_reference_vectors(ctx, golden):
    from phant_b200 import host
    g = golden("ecrecover_kat.json")
    k = g["erecover"]
    pub, addr, ok = run(ctx, [bytes.fromhex(k["hash"])], [bytes.fromhex(k["sig65"])])
    assert ok[0] == 1 and pub[0].tobytes().hex() == k["pubkey65"]
    got = host.get_senders(ctx, [bytes.fromhex(t["encoded"]) for t in g["txs"]], chain_id=1)
    assert [a.hex() for a in got] == [t["sender"] for t in g["txs"]]


def test_random_inputs_vs_oracle(ctx, oracle):
    rng = np.random.default_rng(12)
    hashes, sigs = [], []
    for i in range(3000):
        r, s = int.from_bytes(rng.bytes(32), "big"), int.from_bytes(rng.bytes(32), "big")
        if i % 7 == 0:
            r %= 2 ** 127
        if i % 11 == 0:
            s = [0, 1, N - 1, N, 2 ** 256 - 1][i // 11 % 5]
        if i % 13 == 0:
            r = [0, 1, N - 1, N, 2 ** 256 - 1][i // 13 % 5]
        recid = int(rng.integers(0, 4)) if i % 17 else int(rng.integers(4, 256))
        hashes.append(bytes(rng.bytes(32)))
        sigs.append(r.to_bytes(32, "big") + s.to_bytes(32, "big") + bytes([recid]))
    pub, addr, ok = run(ctx, hashes, sigs)
    n_ok = 0
    for i in range(len(hashes)):
        want = oracle.ecrecover(hashes[i], sigs[i])
        assert bool(ok[i]) == (want is not None), i
        if want:
            assert pub[i].tobytes() == want, i
            assert addr[i].tobytes() == oracle.keccak256(want[1:])[12:], i
            n_ok += 1
        else:
            assert not pub[i].any() and not addr[i].any()
    assert n_ok == 752  # what the oracle says for this seed: recid 2 / 3 almost never lifts, half of the abscissae are off the curve

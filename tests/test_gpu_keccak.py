"""K: batched Keccak-256 on the GPU vs the oracle / reference vectors.  Bit-exact."""
import numpy as np
import pytest

import oracle_lib
from gpu_util import random_csr

pytestmark = pytest.mark.gpu

VARIANTS = {"staged": 0, "staged_nobin": 1 << 6, "direct": 1 << 4, "direct_nobin": (1 << 4) | (1 << 6), "warp": 1 << 5}


@pytest.fixture(scope="module")
def ctx():
    from phant_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


def gpu_hash(ctx, data, off, flags=0):
    n = len(off) - 1
    out = np.zeros((n, 32), np.uint8)
    ctx.set_flags(flags)
    ctx.keccak256_batch(data, off, n, out)
    return out


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_reference_keccak_table(ctx, golden, variant):
    """ethash/test/unittests/test_keccak.cpp:25-195 through the GPU, every prefix length, 8 byte offsets."""
    g = golden("keccak_kat.json")
    text = g["text"].encode()
    for shift in range(8):
        # messages laid out one after another, each preceded by `shift` bytes of garbage, so every case
        # starts at a different residue mod 8 / mod 16
        data, starts = bytearray(), []
        for c in g["cases"]:
            data += b"\xaa" * shift
            starts.append(len(data))
            data += text[:c["len"]]
        data = np.frombuffer(bytes(data) + b"\x00" * 32, np.uint8)
        for j in range(shift, len(g["cases"]), 9):
            c = g["cases"][j]
            o = np.array([starts[j], starts[j] + c["len"]], np.uint64)
            got = gpu_hash(ctx, data, o, VARIANTS[variant])
            assert got[0].tobytes().hex() == c["keccak256"], (variant, shift, c["len"])


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_random_lengths_vs_oracle(ctx, oracle, variant):
    rng = np.random.default_rng(42)
    n = 6000  # >= 4096 so the regrouping path runs for the binned variants
    data, off = random_csr(rng, n)
    got = gpu_hash(ctx, data, off, VARIANTS[variant])
    want = oracle.keccak256_batch(data, off, threads=8)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (variant, bad[:5], (off[bad[:5] + 1] - off[bad[:5]]), off[bad[:5]] % 16)


def test_long_messages(ctx, oracle):
    """contract code sized inputs (keccak(code) in the account leaf): up to 49 KB, odd offsets"""
    rng = np.random.default_rng(5)
    lens = [0, 1, 24576, 49152, 49153, 5000, 13, 136 * 40, 136 * 40 + 135]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[0] = 3
    off[1:] = 3 + np.cumsum(lens)
    data = rng.integers(0, 256, int(off[-1]) + 32, dtype=np.uint8)
    want = oracle.keccak256_batch(data, off)
    for v in VARIANTS.values():
        assert (gpu_hash(ctx, data, off, v) == want).all()


def test_empty_batch_and_invalid(ctx):
    from phant_b200 import gpu
    ctx.set_flags(0)
    ctx.keccak256_batch(None, None, 0, None)  # n == 0 is fine
    off = np.array([5, 3], np.uint64)         # non-monotone offsets
    with pytest.raises(gpu.PhantGpuError) as e:
        ctx.keccak256_batch(np.zeros(8, np.uint8), off, 1, np.zeros(32, np.uint8))
    assert e.value.code == -1


def test_device_pointer_mode_and_size(ctx, oracle):
    """200k full branch nodes (532 B) resident in HBM; digests compared on a strided sample + a linearity-free
    property: identical messages give identical digests, a one-bit change does not."""
    import torch
    from phant_b200 import gpu
    n, size = 200_000, 532
    g = torch.Generator(device="cuda").manual_seed(1)
    data = torch.randint(0, 256, (n * size + 64,), dtype=torch.uint8, device="cuda", generator=g)
    data[size:2 * size] = data[0:size]                       # message 1 == message 0
    off = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * size
    out = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for flags in (0, 1 << 4):
        out.zero_()
        ctx.set_flags(gpu.FLAG_DEVICE_PTRS | flags)
        ctx.keccak256_batch(data, off, n, out)
        ctx.synchronize()
        h = out.cpu().numpy()
        assert (h[0] == h[1]).all() and not (h[1] == h[2]).all()
        host = data.cpu().numpy()
        for i in range(0, n, 997):
            assert h[i].tobytes() == oracle.keccak256(host[i * size:(i + 1) * size].tobytes()), i
    ctx.set_flags(0)
    st = ctx.stats()
    assert st["keccak_msgs"] >= 2 * n and st["keccak_perms"] >= 2 * n * 4 and st["keccak_ms"] > 0


def test_fixture_header_hashes(ctx, golden):
    """the 87 block headers of the fixtures in one K call: keccak256(rlp(header)) == the header's `hash` field
    (what src/blockchain/blockchain.zig:135-137 relies on); no oracle in the loop"""
    g = golden("fixture_states.json.gz")
    blocks = [b for t in g["tests"] for b in t["blocks"]]
    data, off = oracle_lib.csr([bytes.fromhex(b["header_rlp"]) for b in blocks], np.uint64)
    got = gpu_hash(ctx, np.concatenate([data, np.zeros(32, np.uint8)]), off)
    assert [h.tobytes().hex() for h in got] == [b["hash"] for b in blocks] and len(blocks) == 87


def test_regrouping_by_block_count_is_a_permutation(oracle):
    """The two-launch counting sort that regroups messages by rate-block count (keccak_class_kernel /
    keccak_regroup_kernel) must visit every message exactly once: 700k messages of 0..2500 bytes (1..19 blocks, i.e. all
    16 classes incl. the clamped one), poisoned output buffer, device pointers, twice in a row on one context."""
    import torch
    from phant_b200 import gpu
    rng = np.random.default_rng(77)
    n = 700_000
    lens = rng.integers(0, 2501, n)
    lens[rng.random(n) < 0.5] = 532
    lens[rng.random(n) < 0.2] = 112
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens).astype(np.uint64)
    data = rng.integers(0, 256, int(off[-1]) + 64, dtype=np.uint8)
    want = oracle.keccak256_batch(data, off, threads=8)
    ctx = gpu.Context(0, gpu.FLAG_DEVICE_PTRS)
    d_data = torch.from_numpy(data).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    for _ in range(2):
        d_out = torch.full((n, 32), 0xEE, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.keccak256_batch(d_data, d_off, n, d_out)
        ctx.synchronize()
        assert (d_out.cpu().numpy() == want).all()
    d_out = torch.full((n, 32), 0xEE, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.keccak256_batch_async(d_data, d_off, n, int(off[-1]), d_out)   # total supplied: no read-back inside the call
    ctx.synchronize()
    assert (d_out.cpu().numpy() == want).all()
    assert ctx.stats()["keccak_perms"] == 3 * int((lens // 136 + 1).sum())
    ctx.close()

"""N2: the witness wire format in front of the verifier -- host logic only, no GPU.
rlp([headers, codes, state]); strict canonical RLP; anything else is InvalidWitness (never a silent accept)."""
import numpy as np
import pytest

from phant_b200 import host


def sample(rng):
    headers = [host._rlp_list([host._rlp_str(rng.integers(0, 256, 32, dtype=np.uint8).tobytes()), host._rlp_str(b"\x01")]) for _ in range(2)]
    codes = [b"", b"\x00", b"\x7f", b"\x80", rng.integers(0, 256, 300, dtype=np.uint8).tobytes()]
    nodes = [rng.integers(0, 256, int(l), dtype=np.uint8).tobytes() for l in (532, 112, 83, 55, 56, 1, 70000)]
    return headers, codes, nodes


def test_round_trip():
    rng = np.random.default_rng(1)
    h, c, n = sample(rng)
    blob = host.encode_witness(h, c, n)
    assert host.decode_witness(blob) == (h, c, n)
    assert host.decode_witness(host.encode_witness([], [], [])) == ([], [], [])
    assert host.decode_witness(bytes.fromhex("c3c0c0c0")) == ([], [], [])


@pytest.mark.parametrize("bad", [
    "", "80", "c0", "c2c0c0", "c4c0c0c0c0",     # not a list / wrong arity
    "c3c0c080", "c380c0c0",                      # field is a string, not a list
    "c4c0c0c0",                                  # outer length overruns
    "c3c0c0c000",                                # trailing byte
    "c5c0c0c28100",                              # 0x00 must encode itself
    "c6c0c0c3b80100",                            # long form for a 1-byte payload
    "c7c0c0c4b9000100",                          # leading zero in the length
    "c4c0c0c181",                                # truncated node
    "c5c0c0c2c101",                              # node is a list
    "c4c180c0c0",                                # header is a string
])
def test_malformed_is_refused(bad):
    with pytest.raises(host.InvalidWitness):
        host.decode_witness(bytes.fromhex(bad))


def test_every_truncation_and_byte_flip_in_the_framing_is_refused_or_changes_the_nodes():
    rng = np.random.default_rng(2)
    h, c, n = sample(rng)
    n = n[:5]
    blob = host.encode_witness(h, c, n)
    for cut in range(len(blob)):
        with pytest.raises(host.InvalidWitness):
            host.decode_witness(blob[:cut])
    for pos in rng.integers(0, len(blob), 400):
        mut = bytearray(blob)
        mut[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            got = host.decode_witness(bytes(mut))
        except host.InvalidWitness:
            continue
        assert got != (h, c, n)  # decodable, but then it is a different witness (the walk will judge its nodes)


def test_new_payload_v2_over_an_oracle_backed_context(oracle, golden):
    """the handler mirror end to end on the CPU (host logic only; tests/test_gpu_host_py.py runs it on the device): roots of
    the payload's two lists with 32-byte index keys, witness verdicts, senders"""
    from helpers import OracleBackedCtx, secure_account_items
    ctx = OracleBackedCtx(oracle)
    g = golden("fixture_states.json.gz")
    accounts = max(g["tables"].values(), key=len)[:40]
    items = secure_account_items(oracle.keccak256, oracle.mptize, accounts)
    trie = oracle.trie(items)
    keys = [k for k, _ in items[:10]] + [oracle.keccak256(b"nobody")]
    nodes = list({nd: 1 for k in keys for nd in trie.prove(k)})
    txs = [bytes.fromhex(t["encoded"]) for t in golden("ecrecover_kat.json")["txs"]]
    wds = [host._rlp_list([host._rlp_uint(i), host._rlp_uint(7), host._rlp_str(bytes(20)), host._rlp_uint(1000 + i)]) for i in range(5)]
    r = host.new_payload_v2(ctx, txs, wds, host.encode_witness([], [], nodes), trie.root(), keys, chain_id=1)
    assert r["accept"] and r["witness_status"] == [1] * 10 + [2]
    assert [a.hex() for a in r["senders"]] == [t["sender"] for t in golden("ecrecover_kat.json")["txs"]]
    assert r["transactions_root"] == oracle.mptize([(i.to_bytes(32, "big"), t) for i, t in enumerate(txs)])
    assert r["withdrawals_root"] == oracle.mptize([(i.to_bytes(32, "big"), w) for i, w in enumerate(wds)])
    short = host.new_payload_v2(ctx, txs, wds, host.encode_witness([], [], nodes[1:]), trie.root(), keys)
    assert not short["accept"] and set(short["witness_status"]) & {0, 3}
    broken = host.new_payload_v2(ctx, txs, wds, b"\xc1", trie.root(), keys)
    assert not broken["accept"] and broken["witness_error"]
    assert host.new_payload_v2(ctx, [], [])["accept"]


def test_sample_new_payload_request(oracle, golden):
    """src/engine_api/engine_api.zig:87-134 feeds src/engine_api/test_req.json (no transactions, no withdrawals) to
    newPayloadV2Handler: its receiptsRoot is the empty-trie constant, and so are the two roots toBlock builds"""
    from helpers import OracleBackedCtx
    p = golden("engine_payload_kat.json")["payload"]
    assert p["transactions"] == [] and p["receiptsRoot"][2:] == host.EMPTY_MPT_ROOT.hex()
    r = host.new_payload_v2(OracleBackedCtx(oracle), [], [])
    assert r["transactions_root"] == r["withdrawals_root"] == host.EMPTY_MPT_ROOT == oracle.mptize([]) and r["accept"] and r["senders"] == []

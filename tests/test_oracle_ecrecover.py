"""Sender recovery (row N4): the oracle's secp256k1 recovery (oracle/secp256k1.c) against the reference's own vectors and
against OpenSSL, and the host mirror of TxSigner.get_sender (phant_b200/host.py::get_senders) over an oracle-backed
context.  CPU-only; the arithmetic the reference itself uses (libsecp256k1 behind zig-eth-secp256k1) is not in its tree."""
import numpy as np
import pytest

from helpers import OracleBackedCtx

N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
P = 2 ** 256 - 0x1000003D1
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798


def sig65(r, s, recid):
    return r.to_bytes(32, "big") + s.to_bytes(32, "big") + bytes([recid])


def test_reference_erecover_vector(oracle, golden):
    """src/crypto/ecdsa.zig:38-48 (generated with geth)"""
    k = golden("ecrecover_kat.json")["erecover"]
    assert oracle.ecrecover(bytes.fromhex(k["hash"]), bytes.fromhex(k["sig65"])).hex() == k["pubkey65"]


def test_mainnet_senders_through_the_host_mirror(oracle, golden):
    """src/signer/signer.zig:199-227: legacy (EIP-155) and EIP-1559 transactions with their on-chain senders"""
    from phant_b200 import host
    g = golden("ecrecover_kat.json")
    ctx = OracleBackedCtx(oracle)
    got = host.get_senders(ctx, [bytes.fromhex(t["encoded"]) for t in g["txs"]], chain_id=1)
    assert [a.hex() for a in got] == [t["sender"] for t in g["txs"]]
    # the error behaviour of get_sender: wrong chain id -> EIP155_v; high s -> InvalidS; garbage -> InvalidTransaction
    legacy = bytes.fromhex(g["txs"][0]["encoded"])
    assert str(host.get_senders(ctx, [legacy], chain_id=5)[0]) == "EIP155_v"
    assert str(host.get_senders(ctx, [legacy[:-1]], chain_id=1)[0]) == "InvalidTransaction"
    assert str(host.get_senders(ctx, [b"\x02\xc0"], chain_id=1)[0]) == "InvalidTransaction"
    _, r, s, recid = host._tx_signing_parts(legacy, 1)
    high_s = legacy.replace(s.to_bytes(32, "big"), (N - s).to_bytes(32, "big"))
    assert str(host.get_senders(ctx, [high_s], chain_id=1)[0]) == "InvalidS"
    # a flipped payload bit recovers some OTHER address (recovery cannot tell), never the real sender
    other = host.get_senders(ctx, [legacy[:5] + bytes([legacy[5] ^ 1]) + legacy[6:]], chain_id=1)[0]
    assert isinstance(other, (bytes, host.SenderError)) and other != bytes.fromhex(g["txs"][0]["sender"])


def test_against_openssl(oracle):
    """random keys and digests signed by OpenSSL: exactly one of recid 0 / 1 recovers the signer's key, and the oracle's own
    public-key derivation agrees with OpenSSL's"""
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric.utils import Prehashed, decode_dss_signature
    rng = np.random.default_rng(8)
    for _ in range(120):
        priv = int.from_bytes(rng.bytes(32), "big") % (N - 1) + 1
        key = ec.derive_private_key(priv, ec.SECP256K1())
        pub = key.public_key().public_bytes(serialization.Encoding.X962, serialization.PublicFormat.UncompressedPoint)
        assert oracle.secp256k1_pubkey(priv.to_bytes(32, "big")) == pub
        digest = rng.bytes(32)
        r, s = decode_dss_signature(key.sign(digest, ec.ECDSA(Prehashed(hashes.SHA256()))))
        got = [oracle.ecrecover(digest, sig65(r, s, recid)) for recid in (0, 1)]
        assert got.count(pub) == 1, (r, s)
        assert got[0] != got[1]


def test_rejections(oracle):
    h = bytes(range(32))
    good_r = GX  # the generator's abscissa: certainly on the curve
    assert oracle.ecrecover(h, sig65(good_r, 5, 0)) is not None
    for r, s, recid in [(0, 5, 0), (good_r, 0, 0), (N, 5, 0), (good_r, N, 0), (N + 1, 5, 0), (good_r, 5, 4), (good_r, 5, 255),
                        (good_r, 5, 2)]:  # recid 2: x = r + n >= p
        assert oracle.ecrecover(h, sig65(r % 2 ** 256, s % 2 ** 256, recid)) is None, (r, s, recid)
    # about half of all abscissae are not on the curve
    rng = np.random.default_rng(3)
    misses = sum(oracle.ecrecover(h, sig65(int.from_bytes(rng.bytes(32), "big") % (N - 1) + 1, 7, 0)) is None for _ in range(60))
    assert 15 < misses < 45

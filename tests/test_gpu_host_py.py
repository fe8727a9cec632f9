"""The reference's hot-path tests restated against the Python host mirror (phant_b200/host.py)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from phant_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


def test_mpt_correctness(ctx):
    """src/mpt/mpt.zig:316-391 "correctness", line for line."""
    from phant_b200.host import KeyVal, mptize, EMPTY_MPT_ROOT
    cases = [
        ("empty", [], EMPTY_MPT_ROOT.hex()),
        ("single key - root is a leaf node", [KeyVal.init(bytes([1, 2, 3, 4]), b"hello")],
         "6764f7ad0efcbc11b84fe7567773aa4b12bd6b4d35c05bbc3951b58dedb6c8e8"),
        ("two keys - root is a branch node with two (embedded) leaf nodes",
         [KeyVal.init(bytes([1, 2, 3, 4]), b"hello1"), KeyVal.init(bytes([255, 2, 3, 4]), b"hello2")],
         "5c474c00e417f587322ae674c948f04e2c217f95bd1dac806af14fa46f8fa403"),
        ("three keys - two embedded leaves and one hashed node",
         [KeyVal.init(bytes([1 << 4, 2, 3, 4]), b"hello1"), KeyVal.init(bytes([2 << 4, 2, 3, 4]), b"hello2"),
          KeyVal.init(bytes([3 << 4, 2, 3, 4]), b"hello333333333333333333333333333")],
         "86d4d51eedae1cd8ffdfeef48e5f1cd021d84c8d3df0088dfad39e72b37fc4b1"),
        ("two keys - extension node of 3 nibbles and two leaf nodes",
         [KeyVal.init(bytes([0, 0xf1, 3, 4]), b"hello1"), KeyVal.init(bytes([0, 0xf2, 3, 4]), b"hello2")],
         "312b81f16960a816e84679c5b9de49471b07b5c11ef0eff19779b083e418f83b"),
        ("complex - 5 levels, 3 branch nodes, 2 extension nodes, 4 leaf nodes",
         [KeyVal.init(bytes([0x34, 0x57, 0x81]), b"hello1"), KeyVal.init(bytes([0x34, 0x57, 0x83]), b"hello2"),
          KeyVal.init(bytes([0x34, 0x5F, 2, 3]), b"hello3"), KeyVal.init(bytes([0xFF, 1, 2, 3]), b"hello4")],
         "c66c75a03f2b52dfc32b5e229bb2ff7e1d53dcb2b54fe83a1b39418788e0fc66"),
        ("complex - one branch node with a value, 40-byte value",
         [KeyVal.init(bytes([0x34]), b"hello1"), KeyVal.init(bytes([0x34, 0x57, 0x81]), b"hello2"),
          KeyVal.init(bytes([0x34, 0x57, 0x83]), b"hello3"), KeyVal.init(bytes([0x34, 0x5F, 2, 3]), b"hello4"),
          KeyVal.init(bytes([0xEF, 1, 2, 3]), b"0123456789012345678901234567890123456789"),
          KeyVal.init(bytes([0xFF, 1, 2, 3]), b"hello5")],
         "88a4fc29676ebee58aafcd377acd46af6d29044f9bb8220c50ca8dcfe5153fb3"),
    ]
    for name, keyvals, exp in cases:
        assert mptize(ctx, keyvals).hex() == exp, name


def test_block_roots_like_run_block(ctx, golden):
    """src/blockchain/blockchain.zig:76-90 post-checks: transactions / withdrawals roots of every valid fixture
    block through calculate_mpt_root, plus the state-root check phant has commented out (:83-85)."""
    from phant_b200.host import AccountState, StateDB, calculate_mpt_root
    g = golden("fixture_states.json.gz")
    tables = {}
    for key, accounts in g["tables"].items():
        db = StateDB()
        for a in accounts:
            db.db[bytes.fromhex(a["address"])] = AccountState(a["nonce"], int(a["balance"], 16), bytes.fromhex(a["code"]),
                                                              {int(k, 16): int(v, 16) for k, v in a["storage"].items()})
        tables[key] = db.root(ctx).hex()
    for t in g["tests"]:
        assert tables[t["pre"]] == t["pre_root"]
        assert tables[t["post"]] == t["post_root"]
        for b in t["blocks"]:
            assert calculate_mpt_root(ctx, [bytes.fromhex(x) for x in b["tx_values"]]).hex() == b["transactionsTrie"]
            assert calculate_mpt_root(ctx, [bytes.fromhex(x) for x in b["wd_values"]]).hex() == b["withdrawalsRoot"]


def test_payload_list_root_and_hasher(ctx, oracle):
    from phant_b200.host import KeyVal, keccak256, keccak256_with_prefix, payload_list_root
    items = [bytes([i]) * (i + 40) for i in range(130)]
    want = oracle.mptize([(i.to_bytes(32, "big"), v) for i, v in enumerate(items)])
    assert payload_list_root(ctx, items) == want
    assert keccak256(ctx, b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256_with_prefix(ctx, b"\x02", b"abc") == oracle.keccak256(b"\x02abc")
    assert KeyVal.less_than(KeyVal(b"\x01", b""), KeyVal(b"\x01\x00", b""))


def test_verify_witness(ctx, oracle, golden):
    from helpers import secure_account_items
    from phant_b200.host import verify_witness
    g = golden("fixture_states.json.gz")
    accounts = max(g["tables"].values(), key=len)[:60]
    items = secure_account_items(oracle.keccak256, oracle.mptize, accounts)
    trie = oracle.trie(items)
    proofs = [(k, trie.prove(k)) for k, _ in items]
    absent = oracle.keccak256(b"nobody")
    proofs.append((absent, trie.prove(absent)))
    st = verify_witness(ctx, trie.root(), proofs)
    assert st == [1] * len(items) + [2]
    bad = list(proofs)
    bad[3] = (bad[3][0], bad[3][1][:-1])
    assert verify_witness(ctx, trie.root(), bad)[3] == 0


def test_logs_bloom(ctx, oracle, golden):
    """row N3: Receipt.calculateLogsBloom for a whole block (src/types/receipt.zig:37-63)"""
    import numpy as np
    from phant_b200.host import Log, calculate_logs_blooms
    g = golden("logs_bloom_kat.json")
    logs = [Log(bytes.fromhex(a), [bytes.fromhex(t) for t in ts]) for a, ts in g["logs"]]
    blooms, block = calculate_logs_blooms(ctx, [logs, [], logs[:1]])
    assert blooms[0].hex() == g["bloom"] and blooms[1] == bytes(256)
    assert block == bytes(a | b for a, b in zip(blooms[0], blooms[2]))
    # many receipts vs the oracle
    rng = np.random.default_rng(4)
    receipts, items, own = [], [], []
    for r in range(500):
        ls = []
        for _ in range(int(rng.integers(0, 6))):
            lg = Log(rng.integers(0, 256, 20, dtype=np.uint8).tobytes(), [rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(int(rng.integers(0, 5)))])
            ls.append(lg)
            items.append(lg.address); own.append(r)
            for t in lg.topics:
                items.append(t); own.append(r)
        receipts.append(ls)
    blooms, block = calculate_logs_blooms(ctx, receipts)
    want = oracle.logs_bloom(items, own, 500)
    assert all(blooms[i] == want[i].tobytes() for i in range(500))
    assert block == np.bitwise_or.reduce(want, axis=0).tobytes()


def test_tx_hashes_and_addresses(ctx, oracle, golden):
    """row N4 (hashing half): Tx.hash for a block of transactions and address = keccak(pubkey[1:])[12:]"""
    from phant_b200.host import addresses_from_pubkeys, tx_hashes
    cases = golden("tx_hash_kat.json")["cases"]
    assert [h.hex() for h in tx_hashes(ctx, [bytes.fromhex(c["encoded"]) for c in cases])] == [c["hash"] for c in cases]
    pubs = [bytes([4]) + bytes([i]) * 64 for i in range(40)]
    assert addresses_from_pubkeys(ctx, pubs) == [oracle.keccak256(p[1:])[12:] for p in pubs]


def test_batched_list_roots(ctx, golden, oracle):
    """all 174 transaction / withdrawal tries of the fixtures in ONE forest build (phant_gpu_mpt_roots)"""
    from phant_b200 import gpu
    from phant_b200.host import KeyVal, calculate_mpt_roots, mptize_many
    g = golden("fixture_states.json.gz")
    lists, want = [], []
    for t in g["tests"]:
        for b in t["blocks"]:
            lists.append([bytes.fromhex(x) for x in b["tx_values"]]); want.append(b["transactionsTrie"])
            lists.append([bytes.fromhex(x) for x in b["wd_values"]]); want.append(b["withdrawalsRoot"])
    got = calculate_mpt_roots(ctx, lists)
    assert [r.hex() for r in got] == want and len(want) == 174
    # the sortedness check is per trie: [b, a] in one trie is refused, [b] + [a] in two tries is fine
    with pytest.raises(gpu.PhantGpuError):
        mptize_many(ctx, [[KeyVal(b"\x02", b"x"), KeyVal(b"\x01", b"y")]])
    two = mptize_many(ctx, [[KeyVal(b"\x02", b"x")], [KeyVal(b"\x01", b"y")], []])
    assert two[0] == oracle.mptize([(b"\x02", b"x")]) and two[1] == oracle.mptize([(b"\x01", b"y")])
    assert two[2].hex() == "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"


def test_run_block_post_checks(ctx, golden):
    """H12: the post-execution root comparisons of runBlock on every valid fixture block (receipts are not in the
    fixtures, so that root is skipped), with the last block's post state"""
    from phant_b200.host import AccountState, StateDB, run_block_post_checks
    g = golden("fixture_states.json.gz")
    checked = 0
    for t in g["tests"][:25]:
        if not t["blocks"]:
            continue
        db = StateDB()
        for a in g["tables"][t["post"]]:
            db.db[bytes.fromhex(a["address"])] = AccountState(a["nonce"], int(a["balance"], 16), bytes.fromhex(a["code"]),
                                                              {int(k, 16): int(v, 16) for k, v in a["storage"].items()})
        for i, b in enumerate(t["blocks"]):
            last = i == len(t["blocks"]) - 1
            header = {"transactions_root": bytes.fromhex(b["transactionsTrie"]), "withdrawals_root": bytes.fromhex(b["withdrawalsRoot"]),
                      "state_root": bytes.fromhex(t["post_root"])}
            txs = [bytes.fromhex(x) for x in b["tx_values"]]
            wds = [bytes.fromhex(x) for x in b["wd_values"]]
            assert run_block_post_checks(ctx, header, txs, None, wds, db if last else None) == []
            header["withdrawals_root"] = bytes(32)
            assert run_block_post_checks(ctx, header, txs, None, wds) == ["withdrawals_root"]
            checked += 1
    assert checked >= 20


def test_receipts_root_reference_vector(ctx, golden):
    """N3 end to end: logs -> blooms (GPU) -> receipt encodings (host RLP) -> receipts trie (GPU), against the root quoted in
    evmone/test/unittests/state_mpt_hash_test.cpp:192-245 (one legacy receipt with three logs, one log-free EIP-1559 receipt)"""
    from phant_b200.host import Log, Receipt, receipts_root
    g = golden("logs_bloom_kat.json")
    receipts = [Receipt(r["succeeded"], r["gas_used"],
                        [Log(bytes.fromhex(l["address"]), [bytes.fromhex(x) for x in l["topics"]], bytes.fromhex(l["data"])) for l in r["logs"]],
                        tx_type=r["type"]) for r in g["receipts"]]
    root, block_bloom = receipts_root(ctx, receipts)
    assert root.hex() == g["receipts_root"]
    assert receipts[0].bloom.hex() == g["bloom"] and block_bloom.hex() == g["bloom"]


def test_state_root_sharded_by_top_nibble(ctx, oracle, golden):
    """SURVEY.md 8e for S: the account trie split into the 16 subtrees under the root branch (phant_gpu_state_subtree_roots),
    ranks simulated one after another on this GPU; the combined root must equal the unsharded S root and the oracle's,
    for every world size, on random states (storage, code, lone-slot cases) and on the fixture states."""
    import numpy as np
    from phant_b200 import host, shard
    from test_shard_gloo import _random_statedb

    def oracle_root(db):
        accounts = [{"address": a.hex(), "nonce": s.nonce, "balance": "%064x" % s.balance, "code": s.code.hex(),
                     "storage": {"%064x" % k: "%064x" % v for k, v in s.storage.items()}} for a, s in sorted(db.db.items())]
        return oracle.state_root(accounts)

    def sharded(db, world):
        refs, mask = np.zeros((16, 32), np.uint8), 0
        for rank in range(world):
            r, m, _ = db.local_subtree_roots(ctx, rank, world)
            assert mask & m == 0 and not r[[v for v in range(16) if not (m >> v) & 1]].any()
            refs, mask = refs + r, mask | m
        return host.keccak256(ctx, shard.root_branch_rlp(refs, mask)) if bin(mask).count("1") >= 2 else None

    dbs = [_random_statedb(np.random.default_rng(s), n) for s, n in ((1, 2), (2, 17), (3, 400), (4, 5000))]
    g = golden("fixture_states.json.gz")
    for name in list(g["tables"])[:12]:
        db = host.StateDB()
        for a in g["tables"][name]:
            db.db[bytes.fromhex(a["address"])] = host.AccountState(a["nonce"], int(a["balance"], 16), bytes.fromhex(a["code"]),
                                                                 {int(k, 16): int(v, 16) for k, v in a["storage"].items()})
        dbs.append(db)
    branch_rooted = 0
    for db in dbs:
        want = db.root(ctx)
        assert want == oracle_root(db)
        assert db.root_sharded(ctx, 0, 1) == want  # world 1: no process group needed
        for world in (2, 4, 16):
            got = sharded(db, world)
            if got is not None:
                assert got == want, (len(db.db), world)
                branch_rooted += 1
    assert branch_rooted >= 20
    lone = _random_statedb(np.random.default_rng(9), 3, oracle)  # every account under slot 7: root is not a branch
    assert lone.root_sharded(ctx, 0, 1) == lone.root(ctx) == oracle_root(lone)


def test_payload_witness_blob(ctx, oracle, golden):
    """N2: witness bytes as they would arrive in the payload (rlp([headers, codes, state]), nodes as an unordered set)
    -> decode -> verifier; a dropped node shows as status 3, a damaged node as a missing child (3) or a reject (0)"""
    import random
    from helpers import secure_account_items
    from phant_b200 import host
    g = golden("fixture_states.json.gz")
    accounts = max(g["tables"].values(), key=len)[:80]
    items = secure_account_items(oracle.keccak256, oracle.mptize, accounts)
    trie = oracle.trie(items)
    keys = [k for k, _ in items[:50]] + [oracle.keccak256(b"nobody"), oracle.keccak256(b"nothing")]
    nodes = list({nd: 1 for k in keys for nd in trie.prove(k)})
    random.Random(5).shuffle(nodes)
    blob = host.encode_witness([], [bytes.fromhex(a["code"]) for a in accounts[:3]], nodes)
    assert host.verify_payload_witness(ctx, trie.root(), blob, keys) == [1] * 50 + [2, 2]
    leaf = trie.prove(keys[7])[-1]
    short = host.encode_witness([], [], [n for n in nodes if n != leaf])
    st = host.verify_payload_witness(ctx, trie.root(), short, keys)
    assert st[7] == 3 and all(s in (1, 2) for i, s in enumerate(st) if i != 7)
    dmg = [bytes([n[0]]) + bytes([n[1] ^ 0x40]) + n[2:] if n == leaf else n for n in nodes]
    assert host.verify_payload_witness(ctx, trie.root(), host.encode_witness([], [], dmg), keys)[7] in (0, 3)
    with pytest.raises(host.InvalidWitness):
        host.verify_payload_witness(ctx, trie.root(), blob[:-1], keys)


def test_new_payload_v2_on_the_device(ctx, oracle, golden):
    """N2 end to end through the REAL library (execution_payload.zig:125-183 with its TODO filled in): the payload's two index
    tries as one forest (M), the witness blob decoded and every touched key resolved in its node set from the parent state
    root (W), all senders in one recovery call (R).  Same scenario as the CPU test of the host logic
    (tests/test_host_witness.py), here nothing stands in for the GPU.  The witness is cut from a fixture state whose root the
    fixture header pins; the transactions are the reference's mainnet vectors (signer.zig:199-227, transaction.zig:282-303)."""
    from phant_b200 import host
    from helpers import secure_account_items
    g = golden("fixture_states.json.gz")
    root_of = {t["pre"]: t["pre_root"] for t in g["tests"]}
    tab = max((k for k in g["tables"] if k in root_of), key=lambda k: len(g["tables"][k]))
    accounts = g["tables"][tab]
    items = secure_account_items(oracle.keccak256, oracle.mptize, accounts)
    trie = oracle.trie(items)                                   # cuts the witness; the root it must reach is the fixture's
    state_root = bytes.fromhex(root_of[tab])
    keys = [k for k, _ in items[:40]] + [oracle.keccak256(b"nobody"), oracle.keccak256(b"nobody else")]
    nodes = list({nd: 1 for k in keys for nd in trie.prove(k)})
    kat = golden("ecrecover_kat.json")["txs"]
    txs = [bytes.fromhex(t["encoded"]) for t in kat]
    wds = [host._rlp_list([host._rlp_uint(i), host._rlp_uint(7), host._rlp_str(bytes(20)), host._rlp_uint(1000 + i)]) for i in range(5)]
    r = host.new_payload_v2(ctx, txs, wds, host.encode_witness([], [], nodes), state_root, keys, chain_id=1)
    assert r["accept"] and r["witness_status"] == [1] * 40 + [2, 2]
    assert [a.hex() for a in r["senders"]] == [t["sender"] for t in kat]
    assert r["transactions_root"] == oracle.mptize([(i.to_bytes(32, "big"), t) for i, t in enumerate(txs)])
    assert r["withdrawals_root"] == oracle.mptize([(i.to_bytes(32, "big"), w) for i, w in enumerate(wds)])
    # an incomplete witness, a wrong parent root and an undecodable blob all refuse the payload before execution
    short = host.new_payload_v2(ctx, txs, wds, host.encode_witness([], [], nodes[1:]), state_root, keys)
    assert not short["accept"] and set(short["witness_status"]) & {0, 3}
    wrong = host.new_payload_v2(ctx, txs, wds, host.encode_witness([], [], nodes), bytes([state_root[0] ^ 1]) + state_root[1:], keys)
    assert not wrong["accept"] and set(wrong["witness_status"]) == {3}   # the root node is not in the set
    broken = host.new_payload_v2(ctx, txs, wds, b"\xc1", state_root, keys)
    assert not broken["accept"] and broken["witness_error"]
    # the reference's own sample request (src/engine_api/test_req.json via engine_api.zig:87-134): no transactions, no
    # withdrawals -> both roots are the empty-trie constant its receiptsRoot field also shows
    p = golden("engine_payload_kat.json")["payload"]
    e = host.new_payload_v2(ctx, [bytes.fromhex(t[2:]) for t in p["transactions"]], [])
    assert e["accept"] and e["transactions_root"] == e["withdrawals_root"] == host.EMPTY_MPT_ROOT == bytes.fromhex(p["receiptsRoot"][2:])

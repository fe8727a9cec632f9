// host_test.cpp -- the reference's own hot-path tests, restated against the C++ host mirror (needs a B200).
//   mptize "correctness" cases          src/mpt/mpt.zig:316-391
//   keccak constants                    src/blockchain/vm.zig:22, src/mpt/mpt.zig:10, src/types/block.zig:13
//   state roots                         evmone/test/unittests/state_mpt_hash_test.cpp:19-66
// Exit code 0 = all passed.  Built and run by tests/test_gpu_host_cpp.py.
#include "phant_host.hpp"

#include <cstdio>
#include <string>

using namespace phant;

static Bytes B(const std::string& s) { return Bytes(s.begin(), s.end()); }
static Bytes X(const std::string& hex)
{
    Bytes out;
    for (size_t i = 0; i + 1 < hex.size(); i += 2) out.push_back((uint8_t)std::stoi(hex.substr(i, 2), nullptr, 16));
    return out;
}
template <class Bytes_> static std::string H(const Bytes_& h)
{
    static const char* d = "0123456789abcdef";
    std::string s;
    for (uint8_t b : h) { s.push_back(d[b >> 4]); s.push_back(d[b & 15]); }
    return s;
}
static int failures = 0;
static void expect(const std::string& got, const std::string& want, const char* name)
{
    if (got != want) { std::printf("FAIL %s\n  got  %s\n  want %s\n", name, got.c_str(), want.c_str()); ++failures; }
    else std::printf("ok   %s\n", name);
}

int main()
{
    Gpu g(0);
    using mpt::KeyVal;
    // ---- src/mpt/mpt.zig:326-385 ----
    expect(H(mpt::mptize(g, {})), "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421", "mptize empty");
    expect(H(mpt::mptize(g, {KeyVal::init({1, 2, 3, 4}, B("hello"))})), "6764f7ad0efcbc11b84fe7567773aa4b12bd6b4d35c05bbc3951b58dedb6c8e8",
           "single key - root is a leaf node");
    expect(H(mpt::mptize(g, {KeyVal::init({1, 2, 3, 4}, B("hello1")), KeyVal::init({255, 2, 3, 4}, B("hello2"))})),
           "5c474c00e417f587322ae674c948f04e2c217f95bd1dac806af14fa46f8fa403", "two keys - two embedded leaves");
    expect(H(mpt::mptize(g, {KeyVal::init({1 << 4, 2, 3, 4}, B("hello1")), KeyVal::init({2 << 4, 2, 3, 4}, B("hello2")),
                             KeyVal::init({3 << 4, 2, 3, 4}, B("hello333333333333333333333333333"))})),
           "86d4d51eedae1cd8ffdfeef48e5f1cd021d84c8d3df0088dfad39e72b37fc4b1", "three keys - one hashed child");
    expect(H(mpt::mptize(g, {KeyVal::init({0, 0xf1, 3, 4}, B("hello1")), KeyVal::init({0, 0xf2, 3, 4}, B("hello2"))})),
           "312b81f16960a816e84679c5b9de49471b07b5c11ef0eff19779b083e418f83b", "two keys - extension of 3 nibbles");
    expect(H(mpt::mptize(g, {KeyVal::init({0x34, 0x57, 0x81}, B("hello1")), KeyVal::init({0x34, 0x57, 0x83}, B("hello2")),
                             KeyVal::init({0x34, 0x5F, 2, 3}, B("hello3")), KeyVal::init({0xFF, 1, 2, 3}, B("hello4"))})),
           "c66c75a03f2b52dfc32b5e229bb2ff7e1d53dcb2b54fe83a1b39418788e0fc66", "complex - 5 levels");
    expect(H(mpt::mptize(g, {KeyVal::init({0x34}, B("hello1")), KeyVal::init({0x34, 0x57, 0x81}, B("hello2")),
                             KeyVal::init({0x34, 0x57, 0x83}, B("hello3")), KeyVal::init({0x34, 0x5F, 2, 3}, B("hello4")),
                             KeyVal::init({0xEF, 1, 2, 3}, B("0123456789012345678901234567890123456789")),
                             KeyVal::init({0xFF, 1, 2, 3}, B("hello5"))})),
           "88a4fc29676ebee58aafcd377acd46af6d29044f9bb8220c50ca8dcfe5153fb3", "complex - branch with a value");
    // unsorted input must be refused (the reference asserts, mpt.zig:39)
    try {
        mpt::mptize(g, {KeyVal::init({2}, B("a")), KeyVal::init({1}, B("b"))});
        expect("accepted", "GpuError", "unsorted list refused");
    } catch (const GpuError& e) {
        expect(std::to_string(e.code), "-1", "unsorted list refused");
    }
    // ---- keccak constants ----
    expect(H(hasher::keccak256(g, {})), "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470", "keccak('')");
    expect(H(hasher::keccak256(g, {0x80})), "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421", "keccak(0x80)");
    expect(H(hasher::keccak256WithPrefix(g, {}, {0xc0})), "1dcc4de8dec75d7aab85b567b6ccd41ad312451b948a7413f0a142fd40d49347", "keccak(0xc0)");
    // ---- calculateMPTRoot: empty and one-item lists (item 0 is keyed 0x80) ----
    expect(H(blockchain::calculateMPTRoot(g, {})), H(mpt::empty_mpt_root), "calculateMPTRoot empty");
    expect(H(blockchain::calculateMPTRoot(g, {B("x")})), H(mpt::mptize(g, {KeyVal::init({0x80}, B("x"))})), "calculateMPTRoot one item");
    // ---- state roots (go-ethereum derived, state_mpt_hash_test.cpp) ----
    state::StateDB db;
    expect(H(db.root(g)), H(mpt::empty_mpt_root), "StateDB.root empty");
    Address a2{}; a2[19] = 2;
    db.db[a2].balance[31] = 1;
    expect(H(db.root(g)), "084f337237951e425716a04fb0aaa74111eda9d9c61767f2497697d0a201c92e", "single_account_v1");
    state::StateDB db2;
    Address a0{}, a1{}; a1[19] = 1;
    db2.db[a0] = {};
    expect(H(db2.root(g)), "0ce23f3c809de377b008a4a3ee94a0834aac8bec1f86e28ffe4fdb5a15b0c785", "two_accounts step 1");
    auto& acc = db2.db[a1];
    acc.nonce = 1;
    acc.balance.fill(0xff); acc.balance[31] = 0xfe; // -2 as u256
    acc.code = {0x00};
    std::array<uint8_t, 32> k1{}, k2{}, v1{}, v2{};
    k1[31] = 1; k2[31] = 2; v1[31] = 0xfe; v2[31] = 0xfd;
    acc.storage[k1] = v1; acc.storage[k2] = v2;
    expect(H(db2.root(g)), "d3e845156fca75de99712281581304fbde104c0fc5a102b09288c07cdde0b666", "two_accounts step 2");
    // ---- the same root sharded over 2 / 4 / 16 ranks by top nibble: shares or-ed together, root branch hashed once ----
    state::StateDB big;
    for (int i = 0; i < 200; ++i) {
        Address a{}; a[0] = (uint8_t)i; a[7] = (uint8_t)(i * 31); a[19] = (uint8_t)(i >> 1);
        auto& s = big.db[a];
        s.nonce = i; s.balance[31] = (uint8_t)i; s.balance[20] = 1;
        if (i % 3 == 0) s.code = {0x60, (uint8_t)i};
        if (i % 5 == 0) { std::array<uint8_t, 32> k{}, v{}; k[31] = (uint8_t)i; v[15] = 7; s.storage[k] = v; }
    }
    const std::string big_root = H(big.root(g));
    for (int world : {1, 2, 4, 16}) {
        state::StateDB::SubtreeRoots all;
        for (int r = 0; r < world; ++r) {
            const auto part = big.subtreeRoots(g, r, world);
            if (all.mask & part.mask) expect("overlap", "disjoint", "subtree ownership");
            all.mask |= part.mask;
            for (size_t b = 0; b < all.refs.size(); ++b) all.refs[b] |= part.refs[b];
        }
        expect(H(state::StateDB::rootFromSubtreeRoots(g, all)), big_root, ("sharded StateDB.root world " + std::to_string(world)).c_str());
    }
    // ---- StateDB.root() block after block on a resident trie: load `big`, then apply only what a "block" touched ----
    {
        state::ResidentStateTrie rt(g);
        expect(H(rt.root()), H(mpt::empty_mpt_root), "ResidentStateTrie empty");
        std::map<Address, const state::AccountState*> all;
        for (const auto& [a, acc] : big.db) all[a] = &acc;
        expect(H(rt.apply(all)), big_root, "ResidentStateTrie: load == StateDB.root()");
        // a block: two balances change, one storage slot is written, one account is created, one destroyed
        std::map<Address, const state::AccountState*> touched;
        auto it = big.db.begin();
        it->second.balance[31] ^= 0x55; touched[it->first] = &it->second; ++it;
        it->second.nonce += 7; touched[it->first] = &it->second; ++it;
        { std::array<uint8_t, 32> k{}, v{}; k[0] = 9; v[31] = 3; it->second.storage[k] = v; touched[it->first] = &it->second; ++it; }
        const Address gone = it->first;
        touched[gone] = nullptr;
        Address fresh{}; fresh[3] = 0xee; fresh[19] = 0x42;
        big.db[fresh].balance[30] = 1;
        touched[fresh] = &big.db[fresh];
        const Hash32 after = rt.apply(touched);
        big.db.erase(gone);
        expect(H(after), H(big.root(g)), "ResidentStateTrie: after a block == StateDB.root() of the new state");
    }
    // ---- sender recovery: the geth-generated vector of src/crypto/ecdsa.zig:38-48, and a signature that recovers nothing ----
    {
        const Bytes hm = X("05e0e0ff09b01e5626daac3165b82afa42be29197b82e8a5a8800740ee7519d2");
        const Bytes sg = X("5a62891eb3e26f3a2344f93a7bad7fe5e670dc45cbdbf0e5bbdba4399238b5e6614caf592f96ee273a2bf018a976e7bf4b63777f9e53ce819d96c5035611400600");
        const Bytes pk = X("682bade67348db99074fcaaffef29394192e7e227a2bdb49f930c74358060c6a42df70f7ef8aadd94854abe646e047142fad42811e325afbec4753342d630b1e");
        Hash32 h{}; std::array<uint8_t, 65> s1{}, s0{};
        std::copy(hm.begin(), hm.end(), h.begin());
        std::copy(sg.begin(), sg.end(), s1.begin());
        const auto snd = signer::getSenders(g, {h, h}, {s1, s0}); // s0: r = s = 0
        const Hash32 want = hasher::keccak256(g, pk);
        expect(H(Bytes(snd.addresses[0].begin(), snd.addresses[0].end())), H(Bytes(want.begin() + 12, want.end())), "getSenders: erecover vector");
        expect(std::to_string((int)snd.recovered[0]) + std::to_string((int)snd.recovered[1]), "10", "getSenders: recovered flags");
    }
    // ---- witness: an absent key under the empty root, and a bogus chain ----
    engine_api::Witness w;
    Hash32 key{};
    w.add_proof({}, key);
    w.add_proof({Bytes{0xc0}}, key);
    auto st = engine_api::verifyWitness(g, mpt::empty_mpt_root, w);
    expect(std::to_string((int)st[0]) + std::to_string((int)st[1]), "20", "verifyWitness empty trie: absent / reject");
    std::printf(failures ? "FAILED %d\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}

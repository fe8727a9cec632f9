// comm_test.cpp -- world size 2 through the C ABI from C++ (needs two B200s): one process, one context + one host thread per
// GPU (the layout src/main.zig:143-149 suggests: handlers run on worker threads), NCCL behind phant_gpu_comm_init_local.
//   V sharded   phant_gpu_verify_proofs_sharded: every rank ends with the whole accept bitmap
//   C5 counts   phant_gpu_block_reject_counts: per-block reject counts summed over ranks
//   S sharded   phant_gpu_state_root_sharded == StateDB.root() on one GPU
// Exit code 0 = all passed; 77 = fewer than two usable devices.  Built and run by tests/test_gpu_comm.py.
#include "phant_host.hpp"

#include <cstdio>
#include <memory>
#include <string>
#include <thread>

using namespace phant;

static int failures = 0;
static void expect(bool ok, const char* name)
{
    std::printf("%s %s\n", ok ? "ok  " : "FAIL", name);
    if (!ok) ++failures;
}

int main()
{
    constexpr int W = 2;
    std::unique_ptr<Gpu> g[W];
    try {
        for (int r = 0; r < W; ++r) g[r] = std::make_unique<Gpu>(r);
    } catch (const GpuError& e) {
        std::printf("skip: %s\n", e.what());
        return 77;
    }
    phant_gpu_ctx* ctxs[W] = {g[0]->ctx(), g[1]->ctx()};
    int rc = phant_gpu_comm_init_local(ctxs, W);
    if (rc != 0) { std::printf("comm_init_local: %s [%s]\n", phant_gpu_strerror(rc), phant_gpu_last_error(ctxs[0])); return 1; }
    int rank = -1, world = 0, ver = 0;
    phant_gpu_comm_info(ctxs[1], &rank, &world, &ver);
    expect(rank == 1 && world == 2 && ver >= 20000, "comm_info");
    std::printf("NCCL version code %d\n", ver);

    // ---- V sharded: 1000 one-line proofs; proof p = empty chain, accepted (proven absent) iff its root is keccak(0x80) ----
    const uint64_t n = 1000;
    std::vector<uint8_t> keys(32 * n, 0x11), roots(32 * n, 0);
    for (uint64_t p = 0; p < n; ++p)
        if (p % 3) std::copy(mpt::empty_mpt_root.begin(), mpt::empty_mpt_root.end(), roots.begin() + 32 * p);
    const uint64_t words = phant_gpu_sharded_bitmap_words(n, W);
    std::vector<uint64_t> bitmap[W];
    std::vector<uint8_t> status[W];
    std::vector<uint32_t> counts[W];
    int rcs[W] = {0, 0}, rcs2[W] = {0, 0};
    auto verify_rank = [&](int r) {
        uint64_t lo, hi;
        phant_gpu_shard_range(n, r, W, &lo, &hi);
        std::vector<uint64_t> node_off{0}, first(hi - lo + 1, 0);
        phant_gpu_proof_batch b{};
        b.n_proofs = hi - lo;
        uint8_t dummy = 0;
        b.nodes = &dummy; b.node_off = node_off.data(); b.proof_first = first.data();
        b.keys32 = keys.data() + 32 * lo; b.roots32 = roots.data() + 32 * lo; b.n_roots = hi - lo;
        bitmap[r].assign(words, ~0ull);
        status[r].assign(hi - lo, 9);
        rcs[r] = phant_gpu_verify_proofs_sharded(ctxs[r], &b, n, bitmap[r].data(), status[r].data(), nullptr, nullptr);
        // blocks of 50 proofs, sharded with the proofs; counts summed over both ranks
        std::vector<uint32_t> block(hi - lo);
        for (uint64_t p = lo; p < hi; ++p) block[p - lo] = (uint32_t)(p / 50);
        counts[r].assign(n / 50, 99);
        rcs2[r] = phant_gpu_block_reject_counts(ctxs[r], status[r].data(), block.data(), hi - lo, n / 50, counts[r].data());
    };
    {
        std::thread t0(verify_rank, 0), t1(verify_rank, 1);
        t0.join(); t1.join();
    }
    expect(rcs[0] == 0 && rcs[1] == 0 && rcs2[0] == 0 && rcs2[1] == 0, "sharded calls return OK");
    uint64_t lo1, hi1;
    phant_gpu_shard_range(n, 1, W, &lo1, &hi1);
    const uint64_t per = words / W * 64;
    bool bits_ok = true, counts_ok = true;
    for (int r = 0; r < W; ++r) {
        for (uint64_t p = 0; p < n; ++p) {
            const uint64_t pos = p < lo1 ? p : per + (p - lo1); // rank 1's slice starts at word per/64
            const bool bit = (bitmap[r][pos / 64] >> (pos % 64)) & 1;
            bits_ok &= bit == (p % 3 != 0);
        }
        for (uint64_t b = 0; b < n / 50; ++b) {
            uint32_t want = 0;
            for (uint64_t p = 50 * b; p < 50 * b + 50; ++p) want += p % 3 == 0;
            counts_ok &= counts[r][b] == want;
        }
    }
    expect(bits_ok, "gathered accept bitmap identical and right on both ranks");
    expect(counts_ok, "per-block reject counts summed over ranks");

    // ---- S sharded ----
    state::StateDB big;
    for (int i = 0; i < 300; ++i) {
        Address a{}; a[0] = (uint8_t)i; a[7] = (uint8_t)(i * 31); a[19] = (uint8_t)(i >> 1);
        auto& s = big.db[a];
        s.nonce = i; s.balance[31] = (uint8_t)i; s.balance[20] = 1;
        if (i % 3 == 0) s.code = {0x60, (uint8_t)i};
        if (i % 5 == 0) { std::array<uint8_t, 32> k{}, v{}; k[31] = (uint8_t)i; v[15] = 7; s.storage[k] = v; }
    }
    const Hash32 want_root = big.root(*g[0]);
    Hash32 got[W];
    int rcs3[W];
    auto root_rank = [&](int r) {
        try { got[r] = big.rootSharded(*g[r], r, W); rcs3[r] = 0; } catch (const GpuError& e) { std::printf("rank %d: %s\n", r, e.what()); rcs3[r] = e.code; }
    };
    {
        std::thread t0(root_rank, 0), t1(root_rank, 1);
        t0.join(); t1.join();
    }
    expect(rcs3[0] == 0 && rcs3[1] == 0 && got[0] == want_root && got[1] == want_root, "phant_gpu_state_root_sharded == StateDB.root()");
    // a state whose accounts all fall under ONE root-branch slot: the plain-root fallback + broadcast
    state::StateDB lone;
    {
        std::vector<Bytes> addrs;
        for (int i = 0; i < 4000 && lone.db.size() < 3; ++i) {
            Address a{}; a[18] = (uint8_t)(i >> 8); a[19] = (uint8_t)i;
            const Hash32 h = hasher::keccak256(*g[0], Bytes(a.begin(), a.end()));
            if ((h[0] >> 4) == 0xb) lone.db[a].nonce = i + 1;
        }
    }
    const Hash32 lone_root = lone.root(*g[0]);
    auto lone_rank = [&](int r) {
        try { got[r] = lone.rootSharded(*g[r], r, W); rcs3[r] = 0; } catch (const GpuError& e) { std::printf("rank %d: %s\n", r, e.what()); rcs3[r] = e.code; }
    };
    {
        std::thread t0(lone_rank, 0), t1(lone_rank, 1);
        t0.join(); t1.join();
    }
    expect(rcs3[0] == 0 && rcs3[1] == 0 && got[0] == lone_root && got[1] == lone_root, "lone-slot state: plain root on the holder, broadcast");
    std::printf(failures ? "FAILED %d\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}

// ref_shim.cpp -- TEST INFRASTRUCTURE.  A C doorway into the reference's vendored evmone MPT
// (evmone/test/state/mpt.hpp:17-29), compiled by oracle/Makefile against the sources where they lie
// under /root/reference.  Used only to cross-check the oracle's secure-trie roots: evmone's MPT never
// embeds short nodes (evmone/test/state/mpt.cpp:243-252), so it equals mptize only when every node
// is >= 32 bytes (always true for 32-byte hashed keys in practice).
#include <test/state/mpt.hpp>
#include <cstdint>
#include <cstring>

extern "C" void ref_evmone_mpt_root(const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals,
                                    const uint64_t* val_off, uint64_t n, uint8_t out_root[32])
{
    evmone::state::MPT trie;
    for (uint64_t i = 0; i < n; ++i) {
        evmone::bytes v{vals + val_off[i], vals + val_off[i + 1]};
        trie.insert({keys + key_off[i], key_off[i + 1] - key_off[i]}, std::move(v));
    }
    const auto h = trie.hash();
    std::memcpy(out_root, h.bytes, 32);
}

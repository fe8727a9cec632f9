// node_summary.cuh -- the per-node summary the hash kernels leave for the proof walk (device function; also compiled as
// host code by tests/hostcheck/walk_host.cpp so that summary + walk can be fuzzed against the oracle without a GPU).
#pragma once
#include <stdint.h>

namespace phant {

// ------------------------------------------------------------------------------------------------
// node summary for the proof walk (walk_kernel.cu): while a node's bytes sit in shared memory / L1 for hashing,
// classify it once.  A SIMPLE BRANCH is a strictly canonical 17-item list (rule R2 of DESIGN.md) whose 16 children
// are each empty (0x80) or a 32-byte hash (0xa0 ..) and whose value is empty -- by far the common trie node.
// For those the walk needs no parse: summary = mask of hash children, header size, kind 1; the child for nibble n
// sits at hdr + 33*popc(mask & ((1<<n)-1)) + (n - popc(..)).  Everything else gets summary 0 = "walk parses it".
// The summary does not depend on the key, so it is also right for witness nodes shared between proofs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t summarize_node(const uint8_t* p, uint32_t len)
{
    if (len < 18) return 0;
    const uint32_t b0 = p[0];
    uint32_t hdr, pay;
    if (b0 < 0xc0) return 0;
    if (b0 <= 0xf7) { hdr = 1; pay = b0 - 0xc0; }
    else {
        const uint32_t n = b0 - 0xf7;
        if (n > 2 || p[1] == 0) return 0;
        pay = n == 1 ? p[1] : ((uint32_t)p[1] << 8) | p[2];
        if (pay <= 55) return 0;
        hdr = 1 + n;
    }
    if (hdr + pay != len) return 0;
    if (len == 532) { // the full branch (16 hashed children): 17 independent byte probes instead of a dependent scan
        uint32_t ok = p[531] == 0x80;
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) ok &= p[3 + 33 * i] == 0xa0;
        if (ok) return (0xffffu << 8) | (3u << 2) | 1u;
    }
    uint32_t o = hdr, mask = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < 16; ++i) {
        if (o >= len) return 0;
        const uint32_t c = p[o];
        if (c == 0x80) o += 1;
        else if (c == 0xa0) { mask |= 1u << i; o += 33; }
        else return 0;
    }
    if (o + 1 != len || p[o] != 0x80) return 0;
    return (mask << 8) | (hdr << 2) | 1u;
}

} // namespace phant

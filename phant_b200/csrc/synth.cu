// synth.cu -- synthetic witnesses generated in HBM, so that benchmark inputs of
// BASELINE.json's sizes (1M / 10M proofs, 3.9 / 28 GB) never cross PCIe.  Byte-identical twin of
// oracle/synth.c (tests/test_gpu_synth.py compares them); node shapes are the encodings
// src/mpt/mpt.zig:170-281 produces for a secure trie.  One proof per thread, hashing bottom-up with the
// same device Keccak the product kernels use.
#include "common.cuh"
#include "keccak_f1600.cuh"

namespace phant {
namespace {

__device__ __forceinline__ uint64_t sm64(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t stream_init(uint64_t seed, uint64_t tag, uint64_t index)
{
    uint64_t s = seed ^ (tag * 0xA24BAED4963EE407ull) ^ (index * 0xD1342543DE82EF95ull);
    (void)sm64(s);
    return s;
}
__device__ __forceinline__ void put64le(uint8_t* p, uint64_t v)
{
#pragma unroll
    for (int b = 0; b < 8; ++b) p[b] = (uint8_t)(v >> (8 * b));
}
__device__ __forceinline__ void draw32(uint64_t& s, uint8_t* out)
{
    for (int w = 0; w < 4; ++w) put64le(out + 8 * w, sm64(s));
}
__device__ __forceinline__ void put_digest(uint8_t* p, const uint64_t (&h)[4])
{
    for (int w = 0; w < 4; ++w) put64le(p + 8 * w, h[w]);
}
__device__ __forceinline__ uint32_t key_nib(const uint8_t* key, uint32_t k)
{
    return (k & 1) ? (key[k >> 1] & 15u) : (key[k >> 1] >> 4);
}

constexpr uint32_t FULL_BRANCH = 532, SPARSE_BRANCH = 83;

__device__ void put_full_branch(uint8_t* out, uint64_t& s, uint32_t nib, const uint64_t (&child)[4])
{
    out[0] = 0xf9; out[1] = 0x02; out[2] = 0x11;
    for (uint32_t slot = 0; slot < 16; ++slot) {
        uint8_t* p = out + 3 + 33 * slot;
        p[0] = 0xa0;
        if (slot == nib) put_digest(p + 1, child);
        else draw32(s, p + 1);
    }
    out[531] = 0x80;
}
__device__ void put_sparse_branch(uint8_t* out, uint64_t& s, uint32_t nib, const uint64_t (&child)[4])
{
    const uint32_t other = (nib + 1 + (uint32_t)(sm64(s) % 15)) % 16;
    uint64_t sib[4];
    for (int w = 0; w < 4; ++w) sib[w] = sm64(s);
    out[0] = 0xf8; out[1] = 0x51;
    uint32_t o = 2;
    for (uint32_t slot = 0; slot < 16; ++slot) {
        if (slot == nib || slot == other) {
            out[o++] = 0xa0;
            if (slot == nib) put_digest(out + o, child); else put_digest(out + o, sib);
            o += 32;
        } else out[o++] = 0x80;
    }
    out[o++] = 0x80;
}
__device__ uint32_t put_leaf_path(uint8_t* out, const uint8_t* key, uint32_t from)
{
    uint32_t cnt = 64 - from, o = 0, i = from;
    if (cnt & 1) { out[o++] = (uint8_t)(0x30 | key_nib(key, i)); i++; }
    else out[o++] = 0x20;
    for (; i < 64; i += 2) out[o++] = (uint8_t)((key_nib(key, i) << 4) | key_nib(key, i + 1));
    return o;
}
__device__ void maybe_corrupt(uint64_t seed, uint64_t gi, int corrupt, uint8_t* proof, uint64_t n_bytes)
{
    if (!corrupt || gi % 97 != 0) return;
    uint64_t s = stream_init(seed, 0xC0, gi);
    const uint64_t bit = sm64(s) % (8 * n_bytes);
    proof[bit >> 3] ^= (uint8_t)(1u << (bit & 7));
}

__host__ __device__ inline uint32_t c2_leaf_size(uint32_t depth) { return 2 + (1 + 1 + (65 - depth) / 2) + 80; }
__host__ __device__ inline uint64_t c2_bytes(uint32_t depth) { return (uint64_t)(depth - 1) * FULL_BRANCH + c2_leaf_size(depth); }

__global__ void __launch_bounds__(128)
synth_c2_kernel(uint64_t seed, uint64_t first_index, uint64_t n, uint32_t depth, int corrupt, uint8_t* __restrict__ nodes,
                uint64_t* __restrict__ node_off, uint64_t* __restrict__ proof_first, uint8_t* __restrict__ keys32,
                uint8_t* __restrict__ roots32)
{
    const uint64_t per = c2_bytes(depth);
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t gi = first_index + k;
        uint8_t* proof = nodes + per * k;
        uint8_t* key = keys32 + 32 * k;
        uint64_t s = stream_init(seed, 0xC2, gi);
        draw32(s, key);
        uint8_t* leaf = proof + (uint64_t)(depth - 1) * FULL_BRANCH;
        const uint32_t lsz = c2_leaf_size(depth);
        uint32_t o = 0;
        leaf[o++] = 0xf8; leaf[o++] = (uint8_t)(lsz - 2);
        const uint32_t hpn = 1 + (65 - depth) / 2;
        leaf[o++] = (uint8_t)(0x80 + hpn);
        o += put_leaf_path(leaf + o, key, depth - 1);
        leaf[o++] = 0xb8; leaf[o++] = 78;
        leaf[o++] = 0xf8; leaf[o++] = 76;
        leaf[o++] = (uint8_t)(1 + sm64(s) % 127);
        const uint64_t bal = sm64(s) | 0x8000000000000000ull;
        leaf[o++] = 0x88;
        for (int b = 0; b < 8; ++b) leaf[o++] = (uint8_t)(bal >> (8 * (7 - b)));
        leaf[o++] = 0xa0; draw32(s, leaf + o); o += 32;
        leaf[o++] = 0xa0; draw32(s, leaf + o); o += 32;
        uint64_t h[4];
        keccak256_thread<2>(leaf, lsz, h);
        for (int lvl = (int)depth - 2; lvl >= 0; --lvl) {
            uint8_t* br = proof + (uint64_t)lvl * FULL_BRANCH;
            put_full_branch(br, s, key_nib(key, (uint32_t)lvl), h);
            keccak256_thread<2>(br, FULL_BRANCH, h);
        }
        put_digest(roots32 + 32 * k, h);
        maybe_corrupt(seed, gi, corrupt, proof, per);
        proof_first[k] = k * depth;
        for (uint32_t j = 0; j < depth; ++j) node_off[k * depth + j] = per * k + (uint64_t)j * FULL_BRANCH;
        if (k == n - 1) { proof_first[n] = n * depth; node_off[n * depth] = per * n; }
    }
}

__device__ __forceinline__ uint32_t c3_depth(uint64_t seed, uint64_t gi, uint64_t& s)
{
    s = stream_init(seed, 0xC3, gi);
    return 4 + (uint32_t)(sm64(s) % 9);
}
__device__ __forceinline__ uint32_t c3_branch_size(uint32_t lvl) { return lvl < 5 ? FULL_BRANCH : SPARSE_BRANCH; }
__device__ __forceinline__ uint32_t c3_leaf_size(uint32_t d) { return 2 + (1 + 1 + (65 - d) / 2) + 34; }

__global__ void synth_c3_sizes_kernel(uint64_t seed, uint64_t first_index, uint64_t n, uint64_t* __restrict__ n_nodes,
                                      uint64_t* __restrict__ n_bytes)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s;
        const uint32_t d = c3_depth(seed, first_index + k, s);
        uint32_t b = c3_leaf_size(d);
        for (uint32_t l = 0; l + 1 < d; ++l) b += c3_branch_size(l);
        n_nodes[k] = d;
        n_bytes[k] = b;
    }
}

__global__ void __launch_bounds__(128)
synth_c3_kernel(uint64_t seed, uint64_t first_index, uint64_t n, int corrupt, const uint64_t* __restrict__ node_first,
                const uint64_t* __restrict__ byte_first, uint8_t* __restrict__ nodes, uint64_t* __restrict__ node_off,
                uint8_t* __restrict__ keys32, uint8_t* __restrict__ roots32)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t gi = first_index + k;
        uint64_t s;
        const uint32_t d = c3_depth(seed, gi, s);
        uint8_t* key = keys32 + 32 * k;
        draw32(s, key);
        const uint64_t base = byte_first[k];
        uint8_t* proof = nodes + base;
        uint64_t* noff = node_off + node_first[k];
        uint32_t rel[13];
        uint32_t off = 0;
        for (uint32_t l = 0; l + 1 < d; ++l) { rel[l] = off; off += c3_branch_size(l); }
        rel[d - 1] = off;
        const uint32_t lsz = c3_leaf_size(d);
        const uint32_t total = off + lsz;
        uint8_t* leaf = proof + off;
        uint32_t o = 0;
        leaf[o++] = 0xf8; leaf[o++] = (uint8_t)(lsz - 2);
        const uint32_t hpn = 1 + (65 - d) / 2;
        leaf[o++] = (uint8_t)(0x80 + hpn);
        o += put_leaf_path(leaf + o, key, d - 1);
        leaf[o++] = 0xa1; leaf[o++] = 0xa0;
        draw32(s, leaf + o);
        leaf[o] |= 0x80;
        o += 32;
        uint64_t h[4];
        keccak256_thread<2>(leaf, lsz, h);
        for (int lvl = (int)d - 2; lvl >= 0; --lvl) {
            uint8_t* br = proof + rel[lvl];
            if (lvl < 5) put_full_branch(br, s, key_nib(key, (uint32_t)lvl), h);
            else put_sparse_branch(br, s, key_nib(key, (uint32_t)lvl), h);
            keccak256_thread<2>(br, c3_branch_size((uint32_t)lvl), h);
        }
        put_digest(roots32 + 32 * k, h);
        maybe_corrupt(seed, gi, corrupt, proof, total);
        for (uint32_t j = 0; j < d; ++j) noff[j] = base + rel[j];
        if (k == n - 1) noff[d] = base + total;
    }
}

unsigned grid_for(int device, uint64_t n, unsigned block)
{
    uint64_t blocks = (n + block - 1) / block;
    const uint64_t cap = (uint64_t)keccak_num_sms(device) * 8;
    return (unsigned)(blocks > cap ? cap : (blocks ? blocks : 1));
}

} // namespace

cudaError_t launch_synth_c2(cudaStream_t s, int device, uint64_t seed, uint64_t first_index, uint64_t n, uint32_t depth,
                            int corrupt, uint8_t* nodes, uint64_t* node_off, uint64_t* proof_first, uint8_t* keys32,
                            uint8_t* roots32)
{
    if (n == 0) return cudaSuccess;
    synth_c2_kernel<<<grid_for(device, n, 128), 128, 0, s>>>(seed, first_index, n, depth, corrupt, nodes, node_off, proof_first,
                                                          keys32, roots32);
    return cudaGetLastError();
}
cudaError_t launch_synth_c3_sizes(cudaStream_t s, int device, uint64_t seed, uint64_t first_index, uint64_t n,
                                  uint64_t* n_nodes_per, uint64_t* n_bytes_per)
{
    if (n == 0) return cudaSuccess;
    synth_c3_sizes_kernel<<<grid_for(device, n, 256), 256, 0, s>>>(seed, first_index, n, n_nodes_per, n_bytes_per);
    return cudaGetLastError();
}
cudaError_t launch_synth_c3(cudaStream_t s, int device, uint64_t seed, uint64_t first_index, uint64_t n, int corrupt,
                            const uint64_t* node_first, const uint64_t* byte_first, uint8_t* nodes, uint64_t* node_off,
                            uint8_t* keys32, uint8_t* roots32)
{
    if (n == 0) return cudaSuccess;
    synth_c3_kernel<<<grid_for(device, n, 128), 128, 0, s>>>(seed, first_index, n, corrupt, node_first, byte_first, nodes, node_off,
                                                          keys32, roots32);
    return cudaGetLastError();
}

uint64_t synth_c2_bytes_per_proof(uint32_t depth) { return c2_bytes(depth); }

} // namespace phant

/* synth.c -- TEST INFRASTRUCTURE (see oracle.h).  Synthetic witnesses of SURVEY.md 8(d), CPU twin of
 * phant_b200/csrc/synth.cu (tests require the two to be byte-identical).
 *
 * Node shapes are the ones src/mpt/mpt.zig:170-281 produces for a secure (32-byte hashed key) trie:
 *   full branch   f9 0211 | 16 x (a0 hash32) | 80                      = 532 B
 *   sparse branch f8 51   | 14 x 80, 2 x (a0 hash32) | 80              =  83 B
 *   account leaf  f8 6e   | 9d hp(57 nibbles) | b8 4e rlp(account)     = 112 B   (depth 8)
 *   storage leaf  f8 xx   | hp | a1 a0 value32                         = 64..68 B
 * Every proof is its own little trie: siblings off the key's path are PRNG hashes, so the root is
 * per proof.  PRNG: splitmix64 stream keyed by (seed, global proof index) -- any proof can be
 * regenerated alone.
 */
#include "oracle.h"
#include <string.h>

static inline uint64_t sm64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t stream_init(uint64_t seed, uint64_t tag, uint64_t index)
{
    uint64_t s = seed ^ (tag * 0xA24BAED4963EE407ull) ^ (index * 0xD1342543DE82EF95ull);
    (void)sm64(&s);
    return s;
}
static inline void draw_bytes(uint64_t* s, uint8_t* out, int n_words)
{
    for (int w = 0; w < n_words; ++w) {
        uint64_t v = sm64(s);
        for (int b = 0; b < 8; ++b) out[8 * w + b] = (uint8_t)(v >> (8 * b));
    }
}

enum { FULL_BRANCH = 532, SPARSE_BRANCH = 83 };

/* full branch with the child hash at slot `nib`, PRNG hashes elsewhere */
static void put_full_branch(uint8_t* out, uint64_t* s, unsigned nib, const uint8_t child[32])
{
    out[0] = 0xf9; out[1] = 0x02; out[2] = 0x11;
    for (unsigned slot = 0; slot < 16; ++slot) {
        uint8_t* p = out + 3 + 33 * slot;
        p[0] = 0xa0;
        if (slot == nib) memcpy(p + 1, child, 32);
        else draw_bytes(s, p + 1, 4);
    }
    out[531] = 0x80;
}
static void put_sparse_branch(uint8_t* out, uint64_t* s, unsigned nib, const uint8_t child[32])
{
    unsigned other = (nib + 1 + (unsigned)(sm64(s) % 15)) % 16;
    uint8_t sib[32];
    draw_bytes(s, sib, 4);
    out[0] = 0xf8; out[1] = 0x51;
    uint32_t o = 2;
    for (unsigned slot = 0; slot < 16; ++slot) {
        if (slot == nib || slot == other) {
            out[o++] = 0xa0;
            memcpy(out + o, slot == nib ? child : sib, 32);
            o += 32;
        } else out[o++] = 0x80;
    }
    out[o++] = 0x80;
}
/* hex-prefix of key nibbles [from, 64) with the leaf flag; returns byte count */
static uint32_t put_leaf_path(uint8_t* out, const uint8_t key[32], uint32_t from)
{
    uint32_t cnt = 64 - from, o = 0, i = from;
#define KNIB(k) (((k) & 1) ? (key[(k) >> 1] & 15) : (key[(k) >> 1] >> 4))
    if (cnt & 1) { out[o++] = (uint8_t)(0x30 | KNIB(i)); i++; }
    else out[o++] = 0x20;
    for (; i < 64; i += 2) out[o++] = (uint8_t)((KNIB(i) << 4) | KNIB(i + 1));
    return o;
}

static void maybe_corrupt(uint64_t seed, uint64_t gi, int corrupt, uint8_t* proof, uint64_t n_bytes)
{
    if (!corrupt || gi % 97 != 0) return;
    uint64_t s = stream_init(seed, 0xC0, gi);
    uint64_t bit = sm64(&s) % (8 * n_bytes);
    proof[bit >> 3] ^= (uint8_t)(1u << (bit & 7));
}

/* ---------------- C2: account proofs ---------------- */
static uint32_t c2_leaf_size(uint32_t depth)
{
    uint32_t hpn = 1 + (65 - depth) / 2;
    return 2 + (1 + hpn) + 80;
}
uint64_t oracle_synth_c2_bytes_per_proof(uint32_t depth) { return (uint64_t)(depth - 1) * FULL_BRANCH + c2_leaf_size(depth); }

static void c2_one(uint64_t seed, uint64_t gi, uint32_t depth, int corrupt, uint8_t* proof, uint8_t* key, uint8_t* root)
{
    uint64_t s = stream_init(seed, 0xC2, gi);
    draw_bytes(&s, key, 4);
    uint8_t* leaf = proof + (uint64_t)(depth - 1) * FULL_BRANCH;
    uint32_t lsz = c2_leaf_size(depth);
    /* leaf = rlp([hp(path), rlp([nonce, balance, storageRoot, codeHash])]) */
    uint32_t o = 0;
    leaf[o++] = 0xf8; leaf[o++] = (uint8_t)(lsz - 2);
    uint8_t hp[33];
    uint32_t hpn = put_leaf_path(hp, key, depth - 1);
    leaf[o++] = (uint8_t)(0x80 + hpn);
    memcpy(leaf + o, hp, hpn); o += hpn;
    leaf[o++] = 0xb8; leaf[o++] = 78;
    leaf[o++] = 0xf8; leaf[o++] = 76;
    leaf[o++] = (uint8_t)(1 + sm64(&s) % 127);            /* nonce 1..127: a single byte */
    uint64_t bal = sm64(&s) | 0x8000000000000000ull;       /* 8 bytes, top byte non-zero */
    leaf[o++] = 0x88;
    for (int b = 0; b < 8; ++b) leaf[o++] = (uint8_t)(bal >> (8 * (7 - b)));
    leaf[o++] = 0xa0; draw_bytes(&s, leaf + o, 4); o += 32; /* storage root */
    leaf[o++] = 0xa0; draw_bytes(&s, leaf + o, 4); o += 32; /* code hash */
    uint8_t h[32];
    oracle_keccak256(leaf, lsz, h);
    for (int lvl = (int)depth - 2; lvl >= 0; --lvl) {
        uint8_t* br = proof + (uint64_t)lvl * FULL_BRANCH;
        put_full_branch(br, &s, KNIB((uint32_t)lvl), h);
        oracle_keccak256(br, FULL_BRANCH, h);
    }
    memcpy(root, h, 32);
    maybe_corrupt(seed, gi, corrupt, proof, oracle_synth_c2_bytes_per_proof(depth));
}

void oracle_synth_c2(uint64_t seed, uint64_t first_index, uint64_t n, uint32_t depth, int corrupt, uint8_t* nodes,
                     uint64_t* node_off, uint64_t* proof_first, uint8_t* keys32, uint8_t* roots32, int threads)
{
    uint64_t per = oracle_synth_c2_bytes_per_proof(depth);
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        c2_one(seed, first_index + (uint64_t)k, depth, corrupt, nodes + per * k, keys32 + 32 * k, roots32 + 32 * k);
        proof_first[k] = (uint64_t)k * depth;
        for (uint32_t j = 0; j < depth; ++j) node_off[(uint64_t)k * depth + j] = per * k + (uint64_t)j * FULL_BRANCH;
    }
    proof_first[n] = n * depth;
    node_off[n * depth] = per * n;
}

/* ---------------- C3: storage proofs, depth 4..12 ---------------- */
static uint32_t c3_depth(uint64_t seed, uint64_t gi, uint64_t* s_out)
{
    uint64_t s = stream_init(seed, 0xC3, gi);
    uint32_t d = 4 + (uint32_t)(sm64(&s) % 9);
    if (s_out) *s_out = s;
    return d;
}
static uint32_t c3_branch_size(uint32_t lvl) { return lvl < 5 ? FULL_BRANCH : SPARSE_BRANCH; }
static uint32_t c3_leaf_size(uint32_t d) { return 2 + (1 + 1 + (65 - d) / 2) + 34; }
static uint32_t c3_bytes(uint32_t d)
{
    uint32_t b = c3_leaf_size(d);
    for (uint32_t l = 0; l + 1 < d; ++l) b += c3_branch_size(l);
    return b;
}
void oracle_synth_c3_sizes(uint64_t seed, uint64_t first_index, uint64_t n, uint32_t* n_nodes, uint32_t* n_bytes)
{
    for (uint64_t k = 0; k < n; ++k) {
        uint32_t d = c3_depth(seed, first_index + k, NULL);
        n_nodes[k] = d;
        n_bytes[k] = c3_bytes(d);
    }
}
static void c3_one(uint64_t seed, uint64_t gi, int corrupt, uint8_t* proof, uint64_t* noff /* d+1, relative */,
                   uint8_t* key, uint8_t* root)
{
    uint64_t s;
    uint32_t d = c3_depth(seed, gi, &s);
    draw_bytes(&s, key, 4);
    uint32_t off = 0;
    for (uint32_t l = 0; l + 1 < d; ++l) { noff[l] = off; off += c3_branch_size(l); }
    noff[d - 1] = off;
    uint32_t lsz = c3_leaf_size(d);
    noff[d] = off + lsz;
    uint8_t* leaf = proof + off;
    uint32_t o = 0;
    leaf[o++] = 0xf8; leaf[o++] = (uint8_t)(lsz - 2);
    uint8_t hp[33];
    uint32_t hpn = put_leaf_path(hp, key, d - 1);
    leaf[o++] = (uint8_t)(0x80 + hpn);
    memcpy(leaf + o, hp, hpn); o += hpn;
    leaf[o++] = 0xa1; leaf[o++] = 0xa0;
    draw_bytes(&s, leaf + o, 4);
    leaf[o] |= 0x80; /* 32-byte value, top byte non-zero */
    o += 32;
    uint8_t h[32];
    oracle_keccak256(leaf, lsz, h);
    for (int lvl = (int)d - 2; lvl >= 0; --lvl) {
        uint8_t* br = proof + noff[lvl];
        if (lvl < 5) put_full_branch(br, &s, KNIB((uint32_t)lvl), h);
        else put_sparse_branch(br, &s, KNIB((uint32_t)lvl), h);
        oracle_keccak256(br, c3_branch_size((uint32_t)lvl), h);
    }
    memcpy(root, h, 32);
    maybe_corrupt(seed, gi, corrupt, proof, noff[d]);
}
void oracle_synth_c3(uint64_t seed, uint64_t first_index, uint64_t n, int corrupt, uint8_t* nodes, uint64_t* node_off,
                     uint64_t* proof_first, uint8_t* keys32, uint8_t* roots32, int threads)
{
    /* serial prefix over the sizes, then fill in parallel */
    uint64_t nn = 0, nb = 0;
    for (uint64_t k = 0; k < n; ++k) {
        uint32_t d = c3_depth(seed, first_index + k, NULL);
        proof_first[k] = nn;
        node_off[nn] = nb;
        nn += d;
        nb += c3_bytes(d);
    }
    proof_first[n] = nn;
    node_off[nn] = nb;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        uint64_t rel[14];
        uint64_t base = node_off[proof_first[k]];
        c3_one(seed, first_index + (uint64_t)k, corrupt, nodes + base, rel, keys32 + 32 * k, roots32 + 32 * k);
        uint32_t d = (uint32_t)(proof_first[k + 1] - proof_first[k]);
        for (uint32_t j = 0; j < d; ++j) node_off[proof_first[k] + j] = base + rel[j];
    }
}

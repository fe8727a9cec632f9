/* blocks.c -- TEST INFRASTRUCTURE (see oracle.h).  Synthetic block witnesses of BASELINE.json config
 * "1000 synthetic blocks x 300 tx each, full witness verify" (SURVEY.md 8d, row C5).
 *
 * Per block b a VIRTUAL state trie: only the paths of the touched accounts are materialised, every sibling
 * off those paths is a PRF hash of (seed, block, trie, path), so nodes shared between proofs (the root and
 * the top levels) are byte-identical and the witness can be DEDUPLICATED: `nodes` holds each distinct node
 * once and every proof is a list of node indices.  Per transaction t: two account proofs (sender A, contract
 * B; depth 8 = 7 full branches + the 112-byte account leaf, as mpt.zig:218-281 encodes them) and two storage
 * proofs inside B's storage trie (depth 6 = 5 branches + leaf), whose root is the storageRoot in B's account
 * leaf.  Account proofs verify against the block's state root, storage proofs against B's storage root.
 * Key prefixes are distinct by construction, so every path ends in its own leaf at the fixed depth.
 * Blocks with b % 100 == 37 carry one corrupted node (one flipped bit): every proof through it must reject.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t stream3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t s = seed ^ (a * 0xA24BAED4963EE407ull) ^ (b * 0xD1342543DE82EF95ull) ^ (c * 0x9FB21C651E98DF25ull);
    (void)sm64(&s);
    (void)sm64(&s);
    return s;
}
static void prf32(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint8_t out[32])
{
    uint64_t s = stream3(seed, a, b, c);
    for (int w = 0; w < 4; ++w) {
        uint64_t v = sm64(&s);
        for (int i = 0; i < 8; ++i) out[8 * w + i] = (uint8_t)(v >> (8 * i));
    }
}

typedef struct {
    uint8_t* nodes; uint64_t nodes_len, nodes_cap;
    uint64_t* node_off; uint64_t n_nodes, off_cap;   /* n_nodes+1 entries used */
    uint64_t* index; uint64_t n_index, index_cap;
    uint64_t* first; uint64_t n_proofs, first_cap;   /* n_proofs+1 entries used */
    uint8_t* keys; uint8_t* roots;                   /* 32 * n_proofs, same capacity as first */
} blockbuf;

static void bb_init(blockbuf* b)
{
    memset(b, 0, sizeof *b);
    b->nodes_cap = 1 << 20; b->nodes = malloc(b->nodes_cap);
    b->off_cap = 8192; b->node_off = malloc(8 * b->off_cap); b->node_off[0] = 0;
    b->index_cap = 16384; b->index = malloc(8 * b->index_cap);
    b->first_cap = 2048; b->first = malloc(8 * b->first_cap); b->first[0] = 0;
    b->keys = malloc(32 * b->first_cap); b->roots = malloc(32 * b->first_cap);
}
static uint64_t bb_add_node(blockbuf* b, const uint8_t* rlp, uint32_t len)
{
    if (b->nodes_len + len > b->nodes_cap) { b->nodes_cap *= 2; b->nodes = realloc(b->nodes, b->nodes_cap); }
    if (b->n_nodes + 2 > b->off_cap) { b->off_cap *= 2; b->node_off = realloc(b->node_off, 8 * b->off_cap); }
    memcpy(b->nodes + b->nodes_len, rlp, len);
    b->nodes_len += len;
    b->node_off[++b->n_nodes] = b->nodes_len;
    return b->n_nodes - 1;
}
static void bb_add_proof(blockbuf* b, const uint64_t* chain, uint32_t n, const uint8_t key[32], const uint8_t root[32])
{
    if (b->n_index + n > b->index_cap) { b->index_cap *= 2; b->index = realloc(b->index, 8 * b->index_cap); }
    if (b->n_proofs + 2 > b->first_cap) {
        b->first_cap *= 2;
        b->first = realloc(b->first, 8 * b->first_cap);
        b->keys = realloc(b->keys, 32 * b->first_cap);
        b->roots = realloc(b->roots, 32 * b->first_cap);
    }
    memcpy(b->index + b->n_index, chain, 8 * n);
    b->n_index += n;
    memcpy(b->keys + 32 * b->n_proofs, key, 32);
    memcpy(b->roots + 32 * b->n_proofs, root, 32);
    b->first[++b->n_proofs] = b->n_index;
}
static void bb_free(blockbuf* b)
{
    free(b->nodes); free(b->node_off); free(b->index); free(b->first); free(b->keys); free(b->roots);
}

typedef struct {
    uint8_t key[32];
    const uint8_t* leaf; uint32_t leaf_len; /* encoded leaf node */
    uint64_t chain[16];                     /* node indices root..leaf, filled by build */
} vkey;

#define KNIB(key, k) (((k) & 1) ? ((key)[(k) >> 1] & 15) : ((key)[(k) >> 1] >> 4))

static int cmp_vkey(const void* a, const void* b) { return memcmp(((const vkey*)a)->key, ((const vkey*)b)->key, 32); }

/* union trie over sorted keys [lo, hi) sharing `level` nibbles; `depth` = number of branch levels.  Emits the node
 * (pre-order: a parent gets a smaller index than its children, but the index itself carries no meaning), records its
 * index in every key's chain, returns its hash. */
static void build_virtual(blockbuf* out, uint64_t seed, uint64_t blk, uint64_t trie_id, vkey* k, uint32_t lo, uint32_t hi, uint32_t level,
                          uint32_t depth, uint8_t hash[32])
{
    if (level == depth) { /* one key per leaf by construction */
        uint64_t idx = bb_add_node(out, k[lo].leaf, k[lo].leaf_len);
        k[lo].chain[level] = idx;
        oracle_keccak256(k[lo].leaf, k[lo].leaf_len, hash);
        return;
    }
    uint8_t node[532];
    node[0] = 0xf9; node[1] = 0x02; node[2] = 0x11; node[531] = 0x80;
    /* reserve the index first (pre-order), fill the bytes after the children are known */
    uint64_t idx = bb_add_node(out, node, 532);
    uint64_t my_off = out->node_off[idx];
    /* path prefix as an integer for the PRF */
    uint64_t prefix = 0;
    for (uint32_t i = 0; i < level; ++i) prefix = prefix * 16 + KNIB(k[lo].key, i);
    uint32_t start = lo;
    for (uint32_t c = 0; c < 16; ++c) {
        uint8_t* slot = node + 3 + 33 * c;
        slot[0] = 0xa0;
        uint32_t end = start;
        while (end < hi && (uint32_t)KNIB(k[end].key, level) == c) end++;
        if (end > start) build_virtual(out, seed, blk, trie_id, k, start, end, level + 1, depth, slot + 1);
        else prf32(seed, blk, trie_id, (prefix * 16 + c) * 64 + level + 1, slot + 1);
        start = end;
    }
    memcpy(out->nodes + my_off, node, 532); /* out->nodes may have moved: use the offset */
    for (uint32_t i = lo; i < hi; ++i) k[i].chain[level] = idx;
    oracle_keccak256(node, 532, hash);
}

static uint32_t put_leaf_path(uint8_t* out, const uint8_t key[32], uint32_t from)
{
    uint32_t cnt = 64 - from, o = 0, i = from;
    if (cnt & 1) { out[o++] = (uint8_t)(0x30 | KNIB(key, i)); i++; } else out[o++] = 0x20;
    for (; i < 64; i += 2) out[o++] = (uint8_t)((KNIB(key, i) << 4) | KNIB(key, i + 1));
    return o;
}
/* key with a given distinct prefix of `pn` nibbles, the rest from the stream */
static void make_key(uint64_t* s, uint64_t prefix, uint32_t pn, uint8_t key[32])
{
    for (int w = 0; w < 4; ++w) {
        uint64_t v = sm64(s);
        for (int i = 0; i < 8; ++i) key[8 * w + i] = (uint8_t)(v >> (8 * i));
    }
    for (uint32_t i = 0; i < pn; ++i) {
        uint32_t nb = (uint32_t)(prefix >> (4 * (pn - 1 - i))) & 15;
        if (i & 1) key[i >> 1] = (uint8_t)((key[i >> 1] & 0xf0) | nb);
        else key[i >> 1] = (uint8_t)((key[i >> 1] & 0x0f) | (nb << 4));
    }
}

enum { ACC_DEPTH = 7, STO_DEPTH = 5 }; /* branch levels; proofs have depth+1 nodes */

static void gen_block(uint64_t seed, uint64_t blk, uint32_t txs, blockbuf* out)
{
    bb_init(out);
    const uint32_t n_acc = 2 * txs;
    vkey* acc = calloc(n_acc, sizeof *acc);
    uint8_t (*acc_leaf)[112] = malloc((size_t)n_acc * 112);
    vkey* slots = calloc(2 * txs, sizeof *slots);
    uint8_t (*slot_leaf)[72] = malloc((size_t)2 * txs * 72);
    uint8_t (*sroot)[32] = malloc((size_t)txs * 32); /* storage root of contract B of tx t */
    const uint64_t mul28 = 0x9E3779B1ull, off28 = (blk * 0x632BE5ABull) & 0xfffffff;

    /* storage tries first: their roots go into the account leaves */
    for (uint32_t t = 0; t < txs; ++t) {
        uint64_t s = stream3(seed, blk, 0x5107 + t, 1);
        uint64_t slot_prefix0 = 0;
        for (uint32_t j = 0; j < 2; ++j) {
            vkey* k = &slots[2 * t + j];
            uint64_t p0 = (sm64(&s) >> 11) & 0xfffff, prefix = p0;                       /* 5 random nibbles ... */
            if (j == 1 && p0 == slot_prefix0) prefix = p0 ^ 1;                           /* ... distinct inside the pair */
            if (j == 0) slot_prefix0 = p0;
            make_key(&s, prefix, STO_DEPTH, k->key);
            uint8_t* lf = slot_leaf[2 * t + j];
            uint8_t hp[33];
            uint32_t hpn = put_leaf_path(hp, k->key, STO_DEPTH);
            uint32_t payload = 1 + hpn + 34, o = 0;
            lf[o++] = 0xf8; lf[o++] = (uint8_t)payload;
            lf[o++] = (uint8_t)(0x80 + hpn); memcpy(lf + o, hp, hpn); o += hpn;
            lf[o++] = 0xa1; lf[o++] = 0xa0;
            for (int w = 0; w < 4; ++w) { uint64_t v = sm64(&s); for (int i = 0; i < 8; ++i) lf[o + 8 * w + i] = (uint8_t)(v >> (8 * i)); }
            lf[o] |= 0x80;
            o += 32;
            k->leaf = lf; k->leaf_len = o;
        }
        if (memcmp(slots[2 * t].key, slots[2 * t + 1].key, 32) > 0) { vkey tmp = slots[2 * t]; slots[2 * t] = slots[2 * t + 1]; slots[2 * t + 1] = tmp; }
        build_virtual(out, seed, blk, 0x1000000ull + t, slots, 2 * t, 2 * t + 2, 0, STO_DEPTH, sroot[t]);
    }
    /* accounts: A_t = 2t (storage root = PRF), B_t = 2t+1 (storage root = its trie) */
    for (uint32_t a = 0; a < n_acc; ++a) {
        uint64_t s = stream3(seed, blk, 0xACC0 + a, 2);
        uint64_t prefix = ((uint64_t)a * mul28 + off28) & 0xfffffff; /* 7 nibbles, distinct for a < 2^28 */
        make_key(&s, prefix, ACC_DEPTH, acc[a].key);
        uint8_t* lf = acc_leaf[a];
        uint8_t hp[33];
        uint32_t hpn = put_leaf_path(hp, acc[a].key, ACC_DEPTH); /* 57 nibbles -> 29 bytes */
        uint32_t o = 0;
        lf[o++] = 0xf8; lf[o++] = (uint8_t)(1 + hpn + 80);
        lf[o++] = (uint8_t)(0x80 + hpn); memcpy(lf + o, hp, hpn); o += hpn;
        lf[o++] = 0xb8; lf[o++] = 78; lf[o++] = 0xf8; lf[o++] = 76;
        lf[o++] = (uint8_t)(1 + sm64(&s) % 127);
        uint64_t bal = sm64(&s) | 0x8000000000000000ull;
        lf[o++] = 0x88;
        for (int i = 0; i < 8; ++i) lf[o++] = (uint8_t)(bal >> (8 * (7 - i)));
        lf[o++] = 0xa0;
        if (a & 1) memcpy(lf + o, sroot[a >> 1], 32);
        else prf32(seed, blk, 0x5707, a, lf + o);
        o += 32;
        lf[o++] = 0xa0; prf32(seed, blk, 0xC0DE, a, lf + o); o += 32;
        acc[a].leaf = lf; acc[a].leaf_len = o;
        acc[a].chain[15] = a; /* remember the creation order through the sort */
    }
    qsort(acc, n_acc, sizeof *acc, cmp_vkey);
    uint8_t state_root[32];
    build_virtual(out, seed, blk, 0, acc, 0, n_acc, 0, ACC_DEPTH, state_root);
    /* proofs in transaction order: acct(A), acct(B), slot(B,0), slot(B,1) */
    uint32_t* where = malloc(4 * n_acc);
    for (uint32_t i = 0; i < n_acc; ++i) where[acc[i].chain[15]] = i;
    for (uint32_t t = 0; t < txs; ++t) {
        bb_add_proof(out, acc[where[2 * t]].chain, ACC_DEPTH + 1, acc[where[2 * t]].key, state_root);
        bb_add_proof(out, acc[where[2 * t + 1]].chain, ACC_DEPTH + 1, acc[where[2 * t + 1]].key, state_root);
        bb_add_proof(out, slots[2 * t].chain, STO_DEPTH + 1, slots[2 * t].key, sroot[t]);
        bb_add_proof(out, slots[2 * t + 1].chain, STO_DEPTH + 1, slots[2 * t + 1].key, sroot[t]);
    }
    if (blk % 100 == 37) { /* one corrupted node */
        uint64_t s = stream3(seed, blk, 0xBAD, 3);
        uint64_t bit = sm64(&s) % (8 * out->nodes_len);
        out->nodes[bit >> 3] ^= (uint8_t)(1u << (bit & 7));
    }
    free(where); free(acc); free(acc_leaf); free(slots); free(slot_leaf); free(sroot);
}

/* Generate blocks [first_block, first_block + n_blocks).  All outputs are malloc'd and owned by the caller
 * (oracle_free).  totals[0..3] = unique nodes, node bytes, node references, proofs. */
int oracle_synth_blocks(uint64_t seed, uint64_t first_block, uint32_t n_blocks, uint32_t txs_per_block, int threads, uint8_t** nodes,
                        uint64_t** node_off, uint64_t** node_index, uint64_t** proof_first, uint8_t** keys32, uint8_t** roots32,
                        uint32_t** block_of_proof, uint64_t totals[4])
{
    blockbuf* bb = malloc(sizeof(blockbuf) * n_blocks);
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int64_t b = 0; b < (int64_t)n_blocks; ++b) gen_block(seed, first_block + (uint64_t)b, txs_per_block, &bb[b]);
    uint64_t tn = 0, tb = 0, ti = 0, tp = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) { tn += bb[b].n_nodes; tb += bb[b].nodes_len; ti += bb[b].n_index; tp += bb[b].n_proofs; }
    *nodes = malloc(tb + 64);
    *node_off = malloc(8 * (tn + 1));
    *node_index = malloc(8 * (ti + 1));
    *proof_first = malloc(8 * (tp + 1));
    *keys32 = malloc(32 * tp + 32);
    *roots32 = malloc(32 * tp + 32);
    *block_of_proof = malloc(4 * (tp + 1));
    uint64_t on = 0, ob = 0, oi = 0, op = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        memcpy(*nodes + ob, bb[b].nodes, bb[b].nodes_len);
        for (uint64_t j = 0; j < bb[b].n_nodes; ++j) (*node_off)[on + j] = ob + bb[b].node_off[j];
        for (uint64_t j = 0; j < bb[b].n_index; ++j) (*node_index)[oi + j] = on + bb[b].index[j];
        for (uint64_t j = 0; j < bb[b].n_proofs; ++j) { (*proof_first)[op + j] = oi + bb[b].first[j]; (*block_of_proof)[op + j] = b; }
        memcpy(*keys32 + 32 * op, bb[b].keys, 32 * bb[b].n_proofs);
        memcpy(*roots32 + 32 * op, bb[b].roots, 32 * bb[b].n_proofs);
        on += bb[b].n_nodes; ob += bb[b].nodes_len; oi += bb[b].n_index; op += bb[b].n_proofs;
        bb_free(&bb[b]);
    }
    (*node_off)[tn] = tb;
    (*proof_first)[tp] = ti;
    totals[0] = tn; totals[1] = tb; totals[2] = ti; totals[3] = tp;
    free(bb);
    return 0;
}

void oracle_free(void* p) { free(p); }

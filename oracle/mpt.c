/* mpt.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of phant's `mptize`.
 *
 * Follows src/mpt/mpt.zig:
 *   KeyVal.init (key bytes -> nibbles)                     mpt.zig:19-29
 *   mptize (sorted list -> root.hash())                    mpt.zig:38-45
 *   insertNode (recursive partition by nibble)             mpt.zig:47-119
 *     - one item -> leaf with the remaining nibbles        mpt.zig:54-56
 *     - key exhausted at this level -> branch value        mpt.zig:65-69
 *     - all items share the nibble -> extension over the
 *       longest common prefix                              mpt.zig:83-106
 *     - child reference = raw RLP if < 32 B else keccak    mpt.zig:104,112
 *   node encodings leaf / extension / branch               mpt.zig:170-281
 *   hex-prefix encodeNibbles                               mpt.zig:285-314
 *   empty root constant                                    mpt.zig:10
 * Unlike the reference the nodes are kept in a tree so that proofs can be cut from it.
 */
#include "oracle.h"
#include "rlp.h"
#include <stdlib.h>

enum { K_LEAF = 1, K_EXT = 2, K_BRANCH = 3 };

typedef struct node {
    uint8_t kind;
    uint8_t hashed; /* hash[] valid */
    uint8_t hash[32];
    uint8_t* rlp;
    uint64_t rlp_len;
    const uint8_t* path; /* nibbles (leaf / ext) */
    uint32_t path_len;
    struct node* child[16]; /* branch */
    struct node* next;      /* ext */
} node;

struct oracle_trie {
    uint64_t n;
    uint8_t* nib;      /* all key nibbles */
    uint64_t* nib_off; /* n+1 */
    uint8_t* vals;
    uint64_t* val_off;
    node* root;
    uint8_t root_hash[32];
    uint64_t n_nodes, n_hashed, hashed_bytes;
};

static const uint8_t EMPTY_ROOT[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45,
                                       0xe6, 0x92, 0xc0, 0xf8, 0x6e, 0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c,
                                       0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

#define NIB(t, i) ((t)->nib + (t)->nib_off[i])
#define NLEN(t, i) ((uint32_t)((t)->nib_off[(i) + 1] - (t)->nib_off[i]))

/* hex-prefix encoding (mpt.zig:285-314); returns byte count */
static uint32_t hex_prefix(uint8_t* out, const uint8_t* nib, uint32_t n, int is_leaf)
{
    uint32_t o = 0;
    uint32_t i = 0;
    if (n % 2 == 0) {
        out[o++] = (uint8_t)((is_leaf ? 2 : 0) << 4);
    } else {
        out[o++] = (uint8_t)(((is_leaf ? 3 : 1) << 4) | nib[0]);
        i = 1;
    }
    for (; i < n; i += 2) out[o++] = (uint8_t)((nib[i] << 4) | nib[i + 1]);
    return o;
}

/* reference of a child as it appears inside its parent: raw RLP when < 32 B, else str(hash) */
static uint64_t ref_size(const node* c) { return c->rlp_len < 32 ? c->rlp_len : 33; }
static uint64_t put_ref(uint8_t* out, const node* c)
{
    if (c->rlp_len < 32) { memcpy(out, c->rlp, c->rlp_len); return c->rlp_len; }
    out[0] = 0xa0;
    memcpy(out + 1, c->hash, 32);
    return 33;
}

static void finish(oracle_trie* t, node* nd)
{
    t->n_nodes++;
    if (nd->rlp_len >= 32) {
        oracle_keccak256(nd->rlp, nd->rlp_len, nd->hash);
        nd->hashed = 1;
        t->n_hashed++;
        t->hashed_bytes += nd->rlp_len;
    }
}

static node* make_leaf(oracle_trie* t, const uint8_t* path, uint32_t plen, const uint8_t* val, uint64_t vlen)
{
    node* nd = calloc(1, sizeof *nd);
    nd->kind = K_LEAF;
    nd->path = path;
    nd->path_len = plen;
    uint8_t* hp = malloc(plen / 2 + 2);
    uint32_t hpn = hex_prefix(hp, path, plen, 1);
    uint64_t payload = rlp_str_size(hp, hpn) + rlp_str_size(val, vlen);
    nd->rlp = malloc(rlp_list_hdr_size(payload) + payload);
    uint64_t o = rlp_put_list_hdr(nd->rlp, payload);
    o += rlp_put_str(nd->rlp + o, hp, hpn);
    o += rlp_put_str(nd->rlp + o, val, vlen);
    nd->rlp_len = o;
    free(hp);
    finish(t, nd);
    return nd;
}

static node* insert_node(oracle_trie* t, uint64_t lo, uint64_t hi, uint32_t level)
{
    if (hi == lo) return NULL;
    if (hi - lo == 1)
        return make_leaf(t, NIB(t, lo) + level, NLEN(t, lo) - level, t->vals + t->val_off[lo],
                         t->val_off[lo + 1] - t->val_off[lo]);

    node* bn = calloc(1, sizeof *bn);
    bn->kind = K_BRANCH;
    const uint8_t* bval = NULL;
    uint64_t bval_len = 0;
    uint64_t start = lo;
    while (start < hi) {
        if (level == NLEN(t, start)) { /* key ends here: branch value (mpt.zig:65-69) */
            bval = t->vals + t->val_off[start];
            bval_len = t->val_off[start + 1] - t->val_off[start];
            start++;
            continue;
        }
        uint64_t end = start;
        while (end < hi && NIB(t, end)[level] == NIB(t, start)[level]) end++;

        if (start == lo && end == hi) { /* every item shares this nibble: extension (mpt.zig:83-106) */
            uint32_t pi = level + 1;
            for (;;) {
                if (NLEN(t, lo) == pi) break;
                int stop = 0;
                for (uint64_t k = lo + 1; k < hi; ++k)
                    if (pi == NLEN(t, k) || NIB(t, k)[pi] != NIB(t, lo)[pi]) { stop = 1; break; }
                if (stop) break;
                pi++;
            }
            free(bn);
            node* next = insert_node(t, lo, hi, pi);
            node* en = calloc(1, sizeof *en);
            en->kind = K_EXT;
            en->path = NIB(t, lo) + level;
            en->path_len = pi - level;
            en->next = next;
            uint8_t* hp = malloc(en->path_len / 2 + 2);
            uint32_t hpn = hex_prefix(hp, en->path, en->path_len, 0);
            uint64_t payload = rlp_str_size(hp, hpn) + ref_size(next);
            en->rlp = malloc(rlp_list_hdr_size(payload) + payload);
            uint64_t o = rlp_put_list_hdr(en->rlp, payload);
            o += rlp_put_str(en->rlp + o, hp, hpn);
            o += put_ref(en->rlp + o, next);
            en->rlp_len = o;
            free(hp);
            finish(t, en);
            return en;
        }
        bn->child[NIB(t, start)[level]] = insert_node(t, start, end, level + 1);
        start = end;
    }
    uint64_t payload = rlp_str_size(bval, bval_len);
    for (int s = 0; s < 16; ++s) payload += bn->child[s] ? ref_size(bn->child[s]) : 1;
    bn->rlp = malloc(rlp_list_hdr_size(payload) + payload);
    uint64_t o = rlp_put_list_hdr(bn->rlp, payload);
    for (int s = 0; s < 16; ++s) {
        if (bn->child[s]) o += put_ref(bn->rlp + o, bn->child[s]);
        else bn->rlp[o++] = 0x80;
    }
    o += rlp_put_str(bn->rlp + o, bval, bval_len);
    bn->rlp_len = o;
    finish(t, bn);
    return bn;
}

static void free_node(node* nd)
{
    if (!nd) return;
    for (int s = 0; s < 16; ++s) free_node(nd->child[s]);
    free_node(nd->next);
    free(nd->rlp);
    free(nd);
}

oracle_trie* oracle_trie_build(const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals,
                               const uint64_t* val_off, uint64_t n)
{
    /* strictly sorted, byte-lexicographic with a strict prefix first == nibble order (mpt.zig:31-33,39) */
    for (uint64_t i = 0; i + 1 < n; ++i) {
        uint32_t la = key_off[i + 1] - key_off[i], lb = key_off[i + 2] - key_off[i + 1];
        int c = memcmp(keys + key_off[i], keys + key_off[i + 1], la < lb ? la : lb);
        if (c > 0 || (c == 0 && la >= lb)) return NULL;
    }
    oracle_trie* t = calloc(1, sizeof *t);
    t->n = n;
    t->nib_off = malloc((n + 1) * sizeof(uint64_t));
    t->val_off = malloc((n + 1) * sizeof(uint64_t));
    uint64_t total_key = n ? key_off[n] - key_off[0] : 0;
    t->nib = malloc(2 * total_key + 1);
    uint64_t total_val = n ? val_off[n] - val_off[0] : 0;
    t->vals = malloc(total_val + 1);
    if (total_val) memcpy(t->vals, vals + val_off[0], total_val);
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; ++i) {
        t->nib_off[i] = o;
        t->val_off[i] = val_off[i] - val_off[0];
        for (uint32_t k = key_off[i]; k < key_off[i + 1]; ++k) {
            t->nib[o++] = keys[k] >> 4;
            t->nib[o++] = keys[k] & 0x0f;
        }
    }
    t->nib_off[n] = o;
    t->val_off[n] = total_val;
    t->root = insert_node(t, 0, n, 0);
    if (!t->root) {
        memcpy(t->root_hash, EMPTY_ROOT, 32);
    } else {
        if (!t->root->hashed) { /* the root is always hashed, whatever its size (mpt.zig:42) */
            oracle_keccak256(t->root->rlp, t->root->rlp_len, t->root->hash);
            t->root->hashed = 1;
            t->n_hashed++;
            t->hashed_bytes += t->root->rlp_len;
        }
        memcpy(t->root_hash, t->root->hash, 32);
    }
    return t;
}

void oracle_trie_root(const oracle_trie* t, uint8_t out_root[32]) { memcpy(out_root, t->root_hash, 32); }

void oracle_trie_stats(const oracle_trie* t, uint64_t* n_nodes, uint64_t* n_hashed, uint64_t* hashed_bytes)
{
    if (n_nodes) *n_nodes = t->n_nodes;
    if (n_hashed) *n_hashed = t->n_hashed;
    if (hashed_bytes) *hashed_bytes = t->hashed_bytes;
}

void oracle_trie_free(oracle_trie* t)
{
    if (!t) return;
    free_node(t->root);
    free(t->nib);
    free(t->nib_off);
    free(t->vals);
    free(t->val_off);
    free(t);
}

int oracle_mptize(const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals, const uint64_t* val_off,
                  uint64_t n, uint8_t out_root[32])
{
    oracle_trie* t = oracle_trie_build(keys, key_off, vals, val_off, n);
    if (!t) return -1;
    oracle_trie_root(t, out_root);
    oracle_trie_free(t);
    return 0;
}

int oracle_trie_prove(const oracle_trie* t, const uint8_t* key, uint32_t key_len, uint8_t* out, uint64_t cap,
                      uint64_t* node_off, uint32_t cap_nodes)
{
    uint8_t nib[2 * 256];
    if (key_len > 256) return -1;
    for (uint32_t i = 0; i < key_len; ++i) { nib[2 * i] = key[i] >> 4; nib[2 * i + 1] = key[i] & 15; }
    uint32_t nn = 2 * key_len, pos = 0;
    int count = 0;
    uint64_t o = 0;
    node_off[0] = 0;
    const node* nd = t->root;
    int is_root = 1;
    while (nd) {
        if (is_root || nd->rlp_len >= 32) { /* embedded nodes travel inside their parent */
            if ((uint32_t)count >= cap_nodes || o + nd->rlp_len > cap) return -1;
            memcpy(out + o, nd->rlp, nd->rlp_len);
            o += nd->rlp_len;
            node_off[++count] = o;
        }
        is_root = 0;
        if (nd->kind == K_LEAF) break;
        if (nd->kind == K_EXT) {
            if (nn - pos < nd->path_len || memcmp(nib + pos, nd->path, nd->path_len) != 0) break;
            pos += nd->path_len;
            nd = nd->next;
        } else {
            if (pos == nn) break;
            nd = nd->child[nib[pos++]];
        }
    }
    return count;
}

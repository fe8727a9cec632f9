/* verify.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU statement of the Merkle-Patricia proof walk.
 *
 * PARITY UNPINNED: the reference has no proof verifier.  The hook this fills is the TODO at
 * src/engine_api/execution_payload.zig:177-178 ("reconstruct the proof from the ... execution witness
 * and verify it").  The node encodings walked here are exactly the ones src/mpt/mpt.zig:170-281
 * produces (leaf / extension / branch, hex-prefix paths mpt.zig:285-314, children embedded when their
 * RLP is < 32 B, mpt.zig:104,112); the walk is the yellow-paper trie lookup over those encodings.
 *
 * Rules (the CUDA walker must agree bit for bit; DESIGN.md "Proof walk" repeats them):
 *   R1 node i of the chain must hash to the reference the parent holds for the key's next nibble
 *      (node 0 must hash to the root).  References to children < 32 B are embedded lists and are
 *      walked in place without consuming a chain node.
 *   R2 every node is a strictly canonical RLP list whose payload fills the node exactly and holds
 *      17 items (branch) or 2 items (leaf / extension, told apart by the hex-prefix flag).
 *   R3 terminal outcomes: leaf whose path equals the rest of the key -> PRESENT (value = item 1);
 *      leaf or extension whose path diverges, or an empty branch slot, or an empty branch value
 *      with the key exhausted -> ABSENT.  A terminal must be the last node of the chain, and a
 *      hash reference needs a following node; otherwise REJECT.
 *   R4 an empty chain is ABSENT under the empty root keccak(0x80) (mpt.zig:10) and REJECT otherwise.
 */
#include "oracle.h"
#include "rlp.h"
#include <stdlib.h>

enum { ST_REJECT = 0, ST_PRESENT = 1, ST_ABSENT = 2, ST_MISSING = 3 };

/* Bag mode (oracle_verify_bag): the witness is an unordered SET of nodes (the shape of an execution witness), not a
 * chain per key.  Same rules R1/R2/R3 with these changes: a hash reference is resolved by looking its 32 bytes up
 * among the digests of the bag -- not found = ST_MISSING (the witness is incomplete; not accepted); terminals need
 * not be "the last node" (there is no chain) and unused nodes are harmless. */
typedef struct {
    const uint8_t* digests; /* n * 32, digest of node i */
    const uint32_t* sorted; /* node indices sorted by digest */
    uint64_t n;
} bag_index;

static int64_t bag_find(const bag_index* b, const uint8_t h[32])
{
    uint64_t lo = 0, hi = b->n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        int c = memcmp(b->digests + 32ull * b->sorted[mid], h, 32);
        if (c == 0) return b->sorted[mid];
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return -1;
}

static const uint8_t EMPTY_ROOT[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45,
                                       0xe6, 0x92, 0xc0, 0xf8, 0x6e, 0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c,
                                       0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

static int verify_one(const uint8_t* nodes, const uint64_t* node_off, const uint64_t* node_index, uint64_t first, uint64_t last,
                      const uint8_t* key32, const uint8_t* root32, uint64_t* voff, uint32_t* vlen, const bag_index* bag)
{
    *voff = 0;
    *vlen = 0;
    if (bag) { /* no chain: every "is this the last node" test passes, nothing is ever "left over" */
        first = 0;
        last = 1;
        if (memcmp(root32, EMPTY_ROOT, 32) == 0) return ST_ABSENT;
    } else if (first == last) return memcmp(root32, EMPTY_ROOT, 32) == 0 ? ST_ABSENT : ST_REJECT;

    uint8_t nib[64];
    for (int i = 0; i < 32; ++i) { nib[2 * i] = key32[i] >> 4; nib[2 * i + 1] = key32[i] & 15; }
    uint32_t pos = 0;

    uint8_t expect[32];
    memcpy(expect, root32, 32);
    uint64_t i = first;
    /* current node = nodes[cur .. cur+cur_len); either a chain node or an embedded child */
    const uint8_t* cur = NULL;
    uint64_t cur_len = 0;
    int embedded = 0;

    for (;;) {
        if (!embedded && bag) {
            const int64_t ni = bag_find(bag, expect);
            if (ni < 0) return ST_MISSING;
            cur = nodes + node_off[ni];
            cur_len = node_off[ni + 1] - node_off[ni];
            i = last; /* "we are at the last node" for every terminal test below */
        } else if (!embedded) {
            if (i == last) return ST_REJECT; /* hash reference without a node to resolve it */
            const uint64_t ni = node_index ? node_index[i] : i; /* deduplicated witness: chains hold node indices */
            cur = nodes + node_off[ni];
            cur_len = node_off[ni + 1] - node_off[ni];
            uint8_t h[32];
            oracle_keccak256(cur, cur_len, h);
            if (memcmp(h, expect, 32) != 0) return ST_REJECT;
            i++;
        }
        /* R2: one list filling the node */
        int is_list;
        uint64_t po, pl;
        uint64_t tot = rlp_item(cur, cur_len, &is_list, &po, &pl);
        if (tot == 0 || !is_list || tot != cur_len) return ST_REJECT;
        const uint8_t* pay = cur + po;
        /* index the items */
        uint64_t ioff[17], ipo[17], ipl[17];
        int ilist[17];
        int cnt = 0;
        uint64_t o = 0;
        while (o < pl) {
            if (cnt == 17) return ST_REJECT;
            uint64_t t = rlp_item(pay + o, pl - o, &ilist[cnt], &ipo[cnt], &ipl[cnt]);
            if (t == 0) return ST_REJECT;
            ioff[cnt] = o;
            o += t;
            cnt++;
        }
        if (cnt != 17 && cnt != 2) return ST_REJECT;

        int child; /* index of the item holding the next reference */
        if (cnt == 17) {
            if (pos == 64) { /* key exhausted: the branch value decides */
                if (ilist[16]) return ST_REJECT;
                if (i != last) return ST_REJECT;
                if (ipl[16] == 0) return ST_ABSENT;
                *voff = (uint64_t)(pay + ioff[16] + ipo[16] - nodes);
                *vlen = (uint32_t)ipl[16];
                return ST_PRESENT;
            }
            child = nib[pos++];
        } else {
            if (ilist[0] || ipl[0] == 0) return ST_REJECT;
            const uint8_t* hp = pay + ioff[0] + ipo[0];
            uint32_t flag = hp[0] >> 4;
            if (flag > 3) return ST_REJECT;
            if (!(flag & 1) && (hp[0] & 15)) return ST_REJECT; /* padding nibble must be zero */
            uint32_t plen = (uint32_t)(2 * (ipl[0] - 1) + (flag & 1));
            if (plen > 64) return ST_REJECT;
            uint8_t path[64];
            uint32_t k = 0;
            if (flag & 1) path[k++] = hp[0] & 15;
            for (uint64_t b = 1; b < ipl[0]; ++b) { path[k++] = hp[b] >> 4; path[k++] = hp[b] & 15; }
            int match = (64 - pos >= plen) && memcmp(nib + pos, path, plen) == 0;
            if (flag & 2) { /* leaf */
                if (ilist[1]) return ST_REJECT;
                if (i != last) return ST_REJECT;
                if (match && pos + plen == 64) {
                    *voff = (uint64_t)(pay + ioff[1] + ipo[1] - nodes);
                    *vlen = (uint32_t)ipl[1];
                    return ST_PRESENT;
                }
                return ST_ABSENT;
            }
            if (plen == 0) return ST_REJECT; /* extension with an empty path */
            if (!match) return i == last ? ST_ABSENT : ST_REJECT;
            pos += plen;
            child = 1;
        }
        /* follow the reference */
        if (ilist[child]) { /* embedded node: must be < 32 B in total, walked in place */
            uint64_t tot_child = ipo[child] + ipl[child];
            if (tot_child >= 32) return ST_REJECT;
            cur = pay + ioff[child];
            cur_len = tot_child;
            embedded = 1;
            continue;
        }
        embedded = 0;
        if (ipl[child] == 0) { /* empty slot */
            if (cnt == 2) return ST_REJECT; /* an extension must point somewhere */
            return i == last ? ST_ABSENT : ST_REJECT;
        }
        if (ipl[child] != 32) return ST_REJECT;
        memcpy(expect, pay + ioff[child] + ipo[child], 32);
    }
}

void oracle_verify_proofs(const oracle_proof_batch* in, uint64_t* accept_bitmap, uint8_t* status,
                          uint64_t* val_off, uint32_t* val_len, int threads)
{
    uint64_t n = in->n_proofs;
    if (accept_bitmap) memset(accept_bitmap, 0, ((n + 63) / 64) * 8);
    if (threads < 1) threads = 1;
    /* one 64-proof bitmap word per iteration, so threads never share a word */
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
    for (int64_t w = 0; w < (int64_t)((n + 63) / 64); ++w) {
        uint64_t word = 0;
        for (uint64_t p = (uint64_t)w * 64; p < n && p < (uint64_t)(w + 1) * 64; ++p) {
            const uint8_t* root = in->roots32 + (in->n_roots == 1 ? 0 : 32 * p);
            uint64_t vo;
            uint32_t vl;
            int st = verify_one(in->nodes, in->node_off, in->node_index, in->proof_first[p], in->proof_first[p + 1],
                                in->keys32 + 32 * p, root, &vo, &vl, NULL);
            if (status) status[p] = (uint8_t)st;
            if (val_off) val_off[p] = vo;
            if (val_len) val_len[p] = vl;
            if (st != ST_REJECT) word |= 1ull << (p & 63);
        }
        if (accept_bitmap) accept_bitmap[w] = word;
    }
}

/* qsort context (single-threaded use) */
static const uint8_t* g_sort_digests;
static int cmp_by_digest(const void* a, const void* b)
{
    return memcmp(g_sort_digests + 32ull * *(const uint32_t*)a, g_sort_digests + 32ull * *(const uint32_t*)b, 32);
}

/* Bag verification: nodes CSR (unordered set), n_keys keys, roots32 (n_roots == 1 or n_keys).  status: 0 reject (malformed
 * node on the path), 1 present, 2 absent, 3 a node on the path is missing from the bag. */
void oracle_verify_bag(const uint8_t* nodes, const uint64_t* node_off, uint64_t n_nodes, const uint8_t* keys32, uint64_t n_keys,
                       const uint8_t* roots32, uint64_t n_roots, uint8_t* status, uint64_t* val_off, uint32_t* val_len, int threads)
{
    uint8_t* digests = malloc(32 * n_nodes + 32);
    uint32_t* sorted = malloc(4 * n_nodes + 4);
    oracle_keccak256_batch(nodes, node_off, n_nodes, digests, threads);
    for (uint64_t i = 0; i < n_nodes; ++i) sorted[i] = (uint32_t)i;
    g_sort_digests = digests;
    qsort(sorted, n_nodes, 4, cmp_by_digest);
    bag_index bag = {digests, sorted, n_nodes};
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
    for (int64_t p = 0; p < (int64_t)n_keys; ++p) {
        uint64_t vo;
        uint32_t vl;
        int st = verify_one(nodes, node_off, NULL, 0, 0, keys32 + 32 * p, roots32 + (n_roots == 1 ? 0 : 32 * p), &vo, &vl, &bag);
        status[p] = (uint8_t)st;
        if (val_off) val_off[p] = vo;
        if (val_len) val_len[p] = vl;
    }
    free(digests);
    free(sorted);
}

/* state.c -- TEST INFRASTRUCTURE (see oracle.h).  State root over a flat account table.
 *
 * phant has no StateDB.root() (src/state/statedb.zig:16-194; the check is commented out at
 * src/blockchain/blockchain.zig:83-85).  Semantics follow the vendored evmone helper:
 *   storage trie:  key = keccak(slot32), value = rlp(trim(value32)), zero values skipped
 *                  evmone/test/state/mpt_hash.cpp:15-24   (phant deletes zero slots: statedb.zig:112-119)
 *   account leaf:  key = keccak(addr20), value = rlp([nonce, balance, storage_root, keccak(code)])
 *                  evmone/test/state/mpt_hash.cpp:27-36, integer encodings rlp.hpp:47-67
 * and the trie itself is phant's mptize (mpt.c).  Pinned by the fixtures' genesis / last-block
 * stateRoot fields (tests/golden/fixture_states.json.gz).
 */
#include "oracle.h"
#include "rlp.h"
#include <stdlib.h>

typedef struct { uint8_t key[32]; uint64_t idx; } sort_ent;

static int cmp_ent(const void* a, const void* b) { return memcmp(((const sort_ent*)a)->key, ((const sort_ent*)b)->key, 32); }

/* rlp of a big-endian integer with leading zeros trimmed (rlp.hpp:47-67) */
static uint64_t put_trimmed(uint8_t* out, const uint8_t* be, unsigned n)
{
    unsigned z = 0;
    while (z < n && be[z] == 0) z++;
    return rlp_put_str(out, be + z, n - z);
}

static int storage_root(const oracle_accounts* a, uint64_t acc, uint8_t out[32])
{
    uint64_t lo = a->slot_off[acc], hi = a->slot_off[acc + 1];
    uint64_t cnt = 0;
    sort_ent* ents = malloc((hi - lo + 1) * sizeof *ents);
    for (uint64_t s = lo; s < hi; ++s) {
        const uint8_t* v = a->slot_vals32 + 32 * s;
        int zero = 1;
        for (int i = 0; i < 32; ++i) if (v[i]) { zero = 0; break; }
        if (zero) continue;
        oracle_keccak256(a->slot_keys32 + 32 * s, 32, ents[cnt].key);
        ents[cnt].idx = s;
        cnt++;
    }
    qsort(ents, cnt, sizeof *ents, cmp_ent);
    uint8_t* keys = malloc(32 * cnt + 1);
    uint32_t* koff = malloc((cnt + 1) * sizeof *koff);
    uint8_t* vals = malloc(33 * cnt + 1);
    uint64_t* voff = malloc((cnt + 1) * sizeof *voff);
    uint64_t vo = 0;
    for (uint64_t i = 0; i < cnt; ++i) {
        memcpy(keys + 32 * i, ents[i].key, 32);
        koff[i] = (uint32_t)(32 * i);
        voff[i] = vo;
        vo += put_trimmed(vals + vo, a->slot_vals32 + 32 * ents[i].idx, 32);
    }
    koff[cnt] = (uint32_t)(32 * cnt);
    voff[cnt] = vo;
    int rc = oracle_mptize(keys, koff, vals, voff, cnt, out);
    free(ents); free(keys); free(koff); free(vals); free(voff);
    return rc;
}

int oracle_state_root(const oracle_accounts* a, uint8_t out_root[32])
{
    uint64_t n = a->n_accounts;
    sort_ent* ents = malloc((n + 1) * sizeof *ents);
    for (uint64_t i = 0; i < n; ++i) {
        oracle_keccak256(a->addr20 + 20 * i, 20, ents[i].key);
        ents[i].idx = i;
    }
    qsort(ents, n, sizeof *ents, cmp_ent);
    uint8_t* keys = malloc(32 * n + 1);
    uint32_t* koff = malloc((n + 1) * sizeof *koff);
    uint8_t* vals = malloc(128 * n + 1);
    uint64_t* voff = malloc((n + 1) * sizeof *voff);
    uint64_t vo = 0;
    int rc = 0;
    for (uint64_t i = 0; i < n && rc == 0; ++i) {
        uint64_t acc = ents[i].idx;
        memcpy(keys + 32 * i, ents[i].key, 32);
        koff[i] = (uint32_t)(32 * i);
        voff[i] = vo;
        uint8_t sroot[32], chash[32], body[128], nonce_be[8];
        rc = storage_root(a, acc, sroot);
        oracle_keccak256(a->code + a->code_off[acc], a->code_off[acc + 1] - a->code_off[acc], chash);
        for (int b = 0; b < 8; ++b) nonce_be[b] = (uint8_t)(a->nonce[acc] >> (8 * (7 - b)));
        uint64_t bo = 0;
        bo += put_trimmed(body + bo, nonce_be, 8);
        bo += put_trimmed(body + bo, a->balance32 + 32 * acc, 32);
        bo += rlp_put_str(body + bo, sroot, 32);
        bo += rlp_put_str(body + bo, chash, 32);
        vo += rlp_put_list_hdr(vals + vo, bo);
        memcpy(vals + vo, body, bo);
        vo += bo;
    }
    koff[n] = (uint32_t)(32 * n);
    voff[n] = vo;
    if (rc == 0) rc = oracle_mptize(keys, koff, vals, voff, n, out_root);
    free(ents); free(keys); free(koff); free(vals); free(voff);
    return rc;
}

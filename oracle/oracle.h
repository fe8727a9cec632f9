/* oracle.h -- CPU restatement of phant's trie/hash hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and only as the checker
 * (or as the timed CPU baseline; the development benchmarks under tools/ use it the same way -- input generator, checker,
 * CPU timing -- never as the thing measured).  The product path (phant_b200/csrc -> libphantgpu.so) never links,
 * loads or calls it and fails loudly when the CUDA library is missing.
 *
 * Parity status:
 *   - Keccak-256, mptize (trie root from sorted key/values), state root: PINNED by the reference's
 *     own vectors (see tests/test_oracle_*.py and tests/golden/).
 *   - Proof verification (oracle_verify_proofs): PARITY UNPINNED.  phant has no proof verifier
 *     (reference src/engine_api/execution_payload.zig:177-178 is a TODO).  The walk below is a
 *     restatement of the yellow-paper trie lookup; it is anchored indirectly: proofs are cut from
 *     tries whose roots are pinned by the fixtures.
 *
 * All file:line citations are relative to the reference checkout (/root/reference).
 */
#ifndef PHANT_ORACLE_H
#define PHANT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- Keccak-256 (follows ethash/lib/keccak/keccak.c:301-361; == src/crypto/hasher.zig:4-8) ---- */
typedef void (*oracle_keccak_fn)(const uint8_t* data, size_t len, uint8_t out[32]);
void oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
/* Swap the hash used by every other oracle function (used to plug oracle/_ref's compiled reference
 * keccak.c in, so the CPU baseline times the reference's own arithmetic).  NULL restores the port. */
void oracle_set_keccak(oracle_keccak_fn fn);
/* dlopen oracle/_ref/libref_keccak.so (the reference's keccak.c compiled unchanged) and route every
 * oracle hash through it.  0 on success, negative if the library is absent. */
int oracle_use_ref_keccak(const char* so_path);
/* CSR batch: msg i = msgs[off[i] .. off[i+1]).  threads<=1 -> serial. */
void oracle_keccak256_batch(const uint8_t* msgs, const uint64_t* off, uint64_t n, uint8_t* out32, int threads);

/* logs bloom (src/types/receipt.zig:37-63): item i sets 3 bits in bloom bloom_of_item[i]; blooms = n_blooms*256 bytes */
void oracle_logs_bloom(const uint8_t* items, const uint64_t* item_off, const uint32_t* bloom_of_item, uint64_t n_items,
                       uint64_t n_blooms, uint8_t* blooms);

/* ---- secp256k1 public-key recovery (oracle/secp256k1.c): what ecdsa.Signer.erecover returns (src/crypto/ecdsa.zig:19-21,
 * reached from TxSigner.get_sender, src/signer/signer.zig:78).  sig65 = r || s || recid.  0 = pub65 holds 0x04 || X || Y;
 * negative = the signature recovers no key.  The arithmetic lives in libsecp256k1 behind zig-eth-secp256k1 @ 95b7f93
 * (build.zig.zon), absent from the tree: restated from SEC 1 section 4.1.6, pinned by ecdsa.zig:38-48 and signer.zig:199-227. */
int oracle_ecrecover(const uint8_t hash32[32], const uint8_t sig65[65], uint8_t pub65[65]);
int oracle_secp256k1_pubkey(const uint8_t priv32[32], uint8_t pub65[65]);

/* ---- mptize (follows src/mpt/mpt.zig:38-314) ----
 * keys: byte strings sorted lexicographically (a strict prefix sorts first), CSR; values CSR.
 * Returns 0, or -1 if keys are not strictly sorted (the reference asserts sortedness, mpt.zig:39). */
int oracle_mptize(const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals, const uint64_t* val_off,
                  uint64_t n, uint8_t out_root[32]);

/* Trie handle: same construction, nodes kept so proofs can be cut from it. */
typedef struct oracle_trie oracle_trie;
oracle_trie* oracle_trie_build(const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals,
                               const uint64_t* val_off, uint64_t n);
void oracle_trie_root(const oracle_trie*, uint8_t out_root[32]);
/* Census: number of nodes, how many are hashed (RLP >= 32 B or root), total RLP bytes of hashed nodes. */
void oracle_trie_stats(const oracle_trie*, uint64_t* n_nodes, uint64_t* n_hashed, uint64_t* hashed_bytes);
/* Cut the proof for `key` (key_len bytes): the RLP of every *hashed* node on the lookup path, root first,
 * concatenated into out (cap bytes) with node_off[0..n_nodes] byte offsets (cap_nodes+1 entries).
 * Returns number of nodes, or -1 if a buffer is too small.  Works for present and absent keys. */
int oracle_trie_prove(const oracle_trie*, const uint8_t* key, uint32_t key_len, uint8_t* out, uint64_t cap,
                      uint64_t* node_off, uint32_t cap_nodes);
void oracle_trie_free(oracle_trie*);

/* ---- state root (semantics: evmone/test/state/mpt_hash.cpp:15-36, rlp.hpp:47-67; data model
 *      src/state/statedb.zig:16-30, src/state/types.zig:7-33) ----
 * Flat SoA accounts: addr20[n*20], nonce[n], balance32[n*32] (big endian), code CSR,
 * storage CSR of (slot32, value32) pairs per account; zero values are skipped (statedb.zig:112-119). */
typedef struct {
    uint64_t n_accounts;
    const uint8_t* addr20;
    const uint64_t* nonce;
    const uint8_t* balance32;
    const uint8_t* code;        const uint64_t* code_off;     /* n+1 */
    const uint8_t* slot_keys32; const uint8_t* slot_vals32;   const uint64_t* slot_off; /* n+1, in slots */
} oracle_accounts;
int oracle_state_root(const oracle_accounts* a, uint8_t out_root[32]);

/* ---- proof verification (spec walk; see header note: parity unpinned) ----
 * Status per proof: 0 reject, 1 present (value slice returned), 2 proven absent. */
typedef struct {
    uint64_t n_proofs;
    const uint8_t* nodes;        /* concatenated node RLP */
    const uint64_t* node_off;    /* n_nodes+1 byte offsets */
    const uint64_t* proof_first; /* n_proofs+1: proof p uses nodes [first[p], first[p+1]) */
    const uint8_t* keys32;       /* n_proofs*32 */
    const uint8_t* roots32;      /* n_roots*32 */
    uint64_t n_roots;            /* 1 (broadcast) or n_proofs */
    const uint64_t* node_index;  /* NULL, or deduplicated witness: proof p walks nodes node_index[proof_first[p] ..] */
} oracle_proof_batch;
/* status[n_proofs] (may be NULL), bitmap = ceil(n/64) words (may be NULL),
 * val_off/val_len[n_proofs] (may be NULL): byte slice of `nodes` holding the proven value. */
void oracle_verify_proofs(const oracle_proof_batch* in, uint64_t* accept_bitmap, uint8_t* status,
                          uint64_t* val_off, uint32_t* val_len, int threads);

/* Bag verification: the witness is an unordered set of nodes (execution-witness shape); references are resolved by
 * digest lookup.  status 0 reject, 1 present, 2 absent, 3 node missing from the bag.  See verify.c. */
void oracle_verify_bag(const uint8_t* nodes, const uint64_t* node_off, uint64_t n_nodes, const uint8_t* keys32, uint64_t n_keys,
                       const uint8_t* roots32, uint64_t n_roots, uint8_t* status, uint64_t* val_off, uint32_t* val_len, int threads);

/* ---- complete-trie dirty-frontier update (config C4) ---- */
typedef struct oracle_ctrie oracle_ctrie;
/* depth = number of branch levels L (leaves = 16^L); untouched leaf hashes come from the PRNG. */
oracle_ctrie* oracle_ctrie_open(uint32_t depth, uint64_t seed);
void oracle_ctrie_root(const oracle_ctrie*, uint8_t out_root[32]);
/* keys32: n_dirty*32 (leaf index = first `depth` nibbles); leaf_vals CSR: value bytes (account RLP). */
void oracle_ctrie_update(oracle_ctrie*, const uint8_t* keys32, const uint8_t* leaf_vals, const uint32_t* val_off,
                         uint64_t n_dirty, uint8_t out_root[32]);
void oracle_ctrie_free(oracle_ctrie*);

/* ---- synthetic workloads; byte-identical twins of phant_b200/csrc/synth.cu ---- */
#define PHANT_SYNTH_SEED 0x5048414E54ull /* "PHANT" */
/* C2: account proofs, `depth` nodes each (depth-1 full branches + 112-byte leaf). */
uint64_t oracle_synth_c2_bytes_per_proof(uint32_t depth);
void oracle_synth_c2(uint64_t seed, uint64_t first_index, uint64_t n, uint32_t depth, int corrupt,
                     uint8_t* nodes, uint64_t* node_off, uint64_t* proof_first, uint8_t* keys32, uint8_t* roots32,
                     int threads);
/* C3: storage proofs, depth 4..12; sizes pass first (bytes[i], nodes[i] per proof), then fill. */
void oracle_synth_c3_sizes(uint64_t seed, uint64_t first_index, uint64_t n, uint32_t* n_nodes, uint32_t* n_bytes);
void oracle_synth_c3(uint64_t seed, uint64_t first_index, uint64_t n, int corrupt, uint8_t* nodes,
                     uint64_t* node_off, uint64_t* proof_first, uint8_t* keys32, uint8_t* roots32, int threads);

/* C5: deduplicated block witnesses (oracle/blocks.c).  Outputs are malloc'd; release with oracle_free. */
int oracle_synth_blocks(uint64_t seed, uint64_t first_block, uint32_t n_blocks, uint32_t txs_per_block, int threads, uint8_t** nodes,
                        uint64_t** node_off, uint64_t** node_index, uint64_t** proof_first, uint8_t** keys32, uint8_t** roots32,
                        uint32_t** block_of_proof, uint64_t totals[4]);
void oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif

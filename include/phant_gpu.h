/* phant_gpu.h -- C ABI of libphantgpu.so: phant's trie/hash hot path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  phant has no plugin interface for this path -- `mptize`
 * and `keccak256` are ordinary Zig functions -- so each entry point below names the reference
 * function or hook it stands behind.  Style follows the one C plugin ABI phant already links, EVMC
 * (evmone/evmc/include/evmc/evmc.h:1068-1126: version probe, create/destroy, plain structs).
 * INTEGRATION.md shows the Zig `@cImport` binding and the build.zig lines a maintainer adds.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary, every size explicit, caller owns every buffer;
 *   - return 0 = OK, negative = error (phant_gpu_strerror).  Accept / reject of a proof is DATA,
 *     never an error code;
 *   - pointers are HOST pointers and the library copies host<->device, unless the context was
 *     switched to device pointers with PHANT_GPU_FLAG_DEVICE_PTRS (benchmarks, callers that already
 *     hold witnesses in HBM).  Device-pointer inputs must be 16-byte aligned and the byte buffers
 *     (msgs / nodes) must be readable for 16 bytes past their last offset.  Host inputs are
 *     validated (monotone offsets, index ranges) before anything is launched; device-pointer inputs
 *     are TRUSTED to be well-formed CSR -- the node CONTENTS are never trusted in either mode.  Device-pointer calls are
 *     ASYNCHRONOUS on the context's stream (its own non-blocking stream unless phant_gpu_set_stream gave it the caller's):
 *     ordering against the caller's own kernels that produce the inputs / consume the outputs is the caller's job --
 *     share the stream, or synchronise on both sides;
 *   - a context owns one device, one stream and its scratch memory; it is not re-entrant: one
 *     context per host thread, or lock around it (phant calls runBlock from httpz worker threads,
 *     src/main.zig:143-149);
 *   - there is NO CPU fallback inside the library: without a usable CUDA device create() fails with
 *     PHANT_GPU_E_NO_DEVICE and the Zig caller keeps its own CPU path.
 */
#ifndef PHANT_GPU_H
#define PHANT_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PHANT_GPU_ABI_VERSION 2 /* 2: multi-GPU entry points (phant_gpu_comm_*, *_sharded), U kind 1 */

enum {
    PHANT_GPU_OK = 0,
    PHANT_GPU_E_INVALID = -1,   /* bad argument (null pointer, unsorted keys, offsets not monotone ...) */
    PHANT_GPU_E_NO_DEVICE = -2, /* no CUDA device / driver */
    PHANT_GPU_E_OOM = -3,       /* device or pinned-host allocation failed */
    PHANT_GPU_E_CUDA = -4,      /* any other CUDA runtime error (phant_gpu_last_error has the text) */
    PHANT_GPU_E_COMM = -5,      /* NCCL not loadable / communicator missing / a collective failed (phant_gpu_last_error) */
    PHANT_GPU_E_MALFORMED = -6  /* malformed RLP in a *builder* input (never used for proofs) */
};

enum {
    PHANT_GPU_FLAG_DEVICE_PTRS = 1u << 0, /* all data pointers are device pointers; no copies */
    /* Keccak kernel choice (default = staged thread-per-sponge).  See DESIGN.md "Keccak kernels". */
    PHANT_GPU_FLAG_KECCAK_DIRECT = 1u << 4, /* thread-per-sponge, direct global loads */
    PHANT_GPU_FLAG_KECCAK_WARP = 1u << 5,   /* one warp per sponge (north-star layout; slower) */
    PHANT_GPU_FLAG_NO_BINNING = 1u << 6     /* do not regroup messages by absorb-block count */
};

typedef struct phant_gpu_ctx phant_gpu_ctx;

typedef struct {
    int32_t device;      /* CUDA device ordinal */
    uint32_t flags;      /* PHANT_GPU_FLAG_* */
    uint64_t reserved[4];
} phant_gpu_config;

/* Environment overrides read by phant_gpu_create: PHANT_GPU_DEVICE (replaces cfg->device), PHANT_GPU_FLAGS (decimal / 0x..,
 * OR-ed into cfg->flags), PHANT_GPU_NCCL_LIB (NCCL library name for the multi-GPU entry points). */
int phant_gpu_abi_version(void);
int phant_gpu_create(phant_gpu_ctx** out, const phant_gpu_config* cfg);
void phant_gpu_destroy(phant_gpu_ctx* ctx);
int phant_gpu_set_flags(phant_gpu_ctx* ctx, uint32_t flags);
/* Run on the caller's CUDA stream (a cudaStream_t passed as void*) instead of the context's own; NULL
 * restores the private stream.  With device pointers every call is then asynchronous on that stream,
 * so the caller can order it against its own kernels / NCCL collectives without a host sync. */
int phant_gpu_set_stream(phant_gpu_ctx* ctx, void* cuda_stream);
const char* phant_gpu_strerror(int code);
const char* phant_gpu_last_error(const phant_gpu_ctx* ctx); /* text of the last CUDA error on this context */

/* Per-call counters, reset by the caller: what bench.py reports as gpu_launches / h2d / d2h and the
 * device time of the dominant kernels (CUDA events on the context's stream). */
typedef struct {
    uint64_t launches;      /* kernels launched by this library */
    uint64_t h2d_bytes, d2h_bytes;
    double keccak_ms;       /* accumulated device time of the batched Keccak kernels */
    double walk_ms;         /* accumulated device time of the proof-walk kernel */
    uint64_t keccak_msgs, keccak_bytes, keccak_perms; /* work the Keccak kernels were given */
    uint64_t reserved[4];
} phant_gpu_stats;
int phant_gpu_get_stats(phant_gpu_ctx* ctx, phant_gpu_stats* out);
int phant_gpu_reset_stats(phant_gpu_ctx* ctx);
int phant_gpu_synchronize(phant_gpu_ctx* ctx);

/* K -- batched Keccak-256.  Replaces hasher.keccak256 (src/crypto/hasher.zig:4-8) for many inputs at
 * once: message i = msgs[off[i] .. off[i+1]) (CSR byte offsets, any alignment, any length incl. 0);
 * out = n*32 digest bytes.  keccak256WithPrefix (hasher.zig:10-17) is the same call on prefix||data. */
int phant_gpu_keccak256_batch(phant_gpu_ctx* ctx, const uint8_t* msgs, const uint64_t* off, uint64_t n, uint8_t* out);
/* The same for DEVICE pointers when the caller knows total_bytes = off[n] - off[0]: nothing is read back, the call is
 * asynchronous on the context's stream (the entry point above needs one host synchronisation to learn the total). */
int phant_gpu_keccak256_batch_async(phant_gpu_ctx* ctx, const uint8_t* msgs, const uint64_t* off, uint64_t n, uint64_t total_bytes, uint8_t* out);

/* M -- == mptize (src/mpt/mpt.zig:38-45).  Keys are byte strings sorted lexicographically (a strict
 * prefix first: KeyVal.lessThan, mpt.zig:31-33), CSR; values CSR.  n == 0 -> empty_mpt_root
 * (mpt.zig:10).  Unsorted or duplicate keys -> PHANT_GPU_E_INVALID (the reference asserts, mpt.zig:39).
 * In device-pointer mode only out_root stays a host pointer. */
int phant_gpu_mpt_root(phant_gpu_ctx* ctx, const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals,
                       const uint64_t* val_off, uint64_t n, uint8_t out_root[32]);

/* M (batched) -- many independent tries in one call: trie t holds items [seg_off[t], seg_off[t+1]) of the CSR arrays,
 * each segment sorted like a single mptize input.  All tries are built together as one forest (one set of launches
 * whatever n_tries is): e.g. the transaction / receipt / withdrawal tries (src/blockchain/blockchain.zig:200-203)
 * of a whole range of blocks.  out_roots = n_tries * 32 bytes; an empty segment gives empty_mpt_root. */
int phant_gpu_mpt_roots(phant_gpu_ctx* ctx, const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals, const uint64_t* val_off,
                        const uint32_t* seg_off, uint64_t n_tries, uint8_t* out_roots);

/* S -- state root of a flat account table: the body of the missing StateDB.root()
 * (hook: src/blockchain/blockchain.zig:83-85; data model src/state/statedb.zig:16-30,
 * src/state/types.zig:7-33).  Trie contents as in evmone/test/state/mpt_hash.cpp:15-36:
 * key keccak(addr), value rlp([nonce, balance, storage_root, keccak(code)]); storage key
 * keccak(slot), value rlp(trim(value)), zero values skipped (statedb.zig:112-119). */
typedef struct {
    uint64_t n_accounts;
    const uint8_t* addr20;      /* n*20 */
    const uint64_t* nonce;      /* n */
    const uint8_t* balance32;   /* n*32 big endian */
    const uint8_t* code;        /* concatenated code bytes */
    const uint64_t* code_off;   /* n+1 */
    const uint8_t* slot_keys32; /* total_slots*32 */
    const uint8_t* slot_vals32; /* total_slots*32 */
    const uint64_t* slot_off;   /* n+1, in slots */
} phant_gpu_accounts;
int phant_gpu_state_root(phant_gpu_ctx* ctx, const phant_gpu_accounts* accounts, uint8_t out_root[32]);

/* S, sharded across GPUs by the top nibble of keccak(addr): for the accounts handed in, out_roots[v] =
 * hash of the subtree hanging under slot v of the ROOT branch (tries built from key nibble 1 on), bit v of *out_mask set
 * when slot v is populated (unpopulated slots are zero-filled).  Each rank passes the accounts whose top nibble it owns;
 * after one all-gather of 16 x 32 bytes + the masks, every rank hashes the root branch rlp([ref_0 .. ref_15, ""])
 * itself (one K call).  With fewer than two populated slots overall the root is not a branch: the caller falls back to
 * phant_gpu_state_root on the one rank that holds every account.  An account leaf is >= 70 bytes, so a populated slot's
 * reference is always a hash (mpt.zig:104,112 inlining cannot apply). */
int phant_gpu_state_subtree_roots(phant_gpu_ctx* ctx, const phant_gpu_accounts* accounts, uint8_t out_roots[16 * 32], uint32_t* out_mask);

/* V -- batched Merkle-Patricia proof verification: the body of the TODO at
 * src/engine_api/execution_payload.zig:177-178.  Proof p = nodes [proof_first[p], proof_first[p+1]),
 * root first; node j = nodes[node_off[j] .. node_off[j+1]).  n_roots == 1 broadcasts one root.
 * Outputs (each may be NULL): accept_bitmap ceil(n/64) words, bit p = proof p accepted;
 * status[p] 0 reject / 1 present / 2 proven absent; val_off/val_len[p] = slice of `nodes` holding
 * the proven value.  Walk rules: DESIGN.md "Proof walk". */
typedef struct {
    uint64_t n_proofs;
    const uint8_t* nodes;
    const uint64_t* node_off;    /* n_nodes+1 */
    const uint64_t* proof_first; /* n_proofs+1 */
    const uint8_t* keys32;       /* n_proofs*32 */
    const uint8_t* roots32;      /* n_roots*32 */
    uint64_t n_roots;
    uint64_t n_nodes;            /* = proof_first[n_proofs]; may be 0 = "read it from the arrays" (host pointers: */
    uint64_t nodes_bytes;        /* = node_off[n_nodes];      always derived; device pointers: costs a sync)      */
    /* Deduplicated witness (optional): when node_index is not NULL, `nodes` holds every DISTINCT node once
     * (n_nodes of them -- required), each is hashed once, and proof p walks the nodes
     * node_index[proof_first[p] .. proof_first[p+1]).  NULL = chains are contiguous runs of `nodes`. */
    const uint64_t* node_index;
} phant_gpu_proof_batch;
int phant_gpu_verify_proofs(phant_gpu_ctx* ctx, const phant_gpu_proof_batch* in, uint64_t* accept_bitmap,
                            uint8_t* status, uint64_t* val_off, uint32_t* val_len);

/* W -- witness given as an unordered SET of nodes (the shape of an execution witness: `state: [node, ...]`), not as
 * one chain per key.  Every node is hashed once, a device hash table maps digest -> node, and each key is walked
 * from its root resolving every hash reference through the table.  Same node rules as V (R1, R2, R3 without the
 * chain-position clauses); status gets a fourth value: 3 = a node on the key's path is not in the set (incomplete
 * witness).  accept bit = status 1 or 2.  Nodes not on any path are harmless. */
typedef struct {
    uint64_t n_nodes;
    const uint8_t* nodes;
    const uint64_t* node_off; /* n_nodes+1 */
    uint64_t nodes_bytes;     /* = node_off[n_nodes] (device pointers: required; host pointers: derived) */
    uint64_t n_keys;
    const uint8_t* keys32;    /* n_keys*32 */
    const uint8_t* roots32;   /* n_roots*32 */
    uint64_t n_roots;         /* 1 (broadcast) or n_keys */
} phant_gpu_witness;
int phant_gpu_verify_witness(phant_gpu_ctx* ctx, const phant_gpu_witness* in, uint64_t* accept_bitmap, uint8_t* status,
                             uint64_t* val_off, uint32_t* val_len);

/* B -- logs blooms (row N3 of SURVEY.md 8f): Receipt.calculateLogsBloom / addToBloom
 * (src/types/receipt.zig:37-63) for many receipts at once.  items = the bloom inputs (log addresses 20 B, topics
 * 32 B, ...) CSR; bloom_of_item[i] says which of the n_blooms 2048-bit filters item i belongs to.  Each item is
 * hashed (same batched Keccak kernel) and sets 3 bits: for j in 0..2, bit 0x7ff - (be16(hash[2j..2j+2]) & 0x7ff),
 * most significant bit of byte 0 first.  blooms = n_blooms * 256 bytes (zeroed by the call). */
int phant_gpu_logs_bloom(phant_gpu_ctx* ctx, const uint8_t* items, const uint64_t* item_off, const uint32_t* bloom_of_item,
                         uint64_t n_items, uint64_t n_blooms, uint8_t* blooms);

/* R -- batched sender recovery (row N4 of SURVEY.md 8f): for each i, the secp256k1 public key that signed hashes32[i]
 * under sigs65[i] = r(32, big endian) || s(32) || recid(1), and its address keccak256(X || Y)[12..32] -- the tail of
 * TxSigner.get_sender (src/signer/signer.zig:78-79: ecdsa_signer.erecover, src/crypto/ecdsa.zig:19-21, then
 * hasher.keccak256(pubkey[1..])[12..]).  Decisions are those of libsecp256k1's secp256k1_ecdsa_recover, which the
 * reference calls through zig-eth-secp256k1: ok[i] = 1 and pubkeys65[i] = 0x04 || X || Y, or ok[i] = 0 (r or s zero or
 * >= n, recid > 3, x not on the curve, result at infinity) with zero-filled outputs.  The low-s rule and the EIP-155 `v`
 * decoding are the caller's (signer.zig:41-76, ecdsa.zig:28-36), as in the reference.  pubkeys65 / addresses20 may be
 * NULL.  A failed recovery is data, not an error code. */
int phant_gpu_ecrecover_batch(phant_gpu_ctx* ctx, const uint8_t* hashes32, const uint8_t* sigs65, uint64_t n,
                              uint8_t* pubkeys65, uint8_t* addresses20, uint8_t* ok);

/* U -- resident trie + dirty-frontier root recompute (BASELINE.json "state-root recompute").  Hook: StateDB.root() after
 * a block (src/blockchain/blockchain.zig:83-85).  Two kinds:
 *   kind 0  a complete 16-ary trie with `depth` branch levels (16^depth leaves) whose untouched leaf hashes come from the
 *           synthetic PRNG (the benchmark shape); update rewrites n_dirty leaves at DISTINCT leaf positions (a collision is
 *           refused with PHANT_GPU_E_INVALID and leaves the trie untouched) and re-hashes only the dirty frontier;
 *   kind 1  a SPARSE secure trie over arbitrary 32-byte keys (keccak(address) / keccak(slot)) with arbitrary values, initially
 *           empty: phant_gpu_trie_update is an UPSERT of (key, value) pairs in any order, an empty value DELETES the key
 *           (absent keys are ignored), the same key twice in one call is PHANT_GPU_E_INVALID.  Resident on the device: the
 *           sorted key table, the values, and the references of a dense top of L = floor(log16(n / 16)) nibble levels; an
 *           update re-builds only the buckets (keys sharing an L-nibble prefix) that hold a dirty key -- with all of mptize's
 *           rules: extensions, embedded nodes -- and re-hashes the dirty part of the dense levels.  Host pointers only.
 *           The root always equals mptize over the current key set (src/mpt/mpt.zig:38-45). */
typedef struct phant_gpu_trie phant_gpu_trie;
typedef struct {
    uint32_t kind;  /* 0 = complete synthetic trie, 1 = sparse secure trie */
    uint32_t depth; /* kind 0: branch levels; kind 1: ignored */
    uint64_t seed;
    uint64_t reserved[4];
} phant_gpu_trie_desc;
int phant_gpu_trie_open(phant_gpu_ctx* ctx, const phant_gpu_trie_desc* desc, phant_gpu_trie** out);
int phant_gpu_trie_root(phant_gpu_trie* trie, uint8_t out_root[32]);
int phant_gpu_trie_update(phant_gpu_trie* trie, const uint8_t* keys32, const uint8_t* leaf_vals,
                          const uint32_t* val_off, uint64_t n_dirty, uint8_t out_root[32]);
void phant_gpu_trie_close(phant_gpu_trie* trie);

/* ---- multi-GPU: proofs shard by contiguous index range, one context per GPU, the only exchange is the
 * accept bitmap.  The reference runs block processing on httpz worker threads (src/main.zig:143-149): either one process
 * with one context + host thread per GPU (phant_gpu_comm_init_local), or one process per GPU (phant_gpu_comm_init with an
 * id from rank 0, carried by whatever channel the host has).  NCCL is loaded at run time (libnccl.so.2, override with
 * PHANT_GPU_NCCL_LIB); single-GPU users never touch it.  Failures return PHANT_GPU_E_COMM. ---- */
#define PHANT_GPU_COMM_ID_BYTES 128
int phant_gpu_comm_get_unique_id(uint8_t id[PHANT_GPU_COMM_ID_BYTES]);                            /* rank 0, then distribute */
int phant_gpu_comm_init(phant_gpu_ctx* ctx, const uint8_t id[PHANT_GPU_COMM_ID_BYTES], int rank, int world); /* collective */
int phant_gpu_comm_init_local(phant_gpu_ctx** ctxs, int n); /* one process: n contexts on n devices, rank i = ctxs[i]; afterwards
                                                               drive each context from its own host thread */
int phant_gpu_comm_info(const phant_gpu_ctx* ctx, int* rank, int* world, int* nccl_version);
/* Optional PEER TRANSPORT for the gathered accept bitmap (same node, NVLink): a collective call that maps one small symmetric
 * region of every rank into every other rank (cudaIpc; one process per GPU -- contexts that share a process keep NCCL).  Afterwards
 * device-pointer calls of phant_gpu_verify_proofs_sharded with equal, 64-aligned shards of at most max_n_global proofs need
 * no collective launch: the walk kernel's epilogue stores each ballot word straight into every rank's gathered bitmap and
 * publishes the step, and the comm stream only waits for the peers' words and copies the bitmap out.  Everything else keeps
 * using NCCL.  PHANT_GPU_E_COMM when a mapping is not possible on some rank: nothing changes, NCCL stays in use.  A rank
 * that stops answering makes the waiting kernels give up after 4 s; phant_gpu_comm_peer_status reports it. */
int phant_gpu_comm_enable_peer(phant_gpu_ctx* ctx, uint64_t max_n_global);
int phant_gpu_comm_disable_peer(phant_gpu_ctx* ctx); /* collective: unmap, back to the NCCL gather */
int phant_gpu_comm_peer_status(phant_gpu_ctx* ctx, int* enabled, uint64_t* steps, int* timed_out);
int phant_gpu_comm_fence(phant_gpu_ctx* ctx);   /* the context's stream waits (on the device) for every collective issued so far */
int phant_gpu_comm_destroy(phant_gpu_ctx* ctx); /* also done by phant_gpu_destroy */
/* Rank r of `world` owns proofs [lo, hi): contiguous, every boundary but the last a multiple of 64 so that bitmap words
 * are disjoint.  A gathered bitmap has phant_gpu_sharded_bitmap_words(n, world) words (>= ceil(n/64): equal slices). */
int phant_gpu_shard_range(uint64_t n, int rank, int world, uint64_t* lo, uint64_t* hi);
uint64_t phant_gpu_sharded_bitmap_words(uint64_t n, int world);
/* V across GPUs: `local` holds THIS rank's shard of a batch of n_global proofs (local->n_proofs == hi - lo); on return
 * global_bitmap holds every rank's accept bits (bit p = proof p of the global batch); status / val_* cover the local shard.
 * Host pointers: synchronous.  Device pointers: Keccak + walk are enqueued on the context's stream and ONE ncclAllGather on
 * the context's comm stream behind an event -- the next call's Keccak overlaps it; the walk that next writes the same
 * global_bitmap buffer waits for it on the device (alternate two buffers and nothing ever waits); phant_gpu_comm_fence or
 * phant_gpu_synchronize before reading. */
int phant_gpu_verify_proofs_sharded(phant_gpu_ctx* ctx, const phant_gpu_proof_batch* local, uint64_t n_global,
                                    uint64_t* global_bitmap, uint8_t* status, uint64_t* val_off, uint32_t* val_len);
/* Per-block verdicts when BLOCKS are sharded (BASELINE.json "1000 blocks x 300 tx"): counts[b] = proofs of block b that
 * were not accepted (status 0 or 3), summed over ranks with one all-reduce; block b is valid iff counts[b] == 0. */
int phant_gpu_block_reject_counts(phant_gpu_ctx* ctx, const uint8_t* status, const uint32_t* block_of_proof, uint64_t n_proofs,
                                  uint64_t n_blocks, uint32_t* counts);
/* S across GPUs: rank r passes the accounts whose top nibble of keccak(address) it owns (phant_gpu_nibble_owner: contiguous
 * slot ranges); subtree roots are built locally, ONE all-gather of 528 bytes per rank follows and every rank hashes the
 * root branch itself.  Fewer than two populated slots overall: the one rank holding accounts computes the plain root and
 * broadcasts it.  Host pointers only. */
int phant_gpu_nibble_owner(int nibble, int world);
int phant_gpu_state_root_sharded(phant_gpu_ctx* ctx, const phant_gpu_accounts* mine, uint8_t out_root[32]);

/* Synthetic witnesses generated on the device (byte-identical to oracle/synth.c).
 * All pointers are DEVICE pointers regardless of the context flags.  which: 2 = account proofs of
 * `depth` nodes (config C2), 3 = storage proofs depth 4..12 (config C3; depth ignored).
 * phant_gpu_synth_sizes fills host totals so the caller can allocate. */
int phant_gpu_synth_sizes(phant_gpu_ctx* ctx, int which, uint64_t seed, uint64_t first_index, uint64_t n,
                          uint32_t depth, uint64_t* total_nodes, uint64_t* total_bytes);
int phant_gpu_synth(phant_gpu_ctx* ctx, int which, uint64_t seed, uint64_t first_index, uint64_t n, uint32_t depth,
                    int corrupt, uint8_t* nodes, uint64_t* node_off, uint64_t* proof_first, uint8_t* keys32,
                    uint8_t* roots32);

#ifdef __cplusplus
}
#endif
#endif /* PHANT_GPU_H */

// trie.cu -- trie builders on the device: M (== mptize), S (state root), U (resident complete trie).
//
// M restates src/mpt/mpt.zig:38-314 as a level-synchronous FOREST builder over sorted keys:
//   top-down   every (extension+)branch unit is a key range [lo, hi) of the sorted list; its branch depth is the
//              common prefix of the range (mpt.zig:83-106), a key that ends there is the branch value
//              (mpt.zig:65-69), and the 16 children are found by binary search on the next nibble
//              (mpt.zig:72-79) -- 16 threads per unit, one BFS level per launch;
//   bottom-up  leaves first, then unit levels deepest-first: RLP is written into a scratch arena
//              (leaf mpt.zig:255-281, branch :218-247, extension :180-209, hex-prefix :285-314) and hashed by
//              the SAME batched Keccak kernel the verifier uses; a child enters its parent as raw RLP when
//              < 32 bytes, else as its hash (mpt.zig:104,112); the root is always hashed (mpt.zig:42).
// A forest (many tries at once) is what S needs: all storage tries of a state are built together.
// S follows evmone/test/state/mpt_hash.cpp:15-36 for the trie contents (phant has no StateDB.root()).
// U is the dirty-frontier recompute over a resident complete 16-ary trie (BASELINE.json config C4).
#include "../../include/phant_gpu.h"
#include "common.cuh"
#include "ctx.cuh"
#include "keccak_f1600.cuh"

#include <chrono>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <string.h>
#include <new>
#include <vector>

using namespace phant;

#define CU(expr)                                                                \
    do {                                                                        \
        cudaError_t e_ = (expr);                                                \
        if (e_ != cudaSuccess) return ctx->fail(e_, #expr, __FILE__, __LINE__); \
    } while (0)
#define RC(expr)                  \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != 0) return rc_; \
    } while (0)

namespace {

// development knob PHANT_GPU_TRACE=1: wall time of the phases of a sparse-trie update on stderr (adds a synchronisation per phase)
struct PhaseTrace {
    cudaStream_t s;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTrace(cudaStream_t st) : s(st)
    {
        static int env = -1;
        if (env < 0) { const char* e = getenv("PHANT_GPU_TRACE"); env = e && *e == '1'; }
        on = env == 1;
        if (on) { cudaStreamSynchronize(s); t0 = std::chrono::steady_clock::now(); }
    }
    void mark(const char* what)
    {
        if (!on) return;
        cudaStreamSynchronize(s);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[phant trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};


constexpr uint32_t NONE = 0xffffffffu;
constexpr uint32_t KIND_LEAF = 1u << 30, KIND_NODE = 2u << 30, KIND_MASK = 3u << 30, IDX_MASK = ~KIND_MASK;

__constant__ uint8_t EMPTY_ROOT_D[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45,
                                         0xe6, 0x92, 0xc0, 0xf8, 0x6e, 0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c,
                                         0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

// ---------------------------------------------------------------- RLP helpers (device)
__device__ __forceinline__ uint32_t be_len(uint64_t v)
{
    uint32_t n = 0;
    while (v) { ++n; v >>= 8; }
    return n;
}
__device__ __forceinline__ uint64_t str_size(uint64_t len, uint32_t first_byte)
{
    if (len == 1 && first_byte < 0x80) return 1;
    if (len <= 55) return 1 + len;
    return 1 + be_len(len) + len;
}
__device__ __forceinline__ uint32_t hdr_size(uint64_t payload) { return payload <= 55 ? 1 : 1 + be_len(payload); }
__device__ __forceinline__ uint32_t put_hdr(uint8_t* out, uint64_t len, uint32_t short_base, uint32_t long_base)
{
    if (len <= 55) { out[0] = (uint8_t)(short_base + len); return 1; }
    const uint32_t n = be_len(len);
    out[0] = (uint8_t)(long_base + n);
    for (uint32_t i = 0; i < n; ++i) out[1 + i] = (uint8_t)(len >> (8 * (n - 1 - i)));
    return 1 + n;
}

// ---------------------------------------------------------------- key access
struct Keys {
    const uint8_t* bytes;
    const uint32_t* off; // n+1
};
__device__ __forceinline__ uint32_t nlen(const Keys& k, uint32_t i) { return 2 * (k.off[i + 1] - k.off[i]); }
__device__ __forceinline__ uint32_t nib(const Keys& k, uint32_t i, uint32_t pos)
{
    const uint8_t b = k.bytes[k.off[i] + (pos >> 1)];
    return (pos & 1) ? (b & 15u) : (b >> 4);
}
// common prefix of keys i and j in nibbles
__device__ uint32_t lcp(const Keys& k, uint32_t i, uint32_t j)
{
    const uint32_t li = k.off[i + 1] - k.off[i], lj = k.off[j + 1] - k.off[j];
    const uint32_t m = li < lj ? li : lj;
    const uint8_t* a = k.bytes + k.off[i];
    const uint8_t* b = k.bytes + k.off[j];
    uint32_t t = 0;
    while (t < m && a[t] == b[t]) ++t;
    if (t == m) return 2 * m;
    return 2 * t + (((a[t] ^ b[t]) & 0xf0) ? 0 : 1);
}
// hex-prefix bytes of nibbles [from, to) of key i (mpt.zig:285-314); returns count
__device__ uint32_t hp_size(uint32_t cnt) { return 1 + cnt / 2; }
__device__ uint32_t put_hp(uint8_t* out, const Keys& k, uint32_t i, uint32_t from, uint32_t to, bool leaf)
{
    const uint32_t cnt = to - from;
    uint32_t o = 0, q = from;
    if (cnt & 1) { out[o++] = (uint8_t)(((leaf ? 3 : 1) << 4) | nib(k, i, q)); ++q; }
    else out[o++] = (uint8_t)((leaf ? 2 : 0) << 4);
    for (; q < to; q += 2) out[o++] = (uint8_t)((nib(k, i, q) << 4) | nib(k, i, q + 1));
    return o;
}
__device__ uint32_t hp_first_byte(const Keys& k, uint32_t i, uint32_t from, uint32_t to, bool leaf)
{
    const uint32_t cnt = to - from;
    return (cnt & 1) ? (((leaf ? 3u : 1u) << 4) | nib(k, i, from)) : ((leaf ? 2u : 0u) << 4);
}

// ---------------------------------------------------------------- node tables
struct Tables {
    // per unit (capacity = n_keys)
    uint32_t *lo, *hi, *ext_from, *depth, *child; // child[16 * id + v]
    uint8_t* has_value;
    uint32_t* ext_list; // per level: ids of units with an extension, at the level's base
    // per key
    uint32_t* leaf_start; // nibble where the leaf path starts; NONE for branch-value keys
    // references (what a parent copies): 33 bytes + length, for leaves [0, n) and units [n, n + cap)
    uint8_t* ref;
    uint8_t* ref_len;
    uint8_t* top_digest; // per unit: hash of its topmost encoding (extension if any, else branch)
    // per segment
    uint32_t* seg_root; // KIND | idx, or 0 = empty
};

// ---------------------------------------------------------------- validation
__global__ void check_sorted_kernel(Keys k, const uint64_t* __restrict__ val_off, const uint32_t* __restrict__ seg_of_key, uint32_t n,
                                    uint32_t* flags /*[0] unsorted, [1] some key is a prefix of its successor, [2] max key bytes, [3] max value bytes*/)
{
    uint32_t max_k = 0, max_v = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t la = k.off[i + 1] - k.off[i];
        const uint64_t lv = val_off[i + 1] - val_off[i];
        max_k = la > max_k ? la : max_k;
        const uint32_t lv32 = lv > 0xffffffffull ? 0xffffffffu : (uint32_t)lv;
        max_v = lv32 > max_v ? lv32 : max_v;
        if (i + 1 >= n || (seg_of_key && seg_of_key[i] != seg_of_key[i + 1])) continue;
        const uint32_t lb = k.off[i + 2] - k.off[i + 1];
        const uint8_t* a = k.bytes + k.off[i];
        const uint8_t* b = k.bytes + k.off[i + 1];
        const uint32_t m = la < lb ? la : lb;
        uint32_t t = 0;
        while (t < m && a[t] == b[t]) ++t;
        const bool ok = t < m ? a[t] < b[t] : la < lb; // strictly increasing; a strict prefix sorts first
        if (!ok) atomicExch(&flags[0], 1u);
        if (t == la) atomicExch(&flags[1], 1u);        // key i is a prefix of key i+1: that branch will carry a value
    }
    for (int o = 16; o; o >>= 1) {
        const uint32_t ok = __shfl_down_sync(0xffffffffu, max_k, o), ov = __shfl_down_sync(0xffffffffu, max_v, o);
        max_k = ok > max_k ? ok : max_k;
        max_v = ov > max_v ? ov : max_v;
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(&flags[2], max_k); atomicMax(&flags[3], max_v); }
}

// ---------------------------------------------------------------- top-down
__global__ void init_roots_kernel(const uint32_t* __restrict__ seg_off, uint32_t n_seg, Tables t, uint32_t* counters /*[0]=units*/,
                                  uint32_t start /*nibbles already consumed above each segment's root (0 for a whole trie)*/)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        const uint32_t lo = seg_off[s], hi = seg_off[s + 1];
        if (hi == lo) { t.seg_root[s] = 0; continue; }
        if (hi - lo == 1) { t.seg_root[s] = KIND_LEAF | lo; t.leaf_start[lo] = start; continue; }
        const uint32_t id = atomicAdd(&counters[0], 1u);
        t.lo[id] = lo; t.hi[id] = hi; t.ext_from[id] = start;
        t.seg_root[s] = KIND_NODE | id;
    }
}

// One branch unit, 16 cooperating lanes (lane v looks after child nibble v): find the branch depth, the branch value and
// the 16 child ranges; children with >= 2 keys become units of the next level (appended at next_base + counter).
__device__ __forceinline__ void expand_unit(const Keys& k, const Tables& t, uint32_t id, uint32_t v, uint32_t sub, uint32_t next_base,
                                            uint32_t* next_count, uint32_t* ext_count, uint32_t ext_base)
{
    const uint32_t lo = t.lo[id], hi = t.hi[id], from = t.ext_from[id];
    uint32_t p = 0;
    if (v == 0) {
        p = lcp(k, lo, hi - 1);
        const uint32_t l0 = nlen(k, lo);
        if (l0 < p) p = l0;
    }
    p = __shfl_sync(sub, p, 0, 16);
    const bool has_value = nlen(k, lo) == p;
    const uint32_t first = lo + (has_value ? 1 : 0);
    // lower bound of nibble value v at depth p within [first, hi)
    uint32_t a = first, b = hi;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (nib(k, mid, p) < v) a = mid + 1; else b = mid;
    }
    uint32_t ub = __shfl_down_sync(sub, a, 1, 16);
    if (v == 15) ub = hi;
    uint32_t child = 0;
    const uint32_t c = ub - a;
    if (c == 1) {
        child = KIND_LEAF | a;
        t.leaf_start[a] = p + 1;
    } else if (c >= 2) {
        const uint32_t nid = next_base + atomicAdd(next_count, 1u);
        t.lo[nid] = a; t.hi[nid] = ub; t.ext_from[nid] = p + 1;
        child = KIND_NODE | nid;
    }
    t.child[16 * id + v] = child;
    if (v == 0) {
        t.depth[id] = p;
        t.has_value[id] = has_value ? 1 : 0;
        if (has_value) t.leaf_start[lo] = NONE;
        if (p > from) t.ext_list[ext_base + atomicAdd(ext_count, 1u)] = id;
    }
}

// one BFS level per launch (large tries)
__global__ void __launch_bounds__(128)
expand_kernel(Keys k, Tables t, uint32_t beg, uint32_t cnt, uint32_t next_base, uint32_t* counters /*[1]=next units, [2]=ext in this level*/)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sub = 0xffffu << (threadIdx.x & 16); // my half-warp
    for (uint32_t u = gid >> 4; u < cnt; u += (gridDim.x * blockDim.x) >> 4) // the 16 lanes of a half-warp share u
        expand_unit(k, t, beg + u, threadIdx.x & 15, sub, next_base, &counters[1], &counters[2], beg);
}

// the whole BFS in ONE launch of one CTA (small tries are launch-bound): levels[0] = number of levels, then
// (begin, count, extensions) per level; stops at max_levels (the caller sizes it from the key length)
__global__ void __launch_bounds__(1024)
bfs_small_kernel(Keys k, Tables t, uint32_t first_cnt, uint32_t max_levels, uint32_t* __restrict__ levels)
{
    __shared__ uint32_t s_next, s_ext;
    const uint32_t sub = 0xffffu << (threadIdx.x & 16);
    uint32_t beg = 0, cnt = first_cnt, nl = 0;
    while (cnt && nl < max_levels) {
        if (threadIdx.x == 0) { s_next = 0; s_ext = 0; }
        __syncthreads();
        for (uint32_t u = threadIdx.x >> 4; u < cnt; u += blockDim.x >> 4)
            expand_unit(k, t, beg + u, threadIdx.x & 15, sub, beg + cnt, &s_next, &s_ext, beg);
        __syncthreads();
        if (threadIdx.x == 0) { levels[1 + 3 * nl] = beg; levels[2 + 3 * nl] = cnt; levels[3 + 3 * nl] = s_ext; }
        beg += cnt;
        cnt = s_next;
        ++nl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { levels[0] = nl; levels[1 + 3 * max_levels] = cnt; } // cnt != 0: deeper than max_levels (never with max_levels >= key nibbles)
}

// ---------------------------------------------------------------- leaves
struct Vals {
    const uint8_t* bytes;
    const uint64_t* off;
};

// ids (nullable): work item j is key ids[j] (the leaves that still have to be encoded when a leaf-reference cache is in use)
__global__ void leaf_size_kernel(Keys k, Vals vals, Tables t, uint32_t n, uint64_t* __restrict__ size, const uint32_t* __restrict__ ids)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = ids ? ids[j] : j;
        const uint32_t ls = t.leaf_start[i];
        uint64_t sz = 0;
        if (ls != NONE) {
            const uint32_t nl = nlen(k, i);
            const uint32_t hpn = hp_size(nl - ls);
            const uint64_t vl = vals.off[i + 1] - vals.off[i];
            const uint64_t payload = str_size(hpn, hp_first_byte(k, i, ls, nl, true)) + str_size(vl, vl ? vals.bytes[vals.off[i]] : 0);
            sz = hdr_size(payload) + payload;
        }
        size[j] = sz;
    }
}
// one warp per leaf: lane 0 writes the headers and the path, all lanes copy the value
__global__ void __launch_bounds__(256)
leaf_encode_kernel(Keys k, Vals vals, Tables t, uint32_t n, const uint64_t* __restrict__ aoff, uint8_t* __restrict__ arena,
                   uint64_t* __restrict__ size_out /*nullable: slot layout, the encoder reports the sizes itself*/,
                   const uint32_t* __restrict__ ids /*nullable: work item j is key ids[j]*/)
{
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t i = ids ? ids[j] : j;
        const uint32_t ls = t.leaf_start[i];
        if (ls == NONE) { if (size_out && lane == 0) size_out[j] = 0; continue; }
        uint8_t* out = arena + aoff[j];
        const uint32_t nl = nlen(k, i);
        const uint32_t hpn = hp_size(nl - ls);
        const uint64_t vl = vals.off[i + 1] - vals.off[i];
        const uint8_t* v = vals.bytes + vals.off[i];
        const uint32_t hp0 = hp_first_byte(k, i, ls, nl, true);
        const uint64_t s_hp = str_size(hpn, hp0), s_v = str_size(vl, vl ? v[0] : 0);
        const uint32_t h = hdr_size(s_hp + s_v);
        if (lane == 0) {
            if (size_out) size_out[j] = h + s_hp + s_v;
            put_hdr(out, s_hp + s_v, 0xc0, 0xf7);
            uint8_t* q = out + h;
            if (s_hp > hpn) q += put_hdr(q, hpn, 0x80, 0xb7);
            put_hp(q, k, i, ls, nl, true);
            q = out + h + s_hp;
            if (s_v > vl) put_hdr(q, vl, 0x80, 0xb7);
        }
        uint8_t* dst = out + h + s_hp + (s_v - vl);
        for (uint64_t b = lane; b < vl; b += 32) dst[b] = v[b];
    }
}
// Leaf-reference cache (resident tries): row i = [leaf_start + 1 (0 = nothing cached)] + the 32-byte digest of key i's leaf as
// it was last encoded.  A leaf's encoding depends on (key, the nibble its path starts at, value) only, so a row whose depth
// byte matches needs neither encode nor hash: its reference is written here and the key is left out of the to-do list.
__global__ void leaf_cache_probe_kernel(Tables t, uint32_t n, const uint8_t* __restrict__ cache, uint8_t* __restrict__ leaf_digests,
                                        uint32_t* __restrict__ todo_flag)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t ls = t.leaf_start[i];
        const uint8_t* row = cache + 33ull * i;
        const bool hit = ls != NONE && ls < 255 && row[0] == (uint8_t)(ls + 1);
        todo_flag[i] = hit ? 0 : 1;
        if (hit) {
            uint8_t* r = t.ref + 33ull * i;
            r[0] = 0xa0;
            for (uint32_t b = 0; b < 32; ++b) { r[1 + b] = row[1 + b]; leaf_digests[32ull * i + b] = row[1 + b]; }
            t.ref_len[i] = 33;
        }
    }
}
__global__ void leaf_cache_store_kernel(Tables t, uint32_t n, const uint8_t* __restrict__ leaf_digests, uint8_t* __restrict__ cache_out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t ls = t.leaf_start[i];
        uint8_t* row = cache_out + 33ull * i;
        const bool hashed = ls != NONE && ls < 255 && t.ref_len[i] == 33; // embedded leaves (< 32 bytes) are never cached
        row[0] = hashed ? (uint8_t)(ls + 1) : 0;
        if (hashed)
            for (uint32_t b = 0; b < 32; ++b) row[1 + b] = leaf_digests[32ull * i + b];
    }
}
__global__ void compact_iota_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint32_t n, uint32_t* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) out[pos[i]] = i;
}
// reference of item j (leaf or encoded unit): raw RLP when < 32 bytes, else 0xa0 || digest
__global__ void finalize_ref_kernel(uint32_t cnt, const uint32_t* __restrict__ ids /*nullable: identity*/, uint32_t id_base,
                                    const uint64_t* __restrict__ aoff, const uint64_t* __restrict__ alen /*nullable: CSR*/,
                                    const uint8_t* __restrict__ arena,
                                    const uint8_t* __restrict__ digests, uint32_t ref_base, Tables t, uint8_t* __restrict__ top_digest)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const uint32_t id = ids ? ids[j] : id_base + j;
        const uint64_t len = alen ? alen[j] : aoff[j + 1] - aoff[j];
        uint8_t* r = t.ref + 33ull * (ref_base + id);
        if (len == 0) { t.ref_len[ref_base + id] = 0; continue; } // not a leaf (branch-value key)
        if (len < 32) {
            for (uint32_t b = 0; b < len; ++b) r[b] = arena[aoff[j] + b];
            t.ref_len[ref_base + id] = (uint8_t)len;
        } else {
            r[0] = 0xa0;
            for (uint32_t b = 0; b < 32; ++b) r[1 + b] = digests[32ull * j + b];
            t.ref_len[ref_base + id] = 33;
        }
        if (top_digest)
            for (uint32_t b = 0; b < 32; ++b) top_digest[32ull * id + b] = digests[32ull * j + b];
    }
}

// ---------------------------------------------------------------- branches and extensions
__device__ __forceinline__ uint32_t child_ref_index(uint32_t child, uint32_t n_keys)
{
    return (child & KIND_MASK) == KIND_LEAF ? (child & IDX_MASK) : n_keys + (child & IDX_MASK);
}
__global__ void branch_size_kernel(Keys k, Vals vals, Tables t, uint32_t n_keys, uint32_t beg, uint32_t cnt, uint64_t* __restrict__ size)
{
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < cnt; u += gridDim.x * blockDim.x) {
        const uint32_t id = beg + u;
        uint64_t payload = 0;
        for (uint32_t v = 0; v < 16; ++v) {
            const uint32_t c = t.child[16 * id + v];
            payload += c ? t.ref_len[child_ref_index(c, n_keys)] : 1;
        }
        if (t.has_value[id]) {
            const uint32_t i = t.lo[id];
            const uint64_t vl = vals.off[i + 1] - vals.off[i];
            payload += str_size(vl, vl ? vals.bytes[vals.off[i]] : 0);
        } else payload += 1;
        size[u] = hdr_size(payload) + payload;
    }
}
// one warp per unit: lane v < 16 places child v at the prefix sum of the reference sizes, all lanes copy the value
__global__ void __launch_bounds__(256)
branch_encode_kernel(Keys k, Vals vals, Tables t, uint32_t n_keys, uint32_t beg, uint32_t cnt, const uint64_t* __restrict__ aoff,
                     uint64_t* __restrict__ alen /*nullable: CSR*/, bool self_size /*slot layout: compute and store alen here*/,
                     uint8_t* __restrict__ arena)
{
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < cnt; u += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t id = beg + u;
        uint8_t* out = arena + aoff[u];
        uint64_t total = self_size ? 0 : (alen ? alen[u] : aoff[u + 1] - aoff[u]);
        uint32_t c = 0, sz = 0, ri = 0;
        if (lane < 16) {
            c = t.child[16 * id + lane];
            if (c) { ri = child_ref_index(c, n_keys); sz = t.ref_len[ri]; } else sz = 1;
        }
        uint32_t pre = sz; // inclusive scan over lanes 0..15
        for (int o = 1; o < 16; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= (uint32_t)o) pre += y;
        }
        const uint32_t refs_total = __shfl_sync(0xffffffffu, pre, 15);
        const uint64_t payload_wo_value = refs_total;
        if (self_size) {
            uint64_t s_val = 1;
            if (t.has_value[id]) {
                const uint32_t i = t.lo[id];
                const uint64_t vl = vals.off[i + 1] - vals.off[i];
                s_val = str_size(vl, vl ? vals.bytes[vals.off[i]] : 0);
            }
            total = hdr_size(refs_total + s_val) + refs_total + s_val;
            if (lane == 0) alen[u] = total;
        }
        // total = header + payload: the header size (1..5) is the one consistent with the payload it leaves
        uint32_t hdr = 1;
        for (uint32_t cand = 1; cand <= 5; ++cand) {
            const uint64_t pay = total - cand;
            if (hdr_size(pay) == cand) { hdr = cand; break; }
        }
        if (lane == 0) put_hdr(out, total - hdr, 0xc0, 0xf7);
        if (lane < 16) {
            uint8_t* q = out + hdr + (pre - sz);
            if (c) {
                const uint8_t* r = t.ref + 33ull * ri;
                for (uint32_t b = 0; b < sz; ++b) q[b] = r[b];
            } else q[0] = 0x80;
        }
        uint8_t* q = out + hdr + payload_wo_value;
        if (t.has_value[id]) {
            const uint32_t i = t.lo[id];
            const uint64_t vl = vals.off[i + 1] - vals.off[i];
            const uint8_t* v = vals.bytes + vals.off[i];
            const uint64_t s_v = str_size(vl, vl ? v[0] : 0);
            if (lane == 0 && s_v > vl) put_hdr(q, vl, 0x80, 0xb7);
            uint8_t* dst = q + (s_v - vl);
            for (uint64_t b = lane; b < vl; b += 32) dst[b] = v[b];
        } else if (lane == 0) q[0] = 0x80;
    }
}
__global__ void ext_size_kernel(Keys k, Tables t, uint32_t n_keys, const uint32_t* __restrict__ ids, uint32_t cnt, uint64_t* __restrict__ size)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const uint32_t id = ids[j];
        const uint32_t from = t.ext_from[id], to = t.depth[id], i = t.lo[id];
        const uint32_t hpn = hp_size(to - from);
        const uint64_t payload = str_size(hpn, hp_first_byte(k, i, from, to, false)) + t.ref_len[n_keys + id];
        size[j] = hdr_size(payload) + payload;
    }
}
__global__ void ext_encode_kernel(Keys k, Tables t, uint32_t n_keys, const uint32_t* __restrict__ ids, uint32_t cnt,
                                  const uint64_t* __restrict__ aoff, uint8_t* __restrict__ arena, uint64_t* __restrict__ size_out /*nullable*/)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const uint32_t id = ids[j];
        const uint32_t from = t.ext_from[id], to = t.depth[id], i = t.lo[id];
        const uint32_t hpn = hp_size(to - from);
        const uint64_t s_hp = str_size(hpn, hp_first_byte(k, i, from, to, false));
        const uint32_t rl = t.ref_len[n_keys + id];
        uint8_t* out = arena + aoff[j];
        if (size_out) size_out[j] = hdr_size(s_hp + rl) + s_hp + rl;
        uint8_t* q = out + put_hdr(out, s_hp + rl, 0xc0, 0xf7);
        if (s_hp > hpn) q += put_hdr(q, hpn, 0x80, 0xb7);
        q += put_hp(q, k, i, from, to, false);
        const uint8_t* r = t.ref + 33ull * (n_keys + id);
        for (uint32_t b = 0; b < rl; ++b) q[b] = r[b];
    }
}
__global__ void gather_roots_kernel(Tables t, uint32_t n_seg, const uint8_t* __restrict__ leaf_digests, uint8_t* __restrict__ roots)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        const uint32_t r = t.seg_root[s];
        const uint8_t* src = EMPTY_ROOT_D;
        if ((r & KIND_MASK) == KIND_LEAF) src = leaf_digests + 32ull * (r & IDX_MASK);
        else if ((r & KIND_MASK) == KIND_NODE) src = t.top_digest + 32ull * (r & IDX_MASK);
        for (uint32_t b = 0; b < 32; ++b) roots[32ull * s + b] = src[b];
    }
}

__global__ void fixed_offsets_kernel(uint64_t* off, uint64_t n, uint64_t stride)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) off[i] = stride * i;
}

unsigned grid1d(int device, uint64_t work_items, unsigned block, unsigned per_item = 1)
{
    uint64_t blocks = (work_items * per_item + block - 1) / block;
    const uint64_t cap = (uint64_t)keccak_num_sms(device) * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

// exclusive scan of cnt u64 sizes into cnt+1 offsets (in[cnt] must be readable; it is forced to 0 first)
int scan_sizes(phant_gpu_ctx* ctx, uint64_t* sizes, uint64_t* offs, uint64_t cnt)
{
    CU(cudaMemsetAsync(sizes + cnt, 0, 8, ctx->stream));
    size_t temp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, temp, (const uint64_t*)sizes, offs, (int64_t)(cnt + 1), ctx->stream));
    RC(ctx->d_cub.reserve(ctx, temp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.ptr, temp, (const uint64_t*)sizes, offs, (int64_t)(cnt + 1), ctx->stream));
    return 0;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// forest builder: keys/vals on the device, keys sorted inside each segment; roots = n_seg * 32 bytes (device)
// ------------------------------------------------------------------------------------------------
int phant_gpu_ctx::build_forest(const uint8_t* d_keys, const uint32_t* d_key_off, const uint8_t* d_vals, const uint64_t* d_val_off,
                                uint32_t n, const uint32_t* d_seg_off, uint32_t n_seg, const uint32_t* d_seg_of_key, uint8_t* d_roots,
                                int slots_hint, uint32_t start_depth, const uint8_t* d_leaf_cache, uint8_t* d_leaf_cache_out)
{
    phant_gpu_ctx* ctx = this;
    cudaStream_t s = stream;
    if (n_seg == 0) return PHANT_GPU_OK;
    PhaseTrace trf(s);
    const Keys k{d_keys, d_key_off};
    const Vals vals{d_vals, d_val_off};
    const uint32_t cap = n ? n : 1;

    // tables
    RC(d_b0.reserve(ctx, 4ull * cap * 4 + 16ull * 4 * cap + cap)); // lo,hi,ext_from,depth + child + has_value
    uint32_t* base = (uint32_t*)d_b0.ptr;
    Tables t;
    t.lo = base; t.hi = base + cap; t.ext_from = base + 2ull * cap; t.depth = base + 3ull * cap;
    t.child = base + 4ull * cap;
    t.has_value = (uint8_t*)(base + 20ull * cap);
    RC(d_b1.reserve(ctx, 4ull * cap * 2 + 4ull * n_seg)); // ext_list, leaf_start, seg_root
    t.ext_list = (uint32_t*)d_b1.ptr;
    t.leaf_start = t.ext_list + cap;
    t.seg_root = t.leaf_start + cap;
    RC(d_b2.reserve(ctx, 34ull * 2 * cap + 32ull * cap)); // ref, ref_len, top_digest
    t.ref = (uint8_t*)d_b2.ptr;
    t.ref_len = t.ref + 33ull * 2 * cap;
    t.top_digest = t.ref_len + 2ull * cap;
    RC(d_b3.reserve(ctx, 64)); // counters
    uint32_t* counters = (uint32_t*)d_b3.ptr;
    CU(cudaMemsetAsync(counters, 0, 64, s));
    CU(cudaMemsetAsync(t.leaf_start, 0xff, 4ull * cap, s));

    // keys must be strictly sorted inside each segment (mpt.zig:39 asserts)
    if (n) {
        check_sorted_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(k, d_val_off, d_seg_of_key, n, counters + 4);
        stats.launches++;
    }
    init_roots_kernel<<<grid1d(device, n_seg, 256), 256, 0, s>>>(d_seg_off, n_seg, t, counters, start_depth);
    stats.launches++;
    uint32_t h[8];
    CU(cudaMemcpyAsync(h, counters, 32, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (h[4]) return PHANT_GPU_E_INVALID;
    // layout choice (see "Two arena layouts" below), from what the check kernel saw: slots need no branch values and small leaves
    uint32_t leaf_stride = 0;
    if (slots_hint > 0) leaf_stride = (uint32_t)slots_hint;
    else if (slots_hint < 0 && n && !h[5] && h[6] <= 64 && h[7] <= 3072) {
        const uint32_t st = (h[7] + h[6] + 16 + 15) & ~15u;
        if ((uint64_t)st * n <= (1ull << 31)) leaf_stride = st;
    }

    // ---- top-down: one launch per BFS level, or the whole BFS in one single-CTA launch for small tries ----
    std::vector<uint32_t> level_beg, level_cnt, level_ext;
    uint32_t beg = 0, cnt = h[0];
    if (cnt && leaf_stride && n <= 8192) { // slot layout implies keys <= 64 bytes = 128 nibbles: at most 129 levels
        constexpr uint32_t MAXL = 160;
        RC(d_tmp_b.reserve(ctx, 4 * (2 + 3 * MAXL)));
        uint32_t* d_levels = (uint32_t*)d_tmp_b.ptr;
        bfs_small_kernel<<<1, 1024, 0, s>>>(k, t, cnt, MAXL, d_levels);
        stats.launches++;
        std::vector<uint32_t> hl(2 + 3 * MAXL);
        CU(cudaMemcpyAsync(hl.data(), d_levels, 4 * hl.size(), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (hl[1 + 3 * MAXL] != 0) return PHANT_GPU_E_INVALID;
        for (uint32_t l = 0; l < hl[0]; ++l) { level_beg.push_back(hl[1 + 3 * l]); level_cnt.push_back(hl[2 + 3 * l]); level_ext.push_back(hl[3 + 3 * l]); }
        cnt = 0;
    }
    while (cnt) {
        CU(cudaMemsetAsync(counters + 1, 0, 8, s));
        expand_kernel<<<grid1d(device, cnt, 128, 16), 128, 0, s>>>(k, t, beg, cnt, beg + cnt, counters);
        stats.launches++;
        CU(cudaMemcpyAsync(h, counters, 32, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        level_beg.push_back(beg); level_cnt.push_back(cnt); level_ext.push_back(h[2]);
        beg += cnt;
        cnt = h[1];
        if ((uint64_t)beg + cnt > cap) return PHANT_GPU_E_CUDA; // cannot happen: a trie over n keys has < n branch units
    }

    // Two arena layouts.  General: exact sizes, exclusive scan, one read-back of the total per pass.  Slots (leaf_stride != 0:
    // the caller vouches that no key is a prefix of another -- so no branch carries a value -- and bounds the leaf size):
    // every node gets a fixed-stride slot, the Keccak kernel takes (offset, length) pairs, and no pass needs a scan or
    // a host round trip.  Small tries are launch-bound, so this halves their latency.
    const bool slots = leaf_stride != 0;
    constexpr uint32_t BRANCH_STRIDE = 544, EXT_STRIDE = 128;
    uint64_t* fixed_leaf = nullptr; uint64_t* fixed_branch = nullptr; uint64_t* fixed_ext = nullptr;
    if (slots) {
        uint32_t max_level = 1;
        for (uint32_t c : level_cnt) if (c > max_level) max_level = c;
        RC(d_scan_a.reserve(ctx, 8ull * (n + 2) + 8ull * (max_level + 2) * 2));
        fixed_leaf = (uint64_t*)d_scan_a.ptr;
        fixed_branch = fixed_leaf + (n + 2);
        fixed_ext = fixed_branch + (max_level + 2);
        fixed_offsets_kernel<<<grid1d(device, n + 1, 256), 256, 0, s>>>(fixed_leaf, n, leaf_stride);
        fixed_offsets_kernel<<<grid1d(device, max_level + 1, 256), 256, 0, s>>>(fixed_branch, max_level, BRANCH_STRIDE);
        fixed_offsets_kernel<<<grid1d(device, max_level + 1, 256), 256, 0, s>>>(fixed_ext, max_level, EXT_STRIDE);
        stats.launches += 3;
    }

    trf.mark("    forest: check + BFS");
    // ---- leaves: sizes -> offsets -> encode -> hash -> references (only the leaves the cache cannot answer) ----
    uint8_t* leaf_digests = nullptr;
    if (n) {
        RC(d_b6.reserve(ctx, 32ull * n));
        leaf_digests = (uint8_t*)d_b6.ptr;
        uint32_t m = n;                 // leaves to encode
        const uint32_t* ids = nullptr;  // their key indices (nullptr = all, in order)
        if (d_leaf_cache) {
            RC(d_tmp_a.reserve(ctx, 4ull * (n + 2) * 3));
            uint32_t* flag = (uint32_t*)d_tmp_a.ptr;
            uint32_t* pos = flag + n + 2;
            uint32_t* list = pos + n + 2;
            leaf_cache_probe_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(t, n, d_leaf_cache, leaf_digests, flag);
            CU(cudaMemsetAsync(flag + n, 0, 4, s));
            size_t temp = 0;
            CU(cub::DeviceScan::ExclusiveSum(nullptr, temp, (const uint32_t*)flag, pos, (int64_t)(n + 1), s));
            RC(d_cub.reserve(ctx, temp));
            CU(cub::DeviceScan::ExclusiveSum(d_cub.ptr, temp, (const uint32_t*)flag, pos, (int64_t)(n + 1), s));
            compact_iota_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(flag, pos, n, list);
            CU(cudaMemcpyAsync(&m, pos + n, 4, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            ids = list;
            stats.launches += 3;
        }
        if (m) {
            RC(d_b4.reserve(ctx, 8ull * (m + 1) * 2 + 32ull * m + 64));
            uint64_t* sizes = (uint64_t*)d_b4.ptr;
            uint64_t* offs = sizes + (m + 1);
            uint8_t* dg = ids ? (uint8_t*)(((uintptr_t)(offs + (m + 1)) + 15) & ~(uintptr_t)15) : leaf_digests; // compact digests when indirect
            if (!slots) leaf_size_kernel<<<grid1d(device, m, 256), 256, 0, s>>>(k, vals, t, m, sizes, ids);
            uint64_t total = (uint64_t)m * leaf_stride;
            if (slots) offs = fixed_leaf;
            else {
                RC(scan_sizes(ctx, sizes, offs, m));
                CU(cudaMemcpyAsync(&total, offs + m, 8, cudaMemcpyDeviceToHost, s));
                CU(cudaStreamSynchronize(s));
            }
            RC(d_b5.reserve(ctx, total + 64));
            leaf_encode_kernel<<<grid1d(device, m, 256, 32), 256, 0, s>>>(k, vals, t, m, offs, (uint8_t*)d_b5.ptr, slots ? sizes : nullptr, ids);
            stats.launches += slots ? 1 : 2;
            if (slots) RC(hash_slots((const uint8_t*)d_b5.ptr, offs, sizes, m, dg));
            else RC(hash_csr((const uint8_t*)d_b5.ptr, offs, m, total, dg));
            finalize_ref_kernel<<<grid1d(device, m, 256), 256, 0, s>>>(m, ids, 0, offs, slots ? sizes : nullptr, (const uint8_t*)d_b5.ptr, dg, 0, t,
                                                                     ids ? leaf_digests : nullptr);
            stats.launches++;
        }
        if (d_leaf_cache_out) {
            leaf_cache_store_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(t, n, leaf_digests, d_leaf_cache_out);
            stats.launches++;
        }
    }

    trf.mark("    forest: leaves");
    // ---- units, deepest level first: branch, then the extension above it where there is one ----
    for (int L = (int)level_beg.size() - 1; L >= 0; --L) {
        const uint32_t lb = level_beg[L], lc = level_cnt[L], le = level_ext[L];
        RC(d_b7.reserve(ctx, 8ull * (lc + 1) * 2));
        uint64_t* sizes = (uint64_t*)d_b7.ptr;
        uint64_t* offs = sizes + (lc + 1);
        if (!slots) branch_size_kernel<<<grid1d(device, lc, 128), 128, 0, s>>>(k, vals, t, n, lb, lc, sizes);
        uint64_t total = (uint64_t)lc * BRANCH_STRIDE;
        if (slots) offs = fixed_branch;
        else {
            RC(scan_sizes(ctx, sizes, offs, lc));
            CU(cudaMemcpyAsync(&total, offs + lc, 8, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
        }
        RC(d_b8.reserve(ctx, total + 64));
        RC(d_b9.reserve(ctx, 32ull * lc));
        branch_encode_kernel<<<grid1d(device, lc, 256, 32), 256, 0, s>>>(k, vals, t, n, lb, lc, offs, slots ? sizes : nullptr, slots, (uint8_t*)d_b8.ptr);
        stats.launches += slots ? 1 : 2;
        if (slots) RC(hash_slots((const uint8_t*)d_b8.ptr, offs, sizes, lc, (uint8_t*)d_b9.ptr));
        else RC(hash_csr((const uint8_t*)d_b8.ptr, offs, lc, total, (uint8_t*)d_b9.ptr));
        finalize_ref_kernel<<<grid1d(device, lc, 256), 256, 0, s>>>(lc, nullptr, lb, offs, slots ? sizes : nullptr, (const uint8_t*)d_b8.ptr,
                                                                  (const uint8_t*)d_b9.ptr, n, t, t.top_digest);
        stats.launches++;
        if (le) {
            const uint32_t* ids = t.ext_list + lb;
            if (!slots) ext_size_kernel<<<grid1d(device, le, 128), 128, 0, s>>>(k, t, n, ids, le, sizes);
            total = (uint64_t)le * EXT_STRIDE;
            offs = sizes + (lc + 1);
            if (slots) offs = fixed_ext;
            else {
                RC(scan_sizes(ctx, sizes, offs, le));
                CU(cudaMemcpyAsync(&total, offs + le, 8, cudaMemcpyDeviceToHost, s));
                CU(cudaStreamSynchronize(s));
            }
            RC(d_b8.reserve(ctx, total + 64));
            ext_encode_kernel<<<grid1d(device, le, 128), 128, 0, s>>>(k, t, n, ids, le, offs, (uint8_t*)d_b8.ptr, slots ? sizes : nullptr);
            stats.launches += slots ? 1 : 2;
            if (slots) RC(hash_slots((const uint8_t*)d_b8.ptr, offs, sizes, le, (uint8_t*)d_b9.ptr));
            else RC(hash_csr((const uint8_t*)d_b8.ptr, offs, le, total, (uint8_t*)d_b9.ptr));
            finalize_ref_kernel<<<grid1d(device, le, 256), 256, 0, s>>>(le, ids, 0, offs, slots ? sizes : nullptr, (const uint8_t*)d_b8.ptr,
                                                                      (const uint8_t*)d_b9.ptr, n, t, t.top_digest);
            stats.launches++;
        }
    }
    trf.mark("    forest: units bottom-up");
    gather_roots_kernel<<<grid1d(device, n_seg, 128), 128, 0, s>>>(t, n_seg, leaf_digests, d_roots);
    stats.launches++;
    CU(cudaGetLastError());
    return PHANT_GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// M
// ------------------------------------------------------------------------------------------------
extern "C" int phant_gpu_mpt_root(phant_gpu_ctx* ctx, const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals,
                                  const uint64_t* val_off, uint64_t n, uint8_t out_root[32])
{
    if (!ctx || !out_root || (n && (!key_off || !val_off))) return PHANT_GPU_E_INVALID;
    if (n >= (1ull << 30)) return PHANT_GPU_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    static const uint8_t EMPTY[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                      0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};
    if (n == 0) { memcpy(out_root, EMPTY, 32); return PHANT_GPU_OK; }
    cudaStream_t s = ctx->stream;
    const uint8_t* d_keys = keys; const uint32_t* d_koff = key_off; const uint8_t* d_vals = vals; const uint64_t* d_voff = val_off;
    if (!(ctx->flags & PHANT_GPU_FLAG_DEVICE_PTRS)) {
        for (uint64_t i = 0; i < n; ++i)
            if (key_off[i + 1] < key_off[i] || val_off[i + 1] < val_off[i]) return PHANT_GPU_E_INVALID;
        const uint64_t kb = key_off[n], vb = val_off[n];
        if ((kb && !keys) || (vb && !vals)) return PHANT_GPU_E_INVALID;
        RC(ctx->d_msgs.reserve(ctx, kb + vb + 128));
        RC(ctx->d_off.reserve(ctx, 4 * (n + 1) + 8 * (n + 1) + 16));
        uint8_t* dk = (uint8_t*)ctx->d_msgs.ptr;
        uint8_t* dv = dk + ((kb + 63) & ~63ull);
        uint64_t* dvo = (uint64_t*)ctx->d_off.ptr;
        uint32_t* dko = (uint32_t*)(dvo + (n + 1));
        if (kb) CU(cudaMemcpyAsync(dk, keys, kb, cudaMemcpyHostToDevice, s));
        if (vb) CU(cudaMemcpyAsync(dv, vals, vb, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dko, key_off, 4 * (n + 1), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dvo, val_off, 8 * (n + 1), cudaMemcpyHostToDevice, s));
        ctx->stats.h2d_bytes += kb + vb + 12 * (n + 1);
        d_keys = dk; d_koff = dko; d_vals = dv; d_voff = dvo;
    }
    RC(ctx->d_first.reserve(ctx, 64));
    uint32_t seg[2] = {0, (uint32_t)n};
    CU(cudaMemcpyAsync(ctx->d_first.ptr, seg, 8, cudaMemcpyHostToDevice, s));
    RC(ctx->d_roots.reserve(ctx, 32));
    RC(ctx->build_forest(d_keys, d_koff, d_vals, d_voff, (uint32_t)n, (const uint32_t*)ctx->d_first.ptr, 1, nullptr, (uint8_t*)ctx->d_roots.ptr,
                         /*layout decided on the device*/ -1));
    CU(cudaMemcpyAsync(out_root, ctx->d_roots.ptr, 32, cudaMemcpyDeviceToHost, s));
    ctx->stats.d2h_bytes += 32;
    CU(cudaStreamSynchronize(s));
    return PHANT_GPU_OK;
}

extern "C" int phant_gpu_mpt_roots(phant_gpu_ctx* ctx, const uint8_t* keys, const uint32_t* key_off, const uint8_t* vals,
                                   const uint64_t* val_off, const uint32_t* seg_off, uint64_t n_tries, uint8_t* out_roots)
{
    if (!ctx || (n_tries && (!seg_off || !out_roots))) return PHANT_GPU_E_INVALID;
    if (n_tries == 0) return PHANT_GPU_OK;
    if (ctx->flags & PHANT_GPU_FLAG_DEVICE_PTRS) return PHANT_GPU_E_INVALID; // the batched form takes host arrays
    if (n_tries >= (1ull << 30)) return PHANT_GPU_E_INVALID;
    for (uint64_t t = 0; t < n_tries; ++t) if (seg_off[t + 1] < seg_off[t]) return PHANT_GPU_E_INVALID;
    if (seg_off[0] != 0) return PHANT_GPU_E_INVALID;
    const uint64_t n = seg_off[n_tries];
    if (n >= (1ull << 30) || (n && (!key_off || !val_off))) return PHANT_GPU_E_INVALID;
    for (uint64_t i = 0; i < n; ++i)
        if (key_off[i + 1] < key_off[i] || val_off[i + 1] < val_off[i]) return PHANT_GPU_E_INVALID;
    const uint64_t kb = n ? key_off[n] : 0, vb = n ? val_off[n] : 0;
    if ((kb && !keys) || (vb && !vals)) return PHANT_GPU_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // segment id of every key (the sortedness check must not compare across tries)
    std::vector<uint32_t> seg_of_key(n ? n : 1);
    for (uint64_t t = 0; t < n_tries; ++t)
        for (uint32_t i = seg_off[t]; i < seg_off[t + 1]; ++i) seg_of_key[i] = (uint32_t)t;
    RC(ctx->d_msgs.reserve(ctx, kb + vb + 128));
    RC(ctx->d_off.reserve(ctx, 4 * (n + 1) + 8 * (n + 1) + 16));
    RC(ctx->d_first.reserve(ctx, 4 * (n_tries + 1) + 4 * (n + 1) + 16));
    RC(ctx->d_roots.reserve(ctx, 32 * n_tries));
    uint8_t* dk = (uint8_t*)ctx->d_msgs.ptr;
    uint8_t* dv = dk + ((kb + 63) & ~63ull);
    uint64_t* dvo = (uint64_t*)ctx->d_off.ptr;
    uint32_t* dko = (uint32_t*)(dvo + (n + 1));
    uint32_t* dseg = (uint32_t*)ctx->d_first.ptr;
    uint32_t* dsok = dseg + (n_tries + 1);
    static const uint32_t zero_off[2] = {0, 0};
    static const uint64_t zero_off64[2] = {0, 0};
    if (kb) CU(cudaMemcpyAsync(dk, keys, kb, cudaMemcpyHostToDevice, s));
    if (vb) CU(cudaMemcpyAsync(dv, vals, vb, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dko, n ? key_off : zero_off, 4 * (n + 1), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dvo, n ? val_off : zero_off64, 8 * (n + 1), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dseg, seg_off, 4 * (n_tries + 1), cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(dsok, seg_of_key.data(), 4 * n, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s)); // seg_of_key is a local vector: the copy must finish before it goes out of scope
    ctx->stats.h2d_bytes += kb + vb + 16 * (n + 1) + 4 * (n_tries + 1);
    RC(ctx->build_forest(dk, dko, dv, dvo, (uint32_t)n, dseg, (uint32_t)n_tries, dsok, (uint8_t*)ctx->d_roots.ptr, -1));
    CU(cudaMemcpyAsync(out_roots, ctx->d_roots.ptr, 32 * n_tries, cudaMemcpyDeviceToHost, s));
    ctx->stats.d2h_bytes += 32 * n_tries;
    CU(cudaStreamSynchronize(s));
    return PHANT_GPU_OK;
}

// ------------------------------------------------------------------------------------------------
// S: state root
// ------------------------------------------------------------------------------------------------
namespace {

// non-zero slots -> (segment = account, hashed key index); also the rlp(trim(value)) size
__global__ void slot_flag_kernel(const uint8_t* __restrict__ vals32, uint64_t n_slots, uint8_t* __restrict__ keep)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(vals32 + 32 * i);
        const uint4 a = p[0], b = p[1];
        keep[i] = (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) ? 1 : 0;
    }
}
__global__ void slot_account_kernel(const uint64_t* __restrict__ slot_off, uint32_t n_acc, uint32_t* __restrict__ acc_of_slot)
{
    // one warp per account, lanes stride over its slots
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < n_acc; a += (gridDim.x * blockDim.x) >> 5)
        for (uint64_t s = slot_off[a] + lane; s < slot_off[a + 1]; s += 32) acc_of_slot[s] = a;
}
// key word w (0 = most significant 8 bytes, big endian) of the 32-byte hashes, for the LSD radix passes
__global__ void key_word_kernel(const uint8_t* __restrict__ hashes, const uint32_t* __restrict__ perm, uint32_t n, uint32_t w,
                                uint64_t* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint8_t* h = hashes + 32ull * perm[i] + 8 * w;
        uint64_t v = 0;
        for (int b = 0; b < 8; ++b) v = (v << 8) | h[b];
        out[i] = v;
    }
}
__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = src[perm[i]];
}
__global__ void iota_kernel(uint32_t* p, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}
__global__ void storage_value_size_kernel(const uint8_t* __restrict__ vals32, const uint32_t* __restrict__ slot_of_sorted, uint32_t n,
                                          uint64_t* __restrict__ size)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint8_t* v = vals32 + 32ull * slot_of_sorted[i];
        uint32_t z = 0;
        while (z < 32 && v[z] == 0) ++z;
        const uint32_t len = 32 - z;
        size[i] = (len == 1 && v[z] < 0x80) ? 1 : 1 + len;
    }
}
__global__ void storage_fill_kernel(const uint8_t* __restrict__ slot_hash, const uint8_t* __restrict__ vals32,
                                    const uint32_t* __restrict__ slot_of_sorted, uint32_t n, const uint64_t* __restrict__ val_off,
                                    uint8_t* __restrict__ keys_out, uint32_t* __restrict__ key_off, uint8_t* __restrict__ vals_out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        key_off[i] = 32u * i;
        if (i == n) break;
        const uint32_t sl = slot_of_sorted[i];
        for (int b = 0; b < 32; ++b) keys_out[32ull * i + b] = slot_hash[32ull * sl + b];
        const uint8_t* v = vals32 + 32ull * sl;
        uint32_t z = 0;
        while (z < 32 && v[z] == 0) ++z;
        const uint32_t len = 32 - z;
        uint8_t* o = vals_out + val_off[i];
        if (!(len == 1 && v[z] < 0x80)) *o++ = (uint8_t)(0x80 + len);
        for (uint32_t b = 0; b < len; ++b) o[b] = v[z + b];
    }
}
// segment offsets of the sorted, compacted slots: seg_off[a] = first sorted slot of account a (counts then scan)
__global__ void count_per_account_kernel(const uint32_t* __restrict__ acc_sorted, uint32_t n, uint32_t* __restrict__ counts)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&counts[acc_sorted[i]], 1u);
}
// account leaf value rlp([nonce, balance, storage_root, code_hash]) in sorted account order
__global__ void account_size_kernel(const uint64_t* __restrict__ nonce, const uint8_t* __restrict__ balance32,
                                    const uint32_t* __restrict__ acc_of_sorted, uint32_t n, uint64_t* __restrict__ size)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t a = acc_of_sorted[i];
        const uint64_t nn = nonce[a];
        const uint32_t nl = be_len(nn);
        uint64_t payload = (nl == 0) ? 1 : ((nl == 1 && nn < 0x80) ? 1 : 1 + nl);
        const uint8_t* b = balance32 + 32ull * a;
        uint32_t z = 0;
        while (z < 32 && b[z] == 0) ++z;
        const uint32_t bl = 32 - z;
        payload += (bl == 0) ? 1 : ((bl == 1 && b[z] < 0x80) ? 1 : 1 + bl);
        payload += 33 + 33;
        size[i] = hdr_size(payload) + payload;
    }
}
__global__ void account_fill_kernel(const uint64_t* __restrict__ nonce, const uint8_t* __restrict__ balance32,
                                    const uint8_t* __restrict__ storage_roots, const uint8_t* __restrict__ code_hashes,
                                    const uint8_t* __restrict__ addr_hashes, const uint32_t* __restrict__ acc_of_sorted, uint32_t n,
                                    const uint64_t* __restrict__ val_off, uint8_t* __restrict__ keys_out, uint32_t* __restrict__ key_off,
                                    uint8_t* __restrict__ vals_out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        key_off[i] = 32u * i;
        if (i == n) break;
        const uint32_t a = acc_of_sorted[i];
        for (int b = 0; b < 32; ++b) keys_out[32ull * i + b] = addr_hashes[32ull * a + b];
        uint8_t* out = vals_out + val_off[i];
        const uint64_t total = val_off[i + 1] - val_off[i];
        uint32_t hdr = 1;
        for (uint32_t cand = 1; cand <= 3; ++cand)
            if (hdr_size(total - cand) == cand) { hdr = cand; break; }
        uint8_t* q = out + put_hdr(out, total - hdr, 0xc0, 0xf7);
        const uint64_t nn = nonce[a];
        const uint32_t nl = be_len(nn);
        if (nl == 0) *q++ = 0x80;
        else if (nl == 1 && nn < 0x80) *q++ = (uint8_t)nn;
        else { *q++ = (uint8_t)(0x80 + nl); for (uint32_t b = 0; b < nl; ++b) *q++ = (uint8_t)(nn >> (8 * (nl - 1 - b))); }
        const uint8_t* bal = balance32 + 32ull * a;
        uint32_t z = 0;
        while (z < 32 && bal[z] == 0) ++z;
        const uint32_t bl = 32 - z;
        if (bl == 0) *q++ = 0x80;
        else if (bl == 1 && bal[z] < 0x80) *q++ = bal[z];
        else { *q++ = (uint8_t)(0x80 + bl); for (uint32_t b = 0; b < bl; ++b) *q++ = bal[z + b]; }
        *q++ = 0xa0;
        for (int b = 0; b < 32; ++b) *q++ = storage_roots[32ull * a + b];
        *q++ = 0xa0;
        for (int b = 0; b < 32; ++b) *q++ = code_hashes[32ull * a + b];
    }
}
__global__ void gather_rows32_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t cnt, uint8_t* __restrict__ dst)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(src + 32ull * idx[i]);
        uint4* q = reinterpret_cast<uint4*>(dst + 32ull * i);
        q[0] = p[0];
        q[1] = p[1];
    }
}

} // namespace

// stable LSD radix sort of n items by (segment, 32-byte big-endian hash): perm_out[i] = index of the i-th smallest
int phant_gpu_ctx::sort_by_segment_and_hash(const uint8_t* d_hashes, const uint32_t* d_seg /*nullable*/, uint32_t n, uint32_t* d_perm_out,
                                            DevBuf& scratch)
{
    phant_gpu_ctx* ctx = this;
    cudaStream_t s = stream;
    if (n == 0) return PHANT_GPU_OK;
    RC(scratch.reserve(ctx, 8ull * n * 2 + 4ull * n * 3 + 64));
    uint64_t* kin = (uint64_t*)scratch.ptr;
    uint64_t* kout = kin + n;
    uint32_t* pa = (uint32_t*)(kout + n);
    uint32_t* pb = pa + n;
    uint32_t* sk = pb + n;
    iota_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(pa, n);
    size_t temp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint64_t*)kin, kout, (const uint32_t*)pa, pb, (int64_t)n, 0, 64, s));
    RC(d_cub.reserve(ctx, temp));
    uint32_t* cur = pa;
    uint32_t* nxt = pb;
    for (int w = 3; w >= 0; --w) { // least significant word first
        key_word_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(d_hashes, cur, n, (uint32_t)w, kin);
        CU(cub::DeviceRadixSort::SortPairs(d_cub.ptr, temp, (const uint64_t*)kin, kout, (const uint32_t*)cur, nxt, (int64_t)n, 0, 64, s));
        uint32_t* tsw = cur; cur = nxt; nxt = tsw;
        stats.launches += 2;
    }
    if (d_seg) {
        gather_u32_kernel<<<grid1d(device, n, 256), 256, 0, s>>>(d_seg, cur, n, sk);
        size_t temp2 = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, temp2, (const uint32_t*)sk, (uint32_t*)kout, (const uint32_t*)cur, nxt, (int64_t)n, 0, 32, s));
        RC(d_cub.reserve(ctx, temp2));
        // d_cub may have moved: temp storage is only used inside each call, so this is safe
        CU(cub::DeviceRadixSort::SortPairs(d_cub.ptr, temp2, (const uint32_t*)sk, (uint32_t*)kout, (const uint32_t*)cur, nxt, (int64_t)n, 0, 32, s));
        uint32_t* tsw = cur; cur = nxt; nxt = tsw;
        stats.launches += 2;
    }
    CU(cudaMemcpyAsync(d_perm_out, cur, 4ull * n, cudaMemcpyDeviceToDevice, s));
    return PHANT_GPU_OK;
}

namespace {
// seg_off[v] = first sorted key whose top nibble is >= v (v = 0..16): the 16 subtrees under the root branch
__global__ void top_nibble_segments_kernel(const uint8_t* __restrict__ sorted_keys32, uint32_t n, uint32_t* __restrict__ seg_off)
{
    const uint32_t v = threadIdx.x;
    if (v > 16) return;
    uint32_t a = 0, b = n;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if ((uint32_t)(sorted_keys32[32ull * mid] >> 4) < v) a = mid + 1; else b = mid;
    }
    seg_off[v] = a;
}
} // namespace

// S and its sharded form.  subtree_mask == nullptr: out = the state root (32 bytes).  Otherwise: out = 16 x 32 bytes, the
// hashes of the subtrees under the root branch's 16 slots (accounts grouped by the top nibble of keccak(addr), tries built
// from nibble 1 on), *subtree_mask bit v = slot v is populated.  An account leaf is >= 70 bytes, so a populated slot's
// reference is always the 32-byte hash.
static int state_root_impl(phant_gpu_ctx* ctx, const phant_gpu_accounts* a, uint8_t* out_root, uint32_t* subtree_mask)
{
    if (!ctx || !a || !out_root) return PHANT_GPU_E_INVALID;
    if (ctx->flags & PHANT_GPU_FLAG_DEVICE_PTRS) return PHANT_GPU_E_INVALID; // S takes host tables (it is the StateDB flattening)
    const uint64_t n = a->n_accounts;
    static const uint8_t EMPTY[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                      0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};
    if (n == 0) {
        if (subtree_mask) { *subtree_mask = 0; memset(out_root, 0, 16 * 32); }
        else memcpy(out_root, EMPTY, 32);
        return PHANT_GPU_OK;
    }
    if (n >= (1ull << 30) || !a->addr20 || !a->nonce || !a->balance32 || !a->code_off || !a->slot_off) return PHANT_GPU_E_INVALID;
    for (uint64_t i = 0; i < n; ++i)
        if (a->code_off[i + 1] < a->code_off[i] || a->slot_off[i + 1] < a->slot_off[i]) return PHANT_GPU_E_INVALID;
    const uint64_t code_bytes = a->code_off[n] - a->code_off[0], n_slots = a->slot_off[n] - a->slot_off[0];
    if (a->code_off[0] != 0 || a->slot_off[0] != 0) return PHANT_GPU_E_INVALID;
    if ((code_bytes && !a->code) || (n_slots && (!a->slot_keys32 || !a->slot_vals32)) || n_slots >= (1ull << 30)) return PHANT_GPU_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int dev = ctx->device;

    // ---- stage the tables (one arena in st_in) ----
    auto up = [](uint64_t x) { return (x + 255) & ~255ull; };
    const uint64_t o_addr = 0, o_nonce = up(o_addr + 20 * n), o_bal = up(o_nonce + 8 * n), o_code = up(o_bal + 32 * n),
                   o_coff = up(o_code + code_bytes + 64), o_skey = up(o_coff + 8 * (n + 1)), o_sval = up(o_skey + 32 * n_slots + 64),
                   o_soff = up(o_sval + 32 * n_slots + 64), o_end = up(o_soff + 8 * (n + 1));
    RC(ctx->st_in.reserve(ctx, o_end));
    uint8_t* in = (uint8_t*)ctx->st_in.ptr;
    CU(cudaMemcpyAsync(in + o_addr, a->addr20, 20 * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(in + o_nonce, a->nonce, 8 * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(in + o_bal, a->balance32, 32 * n, cudaMemcpyHostToDevice, s));
    if (code_bytes) CU(cudaMemcpyAsync(in + o_code, a->code, code_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(in + o_coff, a->code_off, 8 * (n + 1), cudaMemcpyHostToDevice, s));
    if (n_slots) {
        CU(cudaMemcpyAsync(in + o_skey, a->slot_keys32, 32 * n_slots, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(in + o_sval, a->slot_vals32, 32 * n_slots, cudaMemcpyHostToDevice, s));
    }
    CU(cudaMemcpyAsync(in + o_soff, a->slot_off, 8 * (n + 1), cudaMemcpyHostToDevice, s));
    ctx->stats.h2d_bytes += 60 * n + code_bytes + 64 * n_slots + 16 * (n + 1);

    // ---- hashes: keccak(addr), keccak(code), keccak(slot key) (batched Keccak kernel, three launches) ----
    const uint64_t h_addr = 0, h_code = up(32 * n), h_slot = up(h_code + 32 * n), h_sroot = up(h_slot + 32 * n_slots + 32),
                   h_offs = up(h_sroot + 32 * n), h_end = up(h_offs + 8 * ((n > n_slots ? n : n_slots) + 1));
    RC(ctx->st_hash.reserve(ctx, h_end));
    uint8_t* hb = (uint8_t*)ctx->st_hash.ptr;
    uint64_t* fixed = (uint64_t*)(hb + h_offs);
    fixed_offsets_kernel<<<grid1d(dev, n, 256), 256, 0, s>>>(fixed, n, 20);
    ctx->stats.launches++;
    RC(ctx->hash_csr(in + o_addr, fixed, n, 20 * n, hb + h_addr));
    RC(ctx->hash_csr(in + o_code, (const uint64_t*)(in + o_coff), n, code_bytes, hb + h_code));
    if (n_slots) {
        fixed_offsets_kernel<<<grid1d(dev, n_slots, 256), 256, 0, s>>>(fixed, n_slots, 32);
        ctx->stats.launches++;
        RC(ctx->hash_csr(in + o_skey, fixed, n_slots, 32 * n_slots, hb + h_slot));
    }

    // ---- storage tries: drop zero slots, sort by (account, hashed key), build all tries as one forest ----
    uint8_t* sroots = hb + h_sroot;
    RC(ctx->st_seg.reserve(ctx, 4ull * (n + 2) * 2 + 4ull * (n_slots + 1) * 4 + (n_slots + 1) + 256));
    uint32_t* seg_cnt = (uint32_t*)ctx->st_seg.ptr;
    uint32_t* seg_off = seg_cnt + (n + 2);
    uint32_t* acc_of_slot = seg_off + (n + 2);
    uint32_t* kept = acc_of_slot + (n_slots + 1);   // compacted -> slot
    uint32_t* perm = kept + (n_slots + 1);          // sorted -> compacted
    uint32_t* slot_sorted = perm + (n_slots + 1);   // sorted -> slot
    uint8_t* keep = (uint8_t*)(slot_sorted + (n_slots + 1));
    uint32_t m = 0; // kept slots
    if (n_slots) {
        slot_flag_kernel<<<grid1d(dev, n_slots, 256), 256, 0, s>>>(in + o_sval, n_slots, keep);
        slot_account_kernel<<<grid1d(dev, n, 256, 32), 256, 0, s>>>((const uint64_t*)(in + o_soff), (uint32_t)n, acc_of_slot);
        iota_kernel<<<grid1d(dev, n_slots, 256), 256, 0, s>>>(perm, (uint32_t)n_slots);
        ctx->stats.launches += 3;
        RC(ctx->d_b3.reserve(ctx, 64));
        size_t temp = 0;
        CU(cub::DeviceSelect::Flagged(nullptr, temp, (const uint32_t*)perm, (const uint8_t*)keep, kept, (uint32_t*)ctx->d_b3.ptr, (int64_t)n_slots, s));
        RC(ctx->d_cub.reserve(ctx, temp));
        CU(cub::DeviceSelect::Flagged(ctx->d_cub.ptr, temp, (const uint32_t*)perm, (const uint8_t*)keep, kept, (uint32_t*)ctx->d_b3.ptr, (int64_t)n_slots, s));
        CU(cudaMemcpyAsync(&m, ctx->d_b3.ptr, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    CU(cudaMemsetAsync(seg_cnt, 0, 4ull * (n + 2), s));
    if (m) {
        // gather the kept slots' hashes / accounts into compact arrays, sort, and lay the forest inputs out
        RC(ctx->st_tmp.reserve(ctx, 32ull * m + 4ull * m * 2 + 64 + 32ull * m + 4ull * (m + 1) + 8ull * (m + 1) * 2 + 33ull * m + 256));
        uint8_t* ch = (uint8_t*)ctx->st_tmp.ptr;               // compact hashes
        uint32_t* cacc = (uint32_t*)(ch + 32ull * m);         // compact account ids
        uint32_t* acc_sorted = cacc + m;
        uint8_t* fk = (uint8_t*)(acc_sorted + m + 16);        // forest keys
        uint32_t* fko = (uint32_t*)(fk + 32ull * m);
        uint64_t* fvs = (uint64_t*)(((uintptr_t)(fko + (m + 1)) + 7) & ~(uintptr_t)7);
        uint64_t* fvo = fvs + (m + 1);
        uint8_t* fv = (uint8_t*)(fvo + (m + 1));
        gather_rows32_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(hb + h_slot, kept, m, ch);
        gather_u32_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(acc_of_slot, kept, m, cacc);
        ctx->stats.launches += 2;
        RC(ctx->sort_by_segment_and_hash(ch, cacc, m, perm, ctx->st_sort));
        gather_u32_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(kept, perm, m, slot_sorted);
        gather_u32_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(cacc, perm, m, acc_sorted);
        count_per_account_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(acc_sorted, m, seg_cnt);
        storage_value_size_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(in + o_sval, slot_sorted, m, fvs);
        ctx->stats.launches += 4;
        RC(scan_sizes(ctx, fvs, fvo, m));
        storage_fill_kernel<<<grid1d(dev, m + 1, 256), 256, 0, s>>>(hb + h_slot, in + o_sval, slot_sorted, m, fvo, fk, fko, fv);
        ctx->stats.launches++;
        size_t temp = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, temp, (const uint32_t*)seg_cnt, seg_off, (int64_t)(n + 1), s));
        RC(ctx->d_cub.reserve(ctx, temp));
        CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.ptr, temp, (const uint32_t*)seg_cnt, seg_off, (int64_t)(n + 1), s));
        RC(ctx->build_forest(fk, fko, fv, fvo, m, seg_off, (uint32_t)n, acc_sorted, sroots, /*32-byte keys, values <= 33 B*/ 96));
    } else {
        CU(cudaMemsetAsync(seg_off, 0, 4ull * (n + 2), s));
        RC(ctx->build_forest(nullptr, seg_off, nullptr, nullptr, 0, seg_off, (uint32_t)n, nullptr, sroots)); // every storage trie empty
    }

    // ---- account trie ----
    RC(ctx->st_acc.reserve(ctx, 4ull * n + 32ull * n + 4ull * (n + 1) + 8ull * (n + 1) * 2 + 112ull * n + 256));
    uint32_t* acc_perm = (uint32_t*)ctx->st_acc.ptr;
    uint8_t* ak = (uint8_t*)(acc_perm + n + (n & 1));
    uint32_t* ako = (uint32_t*)(ak + 32ull * n);
    uint64_t* avs = (uint64_t*)(((uintptr_t)(ako + (n + 1)) + 7) & ~(uintptr_t)7);
    uint64_t* avo = avs + (n + 1);
    uint8_t* av = (uint8_t*)(avo + (n + 1));
    RC(ctx->sort_by_segment_and_hash(hb + h_addr, nullptr, (uint32_t)n, acc_perm, ctx->st_sort));
    account_size_kernel<<<grid1d(dev, n, 256), 256, 0, s>>>((const uint64_t*)(in + o_nonce), in + o_bal, acc_perm, (uint32_t)n, avs);
    ctx->stats.launches++;
    RC(scan_sizes(ctx, avs, avo, n));
    account_fill_kernel<<<grid1d(dev, n + 1, 256), 256, 0, s>>>((const uint64_t*)(in + o_nonce), in + o_bal, sroots, hb + h_code, hb + h_addr, acc_perm,
                                                               (uint32_t)n, avo, ak, ako, av);
    ctx->stats.launches++;
    RC(ctx->d_first.reserve(ctx, 128));
    RC(ctx->d_roots.reserve(ctx, 16 * 32));
    if (subtree_mask) {
        uint32_t seg[17];
        top_nibble_segments_kernel<<<1, 32, 0, s>>>(ak, (uint32_t)n, (uint32_t*)ctx->d_first.ptr);
        ctx->stats.launches++;
        CU(cudaMemcpyAsync(seg, ctx->d_first.ptr, sizeof seg, cudaMemcpyDeviceToHost, s));
        RC(ctx->build_forest(ak, ako, av, avo, (uint32_t)n, (const uint32_t*)ctx->d_first.ptr, 16, nullptr, (uint8_t*)ctx->d_roots.ptr,
                             160, /*the root branch consumed nibble 0*/ 1));
        CU(cudaMemcpyAsync(out_root, ctx->d_roots.ptr, 16 * 32, cudaMemcpyDeviceToHost, s));
        ctx->stats.d2h_bytes += 16 * 32 + sizeof seg;
        CU(cudaStreamSynchronize(s));
        uint32_t mask = 0;
        for (int v = 0; v < 16; ++v) {
            if (seg[v + 1] > seg[v]) mask |= 1u << v;
            else memset(out_root + 32 * v, 0, 32);
        }
        *subtree_mask = mask;
        return PHANT_GPU_OK;
    }
    uint32_t seg[2] = {0, (uint32_t)n};
    CU(cudaMemcpyAsync(ctx->d_first.ptr, seg, 8, cudaMemcpyHostToDevice, s));
    RC(ctx->build_forest(ak, ako, av, avo, (uint32_t)n, (const uint32_t*)ctx->d_first.ptr, 1, nullptr, (uint8_t*)ctx->d_roots.ptr,
                         /*32-byte keys, account RLP <= 110 B*/ 160));
    CU(cudaMemcpyAsync(out_root, ctx->d_roots.ptr, 32, cudaMemcpyDeviceToHost, s));
    ctx->stats.d2h_bytes += 32;
    CU(cudaStreamSynchronize(s));
    return PHANT_GPU_OK;
}

extern "C" int phant_gpu_state_root(phant_gpu_ctx* ctx, const phant_gpu_accounts* a, uint8_t out_root[32])
{
    return state_root_impl(ctx, a, out_root, nullptr);
}

extern "C" int phant_gpu_state_subtree_roots(phant_gpu_ctx* ctx, const phant_gpu_accounts* a, uint8_t out_roots[16 * 32], uint32_t* out_mask)
{
    if (!out_mask) return PHANT_GPU_E_INVALID;
    return state_root_impl(ctx, a, out_roots, out_mask);
}

// ------------------------------------------------------------------------------------------------
// U: resident complete trie
// ------------------------------------------------------------------------------------------------
struct SparseTrie;
struct phant_gpu_trie {
    phant_gpu_ctx* ctx;
    uint32_t kind = 0;
    uint32_t depth;
    std::vector<uint8_t*> level; // kind 0: level[l] = 16^l hashes
    DevBuf store, work;
    SparseTrie* sp = nullptr;    // kind 1
};

namespace {

__device__ __forceinline__ uint64_t sm64(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void ctrie_fill_leaves_kernel(uint64_t seed, uint64_t n_leaves, uint8_t* __restrict__ out)
{
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_leaves; j += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = seed ^ (0xC4ull * 0xA24BAED4963EE407ull) ^ (j * 0xD1342543DE82EF95ull);
        (void)sm64(s);
        uint64_t* o = reinterpret_cast<uint64_t*>(out + 32 * j);
        for (int w = 0; w < 4; ++w) o[w] = sm64(s); // little-endian bytes == the oracle's byte order
    }
}
// branch node over 16 resident child hashes: f9 0211 | 16 x (a0 hash) | 80  (mpt.zig:218-247), 532 bytes, written to the arena
// parents: list of parent positions (nullable = identity).  16 threads per parent, 33 bytes each.
__global__ void ctrie_branch_encode_kernel(const uint8_t* __restrict__ child_level, const uint32_t* __restrict__ parents, uint64_t cnt,
                                           uint8_t* __restrict__ arena)
{
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t v = threadIdx.x & 15;
    for (uint64_t u = gid >> 4; u < cnt; u += ((uint64_t)gridDim.x * blockDim.x) >> 4) {
        const uint64_t p = parents ? parents[u] : u;
        uint8_t* out = arena + 532 * u;
        if (v == 0) { out[0] = 0xf9; out[1] = 0x02; out[2] = 0x11; out[531] = 0x80; }
        const uint4* h = reinterpret_cast<const uint4*>(child_level + 32 * (16 * p + v));
        const uint4 a = h[0], b = h[1];
        uint8_t* q = out + 3 + 33 * v;
        q[0] = 0xa0;
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            q[1 + 4 * i] = (uint8_t)w[i]; q[2 + 4 * i] = (uint8_t)(w[i] >> 8);
            q[3 + 4 * i] = (uint8_t)(w[i] >> 16); q[4 + 4 * i] = (uint8_t)(w[i] >> 24);
        }
    }
}
__global__ void ctrie_scatter_kernel(const uint8_t* __restrict__ digests, const uint32_t* __restrict__ pos, uint64_t cnt, uint8_t* __restrict__ level)
{
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < cnt; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(digests + 32 * u);
        uint4* q = reinterpret_cast<uint4*>(level + 32ull * (pos ? pos[u] : u));
        q[0] = p[0];
        q[1] = p[1];
    }
}
// dirty leaves: position = first `depth` nibbles of the key; leaf RLP = list[hp(nibbles[depth..64), leaf), value]
__global__ void ctrie_leaf_pos_kernel(const uint8_t* __restrict__ keys32, uint64_t n, uint32_t depth, uint32_t* __restrict__ pos)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* key = keys32 + 32 * k;
        uint32_t p = 0;
        for (uint32_t i = 0; i < depth; ++i) p = p * 16 + ((i & 1) ? (key[i >> 1] & 15u) : (key[i >> 1] >> 4));
        pos[k] = p;
    }
}
__global__ void ctrie_leaf_size_kernel(const uint8_t* __restrict__ vals, const uint32_t* __restrict__ val_off, uint64_t n, uint32_t depth,
                                       uint64_t* __restrict__ size)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cnt = 64 - depth, hpn = 1 + cnt / 2;
        const uint64_t vl = val_off[k + 1] - val_off[k];
        // the first hex-prefix byte is 0x20 or 0x3n: a lone byte < 0x80 encodes as itself
        const uint64_t payload = (hpn == 1 ? 1 : 1 + hpn) + str_size(vl, vl ? vals[val_off[k]] : 0);
        size[k] = hdr_size(payload) + payload;
    }
}
__global__ void ctrie_leaf_encode_kernel(const uint8_t* __restrict__ keys32, const uint8_t* __restrict__ vals, const uint32_t* __restrict__ val_off,
                                         uint64_t n, uint32_t depth, const uint64_t* __restrict__ aoff, uint8_t* __restrict__ arena)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* key = keys32 + 32 * k;
        const uint32_t cnt = 64 - depth, hpn = 1 + cnt / 2;
        const uint64_t vl = val_off[k + 1] - val_off[k];
        const uint8_t* v = vals + val_off[k];
        const uint64_t s_hp = hpn == 1 ? 1 : 1 + hpn, s_v = str_size(vl, vl ? v[0] : 0);
        uint8_t* out = arena + aoff[k];
        uint8_t* q = out + put_hdr(out, s_hp + s_v, 0xc0, 0xf7);
        if (hpn > 1) *q++ = (uint8_t)(0x80 + hpn);
        uint32_t i = depth;
#define KN(j) (((j) & 1) ? (key[(j) >> 1] & 15u) : (key[(j) >> 1] >> 4))
        if (cnt & 1) { *q++ = (uint8_t)(0x30 | KN(i)); ++i; } else *q++ = 0x20;
        for (; i < 64; i += 2) *q++ = (uint8_t)((KN(i) << 4) | KN(i + 1));
#undef KN
        if (s_v > vl) q += put_hdr(q, vl, 0x80, 0xb7);
        for (uint64_t b = 0; b < vl; ++b) q[b] = v[b];
    }
}
__global__ void shift4_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ out)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) out[k] = in[k] >> 4;
}

// ---- fused frontier kernels: encode in shared memory, hash, store -- one launch per level, no arena, no host sync ----
constexpr int FR_SLOT = 560;  // same geometry as the staged Keccak kernel's slots (16 x 35)
constexpr int FR_WARPS = 4;
constexpr int FR_SMEM = FR_WARPS * 32 * FR_SLOT + 16;

__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

// byte stream -> aligned 32-bit shared-memory words; FILL (bytes pending in acc) is a compile-time constant everywhere
template <int FILL>
__device__ __forceinline__ void push_word(uint32_t& sa, uint32_t& acc, uint32_t h)
{
    if constexpr (FILL == 0) { sts32(sa, h); sa += 4; }
    else { sts32(sa, acc | (h << (8 * FILL))); sa += 4; acc = h >> (32 - 8 * FILL); }
}
template <int S>
__device__ __forceinline__ void push_child(uint32_t& sa, uint32_t& acc, const uint4 a, const uint4 b)
{
    // before child S the stream holds 3 + 33*S bytes: FILL = (3 + S) % 4; the 0xa0 marker goes first
    constexpr int F0 = (3 + S) % 4;
    if constexpr (F0 == 3) { sts32(sa, acc | (0xa0u << 24)); sa += 4; acc = 0; }
    else acc |= 0xa0u << (8 * F0);
    constexpr int F = (F0 + 1) % 4;
    push_word<F>(sa, acc, a.x); push_word<F>(sa, acc, a.y); push_word<F>(sa, acc, a.z); push_word<F>(sa, acc, a.w);
    push_word<F>(sa, acc, b.x); push_word<F>(sa, acc, b.y); push_word<F>(sa, acc, b.z); push_word<F>(sa, acc, b.w);
}

// Re-hash the dirty branch nodes of one level: parent p = parents[i] (i < *count), children = child_level[16p .. 16p+16).
// Each thread writes the 532-byte encoding f9 0211 | 16 x (a0 hash) | 80 (mpt.zig:218-247) into its shared-memory slot as
// aligned words, absorbs it with the product sponge, and stores the digest at level[p].
__global__ void __launch_bounds__(FR_WARPS * 32)
frontier_branch_kernel(const uint8_t* __restrict__ child_level, const uint32_t* __restrict__ parents, const uint32_t* __restrict__ count_ptr,
                       uint32_t bound, uint8_t* __restrict__ level, const uint32_t* __restrict__ refuse)
{
    extern __shared__ __align__(16) uint8_t fr_smem[];
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(fr_smem) + threadIdx.x * FR_SLOT;
    if (*refuse) return; // duplicate leaf positions: the update is refused as a whole
    uint32_t count = *count_ptr;
    if (count > bound) count = bound;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t p = parents[i];
        const uint4* ch = reinterpret_cast<const uint4*>(child_level + 512ull * p);
        uint32_t sa = slot, acc = 0x1102f9u; // f9 02 11 pending (FILL = 3)
#define PC(S) push_child<S>(sa, acc, ch[2 * S], ch[2 * S + 1]);
        PC(0) PC(1) PC(2) PC(3) PC(4) PC(5) PC(6) PC(7) PC(8) PC(9) PC(10) PC(11) PC(12) PC(13) PC(14) PC(15)
#undef PC
        // after 16 children: 3 + 33*16 = 531 bytes, FILL = 3: the empty value 0x80 completes the last word
        sts32(sa, acc | (0x80u << 24));
        uint64_t st[25];
#pragma unroll
        for (int k = 0; k < 25; ++k) st[k] = 0;
        absorb_full_smem<2>(st, slot);
        absorb_full_smem<2>(st, slot + 136);
        absorb_full_smem<2>(st, slot + 272);
        absorb_final_smem<2>(st, slot + 408, 532 - 408, FR_SLOT - 408);
        uint4* o = reinterpret_cast<uint4*>(level + 32ull * p);
        o[0] = make_uint4((uint32_t)st[0], (uint32_t)(st[0] >> 32), (uint32_t)st[1], (uint32_t)(st[1] >> 32));
        o[1] = make_uint4((uint32_t)st[2], (uint32_t)(st[2] >> 32), (uint32_t)st[3], (uint32_t)(st[3] >> 32));
    }
}

// Dirty leaves: rlp([hp(key nibbles [depth, 64), leaf), value]) built in the slot, hashed, stored at the leaf's position.
// Values up to FR_SLOT - 48 bytes (the caller checks); one thread per leaf.
__global__ void __launch_bounds__(FR_WARPS * 32)
frontier_leaf_kernel(const uint8_t* __restrict__ keys32, const uint8_t* __restrict__ vals, const uint32_t* __restrict__ val_off, uint32_t n,
                     uint32_t depth, uint8_t* __restrict__ leaf_level, const uint32_t* __restrict__ pos_in, const uint32_t* __restrict__ refuse)
{
    extern __shared__ __align__(16) uint8_t fr_smem[];
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(fr_smem) + threadIdx.x * FR_SLOT;
    if (*refuse) return; // duplicate leaf positions: the update is refused as a whole
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint8_t* key = keys32 + 32ull * k;
        const uint32_t pos = pos_in[k];
        const uint32_t cnt = 64 - depth, hpn = 1 + cnt / 2;
        const uint32_t vl = val_off[k + 1] - val_off[k];
        const uint8_t* v = vals + val_off[k];
        const uint32_t s_hp = hpn == 1 ? 1 : 1 + hpn, s_v = (uint32_t)str_size(vl, vl ? v[0] : 0);
        const uint32_t payload = s_hp + s_v;
        uint32_t sa = slot;
        if (payload <= 55) { sts8(sa++, 0xc0 + payload); }
        else if (payload < 256) { sts8(sa++, 0xf8); sts8(sa++, payload); }
        else { sts8(sa++, 0xf9); sts8(sa++, payload >> 8); sts8(sa++, payload & 255); }
        if (hpn > 1) sts8(sa++, 0x80 + hpn);
        uint32_t i = depth;
#define KN(j) (((j) & 1) ? (key[(j) >> 1] & 15u) : (key[(j) >> 1] >> 4))
        if (cnt & 1) { sts8(sa++, 0x30 | KN(i)); ++i; } else sts8(sa++, 0x20);
        for (; i < 64; i += 2) sts8(sa++, (KN(i) << 4) | KN(i + 1));
#undef KN
        if (s_v > vl) {
            if (vl <= 55) sts8(sa++, 0x80 + vl);
            else if (vl < 256) { sts8(sa++, 0xb8); sts8(sa++, vl); }
            else { sts8(sa++, 0xb9); sts8(sa++, vl >> 8); sts8(sa++, vl & 255); }
        }
        for (uint32_t b = 0; b < vl; ++b) sts8(sa++, v[b]);
        const uint32_t len = sa - slot;
        uint64_t st[25];
#pragma unroll
        for (int q = 0; q < 25; ++q) st[q] = 0;
        uint32_t at = slot, rem = len;
        while (rem >= 136) { absorb_full_smem<2>(st, at); at += 136; rem -= 136; }
        absorb_final_smem<2>(st, at, rem, slot + FR_SLOT - at);
        uint4* o = reinterpret_cast<uint4*>(leaf_level + 32ull * pos);
        o[0] = make_uint4((uint32_t)st[0], (uint32_t)(st[0] >> 32), (uint32_t)st[1], (uint32_t)(st[1] >> 32));
        o[1] = make_uint4((uint32_t)st[2], (uint32_t)(st[2] >> 32), (uint32_t)st[3], (uint32_t)(st[3] >> 32));
    }
}

// sorted leaf positions: two dirty keys on one leaf position (a repeated key, or keys sharing their first `depth` nibbles)
// would race on leaf_level[pos]; flag it so that every later kernel of the update leaves the trie untouched
__global__ void ctrie_dup_check_kernel(const uint32_t* __restrict__ sorted_pos, uint64_t n, uint32_t* __restrict__ flag)
{
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k + 1 < n; k += (uint64_t)gridDim.x * blockDim.x)
        if (sorted_pos[k] == sorted_pos[k + 1]) *flag = 1;
}

// children (sorted, `*count` valid of `bound`) -> parents = children >> 4, the tail padded with the last valid value so
// that a fixed-size Unique over `bound` items yields exactly the distinct parents without the host knowing `*count`
__global__ void shift4_pad_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ count_ptr, uint32_t bound, uint32_t* __restrict__ out)
{
    uint32_t count = *count_ptr;
    if (count > bound) count = bound;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < bound; k += gridDim.x * blockDim.x)
        out[k] = count ? in[k < count ? k : count - 1] >> 4 : 0;
}

int ctrie_hash_level(phant_gpu_trie* t, uint32_t l, const uint32_t* d_parents, uint64_t cnt)
{
    // re-hash `cnt` branch nodes of level l (positions d_parents) from level l+1
    phant_gpu_ctx* ctx = t->ctx;
    cudaStream_t s = ctx->stream;
    const uint64_t batch = 1ull << 20; // bound the scratch arena (532 B per node)
    for (uint64_t b0 = 0; b0 < cnt; b0 += batch) {
        const uint64_t c = cnt - b0 < batch ? cnt - b0 : batch;
        RC(t->work.reserve(ctx, 532 * c + 64 + 8 * (c + 2) + 32 * c + 256));
        uint8_t* arena = (uint8_t*)t->work.ptr;
        uint64_t* offs = (uint64_t*)(arena + ((532 * c + 64 + 15) & ~15ull));
        uint8_t* dg = (uint8_t*)(offs + ((c + 2) & ~1ull)); // 16-byte aligned: digests are stored as 128-bit words
        const uint32_t* par = d_parents + b0;
        fixed_offsets_kernel<<<grid1d(ctx->device, c, 256), 256, 0, s>>>(offs, c, 532);
        ctrie_branch_encode_kernel<<<grid1d(ctx->device, c, 128, 16), 128, 0, s>>>(t->level[l + 1], par, c, arena);
        ctx->stats.launches += 2;
        const uint32_t saved = ctx->flags;
        ctx->flags |= PHANT_GPU_FLAG_NO_BINNING; // all 532-byte messages: nothing to regroup
        const int rc = ctx->hash_csr(arena, offs, c, 532 * c, dg);
        ctx->flags = saved;
        RC(rc);
        ctrie_scatter_kernel<<<grid1d(ctx->device, c, 256), 256, 0, s>>>(dg, par, c, t->level[l]);
        ctx->stats.launches++;
    }
    CU(cudaGetLastError());
    return PHANT_GPU_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// U kind 1: a SPARSE resident secure trie (32-byte keys, arbitrary values) -- the structure behind StateDB.root() for a real
// state (hook src/blockchain/blockchain.zig:83-85): after a block only the dirty part is re-hashed.
//
// Resident on the device:  the sorted key table (32 B per key) with a value record (offset, length) into an append-only
// value arena, and a DENSE TOP of L nibble levels: level d holds the 16^d node references of depth d.  Level L's entries
// are the roots of the 16^L BUCKETS -- the sparse subtrees of the keys sharing an L-nibble prefix.  Keys are Keccak
// outputs, so with 16..256 keys per bucket (L = floor(log16(n / 16))) every node above the buckets is a real branch with
// >= 2 children and needs no extension / collapse logic; this is CHECKED on the device at every update, and when it does
// not hold (adversarial or tiny key sets) L is lowered and the top rebuilt -- L = 0 is one bucket = a plain rebuild.
//
// An update (upserts; an empty value deletes): sort the dirty keys, merge them into the table (positions by lower bound +
// two scans), re-build ONLY the dirty buckets as one forest with the M builder (build_forest with start_depth = L: leaves,
// extensions, embedded nodes and all of mptize's rules apply inside a bucket), then re-hash the dirty frontier of the L
// dense levels bottom-up (one launch per level, variable child masks).  Pure value updates skip the merge.
// ------------------------------------------------------------------------------------------------
struct SRec { uint64_t off; uint32_t len; uint32_t pad; };

struct SparseTrie {
    uint64_t n = 0;
    DevBuf keys[2], recs[2], cache[2]; // sorted keys / records / leaf-reference cache rows (33 B), ping-pong across merges
    int cur = 0;
    DevBuf arena;
    uint64_t arena_used = 0;
    uint32_t L = 0;
    DevBuf top, present;     // levels 0..L: 32-byte reference + presence byte per node; level d starts at node (16^d - 1) / 15
    DevBuf sa, sb, sc, sd, se, sf, sg, sh, si, sroots, ssort; // scratch
    uint8_t root[32];
    uint64_t updates = 0, rebuilds = 0;
};

namespace {

constexpr uint8_t EMPTY_ROOT_H[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                      0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

inline uint64_t level_base(uint32_t d) { return ((1ull << (4 * d)) - 1) / 15; }

__device__ __forceinline__ int cmp_key32(const uint8_t* a, const uint8_t* b)
{
    const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 16);
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
    const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t x = __byte_perm(aw[i], 0, 0x0123), y = __byte_perm(bw[i], 0, 0x0123); // big-endian order
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}
__device__ __forceinline__ uint32_t key_prefix(const uint8_t* key, uint32_t L) // first L <= 7 nibbles as an integer
{
    const uint32_t w = __byte_perm(*reinterpret_cast<const uint32_t*>(key), 0, 0x0123);
    return L ? w >> (32 - 4 * L) : 0;
}

// dirty keys (sorted): position in the table, whether found; classification into replace / insert / delete
__global__ void st_classify_kernel(const uint8_t* __restrict__ table, uint32_t n, const uint8_t* __restrict__ dk, const uint32_t* __restrict__ dlen,
                                   uint32_t m, uint32_t* __restrict__ lb, uint8_t* __restrict__ kind /*0 no-op, 1 replace, 2 insert, 3 delete*/,
                                   uint32_t* __restrict__ del_flag /*n, nullable when n == 0*/, uint32_t* __restrict__ ins_at /*n+1*/,
                                   uint32_t* __restrict__ ins_flag /*m*/, uint64_t* __restrict__ app_size /*m*/, uint32_t* __restrict__ counters)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const uint8_t* q = dk + 32ull * j;
        uint32_t a = 0, b = n;
        while (a < b) {
            const uint32_t mid = (a + b) >> 1;
            if (cmp_key32(table + 32ull * mid, q) < 0) a = mid + 1; else b = mid;
        }
        const bool found = a < n && cmp_key32(table + 32ull * a, q) == 0;
        const bool del = dlen[j] == 0;
        uint8_t k = 0;
        if (found) k = del ? 3 : 1; else k = del ? 0 : 2;
        lb[j] = a;
        kind[j] = k;
        ins_flag[j] = k == 2;
        app_size[j] = k == 1 || k == 2 ? dlen[j] : 0;
        if (k == 3) { del_flag[a] = 1; atomicAdd(&counters[1], 1u); }
        if (k == 2) { atomicAdd(&ins_at[a], 1u); atomicAdd(&counters[0], 1u); }
        if (j + 1 < m && cmp_key32(q, dk + 32ull * (j + 1)) == 0) counters[2] = 1; // duplicate key in one update
    }
}
// append the new values to the arena; replaced records are rewritten in place
__global__ void st_gather_voff_kernel(const uint32_t* __restrict__ raw_voff, const uint32_t* __restrict__ perm, uint32_t m, uint32_t* __restrict__ dvoff,
                                      uint32_t* __restrict__ dlen)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const uint32_t src = perm[j];
        dvoff[j] = raw_voff[src];
        dlen[j] = raw_voff[src + 1] - raw_voff[src];
    }
}
__global__ void st_append_kernel(const uint8_t* __restrict__ dv, const uint32_t* __restrict__ dvoff, const uint32_t* __restrict__ dlen,
                                 const uint8_t* __restrict__ kind, const uint32_t* __restrict__ lb, const uint64_t* __restrict__ app_off, uint32_t m, uint64_t arena_base,
                                 uint8_t* __restrict__ arena, SRec* __restrict__ recs_cur, SRec* __restrict__ drec, uint8_t* __restrict__ cache_cur)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < m; j += warps) {
        const uint8_t k = kind[j];
        if (k != 1 && k != 2) continue;
        const uint32_t len = dlen[j];
        const uint64_t dst = arena_base + app_off[j];
        for (uint32_t b = lane; b < len; b += 32) arena[dst + b] = dv[dvoff[j] + b];
        if (lane == 0) {
            const SRec r{dst, len, 0};
            drec[j] = r;
            if (k == 1) { recs_cur[lb[j]] = r; cache_cur[33ull * lb[j]] = 0; } // new value: the cached leaf reference is stale
        }
    }
}
__global__ void st_keep_kernel(const uint32_t* __restrict__ del_flag, uint32_t n, uint32_t* __restrict__ keep /*n+1, keep[n] = 0*/)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) keep[i] = i < n && !del_flag[i] ? 1 : 0;
}
// merged table: kept table entries and inserts at their final positions
__global__ void st_merge_table_kernel(const uint8_t* __restrict__ keys, const SRec* __restrict__ recs, uint32_t n, const uint32_t* __restrict__ del_flag,
                                      const uint32_t* __restrict__ K /*excl scan of keep, n*/, const uint32_t* __restrict__ I /*excl scan of ins_at, n+2*/,
                                      uint8_t* __restrict__ keys_out, SRec* __restrict__ recs_out, const uint8_t* __restrict__ cache, uint8_t* __restrict__ cache_out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (del_flag[i]) continue;
        const uint32_t p = K[i] + I[i + 1];
        const uint4* s = reinterpret_cast<const uint4*>(keys + 32ull * i);
        uint4* d = reinterpret_cast<uint4*>(keys_out + 32ull * p);
        d[0] = s[0]; d[1] = s[1];
        recs_out[p] = recs[i];
        for (uint32_t b = 0; b < 33; ++b) cache_out[33ull * p + b] = cache[33ull * i + b];
    }
}
__global__ void st_merge_dirty_kernel(const uint8_t* __restrict__ dk, const SRec* __restrict__ drec, const uint8_t* __restrict__ kind,
                                      const uint32_t* __restrict__ lb, const uint32_t* __restrict__ ins_index, uint32_t m, uint32_t n,
                                      const uint32_t* __restrict__ K /*n+1 entries valid: K[n] = kept total*/, uint8_t* __restrict__ keys_out,
                                      SRec* __restrict__ recs_out, uint8_t* __restrict__ cache_out)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        if (kind[j] != 2) continue;
        const uint32_t p = K[lb[j]] + ins_index[j];
        const uint4* s = reinterpret_cast<const uint4*>(dk + 32ull * j);
        uint4* d = reinterpret_cast<uint4*>(keys_out + 32ull * p);
        d[0] = s[0]; d[1] = s[1];
        recs_out[p] = drec[j];
        cache_out[33ull * p] = 0; // a new key has no cached leaf reference
    }
}
// bucket of each dirty key + "first of its bucket" flag
__global__ void st_bucket_flag_kernel(const uint8_t* __restrict__ dk, uint32_t m, uint32_t L, uint32_t* __restrict__ bucket, uint32_t* __restrict__ first)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const uint32_t b = key_prefix(dk + 32ull * j, L);
        bucket[j] = b;
        first[j] = j == 0 || key_prefix(dk + 32ull * (j - 1), L) != b;
    }
}
__global__ void st_compact_kernel(const uint32_t* __restrict__ val, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint32_t m,
                                  uint32_t* __restrict__ out)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x)
        if (flag[j]) out[pos[j]] = val[j];
}
// table range of each listed bucket (nullptr list = bucket u itself)
__global__ void st_bucket_range_kernel(const uint8_t* __restrict__ table, uint32_t n, uint32_t L, const uint32_t* __restrict__ list, uint32_t nb,
                                       uint32_t* __restrict__ lo_out, uint32_t* __restrict__ cnt_out)
{
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < nb; u += gridDim.x * blockDim.x) {
        const uint32_t b = list ? list[u] : u;
        uint32_t r[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint32_t want = b + e;
            uint32_t lo = 0, hi = n;
            if (L == 0) lo = e ? n : 0;
            else while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (key_prefix(table + 32ull * mid, L) < want) lo = mid + 1; else hi = mid;
            }
            r[e] = lo;
        }
        lo_out[u] = r[0];
        cnt_out[u] = r[1] - r[0];
    }
}
// gather the keys of the listed buckets into a contiguous forest input (warp per bucket)
__global__ void st_gather_keys_kernel(const uint8_t* __restrict__ table, const SRec* __restrict__ recs, const uint32_t* __restrict__ lo,
                                      const uint32_t* __restrict__ seg_off, uint32_t nb, uint8_t* __restrict__ gkeys, uint32_t* __restrict__ gkey_off,
                                      uint32_t* __restrict__ seg_of_key, uint64_t* __restrict__ gval_size, SRec* __restrict__ grec,
                                      const uint8_t* __restrict__ cache, uint8_t* __restrict__ gcache)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < nb; u += warps) {
        const uint32_t base = seg_off[u], cnt = seg_off[u + 1] - base, from = lo[u];
        for (uint32_t t = lane; t < cnt; t += 32) {
            const uint4* s = reinterpret_cast<const uint4*>(table + 32ull * (from + t));
            uint4* d = reinterpret_cast<uint4*>(gkeys + 32ull * (base + t));
            d[0] = s[0]; d[1] = s[1];
            gkey_off[base + t] = 32u * (base + t);
            seg_of_key[base + t] = u;
            const SRec r = recs[from + t];
            gval_size[base + t] = r.len;
            grec[base + t] = r;
            for (uint32_t b = 0; b < 33; ++b) gcache[33ull * (base + t) + b] = cache[33ull * (from + t) + b];
        }
    }
}
// the leaf references of this build back into the table's cache rows
__global__ void st_scatter_cache_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ seg_off, uint32_t nb, const uint8_t* __restrict__ gcache_out,
                                        uint8_t* __restrict__ cache)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < nb; u += warps) {
        const uint32_t base = seg_off[u], cnt = seg_off[u + 1] - base, from = lo[u];
        for (uint32_t t = lane; t < cnt; t += 32)
            for (uint32_t b = 0; b < 33; ++b) cache[33ull * (from + t) + b] = gcache_out[33ull * (base + t) + b];
    }
}
__global__ void st_gather_vals_kernel(const uint8_t* __restrict__ arena, const SRec* __restrict__ grec, const uint64_t* __restrict__ gval_off, uint32_t mk,
                                      uint8_t* __restrict__ gvals)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < mk; t += warps) {
        const SRec r = grec[t];
        const uint64_t dst = gval_off[t];
        for (uint32_t b = lane; b < r.len; b += 32) gvals[dst + b] = arena[r.off + b];
    }
}
__global__ void st_scatter_roots_kernel(const uint8_t* __restrict__ roots, const uint32_t* __restrict__ list, const uint32_t* __restrict__ cnt, uint32_t nb,
                                        uint8_t* __restrict__ level, uint8_t* __restrict__ present)
{
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < nb; u += gridDim.x * blockDim.x) {
        const uint32_t b = list ? list[u] : u;
        const uint4* s = reinterpret_cast<const uint4*>(roots + 32ull * u);
        uint4* d = reinterpret_cast<uint4*>(level + 32ull * b);
        d[0] = s[0]; d[1] = s[1];
        present[b] = cnt[u] ? 1 : 0;
    }
}
// the number of valid children stays on the device (*cnt_ptr <= bound): no host read-back per dense level
__global__ void st_parent_flag_kernel(const uint32_t* __restrict__ child, const uint32_t* __restrict__ cnt_ptr, uint32_t bound, uint32_t* __restrict__ parent,
                                      uint32_t* __restrict__ first)
{
    const uint32_t cnt = *cnt_ptr < bound ? *cnt_ptr : bound;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j <= bound; j += gridDim.x * blockDim.x) {
        if (j < cnt) {
            parent[j] = child[j] >> 4;
            first[j] = j == 0 || (child[j - 1] >> 4) != (child[j] >> 4);
        } else first[j] = 0;
    }
}
// One dense-top node per thread: rlp([ref or "" x 16, ""]) over the children that exist (src/mpt/mpt.zig:218-247), built in the
// thread's shared-memory slot, hashed with the product sponge.  A node with exactly ONE child would have to collapse into
// an extension / its child (mpt.zig:83-106): that breaks the dense-top premise and is reported through *violation.
__global__ void __launch_bounds__(FR_WARPS * 32)
st_top_branch_kernel(const uint8_t* __restrict__ child_level, const uint8_t* __restrict__ child_present, const uint32_t* __restrict__ parents /*nullable*/,
                     uint32_t count, const uint32_t* __restrict__ count_ptr /*nullable: the count lives on the device, `count` is its bound*/,
                     uint8_t* __restrict__ level, uint8_t* __restrict__ present, uint32_t* __restrict__ violation)
{
    extern __shared__ __align__(16) uint8_t fr_smem[];
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(fr_smem) + threadIdx.x * FR_SLOT;
    if (count_ptr && *count_ptr < count) count = *count_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t p = parents ? parents[i] : i;
        uint32_t mask = 0;
        for (uint32_t v = 0; v < 16; ++v) mask |= (child_present[16ull * p + v] ? 1u : 0u) << v;
        const uint32_t c = __popc(mask);
        if (c == 0) { present[p] = 0; continue; }
        if (c == 1) atomicExch(violation, 1u);
        present[p] = 1;
        const uint32_t payload = 33 * c + (16 - c) + 1;
        uint32_t sa = slot;
        if (payload <= 55) sts8(sa++, 0xc0 + payload);
        else if (payload < 256) { sts8(sa++, 0xf8); sts8(sa++, payload); }
        else { sts8(sa++, 0xf9); sts8(sa++, payload >> 8); sts8(sa++, payload & 255); }
        for (uint32_t v = 0; v < 16; ++v) {
            if (!((mask >> v) & 1)) { sts8(sa++, 0x80); continue; }
            sts8(sa++, 0xa0);
            const uint4* h = reinterpret_cast<const uint4*>(child_level + 32ull * (16ull * p + v));
            const uint4 a = h[0], b = h[1];
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) { sts8(sa++, w[k] & 255); sts8(sa++, (w[k] >> 8) & 255); sts8(sa++, (w[k] >> 16) & 255); sts8(sa++, w[k] >> 24); }
        }
        sts8(sa++, 0x80);
        const uint32_t len = sa - slot;
        uint64_t st[25];
#pragma unroll
        for (int q = 0; q < 25; ++q) st[q] = 0;
        uint32_t at = slot, rem = len;
        while (rem >= 136) { absorb_full_smem<2>(st, at); at += 136; rem -= 136; }
        absorb_final_smem<2>(st, at, rem, slot + FR_SLOT - at);
        uint4* o = reinterpret_cast<uint4*>(level + 32ull * p);
        o[0] = make_uint4((uint32_t)st[0], (uint32_t)(st[0] >> 32), (uint32_t)st[1], (uint32_t)(st[1] >> 32));
        o[1] = make_uint4((uint32_t)st[2], (uint32_t)(st[2] >> 32), (uint32_t)st[3], (uint32_t)(st[3] >> 32));
    }
}

int st_scan_u32(phant_gpu_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t cnt)
{
    size_t temp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, temp, in, out, (int64_t)cnt, ctx->stream));
    RC(ctx->d_cub.reserve(ctx, temp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.ptr, temp, in, out, (int64_t)cnt, ctx->stream));
    return 0;
}

uint32_t st_target_L(uint64_t n)
{
    uint32_t L = 0;
    while (L < 6 && (n >> (4 * (L + 1))) >= 16) ++L; // 16 .. 255 keys per bucket
    return L;
}

// (re)build the listed buckets (d_list == nullptr: all 16^L of them) and the dense levels above them; root -> sp->root
int st_rebuild(phant_gpu_trie* t, const uint32_t* d_list, uint32_t nb, bool all)
{
    phant_gpu_ctx* ctx = t->ctx;
    SparseTrie* sp = t->sp;
    cudaStream_t s = ctx->stream;
    const int dev = ctx->device;
    const uint32_t L = sp->L, n = (uint32_t)sp->n;
    const uint8_t* table = (const uint8_t*)sp->keys[sp->cur].ptr;
    const SRec* recs = (const SRec*)sp->recs[sp->cur].ptr;
    uint8_t* top = (uint8_t*)sp->top.ptr;
    uint8_t* pres = (uint8_t*)sp->present.ptr;
    if (n == 0) { memcpy(sp->root, EMPTY_ROOT_H, 32); return PHANT_GPU_OK; }
    PhaseTrace tr(s);
    // ranges and sizes of the buckets
    RC(sp->sa.reserve(ctx, 4ull * (nb + 2) * 3 + 64));
    uint32_t* lo = (uint32_t*)sp->sa.ptr;
    uint32_t* cnt = lo + nb + 2;
    uint32_t* seg_off = cnt + nb + 2;
    st_bucket_range_kernel<<<grid1d(dev, nb, 128), 128, 0, s>>>(table, n, L, d_list, nb, lo, cnt);
    CU(cudaMemsetAsync(cnt + nb, 0, 4, s));
    RC(st_scan_u32(ctx, cnt, seg_off, nb + 1));
    uint32_t mk = 0;
    CU(cudaMemcpyAsync(&mk, seg_off + nb, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    ctx->stats.launches += 2;
    RC(sp->sroots.reserve(ctx, 32ull * nb + 64));
    if (mk) {
        RC(sp->sb.reserve(ctx, 32ull * mk + 64));                       // gathered keys
        RC(sp->sc.reserve(ctx, 4ull * (mk + 2) * 2 + 8ull * (mk + 2) * 2 + 16ull * mk + 64));
        uint8_t* gkeys = (uint8_t*)sp->sb.ptr;
        uint64_t* gsize = (uint64_t*)sp->sc.ptr;
        uint64_t* gvoff = gsize + mk + 2;
        SRec* grec = (SRec*)(gvoff + mk + 2);
        uint32_t* gkoff = (uint32_t*)(grec + mk);
        uint32_t* seg_of_key = gkoff + mk + 2;
        RC(sp->si.reserve(ctx, 66ull * mk + 64));
        uint8_t* gcache = (uint8_t*)sp->si.ptr;
        uint8_t* gcache_out = gcache + 33ull * mk;
        uint8_t* cache_tab = (uint8_t*)sp->cache[sp->cur].ptr;
        st_gather_keys_kernel<<<grid1d(dev, nb, 256, 32), 256, 0, s>>>(table, recs, lo, seg_off, nb, gkeys, gkoff, seg_of_key, gsize, grec, cache_tab, gcache);
        const uint32_t last = 32u * mk;
        CU(cudaMemcpyAsync(gkoff + mk, &last, 4, cudaMemcpyHostToDevice, s));
        RC(scan_sizes(ctx, gsize, gvoff, mk));
        uint64_t vbytes = 0;
        CU(cudaMemcpyAsync(&vbytes, gvoff + mk, 8, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        RC(sp->sd.reserve(ctx, vbytes + 64));
        st_gather_vals_kernel<<<grid1d(dev, mk, 256, 32), 256, 0, s>>>((const uint8_t*)sp->arena.ptr, grec, gvoff, mk, (uint8_t*)sp->sd.ptr);
        ctx->stats.launches += 3;
        tr.mark("  bucket ranges + gathers");
        RC(ctx->build_forest(gkeys, gkoff, (const uint8_t*)sp->sd.ptr, gvoff, mk, seg_off, nb, seg_of_key, (uint8_t*)sp->sroots.ptr, -1, L, gcache, gcache_out));
        tr.mark("  build_forest");
        st_scatter_cache_kernel<<<grid1d(dev, nb, 256, 32), 256, 0, s>>>(lo, seg_off, nb, gcache_out, cache_tab);
        ctx->stats.launches++;
    }
    if (L == 0) { // one bucket: its root is the trie's root
        CU(cudaMemcpyAsync(sp->root, sp->sroots.ptr, 32, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return PHANT_GPU_OK;
    }
    st_scatter_roots_kernel<<<grid1d(dev, nb, 256), 256, 0, s>>>((const uint8_t*)sp->sroots.ptr, d_list, cnt, nb, top + 32 * level_base(L), pres + level_base(L));
    ctx->stats.launches++;
    // dense levels bottom-up: parents of the dirty children
    static bool attr[64] = {false};
    bool& opted = attr[(dev >= 0 && dev < 64) ? dev : 0];
    if (!opted) { CU(cudaFuncSetAttribute(st_top_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM)); opted = true; }
    RC(sp->se.reserve(ctx, 4ull * (nb + 2) * 5 + 64));
    uint32_t* par = (uint32_t*)sp->se.ptr;
    uint32_t* flag = par + nb + 2;
    uint32_t* pos = flag + nb + 2;
    uint32_t* ping[2] = {pos + nb + 2, pos + 2 * (nb + 2)};
    RC(ctx->d_b3.reserve(ctx, 64));
    uint32_t* viol = (uint32_t*)ctx->d_b3.ptr + 12;
    CU(cudaMemsetAsync(viol, 0, 4, s));
    const uint32_t* child = d_list;
    uint32_t cbound = nb;                       // upper bound of the children count; the exact count stays on the device
    uint32_t* cnt_dev = (uint32_t*)ctx->d_b3.ptr + 13;
    CU(cudaMemcpyAsync(cnt_dev, &nb, 4, cudaMemcpyHostToDevice, s));
    const unsigned fr_cap = (unsigned)keccak_num_sms(dev) * 3;
    for (int d = (int)L - 1; d >= 0; --d) {
        const uint32_t* plist = nullptr;
        uint32_t pbound = 1u << (4 * d);
        const uint32_t* pcount_dev = nullptr;
        if (!all) { // distinct parents of the (sorted) dirty children
            uint32_t* uniq = ping[d & 1];
            st_parent_flag_kernel<<<grid1d(dev, cbound + 1, 256), 256, 0, s>>>(child, cnt_dev, cbound, par, flag);
            RC(st_scan_u32(ctx, flag, pos, cbound + 1));
            st_compact_kernel<<<grid1d(dev, cbound, 256), 256, 0, s>>>(par, flag, pos, cbound, uniq);
            CU(cudaMemcpyAsync(cnt_dev, pos + cbound, 4, cudaMemcpyDeviceToDevice, s)); // the parents are the next level's children
            if (cbound < pbound) pbound = cbound;
            plist = uniq;
            child = uniq;
            cbound = pbound;
            pcount_dev = cnt_dev;
            ctx->stats.launches += 3;
        }
        const unsigned g = (pbound + FR_WARPS * 32 - 1) / (FR_WARPS * 32);
        st_top_branch_kernel<<<g < fr_cap ? g : fr_cap, FR_WARPS * 32, FR_SMEM, s>>>(top + 32 * level_base(d + 1), pres + level_base(d + 1), plist, pbound, pcount_dev,
                                                                               top + 32 * level_base(d), pres + level_base(d), viol);
        ctx->stats.launches++;
    }
    CU(cudaGetLastError());
    tr.mark("  dense levels");
    uint32_t v = 0;
    CU(cudaMemcpyAsync(sp->root, top, 32, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&v, viol, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return v ? 1 : PHANT_GPU_OK; // 1 = the dense-top premise does not hold at this L
}

int st_set_L_and_rebuild_all(phant_gpu_trie* t, uint32_t L)
{
    phant_gpu_ctx* ctx = t->ctx;
    SparseTrie* sp = t->sp;
    for (;;) {
        sp->L = L;
        const uint64_t nodes = level_base(L + 1);
        RC(sp->top.reserve(ctx, 32 * nodes + 64));
        RC(sp->present.reserve(ctx, nodes + 64));
        CU(cudaMemsetAsync(sp->present.ptr, 0, nodes, ctx->stream));
        sp->rebuilds++;
        const int rc = st_rebuild(t, nullptr, 1u << (4 * L), true);
        if (rc != 1) return rc;
        if (L == 0) return PHANT_GPU_E_CUDA; // cannot happen: L = 0 has no dense level
        --L; // some top node has a single child: fewer dense levels
    }
}

int strie_update(phant_gpu_trie* t, const uint8_t* keys32, const uint8_t* vals, const uint32_t* val_off, uint64_t m64, uint8_t out_root[32])
{
    phant_gpu_ctx* ctx = t->ctx;
    SparseTrie* sp = t->sp;
    cudaStream_t s = ctx->stream;
    const int dev = ctx->device;
    if (ctx->flags & PHANT_GPU_FLAG_DEVICE_PTRS) return PHANT_GPU_E_INVALID; // host tables (it is the StateDB flattening, as S)
    if (m64 == 0) { memcpy(out_root, sp->root, 32); return PHANT_GPU_OK; }
    if (!keys32 || !val_off || m64 >= (1ull << 28) || sp->n + m64 >= (1ull << 30)) return PHANT_GPU_E_INVALID;
    const uint32_t m = (uint32_t)m64, n = (uint32_t)sp->n;
    for (uint32_t i = 0; i < m; ++i)
        if (val_off[i + 1] < val_off[i]) return PHANT_GPU_E_INVALID;
    const uint64_t vb = val_off[m];
    if (vb && !vals) return PHANT_GPU_E_INVALID;
    // ---- stage + sort the dirty keys ----
    RC(sp->sg.reserve(ctx, 32ull * m * 2 + vb + 64 + 4ull * (m + 2) * 8 + 8ull * (m + 2) * 2 + 16ull * m + m + 256));
    uint8_t* raw_k = (uint8_t*)sp->sg.ptr;
    uint8_t* dk = raw_k + 32ull * m;                         // sorted keys
    uint8_t* dv = dk + 32ull * m;                            // values as given
    uint32_t* u32 = (uint32_t*)(dv + ((vb + 63) & ~63ull));
    uint32_t* raw_voff = u32;                                // m+1 (+1 pad)
    uint32_t* perm = raw_voff + m + 2;
    uint32_t* dvoff = perm + m + 2;                          // sorted: start offset; dvoff[j+1] is NOT the end (values stay in given order)
    uint32_t* dlen = dvoff + m + 2;
    uint32_t* lb = dlen + m + 2;
    uint32_t* ins_flag = lb + m + 2;
    uint32_t* ins_index = ins_flag + m + 2;
    uint32_t* bucket = ins_index + m + 2;
    uint64_t* app_size = (uint64_t*)(bucket + m + 2);
    uint64_t* app_off = app_size + m + 2;
    SRec* drec = (SRec*)(app_off + m + 2);
    uint8_t* kind = (uint8_t*)(drec + m);
    CU(cudaMemcpyAsync(raw_k, keys32, 32ull * m, cudaMemcpyHostToDevice, s));
    if (vb) CU(cudaMemcpyAsync(dv, vals, vb, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(raw_voff, val_off, 4ull * (m + 1), cudaMemcpyHostToDevice, s));
    ctx->stats.h2d_bytes += 32ull * m + vb + 4ull * (m + 1);
    PhaseTrace tr(s);
    tr.mark("stage dirty (H2D)");
    RC(ctx->sort_by_segment_and_hash(raw_k, nullptr, m, perm, sp->ssort));
    tr.mark("sort dirty keys");
    gather_rows32_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(raw_k, perm, m, dk);
    st_gather_voff_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(raw_voff, perm, m, dvoff, dlen);
    // ---- classify against the table ----
    RC(sp->sh.reserve(ctx, 4ull * (n + 3) * 5 + 64));
    uint32_t* del_flag = (uint32_t*)sp->sh.ptr;
    uint32_t* ins_at = del_flag + n + 3;
    uint32_t* keep = ins_at + n + 3;
    uint32_t* Kscan = keep + n + 3;
    uint32_t* Iscan = Kscan + n + 3;
    CU(cudaMemsetAsync(del_flag, 0, 4ull * (n + 3) * 2, s));
    RC(ctx->d_b3.reserve(ctx, 64));
    uint32_t* counters = (uint32_t*)ctx->d_b3.ptr;
    CU(cudaMemsetAsync(counters, 0, 64, s));
    st_classify_kernel<<<grid1d(dev, m, 128), 128, 0, s>>>((const uint8_t*)sp->keys[sp->cur].ptr, n, dk, dlen, m, lb, kind, del_flag, ins_at, ins_flag,
                                                           app_size, counters);
    RC(scan_sizes(ctx, app_size, app_off, m));
    RC(st_scan_u32(ctx, ins_flag, ins_index, m));
    uint32_t hc[4] = {0, 0, 0, 0};
    uint64_t app_bytes = 0;
    CU(cudaMemcpyAsync(hc, counters, 16, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&app_bytes, app_off + m, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    ctx->stats.launches += 5;
    tr.mark("classify + scans + readback");
    if (hc[2]) return PHANT_GPU_E_INVALID; // the same key twice in one update
    const uint32_t n_ins = hc[0], n_del = hc[1];
    // ---- values into the arena (grown with contents preserved; compaction is a rebuild-time concern) ----
    if (sp->arena_used + app_bytes + 64 > sp->arena.cap) {
        DevBuf bigger;
        RC(bigger.reserve(ctx, (sp->arena_used + app_bytes) * 2 + (1 << 20)));
        if (sp->arena_used) CU(cudaMemcpyAsync(bigger.ptr, sp->arena.ptr, sp->arena_used, cudaMemcpyDeviceToDevice, s));
        CU(cudaStreamSynchronize(s));
        sp->arena.release();
        sp->arena = bigger;
    }
    SRec* recs_cur = (SRec*)sp->recs[sp->cur].ptr;
    st_append_kernel<<<grid1d(dev, m, 256, 32), 256, 0, s>>>(dv, dvoff, dlen, kind, lb, app_off, m, sp->arena_used, (uint8_t*)sp->arena.ptr, recs_cur, drec,
                                                            (uint8_t*)sp->cache[sp->cur].ptr);
    sp->arena_used += app_bytes;
    ctx->stats.launches++;
    tr.mark("append values");
    // ---- merge (skipped for pure value updates) ----
    const uint32_t new_n = n - n_del + n_ins;
    if (n_ins || n_del) {
        const int nxt = 1 - sp->cur;
        RC(sp->keys[nxt].reserve(ctx, 32ull * new_n + 64));
        RC(sp->recs[nxt].reserve(ctx, 16ull * new_n + 64));
        RC(sp->cache[nxt].reserve(ctx, 33ull * new_n + 64));
        st_keep_kernel<<<grid1d(dev, n + 1, 256), 256, 0, s>>>(del_flag, n, keep);
        RC(st_scan_u32(ctx, keep, Kscan, n + 1));
        RC(st_scan_u32(ctx, ins_at, Iscan, n + 2));
        if (n)
            st_merge_table_kernel<<<grid1d(dev, n, 256), 256, 0, s>>>((const uint8_t*)sp->keys[sp->cur].ptr, recs_cur, n, del_flag, Kscan, Iscan,
                                                                     (uint8_t*)sp->keys[nxt].ptr, (SRec*)sp->recs[nxt].ptr,
                                                                     (const uint8_t*)sp->cache[sp->cur].ptr, (uint8_t*)sp->cache[nxt].ptr);
        st_merge_dirty_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(dk, drec, kind, lb, ins_index, m, n, Kscan, (uint8_t*)sp->keys[nxt].ptr,
                                                                 (SRec*)sp->recs[nxt].ptr, (uint8_t*)sp->cache[nxt].ptr);
        ctx->stats.launches += 5;
        sp->cur = nxt;
        sp->n = new_n;
    }
    tr.mark("merge");
    sp->updates++;
    if (new_n == 0) {
        sp->L = 0;
        memcpy(sp->root, EMPTY_ROOT_H, 32);
        memcpy(out_root, sp->root, 32);
        return PHANT_GPU_OK;
    }
    // ---- which part of the trie to re-hash ----
    const uint32_t Lt = st_target_L(new_n);
    int rc;
    if (Lt > sp->L || Lt + 1 < sp->L || n == 0) {
        rc = st_set_L_and_rebuild_all(t, Lt);      // the table grew / shrank past a bucket-size bound: new dense depth
    } else {
        uint32_t* first = ins_flag;                 // (classification scratch is free again)
        uint32_t* fpos = ins_index;
        uint32_t* list = lb;
        st_bucket_flag_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(dk, m, sp->L, bucket, first);
        CU(cudaMemsetAsync(first + m, 0, 4, s));
        RC(st_scan_u32(ctx, first, fpos, m + 1));
        st_compact_kernel<<<grid1d(dev, m, 256), 256, 0, s>>>(bucket, first, fpos, m, list);
        uint32_t nb = 0;
        CU(cudaMemcpyAsync(&nb, fpos + m, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        ctx->stats.launches += 3;
        rc = st_rebuild(t, list, nb, false);
        if (rc == 1) rc = st_set_L_and_rebuild_all(t, sp->L - 1); // a dense node lost all but one child: fewer dense levels
    }
    tr.mark("rebuild (buckets + top)");
    if (rc) return rc;
    memcpy(out_root, sp->root, 32);
    ctx->stats.d2h_bytes += 36;
    return PHANT_GPU_OK;
}

} // namespace

extern "C" int phant_gpu_trie_open(phant_gpu_ctx* ctx, const phant_gpu_trie_desc* desc, phant_gpu_trie** out)
{
    if (!ctx || !desc || !out) return PHANT_GPU_E_INVALID;
    *out = nullptr;
    if (desc->kind == 1) { // sparse resident secure trie, initially empty
        CU(cudaSetDevice(ctx->device));
        phant_gpu_trie* t = new (std::nothrow) phant_gpu_trie();
        SparseTrie* sp = new (std::nothrow) SparseTrie();
        if (!t || !sp) { delete t; delete sp; return PHANT_GPU_E_OOM; }
        t->ctx = ctx; t->kind = 1; t->depth = 0; t->sp = sp;
        memcpy(sp->root, EMPTY_ROOT_H, 32);
        *out = t;
        return PHANT_GPU_OK;
    }
    if (desc->kind != 0 || desc->depth < 1 || desc->depth > 7) return PHANT_GPU_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    phant_gpu_trie* t = new (std::nothrow) phant_gpu_trie();
    if (!t) return PHANT_GPU_E_OOM;
    t->ctx = ctx;
    t->depth = desc->depth;
    uint64_t total = 0, cnt = 1;
    std::vector<uint64_t> offs;
    for (uint32_t l = 0; l <= desc->depth; ++l) { offs.push_back(total); total += 32 * cnt; cnt *= 16; }
    if (int rc = t->store.reserve(ctx, total)) { delete t; return rc; }
    for (uint32_t l = 0; l <= desc->depth; ++l) t->level.push_back((uint8_t*)t->store.ptr + offs[l]);
    const uint64_t n_leaves = cnt / 16;
    ctrie_fill_leaves_kernel<<<grid1d(ctx->device, n_leaves, 256), 256, 0, ctx->stream>>>(desc->seed, n_leaves, t->level[desc->depth]);
    ctx->stats.launches++;
    // all branch levels bottom-up; explicit parent lists keep the batches simple
    for (int l = (int)desc->depth - 1; l >= 0; --l) {
        uint64_t n_l = 1;
        for (int k = 0; k < l; ++k) n_l *= 16;
        DevBuf ids;
        if (int rc = ids.reserve(ctx, 4 * n_l + 16)) { t->store.release(); delete t; return rc; }
        iota_kernel<<<grid1d(ctx->device, n_l, 256), 256, 0, ctx->stream>>>((uint32_t*)ids.ptr, (uint32_t)n_l);
        int rc = ctrie_hash_level(t, (uint32_t)l, (const uint32_t*)ids.ptr, n_l);
        cudaStreamSynchronize(ctx->stream);
        ids.release();
        if (rc) { t->store.release(); t->work.release(); delete t; return rc; }
    }
    *out = t;
    return PHANT_GPU_OK;
}

extern "C" int phant_gpu_trie_root(phant_gpu_trie* t, uint8_t out_root[32])
{
    if (!t || !out_root) return PHANT_GPU_E_INVALID;
    phant_gpu_ctx* ctx = t->ctx;
    if (t->kind == 1) { memcpy(out_root, t->sp->root, 32); return PHANT_GPU_OK; }
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(out_root, t->level[0], 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PHANT_GPU_OK;
}

extern "C" int phant_gpu_trie_update(phant_gpu_trie* t, const uint8_t* keys32, const uint8_t* leaf_vals, const uint32_t* val_off,
                                     uint64_t n_dirty, uint8_t out_root[32])
{
    if (!t || !out_root) return PHANT_GPU_E_INVALID;
    phant_gpu_ctx* ctx = t->ctx;
    CU(cudaSetDevice(ctx->device));
    if (t->kind == 1) return strie_update(t, keys32, leaf_vals, val_off, n_dirty, out_root);
    cudaStream_t s = ctx->stream;
    if (n_dirty == 0) return phant_gpu_trie_root(t, out_root);
    if (!keys32 || !val_off || n_dirty >= (1ull << 28)) return PHANT_GPU_E_INVALID;
    const uint8_t* d_keys = keys32; const uint8_t* d_vals = leaf_vals; const uint32_t* d_voff = val_off;
    uint64_t vb = 0, max_val = ~0ull;
    if (!(ctx->flags & PHANT_GPU_FLAG_DEVICE_PTRS)) {
        max_val = 0;
        for (uint64_t i = 0; i < n_dirty; ++i) {
            if (val_off[i + 1] < val_off[i]) return PHANT_GPU_E_INVALID;
            if ((uint64_t)(val_off[i + 1] - val_off[i]) > max_val) max_val = val_off[i + 1] - val_off[i];
        }
        vb = val_off[n_dirty];
        if (vb && !leaf_vals) return PHANT_GPU_E_INVALID;
        RC(ctx->d_msgs.reserve(ctx, 32 * n_dirty + vb + 4 * (n_dirty + 1) + 256));
        uint8_t* dk = (uint8_t*)ctx->d_msgs.ptr;
        uint8_t* dv = dk + 32 * n_dirty;
        uint32_t* dvo = (uint32_t*)(dv + ((vb + 63) & ~63ull));
        CU(cudaMemcpyAsync(dk, keys32, 32 * n_dirty, cudaMemcpyHostToDevice, s));
        if (vb) CU(cudaMemcpyAsync(dv, leaf_vals, vb, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dvo, val_off, 4 * (n_dirty + 1), cudaMemcpyHostToDevice, s));
        ctx->stats.h2d_bytes += 32 * n_dirty + vb + 4 * (n_dirty + 1);
        d_keys = dk; d_vals = dv; d_voff = dvo;
    } else {
        CU(cudaMemcpyAsync(&vb, val_off + n_dirty, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    const uint32_t L = t->depth;
    if (!(ctx->flags & PHANT_GPU_FLAG_DEVICE_PTRS) && max_val <= FR_SLOT - 48) {
        // ---- fused frontier path: one launch for the leaves, then per level shift/pad + unique + one hash launch; the only
        // host synchronisation is the final read of the root ----
        static bool attr[64] = {false}; // function attributes are per device (as launch_staged in keccak_kernels.cu)
        bool& opted = attr[(ctx->device >= 0 && ctx->device < 64) ? ctx->device : 0];
        if (!opted) {
            CU(cudaFuncSetAttribute(frontier_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM));
            CU(cudaFuncSetAttribute(frontier_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM));
            opted = true;
        }
        RC(ctx->d_b0.reserve(ctx, 4 * n_dirty * 4 + 256));
        uint32_t* pos = (uint32_t*)ctx->d_b0.ptr;
        uint32_t* cur = pos + n_dirty;
        uint32_t* tmp = cur + n_dirty;
        uint32_t* uniq = tmp + n_dirty;
        RC(ctx->d_b3.reserve(ctx, 64));
        uint32_t* counts = (uint32_t*)ctx->d_b3.ptr; // counts[0] = valid entries of `cur`; counts[8] = "refused" flag
        uint32_t* refuse = counts + 8;
        const unsigned fr_grid = (unsigned)((n_dirty + FR_WARPS * 32 - 1) / (FR_WARPS * 32));
        const unsigned fr_cap = (unsigned)keccak_num_sms(ctx->device) * 3;
        // positions first, sorted, checked for collisions; only then is anything written to the resident levels
        ctrie_leaf_pos_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(d_keys, n_dirty, L, pos);
        size_t temp = 0;
        CU(cub::DeviceRadixSort::SortKeys(nullptr, temp, (const uint32_t*)pos, cur, (int64_t)n_dirty, 0, 4 * (int)L, s));
        RC(ctx->d_cub.reserve(ctx, temp));
        CU(cub::DeviceRadixSort::SortKeys(ctx->d_cub.ptr, temp, (const uint32_t*)pos, cur, (int64_t)n_dirty, 0, 4 * (int)L, s));
        const uint32_t init[16] = {(uint32_t)n_dirty, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        CU(cudaMemcpyAsync(counts, init, sizeof init, cudaMemcpyHostToDevice, s));
        ctrie_dup_check_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(cur, n_dirty, refuse);
        frontier_leaf_kernel<<<fr_grid < fr_cap ? fr_grid : fr_cap, FR_WARPS * 32, FR_SMEM, s>>>(d_keys, d_vals, d_voff, (uint32_t)n_dirty, L, t->level[L], pos, refuse);
        ctx->stats.launches += 4;
        uint64_t bound = n_dirty;
        for (int l = (int)L - 1; l >= 0; --l) {
            shift4_pad_kernel<<<grid1d(ctx->device, bound, 256), 256, 0, s>>>(cur, counts, (uint32_t)bound, tmp);
            size_t t2 = 0;
            CU(cub::DeviceSelect::Unique(nullptr, t2, (const uint32_t*)tmp, uniq, counts, (int64_t)bound, s));
            RC(ctx->d_cub.reserve(ctx, t2));
            CU(cub::DeviceSelect::Unique(ctx->d_cub.ptr, t2, (const uint32_t*)tmp, uniq, counts, (int64_t)bound, s));
            uint64_t level_nodes = 1;
            for (int q = 0; q < l; ++q) level_nodes *= 16;
            if (bound > level_nodes) bound = level_nodes; // a level cannot have more dirty nodes than nodes
            const unsigned g = (unsigned)((bound + FR_WARPS * 32 - 1) / (FR_WARPS * 32));
            frontier_branch_kernel<<<g < fr_cap ? g : fr_cap, FR_WARPS * 32, FR_SMEM, s>>>(t->level[l + 1], uniq, counts, (uint32_t)bound, t->level[l], refuse);
            ctx->stats.launches += 3;
            uint32_t* x = cur; cur = uniq; uniq = x;
        }
        CU(cudaGetLastError());
        uint32_t refused = 0;
        CU(cudaMemcpyAsync(out_root, t->level[0], 32, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&refused, refuse, 4, cudaMemcpyDeviceToHost, s));
        ctx->stats.d2h_bytes += 36;
        CU(cudaStreamSynchronize(s));
        return refused ? PHANT_GPU_E_INVALID : PHANT_GPU_OK; // refused: the trie is unchanged, out_root = its current root
    }
    // ---- general path (device pointers, or leaf values too large for a staging slot):
    // dirty leaves: encode, hash (batched Keccak), scatter into the leaf level ----
    RC(ctx->d_b0.reserve(ctx, 4 * n_dirty * 4 + 8 * (n_dirty + 2) * 2 + 32 * n_dirty + 256));
    uint32_t* pos = (uint32_t*)ctx->d_b0.ptr;
    uint32_t* pa = pos + n_dirty;
    uint32_t* pb = pa + n_dirty;
    uint32_t* pc = pb + n_dirty;
    uint64_t* sizes = (uint64_t*)(((uintptr_t)(pc + n_dirty) + 15) & ~(uintptr_t)15);
    uint64_t* offs = sizes + ((n_dirty + 2) & ~1ull);
    uint8_t* dg = (uint8_t*)(offs + ((n_dirty + 2) & ~1ull)); // 16-byte aligned
    ctrie_leaf_pos_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(d_keys, n_dirty, L, pos);
    {   // distinct leaf positions, checked before anything is written to the resident levels
        size_t temp0 = 0;
        CU(cub::DeviceRadixSort::SortKeys(nullptr, temp0, (const uint32_t*)pos, pa, (int64_t)n_dirty, 0, 4 * (int)L, s));
        RC(ctx->d_cub.reserve(ctx, temp0));
        CU(cub::DeviceRadixSort::SortKeys(ctx->d_cub.ptr, temp0, (const uint32_t*)pos, pa, (int64_t)n_dirty, 0, 4 * (int)L, s));
        RC(ctx->d_b3.reserve(ctx, 64));
        CU(cudaMemsetAsync(ctx->d_b3.ptr, 0, 64, s));
        ctrie_dup_check_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(pa, n_dirty, (uint32_t*)ctx->d_b3.ptr);
        uint32_t refused = 0;
        CU(cudaMemcpyAsync(&refused, ctx->d_b3.ptr, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        ctx->stats.launches += 2;
        if (refused) return PHANT_GPU_E_INVALID;
    }
    ctrie_leaf_size_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(d_vals, d_voff, n_dirty, L, sizes);
    RC(scan_sizes(ctx, sizes, offs, n_dirty));
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, offs + n_dirty, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    RC(ctx->d_b1.reserve(ctx, total + 64));
    ctrie_leaf_encode_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(d_keys, d_vals, d_voff, n_dirty, L, offs, (uint8_t*)ctx->d_b1.ptr);
    ctx->stats.launches += 3;
    RC(ctx->hash_csr((const uint8_t*)ctx->d_b1.ptr, offs, n_dirty, total, dg));
    ctrie_scatter_kernel<<<grid1d(ctx->device, n_dirty, 256), 256, 0, s>>>(dg, pos, n_dirty, t->level[L]);
    ctx->stats.launches++;
    // ---- frontier: unique parents level by level (`pa` = the sorted positions, then shift + unique) ----
    uint64_t cnt = n_dirty;
    uint32_t* cur = pa;
    uint32_t* tmp = pb;
    uint32_t* uniq = pc;
    for (int l = (int)L - 1; l >= 0; --l) {
        shift4_kernel<<<grid1d(ctx->device, cnt, 256), 256, 0, s>>>(cur, cnt, tmp);
        size_t t2 = 0;
        CU(cub::DeviceSelect::Unique(nullptr, t2, (const uint32_t*)tmp, uniq, (uint32_t*)ctx->d_b3.ptr, (int64_t)cnt, s));
        RC(ctx->d_cub.reserve(ctx, t2));
        CU(cub::DeviceSelect::Unique(ctx->d_cub.ptr, t2, (const uint32_t*)tmp, uniq, (uint32_t*)ctx->d_b3.ptr, (int64_t)cnt, s));
        uint32_t nu = 0;
        CU(cudaMemcpyAsync(&nu, ctx->d_b3.ptr, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        ctx->stats.launches += 2;
        cnt = nu;
        RC(ctrie_hash_level(t, (uint32_t)l, uniq, cnt));
        uint32_t* x = cur; cur = uniq; uniq = x; // the unique parents are the next level's (sorted) children
    }
    CU(cudaMemcpyAsync(out_root, t->level[0], 32, cudaMemcpyDeviceToHost, s));
    ctx->stats.d2h_bytes += 32;
    CU(cudaStreamSynchronize(s));
    return PHANT_GPU_OK;
}

extern "C" void phant_gpu_trie_close(phant_gpu_trie* t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    t->store.release();
    t->work.release();
    if (t->sp) {
        SparseTrie* sp = t->sp;
        for (DevBuf* b : {&sp->keys[0], &sp->keys[1], &sp->recs[0], &sp->recs[1], &sp->cache[0], &sp->cache[1], &sp->si, &sp->arena, &sp->top, &sp->present, &sp->sa, &sp->sb, &sp->sc, &sp->sd,
                          &sp->se, &sp->sf, &sp->sg, &sp->sh, &sp->sroots, &sp->ssort}) b->release();
        delete sp;
    }
    delete t;
}

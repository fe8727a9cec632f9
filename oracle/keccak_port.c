/* keccak_port.c -- TEST INFRASTRUCTURE (see oracle.h).  Plain-C restatement of Keccak-256.
 *
 * Follows the sponge of ethash/lib/keccak/keccak.c:301-354 (rate 136 B, whole blocks absorbed as
 * little-endian 64-bit words, tail + 0x01 pad byte, 0x80 into the last rate byte, one final
 * permutation, first 32 state bytes out) and the Keccak-f[1600] round structure of
 * keccak.c:58-269 (theta, rho+pi, chi, iota; round constants keccak.c:38-46), written as loops
 * over a 5x5 lane array rather than the reference's unrolled Aba..Asu variables.
 * phant's src/crypto/hasher.zig:4-8 delegates to Zig std Keccak256, which is the same function.
 */
#include "oracle.h"
#include <dlfcn.h>
#include <string.h>

static const uint64_t RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
    0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

/* rho rotation of lane (x,y) at index x+5y */
static const unsigned RHO[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                 25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};

static inline uint64_t rol64(uint64_t x, unsigned s) { return s ? (x << s) | (x >> (64 - s)) : x; }

static void keccak_f1600(uint64_t a[25])
{
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        /* theta apply + rho + pi:  B[y, 2x+3y] = rol(A[x,y] ^ D[x], r[x,y]) */
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) {
                int nx = y, ny = (2 * x + 3 * y) % 5;
                b[nx + 5 * ny] = rol64(a[x + 5 * y] ^ d[x], RHO[x + 5 * y]);
            }
        /* chi */
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x)
                a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[round];
    }
}

static inline uint64_t load_le64(const uint8_t* p)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

void oracle_keccak256_port(const uint8_t* data, size_t len, uint8_t out[32])
{
    enum { RATE = 136 };
    uint64_t st[25];
    memset(st, 0, sizeof st);
    while (len >= RATE) {
        for (int i = 0; i < RATE / 8; ++i) st[i] ^= load_le64(data + 8 * i);
        keccak_f1600(st);
        data += RATE;
        len -= RATE;
    }
    uint8_t last[RATE];
    memset(last, 0, sizeof last);
    if (len) memcpy(last, data, len);
    last[len] ^= 0x01;
    last[RATE - 1] ^= 0x80;
    for (int i = 0; i < RATE / 8; ++i) st[i] ^= load_le64(last + 8 * i);
    keccak_f1600(st);
    for (int i = 0; i < 4; ++i)
        for (int b = 0; b < 8; ++b) out[8 * i + b] = (uint8_t)(st[i] >> (8 * b));
}

static oracle_keccak_fn g_keccak = oracle_keccak256_port;

void oracle_set_keccak(oracle_keccak_fn fn) { g_keccak = fn ? fn : oracle_keccak256_port; }

void oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) { g_keccak(data, len, out); }

/* Plug the reference's own compiled Keccak (oracle/_ref/libref_keccak.so, built by oracle/Makefile from
 * ethash/lib/keccak/keccak.c where it lies) under every oracle function.  0 on success. */
typedef struct { uint8_t b[32]; } ref_h256; /* layout of union ethash_hash256, ethash/include/ethash/hash_types.h:15-21 */
static ref_h256 (*g_ref_keccak)(const uint8_t*, size_t);
static void ref_trampoline(const uint8_t* data, size_t len, uint8_t out[32])
{
    ref_h256 h = g_ref_keccak(data, len);
    memcpy(out, h.b, 32);
}
int oracle_use_ref_keccak(const char* so_path)
{
    void* h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    *(void**)(&g_ref_keccak) = dlsym(h, "ethash_keccak256");
    if (!g_ref_keccak) return -2;
    g_keccak = ref_trampoline;
    return 0;
}

void oracle_keccak256_batch(const uint8_t* msgs, const uint64_t* off, uint64_t n, uint8_t* out32, int threads)
{
    oracle_keccak_fn fn = g_keccak;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) fn(msgs + off[i], (size_t)(off[i + 1] - off[i]), out32 + 32 * i);
}

/* logs bloom: Receipt.addToBloom (src/types/receipt.zig:50-63; same arithmetic as evmone/test/state/bloom_filter.cpp:17-31) */
void oracle_logs_bloom(const uint8_t* items, const uint64_t* item_off, const uint32_t* bloom_of_item, uint64_t n_items,
                       uint64_t n_blooms, uint8_t* blooms)
{
    memset(blooms, 0, 256 * n_blooms);
    for (uint64_t i = 0; i < n_items; ++i) {
        uint8_t h[32];
        oracle_keccak256(items + item_off[i], item_off[i + 1] - item_off[i], h);
        uint8_t* bloom = blooms + 256 * (uint64_t)bloom_of_item[i];
        for (int j = 0; j < 3; ++j) {
            unsigned bit_to_set = (((unsigned)h[2 * j] << 8) | h[2 * j + 1]) & 0x07ff;
            unsigned bit_index = 0x07ff - bit_to_set;
            bloom[bit_index / 8] |= (uint8_t)(1u << (7 - bit_index % 8));
        }
    }
}

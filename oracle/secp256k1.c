/* secp256k1.c -- CPU restatement of ECDSA public-key recovery on secp256k1.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * What it states: TxSigner.get_sender's `self.ecdsa_signer.erecover(sig, tx_hash)` (reference src/signer/signer.zig:78,
 * src/crypto/ecdsa.zig:19-21).  The arithmetic itself is NOT in the reference tree: ecdsa.zig:2 imports the third-party
 * package zig-eth-secp256k1 (build.zig.zon:9-12, commit 95b7f93), a thin wrapper over bitcoin-core/libsecp256k1's
 * secp256k1_ecdsa_recover.  This file restates that published algorithm (SEC 1 v2 section 4.1.6):
 *
 *   sig = r(32, big endian) || s(32) || recid(1);  reject unless 0 < r < n, 0 < s < n, recid <= 3
 *   x = r + (recid & 2 ? n : 0), reject unless x < p;  y = sqrt(x^3 + 7) with parity recid & 1, reject if no root
 *   Q = r^-1 (s R - z G), z = the 32-byte message hash as an integer (mod n);  reject if Q is the point at infinity
 *   output 0x04 || X(32) || Y(32)
 *
 * Parity is pinned by the reference's own vectors: the geth-generated erecover vector of ecdsa.zig:38-48 and the two
 * mainnet transactions with known senders of signer.zig:199-227 (tests/golden/ecrecover_kat.json), and cross-checked
 * against OpenSSL (python `cryptography`) on random keys (tests/test_oracle_ecrecover.py).
 *
 * Written for obviousness, not speed: 8 x 32-bit limbs, ONE generic modular multiply (schoolbook product, then fold the
 * high half down with 2^256 mod m until it is gone) used for both the field prime p and the group order n, inversion
 * and square root by plain square-and-multiply, two separate double-and-add scalar multiplications in Jacobian
 * coordinates.  The device code (phant_b200/csrc/secp256k1.cuh) is organised differently on every one of these points.
 */
#include "oracle.h"

#include <string.h>

typedef struct { uint32_t w[8]; } u256; /* little-endian limbs */

static const u256 P = {{0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}};
static const u256 N = {{0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}};
static const u256 GX = {{0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu}};
static const u256 GY = {{0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u}};

static void from_be(u256* r, const uint8_t b[32])
{
    for (int i = 0; i < 8; ++i)
        r->w[i] = ((uint32_t)b[28 - 4 * i] << 24) | ((uint32_t)b[29 - 4 * i] << 16) | ((uint32_t)b[30 - 4 * i] << 8) | b[31 - 4 * i];
}
static void to_be(uint8_t b[32], const u256* a)
{
    for (int i = 0; i < 8; ++i) {
        b[28 - 4 * i] = (uint8_t)(a->w[i] >> 24); b[29 - 4 * i] = (uint8_t)(a->w[i] >> 16);
        b[30 - 4 * i] = (uint8_t)(a->w[i] >> 8);  b[31 - 4 * i] = (uint8_t)a->w[i];
    }
}
static int is_zero(const u256* a) { uint32_t o = 0; for (int i = 0; i < 8; ++i) o |= a->w[i]; return o == 0; }
static int cmp(const u256* a, const u256* b)
{
    for (int i = 7; i >= 0; --i) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}
static uint32_t add_raw(u256* r, const u256* a, const u256* b) /* returns the carry */
{
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) { c += (uint64_t)a->w[i] + b->w[i]; r->w[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
static uint32_t sub_raw(u256* r, const u256* a, const u256* b) /* returns the borrow */
{
    int64_t c = 0;
    for (int i = 0; i < 8; ++i) { c += (int64_t)a->w[i] - b->w[i]; r->w[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)(c & 1);
}
static void add_mod(u256* r, const u256* a, const u256* b, const u256* m)
{
    const uint32_t carry = add_raw(r, a, b);
    if (carry || cmp(r, m) >= 0) sub_raw(r, r, m);
}
static void sub_mod(u256* r, const u256* a, const u256* b, const u256* m)
{
    if (sub_raw(r, a, b)) add_raw(r, r, m);
}
/* r = a * b mod m, for m = 2^256 - k with k < 2^130 (true for p and n): fold the high half down with k until it is gone */
static void mul_mod(u256* r, const u256* a, const u256* b, const u256* m)
{
    uint32_t t[17] = {0};
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
        for (int j = 0; j < 8; ++j) { c += (uint64_t)a->w[i] * b->w[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        t[i + 8] = (uint32_t)c;
    }
    u256 k, zero = {{0}};
    sub_raw(&k, &zero, m); /* 2^256 - m */
    for (;;) {
        uint32_t hi_any = 0;
        for (int i = 8; i < 17; ++i) hi_any |= t[i];
        if (!hi_any) break;
        uint32_t u[17] = {0};
        for (int i = 0; i < 8; ++i) u[i] = t[i];
        for (int i = 0; i < 9; ++i) { /* u += t[8 + i] * k << (32 i) */
            uint64_t c = 0;
            for (int j = 0; j < 8 && i + j < 17; ++j) { c += (uint64_t)t[8 + i] * k.w[j] + u[i + j]; u[i + j] = (uint32_t)c; c >>= 32; }
            for (int q = i + 8; c && q < 17; ++q) { c += u[q]; u[q] = (uint32_t)c; c >>= 32; }
        }
        memcpy(t, u, sizeof t);
    }
    for (int i = 0; i < 8; ++i) r->w[i] = t[i];
    while (cmp(r, m) >= 0) sub_raw(r, r, m);
}
static void pow_mod(u256* r, const u256* a, const u256* e, const u256* m)
{
    u256 acc = {{1}}, base = *a;
    for (int i = 0; i < 256; ++i) {
        if ((e->w[i >> 5] >> (i & 31)) & 1) mul_mod(&acc, &acc, &base, m);
        mul_mod(&base, &base, &base, m);
    }
    *r = acc;
}
static void inv_mod(u256* r, const u256* a, const u256* m) /* m prime: a^(m-2) */
{
    u256 e, two = {{2}};
    sub_raw(&e, m, &two);
    pow_mod(r, a, &e, m);
}

/* ---- curve y^2 = x^3 + 7 over F_p, Jacobian coordinates; inf = (Z == 0) ---- */
typedef struct { u256 x, y, z; } jac;

static void jac_double(jac* r, const jac* a)
{
    if (is_zero(&a->z) || is_zero(&a->y)) { memset(r, 0, sizeof *r); return; }
    u256 s, m, t, y2, x3, y3, z3;
    mul_mod(&y2, &a->y, &a->y, &P);
    mul_mod(&s, &a->x, &y2, &P); add_mod(&s, &s, &s, &P); add_mod(&s, &s, &s, &P);          /* S = 4 X Y^2 */
    mul_mod(&m, &a->x, &a->x, &P); add_mod(&t, &m, &m, &P); add_mod(&m, &t, &m, &P);        /* M = 3 X^2 */
    mul_mod(&x3, &m, &m, &P); sub_mod(&x3, &x3, &s, &P); sub_mod(&x3, &x3, &s, &P);         /* X' = M^2 - 2S */
    mul_mod(&t, &y2, &y2, &P); add_mod(&t, &t, &t, &P); add_mod(&t, &t, &t, &P); add_mod(&t, &t, &t, &P); /* 8 Y^4 */
    sub_mod(&y3, &s, &x3, &P); mul_mod(&y3, &y3, &m, &P); sub_mod(&y3, &y3, &t, &P);        /* Y' = M (S - X') - 8 Y^4 */
    mul_mod(&z3, &a->y, &a->z, &P); add_mod(&z3, &z3, &z3, &P);                             /* Z' = 2 Y Z */
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(jac* r, const jac* a, const jac* b)
{
    if (is_zero(&a->z)) { *r = *b; return; }
    if (is_zero(&b->z)) { *r = *a; return; }
    u256 z1z1, z2z2, u1, u2, s1, s2, h, rr, t, h2, h3, x3, y3, z3;
    mul_mod(&z1z1, &a->z, &a->z, &P); mul_mod(&z2z2, &b->z, &b->z, &P);
    mul_mod(&u1, &a->x, &z2z2, &P); mul_mod(&u2, &b->x, &z1z1, &P);
    mul_mod(&s1, &a->y, &z2z2, &P); mul_mod(&s1, &s1, &b->z, &P);
    mul_mod(&s2, &b->y, &z1z1, &P); mul_mod(&s2, &s2, &a->z, &P);
    sub_mod(&h, &u2, &u1, &P); sub_mod(&rr, &s2, &s1, &P);
    if (is_zero(&h)) {
        if (is_zero(&rr)) { jac_double(r, a); return; }
        memset(r, 0, sizeof *r); return; /* P + (-P) */
    }
    mul_mod(&h2, &h, &h, &P); mul_mod(&h3, &h2, &h, &P);
    mul_mod(&t, &u1, &h2, &P);
    mul_mod(&x3, &rr, &rr, &P); sub_mod(&x3, &x3, &h3, &P); sub_mod(&x3, &x3, &t, &P); sub_mod(&x3, &x3, &t, &P);
    sub_mod(&y3, &t, &x3, &P); mul_mod(&y3, &y3, &rr, &P); mul_mod(&t, &s1, &h3, &P); sub_mod(&y3, &y3, &t, &P);
    mul_mod(&z3, &a->z, &b->z, &P); mul_mod(&z3, &z3, &h, &P);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_mul(jac* r, const jac* p, const u256* k)
{
    jac acc;
    memset(&acc, 0, sizeof acc);
    for (int i = 255; i >= 0; --i) {
        jac_double(&acc, &acc);
        if ((k->w[i >> 5] >> (i & 31)) & 1) jac_add(&acc, &acc, p);
    }
    *r = acc;
}

/* 0 = recovered; negative = the signature does not recover a key (what libsecp256k1 reports as failure) */
int oracle_ecrecover(const uint8_t hash32[32], const uint8_t sig65[65], uint8_t pub65[65])
{
    u256 r, s, z, x, zero = {{0}};
    from_be(&r, sig65); from_be(&s, sig65 + 32); from_be(&z, hash32);
    const uint32_t recid = sig65[64];
    if (recid > 3) return -1;
    if (is_zero(&r) || is_zero(&s) || cmp(&r, &N) >= 0 || cmp(&s, &N) >= 0) return -2;
    x = r;
    if (recid & 2) { if (add_raw(&x, &r, &N) || cmp(&x, &P) >= 0) return -3; }
    /* y^2 = x^3 + 7; p = 3 mod 4, so a root, if any, is (x^3 + 7)^((p+1)/4) */
    u256 y2, y, e, seven = {{7}}, one = {{1}};
    mul_mod(&y2, &x, &x, &P); mul_mod(&y2, &y2, &x, &P); add_mod(&y2, &y2, &seven, &P);
    add_raw(&e, &P, &one); /* p + 1 < 2^256: no carry */
    for (int i = 0; i < 7; ++i) e.w[i] = (e.w[i] >> 2) | (e.w[i + 1] << 30);
    e.w[7] >>= 2;
    pow_mod(&y, &y2, &e, &P);
    u256 chk;
    mul_mod(&chk, &y, &y, &P);
    if (cmp(&chk, &y2) != 0) return -4; /* x is not the abscissa of a curve point */
    if ((y.w[0] & 1) != (recid & 1)) sub_mod(&y, &zero, &y, &P);
    while (cmp(&z, &N) >= 0) sub_raw(&z, &z, &N);
    u256 rinv, u1, u2;
    inv_mod(&rinv, &r, &N);
    mul_mod(&u1, &z, &rinv, &N); sub_mod(&u1, &zero, &u1, &N); /* -z / r */
    mul_mod(&u2, &s, &rinv, &N);                                /*  s / r */
    jac R = {x, y, {{1}}}, G = {GX, GY, {{1}}}, a, b, q;
    jac_mul(&a, &G, &u1);
    jac_mul(&b, &R, &u2);
    jac_add(&q, &a, &b);
    if (is_zero(&q.z)) return -5;
    u256 zi, zi2, zi3, ax, ay;
    inv_mod(&zi, &q.z, &P);
    mul_mod(&zi2, &zi, &zi, &P); mul_mod(&zi3, &zi2, &zi, &P);
    mul_mod(&ax, &q.x, &zi2, &P); mul_mod(&ay, &q.y, &zi3, &P);
    pub65[0] = 0x04;
    to_be(pub65 + 1, &ax); to_be(pub65 + 33, &ay);
    return 0;
}

/* public key of a private key (test helper: lets the tests build signatures' expected answers without OpenSSL) */
int oracle_secp256k1_pubkey(const uint8_t priv32[32], uint8_t pub65[65])
{
    u256 k;
    from_be(&k, priv32);
    if (is_zero(&k) || cmp(&k, &N) >= 0) return -1;
    jac G = {GX, GY, {{1}}}, q;
    jac_mul(&q, &G, &k);
    u256 zi, zi2, zi3, ax, ay;
    inv_mod(&zi, &q.z, &P);
    mul_mod(&zi2, &zi, &zi, &P); mul_mod(&zi3, &zi2, &zi, &P);
    mul_mod(&ax, &q.x, &zi2, &P); mul_mod(&ay, &q.y, &zi3, &P);
    pub65[0] = 0x04;
    to_be(pub65 + 1, &ax); to_be(pub65 + 33, &ay);
    return 0;
}

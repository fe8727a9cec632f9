// secp256k1.cuh -- ECDSA public-key recovery for one signature per thread (sm_100a; also compiles as host code for the
// CPU-side check in tests/hostcheck/ecrecover_host.cpp).
//
// What it replaces: TxSigner.get_sender's `self.ecdsa_signer.erecover(sig, tx_hash)` followed by
// `keccak256(pubkey[1..])[12..]` (reference src/signer/signer.zig:78-79, src/crypto/ecdsa.zig:19-21); the reference
// reaches libsecp256k1's secp256k1_ecdsa_recover through zig-eth-secp256k1 (build.zig.zon, commit 95b7f93), one call per
// transaction.  Same accept / reject decisions (SEC 1 v2 section 4.1.6):
//   0 < r < n, 0 < s < n, recid <= 3;  x = r (+ n if recid & 2) must be < p and on the curve;  y parity = recid & 1;
//   Q = u1 G + u2 R with u1 = -z / r, u2 = s / r (mod n);  Q at infinity fails.
//
// Shape: 4 x 64-bit limbs; p = 2^256 - 0x1000003D1 reduces by folding the high half with a 33-bit constant; the few scalar
// operations (one inversion mod n) use a generic fold with 2^256 - n; the double multiplication is one interleaved pass
// over the bits of (u1, u2) with the table {G, R, G + R} in affine coordinates (Shamir's trick: 256 doublings and at most
// 256 mixed additions), Jacobian accumulator, complete handling of the doubling / cancelling corner cases (inputs are
// adversarial).  Integer work only; no tensor cores, no floating point.
// Tried and dropped (round 2): width-w NAF recoding with a static table of odd multiples of G and a Jacobian table of odd
// multiples of R -- 25% fewer field multiplications on paper, 1.85x SLOWER on the B200 (2.07 M vs 3.82 M signatures/s):
// digit arrays and the Jacobian table live in local memory (2.2 KB per thread), and this kernel is bound by local-memory
// round trips and occupancy, not by multiplications.
#pragma once
#include <stdint.h>

namespace phant {
namespace secp {

#if defined(__CUDA_ARCH__)
#define PHANT_UMULHI(a, b) __umul64hi((a), (b))
#else
#define PHANT_UMULHI(a, b) ((uint64_t)(((unsigned __int128)(a) * (b)) >> 64))
#endif

struct u256 { uint64_t v[4]; }; // little-endian limbs

// ---- limb helpers ----
__device__ __forceinline__ uint64_t adc(uint64_t a, uint64_t b, uint64_t& carry)
{
    const uint64_t s = a + carry;
    const uint64_t c1 = s < carry;
    const uint64_t r = s + b;
    carry = c1 | (uint64_t)(r < b);
    return r;
}
__device__ __forceinline__ uint64_t sbb(uint64_t a, uint64_t b, uint64_t& borrow)
{
    const uint64_t d = a - b;
    const uint64_t b1 = a < b;
    const uint64_t r = d - borrow;
    borrow = b1 | (uint64_t)(d < borrow);
    return r;
}
// (hi, lo) = a * b + c + d   (cannot overflow 128 bits)
__device__ __forceinline__ uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t& hi)
{
    uint64_t lo = a * b;
    uint64_t h = PHANT_UMULHI(a, b);
    lo += c; h += lo < c;
    lo += d; h += lo < d;
    hi = h;
    return lo;
}
__device__ __forceinline__ bool is_zero(const u256& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
__device__ __forceinline__ bool geq(const u256& a, const u256& b)
{
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        if (a.v[i] > b.v[i]) return true;
        if (a.v[i] < b.v[i]) return false;
    }
    return true;
}
__device__ __forceinline__ uint64_t add_raw(u256& r, const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    uint64_t c;
    asm("{\n\t"
        "add.cc.u64 %0, %5, %9;\n\t"
        "addc.cc.u64 %1, %6, %10;\n\t"
        "addc.cc.u64 %2, %7, %11;\n\t"
        "addc.cc.u64 %3, %8, %12;\n\t"
        "addc.u64 %4, 0, 0;\n\t"
        "}"
        : "=&l"(r.v[0]), "=&l"(r.v[1]), "=&l"(r.v[2]), "=&l"(r.v[3]), "=&l"(c)
        : "l"(a.v[0]), "l"(a.v[1]), "l"(a.v[2]), "l"(a.v[3]), "l"(b.v[0]), "l"(b.v[1]), "l"(b.v[2]), "l"(b.v[3]));
    return c;
#else
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = adc(a.v[i], b.v[i], c);
    return c;
#endif
}
__device__ __forceinline__ uint64_t sub_raw(u256& r, const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    uint64_t bw;
    asm("{\n\t"
        "sub.cc.u64 %0, %5, %9;\n\t"
        "subc.cc.u64 %1, %6, %10;\n\t"
        "subc.cc.u64 %2, %7, %11;\n\t"
        "subc.cc.u64 %3, %8, %12;\n\t"
        "subc.u64 %4, 0, 0;\n\t"
        "}"
        : "=&l"(r.v[0]), "=&l"(r.v[1]), "=&l"(r.v[2]), "=&l"(r.v[3]), "=&l"(bw)
        : "l"(a.v[0]), "l"(a.v[1]), "l"(a.v[2]), "l"(a.v[3]), "l"(b.v[0]), "l"(b.v[1]), "l"(b.v[2]), "l"(b.v[3]));
    return bw & 1; // subc of 0 - 0 - borrow: all ones when a borrow came in
#else
    uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = sbb(a.v[i], b.v[i], bw);
    return bw;
#endif
}
// 256 x 256 -> 512 bits.  Device: one asm block per row so that the hardware carry flag links the 64-bit multiply-adds
// (mad.lo.cc / madc.hi.cc); written in C the carries become compare + select pairs, about a third of all instructions of
// the recovery kernel.  Host (tests/hostcheck): the portable statement of the same schoolbook product.
__device__ __forceinline__ void mul_wide(uint64_t (&t)[8], const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    const uint64_t b0 = b.v[0], b1 = b.v[1], b2 = b.v[2], b3 = b.v[3];
    {   // row 0: t[0..4] = a0 * b
        const uint64_t x = a.v[0];
        asm("{\n\t"
            "mul.lo.u64 %0, %5, %6;\n\t"
            "mul.hi.u64 %1, %5, %6;\n\t"
            "mul.hi.u64 %2, %5, %7;\n\t"
            "mul.hi.u64 %3, %5, %8;\n\t"
            "mul.hi.u64 %4, %5, %9;\n\t"
            "mad.lo.cc.u64 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u64 %2, %5, %8, %2;\n\t"
            "madc.lo.cc.u64 %3, %5, %9, %3;\n\t"
            "addc.u64 %4, %4, 0;\n\t"
            "}"
            : "=&l"(t[0]), "=&l"(t[1]), "=&l"(t[2]), "=&l"(t[3]), "=&l"(t[4])
            : "l"(x), "l"(b0), "l"(b1), "l"(b2), "l"(b3));
    }
#pragma unroll
    for (int i = 1; i < 4; ++i) { // row i: t[i..i+4] += a_i * b  (t[i+4] is created by this row)
        const uint64_t x = a.v[i];
        asm("{\n\t"
            "mad.lo.cc.u64 %0, %5, %6, %0;\n\t"
            "madc.lo.cc.u64 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u64 %2, %5, %8, %2;\n\t"
            "madc.lo.cc.u64 %3, %5, %9, %3;\n\t"
            "addc.u64 %4, 0, 0;\n\t"
            "mad.hi.cc.u64 %1, %5, %6, %1;\n\t"
            "madc.hi.cc.u64 %2, %5, %7, %2;\n\t"
            "madc.hi.cc.u64 %3, %5, %8, %3;\n\t"
            "madc.hi.u64 %4, %5, %9, %4;\n\t"
            "}"
            : "+l"(t[i]), "+l"(t[i + 1]), "+l"(t[i + 2]), "+l"(t[i + 3]), "=&l"(t[i + 4])
            : "l"(x), "l"(b0), "l"(b1), "l"(b2), "l"(b3));
    }
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) t[i + j] = mac(a.v[i], b.v[j], t[i + j], carry, carry);
        t[i + 4] = carry;
    }
#endif
}
__device__ __forceinline__ u256 from_be(const uint8_t* b)
{
    u256 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint64_t w = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) w = (w << 8) | b[8 * (3 - i) + k];
        r.v[i] = w;
    }
    return r;
}
__device__ __forceinline__ void to_be(uint8_t* b, const u256& a)
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) b[8 * (3 - i) + k] = (uint8_t)(a.v[i] >> (8 * (7 - k)));
}

// ---- field F_p, p = 2^256 - C ----
constexpr uint64_t FP_C = 0x1000003D1ull;
__device__ __forceinline__ u256 fp_p() { return u256{{0xFFFFFFFEFFFFFC2Full, ~0ull, ~0ull, ~0ull}}; }

__device__ __forceinline__ void fp_norm(u256& r) // r < 2^256 -> r mod p
{
    if (r.v[3] == ~0ull && r.v[2] == ~0ull && r.v[1] == ~0ull && r.v[0] >= 0xFFFFFFFEFFFFFC2Full) {
        r.v[0] -= 0xFFFFFFFEFFFFFC2Full; r.v[1] = 0; r.v[2] = 0; r.v[3] = 0;
    }
}
__device__ __forceinline__ u256 fp_add(const u256& a, const u256& b)
{
    u256 r;
    const uint64_t c = add_raw(r, a, b);
    // r + 2^256 == r + C (mod p); cannot carry again: a, b < p.  Branch-free: add c * C (c is 0 or 1)
    const u256 k{{c ? FP_C : 0, 0, 0, 0}};
    add_raw(r, r, k);
    fp_norm(r);
    return r;
}
__device__ __forceinline__ u256 fp_sub(const u256& a, const u256& b)
{
    u256 r;
    const uint64_t bw = sub_raw(r, a, b);
    // went below zero: add p == subtract C (mod 2^256)
    const u256 k{{bw ? FP_C : 0, 0, 0, 0}};
    sub_raw(r, r, k);
    return r;
}
__device__ __forceinline__ u256 fp_neg(const u256& a) { return is_zero(a) ? a : fp_sub(u256{{0, 0, 0, 0}}, a); }
__device__ __forceinline__ u256 fp_reduce(const uint64_t (&t)[8])
{
#if defined(__CUDA_ARCH__)
    // lo + hi * C (C < 2^33): five limbs r0..r4 with r4 < 2^34, then r4 * C folded in once more; k = the last carry
    uint64_t r0, r1, r2, r3, r4, k;
    u256 o;
    asm("{\n\t"
        "mad.lo.cc.u64 %0, %9, %13, %5;\n\t"
        "madc.lo.cc.u64 %1, %10, %13, %6;\n\t"
        "madc.lo.cc.u64 %2, %11, %13, %7;\n\t"
        "madc.lo.cc.u64 %3, %12, %13, %8;\n\t"
        "addc.u64 %4, 0, 0;\n\t"
        "mad.hi.cc.u64 %1, %9, %13, %1;\n\t"
        "madc.hi.cc.u64 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u64 %3, %11, %13, %3;\n\t"
        "madc.hi.u64 %4, %12, %13, %4;\n\t"
        "}"
        : "=&l"(r0), "=&l"(r1), "=&l"(r2), "=&l"(r3), "=&l"(r4)
        : "l"(t[0]), "l"(t[1]), "l"(t[2]), "l"(t[3]), "l"(t[4]), "l"(t[5]), "l"(t[6]), "l"(t[7]), "l"(FP_C));
    asm("{\n\t"
        ".reg .u64 hi;\n\t"
        "mul.hi.u64 hi, %9, %10;\n\t"
        "mad.lo.cc.u64 %0, %9, %10, %5;\n\t"
        "addc.cc.u64 %1, %6, hi;\n\t"
        "addc.cc.u64 %2, %7, 0;\n\t"
        "addc.cc.u64 %3, %8, 0;\n\t"
        "addc.u64 %4, 0, 0;\n\t"
        "}"
        : "=&l"(o.v[0]), "=&l"(o.v[1]), "=&l"(o.v[2]), "=&l"(o.v[3]), "=&l"(k)
        : "l"(r0), "l"(r1), "l"(r2), "l"(r3), "l"(r4), "l"(FP_C));
    const u256 kc{{k ? FP_C : 0, 0, 0, 0}}; // wrapped once more: the value left is tiny, adding C cannot carry
    add_raw(o, o, kc);
    fp_norm(o);
    return o;
#else
    // lo + hi * C: hi * C is 4 limbs x 33 bits
    uint64_t r[5], carry = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = mac(t[4 + i], FP_C, t[i], carry, carry);
    r[4] = carry; // < 2^34
    // fold limb 4: r[4] * C < 2^67
    uint64_t hi;
    uint64_t lo = mac(r[4], FP_C, r[0], 0, hi);
    u256 o;
    uint64_t c = 0;
    o.v[0] = lo;
    o.v[1] = adc(r[1], hi, c);
    o.v[2] = adc(r[2], 0, c);
    o.v[3] = adc(r[3], 0, c);
    if (c) { // wrapped once more: the value left is tiny, adding C cannot carry
        uint64_t k = 0;
        o.v[0] = adc(o.v[0], FP_C, k); o.v[1] = adc(o.v[1], 0, k); o.v[2] = adc(o.v[2], 0, k); o.v[3] = adc(o.v[3], 0, k);
    }
    fp_norm(o);
    return o;
#endif
}
__device__ __forceinline__ u256 fp_mul(const u256& a, const u256& b)
{
    uint64_t t[8];
    mul_wide(t, a, b);
    return fp_reduce(t);
}
__device__ __forceinline__ u256 fp_sqr(const u256& a) { return fp_mul(a, a); }
__device__ __forceinline__ u256 fp_dbl(const u256& a) { return fp_add(a, a); }
// a^(2^k) (k >= 1)
__device__ __forceinline__ u256 fp_sqr_n(u256 a, int k)
{
#pragma unroll 1
    for (int i = 0; i < k; ++i) a = fp_sqr(a);
    return a;
}
// Both exponents p - 2 and (p + 1) / 4 start with a run of 223 ones, a zero, then 22 ones: build a^(2^223 - 1) once with the
// classic ladder of runs (x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223: 11 multiplications) instead of one
// multiplication per exponent bit -- 255 squarings + 15 multiplications for the inverse, 253 + 13 for the square root.
struct FpRuns { u256 x2, x22, x223; };
__device__ __noinline__ FpRuns fp_runs(const u256& a)
{
    FpRuns r;
    r.x2 = fp_mul(fp_sqr(a), a);
    const u256 x3 = fp_mul(fp_sqr(r.x2), a);
    const u256 x6 = fp_mul(fp_sqr_n(x3, 3), x3);
    const u256 x9 = fp_mul(fp_sqr_n(x6, 3), x3);
    const u256 x11 = fp_mul(fp_sqr_n(x9, 2), r.x2);
    r.x22 = fp_mul(fp_sqr_n(x11, 11), x11);
    const u256 x44 = fp_mul(fp_sqr_n(r.x22, 22), r.x22);
    const u256 x88 = fp_mul(fp_sqr_n(x44, 44), x44);
    const u256 x176 = fp_mul(fp_sqr_n(x88, 88), x88);
    const u256 x220 = fp_mul(fp_sqr_n(x176, 44), x44);
    r.x223 = fp_mul(fp_sqr_n(x220, 3), x3);
    return r;
}
// a^(p-2): p - 2 = 1^223 0 1^22 0000 1 0 1 1 0 1 (binary)
__device__ __forceinline__ u256 fp_inv(const u256& a)
{
    const FpRuns r = fp_runs(a);
    u256 t = fp_mul(fp_sqr_n(r.x223, 23), r.x22);
    t = fp_mul(fp_sqr_n(t, 5), a);
    t = fp_mul(fp_sqr_n(t, 3), r.x2);
    return fp_mul(fp_sqr_n(t, 2), a);
}
// a^((p+1)/4): (p + 1) / 4 = 1^223 0 1^22 0000 1 1 00 (binary)
__device__ __forceinline__ u256 fp_sqrt_candidate(const u256& a)
{
    const FpRuns r = fp_runs(a);
    u256 t = fp_mul(fp_sqr_n(r.x223, 23), r.x22);
    t = fp_mul(fp_sqr_n(t, 6), r.x2);
    return fp_sqr_n(t, 2);
}

// ---- scalars mod n, n = 2^256 - K, K = 0x1_4551231950B75FC4_402DA1732FC9BEBF ----
__device__ __forceinline__ u256 sc_n() { return u256{{0xBFD25E8CD0364141ull, 0xBAAEDCE6AF48A03Bull, 0xFFFFFFFFFFFFFFFEull, ~0ull}}; }
__device__ __noinline__ u256 sc_mul(const u256& a, const u256& b)
{
    const uint64_t K[3] = {0x402DA1732FC9BEBFull, 0x4551231950B75FC4ull, 1ull};
    uint64_t t[8];
    mul_wide(t, a, b);
    // fold the high half down with K until nothing is left above 2^256 (at most four rounds: 512 -> 385 -> 259 -> 257 -> 256 bits)
#pragma unroll 1
    while (t[4] | t[5] | t[6] | t[7]) {
        uint64_t u[8] = {t[0], t[1], t[2], t[3], 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint64_t carry = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) u[i + j] = mac(t[4 + i], K[j], u[i + j], carry, carry);
#pragma unroll
            for (int q = i + 3; q < 8; ++q) u[q] = adc(u[q], 0, carry);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = u[i];
    }
    u256 r{{t[0], t[1], t[2], t[3]}};
    const u256 n = sc_n();
    if (geq(r, n)) sub_raw(r, r, n);
    return r;
}
__device__ __forceinline__ u256 sc_neg(const u256& a)
{
    if (is_zero(a)) return a;
    u256 r;
    sub_raw(r, sc_n(), a);
    return r;
}
__device__ __forceinline__ u256 sc_sqr_n(u256 a, int k)
{
#pragma unroll 1
    for (int i = 0; i < k; ++i) a = sc_mul(a, a);
    return a;
}
// a^(n-2), n prime.  The top half of n - 2 is 127 ones and a zero: a^(2^127 - 1) comes from a ladder of runs (10
// multiplications); the bottom 128 bits are taken one by one (64 of them set): 255 squarings + 74 multiplications.
__device__ __noinline__ u256 sc_inv(const u256& a)
{
    const u256 x2 = sc_mul(sc_mul(a, a), a);
    const u256 x3 = sc_mul(sc_mul(x2, x2), a);
    const u256 x6 = sc_mul(sc_sqr_n(x3, 3), x3);
    const u256 x12 = sc_mul(sc_sqr_n(x6, 6), x6);
    const u256 x24 = sc_mul(sc_sqr_n(x12, 12), x12);
    const u256 x48 = sc_mul(sc_sqr_n(x24, 24), x24);
    const u256 x96 = sc_mul(sc_sqr_n(x48, 48), x48);
    const u256 x120 = sc_mul(sc_sqr_n(x96, 24), x24);
    const u256 x126 = sc_mul(sc_sqr_n(x120, 6), x6);
    u256 acc = sc_mul(sc_mul(x126, x126), a); // x127
    acc = sc_mul(acc, acc);                  // the zero bit
    const uint64_t e[2] = {0xBFD25E8CD036413Full, 0xBAAEDCE6AF48A03Bull};
#pragma unroll 1
    for (int i = 127; i >= 0; --i) {
        acc = sc_mul(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = sc_mul(acc, a);
    }
    return acc;
}

// ---- curve y^2 = x^3 + 7 ----
struct Affine { u256 x, y; bool inf; };
struct Jac { u256 x, y, z; }; // z == 0: point at infinity

__device__ __forceinline__ Affine gen()
{
    return Affine{u256{{0x59F2815B16F81798ull, 0x029BFCDB2DCE28D9ull, 0x55A06295CE870B07ull, 0x79BE667EF9DCBBACull}},
                  u256{{0x9C47D08FFB10D4B8ull, 0xFD17B448A6855419ull, 0x5DA4FBFC0E1108A8ull, 0x483ADA7726A3C465ull}}, false};
}
__device__ __forceinline__ Jac jac_double(const Jac& p)
{
    if (is_zero(p.z) || is_zero(p.y)) return Jac{u256{{0, 0, 0, 0}}, u256{{1, 0, 0, 0}}, u256{{0, 0, 0, 0}}};
    // dbl-2009-l (a = 0): A = X^2, B = Y^2, C = B^2, D = 2((X+B)^2 - A - C), E = 3A, X3 = E^2 - 2D, Y3 = E(D - X3) - 8C, Z3 = 2YZ
    const u256 A = fp_sqr(p.x), B = fp_sqr(p.y), C = fp_sqr(B);
    u256 D = fp_sub(fp_sub(fp_sqr(fp_add(p.x, B)), A), C);
    D = fp_dbl(D);
    const u256 E = fp_add(fp_dbl(A), A);
    Jac r;
    r.x = fp_sub(fp_sqr(E), fp_dbl(D));
    r.y = fp_sub(fp_mul(E, fp_sub(D, r.x)), fp_dbl(fp_dbl(fp_dbl(C))));
    r.z = fp_dbl(fp_mul(p.y, p.z));
    return r;
}
// out-of-line copy for the places that are not the hot loop (table set-up, the same-point corner case of an addition): the
// loop body holds exactly ONE inlined doubling and ONE inlined addition, so it stays inside the instruction cache and the
// accumulator never leaves the registers (as a noinline call the Jacobian point went through local memory on every call)
__device__ __noinline__ Jac jac_double_once(const Jac& p) { return jac_double(p); }
// Jacobian + affine (q not at infinity), every corner case handled
__device__ __forceinline__ Jac jac_add_affine(const Jac& p, const Affine& q)
{
    if (is_zero(p.z)) return Jac{q.x, q.y, u256{{1, 0, 0, 0}}};
    const u256 z2 = fp_sqr(p.z);
    const u256 u2 = fp_mul(q.x, z2), s2 = fp_mul(fp_mul(q.y, z2), p.z);
    const u256 h = fp_sub(u2, p.x), r = fp_sub(s2, p.y);
    if (is_zero(h)) {
        if (is_zero(r)) return jac_double_once(p);                                      // the same point
        return Jac{u256{{0, 0, 0, 0}}, u256{{1, 0, 0, 0}}, u256{{0, 0, 0, 0}}};         // opposite points
    }
    const u256 h2 = fp_sqr(h), h3 = fp_mul(h2, h), v = fp_mul(p.x, h2);
    Jac o;
    o.x = fp_sub(fp_sub(fp_sqr(r), h3), fp_dbl(v));
    o.y = fp_sub(fp_mul(r, fp_sub(v, o.x)), fp_mul(p.y, h3));
    o.z = fp_mul(p.z, h);
    return o;
}
__device__ __noinline__ Jac jac_add_affine_once(const Jac& p, const Affine& q) { return jac_add_affine(p, q); }
__device__ __forceinline__ Affine to_affine(const Jac& p)
{
    if (is_zero(p.z)) return Affine{u256{{0, 0, 0, 0}}, u256{{0, 0, 0, 0}}, true};
    const u256 zi = fp_inv(p.z), zi2 = fp_sqr(zi);
    return Affine{fp_mul(p.x, zi2), fp_mul(p.y, fp_mul(zi2, zi)), false};
}
__device__ __forceinline__ uint32_t bit_of(const u256& k, int i) { return (uint32_t)(k.v[i >> 6] >> (i & 63)) & 1u; }

// hash32 = message hash, sig65 = r || s || recid.  true: pub64 = X || Y (big endian, without the 0x04 prefix).
__device__ __forceinline__ bool ecrecover(const uint8_t* hash32, const uint8_t* sig65, uint8_t* pub64)
{
    const u256 r = from_be(sig65), s = from_be(sig65 + 32);
    u256 z = from_be(hash32);
    const uint32_t recid = sig65[64];
    const u256 n = sc_n(), p = fp_p();
    if (recid > 3 || is_zero(r) || is_zero(s) || geq(r, n) || geq(s, n)) return false;
    u256 x = r;
    if (recid & 2) {
        if (add_raw(x, r, n) || geq(x, p)) return false;
    }
    const u256 y2 = fp_add(fp_mul(fp_sqr(x), x), u256{{7, 0, 0, 0}});
    u256 y = fp_sqrt_candidate(y2);
    const u256 chk = fp_sqr(y);
    if (chk.v[0] != y2.v[0] || chk.v[1] != y2.v[1] || chk.v[2] != y2.v[2] || chk.v[3] != y2.v[3]) return false; // not on the curve
    if ((y.v[0] & 1) != (recid & 1)) y = fp_neg(y);
    if (geq(z, n)) sub_raw(z, z, n);
    const u256 rinv = sc_inv(r);
    const u256 u1 = sc_neg(sc_mul(z, rinv)), u2 = sc_mul(s, rinv);

    // table: 1 = G, 2 = R, 3 = G + R
    const Affine G = gen(), R{x, y, false};
    const Affine GR = to_affine(jac_add_affine_once(Jac{G.x, G.y, u256{{1, 0, 0, 0}}}, R));
    Jac acc{u256{{0, 0, 0, 0}}, u256{{1, 0, 0, 0}}, u256{{0, 0, 0, 0}}};
#pragma unroll 1
    for (int i = 255; i >= 0; --i) {
        acc = jac_double(acc);
        const uint32_t sel = bit_of(u1, i) | (bit_of(u2, i) << 1);
        if (sel && !(sel == 3 && GR.inf)) { // one addition site: the table entry is selected word by word
            Affine t;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                t.x.v[w] = sel == 1 ? G.x.v[w] : (sel == 2 ? R.x.v[w] : GR.x.v[w]);
                t.y.v[w] = sel == 1 ? G.y.v[w] : (sel == 2 ? R.y.v[w] : GR.y.v[w]);
            }
            t.inf = false;
            acc = jac_add_affine(acc, t);
        }
    }
    const Affine q = to_affine(acc);
    if (q.inf) return false;
    to_be(pub64, q.x);
    to_be(pub64 + 32, q.y);
    return true;
}

} // namespace secp
} // namespace phant

/* rlp.h -- TEST INFRASTRUCTURE (see oracle.h).  Textbook RLP encode/decode helpers.
 *
 * phant encodes through the un-vendored zig-rlp v0.1.1-beta7 (reference build.zig.zon:5-8); its
 * output is pinned by the seven roots of src/mpt/mpt.zig:326-385 and the fixture roots, all of which
 * standard RLP reproduces.  Same rules as evmone/test/state/rlp.hpp:21-67.
 */
#ifndef PHANT_ORACLE_RLP_H
#define PHANT_ORACLE_RLP_H
#include <stdint.h>
#include <string.h>

static inline unsigned rlp_be_len(uint64_t v)
{
    unsigned n = 0;
    while (v) { ++n; v >>= 8; }
    return n;
}
/* encoded size of a byte string */
static inline uint64_t rlp_str_size(const uint8_t* s, uint64_t len)
{
    if (len == 1 && s[0] < 0x80) return 1;
    if (len <= 55) return 1 + len;
    return 1 + rlp_be_len(len) + len;
}
static inline uint64_t rlp_put_hdr(uint8_t* out, uint64_t len, uint8_t short_base, uint8_t long_base)
{
    if (len <= 55) { out[0] = (uint8_t)(short_base + len); return 1; }
    unsigned n = rlp_be_len(len);
    out[0] = (uint8_t)(long_base + n);
    for (unsigned i = 0; i < n; ++i) out[1 + i] = (uint8_t)(len >> (8 * (n - 1 - i)));
    return 1 + n;
}
static inline uint64_t rlp_put_str(uint8_t* out, const uint8_t* s, uint64_t len)
{
    if (len == 1 && s[0] < 0x80) { out[0] = s[0]; return 1; }
    uint64_t h = rlp_put_hdr(out, len, 0x80, 0xb7);
    if (len) memcpy(out + h, s, len);
    return h + len;
}
static inline uint64_t rlp_list_hdr_size(uint64_t payload) { return payload <= 55 ? 1 : 1 + rlp_be_len(payload); }
static inline uint64_t rlp_put_list_hdr(uint8_t* out, uint64_t payload) { return rlp_put_hdr(out, payload, 0xc0, 0xf7); }

/* Strict item decode at p (bytes available: avail).  Returns total encoded size of the item or 0 when
 * malformed / out of bounds / non-canonical.  *is_list, *pay_off (offset of payload from p), *pay_len. */
static inline uint64_t rlp_item(const uint8_t* p, uint64_t avail, int* is_list, uint64_t* pay_off, uint64_t* pay_len)
{
    if (avail == 0) return 0;
    uint8_t b = p[0];
    if (b < 0x80) { *is_list = 0; *pay_off = 0; *pay_len = 1; return 1; }
    uint8_t base_short = b < 0xc0 ? 0x80 : 0xc0;
    uint8_t base_long = b < 0xc0 ? 0xb7 : 0xf7;
    *is_list = b >= 0xc0;
    if (b <= base_long) {
        uint64_t len = (uint64_t)(b - base_short);
        if (1 + len > avail) return 0;
        if (!*is_list && len == 1 && p[1] < 0x80) return 0; /* non-canonical single byte */
        *pay_off = 1; *pay_len = len;
        return 1 + len;
    }
    unsigned n = (unsigned)(b - base_long);
    if (n > 4 || 1 + (uint64_t)n > avail) return 0; /* lengths >= 2^32 not supported */
    if (p[1] == 0) return 0;                        /* leading zero in length */
    uint64_t len = 0;
    for (unsigned i = 0; i < n; ++i) len = (len << 8) | p[1 + i];
    if (len <= 55) return 0; /* should have used the short form */
    if (1 + n + len > avail) return 0;
    *pay_off = 1 + n; *pay_len = len;
    return 1 + n + len;
}
#endif

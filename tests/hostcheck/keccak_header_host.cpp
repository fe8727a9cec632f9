// Test-only: compiles phant_b200/csrc/keccak_f1600.cuh AS HOST CODE (CUDA qualifiers and the two intrinsics it uses are
// defined away below) so that the permutation's index tables -- including the pruned digest-only last round -- can be
// checked on a machine without a GPU.  Not part of the product: nothing links this.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define __constant__ static const
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) { n &= 31; return n ? (hi << n) | (lo >> (32 - n)) : hi; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t n) { n &= 31; return n ? (lo >> n) | (hi << (32 - n)) : lo; }
#include "../../phant_b200/csrc/keccak_f1600.cuh"

using namespace phant;

// stdin: hex message per line ("-" = empty); stdout: digest via the full permutation, digest via the digest-only one,
// and whether lanes 0..3 of a random state agree between the two permutations
int main()
{
    char line[1 << 16];
    while (fgets(line, sizeof line, stdin)) {
        size_t hl = strlen(line);
        while (hl && (line[hl - 1] == '\n' || line[hl - 1] == '\r')) line[--hl] = 0;
        static uint8_t msg[1 << 15];
        size_t n = 0;
        if (strcmp(line, "-") != 0)
            for (; 2 * n + 1 < hl; ++n) { unsigned v; sscanf(line + 2 * n, "%2x", &v); msg[n] = (uint8_t)v; }
        // sponge by hand over the two permutation variants
        for (int variant = 0; variant < 2; ++variant) {
            uint64_t st[25] = {0};
            size_t pos = 0;
            uint8_t block[136];
            for (;;) {
                const size_t take = n - pos < 136 ? n - pos : 136;
                memset(block, 0, sizeof block);
                memcpy(block, msg + pos, take);
                const bool last = take < 136;
                if (last) { block[take] ^= 0x01; block[135] ^= 0x80; }
                for (int k = 0; k < 17; ++k) { uint64_t w; memcpy(&w, block + 8 * k, 8); st[k] ^= w; }
                if (last && variant == 1) keccak_f1600<2, true>(st); else keccak_f1600<2, false>(st);
                pos += take;
                if (last) break;
            }
            for (int i = 0; i < 32; ++i) printf("%02x", (unsigned)((st[i / 8] >> (8 * (i % 8))) & 0xff));
            printf(variant ? "\n" : " ");
        }
    }
    return 0;
}

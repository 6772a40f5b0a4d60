/* oracle/philox.h - Philox4x32-10 (Salmon et al., SC'11) and the rlsim-b200
 * counter-based stream layout ("--rng=philox" normative spec).
 *
 * TEST INFRASTRUCTURE ONLY: CPU oracle (see oracle/README.md).  The product has
 * its own independent implementation in rlsim_b200/csrc/philox.cuh; parity
 * tests require the two to agree bit for bit.
 */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <stdint.h>

typedef struct { uint32_t v[4]; } px_block;

static inline px_block px_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                        uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    px_block b = {{c0, c1, c2, c3}};
    return b;
}

/* stream tags */
enum { PX_TAG_TARGET = 1, PX_TAG_FRAG = 2, PX_TAG_PCR = 3, PX_TAG_SELECT = 4, PX_TAG_STRAND = 5 };
/* sub-streams of a molecule's fragmentation stream (top byte of c3) */
enum { PX_SUB_HEADER = 0, PX_SUB_BREAKS = 1, PX_SUB_INTERVAL = 2 };

typedef struct { uint32_t k0, k1; } px_key;

/* Stream key derivation: key(tag) = first two words of
 * Philox(ctr = (tag, 'RLSI', 'MB20', 0), key = (seed_lo, seed_hi)). */
static inline px_key px_derive_key(int64_t seed, uint32_t tag) {
    px_block b = px_philox4x32_10(tag, 0x524c5349u, 0x4d423230u, 0u,
                                  (uint32_t)((uint64_t)seed & 0xffffffffu), (uint32_t)((uint64_t)seed >> 32));
    px_key k = {b.v[0], b.v[1]};
    return k;
}

/* A stream = key + the three upper counter words; c0 is the block index. */
typedef struct {
    px_key key;
    uint32_t c1, c2, c3;
    /* sequential cursor (in 64-bit halves) and one-block cache */
    uint64_t cur;
    uint32_t curb; /* sequential cursor in whole blocks (PCR stream) */
    uint32_t cached_block;
    int have;
    px_block blk;
} px_stream;

static inline void px_stream_init(px_stream *s, px_key key, uint32_t c1, uint32_t c2, uint32_t c3) {
    s->key = key; s->c1 = c1; s->c2 = c2; s->c3 = c3; s->cur = 0; s->curb = 0; s->have = 0; s->cached_block = 0;
}

static inline px_block px_stream_block(const px_stream *s, uint32_t b) {
    return px_philox4x32_10(b, s->c1, s->c2, s->c3, s->key.k0, s->key.k1);
}

/* 64-bit draw number d (random access): block d>>1, half d&1. */
static inline uint64_t px_draw64(px_stream *s, uint64_t d) {
    uint32_t b = (uint32_t)(d >> 1);
    if (!s->have || s->cached_block != b) { s->blk = px_stream_block(s, b); s->cached_block = b; s->have = 1; }
    return (d & 1) ? ((uint64_t)s->blk.v[2] | ((uint64_t)s->blk.v[3] << 32))
                   : ((uint64_t)s->blk.v[0] | ((uint64_t)s->blk.v[1] << 32));
}
static inline uint64_t px_next64(px_stream *s) { return px_draw64(s, s->cur++); }
static inline px_block px_next_block(px_stream *s) { return px_stream_block(s, s->curb++); }
static inline uint64_t px_blk_lo(px_block b) { return (uint64_t)b.v[0] | ((uint64_t)b.v[1] << 32); }
static inline uint64_t px_blk_hi(px_block b) { return (uint64_t)b.v[2] | ((uint64_t)b.v[3] << 32); }

/* U in [0,1): top 53 bits. */
static inline double px_u01(uint64_t x) { return (double)(x >> 11) * 1.1102230246251565404e-16; }
/* U in (0,1): top 52 bits + 1/2, exactly representable. */
static inline double px_u01_open(uint64_t x) { return ((double)(x >> 12) + 0.5) * 2.2204460492503130808e-16; }

/* Unbiased-to-2^-64 bounded integer from one whole block (128 random bits):
 * floor(X * n / 2^128), X = x0 * 2^64 + x1 with x0 = v0|v1<<32, x1 = v2|v3<<32. */
static inline uint64_t px_below_block(px_block b, uint64_t n) {
    uint64_t x0 = (uint64_t)b.v[0] | ((uint64_t)b.v[1] << 32);
    uint64_t x1 = (uint64_t)b.v[2] | ((uint64_t)b.v[3] << 32);
    unsigned __int128 lo = (unsigned __int128)x1 * n;
    unsigned __int128 hi = (unsigned __int128)x0 * n + (uint64_t)(lo >> 64);
    return (uint64_t)(hi >> 64);
}
static inline uint64_t px_below(const px_stream *s, uint32_t block, uint64_t n) {
    return px_below_block(px_stream_block(s, block), n);
}
#endif

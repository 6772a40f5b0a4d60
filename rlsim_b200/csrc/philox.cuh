// Philox4x32-10 counter-based streams of the "--rng=philox" spec.
// Independent implementation of the same spec the CPU oracle restates; parity tests require bit equality.
#pragma once
#include <stdint.h>

namespace rl {

enum : uint32_t { TAG_TARGET = 1, TAG_FRAG = 2, TAG_PCR = 3, TAG_SELECT = 4, TAG_STRAND = 5 };
enum : uint32_t { SUB_HEADER = 0, SUB_BREAKS = 1, SUB_INTERVAL = 2, SUB_JUMPLEN = 3 };

struct PxKey { uint32_t k0, k1; };
struct PxBlock { uint32_t v0, v1, v2, v3; };

__host__ __device__ __forceinline__ void px_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

__host__ __device__ __forceinline__ PxBlock philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        px_mulhilo(0xD2511F53u, c0, hi0, lo0);
        px_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return PxBlock{c0, c1, c2, c3};
}

// key(tag) = first two words of Philox(ctr=(tag,'RLSI','MB20',0), key=(seed_lo, seed_hi))
__host__ __device__ inline PxKey px_derive_key(int64_t seed, uint32_t tag) {
    PxBlock b = philox4x32_10(tag, 0x524c5349u, 0x4d423230u, 0u, (uint32_t)((uint64_t)seed & 0xffffffffu),
                              (uint32_t)((uint64_t)seed >> 32));
    return PxKey{b.v0, b.v1};
}

__host__ __device__ __forceinline__ uint64_t px_lo64(const PxBlock &b) { return (uint64_t)b.v0 | ((uint64_t)b.v1 << 32); }
__host__ __device__ __forceinline__ uint64_t px_hi64(const PxBlock &b) { return (uint64_t)b.v2 | ((uint64_t)b.v3 << 32); }

// U in [0,1): top 53 bits;  U in (0,1): top 52 bits + 1/2
__host__ __device__ __forceinline__ double px_u01(uint64_t x) { return (double)(x >> 11) * 1.1102230246251565404e-16; }
__host__ __device__ __forceinline__ double px_u01_open(uint64_t x) {
    return ((double)(x >> 12) + 0.5) * 2.2204460492503130808e-16;
}

// floor(X * n / 2^128), X = lo64 * 2^64 + hi64 : bounded integer from one whole block
__device__ __forceinline__ uint64_t px_below(const PxBlock &b, uint64_t n) {
    uint64_t x0 = px_lo64(b), x1 = px_hi64(b);
    uint64_t lo_hi = __umul64hi(x1, n);       // floor(x1*n / 2^64)
    uint64_t p_lo = x0 * n, p_hi = __umul64hi(x0, n);
    uint64_t s = p_lo + lo_hi;
    return p_hi + (s < p_lo ? 1ull : 0ull);
}

// A stream: key + upper three counter words; c0 is the block index.
struct PxStream {
    PxKey key;
    uint32_t c1, c2, c3;
    __device__ __forceinline__ PxBlock block(uint32_t b) const { return philox4x32_10(b, c1, c2, c3, key.k0, key.k1); }
    __device__ __forceinline__ uint64_t below(uint32_t b, uint64_t n) const { return px_below(block(b), n); }
};

// Sequential cursor over the 64-bit halves of a stream.
struct PxCursor {
    PxStream s;
    uint32_t cur;
    PxBlock blk;
    __device__ __forceinline__ PxCursor(const PxStream &st) : s(st), cur(0) {}
    __device__ __forceinline__ uint64_t next64() {
        uint32_t d = cur++;
        if ((d & 1u) == 0u) { blk = s.block(d >> 1); return px_lo64(blk); }
        return px_hi64(blk);
    }
    __device__ __forceinline__ double u01() { return px_u01(next64()); }
    __device__ __forceinline__ double u01_open() { return px_u01_open(next64()); }
};

// Sequential cursor whose first three blocks (six draws) are generated up front, in convergent code: the header
// draws of a molecule (poly-A, component, Poisson) are consumed inside data-dependent loops where generating a
// block on demand runs with a fraction of the warp.  Draws beyond the sixth fall back to on-demand blocks.
struct PxCursor3 {
    PxStream s;
    uint32_t cur;
    uint64_t d0, d1, d2, d3, d4, d5;  // scalars, selected with predicated moves (no local-memory array)
    PxBlock blk;
    __device__ __forceinline__ PxCursor3(const PxStream &st) : s(st), cur(0) {
        PxBlock a = s.block(0), b = s.block(1), c = s.block(2);
        d0 = px_lo64(a); d1 = px_hi64(a); d2 = px_lo64(b); d3 = px_hi64(b); d4 = px_lo64(c); d5 = px_hi64(c);
    }
    __device__ __forceinline__ uint64_t next64() {
        uint32_t d = cur++;
        if (d >= 6) {
            if ((d & 1u) == 0u) blk = s.block(d >> 1);
            return (d & 1u) ? px_hi64(blk) : px_lo64(blk);
        }
        uint64_t lo = d < 2 ? d0 : (d < 4 ? d2 : d4), hi = d < 2 ? d1 : (d < 4 ? d3 : d5);
        return (d & 1u) ? hi : lo;
    }
    __device__ __forceinline__ double u01() { return px_u01(next64()); }
    __device__ __forceinline__ double u01_open() { return px_u01_open(next64()); }
};

// random-access 64-bit draw d of a stream
__device__ __forceinline__ uint64_t px_draw64(const PxStream &s, uint64_t d) {
    PxBlock b = s.block((uint32_t)(d >> 1));
    return (d & 1ull) ? px_hi64(b) : px_lo64(b);
}

}  // namespace rl

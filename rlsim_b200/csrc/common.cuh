// Shared device-side types of librlsim_gpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rlsim_gpu.h"
#include "philox.cuh"

namespace rl {

// Model parameters as the kernels see them (passed by value as a kernel argument).
struct DevModel {
    int32_t k;             // primer length
    int32_t method;        // RLSIM_AFTER_PRIM ...
    int32_t frag_param;
    int32_t cycles;
    double T;              // priming temperature
    double loss;           // fragment loss probability
    double fixed_eff;
    double strand_bias;
    int32_t gc_mode;
    int32_t n_target;      // 0 => raw target
    int32_t n_polya;
    int32_t n_raw;
    double gc_shape, gc_min, gc_max;
    double len_shape, len_min, len_max;
    rlsim_mixcomp target[RLSIM_MAX_COMPS];
    rlsim_mixcomp polya[RLSIM_MAX_COMPS];
    double raw_location;   // RawTarget pseudo component mean (raw_target.go:76-90)
    uint32_t low, high;    // Targeter.GetLow/GetHigh
    uint32_t polya_max;
    uint32_t first_index;  // global index of this shard's first transcript
    PxKey key_target, key_frag, key_pcr, key_select, key_strand;
    // device tables
    const double *gc_raw;      // [101]
    const double *len_eff;     // [high-low+1]
    const uint32_t *raw_len;   // raw target lengths
    const double *raw_prob;    // raw target probabilities (file order)
};

// one 32-byte record per 64-entry block of forward prefix sums (layout described below, at RProfile)
struct __align__(16) Rec32 { uint64_t coarse; uint32_t half; uint32_t codes[5]; };
__host__ __device__ __forceinline__ uint64_t rec_base(uint64_t off, uint32_t t) { return (off >> 6) + 2ull * t; }

// Transcript store, SoA in HBM.
struct DevTranscripts {
    uint32_t n;
    const uint8_t *bases;     // plus strands with polyAmax 'A's appended, each starting at off[t] (multiple of 32)
    const uint32_t *packed;   // 2-bit codes, 16 bases per word, word index = (off[t] + p) >> 4
    const uint32_t *nmask;    // 1 bit per base (non-ACGT), 32 bases per word
    const uint64_t *off;      // [n+1] padded base offsets
    const uint32_t *len;      // [n] length including the maximal poly-A tail
    const uint64_t *expr;     // [n]
    const uint64_t *mol_off;  // [n+1] exclusive scan of expr
    // priming prefix sums, two levels: fine = u32 inclusive sums inside 64-entry blocks at off[t],
    // coarse = u64 sum of everything before each block at coarse_base(off[t], t)
    const uint32_t *Ff, *Fr;
    const uint64_t *Cf, *Cr;
    const uint32_t *dirty;    // per transcript: non-zero when a letter outside ACGT occurs (incl. the poly(A) tail: never)
    int sym;                  // quantised k-mer table is reverse-complement symmetric: clean transcripts carry no reverse
                              // arrays, their forward arrays hold one more entry (window len-k) and are read mirrored
    const Rec32 *Rf;          // record layout of the forward sums (transcripts with dirty[t] == 0), or null: two-level arrays for all
    const uint32_t *lut;      // quantised k-mer weights in global memory (the record layout recomputes weights from the codes)
    const uint2 *gc;          // per 32-base group {#GC before the group, GC bit mask of the group}; transcript t has
                              // padded_len/32 + 1 entries starting at off[t]/32 + t (see gc_before)
};

// Two-level prefix sums of the fixed-point priming weights w (< 2^26, so a 64-entry block sums to < 2^32):
//   fine[i]   = w[64*(i/64)] + ... + w[i]            (u32, at entry off[t] + i)
//   coarse[b] = w[0] + ... + w[64*b - 1]             (u64, at coarse_base(off[t], t) + b; coarse[nblk] = total)
// P(i) = sum of w[j], j < i  =  coarse[i >> 6] + (i & 63 ? fine[i - 1] : 0).
constexpr uint32_t PFX_BLOCK_SHIFT = 6, PFX_BLOCK = 64;
__host__ __device__ __forceinline__ uint64_t coarse_base(uint64_t off, uint32_t t) { return (off >> 5) + 2ull * t; }

struct Profile {  // one strand of one transcript
    const uint32_t *F;
    const uint64_t *C;
    __device__ __forceinline__ uint64_t P(uint32_t i) const {
        uint64_t c = C[i >> PFX_BLOCK_SHIFT];
        return (i & (PFX_BLOCK - 1)) ? c + F[i - 1] : c;
    }
};

// ---- Record layout of the forward prefix sums (transcripts without letters outside ACGT, symmetric k-mer table, k <= 6).
// One 32-byte record per 64-entry block - exactly one DRAM sector - holds everything a search needs from that block:
//   coarse : sum of the weights of all entries before the block
//   half   : sum of the block's first 32 entries
//   codes  : the 2-bit codes of the 80 bases from the block's first window on (64 windows of k <= 17 bases)
// P(i) = coarse + (i & 32 ? half : 0) + at most 31 weights recomputed from the codes and the k-mer LUT: 0.5 B per entry
// in HBM instead of 4.1, one sector per P() instead of two, and the block that holds the answer needs no further read.
// Record nblk (coarse only) closes a transcript; transcript t's records start at rec_base(off[t], t).

struct RProfile {
    const Rec32 *R;
    const uint32_t *lut;   // quantised weights by 2-bit code (shared memory in k_intervals, global elsewhere)
    uint32_t kmask;
    // sum of the weights of the entries [j0, j0 + n) of a block, j0 in {0, 32}, n <= 32; w0..w2 = the 48 bases from j0
    __device__ __forceinline__ uint32_t run(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t n) const {
        const uint64_t a = (uint64_t)w0 | ((uint64_t)w1 << 32), b = (uint64_t)w1 | ((uint64_t)w2 << 32);
        uint32_t acc = 0;
        const uint32_t na = n < 16u ? n : 16u;
        for (uint32_t j = 0; j < na; j++) acc += lut[(uint32_t)(a >> (2 * j)) & kmask];
        for (uint32_t j = 16; j < n; j++) acc += lut[(uint32_t)(b >> (2 * (j - 16))) & kmask];
        return acc;
    }
    __device__ __forceinline__ uint64_t P(uint32_t i) const {
        const uint4 *q = reinterpret_cast<const uint4 *>(R + (i >> PFX_BLOCK_SHIFT));
        const uint4 u = q[0];
        uint64_t c = (uint64_t)u.x | ((uint64_t)u.y << 32);
        const uint32_t r = i & (PFX_BLOCK - 1);
        if (r == 0) return c;
        const uint4 v = q[1];  // codes[1..4]
        const bool hi = r >= 32;
        if (hi) c += u.z;
        return c + run(hi ? v.y : u.w, hi ? v.z : v.x, hi ? v.w : v.y, r & 31u);
    }
};

// number of G/C among the first x bases of a transcript whose GC groups start at g
__device__ __forceinline__ uint32_t gc_before(const uint2 *g, uint32_t x) {
    uint2 e = g[x >> 5];
    return e.x + __popc(e.y & ((1u << (x & 31)) - 1u));
}

// bits of the length field of a fragment key (lengths are <= high < 2^24)
__host__ __device__ __forceinline__ uint32_t key_len_bits(uint32_t high) {
    uint32_t b = 1;
    while (b < 24 && (high >> b)) b++;
    return b;
}

__device__ __forceinline__ uint32_t upper_bound_u64(const uint64_t *a, uint32_t n, uint64_t x) {
    uint32_t lo = 0, hi = n;  // first i with a[i] > x
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (!(a[mid] > x)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace rl

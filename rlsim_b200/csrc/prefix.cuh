// Priming-affinity building blocks shared by k_prefix (prefix sums materialised in HBM, k_transcripts.cu) and the fused
// prefix + fragmentation kernel (prefix sums kept in shared memory, k_fragment.cu).
//   nnthermo.go:68-104   doublet tables
//   nnthermo.go:126-175  KmerAffinity
//   nnthermo.go:177-201  GetBindingProfile: entry i of the forward profile is the k-mer at plus[i, i+k); entry i of the
//                        reverse profile is the reverse complement of plus[j, j+k), j = L-k-i
#pragma once
#include "common.cuh"
#include "detmath.cuh"

namespace rl {

__device__ __forceinline__ int base_code(uint8_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

// nnthermo.go:68-104 doublet tables, index 4*a+b with A C G T = 0 1 2 3
__constant__ double NN_DH[16] = {-7.6, -8.4, -7.8, -7.2, -8.5, -8.0, -10.6, -7.8, -8.2, -9.8, -8.0, -8.4, -7.2, -8.2, -8.5, -7.6};
__constant__ double NN_DS[16] = {-21.3, -22.4, -21.0, -20.4, -22.7, -19.9, -27.2, -21.0, -22.2, -24.4, -19.9, -22.4, -21.3, -22.2, -22.7, -21.3};

// KmerAffinity over k bytes (any alphabet), nnthermo.go:126-175
static __device__ __noinline__ double kmer_affinity_bytes(const uint8_t *kmer, int kl, double T) {
    double dH = +0.2, dS = -5.7;
    for (int i = 0; i <= kl - 2; i++) {
        int a = base_code(kmer[i]), b = base_code(kmer[i + 1]);
        if (a >= 0 && b >= 0) { dH += NN_DH[4 * a + b]; dS += NN_DS[4 * a + b]; }
    }
    if (kmer[0] == 'A' || kmer[0] == 'T') { dH += 2.2; dS += 6.9; }
    if (kmer[kl - 1] == 'A' || kmer[kl - 1] == 'T') { dH += 2.2; dS += 6.9; }
    bool sym = true;
    for (int j = 0; j <= kl / 2; j++)
        if (kmer[j] != kmer[kl - j - 1]) { sym = false; break; }
    if (sym) dS += -1.4;
    double e = dH - (T * dS) / 1000.0;
    return det_exp(-e / (1.9872 * T));
}

// spec 4.3: w = clamp(floor(K * 2^26 / Kmax + 1/2), 1, 2^26 - 1).  The lower clamp keeps every window's total positive:
// SampleIndexFloat64 (random.go:369-401) accepts any positive weights, so a low -p or a run of N must not abort the run.
#define PFX_SCALE 67108864.0
__device__ __forceinline__ uint32_t quantize(double K, double scale) {
    double x = K * scale + 0.5;
    if (!(x < 67108863.0)) return 67108863u;
    if (x < 1.0) return 1u;
    return (uint32_t)x;
}

constexpr int PFX_ITEMS = 8;                     // profile entries per lane and sub-tile
constexpr int PFX_SUB = 32 * PFX_ITEMS;          // 256 entries per warp sub-tile = 4 blocks of 64

struct PfxWords { uint32_t p0, p1, p2, m0, m1; };

__device__ __forceinline__ uint32_t pfx_jb(int reverse, uint32_t L, int k, uint32_t proflen, uint32_t ib) {
    // lowest plus-strand window start among the (valid) entries [ib, ib+8) of this lane
    if (!reverse) return ib;
    uint32_t nvalid = min((uint32_t)PFX_ITEMS, proflen - ib);
    return L - k - (ib + nvalid - 1);
}
__device__ __forceinline__ PfxWords pfx_load(const DevTranscripts &tr, uint64_t o, uint32_t jb) {
    const uint32_t *pk = tr.packed + (o >> 4) + (jb >> 4);
    const uint32_t *mk = tr.nmask + (o >> 5) + (jb >> 5);
    return PfxWords{__ldg(pk), __ldg(pk + 1), __ldg(pk + 2), __ldg(mk), __ldg(mk + 1)};
}

// One strand of one transcript as the lanes see it.
struct PfxCtx {
    uint64_t o;          // off[t]
    uint32_t L, proflen; // length incl. poly(A); entries of this profile (forward holds one more in symmetric mode)
    int k, reverse;
    double T, scale;
    const uint32_t *lut; // quantised weights by 2-bit code (shared or global memory)
    uint32_t kmask, mmask;
};

// Inclusive partial sums x[0..8) of the weights of this lane's 8 entries [ib, ib+8) (0 beyond proflen); returns their total.
// w5 = the packed / mask words at pfx_jb(ib) (pfx_load).  Windows containing a letter outside ACGT are evaluated on the raw
// bytes (nnthermo.go:126-175 with Go's map misses, utils.go:37-59 for the reverse strand).
__device__ __forceinline__ uint32_t pfx_lane_sums(const DevTranscripts &tr, const PfxCtx &c, uint32_t ib, const PfxWords &w5,
                                                  uint32_t (&x)[PFX_ITEMS]) {
    uint32_t run = 0;
    if (ib >= c.proflen) {
#pragma unroll
        for (int r = 0; r < PFX_ITEMS; r++) x[r] = 0;
        return 0;
    }
    const int k = c.k, reverse = c.reverse;
    const uint32_t nvalid = min((uint32_t)PFX_ITEMS, c.proflen - ib);
    const uint32_t jb = pfx_jb(reverse, c.L, k, c.proflen, ib);
    const uint32_t ob = jb & 15, omb = jb & 31;
    // 64 bits of 2-bit codes starting at base jb (covers the 8 windows: 7 + k <= 19 bases), and their N-mask
    const uint64_t lo64 = (uint64_t)w5.p0 | ((uint64_t)w5.p1 << 32);
    uint64_t win = lo64 >> (2 * ob);
    if (ob) win |= (uint64_t)w5.p2 << (64 - 2 * ob);
    const uint32_t win_lo = (uint32_t)win, win_hi = (uint32_t)(win >> 32);
    const uint32_t mwin = (uint32_t)((((uint64_t)w5.m0 | ((uint64_t)w5.m1 << 32)) >> omb));
    const bool clean = (mwin & ((1u << (PFX_ITEMS - 1 + k)) - 1u)) == 0;
    if (clean && nvalid == PFX_ITEMS) {
        // fast path: no non-ACGT letter near, all 8 entries valid - constant shifts, one LUT gather each
#pragma unroll
        for (int r = 0; r < PFX_ITEMS; r++) {
            const int d = reverse ? (PFX_ITEMS - 1 - r) : r;  // window start = jb + d
            const uint32_t code = __funnelshift_r(win_lo, win_hi, 2 * d) & c.kmask;
            run += c.lut[code];
            x[r] = run;
        }
    } else {
#pragma unroll
        for (int r = 0; r < PFX_ITEMS; r++) {
            uint32_t w = 0;
            if ((uint32_t)r < nvalid) {
                const uint32_t d = reverse ? (nvalid - 1 - r) : (uint32_t)r;
                const uint32_t code = (uint32_t)(win >> (2 * d)) & c.kmask;
                const uint32_t mbits = (mwin >> d) & c.mmask;
                if (mbits == 0) {
                    w = c.lut[code];
                } else {
                    // non-ACGT letters in the window: evaluate KmerAffinity on the raw bytes
                    uint8_t kmer[16];
                    const uint8_t *bp = tr.bases + c.o + jb + d;
                    if (!reverse) {
                        for (int q = 0; q < k; q++) kmer[q] = bp[q];
                    } else {
                        for (int q = 0; q < k; q++) {  // utils.go:37-59
                            uint8_t ch = bp[k - 1 - q], oc;
                            switch (ch) { case 'A': oc = 'T'; break; case 'T': oc = 'A'; break; case 'G': oc = 'C'; break; case 'C': oc = 'G'; break; default: oc = 'N'; }
                            kmer[q] = oc;
                        }
                    }
                    w = quantize(kmer_affinity_bytes(kmer, k, c.T), c.scale);
                }
            }
            run += w;
            x[r] = run;
        }
    }
    return run;
}

}  // namespace rl

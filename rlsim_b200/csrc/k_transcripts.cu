// Transcript ingest and priming-affinity kernels (kernel group 1a).
//
//   k_ingest       : transcript.go:98-117 (append polyAmax 'A's) + 2-bit packing + non-ACGT mask + GC prefix counts
//                    (replaces the per-fragment byte scan of thermocycler.go:161-166)
//   k_kmer_lut     : nnthermo.go:126-175 KmerAffinity for all 4^k ACGT k-mers (replaces the string-keyed memo map)
//   k_prefix       : nnthermo.go:177-201 binding profiles, materialised as fixed-point prefix sums so that
//                    SimulatePriming (nnthermo.go:203-217) becomes a binary search instead of an O(L) cumsum
#include <algorithm>
#include "kernels.cuh"
#include "detmath.cuh"

namespace rl {

__device__ __forceinline__ int base_code(uint8_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

// ---------------------------------------------------------------- ingest
// One CTA per transcript; each thread handles 32 consecutive bases per tile.
constexpr int INGEST_THREADS = 128;

__global__ void __launch_bounds__(INGEST_THREADS)
k_ingest(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off, uint32_t n, uint32_t polya_max,
         const uint64_t *__restrict__ off, uint8_t *__restrict__ bases, uint32_t *__restrict__ packed,
         uint32_t *__restrict__ nmask, uint32_t *__restrict__ gc) {
    __shared__ uint32_t warp_tot[INGEST_THREADS / 32];
    __shared__ uint32_t carry_s;
    uint32_t t = blockIdx.x;
    if (t >= n) return;
    const uint64_t so = src_off[t];
    const uint32_t l = (uint32_t)(src_off[t + 1] - so);
    const uint32_t L = l + polya_max;
    const uint64_t o = off[t];
    const uint32_t Lpad = (L + 31u) & ~31u;
    uint32_t *gcp = gc + o + t;  // (L+1) entries
    if (threadIdx.x == 0) { carry_s = 0; gcp[0] = 0; }
    __syncthreads();
    for (uint32_t tile = 0; tile < Lpad; tile += INGEST_THREADS * 32) {
        uint32_t p0 = tile + threadIdx.x * 32;
        uint32_t pk0 = 0, pk1 = 0, mk = 0, cnt = 0;
        uint32_t bytes[8];
        if (p0 < Lpad) {
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t word = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t p = p0 + w * 4 + b;
                    uint8_t c = p < l ? src[so + p] : (p < L ? (uint8_t)'A' : (uint8_t)0);
                    word |= (uint32_t)c << (8 * b);
                    int code = base_code(c);
                    int i = w * 4 + b;
                    if (code < 0) { mk |= 1u << i; code = 0; }
                    if (i < 16) pk0 |= (uint32_t)code << (2 * i); else pk1 |= (uint32_t)code << (2 * (i - 16));
                    cnt += (c == 'G' || c == 'C');
                }
                bytes[w] = word;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(bases + o + p0);
            dst[0] = make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]);
            dst[1] = make_uint4(bytes[4], bytes[5], bytes[6], bytes[7]);
            packed[((o + p0) >> 4)] = pk0;
            packed[((o + p0) >> 4) + 1] = pk1;
            nmask[(o + p0) >> 5] = mk;
        }
        // block exclusive scan of cnt
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += v;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t base = carry_s;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) base += warp_tot[w];
        uint32_t run = base + incl - cnt;
        if (p0 < L) {
#pragma unroll
            for (int w = 0; w < 8; w++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t p = p0 + w * 4 + b;
                    uint8_t c = (uint8_t)(bytes[w] >> (8 * b));
                    run += (c == 'G' || c == 'C');
                    if (p < L) gcp[p + 1] = run;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == INGEST_THREADS - 1) carry_s = base + incl;
        __syncthreads();
    }
}

void launch_ingest(cudaStream_t st, const uint8_t *src, const uint64_t *src_off, uint32_t n, uint32_t polya_max,
                   const uint64_t *off, uint8_t *bases, uint32_t *packed, uint32_t *nmask, uint32_t *gc) {
    if (n) k_ingest<<<n, INGEST_THREADS, 0, st>>>(src, src_off, n, polya_max, off, bases, packed, nmask, gc);
}

// ---------------------------------------------------------------- k-mer affinities
// nnthermo.go:68-104 doublet tables, index 4*a+b with A C G T = 0 1 2 3
__constant__ double NN_DH[16] = {-7.6, -8.4, -7.8, -7.2, -8.5, -8.0, -10.6, -7.8, -8.2, -9.8, -8.0, -8.4, -7.2, -8.2, -8.5, -7.6};
__constant__ double NN_DS[16] = {-21.3, -22.4, -21.0, -20.4, -22.7, -19.9, -27.2, -21.0, -22.2, -24.4, -19.9, -22.4, -21.3, -22.2, -22.7, -21.3};

// KmerAffinity over k bytes (any alphabet), nnthermo.go:126-175
__device__ double kmer_affinity_bytes(const uint8_t *kmer, int kl, double T) {
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

// code: base i of the k-mer in bits [2i, 2i+2)
__global__ void k_kmer_lut(int k, double T, double *__restrict__ K, unsigned long long *__restrict__ kmax_bits) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 1u << (2 * k);
    if (c >= n) return;
    uint8_t kmer[16];
    for (int i = 0; i < k; i++) kmer[i] = (uint8_t)("ACGT"[(c >> (2 * i)) & 3]);
    double v = kmer_affinity_bytes(kmer, k, T);
    K[c] = v;
    atomicMax(kmax_bits, (unsigned long long)__double_as_longlong(v));  // positive doubles order like their bit patterns
}

// spec 4.3: w = min(floor(K * 2^26 / Kmax + 1/2), 2^26 - 1)
#define PFX_SCALE 67108864.0
__device__ __forceinline__ uint32_t quantize(double K, double scale) {
    double x = K * scale + 0.5;
    if (!(x < 67108863.0)) return 67108863u;
    return (uint32_t)x;
}

__device__ __forceinline__ uint32_t rc_code(uint32_t c, int k) {
    uint32_t r = 0;
    for (int i = 0; i < k; i++) r |= (3u - ((c >> (2 * i)) & 3u)) << (2 * (k - 1 - i));
    return r;
}

// wq_f[c] = weight of the k-mer c;  wq_r[c] = weight of its reverse complement
__global__ void k_kmer_quant(int k, const double *__restrict__ K, const unsigned long long *__restrict__ kmax_bits,
                             uint32_t *__restrict__ wq_f, uint32_t *__restrict__ wq_r) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 1u << (2 * k);
    if (c >= n) return;
    double scale = PFX_SCALE / __longlong_as_double((long long)*kmax_bits);
    wq_f[c] = quantize(K[c], scale);
    wq_r[c] = quantize(K[rc_code(c, k)], scale);
}

void launch_kmer_lut(cudaStream_t st, int k, double T, double *K, unsigned long long *kmax_bits, uint32_t *wq_f,
                     uint32_t *wq_r) {
    uint32_t n = 1u << (2 * k);
    cudaMemsetAsync(kmax_bits, 0, 8, st);
    k_kmer_lut<<<(n + 127) / 128, 128, 0, st>>>(k, T, K, kmax_bits);
    k_kmer_quant<<<(n + 127) / 128, 128, 0, st>>>(k, K, kmax_bits, wq_f, wq_r);
}

// ---------------------------------------------------------------- priming prefix sums
constexpr int PFX_THREADS = 256;                 // 8 independent warps per CTA
constexpr int PFX_ITEMS = 8;                     // profile entries per lane and sub-tile
constexpr int PFX_SUB = 32 * PFX_ITEMS;          // 256 entries per warp sub-tile = 4 blocks of 64
constexpr int PFX_ROW_U32 = PFX_ITEMS + 4;       // padded row of one lane in the transpose tile (48 B: 16 B accesses stay conflict-free)
constexpr int PFX_TILE_U32 = 32 * PFX_ROW_U32;

// One WARP per (transcript, strand); persistent CTAs fetch work from a global counter, so the 16 KB LUT is loaded
// once per CTA and there is no block-level barrier in the loop.  Forward profile entry i is the k-mer at
// plus[i, i+k); reverse entry i is the reverse complement of plus[j, j+k), j = L-k-i (= minus[i, i+k),
// transcript.go:110 + nnthermo.go:184-199).  Both have L-k entries (nnthermo.go:193).  Output: the two-level prefix
// sums of common.cuh (4 B per entry + 8 B per 64 entries).
//
// Each lane owns 8 consecutive entries of a 256-entry sub-tile: it loads the 3 packed words and 2 mask words covering
// its 8 windows (read-only path; neighbouring lanes share the words), extracts the 8 k-mer codes by shifts, gathers
// the weights from the shared-memory LUT; an 8-lane segmented shuffle scan gives the in-block sums, four shuffles the
// block totals.  The words of the next sub-tile are requested before the current one is processed; the sums go
// through a per-warp shared-memory tile so that every store instruction covers 512 contiguous bytes.
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

__global__ void __launch_bounds__(PFX_THREADS, 4)
k_prefix(DevTranscripts tr, uint32_t t_begin, uint32_t t_end, int k, double T, const uint32_t *__restrict__ wq_f,
         const uint32_t *__restrict__ wq_r, const unsigned long long *__restrict__ kmax_bits, uint32_t *__restrict__ Ff,
         uint32_t *__restrict__ Fr, uint64_t *__restrict__ Cf, uint64_t *__restrict__ Cr, int lut_in_smem,
         unsigned int *__restrict__ work) {
    extern __shared__ __align__(16) unsigned char pfx_smem[];
    uint32_t *tile_s = reinterpret_cast<uint32_t *>(pfx_smem);                 // 8 warps x PFX_TILE_U32
    uint32_t *lut_s = tile_s + (PFX_THREADS / 32) * PFX_TILE_U32;
    const int reverse = blockIdx.y;
    const uint32_t *lut = reverse ? wq_r : wq_f;
    if (lut_in_smem) {
        uint32_t n = 1u << (2 * k);
        for (uint32_t i = threadIdx.x; i < n; i += PFX_THREADS) lut_s[i] = lut[i];
        lut = lut_s;
    }
    __syncthreads();
    const double scale = PFX_SCALE / __longlong_as_double((long long)*kmax_bits);
    const uint32_t kmask = (1u << (2 * k)) - 1u;
    const uint32_t mmask = (1u << k) - 1u;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t *tile = tile_s + (threadIdx.x >> 5) * PFX_TILE_U32;

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = t_begin + atomicAdd(&work[reverse], 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= t_end) break;
        if (tr.expr[t] == 0) continue;  // pool.go:104: never fragmented
        const uint32_t L = tr.len[t];
        const uint64_t o = tr.off[t];
        uint32_t *F = (reverse ? Fr : Ff) + o;
        uint64_t *C = (reverse ? Cr : Cf) + coarse_base(o, t);
        if (lane == 0) C[0] = 0;
        if (L <= (uint32_t)k) continue;
        const uint32_t proflen = L - (uint32_t)k;
        uint64_t carry = 0;  // sum of all weights before this sub-tile

        uint32_t ib = lane * PFX_ITEMS;
        PfxWords nxt{0, 0, 0, 0, 0};
        if (ib < proflen) nxt = pfx_load(tr, o, pfx_jb(reverse, L, k, proflen, ib));
        for (uint32_t i0 = 0; i0 < proflen; i0 += PFX_SUB) {
            ib = i0 + lane * PFX_ITEMS;
            const PfxWords w5 = nxt;
            const uint32_t ibn = ib + PFX_SUB;
            if (ibn < proflen) nxt = pfx_load(tr, o, pfx_jb(reverse, L, k, proflen, ibn));  // prefetch the next sub-tile
            uint32_t x[PFX_ITEMS];
            uint32_t run = 0;
            if (ib < proflen) {
                const uint32_t nvalid = min((uint32_t)PFX_ITEMS, proflen - ib);
                const uint32_t jb = pfx_jb(reverse, L, k, proflen, ib);
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
                        const uint32_t code = __funnelshift_r(win_lo, win_hi, 2 * d) & kmask;
                        run += lut[code];
                        x[r] = run;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < PFX_ITEMS; r++) {
                        uint32_t w = 0;
                        if ((uint32_t)r < nvalid) {
                            const uint32_t d = reverse ? (nvalid - 1 - r) : (uint32_t)r;
                            const uint32_t code = (uint32_t)(win >> (2 * d)) & kmask;
                            const uint32_t mbits = (mwin >> d) & mmask;
                            if (mbits == 0) {
                                w = lut[code];
                            } else {
                                // non-ACGT letters in the window: evaluate KmerAffinity on the raw bytes
                                uint8_t kmer[16];
                                const uint8_t *bp = tr.bases + o + jb + d;
                                if (!reverse) {
                                    for (int q = 0; q < k; q++) kmer[q] = bp[q];
                                } else {
                                    for (int q = 0; q < k; q++) {  // utils.go:37-59
                                        uint8_t c = bp[k - 1 - q], oc;
                                        switch (c) { case 'A': oc = 'T'; break; case 'T': oc = 'A'; break; case 'G': oc = 'C'; break; case 'C': oc = 'G'; break; default: oc = 'N'; }
                                        kmer[q] = oc;
                                    }
                                }
                                w = quantize(kmer_affinity_bytes(kmer, k, T), scale);
                            }
                        }
                        run += w;
                        x[r] = run;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < PFX_ITEMS; r++) x[r] = 0;
            }
            // in-block sums: inclusive scan of the lane totals inside each group of 8 lanes (= one 64-entry block)
            uint32_t incl = run;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d, 8);
                if ((lane & 7) >= (uint32_t)d) incl += v;
            }
            const uint32_t excl = incl - run;
            // block totals of the 4 blocks of this sub-tile -> coarse sums
            const uint32_t g0 = __shfl_sync(0xffffffffu, incl, 7), g1 = __shfl_sync(0xffffffffu, incl, 15),
                           g2 = __shfl_sync(0xffffffffu, incl, 23), g3 = __shfl_sync(0xffffffffu, incl, 31);
            if (lane < 4) {  // coarse[(i0 >> 6) + lane + 1] = sum of everything before block lane+1 of this sub-tile
                uint64_t cs = carry + g0;
                if (lane >= 1) cs += g1;
                if (lane >= 2) cs += g2;
                if (lane >= 3) cs += g3;
                if (i0 + lane * PFX_BLOCK < proflen) C[(i0 >> PFX_BLOCK_SHIFT) + lane + 1] = cs;
            }
            carry += (uint64_t)g0 + g1 + g2 + g3;
            // transpose through this warp's shared-memory tile so that every store instruction covers 512 contiguous bytes
            {
                uint4 *row = reinterpret_cast<uint4 *>(tile + lane * PFX_ROW_U32);
                row[0] = make_uint4(excl + x[0], excl + x[1], excl + x[2], excl + x[3]);
                row[1] = make_uint4(excl + x[4], excl + x[5], excl + x[6], excl + x[7]);
            }
            __syncwarp();
            const uint32_t nsub = min((uint32_t)PFX_SUB, proflen - i0);
            uint32_t *dst = F + i0;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t e = 4 * (q * 32 + lane);  // first of the four entries this lane stores
                const uint32_t *src = tile + (e >> 3) * PFX_ROW_U32 + (e & 7);
                if (e + 3 < nsub) *reinterpret_cast<uint4 *>(dst + e) = *reinterpret_cast<const uint4 *>(src);
                else for (uint32_t z = 0; z < 4 && e + z < nsub; z++) dst[e + z] = src[z];
            }
            __syncwarp();
        }
    }
}

void launch_prefix(cudaStream_t st, const DevTranscripts &tr, uint32_t t_begin, uint32_t t_end, int k, double T,
                   const uint32_t *wq_f, const uint32_t *wq_r, const unsigned long long *kmax_bits, uint32_t *Ff, uint32_t *Fr,
                   uint64_t *Cf, uint64_t *Cr, int strands, unsigned int *work) {
    if (t_end <= t_begin) return;
    int lut_in_smem = (k <= 6);
    size_t smem = (size_t)(PFX_THREADS / 32) * PFX_TILE_U32 * 4 + (lut_in_smem ? ((size_t)4 << (2 * k)) : 16);
    uint32_t warps_needed = t_end - t_begin;
    uint32_t blocks = std::min<uint32_t>((warps_needed + 7) / 8, 148u * 6u);
    cudaMemsetAsync(work, 0, 2 * sizeof(unsigned int), st);
    dim3 grid(blocks, strands);
    k_prefix<<<grid, PFX_THREADS, smem, st>>>(tr, t_begin, t_end, k, T, wq_f, wq_r, kmax_bits, Ff, Fr, Cf, Cr, lut_in_smem, work);
}

}  // namespace rl

// Transcript ingest and priming-affinity kernels (kernel group 1a).
//
//   k_ingest       : transcript.go:98-117 (append polyAmax 'A's) + 2-bit packing + non-ACGT mask + GC prefix counts
//                    (replaces the per-fragment byte scan of thermocycler.go:161-166)
//   k_kmer_lut     : nnthermo.go:126-175 KmerAffinity for all 4^k ACGT k-mers (replaces the string-keyed memo map)
//   k_prefix       : nnthermo.go:177-201 binding profiles, materialised as fixed-point prefix sums so that
//                    SimulatePriming (nnthermo.go:203-217) becomes a binary search instead of an O(L) cumsum
#include <algorithm>
#include "kernels.cuh"
#include "prefix.cuh"

namespace rl {

// ---------------------------------------------------------------- ingest
// Pass 1 (k_ingest): one thread per 32-base group of the padded layout.  The source batch is the plain concatenation of
// the transcripts, so consecutive threads read consecutive source bytes (aligned word loads + funnel shift for the
// arbitrary source alignment) and write consecutive words of every output array.  Four bases are classified at a time
// with byte-parallel integer operations.
// Pass 2 (k_gc_scan): one warp per transcript turns the per-group G/C counts into exclusive prefix counts.
constexpr int INGEST_THREADS = 256;

__device__ __forceinline__ uint32_t byte_mask(int64_t k) {  // low k bytes set, k clamped to [0, 4]
    return k <= 0 ? 0u : (k >= 4 ? 0xffffffffu : ((1u << (8 * (int)k)) - 1u));
}
__device__ __forceinline__ uint32_t gather_bit0(uint32_t z) {  // bit 0 of each byte -> 4 adjacent bits
    z &= 0x01010101u;
    return (z | (z >> 7) | (z >> 14) | (z >> 21)) & 0xfu;
}

// Transcript of the warp's first 32-base group (padded position Pw): 32-ary search, every lane probes one split point per
// round.  Returns lo with off[lo] <= Pw < off[hi]; must be called by all 32 lanes.
__device__ __forceinline__ uint32_t warp_first_transcript(const uint64_t *__restrict__ off, uint32_t n, uint64_t Pw, int lane) {
    uint32_t lo = 0, hi = n;  // invariant: off[lo] <= Pw < off[hi]
    while (hi - lo > 1) {
        uint32_t span = hi - lo;
        uint32_t probe = lo + (uint32_t)(((uint64_t)span * (lane + 1)) / 33);  // lo <= probe < hi
        bool le = off[probe] <= Pw;
        uint32_t b = __ballot_sync(0xffffffffu, le);
        // the probes are nondecreasing in the lane index; off[] is nondecreasing: ballot is a prefix of ones
        int k = __popc(b);  // lanes 0..k-1 satisfy off[probe] <= Pw
        uint32_t new_lo = k ? __shfl_sync(0xffffffffu, probe, k - 1) : lo;
        uint32_t new_hi = k < 32 ? __shfl_sync(0xffffffffu, probe, k) : hi;
        if (new_lo == lo && new_hi == hi) {  // span too small for the probes to separate: finish linearly
            while (lo + 1 < hi && off[lo + 1] <= Pw) lo++;
            hi = lo + 1;
        } else { lo = new_lo; hi = new_hi; }
    }
    return lo;
}

__global__ void __launch_bounds__(INGEST_THREADS)
k_ingest(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off, uint32_t n, uint32_t t_index0, uint32_t polya_max,
         const uint64_t *__restrict__ off, uint8_t *__restrict__ bases, uint32_t *__restrict__ packed,
         uint32_t *__restrict__ nmask, uint2 *__restrict__ gc, uint32_t *__restrict__ dirty) {
    const uint64_t P0 = off[0], n_groups = (off[n] - P0) >> 5;
    const uint64_t g = (uint64_t)blockIdx.x * INGEST_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint64_t gw = min(g - lane, n_groups ? n_groups - 1 : 0);
    const uint32_t lo = warp_first_transcript(off, n, P0 + (gw << 5), lane);
    if (g >= n_groups) return;
    const uint64_t P = P0 + (g << 5);
    uint32_t t = lo;
    while (off[t + 1] <= P) t++;  // later lanes may sit in later transcripts (zero-length ones are skipped)
    const uint64_t so = src_off[t];
    const uint32_t l = (uint32_t)(src_off[t + 1] - so);
    const uint32_t L = l + polya_max;
    const uint32_t p0 = (uint32_t)(P - off[t]);
    uint32_t w[9];
#pragma unroll
    for (int j = 0; j < 9; j++) w[j] = 0;
    uint32_t sh = 0;
    if (p0 < l) {
        const uint8_t *a = src + so + p0;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(a) & 3u);
        const uint32_t *aw = reinterpret_cast<const uint32_t *>(a - mis);
        const uint32_t n_words = (min(l - p0, 32u) + mis + 3u) >> 2;
#pragma unroll
        for (int j = 0; j < 9; j++)
            if ((uint32_t)j < n_words) w[j] = __ldg(aw + j);
        sh = 8 * mis;
    }
    uint32_t out[8], pk0 = 0, pk1 = 0, mk = 0, gm = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t x = __funnelshift_r(w[j], w[j + 1], sh);
        const int64_t q = (int64_t)p0 + 4 * j;
        const uint32_t ms = byte_mask((int64_t)l - q), ma = byte_mask((int64_t)L - q);
        x = (x & ms) | (0x41414141u & ma & ~ms);  // transcript.go:98-117: the poly(A) tail; zero padding after it
        out[j] = x;
        const uint32_t v = __vcmpeq4(x, 0x41414141u) | __vcmpeq4(x, 0x43434343u) | __vcmpeq4(x, 0x47474747u) | __vcmpeq4(x, 0x54545454u);
        const uint32_t y = ((x >> 1) ^ (x >> 2)) & 0x03030303u & v;  // A C G T -> 0 1 2 3, anything else 0 + mask bit
        const uint32_t code8 = (y | (y >> 6) | (y >> 12) | (y >> 18)) & 0xffu;
        if (j < 4) pk0 |= code8 << (8 * j); else pk1 |= code8 << (8 * (j - 4));
        mk |= gather_bit0(~v) << (4 * j);
        gm |= gather_bit0(y ^ (y >> 1)) << (4 * j);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(bases + P);
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    *reinterpret_cast<uint2 *>(packed + (P >> 4)) = make_uint2(pk0, pk1);
    nmask[P >> 5] = mk;
    {   // letters outside ACGT among the real positions [p0, L) of this group
        const uint32_t nreal = L > p0 ? min(L - p0, 32u) : 0u;
        const uint32_t real = nreal >= 32 ? 0xffffffffu : ((1u << nreal) - 1u);
        if (mk & real) atomicOr(&dirty[t_index0 + t], 1u);
    }
    gc[(P >> 5) + t_index0 + t] = make_uint2(__popc(gm), gm);
}

__global__ void __launch_bounds__(INGEST_THREADS)
k_gc_scan(uint32_t n, uint32_t t_index0, const uint64_t *__restrict__ off, uint2 *__restrict__ gc) {
    const uint32_t t = blockIdx.x * (INGEST_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n) return;
    const uint32_t J = (uint32_t)((off[t + 1] - off[t]) >> 5);
    uint2 *g = gc + (off[t] >> 5) + t_index0 + t;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < J; base += 32) {
        const uint32_t j = base + lane;
        uint2 e = j < J ? g[j] : make_uint2(0, 0);
        uint32_t incl = e.x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (j < J) g[j] = make_uint2(carry + incl - e.x, e.y);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) g[J] = make_uint2(carry, 0);
}

// src_off / off point at the first transcript of the batch; t_index0 is its index on this handle
void launch_ingest(cudaStream_t st, const uint8_t *src, const uint64_t *src_off, uint32_t n, uint32_t t_index0, uint64_t n_groups,
                   uint32_t polya_max, const uint64_t *off, uint8_t *bases, uint32_t *packed, uint32_t *nmask, uint2 *gc,
                   uint32_t *dirty) {
    if (!n) return;
    cudaMemsetAsync(dirty + t_index0, 0, (size_t)n * 4, st);
    if (n_groups)
        k_ingest<<<(unsigned)((n_groups + INGEST_THREADS - 1) / INGEST_THREADS), INGEST_THREADS, 0, st>>>(
            src, src_off, n, t_index0, polya_max, off, bases, packed, nmask, gc, dirty);
    k_gc_scan<<<(n + INGEST_THREADS / 32 - 1) / (INGEST_THREADS / 32), INGEST_THREADS, 0, st>>>(n, t_index0, off, gc);
}

// ---------------------------------------------------------------- ingest of a packed batch (rlsim_add_transcripts_packed)
// The same outputs as k_ingest from a source that is already 2-bit packed: `codes` = 4 bases per byte (base i of the batch
// at bits 2*(i&3) of byte i>>2; A C G T = 0 1 2 3), `mask` = 1 bit per base (bit i&7 of byte i>>3: a letter outside ACGT,
// carried as 'N') or null.  Both device copies start at a 4-byte boundary of the source arrays: c_shift / m_shift = bases
// between that boundary and the first base of the batch.  0.375 B per base on the host link instead of 1.
__device__ __forceinline__ uint32_t even_bits(uint64_t x) {  // bits 0, 2, 4 ... 62 -> a 32-bit word
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}
__device__ __forceinline__ uint64_t spread_bits(uint32_t m) {  // bit i -> bits 2i and 2i+1
    uint64_t x = m;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x | (x << 1);
}

__global__ void __launch_bounds__(INGEST_THREADS)
k_ingest_packed(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ mask, uint32_t c_shift, uint32_t m_shift,
                const uint64_t *__restrict__ src_off, uint32_t n, uint32_t t_index0, uint32_t polya_max,
                const uint64_t *__restrict__ off, uint8_t *__restrict__ bases, uint32_t *__restrict__ packed,
                uint32_t *__restrict__ nmask, uint2 *__restrict__ gc, uint32_t *__restrict__ dirty) {
    const uint64_t P0 = off[0], n_groups = (off[n] - P0) >> 5;
    const uint64_t g = (uint64_t)blockIdx.x * INGEST_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint64_t gw = min(g - lane, n_groups ? n_groups - 1 : 0);
    uint32_t t = warp_first_transcript(off, n, P0 + (gw << 5), lane);
    if (g >= n_groups) return;
    const uint64_t P = P0 + (g << 5);
    while (off[t + 1] <= P) t++;  // later lanes may sit in later transcripts (zero-length ones are skipped)
    const uint64_t so = src_off[t];
    const uint32_t l = (uint32_t)(src_off[t + 1] - so);
    const uint32_t L = l + polya_max;
    const uint32_t p0 = (uint32_t)(P - off[t]);
    uint64_t pk = 0;
    uint32_t mk = 0;
    if (p0 < l) {
        const uint32_t nsrc = min(l - p0, 32u);
        const uint64_t real2 = nsrc >= 32 ? ~0ull : ((1ull << (2 * nsrc)) - 1ull);
        const uint32_t real1 = nsrc >= 32 ? 0xffffffffu : ((1u << nsrc) - 1u);
        const uint64_t qc = so + p0 + c_shift;  // base position in the device copy of the codes
        const uint32_t *cw = codes + (qc >> 4);
        const uint32_t sh = (uint32_t)(qc & 15) * 2;
        const uint32_t w0 = __ldg(cw), w1 = __ldg(cw + 1), w2 = sh ? __ldg(cw + 2) : 0u;
        pk = ((uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh)) & real2;
        if (mask) {
            const uint64_t qm = so + p0 + m_shift;
            const uint32_t *mw = mask + (qm >> 5);
            const uint32_t ms = (uint32_t)(qm & 31);
            mk = __funnelshift_r(__ldg(mw), ms ? __ldg(mw + 1) : 0u, ms) & real1;
            pk &= ~spread_bits(mk);  // masked letters carry code 0, as k_ingest writes them
        }
    }
    const uint32_t gm = even_bits(pk ^ (pk >> 1));  // G or C: exactly one of the two code bits
    uint32_t out[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t c8 = (uint32_t)(pk >> (8 * j)) & 0xffu;
        const uint32_t sel = (c8 & 3u) | ((c8 & 0xcu) << 2) | ((c8 & 0x30u) << 4) | ((c8 & 0xc0u) << 6);
        uint32_t x = __byte_perm(0x54474341u, 0u, sel);  // "ACGT"[code] in each byte
        const uint32_t m4 = (mk >> (4 * j)) & 0xfu;
        const uint32_t mb = ((m4 & 1u) | ((m4 & 2u) << 7) | ((m4 & 4u) << 14) | ((m4 & 8u) << 21)) * 0xffu;
        x = (x & ~mb) | (0x4e4e4e4eu & mb);
        const int64_t q = (int64_t)p0 + 4 * j;
        out[j] = x & byte_mask((int64_t)L - q);  // source letters, then the poly(A) tail (code 0 = 'A'), then zero padding
    }
    uint4 *dst = reinterpret_cast<uint4 *>(bases + P);
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    *reinterpret_cast<uint2 *>(packed + (P >> 4)) = make_uint2((uint32_t)pk, (uint32_t)(pk >> 32));
    // beyond L the ASCII path marks the zero padding as "not ACGT" in nmask (mask bit of a zero byte): keep that
    const uint32_t nreal = L > p0 ? min(L - p0, 32u) : 0u;
    const uint32_t real = nreal >= 32 ? 0xffffffffu : ((1u << nreal) - 1u);
    nmask[P >> 5] = mk | ~real;
    if (mk & real) atomicOr(&dirty[t_index0 + t], 1u);
    gc[(P >> 5) + t_index0 + t] = make_uint2(__popc(gm), gm);
}

void launch_ingest_packed(cudaStream_t st, const uint32_t *codes, const uint32_t *mask, uint32_t c_shift, uint32_t m_shift,
                          const uint64_t *src_off, uint32_t n, uint32_t t_index0, uint64_t n_groups, uint32_t polya_max,
                          const uint64_t *off, uint8_t *bases, uint32_t *packed, uint32_t *nmask, uint2 *gc, uint32_t *dirty) {
    if (!n) return;
    cudaMemsetAsync(dirty + t_index0, 0, (size_t)n * 4, st);
    if (n_groups)
        k_ingest_packed<<<(unsigned)((n_groups + INGEST_THREADS - 1) / INGEST_THREADS), INGEST_THREADS, 0, st>>>(
            codes, mask, c_shift, m_shift, src_off, n, t_index0, polya_max, off, bases, packed, nmask, gc, dirty);
    k_gc_scan<<<(n + INGEST_THREADS / 32 - 1) / (INGEST_THREADS / 32), INGEST_THREADS, 0, st>>>(n, t_index0, off, gc);
}

// ---------------------------------------------------------------- k-mer affinities
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

__global__ void k_lut_asym(int k, const uint32_t *__restrict__ wq_f, const uint32_t *__restrict__ wq_r, unsigned int *asym) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < (1u << (2 * k)) && wq_f[c] != wq_r[c]) atomicOr(asym, 1u);
}

// *asym != 0 afterwards: some k-mer and its reverse complement quantise to different weights
void launch_kmer_lut(cudaStream_t st, int k, double T, double *K, unsigned long long *kmax_bits, uint32_t *wq_f,
                     uint32_t *wq_r, unsigned int *asym) {
    uint32_t n = 1u << (2 * k);
    cudaMemsetAsync(kmax_bits, 0, 8, st);
    k_kmer_lut<<<(n + 127) / 128, 128, 0, st>>>(k, T, K, kmax_bits);
    k_kmer_quant<<<(n + 127) / 128, 128, 0, st>>>(k, K, kmax_bits, wq_f, wq_r);
    cudaMemsetAsync(asym, 0, 4, st);
    k_lut_asym<<<(n + 127) / 128, 128, 0, st>>>(k, wq_f, wq_r, asym);
}

// ---------------------------------------------------------------- priming prefix sums
constexpr int PFX_THREADS = 256;                 // 8 independent warps per CTA
constexpr uint32_t PFX_LONG = 8192;              // profile entries from which a transcript is summed by a whole CTA (k_prefix_long)
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
// BULK: the 1 KB of fine sums a warp produces per sub-tile leaves shared memory as ONE bulk asynchronous copy
// (cp.async.bulk.global.shared::cta, the TMA engine; SASS UBLKCP) issued by lane 0, double-buffered, instead of two 16-byte
// stores per lane behind a transposing read of the tile.
constexpr int PFX_BULK_TILE_U32 = 2 * PFX_SUB;   // two contiguous 256-entry buffers per warp
template <bool BULK>
__global__ void __launch_bounds__(PFX_THREADS, 4)
k_prefix(DevTranscripts tr, uint32_t t_begin, uint32_t t_end, int k, double T, const uint32_t *__restrict__ wq_f,
         const uint32_t *__restrict__ wq_r, const unsigned long long *__restrict__ kmax_bits, uint32_t *__restrict__ Ff,
         uint32_t *__restrict__ Fr, uint64_t *__restrict__ Cf, uint64_t *__restrict__ Cr, Rec32 *__restrict__ Rf, int lut_in_smem,
         unsigned int *__restrict__ work, uint32_t *__restrict__ long_list, uint32_t long_cap) {
    extern __shared__ __align__(16) unsigned char pfx_smem[];
    constexpr int TILE_U32 = BULK ? PFX_BULK_TILE_U32 : PFX_TILE_U32;
    uint32_t *tile_s = reinterpret_cast<uint32_t *>(pfx_smem);                 // 8 warps x TILE_U32
    uint32_t *lut_s = tile_s + (PFX_THREADS / 32) * TILE_U32;
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
    uint32_t *tile = tile_s + (threadIdx.x >> 5) * TILE_U32;
    uint32_t bulk_buf = 0;  // BULK: which of the warp's two buffers the next sub-tile uses

    // Transcripts are short (~9 sub-tiles): the chain work counter -> transcript metadata -> first words is three dependent
    // round trips per transcript, as long as the sums themselves.  The loop is software-pipelined two deep: the counter
    // for the transcript after next and the metadata of the next one are in flight while the current one is summed.
    auto fetch = [&]() -> uint32_t { return lane == 0 ? t_begin + atomicAdd(&work[reverse], 1u) : 0u; };  // result pending on lane 0
    struct Meta { uint32_t t, L, expressed, dirty; uint64_t o; };
    auto load_meta = [&](uint32_t t) {
        Meta mt{t, 0u, 0u, 0u, 0ull};
        if (t < t_end) { mt.L = tr.len[t]; mt.o = tr.off[t]; mt.expressed = tr.expr[t] != 0; mt.dirty = tr.dirty[t]; }
        return mt;
    };
    Meta cur = load_meta(__shfl_sync(0xffffffffu, fetch(), 0));
    uint32_t pending = fetch();
    for (;; ) {
        if (cur.t >= t_end) break;
        const Meta mt = cur;
        cur = load_meta(__shfl_sync(0xffffffffu, pending, 0));  // the next transcript's metadata: in flight from here on
        pending = fetch();
        const uint32_t t = mt.t;
        if (!mt.expressed) continue;  // pool.go:104: never fragmented
        const uint32_t L = mt.L;
        const uint64_t o = mt.o;
        uint32_t *F = (reverse ? Fr : Ff) + o;
        uint64_t *C = (reverse ? Cr : Cf) + coarse_base(o, t);
        if (tr.sym && reverse && mt.dirty == 0) continue;  // read mirrored from the forward arrays (priming_mirror)
        // record layout (common.cuh): clean transcripts write one 32-byte record per block instead of 4 B per entry
        const bool use_rec = Rf != nullptr && mt.dirty == 0;
        Rec32 *R = use_rec ? Rf + rec_base(o, t) : nullptr;
        if (lane == 0) { if (use_rec) R[0].coarse = 0; else C[0] = 0; }
        if (L < (uint32_t)k + (tr.sym && !reverse ? 0u : 1u)) continue;
        // symmetric mode: the forward arrays also hold window len-k, the first window of the reverse profile
        const uint32_t proflen = L - (uint32_t)k + (tr.sym && !reverse ? 1u : 0u);
        if (long_list != nullptr && proflen > PFX_LONG && !use_rec) {  // a whole CTA sums it (k_prefix_long): no one-warp tail
            if (lane == 0) { const uint32_t slot = atomicAdd(&work[2 + reverse], 1u); if (slot < long_cap) long_list[(size_t)reverse * long_cap + slot] = t; }
            continue;
        }
        uint64_t carry = 0;  // sum of all weights before this sub-tile
        const PfxCtx pc{o, L, proflen, k, reverse, T, scale, lut, kmask, mmask};

        uint32_t ib = lane * PFX_ITEMS;
        PfxWords nxt{0, 0, 0, 0, 0};
        if (ib < proflen) nxt = pfx_load(tr, o, pfx_jb(reverse, L, k, proflen, ib));
        for (uint32_t i0 = 0; i0 < proflen; i0 += PFX_SUB) {
            ib = i0 + lane * PFX_ITEMS;
            const PfxWords w5 = nxt;
            const uint32_t ibn = ib + PFX_SUB;
            if (ibn < proflen) nxt = pfx_load(tr, o, pfx_jb(reverse, L, k, proflen, ibn));  // prefetch the next sub-tile
            uint32_t x[PFX_ITEMS];
            const uint32_t run = pfx_lane_sums(tr, pc, ib, w5, x);
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
            if (use_rec) {
                // first-half sums of the four blocks sit in lanes 3, 11, 19, 27 of the 8-lane scans
                const uint32_t h0 = __shfl_sync(0xffffffffu, incl, 3), h1 = __shfl_sync(0xffffffffu, incl, 11),
                               h2 = __shfl_sync(0xffffffffu, incl, 19), h3 = __shfl_sync(0xffffffffu, incl, 27);
                if (lane < 4 && i0 + lane * PFX_BLOCK < proflen) {
                    uint64_t before = carry;
                    if (lane >= 1) before += g0;
                    if (lane >= 2) before += g1;
                    if (lane >= 3) before += g2;
                    const uint32_t half = lane == 0 ? h0 : lane == 1 ? h1 : lane == 2 ? h2 : h3;
                    const uint32_t blk = (i0 >> PFX_BLOCK_SHIFT) + lane;
                    const uint32_t *pk = tr.packed + (o >> 4) + 4 * blk;  // 16 bases per word: the block's 80 bases are 5 words
                    uint4 *dst = reinterpret_cast<uint4 *>(R + blk);
                    dst[0] = make_uint4((uint32_t)before, (uint32_t)(before >> 32), half, __ldg(pk));
                    dst[1] = make_uint4(__ldg(pk + 1), __ldg(pk + 2), __ldg(pk + 3), __ldg(pk + 4));
                }
                carry += (uint64_t)g0 + g1 + g2 + g3;
                continue;
            }
            if (lane < 4) {  // coarse[(i0 >> 6) + lane + 1] = sum of everything before block lane+1 of this sub-tile
                uint64_t cs = carry + g0;
                if (lane >= 1) cs += g1;
                if (lane >= 2) cs += g2;
                if (lane >= 3) cs += g3;
                if (i0 + lane * PFX_BLOCK < proflen) C[(i0 >> PFX_BLOCK_SHIFT) + lane + 1] = cs;
            }
            carry += (uint64_t)g0 + g1 + g2 + g3;
            if (BULK) {
                uint32_t *tb = tile + bulk_buf * PFX_SUB;
                bulk_buf ^= 1u;
                // the copy issued two sub-tiles ago has read this buffer (at most the last one may still be pending)
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                uint4 *row = reinterpret_cast<uint4 *>(tb + lane * PFX_ITEMS);
                row[0] = make_uint4(excl + x[0], excl + x[1], excl + x[2], excl + x[3]);
                row[1] = make_uint4(excl + x[4], excl + x[5], excl + x[6], excl + x[7]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the copy engine
                __syncwarp();
                if (lane == 0) {
                    const uint32_t nsub = min((uint32_t)PFX_SUB, proflen - i0);
                    const uint32_t bytes = ((nsub + 3u) & ~3u) * 4u;  // whole 16-byte units: up to 3 entries of the transcript's padding
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(F + i0), "r"((uint32_t)__cvta_generic_to_shared(tb)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
            }
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
        if (use_rec && lane == 0) R[(proflen + PFX_BLOCK - 1) >> PFX_BLOCK_SHIFT].coarse = carry;  // closing record: the total
    }
    if (BULK && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory outlives the copies
}

// Long transcripts (more than PFX_LONG profile entries; k_prefix lists them): one CTA per transcript, its 8 warps sum 8
// consecutive sub-tiles per step.  The in-block sums need nothing from outside their block; the coarse sums need the total
// of everything before, which the warps exchange through shared memory once per step (double-buffered: one barrier).
// A 30 000-base transcript is 15 steps here instead of 117 dependent iterations of one warp - which, chunk by chunk, was
// the floor of every k_prefix launch of the chunk pipeline.
__global__ void __launch_bounds__(PFX_THREADS, 4)
k_prefix_long(DevTranscripts tr, int k, double T, const uint32_t *__restrict__ wq_f, const uint32_t *__restrict__ wq_r,
              const unsigned long long *__restrict__ kmax_bits, uint32_t *__restrict__ Ff, uint32_t *__restrict__ Fr,
              uint64_t *__restrict__ Cf, uint64_t *__restrict__ Cr, int lut_in_smem, const unsigned int *__restrict__ work,
              const uint32_t *__restrict__ long_list, uint32_t long_cap) {
    extern __shared__ __align__(16) unsigned char pfx_smem[];
    __shared__ unsigned long long totals[2][PFX_THREADS / 32];
    uint32_t *tile_s = reinterpret_cast<uint32_t *>(pfx_smem);
    uint32_t *lut_s = tile_s + (PFX_THREADS / 32) * PFX_TILE_U32;
    const int reverse = blockIdx.y;
    const uint32_t n_long = min(work[2 + reverse], long_cap);
    if (blockIdx.x >= n_long) return;
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
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *tile = tile_s + warp * PFX_TILE_U32;
    int buf = 0;
    for (uint32_t item = blockIdx.x; item < n_long; item += gridDim.x) {
        const uint32_t t = long_list[(size_t)reverse * long_cap + item];
        const uint32_t L = tr.len[t];
        const uint64_t o = tr.off[t];
        uint32_t *F = (reverse ? Fr : Ff) + o;
        uint64_t *C = (reverse ? Cr : Cf) + coarse_base(o, t);
        const uint32_t proflen = L - (uint32_t)k + (tr.sym && !reverse ? 1u : 0u);
        const PfxCtx pc{o, L, proflen, k, reverse, T, scale, lut, kmask, mmask};
        uint64_t carry = 0;  // sum of all weights before this step
        for (uint32_t s0 = 0; s0 < proflen; s0 += (PFX_THREADS / 32) * PFX_SUB) {
            const uint32_t i0 = s0 + warp * PFX_SUB;
            const bool active = i0 < proflen;
            uint32_t x[PFX_ITEMS];
            uint32_t excl = 0, g0 = 0, g1 = 0, g2 = 0, g3 = 0;
            if (active) {
                const uint32_t ib = i0 + lane * PFX_ITEMS;
                PfxWords w5{0, 0, 0, 0, 0};
                if (ib < proflen) w5 = pfx_load(tr, o, pfx_jb(reverse, L, k, proflen, ib));
                const uint32_t run = pfx_lane_sums(tr, pc, ib, w5, x);
                uint32_t incl = run;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, incl, d, 8);
                    if ((lane & 7) >= (uint32_t)d) incl += v;
                }
                excl = incl - run;
                g0 = __shfl_sync(0xffffffffu, incl, 7); g1 = __shfl_sync(0xffffffffu, incl, 15);
                g2 = __shfl_sync(0xffffffffu, incl, 23); g3 = __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) totals[buf][warp] = (unsigned long long)g0 + g1 + g2 + g3;
            __syncthreads();
            uint64_t before = carry, all = 0;
#pragma unroll
            for (int w = 0; w < PFX_THREADS / 32; w++) {
                const uint64_t v = totals[buf][w];
                if ((uint32_t)w < warp) before += v;
                all += v;
            }
            carry += all;
            buf ^= 1;
            if (!active) continue;
            if (lane < 4) {  // coarse[(i0 >> 6) + lane + 1] = sum of everything before block lane+1 of this sub-tile
                uint64_t cs = before + g0;
                if (lane >= 1) cs += g1;
                if (lane >= 2) cs += g2;
                if (lane >= 3) cs += g3;
                if (i0 + lane * PFX_BLOCK < proflen) C[(i0 >> PFX_BLOCK_SHIFT) + lane + 1] = cs;
            }
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
                const uint32_t e = 4 * (q * 32 + lane);
                const uint32_t *src = tile + (e >> 3) * PFX_ROW_U32 + (e & 7);
                if (e + 3 < nsub) *reinterpret_cast<uint4 *>(dst + e) = *reinterpret_cast<const uint4 *>(src);
                else for (uint32_t z = 0; z < 4 && e + z < nsub; z++) dst[e + z] = src[z];
            }
            __syncwarp();
        }
    }
}

// work: 4 counters (next transcript per strand, listed long transcripts per strand); long_list: 2 x long_cap entries or null
void launch_prefix(cudaStream_t st, const DevTranscripts &tr, uint32_t t_begin, uint32_t t_end, int k, double T,
                   const uint32_t *wq_f, const uint32_t *wq_r, const unsigned long long *kmax_bits, uint32_t *Ff, uint32_t *Fr,
                   uint64_t *Cf, uint64_t *Cr, Rec32 *Rf, int strands, unsigned int *work, uint32_t *long_list, uint32_t long_cap,
                   bool bulk_stores) {
    if (t_end <= t_begin) return;
    int lut_in_smem = (k <= 6);
    size_t smem = (size_t)(PFX_THREADS / 32) * PFX_TILE_U32 * 4 + (lut_in_smem ? ((size_t)4 << (2 * k)) : 16);
    uint32_t warps_needed = t_end - t_begin;
    uint32_t blocks = std::min<uint32_t>((warps_needed + 7) / 8, 148u * 6u);
    cudaMemsetAsync(work, 0, 4 * sizeof(unsigned int), st);
    dim3 grid(blocks, strands);
    if (Rf != nullptr) long_list = nullptr;  // record layout (opt-in): every transcript stays with its warp
    if (bulk_stores && Rf == nullptr) {
        const size_t smem_b = (size_t)(PFX_THREADS / 32) * PFX_BULK_TILE_U32 * 4 + (lut_in_smem ? ((size_t)4 << (2 * k)) : 16);
        k_prefix<true><<<grid, PFX_THREADS, smem_b, st>>>(tr, t_begin, t_end, k, T, wq_f, wq_r, kmax_bits, Ff, Fr, Cf, Cr, Rf, lut_in_smem, work,
                                                          long_list, long_cap);
    } else
    k_prefix<false><<<grid, PFX_THREADS, smem, st>>>(tr, t_begin, t_end, k, T, wq_f, wq_r, kmax_bits, Ff, Fr, Cf, Cr, Rf, lut_in_smem, work,
                                                     long_list, long_cap);
    if (long_list != nullptr) {
        dim3 lgrid(std::min<uint32_t>(std::max<uint32_t>(long_cap, 1u), 148u * 4u), strands);
        k_prefix_long<<<lgrid, PFX_THREADS, smem, st>>>(tr, k, T, wq_f, wq_r, kmax_bits, Ff, Fr, Cf, Cr, lut_in_smem, work, long_list, long_cap);
    }
}

}  // namespace rl

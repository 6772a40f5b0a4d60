// Device-side FASTA reader: the raw text of the input files in, compact upper-cased sequences plus
// a record table out.  Replaces the reference's byte-at-a-time reader:
//
//   fasta.go:113-160   ReadFasta.NextFasta : a '>' opens a new record once the record buffer holds more than one byte
//   fasta.go:176-197   FastaToSeq.NextSeq  : name = first line without '>', sequence = rest, despaced, upper-cased
//   fasta.go:162-168   isSpace             : ' ', TAB, CR, LF
//   input.go:82-101    ParseSeqName        : exactly one '$', non-empty name, TrimSpace + Sscanf("%d") into a uint64
//   input.go:112-113   level = uint64(float64(level) * exprMul)
//
// The reader is a two-state machine over the bytes (HDR: inside the name line, SEQ: inside the sequence).  A tile of
// bytes is summarised as a function of its incoming state -> (outgoing state, sequence bytes kept, records opened); the
// summaries compose associatively, so the text is processed as scan-of-tiles: k_fasta_tiles (summaries),
// k_fasta_carry (one CTA: exclusive scan over the tiles), k_fasta_emit (records + compacted bytes), k_fasta_headers
// (one thread per record: name length, expression level).
#include "kernels.cuh"

namespace rl {

constexpr int FA_THREADS = 256;
constexpr int FA_BYTES = 16;                       // bytes per thread
constexpr int FA_TILE = FA_THREADS * FA_BYTES;     // bytes per CTA
constexpr int FA_MAX_RUN = 64;                     // longest run of consecutive '>' the device reader accepts

// out: bit s = state after the span when entered in state s (0 = HDR, 1 = SEQ); k[s] = sequence bytes kept
struct FaSum { uint64_t k0, k1; uint32_t out, nb; };

__device__ __forceinline__ FaSum fa_identity() { return FaSum{0, 0, 2u, 0}; }  // out: 0 -> 0, 1 -> 1
__device__ __forceinline__ FaSum fa_compose(const FaSum &a, const FaSum &b) {   // a first, then b
    FaSum r;
    const uint32_t a0 = a.out & 1u, a1 = (a.out >> 1) & 1u;
    r.k0 = a.k0 + (a0 ? b.k1 : b.k0);
    r.k1 = a.k1 + (a1 ? b.k1 : b.k0);
    r.out = ((b.out >> a0) & 1u) | (((b.out >> a1) & 1u) << 1);
    r.nb = a.nb + b.nb;
    return r;
}

// fasta.go:138: `c == '>' && pos > 1`.  Inside a run of consecutive '>' every second one opens a record (the one right
// after an opening '>' lands at buffer position 1 and is kept as part of the name).
__device__ __forceinline__ bool fa_boundary(const uint8_t *__restrict__ text, uint64_t i, int *err) {
    uint64_t j = i;
    int run = 0;
    while (j > 0 && text[j - 1] == '>') {
        j--;
        if (++run > FA_MAX_RUN) { atomicExch(err, RLSIM_E_UNSUPPORTED); return true; }
    }
    return (run & 1) == 0;
}

__device__ __forceinline__ bool fa_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

__device__ __forceinline__ void fa_load(const uint8_t *__restrict__ text, uint64_t n, uint64_t i0, uint8_t (&c)[FA_BYTES]) {
    if (i0 + FA_BYTES <= n) {
        uint4 v = *reinterpret_cast<const uint4 *>(text + i0);  // the text buffer is 16-byte aligned
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < FA_BYTES; j++) c[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
    } else {
#pragma unroll
        for (int j = 0; j < FA_BYTES; j++) c[j] = i0 + j < n ? text[i0 + j] : (uint8_t)' ';
    }
}

__device__ __forceinline__ FaSum fa_thread_summary(const uint8_t *__restrict__ text, uint64_t n, uint64_t i0, const uint8_t (&c)[FA_BYTES], int *err) {
    uint32_t s0 = 0, s1 = 1, k0 = 0, k1 = 0, nb = 0;
#pragma unroll
    for (int j = 0; j < FA_BYTES; j++) {
        const uint8_t ch = c[j];
        if (ch == '>' && i0 + j < n && fa_boundary(text, i0 + j, err)) { nb++; s0 = s1 = 0; }
        else if (ch == '\n') { s0 = s1 = 1; }
        else if (!fa_space(ch)) { k0 += s0; k1 += s1; }
    }
    return FaSum{k0, k1, s0 | (s1 << 1), nb};
}

__device__ __forceinline__ FaSum fa_shfl_up(const FaSum &v, int d) {
    FaSum r;
    r.k0 = __shfl_up_sync(0xffffffffu, v.k0, d); r.k1 = __shfl_up_sync(0xffffffffu, v.k1, d);
    r.out = __shfl_up_sync(0xffffffffu, v.out, d); r.nb = __shfl_up_sync(0xffffffffu, v.nb, d);
    return r;
}

// inclusive scan over the CTA; returns this thread's inclusive value, *total = the CTA's composition
template <int THREADS>
__device__ FaSum fa_block_scan(FaSum v, FaSum *warp_sums, FaSum *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        FaSum u = fa_shfl_up(v, d);
        if (lane >= d) v = fa_compose(u, v);
    }
    if (lane == 31) warp_sums[w] = v;
    __syncthreads();
    FaSum pre = fa_identity();
    for (int i = 0; i < w; i++) pre = fa_compose(pre, warp_sums[i]);
    v = fa_compose(pre, v);
    if (total) {
        FaSum t = fa_identity();
        for (int i = 0; i < THREADS / 32; i++) t = fa_compose(t, warp_sums[i]);
        *total = t;
    }
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(FA_THREADS)
k_fasta_tiles(const uint8_t *__restrict__ text, uint64_t n, FaSum *__restrict__ tiles, int *err) {
    __shared__ FaSum warp_sums[FA_THREADS / 32];
    const uint64_t i0 = ((uint64_t)blockIdx.x * FA_THREADS + threadIdx.x) * FA_BYTES;
    uint8_t c[FA_BYTES];
    fa_load(text, n, i0, c);
    FaSum total;
    fa_block_scan<FA_THREADS>(fa_thread_summary(text, n, i0, c, err), warp_sums, &total);
    if (threadIdx.x == 0) tiles[blockIdx.x] = total;
}

// One CTA: exclusive scan of the tile summaries from the state at the start of the text (HDR: text[0] is '>').
// carry[t] = {incoming state, records before, kept bytes before}; totals = {records, kept bytes}
constexpr int FC_THREADS = 1024;
__global__ void __launch_bounds__(FC_THREADS)
k_fasta_carry(const FaSum *__restrict__ tiles, uint32_t n_tiles, uint4 *__restrict__ carry, unsigned long long *totals) {
    __shared__ FaSum warp_sums[FC_THREADS / 32];
    const uint32_t per = (n_tiles + FC_THREADS - 1) / FC_THREADS;
    const uint32_t t0 = min(threadIdx.x * per, n_tiles), t1 = min(t0 + per, n_tiles);
    FaSum mine = fa_identity();
    for (uint32_t t = t0; t < t1; t++) mine = fa_compose(mine, tiles[t]);
    FaSum incl = fa_block_scan<FC_THREADS>(mine, warp_sums, nullptr);
    // exclusive value = everything before this thread's range, evaluated from state HDR
    FaSum excl;
    {
        FaSum up = fa_shfl_up(incl, 1);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) {
            up = fa_identity();
            for (int i = 0; i < w; i++) up = fa_compose(up, warp_sums[i]);
        }
        excl = up;
    }
    uint32_t state = excl.out & 1u;
    uint64_t kept = excl.k0;
    uint32_t rec = excl.nb;
    for (uint32_t t = t0; t < t1; t++) {
        carry[t] = make_uint4(state, rec, (uint32_t)kept, (uint32_t)(kept >> 32));
        const FaSum s = tiles[t];
        kept += state ? s.k1 : s.k0;
        rec += s.nb;
        state = (s.out >> state) & 1u;
    }
    if (t1 == n_tiles && t0 < t1) { totals[0] = rec; totals[1] = kept; }
    if (n_tiles == 0 && threadIdx.x == 0) { totals[0] = 0; totals[1] = 0; }
}

__global__ void __launch_bounds__(FA_THREADS)
k_fasta_emit(const uint8_t *__restrict__ text, uint64_t n, const uint4 *__restrict__ carry, uint64_t *__restrict__ rec_start,
             uint64_t *__restrict__ hdr_end, uint64_t *__restrict__ seq_off, uint8_t *__restrict__ seq, int *err) {
    __shared__ FaSum warp_sums[FA_THREADS / 32];
    __shared__ uint8_t stage[FA_TILE];
    const uint64_t i0 = ((uint64_t)blockIdx.x * FA_THREADS + threadIdx.x) * FA_BYTES;
    uint8_t c[FA_BYTES];
    fa_load(text, n, i0, c);
    const FaSum mine = fa_thread_summary(text, n, i0, c, err);
    FaSum total;
    const FaSum incl = fa_block_scan<FA_THREADS>(mine, warp_sums, &total);
    // exclusive prefix of this thread within the tile
    FaSum excl = fa_shfl_up(incl, 1);
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) {
            excl = fa_identity();
            for (int i = 0; i < w; i++) excl = fa_compose(excl, warp_sums[i]);
        }
    }
    const uint4 cr = carry[blockIdx.x];
    const uint32_t st_in = cr.x;
    const uint64_t tile_kept0 = (uint64_t)cr.z | ((uint64_t)cr.w << 32);
    uint32_t state = (excl.out >> st_in) & 1u;
    uint32_t local = (uint32_t)(st_in ? excl.k1 : excl.k0);  // kept bytes of the tile before this thread
    uint32_t rec = cr.y + excl.nb;                           // records opened before this thread
#pragma unroll
    for (int j = 0; j < FA_BYTES; j++) {
        const uint8_t ch = c[j];
        const uint64_t i = i0 + j;
        if (ch == '>' && i < n && fa_boundary(text, i, err)) {
            rec_start[rec] = i;
            seq_off[rec] = tile_kept0 + local;
            rec++;
            state = 0;
        } else if (ch == '\n') {
            if (state == 0 && rec > 0) hdr_end[rec - 1] = i;
            state = 1;
        } else if (!fa_space(ch) && state) {
            stage[local++] = (ch >= 'a' && ch <= 'z') ? (uint8_t)(ch - 32) : ch;  // bytes.ToUpper (ASCII)
        }
    }
    __syncthreads();
    const uint32_t tile_kept = (uint32_t)(st_in ? total.k1 : total.k0);
    for (uint32_t j = threadIdx.x; j < tile_kept; j += FA_THREADS) seq[tile_kept0 + j] = stage[j];
}

__device__ __forceinline__ bool fa_trim_space(uint8_t c) {  // strings.TrimSpace over ASCII
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

__global__ void k_fasta_headers(const uint8_t *__restrict__ text, uint64_t n, uint32_t n_rec, const uint64_t *__restrict__ rec_start,
                                const uint64_t *__restrict__ hdr_end, const uint64_t *__restrict__ seq_off, uint64_t total_kept,
                                double expr_mul, uint32_t *__restrict__ hdr_len, uint32_t *__restrict__ name_len,
                                uint64_t *__restrict__ level, uint64_t *__restrict__ seq_len, uint8_t *__restrict__ valid) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t b = rec_start[r] + 1;
    const uint64_t rec_end = r + 1 < n_rec ? rec_start[r + 1] : n;
    uint64_t e = hdr_end[r];
    if (e == ~0ull) e = rec_end;  // no newline in the record: everything is the name, the sequence is empty
    seq_len[r] = (r + 1 < n_rec ? seq_off[r + 1] : total_kept) - seq_off[r];
    hdr_len[r] = (uint32_t)min(e - b, (uint64_t)0xffffffffu);
    // input.go:84-88: exactly one '$', non-empty name
    uint64_t d = e;
    int dollars = 0;
    for (uint64_t i = b; i < e; i++)
        if (text[i] == '$') { if (dollars++ == 0) d = i; }
    bool ok = dollars == 1 && d > b;
    uint64_t v = 0;
    if (ok) {
        uint64_t i = d + 1;
        while (i < e && fa_trim_space(text[i])) i++;
        // input.go:92-93: Sscanf("%d") into a uint64: at least one digit, no sign, must fit 64 bits; the rest is ignored
        int digits = 0;
        while (i < e && text[i] >= '0' && text[i] <= '9') {
            const uint64_t dgt = text[i] - '0';
            if (v > (~0ull - dgt) / 10) { ok = false; break; }
            v = v * 10 + dgt;
            digits++; i++;
        }
        if (!digits) ok = false;
    }
    valid[r] = ok ? 1 : 0;
    name_len[r] = ok ? (uint32_t)(d - b) : 0;
    level[r] = ok ? (uint64_t)((double)v * expr_mul) : 0;  // input.go:113
}

// rare path: records with malformed names are dropped, the remaining sequences are packed again
__global__ void k_fasta_gather(const uint8_t *__restrict__ src, const uint64_t *__restrict__ from, const uint64_t *__restrict__ to,
                               uint32_t n, uint8_t *__restrict__ dst) {
    const uint32_t r = blockIdx.x;
    if (r >= n) return;
    const uint64_t a = from[r], b = to[r], len = to[r + 1] - b;
    for (uint64_t j = threadIdx.x; j < len; j += blockDim.x) dst[b + j] = src[a + j];
}

uint32_t fasta_tiles(uint64_t n) { return (uint32_t)((n + FA_TILE - 1) / FA_TILE); }
size_t fasta_tile_bytes() { return sizeof(FaSum); }

void launch_fasta_parse(cudaStream_t st, const uint8_t *text, uint64_t n, void *tiles, uint4 *carry, unsigned long long *totals,
                        uint64_t *rec_start, uint64_t *hdr_end, uint64_t *seq_off, uint8_t *seq, int *err, int pass) {
    const uint32_t nt = fasta_tiles(n);
    if (pass == 0) {
        if (nt) k_fasta_tiles<<<nt, FA_THREADS, 0, st>>>(text, n, (FaSum *)tiles, err);
        k_fasta_carry<<<1, FC_THREADS, 0, st>>>((const FaSum *)tiles, nt, carry, totals);
    } else if (nt) {
        k_fasta_emit<<<nt, FA_THREADS, 0, st>>>(text, n, carry, rec_start, hdr_end, seq_off, seq, err);
    }
}

void launch_fasta_headers(cudaStream_t st, const uint8_t *text, uint64_t n, uint32_t n_rec, const uint64_t *rec_start,
                          const uint64_t *hdr_end, const uint64_t *seq_off, uint64_t total_kept, double expr_mul,
                          uint32_t *hdr_len, uint32_t *name_len, uint64_t *level, uint64_t *seq_len, uint8_t *valid) {
    if (n_rec) k_fasta_headers<<<(n_rec + 127) / 128, 128, 0, st>>>(text, n, n_rec, rec_start, hdr_end, seq_off, total_kept, expr_mul,
                                                                 hdr_len, name_len, level, seq_len, valid);
}

void launch_fasta_gather(cudaStream_t st, const uint8_t *src, const uint64_t *from, const uint64_t *to, uint32_t n, uint8_t *dst) {
    if (n) k_fasta_gather<<<n, 128, 0, st>>>(src, from, to, n, dst);
}

}  // namespace rl

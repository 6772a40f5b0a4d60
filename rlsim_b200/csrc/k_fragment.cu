// Fragmentation kernels (kernel group 1b): one thread per molecule.
//
// Restates, per molecule, transcript.go:139-150 (poly-A tail + Fragment call), transcript.go:279-284
// (SimulatePolyA), fragmentor.go:95-176 (after_prim / after_prim_double / after_noprim / after_noprim_double),
// fragmentor.go:205-270 (pre_prim), fragmentor.go:303-351 (prim_jump), nnthermo.go:203-217 (SimulatePriming, as a
// binary search in the fixed-point prefix sums), fragfilter.go:45-50 (loss filter) and the AfterFrag / PolyALen
// histogram updates of fragstats.go:80,116.  Kept fragments are emitted as 64-bit keys
//     key = (global start position << len_bits) | length    (global start = off[t] + start; len_bits = bits of `high`)
// so that a radix sort of the keys orders them by (transcript, start, end) and run-length encoding yields the
// species table (transcript.go:156-215).
//
// Molecules whose Poisson draw asks for more than FRAG_LOCAL_BREAKS breakpoints are deferred to k_fragment_long
// (one CTA per molecule, breakpoints sorted in shared memory).  Every draw is addressed by
// (transcript, molecule, sub-stream, index), so both kernels produce the same fragments.
#include <algorithm>
#include "kernels.cuh"
#include "prefix.cuh"
#include "variates.cuh"

namespace rl {

constexpr int FRAG_THREADS = 128;
constexpr int FRAG_LOCAL_BREAKS = 48;

struct FragOut {
    uint64_t *keys;             // fragment keys
    uint64_t cap;
    uint32_t len_bits;          // key = (off[t] + start) << len_bits | length: as few sort passes as the model allows
    unsigned long long *count;  // number of keys written (may exceed cap: overflow => caller retries)
    unsigned long long *hist_frag;   // [high+1]
    unsigned long long *hist_polya;  // [polya_max+1]
    uint2 *long_list;           // deferred molecules (t, m_lo) + m_hi in long_hi
    uint32_t *long_hi;
    uint32_t long_cap;
    unsigned int *long_count;
    int *error;                 // RLSIM_E_* raised on the device
    // interval work queue (after_* methods): filled by k_fragment, drained by k_intervals
    uint4 *q_a;                 // (t, m_lo, interval index, start)
    uint2 *q_b;                 // (end, m_hi)
    uint64_t q_cap;
    unsigned long long *q_count;
};

__device__ __forceinline__ PxStream mol_stream(const DevModel &m, uint32_t tg, uint64_t mol, uint32_t sub) {
    return PxStream{m.key_frag, tg, (uint32_t)(mol & 0xffffffffu), (sub << 24) | (uint32_t)((mol >> 32) & 0xffffffu)};
}

// SimulatePriming, nnthermo.go:203-217, over the prefix sums P(0..proflen): the smallest i in [start,end) with
// P(i+1) > P(start) + u.  The answer is defined by the integers alone; HOW it is searched is free.  The weights of
// neighbouring k-mers vary by less than one order of magnitude, so P is close to linear inside a window: an
// interpolated guess in the coarse (per 64 entries) array picks the block, an interpolated guess inside the block's
// 64 u32 sums lands within a few entries, and a short walk finishes - ~3 dependent round trips where bisection of a
// flat u64 array needs ~11.
constexpr int PRIME_INTERP_STEPS = 4;
// smallest i in [start, end) with P(i+1) > tgt, given P(start) <= tgt < P(end); plen = number of profile entries held
__device__ __forceinline__ uint32_t prime_search(const Profile &pf, uint32_t plen, uint32_t start, uint32_t end, uint64_t tgt) {
    // ---- block b in [bs, be] with C[b] <= tgt < C[b+1]   (C[bs] <= P(start) <= tgt < P(end) <= C[be+1])
    uint32_t lo = start >> PFX_BLOCK_SHIFT, hi = ((end - 1) >> PFX_BLOCK_SHIFT) + 1;
    uint64_t clo = 0, chi = 0;  // = C[lo], C[hi] whenever `have`
    const bool have = hi - lo > 1;
    if (have) {
        clo = pf.C[lo]; chi = pf.C[hi];
#pragma unroll 1
        for (int it = 0; it < PRIME_INTERP_STEPS && hi - lo > 1; it++) {
            double f = (double)(tgt - clo) / (double)(chi - clo);
            uint32_t g = lo + (uint32_t)(f * (double)(hi - lo));
            if (g >= hi) g = hi - 1;
            uint64_t a = pf.C[g], b = pf.C[g + 1];
            if (tgt < a) { hi = g; chi = a; }
            else if (tgt >= b) { lo = g + 1; clo = b; }
            else { lo = g; hi = g + 1; clo = a; chi = b; }
        }
        while (hi - lo > 1) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            uint64_t c = pf.C[mid];
            if (c <= tgt) { lo = mid; clo = c; } else { hi = mid; chi = c; }
        }
    }
    const uint32_t b = lo;
    if (!have) { clo = pf.C[b]; chi = pf.C[b + 1]; }
    // ---- entry j of the block: smallest j with fine[64 b + j] > r
    const uint32_t r = (uint32_t)(tgt - clo);
    const uint32_t base = b << PFX_BLOCK_SHIFT;
    const uint32_t nb = min(PFX_BLOCK, plen - base);
    const uint32_t *F = pf.F + base;
    uint32_t j = (uint32_t)(((uint64_t)r * nb) / (chi - clo));  // chi - clo = block sum > r
    if (j >= nb) j = nb - 1;
    // walk from the guess four entries at a time (one 16-byte load per step; F + base is 256-byte aligned): c = how
    // many entries of the quad are <= r.  Entries at or beyond nb count as > r (F[nb-1] = block sum > r).
    uint32_t q = j & ~3u, c;
    auto quad = [&](uint32_t q0) {
        const uint4 v = *reinterpret_cast<const uint4 *>(F + q0);
        return (uint32_t)(v.x <= r) + (uint32_t)(q0 + 1 < nb && v.y <= r) + (uint32_t)(q0 + 2 < nb && v.z <= r) +
               (uint32_t)(q0 + 3 < nb && v.w <= r);
    };
    c = quad(q);
    if (c == 4) { do { q += 4; c = quad(q); } while (c == 4); }          // the sums are increasing: all four <= r, go up
    else if (c == 0) { while (q > 0) { q -= 4; c = quad(q); if (c) break; } }  // first entry already > r: go down
    return base + q + c;
}

// The same search on the record layout (common.cuh): interpolate over the records' coarse sums, then walk the half of the
// block that holds the target, recomputing weights from the record's own codes - no read beyond the block's record.
__device__ __forceinline__ uint32_t prime_search(const RProfile &pf, uint32_t plen, uint32_t start, uint32_t end, uint64_t tgt) {
    uint32_t lo = start >> PFX_BLOCK_SHIFT, hi = ((end - 1) >> PFX_BLOCK_SHIFT) + 1;  // coarse[lo] <= tgt < coarse[hi]
    if (hi - lo > 1) {
        uint64_t clo = pf.R[lo].coarse, chi = pf.R[hi].coarse;
#pragma unroll 1
        for (int it = 0; it < PRIME_INTERP_STEPS && hi - lo > 1; it++) {
            double f = (double)(tgt - clo) / (double)(chi - clo);
            uint32_t g = lo + (uint32_t)(f * (double)(hi - lo));
            if (g >= hi) g = hi - 1;
            uint64_t a = pf.R[g].coarse, b = pf.R[g + 1].coarse;
            if (tgt < a) { hi = g; chi = a; }
            else if (tgt >= b) { lo = g + 1; clo = b; }
            else { lo = g; hi = g + 1; clo = a; chi = b; }
        }
        while (hi - lo > 1) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            uint64_t c = pf.R[mid].coarse;
            if (c <= tgt) { lo = mid; clo = c; } else { hi = mid; chi = c; }
        }
    }
    const uint4 *q = reinterpret_cast<const uint4 *>(pf.R + lo);
    const uint4 u = q[0], v = q[1];
    uint32_t r = (uint32_t)(tgt - ((uint64_t)u.x | ((uint64_t)u.y << 32)));  // < block sum
    const bool up = r >= u.z;  // the target lies in the block's second half
    if (up) r -= u.z;
    const uint32_t w0 = up ? v.y : u.w, w1 = up ? v.z : v.x, w2 = up ? v.w : v.y;
    const uint64_t a = (uint64_t)w0 | ((uint64_t)w1 << 32), b = (uint64_t)w1 | ((uint64_t)w2 << 32);
    uint32_t acc = 0, j = 0;
    for (; j < 31; j++) {
        const uint64_t win = j < 16 ? a >> (2 * j) : b >> (2 * (j - 16));
        acc += pf.lut[(uint32_t)win & pf.kmask];
        if (acc > r) break;
    }
    return (lo << PFX_BLOCK_SHIFT) + (up ? 32u : 0u) + j;
}

// ---------------------------------------------------------------------------------------------------------------------
// Compact profile: what one WARP keeps in shared memory for one transcript (k_fused_warp).  0.625 B per entry instead of
// the 4.1 B of the two-level arrays, so that 40 warps per SM each hold their own transcript:
//   codes  : the 2-bit packed bases themselves (16 per word, copied from the transcript store)
//   gsum   : per group of 16 entries, the sum of the weights from the start of its 64-entry block to the group's end
//   coarse : per 64-entry block, the sum of everything before it (coarse[nblk] = total)
// A prefix sum P(i) is coarse + group sum + at most 15 weights recomputed from the codes and the shared-memory LUT; the
// search walks one group.  Same integers as the materialised sums, hence the same draws.  Transcripts with letters
// outside ACGT do not take this path (their windows need the byte-level affinity and explicit reverse sums).
struct CProfile {
    const uint32_t *codes, *gsum;
    const uint64_t *coarse;
    const uint32_t *lut;
    uint32_t kmask;
    // weights of the entries [16 grp, 16 grp + n) summed; n <= 16
    __device__ __forceinline__ uint32_t partial(uint32_t grp, uint32_t n) const {
        if (!n) return 0;
        const uint64_t win = (uint64_t)codes[grp] | ((uint64_t)codes[grp + 1] << 32);  // 32 bases from 16 grp: 16 windows of k <= 17
        uint32_t acc = 0;
#pragma unroll 4
        for (uint32_t j = 0; j < n; j++) acc += lut[(uint32_t)(win >> (2 * j)) & kmask];
        return acc;
    }
    __device__ __forceinline__ uint64_t P(uint32_t i) const {
        const uint32_t grp = i >> 4;
        uint64_t c = coarse[i >> PFX_BLOCK_SHIFT];
        if (grp & 3u) c += gsum[grp - 1];
        return c + partial(grp, i & 15u);
    }
};

// smallest i in [start, end) with P(i+1) > tgt, given P(start) <= tgt < P(end)
__device__ __forceinline__ uint32_t prime_search(const CProfile &pf, uint32_t plen, uint32_t start, uint32_t end, uint64_t tgt) {
    uint32_t lo = start >> PFX_BLOCK_SHIFT, hi = ((end - 1) >> PFX_BLOCK_SHIFT) + 1;  // coarse[lo] <= tgt < coarse[hi]
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (pf.coarse[mid] <= tgt) lo = mid; else hi = mid;
    }
    const uint32_t r = (uint32_t)(tgt - pf.coarse[lo]);  // < block sum
    uint32_t grp = lo << 2, before = 0;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        const uint32_t s = pf.gsum[(lo << 2) + g];
        if (s <= r) { grp = (lo << 2) + g + 1; before = s; }  // group sums are nondecreasing inside the block
    }
    const uint32_t rem = r - before;
    const uint64_t win = (uint64_t)pf.codes[grp] | ((uint64_t)pf.codes[grp + 1] << 32);
    uint32_t acc = 0, j = 0;
    for (; j < 15; j++) {
        acc += pf.lut[(uint32_t)(win >> (2 * j)) & pf.kmask];
        if (acc > rem) break;
    }
    return (grp << 4) + j;
}

template <class PF>
__device__ __forceinline__ bool priming(const PxStream &iv, uint32_t block, const PF &pf, uint32_t proflen,
                                        uint32_t start, uint32_t end, uint32_t &res) {
    if (end > proflen) end = proflen;
    if (start >= end) { res = 0; return true; }
    const uint64_t ps = pf.P(start), pe = pf.P(end);
    const uint64_t total = pe - ps;
    if (total == 0) return false;
    res = prime_search(pf, proflen, start, end, ps + iv.below(block, total));
    return true;
}

// The same draw on the reverse profile of a transcript without non-ACGT letters when the quantised k-mer table is
// reverse-complement symmetric: reverse entry j is forward window n - j (n = proflen), so with W = forward prefix sums
// over the n + 1 windows, P_rev(j) = W(n+1) - W(n+1-j).  "Smallest j in [s, e) with P_rev(j+1) > P_rev(s) + u" becomes
// "smallest x in [n-e+1, n-s+1) with W(x+1) > W(n-s+1) - u - 1", j = n - x: no reverse arrays are needed.
template <class PF>
__device__ __forceinline__ bool priming_reverse(const PxStream &iv, uint32_t block, const PF &pfwd, const PF &prev,
                                                bool mirror, uint32_t proflen, uint32_t start, uint32_t end, uint32_t &res) {
    if (end > proflen) end = proflen;
    if (start >= end) { res = 0; return true; }
    // one search serves both representations (a single inlined copy keeps the register count of k_intervals down)
    const PF pf = mirror ? pfwd : prev;
    const uint32_t plen = mirror ? proflen + 1 : proflen;
    const uint32_t lo = mirror ? proflen - end + 1 : start, hi = mirror ? proflen - start + 1 : end;
    const uint64_t plo = pf.P(lo), phi = pf.P(hi);
    const uint64_t total = phi - plo;
    if (total == 0) return false;
    const uint64_t u = iv.below(block, total);
    const uint32_t x = prime_search(pf, plen, lo, hi, mirror ? phi - u - 1 : plo + u);
    res = mirror ? proflen - x : x;
    return true;
}

__device__ __forceinline__ bool keep_fragment(const DevModel &m, const PxStream &iv, uint32_t i) {  // fragfilter.go:45-50
    if (m.loss <= 0.0) return true;  // u >= 0 always: the draw cannot change the outcome
    PxBlock b = iv.block(4 * i + 2);
    return !(px_u01(px_lo64(b)) < m.loss);
}

// Emit one fragment: warp-aggregated append + histogram
__device__ __forceinline__ void emit(const FragOut &o, uint64_t gstart, uint32_t size, unsigned int *hist_s,
                                     uint32_t hist_s_n, uint32_t low) {
    unsigned mask = __activemask();
    int leader = __ffs(mask) - 1;
    int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(o.count, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    unsigned long long pos = base + __popc(mask & ((1u << lane) - 1u));
    if (pos < o.cap) o.keys[pos] = (gstart << o.len_bits) | (uint64_t)size;
    uint32_t bin = size - low;
    if (bin < hist_s_n) atomicAdd(&hist_s[bin], 1u);
    else atomicAdd(&o.hist_frag[size], 1ull);
}

// Queue one interval that passed the `end - start >= low` pre-filter (fragmentor.go:136-138) for k_intervals.
__device__ __forceinline__ void enqueue(const FragOut &o, uint32_t t, uint64_t mol, uint32_t i, uint32_t start, uint32_t end) {
    unsigned mask = __activemask();
    int leader = __ffs(mask) - 1;
    int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(o.q_count, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    unsigned long long pos = base + __popc(mask & ((1u << lane) - 1u));
    if (pos < o.q_cap) {
        o.q_a[pos] = make_uint4(t, (uint32_t)(mol & 0xffffffffu), i, start);
        o.q_b[pos] = make_uint2(end, (uint32_t)(mol >> 32));
    }
}


template <class PF>
struct MolCtxT {
    uint32_t t, tg, L, proflen;
    uint64_t off;
    PF Pf, Pr;
    bool mirror;  // reverse profile = forward arrays read backwards (symmetric table, no non-ACGT letter)
};
using MolCtx = MolCtxT<Profile>;

// One interval of the after_* methods, fragmentor.go:132-175
template <class PF>
__device__ __forceinline__ void after_interval(const DevModel &m, const MolCtxT<PF> &c, const PxStream &iv, uint32_t i,
                                               uint32_t start, uint32_t end, const FragOut &o, unsigned int *hist_s,
                                               uint32_t hist_s_n) {
    if (end - start < m.low) return;
    const bool simPriming = (m.method == RLSIM_AFTER_PRIM || m.method == RLSIM_AFTER_PRIM_DOUBLE);
    const bool doublePrime = (m.method == RLSIM_AFTER_PRIM_DOUBLE || m.method == RLSIM_AFTER_NOPRIM_DOUBLE);
    uint32_t new_start;
    if (simPriming) {
        if (!priming(iv, 4 * i, c.Pf, c.proflen, start, end, new_start)) { atomicExch(o.error, RLSIM_E_PRIMING); return; }
        if (doublePrime) {
            uint32_t final = c.L - 1;
            uint32_t ss = final - end + 1, ee = final - new_start + 1, ns;
            if (!priming_reverse(iv, 4 * i + 1, c.Pf, c.Pr, c.mirror, c.proflen, ss, ee, ns)) { atomicExch(o.error, RLSIM_E_PRIMING); return; }
            end = final - ns + 1;
        }
    } else {
        new_start = start + (uint32_t)iv.below(4 * i, end - start);
        if (doublePrime) end = new_start + (uint32_t)iv.below(4 * i + 1, end - new_start);
    }
    uint32_t size = end - new_start;
    if (size < m.low || size > m.high) return;
    if (!keep_fragment(m, iv, i)) return;
    emit(o, c.off + new_start, size, hist_s, hist_s_n, m.low);
}

__device__ __forceinline__ void plain_interval(const DevModel &m, const MolCtx &c, const PxStream &iv, uint32_t i,
                                               uint32_t start, uint32_t end, const FragOut &o, unsigned int *hist_s,
                                               uint32_t hist_s_n) {  // fragmentor.go:252-269
    uint32_t size = end - start;
    if (size < m.low || size > m.high) return;
    if (!keep_fragment(m, iv, i)) return;
    emit(o, c.off + start, size, hist_s, hist_s_n, m.low);
}

// Header draws shared by both kernels: poly-A tail, mixture component, (pre_prim: elongation), Poisson.
struct Header {
    int polyAend;
    uint32_t tail;
    int32_t nb;       // number of random breakpoints (>= 1), 0 => molecule yields nothing
    uint32_t elong;   // pre_prim: first breakpoint
};

template <class Cur> __device__ __forceinline__ double target_location(const DevModel &m, Cur &hdr) {
    if (m.n_target == 0) return m.raw_location;
    return m.target[px_mix_comp(hdr, m.target, m.n_target)].location;
}

// KIND: 0 = after_* methods, 1 = pre_prim, 2 = prim_jump (compile-time specialisations of the molecule kernel), -1 = runtime
template <int KIND, class CTX>
__device__ __forceinline__ Header mol_header(const DevModel &m, const CTX &c, uint64_t mol) {
    const bool is_jump = KIND < 0 ? m.method == RLSIM_PRIM_JUMP : KIND == 2;
    const bool is_pre = KIND < 0 ? m.method == RLSIM_PRE_PRIM : KIND == 1;
    Header h;
    PxCursor3 hdr(mol_stream(m, c.tg, mol, SUB_HEADER));  // first six header draws generated convergently
    h.tail = px_comp_len(hdr, m.polya[px_mix_comp(hdr, m.polya, m.n_polya)]);
    h.polyAend = (int)c.L - (int)m.polya_max + (int)h.tail;  // transcript.go:283
    h.nb = 0;
    h.elong = 0;
    if (is_jump) return h;
    int length = h.polyAend;
    if (length < 1) return h;
    double loc = target_location(m, hdr);
    double rate;
    if (is_pre) {
        double fp = 1.0 / (double)(uint32_t)m.frag_param;
        double ex = px_expo(hdr) / fp;
        int el = ex >= (double)length ? length : (int)ex;
        h.elong = (uint32_t)(length - el);
        rate = (double)length / loc;
    } else {
        rate = m.frag_param == 0 ? ((double)length / loc) / 2.0 : (double)length / (double)(uint32_t)m.frag_param;
    }
    int32_t nb = px_poisson(hdr, rate) - 1;  // fragmentor.go:113-117
    if (nb <= 0) nb = 1;
    h.nb = nb;
    if (is_pre && (int64_t)length - (int64_t)h.elong < 1) h.nb = 0;  // fragmentor.go:240-243
    return h;
}

template <class PF>
__device__ void prim_jump(const DevModel &m, const MolCtxT<PF> &c, uint64_t mol, int polyAend, const FragOut &o,
                          unsigned int *hist_s, uint32_t hist_s_n) {  // fragmentor.go:303-351
    int ltmp = (int)c.L - (int)c.proflen;
    if (polyAend < ltmp) return;
    uint32_t final = (uint32_t)(polyAend - ltmp);
    double fp = m.frag_param != 0 ? 1.0 / (double)(uint32_t)m.frag_param : 0.0;
    PxStream iv = mol_stream(m, c.tg, mol, SUB_INTERVAL);
    PxCursor jl(mol_stream(m, c.tg, mol, SUB_JUMPLEN));
    uint32_t start = 0, end = 0, j = 0;
    while (end < final) {
        if (!priming(iv, 4 * j, c.Pf, c.proflen, end, final, start)) { atomicExch(o.error, RLSIM_E_PRIMING); return; }
        uint64_t new_end;
        if (fp != 0.0) {
            double ex = px_expo(jl) / fp;
            new_end = (uint64_t)start + (ex >= 4294967296.0 ? 4294967295ull : (uint64_t)ex);
        } else {
            uint32_t l = m.n_target == 0 ? px_raw_len(jl, m) : px_comp_len(jl, m.target[px_mix_comp(jl, m.target, m.n_target)]);
            new_end = (uint64_t)start + l;
        }
        if (new_end > final) new_end = final;
        end = (uint32_t)new_end;
        uint32_t size = end - start;
        uint32_t jj = j++;
        if (size < m.low || size > m.high) continue;
        if (!keep_fragment(m, iv, jj)) continue;
        emit(o, c.off + start, size, hist_s, hist_s_n, m.low);
    }
}

__device__ __forceinline__ MolCtx make_ctx(const DevModel &m, const DevTranscripts &tr, uint32_t t) {
    MolCtx c;
    c.t = t;
    c.tg = m.first_index + t;
    c.L = tr.len[t];
    c.proflen = c.L > (uint32_t)m.k ? c.L - (uint32_t)m.k : 0;
    c.off = tr.off[t];
    c.Pf = Profile{tr.Ff ? tr.Ff + c.off : nullptr, tr.Cf ? tr.Cf + coarse_base(c.off, t) : nullptr};
    c.Pr = Profile{tr.Fr ? tr.Fr + c.off : nullptr, tr.Cr ? tr.Cr + coarse_base(c.off, t) : nullptr};
    c.mirror = tr.sym && tr.Fr && tr.dirty[t] == 0;
    return c;
}
// the record layout serves the transcripts without letters outside ACGT (forward sums; the reverse draw reads them mirrored)
__device__ __forceinline__ bool uses_records(const DevTranscripts &tr, uint32_t t) { return tr.Rf != nullptr && tr.dirty[t] == 0; }
__device__ __forceinline__ MolCtxT<RProfile> make_ctx_rec(const DevModel &m, const DevTranscripts &tr, uint32_t t, const uint32_t *lut) {
    MolCtxT<RProfile> c;
    c.t = t;
    c.tg = m.first_index + t;
    c.L = tr.len[t];
    c.proflen = c.L > (uint32_t)m.k ? c.L - (uint32_t)m.k : 0;
    c.off = tr.off[t];
    c.Pf = RProfile{tr.Rf + rec_base(c.off, t), lut, (1u << (2 * m.k)) - 1u};
    c.Pr = c.Pf;
    c.mirror = true;  // records exist only with a symmetric table
    return c;
}

// warp_t[w] = transcript of the first molecule of warp w of k_fragment (molecule g_begin + 32 w): one parallel bisection per
// warp here instead of a one-lane bisection at the head of every warp there (10 % of k_fragment's samples)
__global__ void k_warp_transcripts(DevTranscripts tr, uint64_t g_begin, uint64_t g_end, uint32_t *__restrict__ warp_t) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t g = g_begin + w * 32;
    if (g < g_end) warp_t[w] = upper_bound_u64(tr.mol_off, tr.n + 1, g) - 1;
}
void launch_warp_transcripts(cudaStream_t st, const DevTranscripts &tr, uint64_t g_begin, uint64_t g_end, uint32_t *warp_t) {
    if (g_end <= g_begin) return;
    const uint64_t nw = (g_end - g_begin + 31) / 32;
    k_warp_transcripts<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(tr, g_begin, g_end, warp_t);
}

template <int KIND>
__global__ void __launch_bounds__(FRAG_THREADS, 10)
k_fragment(const __grid_constant__ DevModel m, DevTranscripts tr, uint64_t g_begin, uint64_t g_end, FragOut o, uint32_t hist_s_n,
           const uint32_t *__restrict__ warp_t) {
    extern __shared__ unsigned int hist_s[];  // [hist_s_n] after-frag bins from `low`, then [polya_max+1]
    unsigned int *hist_pa = hist_s + hist_s_n;
    for (uint32_t i = threadIdx.x; i < hist_s_n + m.polya_max + 1; i += FRAG_THREADS) hist_s[i] = 0;
    __syncthreads();
    uint64_t g = g_begin + (uint64_t)blockIdx.x * FRAG_THREADS + threadIdx.x;  // global molecule index
    const unsigned in_range = __ballot_sync(0xffffffffu, g < g_end);  // taken while the warp is still converged
    if (g < g_end) {
        // molecule -> transcript: the warp's first lane searches; a warp rarely spans more than a few transcripts.
        // (measured: a 32-ary search by the whole warp and a tile loop per CTA are both slower here)
        const uint64_t g0 = g - (threadIdx.x & 31);
        uint32_t t = 0;
        if (warp_t) t = warp_t[(g0 - g_begin) >> 5];  // precomputed (k_warp_transcripts)
        else {
            if ((threadIdx.x & 31) == 0) t = upper_bound_u64(tr.mol_off, tr.n + 1, g0) - 1;
            t = __shfl_sync(in_range, t, 0);  // lane 0 holds the smallest g of the warp: it is in range whenever any lane is
        }
        while (tr.mol_off[t + 1] <= g) t++;
        uint64_t mol = g - tr.mol_off[t];
        MolCtx c = make_ctx(m, tr, t);
        Header h = mol_header<KIND>(m, c, mol);
        atomicAdd(&hist_pa[h.tail], 1u);
        if (KIND == 2) {
            if (uses_records(tr, t)) prim_jump(m, make_ctx_rec(m, tr, t, tr.lut), mol, h.polyAend, o, hist_s, hist_s_n);
            else prim_jump(m, c, mol, h.polyAend, o, hist_s, hist_s_n);
        } else if (h.nb > FRAG_LOCAL_BREAKS) {
            unsigned int slot = atomicAdd(o.long_count, 1u);
            if (slot < o.long_cap) { o.long_list[slot] = make_uint2(t, (uint32_t)(mol & 0xffffffffu)); o.long_hi[slot] = (uint32_t)(mol >> 32); }
        } else if (h.nb > 0) {
            uint32_t length = (uint32_t)h.polyAend;
            uint32_t br[FRAG_LOCAL_BREAKS + 2];
            PxStream bs = mol_stream(m, c.tg, mol, SUB_BREAKS);
            uint32_t first = KIND == 1 ? h.elong : 0u;
            uint64_t span = (uint64_t)length - first;
            br[0] = first;
            // insertion sort while drawing (fragmentor.go:125-129)
            int cnt = 1;
            for (int32_t j = 0; j < h.nb; j++) {
                uint32_t v = first + (uint32_t)bs.below((uint32_t)j, span);
                int q = cnt++;
                while (q > 0 && br[q - 1] > v) { br[q] = br[q - 1]; q--; }
                br[q] = v;
            }
            br[cnt] = length;  // length is >= every draw, so it sorts last
            PxStream iv = mol_stream(m, c.tg, mol, SUB_INTERVAL);
            uint32_t n_queued = 0;
            for (int32_t i = 0; i < h.nb; i++) {  // the last interval is never visited (fragmentor.go:132)
                if (KIND == 1) plain_interval(m, c, iv, (uint32_t)i, br[i], br[i + 1], o, hist_s, hist_s_n);
                else if (br[i + 1] - br[i] >= m.low) { enqueue(o, t, mol, (uint32_t)i, br[i], br[i + 1]); n_queued++; }
            }
            if (n_queued && tr.Rf != nullptr && tr.dirty[t] != 0) atomicAdd(o.q_count + 1, (unsigned long long)n_queued);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < hist_s_n; i += FRAG_THREADS)
        if (hist_s[i]) atomicAdd(&o.hist_frag[m.low + i], (unsigned long long)hist_s[i]);
    for (uint32_t i = threadIdx.x; i <= m.polya_max; i += FRAG_THREADS)
        if (hist_pa[i]) atomicAdd(&o.hist_polya[i], (unsigned long long)hist_pa[i]);
}

// One thread per queued interval: the two dependent priming searches (nnthermo.go:203-217) + filters + emit.
// Kept separate from k_fragment so that this latency-bound part runs lean (few registers, full occupancy).
constexpr int IV_THREADS = 256;
// REC = 1: intervals of transcripts on the record layout (k-mer LUT staged in shared memory); REC = 0: the two-level arrays
// (all transcripts when there are no records, else only those with letters outside ACGT).
template <int REC>
__global__ void __launch_bounds__(IV_THREADS, 5)
k_intervals(const __grid_constant__ DevModel m, DevTranscripts tr, FragOut o, uint64_t n_q, uint32_t hist_s_n,
            const unsigned long long *__restrict__ n_q_dev) {
    extern __shared__ unsigned int hist_s[];
    if (n_q_dev) {  // asynchronous chunk pipeline: the queue length never visited the host
        n_q = min((uint64_t)*n_q_dev, o.q_cap);
        if ((uint64_t)blockIdx.x * IV_THREADS >= n_q) return;
    }
    for (uint32_t i = threadIdx.x; i < hist_s_n; i += IV_THREADS) hist_s[i] = 0;
    uint32_t *lut_s = hist_s + hist_s_n;
    if (REC) for (uint32_t i = threadIdx.x; i < (1u << (2 * m.k)); i += IV_THREADS) lut_s[i] = tr.lut[i];
    __syncthreads();
    for (uint64_t g = (uint64_t)blockIdx.x * IV_THREADS + threadIdx.x; g < n_q; g += (uint64_t)gridDim.x * IV_THREADS) {
        uint4 a = o.q_a[g];
        if (tr.Rf != nullptr && (tr.dirty[a.x] == 0) != (REC != 0)) continue;  // the other launch's interval
        uint2 b = o.q_b[g];
        uint64_t mol = (uint64_t)a.y | ((uint64_t)b.y << 32);
        PxStream iv = mol_stream(m, m.first_index + a.x, mol, SUB_INTERVAL);
        if (REC) after_interval(m, make_ctx_rec(m, tr, a.x, lut_s), iv, a.z, a.w, b.x, o, hist_s, hist_s_n);
        else after_interval(m, make_ctx(m, tr, a.x), iv, a.z, a.w, b.x, o, hist_s, hist_s_n);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < hist_s_n; i += IV_THREADS)
        if (hist_s[i]) atomicAdd(&o.hist_frag[m.low + i], (unsigned long long)hist_s[i]);
}

// One CTA per deferred molecule: breakpoints drawn in parallel, bitonic-sorted in shared memory, intervals in parallel.
constexpr int LONG_THREADS = 256;
__global__ void __launch_bounds__(LONG_THREADS)
k_fragment_long(const __grid_constant__ DevModel m, DevTranscripts tr, FragOut o, uint32_t n_long, uint32_t smem_cap,
                const unsigned int *__restrict__ n_long_dev) {
    extern __shared__ uint32_t brs[];  // smem_cap entries (power of two)
    __shared__ Header hs;
    if (n_long_dev) n_long = min(*n_long_dev, n_long);  // asynchronous chunk pipeline: the count never visited the host
  for (uint32_t slot = blockIdx.x; slot < n_long; slot += gridDim.x) {
    __syncthreads();  // the previous molecule of this CTA is done with hs / brs
    uint32_t t = o.long_list[slot].x;
    uint64_t mol = (uint64_t)o.long_list[slot].y | ((uint64_t)o.long_hi[slot] << 32);
    MolCtx c = make_ctx(m, tr, t);
    if (threadIdx.x == 0) hs = mol_header<-1>(m, c, mol);
    __syncthreads();
    Header h = hs;
    uint32_t n = (uint32_t)h.nb + 2;
    uint32_t npow = 1;
    while (npow < n) npow <<= 1;
    if (npow > smem_cap) { if (threadIdx.x == 0) atomicExch(o.error, RLSIM_E_UNSUPPORTED); continue; }
    uint32_t length = (uint32_t)h.polyAend;
    uint32_t first = m.method == RLSIM_PRE_PRIM ? h.elong : 0u;
    uint64_t span = (uint64_t)length - first;
    PxStream bs = mol_stream(m, c.tg, mol, SUB_BREAKS);
    for (uint32_t j = threadIdx.x; j < npow; j += LONG_THREADS) {
        uint32_t v;
        if (j < (uint32_t)h.nb) v = first + (uint32_t)bs.below(j, span);
        else if (j == (uint32_t)h.nb) v = first;
        else if (j == (uint32_t)h.nb + 1) v = length;
        else v = 0xffffffffu;
        brs[j] = v;
    }
    __syncthreads();
    for (uint32_t k2 = 2; k2 <= npow; k2 <<= 1) {
        for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (uint32_t i = threadIdx.x; i < npow; i += LONG_THREADS) {
                uint32_t ixj = i ^ j2;
                if (ixj > i) {
                    uint32_t a = brs[i], b = brs[ixj];
                    bool up = ((i & k2) == 0);
                    if ((a > b) == up) { brs[i] = b; brs[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    PxStream iv = mol_stream(m, c.tg, mol, SUB_INTERVAL);
    const bool rec = uses_records(tr, t);
    const MolCtxT<RProfile> cr = rec ? make_ctx_rec(m, tr, t, tr.lut) : MolCtxT<RProfile>{};
    for (uint32_t i = threadIdx.x; i < (uint32_t)h.nb; i += LONG_THREADS) {
        if (m.method == RLSIM_PRE_PRIM) plain_interval(m, c, iv, i, brs[i], brs[i + 1], o, nullptr, 0);
        else if (rec) after_interval(m, cr, iv, i, brs[i], brs[i + 1], o, nullptr, 0);
        else after_interval(m, c, iv, i, brs[i], brs[i + 1], o, nullptr, 0);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused path (after_prim / after_prim_double): priming prefix sums live in SHARED memory, never in HBM.
//
// A work item is a run of consecutive transcripts (+ a molecule range) whose prefix sums fit one CTA's shared memory.
// The CTA (persistent, items fetched from a global counter) builds the two-level prefix sums of the item's transcripts
// from the 2-bit packed bases (0.25 B/nt of HBM reads instead of 4.1 B/nt written and ~0.5 KB per interval read back at
// random), then fragments the item's molecules with every priming search served from shared memory.  The draws are
// addressed by (transcript, molecule, sub-stream, index) and the search result is defined by the integers alone, so the
// fragments are those of k_prefix + k_fragment + k_intervals bit for bit.
//
// Shared-memory layout (dynamic): [fine u32 x cap_entries][coarse u64 x (cap_entries/64 + 2*FUSED_MAX_TR)][LUT u32 x 4^k]
// [histograms u32].  Inside it every profile (transcript, strand) owns a 64-entry aligned run of fine sums and
// nblk + 1 coarse sums, exactly the per-transcript layout of the HBM arrays (common.cuh), so prime_search is shared.
constexpr int FUSED_THREADS = 512;
constexpr int FUSED_WARPS = FUSED_THREADS / 32;
constexpr int FUSED_MAX_TR = 32;       // transcripts per sub-batch
constexpr int FUSED_SEG = 1024;        // entries per prefix task (one warp, four sub-tiles)

struct FusedBatch {                    // the sub-batch being processed (shared memory)
    uint32_t n, t_next;                // transcripts in it; first transcript of the next sub-batch
    uint32_t nprof, ntask;
    int fail;
    uint32_t t[FUSED_MAX_TR];          // transcript index
    uint32_t L[FUSED_MAX_TR];
    uint32_t fo[FUSED_MAX_TR], co[FUSED_MAX_TR];   // forward arrays: fine offset (entries), coarse offset (u64 slots)
    uint32_t ro[FUSED_MAX_TR], rco[FUSED_MAX_TR];  // reverse arrays (0xffffffff: none / mirrored)
    uint32_t mirror[FUSED_MAX_TR];
    uint64_t off[FUSED_MAX_TR];
    uint64_t mol[FUSED_MAX_TR + 1];    // mol_off of the transcripts (mol[n] = end of the last one)
    // profiles = (transcript slot, strand): prefix tasks are enumerated over them
    uint32_t p_slot[2 * FUSED_MAX_TR], p_rev[2 * FUSED_MAX_TR], p_ent[2 * FUSED_MAX_TR], p_task0[2 * FUSED_MAX_TR + 1];
};

__device__ __forceinline__ uint32_t pad64(uint32_t x) { return (x + 63u) & ~63u; }

// Warp 0 takes transcripts from t_cur while their prefix sums fit `cap` entries: lane i looks at transcript t_cur + i, warp
// scans give every taken transcript its offsets.  (A serial loop here costs more than the prefix sums themselves.)
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += u;
    }
    return v;
}
__device__ void fused_pack(FusedBatch &B, const DevModel &m, const DevTranscripts &tr, uint32_t t_cur, uint32_t t_end,
                           uint32_t cap, bool dbl, uint32_t lane) {
    const uint32_t t = t_cur + lane, k = (uint32_t)m.k;
    const bool in = t < t_end;
    const bool expressed = in && tr.expr[t] != 0;  // pool.go:104: the others are never fragmented
    uint32_t L = 0, entF = 0, entR = 0;
    bool mirror = false;
    if (expressed) {
        L = tr.len[t];
        mirror = dbl && tr.sym && tr.dirty[t] == 0;
        // forward arrays hold one more window in symmetric mode (the first window of the reverse profile)
        entF = L >= k + (tr.sym ? 0u : 1u) ? L - k + (tr.sym ? 1u : 0u) : 0u;
        entR = (dbl && !mirror && L > k) ? L - k : 0u;
    }
    const bool has_rev = expressed && dbl && !mirror;
    const uint32_t eF = pad64(entF), eR = pad64(entR), need = expressed ? eF + eR : 0u;
    const uint32_t pre = warp_incl_scan(need, lane);
    // the batch ends before the first expressed transcript whose sums no longer fit
    const unsigned over = __ballot_sync(0xffffffffu, expressed && pre > cap);
    const uint32_t cut = over ? (uint32_t)__ffs(over) - 1u : 32u;
    const bool take = expressed && lane < cut;
    const unsigned takes = __ballot_sync(0xffffffffu, take);
    const uint32_t n = __popc(takes), slot = __popc(takes & ((1u << lane) - 1u));
    const uint32_t cs = take ? eF / 64 + 1 + (has_rev ? eR / 64 + 1 : 0u) : 0u;
    const uint32_t cpre = warp_incl_scan(cs, lane);
    const uint32_t np_l = take ? (has_rev ? 2u : 1u) : 0u;
    const uint32_t ppre = warp_incl_scan(np_l, lane);
    const uint32_t tf = (entF + FUSED_SEG - 1) / FUSED_SEG, trv = (entR + FUSED_SEG - 1) / FUSED_SEG;
    const uint32_t nt_l = take ? tf + (has_rev ? trv : 0u) : 0u;
    const uint32_t tpre = warp_incl_scan(nt_l, lane);
    if (take) {
        const uint32_t fo = pre - need, co = cpre - cs, p0 = ppre - np_l, k0 = tpre - nt_l;
        B.t[slot] = t; B.L[slot] = L; B.off[slot] = tr.off[t]; B.mol[slot] = tr.mol_off[t]; B.mirror[slot] = mirror ? 1u : 0u;
        B.fo[slot] = fo; B.co[slot] = co;
        B.p_slot[p0] = slot; B.p_rev[p0] = 0; B.p_ent[p0] = entF; B.p_task0[p0] = k0;
        if (has_rev) {
            B.ro[slot] = fo + eF; B.rco[slot] = co + eF / 64 + 1;
            B.p_slot[p0 + 1] = slot; B.p_rev[p0 + 1] = 1; B.p_ent[p0 + 1] = entR; B.p_task0[p0 + 1] = k0 + tf;
        } else { B.ro[slot] = 0xffffffffu; B.rco[slot] = 0xffffffffu; }
        if (slot + 1 == n) B.mol[n] = tr.mol_off[t + 1];
    }
    if (lane == 31) { B.nprof = ppre; B.ntask = tpre; B.p_task0[ppre] = tpre; }
    if (lane == 0) {
        B.n = n;
        B.t_next = min(t_end, t_cur + cut);
        // nothing taken although an expressed transcript is waiting: it does not fit alone - the host reruns on the HBM path
        B.fail = (n == 0 && over != 0u) ? 1 : 0;
    }
}

// pass 1 (all warps): in-block sums of every 1024-entry segment + raw block totals into coarse[b + 1]
// pass 2 (one warp per profile): coarse[] <- running sums of the block totals.  Both separated by CTA barriers (caller).
__device__ void fused_prefix_pass1(const FusedBatch &B, const DevModel &m, const DevTranscripts &tr, uint32_t *fine_s,
                                   uint64_t *coarse_s, const uint32_t *lut_f, const uint32_t *lut_r, double scale,
                                   uint32_t warp, uint32_t nwarps, uint32_t lane) {
    const uint32_t kmask = (1u << (2 * m.k)) - 1u, mmask = (1u << m.k) - 1u;
    for (uint32_t task = warp; task < B.ntask; task += nwarps) {
        uint32_t p = 0;
        while (B.p_task0[p + 1] <= task) p++;
        const uint32_t slot = B.p_slot[p], rev = B.p_rev[p], ent = B.p_ent[p];
        const uint32_t seg0 = (task - B.p_task0[p]) * FUSED_SEG;
        uint32_t *F = fine_s + (rev ? B.ro[slot] : B.fo[slot]);
        uint64_t *C = coarse_s + (rev ? B.rco[slot] : B.co[slot]);
        const PfxCtx pc{B.off[slot], B.L[slot], ent, m.k, (int)rev, m.T, scale, rev ? lut_r : lut_f, kmask, mmask};
        const uint32_t seg1 = min(ent, seg0 + FUSED_SEG);
        PfxWords nxt{0, 0, 0, 0, 0};
        uint32_t ib = seg0 + lane * PFX_ITEMS;
        if (ib < ent) nxt = pfx_load(tr, pc.o, pfx_jb(rev, pc.L, m.k, ent, ib));
        for (uint32_t i0 = seg0; i0 < seg1; i0 += PFX_SUB) {
            ib = i0 + lane * PFX_ITEMS;
            const PfxWords w5 = nxt;
            const uint32_t ibn = ib + PFX_SUB;
            if (ibn < seg1 && ibn < ent) nxt = pfx_load(tr, pc.o, pfx_jb(rev, pc.L, m.k, ent, ibn));
            uint32_t x[PFX_ITEMS];
            const uint32_t run = pfx_lane_sums(tr, pc, ib, w5, x);
            uint32_t incl = run;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d, 8);
                if ((lane & 7) >= (uint32_t)d) incl += v;
            }
            const uint32_t excl = incl - run;
            if (ib < ent) {  // the fine region is padded to 64 entries: whole 8-entry groups may be stored
                uint4 *dst = reinterpret_cast<uint4 *>(F + ib);
                dst[0] = make_uint4(excl + x[0], excl + x[1], excl + x[2], excl + x[3]);
                dst[1] = make_uint4(excl + x[4], excl + x[5], excl + x[6], excl + x[7]);
            }
            if ((lane & 7) == 7 && i0 + (lane >> 3) * PFX_BLOCK < ent) C[(i0 >> PFX_BLOCK_SHIFT) + (lane >> 3) + 1] = incl;  // block total
        }
    }
}
__device__ void fused_prefix_pass2(const FusedBatch &B, uint64_t *coarse_s, uint32_t warp, uint32_t nwarps, uint32_t lane) {
    for (uint32_t p = warp; p < B.nprof; p += nwarps) {
        const uint32_t slot = B.p_slot[p];
        uint64_t *C = coarse_s + (B.p_rev[p] ? B.rco[slot] : B.co[slot]);
        const uint32_t nblk = (B.p_ent[p] + PFX_BLOCK - 1) >> PFX_BLOCK_SHIFT;
        uint64_t carry = 0;
        if (lane == 0) C[0] = 0;
        for (uint32_t b0 = 0; b0 < nblk; b0 += 32) {
            const uint32_t b = b0 + lane;
            uint64_t v = b < nblk ? C[b + 1] : 0ull, incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += u;
            }
            if (b < nblk) C[b + 1] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

struct FusedArgs {
    const FusedItem *items; uint32_t n_items; uint32_t cap_entries;
    unsigned int *work;                       // item counter
    const uint32_t *wq_f, *wq_r; const unsigned long long *kmax_bits; int lut_in_smem;
    int *fallback;                            // set when a transcript's sums do not fit: the host reruns on the HBM path
};

__global__ void __launch_bounds__(FUSED_THREADS, 2)
k_fused(const __grid_constant__ DevModel m, DevTranscripts tr, FragOut o, FusedArgs a, uint32_t hist_s_n) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ FusedBatch B;
    __shared__ uint32_t s_item;
    uint32_t *fine_s = reinterpret_cast<uint32_t *>(fsm);
    uint64_t *coarse_s = reinterpret_cast<uint64_t *>(fine_s + a.cap_entries);
    uint32_t *lut_s = reinterpret_cast<uint32_t *>(coarse_s + (a.cap_entries / 64 + 2 * FUSED_MAX_TR));
    const uint32_t nlut = a.lut_in_smem ? (1u << (2 * m.k)) : 0u;
    const bool dbl = m.method == RLSIM_AFTER_PRIM_DOUBLE;
    // a symmetric table (w[c] == w[rc(c)]) serves explicit reverse profiles too; otherwise the reverse table sits beside it
    const bool two_luts = dbl && !tr.sym;
    unsigned int *hist_s = lut_s + (two_luts ? 2 : 1) * nlut;  // [hist_s_n] after-fragmentation bins from `low`, then [polya_max+1]
    unsigned int *hist_pa = hist_s + hist_s_n;
    for (uint32_t i = threadIdx.x; i < nlut; i += FUSED_THREADS) { lut_s[i] = a.wq_f[i]; if (two_luts) lut_s[nlut + i] = a.wq_r[i]; }
    for (uint32_t i = threadIdx.x; i < hist_s_n + m.polya_max + 1; i += FUSED_THREADS) hist_s[i] = 0;
    const uint32_t *lut_f = nlut ? lut_s : a.wq_f, *lut_r = nlut ? (two_luts ? lut_s + nlut : lut_s) : a.wq_r;
    const double scale = PFX_SCALE / __longlong_as_double((long long)*a.kmax_bits);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.work, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= a.n_items) break;
        const FusedItem it = a.items[item];
        uint32_t t_cur = it.t0;
        while (t_cur < it.t1) {
            __syncthreads();  // everyone is done with the previous sub-batch: its table and arrays may be rebuilt
            if (warp == 0) fused_pack(B, m, tr, t_cur, it.t1, a.cap_entries, dbl, lane);
            __syncthreads();
            if (B.fail) { if (threadIdx.x == 0) atomicExch(a.fallback, 1); break; }
            t_cur = B.t_next;
            if (B.n == 0) continue;  // a window of unexpressed transcripts
            const uint64_t gs = max(it.g0, B.mol[0]), ge = min(it.g1, B.mol[B.n]);
            if (gs < ge) {
                fused_prefix_pass1(B, m, tr, fine_s, coarse_s, lut_f, lut_r, scale, warp, FUSED_WARPS, lane);
                __syncthreads();
                fused_prefix_pass2(B, coarse_s, warp, FUSED_WARPS, lane);
                __syncthreads();
                for (uint64_t g = gs + threadIdx.x; g < ge; g += FUSED_THREADS) {
                    uint32_t s = 0;  // transcript slot of molecule g
                    while (B.mol[s + 1] <= g) s++;
                    const uint64_t mol = g - B.mol[s];
                    MolCtx c;
                    c.t = B.t[s]; c.tg = m.first_index + c.t; c.L = B.L[s];
                    c.proflen = c.L > (uint32_t)m.k ? c.L - (uint32_t)m.k : 0;
                    c.off = B.off[s];
                    c.Pf = Profile{fine_s + B.fo[s], coarse_s + B.co[s]};
                    c.Pr = B.ro[s] != 0xffffffffu ? Profile{fine_s + B.ro[s], coarse_s + B.rco[s]} : Profile{nullptr, nullptr};
                    c.mirror = B.mirror[s] != 0;
                    Header h = mol_header<0>(m, c, mol);
                    atomicAdd(&hist_pa[h.tail], 1u);
                    if (h.nb > FRAG_LOCAL_BREAKS) {
                        unsigned int slot = atomicAdd(o.long_count, 1u);
                        if (slot < o.long_cap) { o.long_list[slot] = make_uint2(c.t, (uint32_t)(mol & 0xffffffffu)); o.long_hi[slot] = (uint32_t)(mol >> 32); }
                    } else if (h.nb > 0) {
                        const uint32_t length = (uint32_t)h.polyAend;
                        uint32_t br[FRAG_LOCAL_BREAKS + 2];
                        PxStream bs = mol_stream(m, c.tg, mol, SUB_BREAKS);
                        br[0] = 0;
                        int cnt = 1;
                        for (int32_t j = 0; j < h.nb; j++) {  // insertion sort while drawing (fragmentor.go:125-129)
                            uint32_t v = (uint32_t)bs.below((uint32_t)j, (uint64_t)length);
                            int q = cnt++;
                            while (q > 0 && br[q - 1] > v) { br[q] = br[q - 1]; q--; }
                            br[q] = v;
                        }
                        br[cnt] = length;
                        PxStream iv = mol_stream(m, c.tg, mol, SUB_INTERVAL);
                        for (int32_t i = 0; i < h.nb; i++)  // the last interval is never visited (fragmentor.go:132)
                            after_interval(m, c, iv, (uint32_t)i, br[i], br[i + 1], o, hist_s, hist_s_n);
                    }
                }
            }
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < hist_s_n; i += FUSED_THREADS)
        if (hist_s[i]) atomicAdd(&o.hist_frag[m.low + i], (unsigned long long)hist_s[i]);
    for (uint32_t i = threadIdx.x; i <= m.polya_max; i += FUSED_THREADS)
        if (hist_pa[i]) atomicAdd(&o.hist_polya[i], (unsigned long long)hist_pa[i]);
}

// ---- one WARP per transcript, no CTA barrier in the loop (k_fused_warp).
// The CTA-cooperative kernel above pays a barrier tail per item; here warps are independent: a warp fetches an item (one
// transcript + a molecule range), builds the transcript's compact profile in its private slice of shared memory and
// fragments the item's molecules 32 at a time.  The LUT and the histograms are shared by the CTA's warps.
// Transcripts that cannot take this path (letters outside ACGT, or longer than the slice) are appended to `redo` and
// served by k_fused afterwards.
struct WarpArgs {
    uint64_t g_begin, g_end;                  // global molecule range of this launch, cut into units of FW_UNIT molecules
    uint32_t cap_lo, cap_entries;             // this launch serves transcripts with cap_lo < pad64(entries) <= cap_entries
    unsigned int *work;                       // unit counter
    const uint32_t *wq_f;
    FusedItem *redo; unsigned int *redo_count; uint32_t redo_cap;
};
constexpr uint32_t FW_UNIT = 256;             // molecules per unit of work (a transcript with more is shared by several warps)
__host__ __device__ __forceinline__ uint32_t fw_slice_bytes(uint32_t cap) { return (cap / 16 + 4) * 4 + (cap / 16) * 4 + (cap / 64 + 2) * 8; }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 5 : 1)
k_fused_warp(const __grid_constant__ DevModel m, DevTranscripts tr, FragOut o, WarpArgs a, uint32_t hist_s_n) {
    extern __shared__ __align__(16) unsigned char fsm[];
    const uint32_t nlut = 1u << (2 * m.k);
    uint32_t *lut_s = reinterpret_cast<uint32_t *>(fsm);
    unsigned int *hist_s = lut_s + nlut;  // [hist_s_n] after-fragmentation bins from `low`, then [polya_max+1]
    unsigned int *hist_pa = hist_s + hist_s_n;
    const uint32_t hist_words = (hist_s_n + m.polya_max + 1 + 3u) & ~3u;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *slice = reinterpret_cast<unsigned char *>(hist_s + hist_words) + (size_t)warp * fw_slice_bytes(a.cap_entries);
    uint64_t *coarse = reinterpret_cast<uint64_t *>(slice);
    uint32_t *codes = reinterpret_cast<uint32_t *>(coarse + (a.cap_entries / 64 + 2));
    uint32_t *gsum = codes + (a.cap_entries / 16 + 4);
    for (uint32_t i = threadIdx.x; i < nlut; i += WARPS * 32) lut_s[i] = a.wq_f[i];
    for (uint32_t i = threadIdx.x; i < hist_s_n + m.polya_max + 1; i += WARPS * 32) hist_s[i] = 0;
    __syncthreads();
    const bool dbl = m.method == RLSIM_AFTER_PRIM_DOUBLE;
    const uint32_t k = (uint32_t)m.k, kmask = nlut - 1u;
    const CProfile pf{codes, gsum, coarse, lut_s, kmask};
    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(a.work, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        uint64_t gs = a.g_begin + (uint64_t)unit * FW_UNIT;
        if (gs >= a.g_end) break;
        const uint64_t ge = min(a.g_end, gs + FW_UNIT);
        uint32_t t = upper_bound_u64(tr.mol_off, tr.n + 1, gs) - 1;  // the transcript that owns molecule gs
      while (gs < ge) {
        while (tr.mol_off[t + 1] <= gs) t++;
        const uint64_t me = min(ge, tr.mol_off[t + 1]);  // this unit's molecules of transcript t: [gs, me)
        const FusedItem it{t, t + 1, gs, me};
        gs = me;
        const uint32_t L = tr.len[t];
        const uint32_t ent = L >= k ? L - k + 1 : 0u;  // symmetric mode: the forward sums also hold window L-k
        const uint32_t e1 = pad64(ent);
        if (tr.dirty[t] != 0) {  // letters outside ACGT: explicit reverse sums, k_fused (the first size class reports them)
            if (a.cap_lo == 0 && lane == 0) { unsigned int r = atomicAdd(a.redo_count, 1u); if (r < a.redo_cap) a.redo[r] = it; }
            continue;
        }
        if (e1 <= a.cap_lo && !(a.cap_lo == 0)) continue;  // another launch's size class
        if (e1 > a.cap_entries) continue;
        const uint64_t off = tr.off[t];
        // ---- compact profile of transcript t
        const uint32_t nwords = (L + 15) >> 4;
        const uint32_t *pk = tr.packed + (off >> 4);
        for (uint32_t w = lane; w < (pad64(ent) >> 4) + 4; w += 32) codes[w] = w < nwords ? __ldg(pk + w) : 0u;
        __syncwarp();
        uint64_t carry = 0;
        if (lane == 0) coarse[0] = 0;
        for (uint32_t g0 = 0; (g0 << 4) < ent; g0 += 32) {  // 32 groups = 8 blocks per step
            const uint32_t g = g0 + lane, e0 = g << 4;
            const uint32_t nv = e0 < ent ? min(16u, ent - e0) : 0u;
            uint32_t incl = pf.partial(g, nv);
#pragma unroll
            for (int d = 1; d < 4; d <<= 1) {  // inclusive sums inside each block (4 groups)
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d, 4);
                if ((lane & 3u) >= (uint32_t)d) incl += v;
            }
            if (((g >> 2) << 6) < ent) gsum[g] = incl;  // every group of a block that holds entries (groups past the end repeat the block total)
            uint64_t bt = (lane & 3u) == 3u ? (uint64_t)incl : 0ull;  // block totals, running over the step's 8 blocks
#pragma unroll
            for (int d = 4; d < 32; d <<= 1) {
                uint64_t v = __shfl_up_sync(0xffffffffu, bt, d);
                if (lane >= (uint32_t)d) bt += v;
            }
            if ((lane & 3u) == 3u && ((g >> 2) << 6) < ent) coarse[(g >> 2) + 1] = carry + bt;
            carry += __shfl_sync(0xffffffffu, bt, 31);
        }
        __syncwarp();
        // ---- molecules of the item, 32 at a time
        MolCtxT<CProfile> c;
        c.t = t; c.tg = m.first_index + t; c.L = L;
        c.proflen = L > k ? L - k : 0;
        c.off = off; c.Pf = pf; c.Pr = pf; c.mirror = dbl;
        for (uint64_t gb = it.g0; gb < it.g1; gb += 32) {
            const uint64_t g = gb + lane;
            if (g < it.g1) {
                const uint64_t mol = g - tr.mol_off[t];
                Header h = mol_header<0>(m, c, mol);
                atomicAdd(&hist_pa[h.tail], 1u);
                if (h.nb > FRAG_LOCAL_BREAKS) {
                    unsigned int slot = atomicAdd(o.long_count, 1u);
                    if (slot < o.long_cap) { o.long_list[slot] = make_uint2(t, (uint32_t)(mol & 0xffffffffu)); o.long_hi[slot] = (uint32_t)(mol >> 32); }
                } else if (h.nb > 0) {
                    const uint32_t length = (uint32_t)h.polyAend;
                    uint32_t br[FRAG_LOCAL_BREAKS + 2];
                    PxStream bs = mol_stream(m, c.tg, mol, SUB_BREAKS);
                    br[0] = 0;
                    int cnt = 1;
                    for (int32_t j = 0; j < h.nb; j++) {  // insertion sort while drawing (fragmentor.go:125-129)
                        uint32_t v = (uint32_t)bs.below((uint32_t)j, (uint64_t)length);
                        int q = cnt++;
                        while (q > 0 && br[q - 1] > v) { br[q] = br[q - 1]; q--; }
                        br[q] = v;
                    }
                    br[cnt] = length;
                    PxStream iv = mol_stream(m, c.tg, mol, SUB_INTERVAL);
                    for (int32_t i = 0; i < h.nb; i++)  // the last interval is never visited (fragmentor.go:132)
                        after_interval(m, c, iv, (uint32_t)i, br[i], br[i + 1], o, hist_s, hist_s_n);
                }
            }
            __syncwarp();
        }
        __syncwarp();  // the slice is rebuilt for the next transcript
      }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < hist_s_n; i += WARPS * 32)
        if (hist_s[i]) atomicAdd(&o.hist_frag[m.low + i], (unsigned long long)hist_s[i]);
    for (uint32_t i = threadIdx.x; i <= m.polya_max; i += WARPS * 32)
        if (hist_pa[i]) atomicAdd(&o.hist_polya[i], (unsigned long long)hist_pa[i]);
}

// Deferred molecules of the fused path (more breakpoints than a thread sorts locally): one CTA per molecule, its
// transcript's prefix sums rebuilt in shared memory, breakpoints bitonic-sorted in shared memory.
constexpr int FLONG_THREADS = 256;
constexpr uint32_t FLONG_BREAKS = 4096;  // power of two; the host only takes the fused path when far fewer are expected
__global__ void __launch_bounds__(FLONG_THREADS)
k_fused_long(const __grid_constant__ DevModel m, DevTranscripts tr, FragOut o, FusedArgs a, uint32_t n_long) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ FusedBatch B;
    __shared__ Header hs;
    uint32_t *fine_s = reinterpret_cast<uint32_t *>(fsm);
    uint64_t *coarse_s = reinterpret_cast<uint64_t *>(fine_s + a.cap_entries);
    uint32_t *brs = reinterpret_cast<uint32_t *>(coarse_s + (a.cap_entries / 64 + 2 * FUSED_MAX_TR));
    const uint32_t slot = blockIdx.x;
    if (slot >= n_long) return;
    const uint32_t t = o.long_list[slot].x;
    const uint64_t mol = (uint64_t)o.long_list[slot].y | ((uint64_t)o.long_hi[slot] << 32);
    const bool dbl = m.method == RLSIM_AFTER_PRIM_DOUBLE;
    if (threadIdx.x < 32) fused_pack(B, m, tr, t, t + 1, a.cap_entries, dbl, threadIdx.x);
    __syncthreads();
    if (B.fail || B.n != 1) { if (threadIdx.x == 0) atomicExch(a.fallback, 1); return; }
    const double scale = PFX_SCALE / __longlong_as_double((long long)*a.kmax_bits);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fused_prefix_pass1(B, m, tr, fine_s, coarse_s, a.wq_f, a.wq_r, scale, warp, FLONG_THREADS / 32, lane);
    __syncthreads();
    fused_prefix_pass2(B, coarse_s, warp, FLONG_THREADS / 32, lane);
    MolCtx c;
    c.t = t; c.tg = m.first_index + t; c.L = B.L[0];
    c.proflen = c.L > (uint32_t)m.k ? c.L - (uint32_t)m.k : 0;
    c.off = B.off[0];
    c.Pf = Profile{fine_s + B.fo[0], coarse_s + B.co[0]};
    c.Pr = B.ro[0] != 0xffffffffu ? Profile{fine_s + B.ro[0], coarse_s + B.rco[0]} : Profile{nullptr, nullptr};
    c.mirror = B.mirror[0] != 0;
    if (threadIdx.x == 0) hs = mol_header<0>(m, c, mol);
    __syncthreads();
    const Header h = hs;
    const uint32_t n = (uint32_t)h.nb + 2;
    uint32_t npow = 1;
    while (npow < n) npow <<= 1;
    if (npow > FLONG_BREAKS) { if (threadIdx.x == 0) atomicExch(a.fallback, 1); return; }
    const uint32_t length = (uint32_t)h.polyAend;
    PxStream bs = mol_stream(m, c.tg, mol, SUB_BREAKS);
    for (uint32_t j = threadIdx.x; j < npow; j += FLONG_THREADS) {
        uint32_t v;
        if (j < (uint32_t)h.nb) v = (uint32_t)bs.below(j, (uint64_t)length);
        else if (j == (uint32_t)h.nb) v = 0u;
        else if (j == (uint32_t)h.nb + 1) v = length;
        else v = 0xffffffffu;
        brs[j] = v;
    }
    __syncthreads();
    for (uint32_t k2 = 2; k2 <= npow; k2 <<= 1) {
        for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (uint32_t i = threadIdx.x; i < npow; i += FLONG_THREADS) {
                uint32_t ixj = i ^ j2;
                if (ixj > i) {
                    uint32_t x = brs[i], y = brs[ixj];
                    bool up = ((i & k2) == 0);
                    if ((x > y) == up) { brs[i] = y; brs[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    PxStream iv = mol_stream(m, c.tg, mol, SUB_INTERVAL);
    for (uint32_t i = threadIdx.x; i < (uint32_t)h.nb; i += FLONG_THREADS)
        after_interval(m, c, iv, i, brs[i], brs[i + 1], o, nullptr, 0);
}

static uint32_t frag_hist_bins(const DevModel &m) {
    uint32_t n = m.high - m.low + 1;
    return n > 8192 ? 0 : n;
}

int launch_fragment(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, uint64_t g_begin, uint64_t g_end,
                    const FragBuffers &fb, const uint32_t *warp_t) {
    if (g_end <= g_begin) return 0;
    const uint64_t n_mol = g_end - g_begin;
    FragOut o{fb.keys, fb.cap, key_len_bits(m.high), fb.count, fb.hist_frag, fb.hist_polya, fb.long_list, fb.long_hi, fb.long_cap, fb.long_count,
              fb.error, fb.q_a, fb.q_b, fb.q_cap, fb.q_count};
    // after_* methods only queue intervals here (k_intervals owns the after-fragmentation histogram)
    uint32_t hist_s_n = m.method <= RLSIM_AFTER_NOPRIM_DOUBLE ? 0 : frag_hist_bins(m);
    size_t smem = ((size_t)hist_s_n + m.polya_max + 1) * 4;
    if (smem > 48 * 1024) { hist_s_n = 0; smem = ((size_t)m.polya_max + 1) * 4; }
    if (smem > 48 * 1024) return RLSIM_E_UNSUPPORTED;
    uint64_t blocks = (n_mol + FRAG_THREADS - 1) / FRAG_THREADS;
    if (blocks > 0x7fffffffull) return RLSIM_E_UNSUPPORTED;
    if (m.method == RLSIM_PRIM_JUMP) k_fragment<2><<<(unsigned)blocks, FRAG_THREADS, smem, st>>>(m, tr, g_begin, g_end, o, hist_s_n, warp_t);
    else if (m.method == RLSIM_PRE_PRIM) k_fragment<1><<<(unsigned)blocks, FRAG_THREADS, smem, st>>>(m, tr, g_begin, g_end, o, hist_s_n, warp_t);
    else k_fragment<0><<<(unsigned)blocks, FRAG_THREADS, smem, st>>>(m, tr, g_begin, g_end, o, hist_s_n, warp_t);
    return 0;
}

// n_q_dirty: queued intervals of transcripts that stay on the two-level arrays although records exist (q_count[1])
int launch_intervals(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint64_t n_q, uint64_t n_q_dirty,
                     const unsigned long long *n_q_dev) {
    if (n_q == 0) return 0;
    FragOut o{fb.keys, fb.cap, key_len_bits(m.high), fb.count, fb.hist_frag, fb.hist_polya, fb.long_list, fb.long_hi, fb.long_cap, fb.long_count,
              fb.error, fb.q_a, fb.q_b, fb.q_cap, fb.q_count};
    uint32_t hist_s_n = frag_hist_bins(m);
    uint64_t blocks = (n_q + IV_THREADS - 1) / IV_THREADS;
    if (blocks > 0x7fffffffull) return RLSIM_E_UNSUPPORTED;
    const uint64_t max_blocks = 148ull * 5 * 6;
    if (tr.Rf != nullptr) {
        const size_t smem = (size_t)hist_s_n * 4 + ((size_t)4 << (2 * m.k));
        k_intervals<1><<<(unsigned)std::min(blocks, max_blocks), IV_THREADS, smem, st>>>(m, tr, o, n_q, hist_s_n, n_q_dev);
        if (n_q_dirty == 0) return 0;
    }
    k_intervals<0><<<(unsigned)std::min(blocks, max_blocks), IV_THREADS, (size_t)hist_s_n * 4, st>>>(m, tr, o, n_q, hist_s_n, n_q_dev);
    return 0;
}

// P(0..n) of one profile, whichever layout holds it (parity probe)
__global__ void k_probe_prefix(const __grid_constant__ DevModel m, DevTranscripts tr, uint32_t t, int reverse, uint64_t *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t L = tr.len[t], proflen = L > (uint32_t)m.k ? L - (uint32_t)m.k : 0;
    if (i > proflen) { out[i] = 0; return; }
    if (uses_records(tr, t)) {
        const MolCtxT<RProfile> c = make_ctx_rec(m, tr, t, tr.lut);
        out[i] = reverse ? c.Pf.P(proflen + 1) - c.Pf.P(proflen + 1 - i) : c.Pf.P(i);  // mirrored: P_rev(i) = W(n+1) - W(n+1-i)
    } else {
        const MolCtx c = make_ctx(m, tr, t);
        if (!reverse) out[i] = c.Pf.P(i);
        else out[i] = c.mirror ? c.Pf.P(proflen + 1) - c.Pf.P(proflen + 1 - i) : c.Pr.P(i);
    }
}
void launch_probe_prefix(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, uint32_t t, int reverse, uint64_t *out, uint32_t n) {
    if (n) k_probe_prefix<<<(n + 127) / 128, 128, 0, st>>>(m, tr, t, reverse, out, n);
}

// ---- fused path launches.  smem_optin = the device's opt-in shared memory per block.
size_t fused_smem_bytes(const DevModel &m, uint32_t cap_entries, int lut_tables, uint32_t hist_s_n) {
    return (size_t)cap_entries * 4 + ((size_t)cap_entries / 64 + 2 * FUSED_MAX_TR) * 8 + (size_t)lut_tables * ((size_t)4 << (2 * m.k)) +
           ((size_t)hist_s_n + m.polya_max + 1) * 4 + 16;
}
int launch_fused(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, const FusedItem *items,
                 uint32_t n_items, uint32_t cap_entries, unsigned int *work, const uint32_t *wq_f, const uint32_t *wq_r,
                 const unsigned long long *kmax_bits, int *fallback, int sm_count) {
    if (!n_items) return 0;
    FragOut o{fb.keys, fb.cap, key_len_bits(m.high), fb.count, fb.hist_frag, fb.hist_polya, fb.long_list, fb.long_hi, fb.long_cap, fb.long_count,
              fb.error, fb.q_a, fb.q_b, fb.q_cap, fb.q_count};
    bool lut_in_smem = m.k <= 6;
    const int tables = (m.method == RLSIM_AFTER_PRIM_DOUBLE && !tr.sym) ? 2 : 1;
    uint32_t hist_s_n = frag_hist_bins(m);
    size_t smem = fused_smem_bytes(m, cap_entries, lut_in_smem ? tables : 0, hist_s_n);
    if (smem > 227 * 1024 && lut_in_smem) { lut_in_smem = false; smem = fused_smem_bytes(m, cap_entries, 0, hist_s_n); }
    if (smem > 227 * 1024) { hist_s_n = 0; smem = fused_smem_bytes(m, cap_entries, 0, 0); }
    if (smem > 227 * 1024) return RLSIM_E_UNSUPPORTED;
    cudaFuncSetAttribute(k_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused, FUSED_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    const uint32_t blocks = std::min<uint32_t>(n_items, (uint32_t)(sm_count * per_sm));
    cudaMemsetAsync(work, 0, sizeof(unsigned int), st);
    FusedArgs a{items, n_items, cap_entries, work, wq_f, wq_r, kmax_bits, lut_in_smem ? 1 : 0, fallback};
    k_fused<<<blocks, FUSED_THREADS, smem, st>>>(m, tr, o, a, hist_s_n);
    return 0;
}
// per-warp slices of `cap_entries`; warps = 8 (small / medium classes) or 4 (large class)
int launch_fused_warp(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint64_t g_begin, uint64_t g_end,
                      uint32_t cap_lo, uint32_t cap_entries, int warps, unsigned int *work, const uint32_t *wq_f, FusedItem *redo,
                      unsigned int *redo_count, uint32_t redo_cap, int sm_count) {
    if (g_end <= g_begin) return 0;
    if (m.k > 6) return RLSIM_E_UNSUPPORTED;  // the LUT must fit shared memory beside the slices
    FragOut o{fb.keys, fb.cap, key_len_bits(m.high), fb.count, fb.hist_frag, fb.hist_polya, fb.long_list, fb.long_hi, fb.long_cap, fb.long_count,
              fb.error, fb.q_a, fb.q_b, fb.q_cap, fb.q_count};
    uint32_t hist_s_n = frag_hist_bins(m);
    auto bytes = [&](uint32_t hs) {
        const size_t hist_words = ((size_t)hs + m.polya_max + 1 + 3) & ~(size_t)3;
        return ((size_t)4 << (2 * m.k)) + hist_words * 4 + (size_t)warps * fw_slice_bytes(cap_entries) + 16;
    };
    size_t smem = bytes(hist_s_n);
    if (smem > 227 * 1024) { hist_s_n = 0; smem = bytes(0); }
    if (smem > 227 * 1024) return RLSIM_E_UNSUPPORTED;
    WarpArgs a{g_begin, g_end, cap_lo, cap_entries, work, wq_f, redo, redo_count, redo_cap};
    cudaMemsetAsync(work, 0, sizeof(unsigned int), st);
    const uint64_t units = (g_end - g_begin + FW_UNIT - 1) / FW_UNIT;
    int per_sm = 1;
    if (warps == 8) {
        cudaFuncSetAttribute(k_fused_warp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_warp<8>, 256, smem);
        const uint32_t blocks = (uint32_t)std::min<uint64_t>((units + 7) / 8, (uint64_t)(sm_count * std::max(per_sm, 1)));
        k_fused_warp<8><<<blocks, 256, smem, st>>>(m, tr, o, a, hist_s_n);
    } else {
        cudaFuncSetAttribute(k_fused_warp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_warp<4>, 128, smem);
        const uint32_t blocks = (uint32_t)std::min<uint64_t>((units + 3) / 4, (uint64_t)(sm_count * std::max(per_sm, 1)));
        k_fused_warp<4><<<blocks, 128, smem, st>>>(m, tr, o, a, hist_s_n);
    }
    return 0;
}
uint64_t fused_warp_units(uint64_t n_mol) { return (n_mol + FW_UNIT - 1) / FW_UNIT; }

int launch_fused_long(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint32_t n_long,
                      uint32_t cap_entries, const uint32_t *wq_f, const uint32_t *wq_r, const unsigned long long *kmax_bits, int *fallback) {
    if (!n_long) return 0;
    FragOut o{fb.keys, fb.cap, key_len_bits(m.high), fb.count, fb.hist_frag, fb.hist_polya, fb.long_list, fb.long_hi, n_long, nullptr,
              fb.error, fb.q_a, fb.q_b, fb.q_cap, fb.q_count};
    const size_t smem = (size_t)cap_entries * 4 + ((size_t)cap_entries / 64 + 2 * FUSED_MAX_TR) * 8 + (size_t)FLONG_BREAKS * 4 + 16;
    if (smem > 227 * 1024) return RLSIM_E_UNSUPPORTED;
    cudaFuncSetAttribute(k_fused_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FusedArgs a{nullptr, 0, cap_entries, nullptr, wq_f, wq_r, kmax_bits, 0, fallback};
    k_fused_long<<<n_long, FLONG_THREADS, smem, st>>>(m, tr, o, a, n_long);
    return 0;
}

int launch_fragment_long(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint32_t n_long,
                         const unsigned int *n_long_dev) {
    if (n_long == 0) return 0;
    FragOut o{fb.keys, fb.cap, key_len_bits(m.high), fb.count, fb.hist_frag, fb.hist_polya, fb.long_list, fb.long_hi, n_long, nullptr,
              fb.error, fb.q_a, fb.q_b, fb.q_cap, fb.q_count};
    const uint32_t smem_cap = 32768;  // breakpoints per molecule (128 KB of shared memory)
    cudaFuncSetAttribute(k_fragment_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_cap * 4));
    // device-side count: a fixed grid strides over however many molecules were deferred (normally none)
    k_fragment_long<<<n_long_dev ? std::min(n_long, 296u) : n_long, LONG_THREADS, smem_cap * 4, st>>>(m, tr, o, n_long, smem_cap, n_long_dev);
    return 0;
}

}  // namespace rl

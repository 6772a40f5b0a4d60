// Target sampling, size selection and final draw (kernels 3 + 4, DESIGN.md sections 5.4-5.5).
//
//   k_target        : target.go:102-123 / raw_target.go:104-129 - req_frags iid target lengths -> histogram
//   k_candidates .. : sampler.go:65-106 + pool.go:191-204 + transcript.go:172-187 - for each requested length l,
//                     draw min(req_l, T_l) of the T_l pooled copies uniformly WITHOUT replacement.  The reference does
//                     it by sequential weighted draws with decrement (transcript, then fragment); that two-stage
//                     scheme is a multivariate hypergeometric draw over species, realised here in parallel as
//                     "the first `pick` distinct values of below(block j, T_l), j = 0,1,..." (spec 4.6):
//                     candidates -> stable radix sort -> first-occurrence flags -> scan in draw order -> cut at `pick`
//                     -> rank-to-species binary search in the per-class prefix sums of the amplified counts.
//   k_expand        : sampler.go:187-227 + frag.go - one record per drawn copy, strand ~ Bernoulli(bias),
//                     per-transcript ids, output ordered by (transcript, start, end).
#include <algorithm>
#include "kernels.cuh"
#include "variates.cuh"

namespace rl {

// ---------------------------------------------------------------- target lengths
constexpr int TGT_THREADS = 256;
__global__ void __launch_bounds__(TGT_THREADS, 4)
k_target(const __grid_constant__ DevModel m, uint64_t first, uint64_t n, unsigned long long *__restrict__ hist, uint32_t hist_n,
         uint32_t smem_bins) {
    extern __shared__ unsigned int hs[];
    for (uint32_t i = threadIdx.x; i < smem_bins; i += TGT_THREADS) hs[i] = 0;
    __syncthreads();
    for (uint64_t q = (uint64_t)blockIdx.x * TGT_THREADS + threadIdx.x; q < n; q += (uint64_t)gridDim.x * TGT_THREADS) {
        const uint64_t i = first + q;  // global draw index: the stream address
        PxCursor cur(PxStream{m.key_target, (uint32_t)(i & 0xffffffffu), (uint32_t)(i >> 32), 0u});
        uint32_t l = m.n_target == 0 ? px_raw_len(cur, m) : px_comp_len(cur, m.target[px_mix_comp(cur, m.target, m.n_target)]);
        if (l < smem_bins) atomicAdd(&hs[l], 1u);
        else if (l < hist_n) atomicAdd(&hist[l], 1ull);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < smem_bins; i += TGT_THREADS)
        if (hs[i]) atomicAdd(&hist[i], (unsigned long long)hs[i]);
}
void launch_target(cudaStream_t st, const DevModel &m, uint64_t first, uint64_t n, unsigned long long *hist, uint32_t hist_n) {
    if (!n) return;
    uint32_t smem_bins = hist_n <= 12000 ? hist_n : 12000;
    uint64_t blocks = (n + TGT_THREADS - 1) / TGT_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_target<<<(unsigned)blocks, TGT_THREADS, smem_bins * 4, st>>>(m, first, n, hist, hist_n, smem_bins);
}

// ---------------------------------------------------------------- per-length classes
__global__ void k_gather_counts(const uint32_t *__restrict__ order, const uint64_t *__restrict__ count,
                                uint64_t *__restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = count[order[i]];
}
void launch_gather_counts(cudaStream_t st, const uint32_t *order, const uint64_t *count, uint64_t *out, uint64_t n) {
    if (n) k_gather_counts<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order, count, out, n);
}

// bounds[l] = first position in the length-sorted species view with length >= l, l = 0..n_len
__global__ void k_class_bounds(const uint32_t *__restrict__ sorted_len, uint64_t n, uint32_t n_len, uint64_t *__restrict__ bounds) {
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > n_len) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (sorted_len[mid] < l) lo = mid + 1; else hi = mid;
    }
    bounds[l] = lo;
}
void launch_class_bounds(cudaStream_t st, const uint32_t *sorted_len, uint64_t n, uint32_t n_len, uint64_t *bounds) {
    k_class_bounds<<<(n_len + 1 + 127) / 128, 128, 0, st>>>(sorted_len, n, n_len, bounds);
}

// ---------------------------------------------------------------- without-replacement draw
__device__ __forceinline__ uint32_t class_of(const DevClassPlan &p, uint64_t g) {
    return upper_bound_u64(p.cand_off, p.n_classes + 1, g) - 1;
}

// keys32: optional narrow copy of the keys for the sort (all keys < 2^32: a third less sort traffic)
__global__ void k_candidates(const __grid_constant__ DevModel m, DevClassPlan plan, uint64_t n_cand, uint64_t *__restrict__ keys,
                             uint32_t *__restrict__ keys32, uint32_t *__restrict__ idx) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_cand) return;
    uint32_t c = class_of(plan, g);
    uint32_t j = (uint32_t)(g - plan.cand_off[c]);
    PxStream s{m.key_select, plan.len[c], 0u, 0u};
    const uint64_t key = plan.G[c] + s.below(j, plan.T[c]);
    keys[g] = key;
    if (keys32) keys32[g] = (uint32_t)key;
    idx[g] = (uint32_t)g;
}
void launch_candidates(cudaStream_t st, const DevModel &m, DevClassPlan plan, uint64_t n_cand, uint64_t *keys, uint32_t *keys32,
                       uint32_t *idx) {
    if (n_cand) k_candidates<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(m, plan, n_cand, keys, keys32, idx);
}

// after the stable sort by key: the first entry of each run of equal keys is the earliest draw of that rank
// The sort only has to bring equal keys together in draw order, so it may stop at the low `sort_bits` bits (one radix
// pass fewer): a run of candidates that agree in those bits is short (about one candidate per 2^sort_bits values) and keeps
// the draw order (stable sort), and a candidate is a first occurrence iff no earlier member of its run has the same key.
template <typename K>
__device__ __forceinline__ bool first_in_run(const K *__restrict__ sk, uint64_t p, K mask) {
    const K key = sk[p], low = key & mask;
    for (uint64_t q = p; q > 0; q--) {
        const K prev = sk[q - 1];
        if ((prev & mask) != low) break;
        if (prev == key) return false;
    }
    return true;
}
template <typename K>
__global__ void k_first_flags(const K *__restrict__ sk, const uint32_t *__restrict__ si, uint64_t n,
                              uint32_t *__restrict__ flags, K mask) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    flags[si[p]] = first_in_run(sk, p, mask) ? 1u : 0u;
}
void launch_first_flags(cudaStream_t st, const uint64_t *sk, const uint32_t *si, uint64_t n, uint32_t *flags, int sort_bits) {
    const uint64_t mask = sort_bits >= 64 ? ~0ull : ((1ull << sort_bits) - 1ull);
    if (n) k_first_flags<uint64_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sk, si, n, flags, mask);
}
void launch_first_flags32(cudaStream_t st, const uint32_t *sk, const uint32_t *si, uint64_t n, uint32_t *flags, int sort_bits) {
    const uint32_t mask = sort_bits >= 32 ? ~0u : ((1u << sort_bits) - 1u);
    if (n) k_first_flags<uint32_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sk, si, n, flags, mask);
}

__global__ void k_apply_picks(DevClassPlan plan, uint64_t n_cand, const uint64_t *__restrict__ keys,
                              const uint32_t *__restrict__ flags, const uint64_t *__restrict__ fscan,
                              const uint64_t *__restrict__ cls_pre, const uint64_t *__restrict__ bounds,
                              const uint32_t *__restrict__ order, unsigned long long *__restrict__ hits, int *short_flag) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_cand) return;
    uint32_t c = class_of(plan, g);
    uint64_t c0 = plan.cand_off[c];
    if (g == c0) {  // enough distinct candidates in this class?
        uint64_t distinct = fscan[plan.cand_off[c + 1]] - fscan[c0];
        if (distinct < plan.pick[c]) short_flag[c] = 1;
    }
    if (!flags[g]) return;
    if (fscan[g] - fscan[c0] >= plan.pick[c]) return;  // beyond the first `pick` distinct values
    uint64_t x = keys[g] - plan.G[c];                 // global rank within the class
    uint32_t l = plan.len[c];
    uint64_t b = bounds[l], e = bounds[l + 1];
    uint64_t t_local = cls_pre[e] - cls_pre[b];
    if (x < plan.base[c] || x - plan.base[c] >= t_local) return;  // another rank's copy
    uint64_t tgt = cls_pre[b] + (x - plan.base[c]);
    uint64_t lo = b, hi = e;  // species position with cls_pre[pos] <= tgt < cls_pre[pos+1]
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (!(cls_pre[mid + 1] > tgt)) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&hits[order[lo]], 1ull);
}
// The same in SORTED candidate order (single-rank form): neighbouring threads hold neighbouring copy ranks, so their
// bisections of cls_pre walk the same cache lines and end in adjacent species - the per-candidate cost becomes one random
// 8-byte read (the candidate's position among the first occurrences in draw order) instead of an 11-step random search.
// A candidate is a first occurrence iff its sorted neighbour on the left has another key (the sort is stable).
// cut[c] = the first candidate index of class c that lies beyond the class's first `pick` distinct values: the number of
// first occurrences before a candidate (fscan) does not decrease along the draws, so "among the first pick distinct values"
// is "first occurrence and index below cut[c]" - one bisection per class instead of a random read per candidate.
__global__ void k_cut_index(DevClassPlan plan, const uint64_t *__restrict__ fscan, uint64_t *__restrict__ cut, int *short_flag) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plan.n_classes) return;
    const uint64_t c0 = plan.cand_off[c], c1 = plan.cand_off[c + 1], tgt = fscan[c0] + plan.pick[c];
    if (fscan[c1] < tgt) short_flag[c] = 1;  // not enough distinct candidates: the host draws more and repeats
    uint64_t lo = c0, hi = c1;  // first g with fscan[g] >= tgt
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (fscan[mid] < tgt) lo = mid + 1; else hi = mid;
    }
    cut[c] = lo;
}

template <typename K>
__global__ void k_apply_picks_sorted(DevClassPlan plan, uint64_t n_cand, const K *__restrict__ sk, const uint32_t *__restrict__ sidx,
                                     const uint64_t *__restrict__ cut, const uint64_t *__restrict__ cls_pre,
                                     const uint64_t *__restrict__ bounds, const uint32_t *__restrict__ order,
                                     unsigned long long *__restrict__ hits, K mask) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_cand) return;
    const uint64_t key = (uint64_t)sk[p];
    const uint32_t c = upper_bound_u64(plan.G, plan.n_classes, key) - 1;  // keys of class c lie in [G[c], G[c] + T[c])
    if ((uint64_t)sidx[p] >= cut[c]) return;  // beyond the first `pick` distinct values
    if (!first_in_run(sk, p, mask)) return;   // a later draw of the same copy rank
    const uint64_t x = key - plan.G[c];
    const uint32_t l = plan.len[c];
    const uint64_t b = bounds[l], e = bounds[l + 1];
    const uint64_t t_local = cls_pre[e] - cls_pre[b];
    if (x < plan.base[c] || x - plan.base[c] >= t_local) return;  // another rank's copy
    const uint64_t tgt = cls_pre[b] + (x - plan.base[c]);
    uint64_t lo = b, hi = e;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (!(cls_pre[mid + 1] > tgt)) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&hits[order[lo]], 1ull);
}
void launch_apply_picks_sorted(cudaStream_t st, DevClassPlan plan, uint64_t n_cand, const uint64_t *sk64, const uint32_t *sk32,
                               const uint32_t *sidx, const uint64_t *fscan, uint64_t *cut, const uint64_t *cls_pre, const uint64_t *bounds,
                               const uint32_t *order, unsigned long long *hits, int *short_flag, int sort_bits) {
    if (!n_cand) return;
    k_cut_index<<<(plan.n_classes + 127) / 128, 128, 0, st>>>(plan, fscan, cut, short_flag);
    const unsigned blocks = (unsigned)((n_cand + 255) / 256);
    if (sk32) k_apply_picks_sorted<uint32_t><<<blocks, 256, 0, st>>>(plan, n_cand, sk32, sidx, cut, cls_pre, bounds, order, hits,
                                                                     sort_bits >= 32 ? ~0u : ((1u << sort_bits) - 1u));
    else k_apply_picks_sorted<uint64_t><<<blocks, 256, 0, st>>>(plan, n_cand, sk64, sidx, cut, cls_pre, bounds, order, hits,
                                                                sort_bits >= 64 ? ~0ull : ((1ull << sort_bits) - 1ull));
}

void launch_apply_picks(cudaStream_t st, DevClassPlan plan, uint64_t n_cand, const uint64_t *keys, const uint32_t *flags,
                        const uint64_t *fscan, const uint64_t *cls_pre, const uint64_t *bounds, const uint32_t *order,
                        unsigned long long *hits, int *short_flag) {
    if (n_cand)
        k_apply_picks<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(plan, n_cand, keys, flags, fscan, cls_pre, bounds,
                                                                        order, hits, short_flag);
}

// ---- multi-GPU: length classes are partitioned over the ranks; the rank that draws a class routes every drawn
// (or, in complement mode, excluded) copy rank to the rank that owns the copy.  Record = {len, 0, x_lo, x_hi}.
__device__ __forceinline__ bool picked(const DevClassPlan &plan, uint64_t g, uint32_t c, const uint32_t *flags,
                                       const uint64_t *fscan) {
    return flags[g] && (fscan[g] - fscan[plan.cand_off[c]] < plan.pick[c]);
}
__device__ __forceinline__ uint32_t owner_of(const uint64_t *all_base, uint32_t nparts, uint32_t n_len, uint32_t l, uint64_t x) {
    uint32_t o = 0;  // last rank whose first copy rank of this class is <= x (empty ranks have base == next base)
    for (uint32_t r = 1; r < nparts; r++)
        if (all_base[(uint64_t)r * n_len + l] <= x) o = r;
    return o;
}
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_MAX_PARTS = 64;
// pass 0: per-owner counts (+ per-class "not enough distinct candidates" flags); pass 1: scatter into the owner groups;
// pass 2: both in one sweep - the owner groups have capacities sized by the host from the expected split (seg_off[64 + r]),
// a record that does not fit raises *overflow and the host falls back to passes 0 + 1
__global__ void __launch_bounds__(EMIT_THREADS)
k_emit_selected(DevClassPlan plan, uint64_t n_cand, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ flags,
                const uint64_t *__restrict__ fscan, const uint64_t *__restrict__ all_base, uint32_t nparts, uint32_t n_len,
                int pass, unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ seg_off,
                unsigned long long *__restrict__ cursor, uint4 *__restrict__ records, int *short_flag, int compact, int *overflow) {
    __shared__ unsigned int cnt_s[EMIT_MAX_PARTS];
    __shared__ unsigned long long base_s[EMIT_MAX_PARTS];
    if (threadIdx.x < EMIT_MAX_PARTS) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    uint64_t g = (uint64_t)blockIdx.x * EMIT_THREADS + threadIdx.x;
    bool sel = false;
    uint32_t owner = 0, l = 0, local = 0;
    uint64_t x = 0;
    if (g < n_cand) {
        uint32_t c = class_of(plan, g);
        if (pass != 1 && g == plan.cand_off[c]) {
            uint64_t distinct = fscan[plan.cand_off[c + 1]] - fscan[plan.cand_off[c]];
            if (distinct < plan.pick[c]) short_flag[c] = 1;
        }
        if (picked(plan, g, c, flags, fscan)) {
            sel = true;
            x = keys[g] - plan.G[c];
            l = plan.len[c];
            owner = owner_of(all_base, nparts, n_len, l, x);
            local = atomicAdd(&cnt_s[owner], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < nparts && cnt_s[threadIdx.x]) {
        if (pass == 0) atomicAdd(&counts[threadIdx.x], (unsigned long long)cnt_s[threadIdx.x]);
        else base_s[threadIdx.x] = seg_off[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt_s[threadIdx.x]);
    }
    if (pass == 0) return;
    __syncthreads();
    if (sel) {
        const unsigned long long pos = base_s[owner] + local;
        if (pass == 2 && pos - seg_off[owner] >= seg_off[EMIT_MAX_PARTS + owner]) { atomicExch(overflow, 1); return; }
        // compact form (the library's own exchange): 8 bytes, length in the top 24 bits - every class total is below 2^40
        if (compact) reinterpret_cast<uint64_t *>(records)[pos] = ((uint64_t)l << 40) | x;
        else records[pos] = make_uint4(l, 0u, (uint32_t)(x & 0xffffffffu), (uint32_t)(x >> 32));
    }
}
void launch_emit_selected(cudaStream_t st, DevClassPlan plan, uint64_t n_cand, const uint64_t *keys, const uint32_t *flags,
                          const uint64_t *fscan, const uint64_t *all_base, uint32_t nparts, uint32_t n_len, int pass,
                          unsigned long long *counts, const unsigned long long *seg_off, unsigned long long *cursor,
                          uint4 *records, int *short_flag, int compact, int *overflow) {
    if (n_cand)
        k_emit_selected<<<(unsigned)((n_cand + EMIT_THREADS - 1) / EMIT_THREADS), EMIT_THREADS, 0, st>>>(
            plan, n_cand, keys, flags, fscan, all_base, nparts, n_len, pass, counts, seg_off, cursor, records, short_flag, compact, overflow);
}

// records received from the drawing ranks -> hits on this rank's species
__global__ void k_apply_records(const uint4 *__restrict__ records, uint64_t n, const uint64_t *__restrict__ my_base,
                                const uint64_t *__restrict__ cls_pre, const uint64_t *__restrict__ bounds,
                                const uint32_t *__restrict__ order, unsigned long long *__restrict__ hits, int *error, int compact) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t l;
    uint64_t x;
    if (compact) { const uint64_t r = reinterpret_cast<const uint64_t *>(records)[i]; l = (uint32_t)(r >> 40); x = r & ((1ull << 40) - 1ull); }
    else { const uint4 r = records[i]; l = r.x; x = (uint64_t)r.z | ((uint64_t)r.w << 32); }
    uint64_t b = bounds[l], e = bounds[l + 1];
    uint64_t t_local = cls_pre[e] - cls_pre[b];
    if (x < my_base[l] || x - my_base[l] >= t_local) { atomicExch(error, RLSIM_E_ARG); return; }  // mis-routed record
    uint64_t tgt = cls_pre[b] + (x - my_base[l]);
    uint64_t lo = b, hi = e;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (!(cls_pre[mid + 1] > tgt)) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&hits[order[lo]], 1ull);
}
void launch_apply_records(cudaStream_t st, const uint4 *records, uint64_t n, const uint64_t *my_base, const uint64_t *cls_pre,
                          const uint64_t *bounds, const uint32_t *order, unsigned long long *hits, int *error, int compact) {
    if (n) k_apply_records<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(records, n, my_base, cls_pre, bounds, order, hits, error, compact);
}

// mode per length: 0 = nothing drawn, 1 = whole class drawn, 2 = hits are the drawn copies, 3 = hits are the excluded copies
__global__ void k_taken(DevSpecies sp, const uint8_t *__restrict__ mode_by_len, uint32_t n_len,
                        const unsigned long long *__restrict__ hits, uint64_t *__restrict__ taken) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    uint32_t l = sp.len[i];
    uint8_t mode = l < n_len ? mode_by_len[l] : 0;
    uint64_t c = sp.count[i], h = hits[i];
    taken[i] = mode == 1 ? c : mode == 2 ? h : mode == 3 ? c - h : 0;
}
void launch_taken(cudaStream_t st, const DevSpecies &sp, const uint8_t *mode_by_len, uint32_t n_len,
                  const unsigned long long *hits, uint64_t *taken) {
    if (sp.n) k_taken<<<(unsigned)((sp.n + 255) / 256), 256, 0, st>>>(sp, mode_by_len, n_len, hits, taken);
}

// ---------------------------------------------------------------- final records
// first output record of every transcript that has species: the per-transcript fragment ids count from there
__global__ void k_tr_out_base(DevSpecies sp, const uint64_t *__restrict__ out_off, uint64_t *__restrict__ tr_out_base) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sp.n) return;
    const uint32_t t = sp.tr[s];
    if (s == 0 || sp.tr[s - 1] != t) tr_out_base[t] = out_off[s];
}

// records per transcript (the packed host format replaces the per-record transcript index and id by these counts):
// the last species of a transcript closes its run of output records
__global__ void k_tr_counts(DevSpecies sp, const uint64_t *__restrict__ out_off, const uint64_t *__restrict__ tr_out_base,
                            uint32_t *__restrict__ tr_counts) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sp.n) return;
    const uint32_t t = sp.tr[s];
    if (s + 1 == sp.n || sp.tr[s + 1] != t) tr_counts[t] = (uint32_t)(out_off[s + 1] - tr_out_base[t]);
}

// One warp per 32 consecutive species: their output records are consecutive (out_off is the running sum of the copies
// taken), so the warp writes them 32 at a time, coalesced, and finds the species of a record by a 5-step search over the
// lanes' own offsets with shuffles - no 22-step bisection of out_off in memory per record.
__global__ void __launch_bounds__(256)
k_expand(const __grid_constant__ DevModel m, DevSpecies sp, const uint64_t *__restrict__ out_off,
         const uint64_t *__restrict__ tr_out_base, uint64_t n_out, DevSampled out) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t s0 = w * 32;
    if (s0 >= sp.n) return;
    const uint64_t s = s0 + lane;
    const uint64_t my_off = out_off[min(s, sp.n)], my_end = out_off[min(s + 1, sp.n)];
    uint32_t t = 0, start = 0, len = 0;
    uint64_t trb = 0;
    if (s < sp.n && my_end > my_off) { t = sp.tr[s]; start = sp.start[s]; len = sp.len[s]; trb = tr_out_base[t]; }
    const uint64_t base = __shfl_sync(0xffffffffu, my_off, 0);
    const uint64_t total = __shfl_sync(0xffffffffu, my_end, 31) - base;
    for (uint64_t r0 = 0; r0 < total; r0 += 32) {
        const uint64_t o = base + r0 + lane;
        const bool active = r0 + lane < total;
        // owner = the last lane whose first record is at or before o (lanes of species with no record share an offset)
        uint32_t lo = 0, hi = 31;
#pragma unroll
        for (int it = 0; it < 5; it++) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            const uint64_t v = __shfl_sync(0xffffffffu, my_off, mid);
            if (v <= o) lo = mid; else hi = mid - 1;
        }
        const uint64_t ooff = __shfl_sync(0xffffffffu, my_off, lo);
        const uint32_t ot = __shfl_sync(0xffffffffu, t, lo), ostart = __shfl_sync(0xffffffffu, start, lo), olen = __shfl_sync(0xffffffffu, len, lo);
        const uint64_t otrb = __shfl_sync(0xffffffffu, trb, lo);
        if (!active) continue;
        const uint64_t j = o - ooff;
        const uint32_t end = ostart + olen;
        PxStream str{m.key_strand, m.first_index + ot, ostart, end};
        const bool minus = px_u01(px_draw64(str, j)) < m.strand_bias;  // sampler.go:219-227
        out.tr[o] = m.first_index + ot;  // global transcript index
        out.start[o] = ostart;
        out.end[o] = end;
        out.minus[o] = minus ? 1 : 0;
        out.id[o] = o - otrb;  // per-transcript fragment id (sampler.go:189,204)
        // packed host format: length with the strand in the top bit, 2 bytes when every length fits 15 bits
        if (out.ls16) out.ls16[o] = (uint16_t)(olen | (minus ? 0x8000u : 0u));
        else if (out.ls32) out.ls32[o] = olen | (minus ? 0x80000000u : 0u);
    }
}
void launch_expand(cudaStream_t st, const DevModel &m, const DevSpecies &sp, const uint64_t *out_off, uint64_t *tr_out_base,
                   uint64_t n_out, DevSampled out) {
    if (!n_out) return;
    k_tr_out_base<<<(unsigned)((sp.n + 255) / 256), 256, 0, st>>>(sp, out_off, tr_out_base);
    if (out.tr_counts) k_tr_counts<<<(unsigned)((sp.n + 255) / 256), 256, 0, st>>>(sp, out_off, tr_out_base, out.tr_counts);
    k_expand<<<(unsigned)((sp.n + 255) / 256), 256, 0, st>>>(m, sp, out_off, tr_out_base, n_out, out);  // a thread per species slot
}

}  // namespace rl

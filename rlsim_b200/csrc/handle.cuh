// Host-side state of one rlsim_handle and the helpers shared by the translation units of the C ABI (api.cu, comm.cu).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include "kernels.cuh"

using namespace rl;

struct ncclComm;  // opaque: NCCL is bound at run time (comm.cu)

namespace {

// Host array in pinned memory with the few std::vector operations the bookkeeping needs: copies from it to the device
// are real asynchronous DMA (a pageable source would be staged by the driver, with the host waiting for the staging).
template <typename T>
struct PinVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    PinVec() = default;
    PinVec(const PinVec &) = delete;
    PinVec &operator=(const PinVec &) = delete;
    ~PinVec() { if (p) cudaFreeHost(p); }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T *data() { return p; }
    const T *data() const { return p; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    T *begin() { return p; }
    T *end() { return p + n; }
    const T *begin() const { return p; }
    const T *end() const { return p + n; }
    void clear() { n = 0; }
    void reserve(size_t want) {
        if (want <= cap) return;
        size_t c = std::max<size_t>(want + want / 2, 1024);
        T *q = nullptr;
        if (cudaHostAlloc((void **)&q, c * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
        if (n) memcpy(q, p, n * sizeof(T));
        if (p) cudaFreeHost(p);
        p = q; cap = c;
    }
    void resize(size_t want) { reserve(want); if (want > n) memset(p + n, 0, (want - n) * sizeof(T)); n = want; }
    void resize_uninit(size_t want) { reserve(want); n = want; }  // the caller overwrites every new element
    void push_back(const T &v) { reserve(n + 1); p[n++] = v; }
};

template <typename T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t ensure(size_t want) {  // contents are NOT preserved; grows with 1/8 slack so sizes that
        if (p && want <= n) return cudaSuccess;  // fluctuate from run to run do not reallocate every time
        release();
        size_t cap = want + want / 8 + 16;
        cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
        if (e != cudaSuccess) { cap = std::max<size_t>(want, 1); e = cudaMalloc(&p, cap * sizeof(T)); }
        if (e == cudaSuccess) n = cap;
        return e;
    }
    // grow to at least `want` elements keeping the first `used` (device-to-device copy on `st`)
    cudaError_t grow_preserve(size_t want, size_t used, cudaStream_t st) {
        if (p && want <= n) return cudaSuccess;
        size_t cap = want + want / 2 + 16;
        T *q = nullptr;
        cudaError_t e = cudaMalloc(&q, cap * sizeof(T));
        if (e != cudaSuccess) { cap = want; e = cudaMalloc(&q, cap * sizeof(T)); }
        if (e != cudaSuccess) return e;
        if (p && used) {
            e = cudaMemcpyAsync(q, p, std::min(used, n) * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = q; n = cap;
        return e;
    }
};

struct Staging {  // reusable device staging of the batch being ingested (kept across rlsim_clear_transcripts)
    DBuf<uint8_t> bases;
    DBuf<uint64_t> off;
};

// A small team of host threads that memcpy slices of one buffer cooperatively.  Callers hand the library pageable memory
// (a Go slice, a numpy array): DMA needs pinned pages, and the driver's own staging of a pageable cudaMemcpyAsync is
// single-threaded and synchronous.  The library stages through a ring of pinned slots instead, filled (or drained) by this
// team at several times the single-thread memcpy rate while the previous slot is on the wire.
extern "C" void rl_stream_copy(void *dst, const void *src, size_t n);  // hostcopy.cpp: memcpy with non-temporal stores
class CopyTeam {
public:
    explicit CopyTeam(int threads) : T(std::max(1, threads)) {
        for (int j = 1; j < T; j++) th.emplace_back([this, j] { worker(j); });
    }
    ~CopyTeam() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
    int size() const { return T; }
    void copy(void *dst_, const void *src_, size_t n_) {  // returns when all n bytes are in place
        if (T == 1 || n_ < (256u << 10)) { rl_stream_copy(dst_, src_, n_); return; }
        { std::lock_guard<std::mutex> lk(m); dst = (uint8_t *)dst_; src = (const uint8_t *)src_; n = n_; pending = T - 1; gen++; }
        cv.notify_all();
        slice(0, (uint8_t *)dst_, (const uint8_t *)src_, n_);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [this] { return pending == 0; });
    }
private:
    void slice(int j, uint8_t *d, const uint8_t *s, size_t nn) const {
        size_t per = ((nn + T - 1) / T + 4095) & ~(size_t)4095, a = std::min(nn, per * j), b = std::min(nn, a + per);
        if (b > a) rl_stream_copy(d + a, s + a, b - a);
    }
    void worker(int j) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            uint8_t *d = dst; const uint8_t *s_ = src; size_t nn = n;
            lk.unlock();
            slice(j, d, s_, nn);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    int T;
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, cv_done;
    uint8_t *dst = nullptr; const uint8_t *src = nullptr; size_t n = 0;
    uint64_t gen = 0; int pending = 0; bool stop = false;
};

inline int stage_threads() {  // RLSIM_STAGE_THREADS, default min(6, cores - 1)
    if (const char *e = getenv("RLSIM_STAGE_THREADS")) return std::max(1, atoi(e));
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(6u, std::max(1u, hw > 1 ? hw - 1 : 1u));
}

// does the host pointer refer to pinned (page-locked / registered) memory, i.e. can the copy engine read it directly?
inline bool host_pinned(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

inline uint32_t bits_for(uint64_t v) {
    uint32_t b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

// kernels launched by one cub::DeviceRadixSort call over `bits` key bits (histogram + exclusive sum + onesweep passes)
inline uint64_t radix_launches(int bits) { return 2 + (uint64_t)((bits + 7) / 8); }

}  // namespace

struct PendingTime { int stage; cudaEvent_t a, b; };  // a stage timer whose events have not been resolved yet
struct rlsim_handle {
    int device = 0;
    cudaStream_t st = nullptr;
    std::string err;
    rlsim_params p{};
    bool have_model = false;
    // target
    std::vector<uint32_t> raw_len;
    std::vector<double> raw_prob;
    std::vector<uint64_t> h_target;  // histogram by length
    bool have_target = false;
    uint64_t req_frags = 0;
    uint64_t low = 0, high = 0;
    double raw_location = 0.0;
    uint32_t polya_max = 0;
    std::vector<double> h_len_eff;  // host-supplied table
    uint64_t len_eff_low = 0, len_eff_high = 0;
    // transcripts
    uint32_t first_index = 0;
    Staging staging;
    cudaStream_t st_copy = nullptr;  // H2D of the next chunk overlaps ingest/prefix of the previous one
    cudaStream_t st_pcr = nullptr;   // eager pipeline: PCR of group g (FP64 issue bound) overlaps the memory-bound stages of group g+1
    cudaEvent_t ev_sp = nullptr, ev_pcr = nullptr, ev_ingest = nullptr, ev_aux = nullptr;
    cudaStream_t st_aux = nullptr;   // target-length sampling runs beside the transcript pipeline
    bool target_pending = false;     // the sampled histogram is still on the device (target_host)
    uint32_t target_hist_n = 0;
    bool pcr_in_flight = false;
    unsigned long long *pin = nullptr;  // pinned scratch for the small device-to-host count reads
    // staging of pageable caller buffers: ring of pinned slots + a team of copy threads (created on first use)
    uint8_t *stage_ring = nullptr; size_t stage_ring_bytes = 0;
    CopyTeam *team = nullptr;
    std::vector<cudaEvent_t> ev_slot;
    std::vector<cudaEvent_t> ev_chunk, ev_free;
    std::vector<PendingTime> pending_times;  // stage timers not yet resolved (drain_timers)
    std::vector<uint64_t> h_srclen;  // unpadded lengths as added
    PinVec<uint64_t> h_expr, h_rel;   // pinned: sources of the per-batch metadata copies
    PinVec<uint64_t> h_off;          // padded offsets [n+1]
    PinVec<uint32_t> h_len;          // with polyA
    uint64_t n_mol = 0, total_padded = 0;
    double sum_expr = 0, sum_expr_len = 0;  // for the fragment-capacity estimate
    bool mol_off_valid = false;
    PinVec<uint64_t> h_mol_off;             // exclusive scan of the expression levels [n+1]
    // eager pipeline: while later chunks are still being copied, earlier chunks are already fragmented and amplified
    uint32_t eager_upto = 0;                // transcripts [0, eager_upto) are done under the current model
    bool eager_failed = false;              // a buffer estimate was too small: the monolithic path takes over
    bool eager_started = false;             // counters / histograms zeroed for this transcript set
    bool eager_pcr = false;                 // species counts already amplified (consumed by the next rlsim_pcr)
    bool eager_used = false;                // the current species table came from the eager pipeline (rlsim_stats.flags)
    uint32_t eager_groups = 0;              // groups started since the last reset (test hook)
    bool staged_pageable = false;           // the last add call staged a pageable caller buffer through pinned memory
    uint64_t eager_keys = 0, eager_sp = 0;  // running fragment-key / species counts
    // asynchronous form of the pipeline (default): per-chunk counts stay on the device, the host never waits inside a batch
    DBuf<rl::EagerChunk> d_ecs;             // one per chunk handed to the pipeline since the last reset
    DBuf<rl::EagerGlobal> d_eglobal;        // species cursor, fragment total, overflow flags
    DBuf<uint32_t> d_rle_cnt;               // run lengths of the chunk being deduplicated
    uint32_t ecs_used = 0;                  // chunks handed out
    uint64_t eager_sp_ub = 0;               // upper bound of the species placed so far (sum of the chunks' fragment counts)
    bool eager_async = false;               // this transcript set runs through the asynchronous form
    cudaStream_t st_back = nullptr;         // back half of a chunk (sort, run-length pass, species decode)
    cudaStream_t st_in = nullptr;           // ingest + prefix sums of a chunk, beside the fragmentation of the chunk before
    std::vector<cudaEvent_t> ev_in;         // per chunk of a batch: ingest + prefix sums queued
    std::vector<cudaEvent_t> ev_cnt, ev_sorted;  // per chunk: front half done (counts on their way) / key buffer read
    PinVec<rl::EagerChunk> ecs_pin;         // landing zone of the chunks' records
    struct ChunkBound { uint64_t bound, q_cap; uint32_t long_cap, t1; };
    std::vector<ChunkBound> ecs_bound;
    DBuf<uint64_t> d_ekeys[rl::EAGER_BUFS];     // fragment keys of the chunks between their front half and the end of their sort
    struct TraceMark { const char *what; int chunk; cudaEvent_t ev; };
    std::vector<TraceMark> trace_marks;     // RLSIM_TRACE=2: device-side timeline of the chunk pipeline (timing events)
    void mark(const char *what, int chunk, cudaStream_t s) {
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); trace_marks.push_back({what, chunk, e});
    }
    DBuf<uint2> d_elong_list[rl::EAGER_BUFS];   // deferred (long) molecules of those chunks
    DBuf<uint32_t> d_elong_hi[rl::EAGER_BUFS];
    bool len_eff_ready = false;
    DBuf<uint8_t> d_bases;
    DBuf<uint32_t> d_packed, d_nmask, d_len;
    DBuf<uint2> d_gc;  // GC groups: {prefix count, bit mask} per 32 bases
    // device FASTA reader (rlsim_add_fasta)
    DBuf<uint8_t> d_fa_text, d_fa_tiles, d_fa_seq, d_fa_seq2, d_fa_valid;
    DBuf<uint4> d_fa_carry;
    DBuf<unsigned long long> d_fa_tot;
    DBuf<uint64_t> d_fa_rec_start, d_fa_hdr_end, d_fa_seq_off, d_fa_level, d_fa_seq_len;
    DBuf<uint32_t> d_fa_hdr_len, d_fa_name_len;
    std::vector<uint64_t> fa_start, fa_level, fa_seq_len;
    std::vector<uint32_t> fa_hdr_len, fa_name_len;
    std::vector<uint8_t> fa_valid;
    uint32_t fa_n = 0;
    DBuf<uint64_t> d_off, d_expr, d_mol_off, d_Cf, d_Cr;  // coarse prefix sums
    DBuf<uint32_t> d_Ff, d_Fr;                          // fine (in-block) prefix sums
    DBuf<uint32_t> d_warp_t;                            // first transcript of every warp of k_fragment (k_warp_transcripts)
    uint64_t warp_t_g0 = 0, warp_t_g1 = 0; uint32_t warp_t_n = 0;  // what the table covers (molecule range, transcripts)
    DBuf<Rec32> d_Rf;                                   // record layout of the forward sums (clean transcripts; common.cuh)
    bool rec_layout = false;                            // records in use for the current model (symmetric table, k <= 6)
    uint32_t prefix_upto = 0;  // transcripts [0, prefix_upto) have valid priming prefix sums IN HBM for the current model
    // fused path (after_prim / after_prim_double): prefix sums in shared memory, work items = runs of transcripts
    PinVec<FusedItem> h_fitems_s, h_fitems_l;     // small-class / large-class items of the last build (CTA-cooperative kernel)
    DBuf<FusedItem> d_fitems_s, d_fitems_l;
    DBuf<FusedItem> d_redo;                       // what the warp kernel passed on to the CTA-cooperative one
    uint32_t fitems_t0 = 0, fitems_t1 = 0; int fitems_key = -1;  // what the device item lists cover (transcript range, layout key)
    bool fused_disabled = false;                  // a transcript did not fit shared memory: this library stays on the HBM path
    bool fused_used = false;
    uint32_t max_len = 0;                         // longest transcript (with tail) added so far
    int sm_count = 148;
    bool lut_valid = false;
    int lut_k = 0; double lut_T = 0; bool lut_sym = false, want_mirror = true;  // cached k-mer table; reverse-complement symmetric after quantisation
    DBuf<uint32_t> d_dirty;  // per transcript: letters outside ACGT present
    // model tables
    DBuf<double> d_gc_raw, d_len_eff, d_raw_prob, d_K;
    DBuf<uint32_t> d_raw_len, d_wq_f, d_wq_r;
    DBuf<unsigned long long> d_kmax;
    // fragmentation
    DBuf<uint64_t> d_keys, d_keys_sorted, d_uniq;
    DBuf<unsigned long long> d_counters;  // [0] fragment count, [1] run count
    DBuf<unsigned long long> d_hist_frag, d_hist_polya, d_hist_target;
    DBuf<uint2> d_long_list, d_q_b;
    DBuf<uint4> d_q_a;
    DBuf<uint32_t> d_long_hi;
    DBuf<unsigned int> d_long_count, d_work;  // d_work: [0..7] fused-path counters and flags, [8..11] k_prefix work counters
    DBuf<uint32_t> d_pfx_long;               // transcripts k_prefix leaves to k_prefix_long (2 strands)
    DBuf<int> d_error;
    DBuf<unsigned char> d_cub;
    uint64_t n_frag = 0;
    bool fragmented = false, amplified = false, sampled = false;
    std::vector<uint64_t> h_hist_frag, h_hist_polya;
    // species
    uint64_t n_sp = 0;
    DBuf<uint32_t> d_sp_tr, d_sp_start, d_sp_len, d_sp_count0, d_sp_gc;
    DBuf<uint64_t> d_sp_count;
    DBuf<double> d_sp_eff;
    bool keep_debug = true;  // keep per-species eff / gc for rlsim_get_species
    // classes
    DBuf<uint32_t> d_iota, d_sorted_len, d_order;
    DBuf<uint64_t> d_cls_cnt, d_cls_pre, d_bounds, d_totals;
    uint32_t n_len = 0;
    std::vector<uint64_t> h_totals;  // local per-length totals
    std::vector<uint64_t> h_global_totals;  // summed over the ranks (rlsim_sample_distributed)
    uint64_t pool_total = 0;
    // selection
    std::vector<uint64_t> h_sampled, h_missing, h_after_sampling, h_after_pcr;
    DBuf<uint64_t> d_cand_keys, d_cand_keys_sorted, d_fscan, d_taken, d_out_off, d_plan_u64, d_plan_cut;
    DBuf<uint32_t> d_cand_idx, d_cand_idx_sorted, d_flags, d_plan_len;
    DBuf<unsigned long long> d_hits;
    DBuf<uint8_t> d_mode;
    DBuf<unsigned long long> d_emit;  // [0..64) counts, [64..128) segment offsets, [128..192) cursors
    DBuf<uint64_t> d_all_base, d_my_base;
    DBuf<int> d_short;
    uint64_t n_out = 0;
    DBuf<uint32_t> d_o_tr, d_o_start, d_o_end;
    DBuf<uint8_t> d_o_minus;
    DBuf<uint64_t> d_o_id, d_tr_out_base;
    DBuf<uint16_t> d_o_ls16;        // packed host format: length | strand << 15 (lengths below 2^15)
    DBuf<uint32_t> d_o_ls32, d_o_trcnt;  // ... or length | strand << 31; records per transcript
    DBuf<uint32_t> d_cand_k32, d_cand_k32_sorted;  // narrow sort keys of the selection candidates
    // fasta
    DBuf<uint8_t> d_names, d_fasta;
    DBuf<uint64_t> d_name_off, d_rec_sizes, d_rec_off;
    // multi-GPU exchange inside the library (comm.cu): NCCL communicator bound at run time, buffers of the exchange
    bool rec8 = false;                   // the exchange in flight uses 8-byte records (length << 40 | rank; all class totals < 2^40)
    ncclComm *comm = nullptr;
    int comm_rank = 0, comm_n = 1;
    DBuf<uint64_t> d_x_send, d_x_recv;   // [totals | partial target] rows, count rows
    DBuf<uint4> d_rec_send, d_rec_recv;  // selection records {len, 0, rank_lo, rank_hi}
    uint64_t *x_pin = nullptr; size_t x_pin_n = 0;  // pinned landing zone of the gathered rows
    // timings
    float ms[RLSIM_T_COUNT] = {0};
    uint64_t launches[RLSIM_T_COUNT] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

#define FAIL(h, code, msg) do { (h)->err = (msg); return (code); } while (0)
#define CK(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__); return RLSIM_E_CUDA; } } while (0)

// ---- host functions with one definition in api.cu, used by comm.cu as well
namespace rl_api {
struct StageTimer {  // stage timing never blocks the host: event pairs are parked and resolved when the timings are read
    rlsim_handle *h;
    int stage;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(rlsim_handle *h_, int s, uint64_t nl, cudaStream_t stream = nullptr);
    void stop();
};
void drain_timers(rlsim_handle *h);
int check_device_error(rlsim_handle *h);
int cub_temp(rlsim_handle *h, size_t bytes);
int target_host(rlsim_handle *h);
int d2h_any(rlsim_handle *h, void *dst, const void *src, size_t bytes);
void set_target_from_hist(rlsim_handle *h, const uint64_t *hist, uint32_t n);
// seg_out != null (the library's own exchange): one sweep into owner groups whose capacities come from the expected split
// (all_tot = every rank's per-length totals, [nparts][n_len]); seg_out[r] = where rank r's group starts in `records`
int select_core(rlsim_handle *h, const std::vector<uint64_t> &T, const std::vector<uint64_t> &base, uint32_t nparts,
                uint32_t part, const uint64_t *d_all_base, uint4 *records, uint64_t cap, uint64_t *counts_out,
                uint64_t *needed_out, bool sync_records = true, const uint64_t *all_tot = nullptr, uint64_t *seg_out = nullptr);
int select_partitioned(rlsim_handle *h, const uint64_t *all_totals, uint32_t nparts, uint32_t part, uint32_t n_len_in,
                       void *records_dev, uint64_t cap, uint64_t *counts_out, uint64_t *needed_out, bool sync_records,
                       uint64_t *seg_out = nullptr);
int finish_sample(rlsim_handle *h);
int apply_records(rlsim_handle *h, const void *records_dev, uint64_t n);
void comm_destroy(rlsim_handle *h);  // comm.cu
}  // namespace rl_api

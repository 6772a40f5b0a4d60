// C ABI of librlsim_gpu.so (include/rlsim_gpu.h): handle, HBM layout, stage orchestration.
//
// Host glue only - every per-molecule / per-species / per-fragment computation is in the k_*.cu kernels.
// Library plumbing used here: cub::DeviceRadixSort / DeviceRunLengthEncode / DeviceScan for the species
// dedup (transcript.go:156-215), the per-length class view (pool.go:156-189) and the selection sort.
#include <cub/cub.cuh>
#include "handle.cuh"

namespace {

__global__ void k_iota(uint32_t *p, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void k_class_totals(const uint64_t *cls_pre, const uint64_t *bounds, uint32_t n_len, uint64_t *totals) {
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n_len) totals[l] = cls_pre[bounds[l + 1]] - cls_pre[bounds[l]];
}

}  // namespace

static std::string g_create_error;

namespace rl_api {

StageTimer::StageTimer(rlsim_handle *h_, int s, uint64_t nl, cudaStream_t stream) : h(h_), stage(s), st(stream ? stream : h_->st) {
    for (cudaEvent_t *e : {&a, &b}) {
        if (!h->ev_free.empty()) { *e = h->ev_free.back(); h->ev_free.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(a, st);
    h->launches[s] += nl;
}
void StageTimer::stop() {
    cudaEventRecord(b, st);
    h->pending_times.push_back(PendingTime{stage, a, b});
    if (h->pending_times.size() > 512) drain_timers(h);
}
void drain_timers(rlsim_handle *h) {
    for (auto &p : h->pending_times) {
        float t = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) h->ms[p.stage] += t;
        h->ev_free.push_back(p.a); h->ev_free.push_back(p.b);
    }
    h->pending_times.clear();
}

int check_device_error(rlsim_handle *h) {
    int e = 0;
    CK(h, cudaMemcpyAsync(&e, h->d_error.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    switch (e) {
    case 0: return 0;
    case RLSIM_E_PRIMING: FAIL(h, e, "Error when simulating priming!");
    case RLSIM_E_PCR_OVERFLOW: FAIL(h, e, "Integer overflow detected when amplifying fragments!");
    case RLSIM_E_UNSUPPORTED: FAIL(h, e, "a molecule needs more breakpoints than the shared-memory sorter holds (32766)");
    default: FAIL(h, e, "device-side error");
    }
}

void mix_minmax(const rlsim_mixcomp *c, int n, uint64_t *mn, uint64_t *mx) {  // target.go:72-87
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i < n; i++) {
        if (i == 0) lo = c[i].low;
        if (hi < c[i].high) hi = c[i].high;
        if (lo > c[i].low) lo = c[i].low;
    }
    *mn = lo; *mx = hi;
}

bool needs_fwd(int m) { return m == RLSIM_AFTER_PRIM || m == RLSIM_AFTER_PRIM_DOUBLE || m == RLSIM_PRIM_JUMP; }
bool needs_rev(int m) { return m == RLSIM_AFTER_PRIM_DOUBLE; }

DevModel make_model(rlsim_handle *h) {
    DevModel m{};
    const rlsim_params &p = h->p;
    m.k = p.kmer_len; m.method = p.frag_method; m.frag_param = p.frag_param; m.cycles = p.nr_cycles;
    m.T = p.priming_temp; m.loss = p.frag_loss_prob; m.fixed_eff = p.fixed_eff; m.strand_bias = p.strand_bias;
    m.gc_mode = p.gc_mode; m.n_target = p.n_target; m.n_polya = p.n_polya; m.n_raw = (int32_t)h->raw_len.size();
    m.gc_shape = p.gc_shape; m.gc_min = p.gc_min; m.gc_max = p.gc_max;
    m.len_shape = p.len_shape; m.len_min = p.len_min; m.len_max = p.len_max;
    memcpy(m.target, p.target, sizeof m.target);
    memcpy(m.polya, p.polya, sizeof m.polya);
    m.raw_location = h->raw_location;
    m.low = (uint32_t)h->low; m.high = (uint32_t)h->high; m.polya_max = h->polya_max; m.first_index = h->first_index;
    // seed policy of main.go:72-81,129-149
    int64_t si = p.seed_init, sp = p.seed_pcr ? p.seed_pcr : si, ss = p.seed_sampling ? p.seed_sampling : sp;
    m.key_target = px_derive_key(si, TAG_TARGET);
    m.key_frag = px_derive_key(si, TAG_FRAG);
    m.key_pcr = px_derive_key(sp, TAG_PCR);
    m.key_select = px_derive_key(ss, TAG_SELECT);
    m.key_strand = px_derive_key(ss, TAG_STRAND);
    m.gc_raw = h->d_gc_raw.p; m.len_eff = h->d_len_eff.p; m.raw_len = h->d_raw_len.p; m.raw_prob = h->d_raw_prob.p;
    return m;
}

DevTranscripts make_tr(rlsim_handle *h) {
    DevTranscripts t{};
    t.n = (uint32_t)h->h_len.size();
    t.bases = h->d_bases.p; t.packed = h->d_packed.p; t.nmask = h->d_nmask.p; t.off = h->d_off.p; t.len = h->d_len.p;
    t.expr = h->d_expr.p; t.mol_off = h->d_mol_off.p;
    t.Ff = needs_fwd(h->p.frag_method) ? h->d_Ff.p : nullptr;
    t.Cf = needs_fwd(h->p.frag_method) ? h->d_Cf.p : nullptr;
    t.Fr = needs_rev(h->p.frag_method) ? h->d_Fr.p : nullptr;
    t.Cr = needs_rev(h->p.frag_method) ? h->d_Cr.p : nullptr;
    t.gc = h->d_gc.p;
    t.dirty = h->d_dirty.p;
    // record layout of the forward sums: needs the mirrored reverse draw (symmetric table) and a LUT that fits shared memory
    h->rec_layout = needs_fwd(h->p.frag_method) && h->lut_valid && h->lut_sym && h->want_mirror && h->p.kmer_len <= 6 &&
                    getenv("RLSIM_RECORDS") != nullptr;  // opt-in: measured slower than the two-level arrays (profiles/r2_experiments.md)
    t.Rf = h->rec_layout ? h->d_Rf.p : nullptr;
    t.lut = h->d_wq_f.p;
    t.sym = h->lut_sym && h->want_mirror ? 1 : 0;
    return t;
}

DevSpecies make_sp(rlsim_handle *h) {
    DevSpecies s{};
    s.n = h->n_sp; s.tr = h->d_sp_tr.p; s.start = h->d_sp_start.p; s.len = h->d_sp_len.p; s.count0 = h->d_sp_count0.p;
    s.count = h->d_sp_count.p; s.eff = h->keep_debug ? h->d_sp_eff.p : nullptr; s.gc = h->keep_debug ? h->d_sp_gc.p : nullptr;
    return s;
}

int raw_target_finish(rlsim_handle *h) {  // raw_target.go:54,59-90
    bool first = true;
    uint64_t mn = 0, mx = 0;
    double mean = 0.0, t = (double)h->req_frags;
    for (size_t l = 0; l < h->h_target.size(); l++) {
        uint64_t c = h->h_target[l];
        if (!c) continue;
        if (first) { mn = mx = l; first = false; }
        mn = std::min<uint64_t>(mn, l); mx = std::max<uint64_t>(mx, l);
        mean += ((double)c / t) * (double)l;
    }
    h->low = mn; h->high = mx; h->raw_location = mean;
    return 0;
}

// capacity of the two-level prefix-sum arrays for `tot` padded bases in `n` transcripts, keeping the first old_tot / old_n
int grow_prefix(rlsim_handle *h, uint64_t tot, uint32_t n, uint64_t old_tot, uint32_t old_n) {
    CK(h, h->d_Rf.grow_preserve(tot / 64 + 2ull * n + 8, old_n ? old_tot / 64 + 2ull * old_n + 2 : 0, h->st));
    CK(h, h->d_Ff.grow_preserve(tot + 64, old_tot, h->st));
    CK(h, h->d_pfx_long.ensure(2ull * n + 2));  // k_prefix's list of long transcripts, per strand
    CK(h, h->d_Cf.grow_preserve(tot / 32 + 2ull * n + 8, old_n ? old_tot / 32 + 2ull * old_n + 2 : 0, h->st));
    if (needs_rev(h->p.frag_method)) {
        CK(h, h->d_Fr.grow_preserve(tot + 64, old_tot, h->st));
        CK(h, h->d_Cr.grow_preserve(tot / 32 + 2ull * n + 8, old_n ? old_tot / 32 + 2ull * old_n + 2 : 0, h->st));
    }
    return 0;
}

// priming affinities -> prefix sums for transcripts [t0, t1) (nnthermo.go:177-201); launches only
int ensure_lut(rlsim_handle *h) {
    if (!h->lut_valid) {
        uint32_t nk = 1u << (2 * h->p.kmer_len);
        CK(h, h->d_K.ensure(nk)); CK(h, h->d_wq_f.ensure(nk)); CK(h, h->d_wq_r.ensure(nk));
        launch_kmer_lut(h->st, h->p.kmer_len, h->p.priming_temp, h->d_K.p, h->d_kmax.p, h->d_wq_f.p, h->d_wq_r.p, h->d_work.p + 2);
        CK(h, cudaMemcpyAsync(h->pin + 6, h->d_work.p + 2, 4, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        h->lut_sym = *reinterpret_cast<unsigned int *>(h->pin + 6) == 0;
        h->lut_valid = true; h->lut_k = h->p.kmer_len; h->lut_T = h->p.priming_temp;
        h->launches[RLSIM_T_PREFIX] += 3;
    }
    return 0;
}
int launch_prefix_range(rlsim_handle *h, uint32_t t0, uint32_t t1, cudaStream_t st = nullptr) {
    int method = h->p.frag_method;
    int rcl = ensure_lut(h);  // (on h->st; a caller with its own stream has made it wait for that)
    if (rcl) return rcl;
    const DevTranscripts trp = make_tr(h);
    launch_prefix(st ? st : h->st, trp, t0, t1, h->p.kmer_len, h->p.priming_temp, h->d_wq_f.p, h->d_wq_r.p, h->d_kmax.p,
                  h->d_Ff.p, h->d_Fr.p, h->d_Cf.p, h->d_Cr.p, const_cast<Rec32 *>(trp.Rf), needs_rev(method) ? 2 : 1, h->d_work.p + 8,
                  // opt-in (RLSIM_PREFIX_LONG=1): long transcripts summed by whole CTAs (k_prefix_long).  Measured without
                  // effect on the chunk pipeline (the one-warp tail of a launch overlaps the other streams' kernels) and
                  // 0.06 ms slower over a whole transcriptome, where enough warps are in flight to hide it.
                  getenv("RLSIM_PREFIX_LONG") ? h->d_pfx_long.p : nullptr, (uint32_t)(h->d_pfx_long.n / 2),
                  getenv("RLSIM_NO_PREFIX_BULK") == nullptr);  // fine sums leave shared memory as bulk asynchronous copies
    h->launches[RLSIM_T_PREFIX] += 2;
    return 0;
}

// Device -> caller's host buffer on h->st.  Pinned destinations are written by the copy engine directly (asynchronous:
// the caller synchronises); pageable ones are staged through the pinned ring and drained by the copy team while the next
// slot is on the wire (complete on return).
int d2h_any(rlsim_handle *h, void *dst_, const void *src_, size_t bytes) {
    if (!bytes || !dst_) return 0;
    uint8_t *dst = (uint8_t *)dst_;
    const uint8_t *src = (const uint8_t *)src_;
    if (bytes < (1u << 20) || getenv("RLSIM_NO_STAGING") || host_pinned(dst)) {
        CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st));
        return 0;
    }
    constexpr size_t SLOT = 8u << 20;
    constexpr uint32_t RING = 3;
    if (h->stage_ring_bytes < RING * SLOT) {
        if (h->stage_ring) cudaFreeHost(h->stage_ring);
        h->stage_ring = nullptr; h->stage_ring_bytes = 0;
        CK(h, cudaHostAlloc((void **)&h->stage_ring, RING * SLOT, cudaHostAllocDefault));
        h->stage_ring_bytes = RING * SLOT;
    }
    while (h->ev_slot.size() < RING) { cudaEvent_t e; CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_slot.push_back(e); }
    if (!h->team) h->team = new CopyTeam(stage_threads());
    const size_t nch = (bytes + SLOT - 1) / SLOT;
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t a = c * SLOT, len = std::min(SLOT, bytes - a);
        cudaError_t e = cudaMemcpyAsync(h->stage_ring + (c % RING) * SLOT, src + a, len, cudaMemcpyDeviceToHost, h->st);
        return e != cudaSuccess ? e : cudaEventRecord(h->ev_slot[c % RING], h->st);
    };
    for (size_t c = 0; c < std::min<size_t>(nch, RING - 1); c++) CK(h, issue(c));
    for (size_t c = 0; c < nch; c++) {
        if (c + RING - 1 < nch) CK(h, issue(c + RING - 1));  // its slot was drained in the previous iteration
        CK(h, cudaEventSynchronize(h->ev_slot[c % RING]));
        const size_t a = c * SLOT, len = std::min(SLOT, bytes - a);
        h->team->copy(dst + a, h->stage_ring + (c % RING) * SLOT, len);
    }
    return 0;
}

#define D2H(dst, src, bytes) do { if (dst) CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st)); } while (0)

void eager_reset(rlsim_handle *h) {
    h->fused_used = false;
    h->eager_upto = 0; h->eager_failed = h->eager_started = h->eager_pcr = false; h->eager_keys = h->eager_sp = 0;
    h->eager_groups = 0;
    h->ecs_used = 0; h->eager_sp_ub = 0; h->eager_async = false;
}

double frag_denominator(rlsim_handle *h) {  // expected fragments per molecule <= L / denominator + 1
    int method = h->p.frag_method;
    double loc = h->raw_location;
    if (h->p.n_target > 0) {
        double lmin = 1e300;
        for (int i = 0; i < h->p.n_target; i++) lmin = std::min(lmin, h->p.target[i].location);
        loc = lmin;
    }
    if (!(loc > 0)) loc = 1.0;
    return method == RLSIM_PRIM_JUMP ? (double)std::max<uint64_t>(h->low, 1)
         : method == RLSIM_PRE_PRIM ? loc
         : (h->p.frag_param == 0 ? 2.0 * loc : (double)h->p.frag_param);
}

// Techne.CalcLengthEff table on the device: the host's own (rlsim_set_len_eff) or the library's
int ensure_len_eff(rlsim_handle *h) {
    if (h->len_eff_ready) return 0;
    uint32_t nle = (uint32_t)(h->high - h->low + 1);
    CK(h, h->d_len_eff.ensure(nle));
    if (!h->h_len_eff.empty()) {
        if (h->len_eff_low != h->low || h->len_eff_high != h->high) FAIL(h, RLSIM_E_ARG, "length-efficiency table does not span [low, high]");
        CK(h, cudaMemcpyAsync(h->d_len_eff.p, h->h_len_eff.data(), (size_t)nle * 8, cudaMemcpyHostToDevice, h->st));
    } else {
        launch_len_eff(h->st, make_model(h), h->d_len_eff.p);
    }
    h->len_eff_ready = true;
    return 0;
}

bool eager_possible(rlsim_handle *h, uint32_t n0) {
    const bool disabled = getenv("RLSIM_NO_EAGER") != nullptr;  // diagnostic switch: always take the monolithic path
    return !disabled && h->have_model && h->p.frag_method <= RLSIM_AFTER_NOPRIM_DOUBLE && (h->p.n_target > 0 || h->have_target) &&
           h->low > 0 && h->high >= h->low && !h->eager_failed && h->eager_upto == n0;
}

bool ns_need(rlsim_handle *h, uint64_t ns) {  // would any per-species array be reallocated for ns species?
    return ns > h->d_sp_tr.n || ns > h->d_sp_start.n || ns > h->d_sp_len.n || ns > h->d_sp_count.n || ns > h->d_sp_eff.n || ns > h->d_sp_gc.n;
}

// table of the first transcript of every warp of k_fragment over the molecules [g0, g1): depends on the expression levels
// only, so the resident steps of one transcript set build it once
int ensure_warp_table(rlsim_handle *h, const DevTranscripts &tr, uint64_t g0, uint64_t g1) {
    const uint32_t n = tr.n;
    if (h->warp_t_n == n && h->warp_t_g0 == g0 && h->warp_t_g1 == g1) return 0;
    CK(h, h->d_warp_t.ensure((g1 - g0 + 31) / 32 + 1));
    launch_warp_transcripts(h->st, tr, g0, g1, h->d_warp_t.p);
    h->launches[RLSIM_T_FRAGMENT]++;
    h->warp_t_n = n; h->warp_t_g0 = g0; h->warp_t_g1 = g1;
    return 0;
}

// ---- fused path: prefix sums in shared memory (k_fused).  Eligible: the priming methods whose intervals are searched
// once per breakpoint pair; every transcript of the range must fit one CTA's shared memory and produce few enough
// breakpoints for the deferred-molecule kernel.
bool fused_method(rlsim_handle *h) {
    // opt-in (measured, profiles/r2_experiments.md: at C3 the three HBM-path kernels are faster than either fused variant)
    const bool disabled = getenv("RLSIM_NO_FUSED") != nullptr || getenv("RLSIM_FUSED") == nullptr;
    return !disabled && !h->fused_disabled && (h->p.frag_method == RLSIM_AFTER_PRIM || h->p.frag_method == RLSIM_AFTER_PRIM_DOUBLE);
}
bool fused_len_ok(rlsim_handle *h, uint32_t L) {
    return (((uint64_t)L + 63) & ~63ull) <= FUSED_CAP_LARGE && (double)L / frag_denominator(h) <= (double)FUSED_LONG_BREAKS;
}
bool fused_range_ok(rlsim_handle *h, uint32_t t0, uint32_t t1) {
    if (!fused_method(h)) return false;
    if (fused_len_ok(h, h->max_len)) return true;
    for (uint32_t t = t0; t < t1; t++)
        if (h->h_expr[t] && !fused_len_ok(h, h->h_len[t])) return false;
    return true;
}

// Work items for the transcripts [t0, t1): runs of consecutive transcripts whose prefix sums fit the small class together
// (at most FUSED_MAX molecules each; heavily expressed transcripts are split by molecule range), one transcript per item
// in the large class.  The kernel sub-batches an item if it turns out not to fit (transcripts with letters outside ACGT
// carry explicit reverse arrays, which the host does not know).
int build_fused_items(rlsim_handle *h, uint32_t t0, uint32_t t1) {
    const bool dbl = h->p.frag_method == RLSIM_AFTER_PRIM_DOUBLE;
    const bool sym = h->lut_sym && h->want_mirror;
    const int key = (dbl ? 1 : 0) | (sym ? 2 : 0);
    if (h->fitems_key == key && h->fitems_t0 == t0 && h->fitems_t1 == t1) return 0;  // still on the device
    constexpr uint64_t M_MAX = 4096;
    PinVec<FusedItem> &S = h->h_fitems_s, &Lg = h->h_fitems_l;
    S.clear(); Lg.clear();
    const uint64_t *mo = h->h_mol_off.data();
    uint32_t ts = 0, te = 0, cnt = 0;
    uint64_t sumE = 0, sumM = 0;
    auto flush = [&] { if (cnt) S.push_back(FusedItem{ts, te, mo[ts], mo[te]}); cnt = 0; sumE = 0; sumM = 0; };
    for (uint32_t t = t0; t < t1; t++) {
        const uint64_t ex = h->h_expr[t];
        if (!ex) continue;
        const uint64_t e1 = ((uint64_t)h->h_len[t] + 63) & ~63ull;
        const bool small = (dbl ? 2 : 1) * e1 <= FUSED_CAP_SMALL;  // fits alone even with explicit reverse arrays
        const uint64_t est = e1 * (dbl && !sym ? 2 : 1);
        if (!small || ex > M_MAX) {
            flush();
            PinVec<FusedItem> &dst = small ? S : Lg;
            for (uint64_t g = mo[t]; g < mo[t + 1]; g += M_MAX) dst.push_back(FusedItem{t, t + 1, g, std::min<uint64_t>(g + M_MAX, mo[t + 1])});
            continue;
        }
        if (cnt && (sumE + est > FUSED_CAP_SMALL || sumM + ex > M_MAX || cnt >= 32)) flush();
        if (!cnt) ts = t;
        te = t + 1; sumE += est; sumM += ex; cnt++;
    }
    flush();
    CK(h, h->d_fitems_s.ensure(S.size() + 1)); CK(h, h->d_fitems_l.ensure(Lg.size() + 1));
    if (S.size()) CK(h, cudaMemcpyAsync(h->d_fitems_s.p, S.data(), S.size() * sizeof(FusedItem), cudaMemcpyHostToDevice, h->st));
    if (Lg.size()) CK(h, cudaMemcpyAsync(h->d_fitems_l.p, Lg.data(), Lg.size() * sizeof(FusedItem), cudaMemcpyHostToDevice, h->st));
    h->fitems_key = key; h->fitems_t0 = t0; h->fitems_t1 = t1;
    return 0;
}

// Run the fused kernels over the transcripts [t0, t1) up to (not including) the deferred molecules: on return the host
// knows how many molecules were deferred (*n_long) and whether a transcript did not fit shared memory (*fell_back: the
// caller reruns the range on the HBM path).
int run_fused_range(rlsim_handle *h, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint32_t t0, uint32_t t1,
                    unsigned int *n_long, int *fell_back) {
    int rc;
    if ((rc = ensure_lut(h))) return rc;
    DevTranscripts trs = tr;
    trs.sym = h->lut_sym && h->want_mirror ? 1 : 0;  // ensure_lut may just have established it
    int *fallback = reinterpret_cast<int *>(h->d_work.p + 6);
    unsigned int *redo_count = h->d_work.p + 7;
    CK(h, cudaMemsetAsync(h->d_work.p + 6, 0, 8, h->st));
    const bool warp_kernel = trs.sym && h->p.kmer_len <= 6 && !getenv("RLSIM_FUSED_CTA");
    unsigned int n_redo = 0;
    if (warp_kernel) {
        // no item lists: the molecule range is cut into fixed units, every launch serves one slice-size class
        const uint64_t g0 = h->h_mol_off[t0], g1 = h->h_mol_off[t1];
        CK(h, h->d_redo.ensure((size_t)(t1 - t0) + fused_warp_units(g1 - g0) + 1));
        static const uint32_t caps[4] = {0, FUSED_WARP_CAP_S, FUSED_WARP_CAP_M, FUSED_WARP_CAP_L};
        static const int warps[3] = {8, 8, 4};
        uint32_t longest = 0;
        if (t1 - t0 == (uint32_t)h->h_len.size()) longest = h->max_len;
        else for (uint32_t t = t0; t < t1; t++) if (h->h_expr[t]) longest = std::max(longest, h->h_len[t]);
        for (int c = 0; c < 3; c++) {
            if (c > 0 && (((uint64_t)longest + 63) & ~63ull) <= caps[c]) break;  // no transcript of this size class or a larger one
            if ((rc = launch_fused_warp(h->st, m, trs, fb, g0, g1, caps[c], caps[c + 1], warps[c], h->d_work.p + 3 + c, h->d_wq_f.p, h->d_redo.p,
                                        redo_count, (uint32_t)h->d_redo.n, h->sm_count))) return rc;
            h->launches[RLSIM_T_FRAGMENT]++;
        }
        CK(h, cudaMemcpyAsync(h->pin, h->d_long_count.p, 4, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaMemcpyAsync(h->pin + 7, h->d_work.p + 6, 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        n_redo = reinterpret_cast<unsigned int *>(h->pin + 7)[1];
        if (n_redo) {  // transcripts with letters outside ACGT: explicit reverse sums, CTA-cooperative kernel
            if ((rc = launch_fused(h->st, m, trs, fb, h->d_redo.p, n_redo, FUSED_CAP_LARGE, h->d_work.p + 3, h->d_wq_f.p, h->d_wq_r.p,
                                   h->d_kmax.p, fallback, h->sm_count))) return rc;
            h->launches[RLSIM_T_FRAGMENT]++;
        }
    } else {
        if ((rc = build_fused_items(h, t0, t1))) return rc;
        if ((rc = launch_fused(h->st, m, trs, fb, h->d_fitems_s.p, (uint32_t)h->h_fitems_s.size(), FUSED_CAP_SMALL, h->d_work.p + 4, h->d_wq_f.p,
                               h->d_wq_r.p, h->d_kmax.p, fallback, h->sm_count))) return rc;
        if ((rc = launch_fused(h->st, m, trs, fb, h->d_fitems_l.p, (uint32_t)h->h_fitems_l.size(), FUSED_CAP_LARGE, h->d_work.p + 5, h->d_wq_f.p,
                               h->d_wq_r.p, h->d_kmax.p, fallback, h->sm_count))) return rc;
        h->launches[RLSIM_T_FRAGMENT] += (h->h_fitems_s.size() ? 1 : 0) + (h->h_fitems_l.size() ? 1 : 0);
    }
    if (!warp_kernel || n_redo) {
        CK(h, cudaMemcpyAsync(h->pin, h->d_long_count.p, 4, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaMemcpyAsync(h->pin + 7, h->d_work.p + 6, 4, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
    }
    *n_long = *reinterpret_cast<unsigned int *>(h->pin);
    *fell_back = *reinterpret_cast<int *>(h->pin + 7);
    return 0;
}

// Fragment, dedup and amplify the transcripts [t0, t1) of a chunk whose ingest + prefix sums are already queued on
// the compute stream.  The host waits on the compute stream only; the copy stream keeps feeding later chunks.
// timed: record stage timers (only when no chunk copies are in flight: see the note on timing events in add_core)
int eager_chunk(rlsim_handle *h, uint32_t t0, uint32_t t1, bool timed = false) {
    int rc;
    struct OptTimer {  // a StageTimer that exists only when `timed`
        StageTimer *t = nullptr;
        OptTimer(bool on, rlsim_handle *h, int stage, cudaStream_t st = nullptr) { if (on) t = new StageTimer(h, stage, 0, st); }
        void stop() { if (t) { t->stop(); delete t; t = nullptr; } }
        ~OptTimer() { stop(); }
    };
    const bool trace = getenv("RLSIM_TRACE") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tt0 = now_ms();
    double tt1 = 0, tt2 = 0, tt3 = 0;
    const uint64_t g0 = h->h_mol_off[t0], g1 = h->h_mol_off[t1];
    if (!h->eager_started) {  // first chunk: histograms, counters
        h->eager_started = true;
        uint32_t hist_frag_n = (uint32_t)h->high + 1, hist_pa_n = h->polya_max + 1;
        CK(h, h->d_hist_frag.ensure(hist_frag_n)); CK(h, h->d_hist_polya.ensure(hist_pa_n));
        CK(h, cudaMemsetAsync(h->d_hist_frag.p, 0, (size_t)hist_frag_n * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_hist_polya.p, 0, (size_t)hist_pa_n * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_counters.p, 0, 4 * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_error.p, 0, 4, h->st));
    }
    if (g1 == g0) { h->eager_upto = t1; return 0; }
    double sel = 0, se = 0;
    for (uint32_t t = t0; t < t1; t++) { se += (double)h->h_expr[t]; sel += (double)h->h_expr[t] * (double)h->h_len[t]; }
    const uint64_t est = (uint64_t)((sel / frag_denominator(h) + se) * 1.25) + 8192;
    CK(h, h->d_keys.grow_preserve(h->eager_keys + est, h->eager_keys, h->st));
    CK(h, h->d_q_a.ensure(est)); CK(h, h->d_q_b.ensure(est));
    const uint32_t long_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((g1 - g0) / 16, 4096), 1u << 24);
    CK(h, h->d_long_list.ensure(long_cap)); CK(h, h->d_long_hi.ensure(long_cap));
    CK(h, cudaMemsetAsync(h->d_counters.p + 2, 0, 16, h->st));
    CK(h, cudaMemsetAsync(h->d_long_count.p, 0, 4, h->st));
    DevModel m = make_model(h);
    DevTranscripts tr = make_tr(h);
    tr.n = t1;  // the metadata of later transcripts reaches the device chunk by chunk: searches over off / mol_off stop here
    FragBuffers fb{h->d_keys.p, h->d_keys.n, h->d_counters.p, h->d_hist_frag.p, h->d_hist_polya.p, h->d_long_list.p, h->d_long_hi.p,
                   long_cap, h->d_long_count.p, h->d_error.p, h->d_q_a.p, h->d_q_b.p, h->d_q_a.n, h->d_counters.p + 2};
    const bool fused = fused_range_ok(h, t0, t1);
    unsigned int n_long = 0;
    unsigned long long n_q = 0, n_q_dirty = 0;
    if (fused) {
        int fell_back = 0;
        if ((rc = run_fused_range(h, m, tr, fb, t0, t1, &n_long, &fell_back)) || fell_back) { h->fused_disabled = true; h->eager_failed = true; return 0; }
    } else {
        if ((rc = ensure_warp_table(h, tr, g0, g1))) return rc;
        OptTimer tf(timed, h, RLSIM_T_FRAGMENT);
        if ((rc = launch_fragment(h->st, m, tr, g0, g1, fb, h->d_warp_t.p))) { h->eager_failed = true; return 0; }
        tf.stop();
        CK(h, cudaMemcpyAsync(h->pin, h->d_long_count.p, 4, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaMemcpyAsync(h->pin + 1, h->d_counters.p + 2, 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaMemcpyAsync(h->pin + 3, h->d_counters.p + 3, 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        n_long = *reinterpret_cast<unsigned int *>(h->pin);
        n_q = h->pin[1];
        n_q_dirty = h->pin[3];
    }
    tt1 = now_ms();
    // test hook: pretend the estimate of group number RLSIM_EAGER_TEST_FAIL (0-based) was too small
    const char *fail_at = getenv("RLSIM_EAGER_TEST_FAIL");
    const bool forced = fail_at && (uint32_t)atoi(fail_at) == h->eager_groups;
    h->eager_groups++;
    if (forced || n_long > long_cap || n_q > h->d_q_a.n) { h->eager_failed = true; return 0; }
    if (fused) {
        if (n_long) {
            DevTranscripts trs = tr; trs.sym = h->lut_sym && h->want_mirror ? 1 : 0;
            launch_fused_long(h->st, m, trs, fb, n_long, FUSED_CAP_LARGE, h->d_wq_f.p, h->d_wq_r.p, h->d_kmax.p, reinterpret_cast<int *>(h->d_work.p + 6));
            CK(h, cudaMemcpyAsync(h->pin + 7, h->d_work.p + 6, 4, cudaMemcpyDeviceToHost, h->st));
            h->launches[RLSIM_T_FRAGMENT] += 1;
        }
        h->fused_used = true;
    } else {
        OptTimer ti(timed, h, RLSIM_T_INTERVALS);
        launch_intervals(h->st, m, tr, fb, n_q, n_q_dirty);
        launch_fragment_long(h->st, m, tr, fb, n_long);
        ti.stop();
        h->launches[RLSIM_T_FRAGMENT] += 1 + (n_long ? 1 : 0); h->launches[RLSIM_T_INTERVALS] += n_q ? 1 : 0;
    }
    CK(h, cudaMemcpyAsync(h->pin + 2, h->d_counters.p, 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    if (fused && *reinterpret_cast<int *>(h->pin + 7)) { h->fused_disabled = true; h->eager_failed = true; return 0; }
    const unsigned long long cnt = h->pin[2];
    tt2 = now_ms();
    CK(h, cudaGetLastError());
    if (cnt > h->d_keys.n) { h->eager_failed = true; return 0; }
    const uint64_t nf = cnt - h->eager_keys;
    if (nf >= (1ull << 31)) { h->eager_failed = true; return 0; }
    uint64_t runs = 0;
    if (nf) {
        // species of this chunk: sort + run-length encode its keys (transcript.go:156-215), appended to the species table
        CK(h, h->d_keys_sorted.ensure(nf)); CK(h, h->d_uniq.ensure(nf));
        const uint64_t need = h->eager_sp + nf;
        if (h->pcr_in_flight && need > h->d_sp_count0.n) CK(h, cudaStreamWaitEvent(h->st, h->ev_pcr, 0));
        CK(h, h->d_sp_count0.grow_preserve(need, h->eager_sp, h->st));
        int end_bit = (int)std::min<uint32_t>(64, key_len_bits((uint32_t)h->high) + bits_for(h->total_padded));
        size_t tb = 0;
        OptTimer td(timed, h, RLSIM_T_DEDUP);
        CK(h, cub::DeviceRadixSort::SortKeys(nullptr, tb, h->d_keys.p + h->eager_keys, h->d_keys_sorted.p, (int)nf, 0, end_bit, h->st));
        if ((rc = cub_temp(h, tb))) return rc;
        CK(h, cub::DeviceRadixSort::SortKeys(h->d_cub.p, tb, h->d_keys.p + h->eager_keys, h->d_keys_sorted.p, (int)nf, 0, end_bit, h->st));
        tb = 0;
        CK(h, cub::DeviceRunLengthEncode::Encode(nullptr, tb, h->d_keys_sorted.p, h->d_uniq.p, h->d_sp_count0.p + h->eager_sp,
                                                 h->d_counters.p + 1, (int)nf, h->st));
        if ((rc = cub_temp(h, tb))) return rc;
        CK(h, cub::DeviceRunLengthEncode::Encode(h->d_cub.p, tb, h->d_keys_sorted.p, h->d_uniq.p, h->d_sp_count0.p + h->eager_sp,
                                                 h->d_counters.p + 1, (int)nf, h->st));
        CK(h, cudaMemcpyAsync(h->pin + 3, h->d_counters.p + 1, 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        tt3 = now_ms();
        runs = h->pin[3];
        // the species arrays move when they grow: the previous group's PCR must have finished with them first
        if (h->pcr_in_flight && (ns_need(h, h->eager_sp + runs)))
            CK(h, cudaStreamWaitEvent(h->st, h->ev_pcr, 0));
        h->launches[RLSIM_T_DEDUP] += 3 + radix_launches(end_bit);
        const uint64_t ns = h->eager_sp + runs;
        CK(h, h->d_sp_tr.grow_preserve(ns, h->eager_sp, h->st)); CK(h, h->d_sp_start.grow_preserve(ns, h->eager_sp, h->st));
        CK(h, h->d_sp_len.grow_preserve(ns, h->eager_sp, h->st)); CK(h, h->d_sp_count.grow_preserve(ns, h->eager_sp, h->st));
        CK(h, h->d_sp_eff.grow_preserve(ns, h->eager_sp, h->st)); CK(h, h->d_sp_gc.grow_preserve(ns, h->eager_sp, h->st));
        DevSpecies sp{};
        sp.n = runs;
        sp.tr = h->d_sp_tr.p + h->eager_sp; sp.start = h->d_sp_start.p + h->eager_sp; sp.len = h->d_sp_len.p + h->eager_sp;
        sp.count0 = h->d_sp_count0.p + h->eager_sp; sp.count = h->d_sp_count.p + h->eager_sp;
        sp.eff = h->keep_debug ? h->d_sp_eff.p + h->eager_sp : nullptr; sp.gc = h->keep_debug ? h->d_sp_gc.p + h->eager_sp : nullptr;
        launch_decode_species(h->st, tr, h->d_uniq.p, sp, (uint32_t)h->high);
        td.stop();
        if ((rc = ensure_len_eff(h))) return rc;
        m = make_model(h);  // picks up the length-efficiency table pointer
        CK(h, cudaEventRecord(h->ev_sp, h->st));
        CK(h, cudaStreamWaitEvent(h->st_pcr, h->ev_sp, 0));
        OptTimer tp(timed, h, RLSIM_T_PCR, h->st_pcr);
        launch_pcr(h->st_pcr, m, tr, sp, h->d_error.p);  // thermocycler.go:106-147 for this group's species
        tp.stop();
        CK(h, cudaEventRecord(h->ev_pcr, h->st_pcr));
        h->pcr_in_flight = true;
        h->launches[RLSIM_T_PCR] += 1;
        CK(h, cudaGetLastError());
    }
    if (trace) fprintf(stderr, "[rlsim trace]   layout+fragment %.3f, intervals %.3f, sort+rle %.3f, enqueue pcr %.3f ms (%llu frags, %llu species)\n", tt1 - tt0, tt2 - tt1, tt3 - tt2, now_ms() - tt3, (unsigned long long)nf, (unsigned long long)runs);
    h->eager_keys = cnt;
    h->eager_sp += runs;
    h->eager_upto = t1;
    return 0;
}

// ---- asynchronous form of the chunk pipeline ------------------------------------------------------------------------
// eager_chunk above reads three counts back per group (queued intervals, fragment keys, species) and the GPU idles while
// the host turns each of them into the next launch: measured, half of the time of a batch.  Here a chunk is split into a
// FRONT half (fragmentation: k_fragment, k_intervals on h->st, fed by device-side counts) and a BACK half (dedup + PCR:
// sort, run-length pass, species decode on h->st_back, k_pcr on h->st_pcr).  The only number the host needs is the
// chunk's fragment count (cub takes its item count by value); it is copied to pinned memory behind the front half and
// read one chunk later, after the next chunk's front half has been queued - so the GPU always has work while the host
// waits.  Where the species of a chunk land in the species table is decided on the device (k_species_begin).
// Fragment keys rotate through EAGER_BUFS buffers, so a slow sort (small, latency-bound) does not hold back the front half.  Any bound that proves too small raises a flag and the monolithic path
// recomputes the library, as with eager_chunk.
uint64_t eager_chunk_bound(rlsim_handle *h, double sum_expr_len, double sum_expr) {
    return (uint64_t)((sum_expr_len / frag_denominator(h) + sum_expr) * 1.25) + 2048;
}

// Sizes the buffers of the front half for `n_chunks` more chunks (largest fragment bound: `largest`, most molecules in
// one chunk: `max_mol`): nothing of it is reallocated while chunks are in flight.
int eager_async_reserve(rlsim_handle *h, uint32_t n_chunks, uint64_t largest, uint64_t max_mol, uint64_t total_bound) {
    // The short stages that gate the others run at high stream priority: their blocks are placed before the pending blocks
    // of the long k_fragment / k_pcr grids (measured: the radix sort of a chunk - decoupled look-back between its blocks -
    // took 0.5-0.9 ms beside those grids at equal priority, more than the whole chunk takes otherwise).
    int prio_lo = 0, prio_hi = 0;
    CK(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (getenv("RLSIM_NO_PRIORITY")) prio_hi = prio_lo;
    if (!h->st_back) CK(h, cudaStreamCreateWithPriority(&h->st_back, cudaStreamNonBlocking, prio_hi));
    if (!h->st_in) CK(h, cudaStreamCreateWithPriority(&h->st_in, cudaStreamNonBlocking, prio_hi));
    while (h->ev_in.size() < n_chunks) { cudaEvent_t e; CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_in.push_back(e); }
    const uint32_t slots = h->ecs_used + n_chunks;
    if (slots > h->d_ecs.n) {  // the records move: everything that reads them must be done
        CK(h, cudaStreamSynchronize(h->st_back));
        CK(h, cudaStreamSynchronize(h->st_pcr));
        h->pcr_in_flight = false;
    }
    if (!h->eager_started) {  // first chunk of this transcript set: histograms, counters
        h->eager_started = true;
        h->eager_async = true;
        uint32_t hist_frag_n = (uint32_t)h->high + 1, hist_pa_n = h->polya_max + 1;
        CK(h, h->d_hist_frag.ensure(hist_frag_n)); CK(h, h->d_hist_polya.ensure(hist_pa_n));
        CK(h, h->d_eglobal.ensure(1));
        CK(h, cudaMemsetAsync(h->d_hist_frag.p, 0, (size_t)hist_frag_n * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_hist_polya.p, 0, (size_t)hist_pa_n * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_counters.p, 0, 4 * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_eglobal.p, 0, sizeof(rl::EagerGlobal), h->st));
        CK(h, cudaMemsetAsync(h->d_error.p, 0, 4, h->st));
    }
    CK(h, h->d_ecs.grow_preserve(slots, h->ecs_used, h->st));
    CK(h, cudaMemsetAsync(h->d_ecs.p + h->ecs_used, 0, (size_t)n_chunks * sizeof(rl::EagerChunk), h->st));
    h->ecs_pin.resize(slots);
    while (h->ev_cnt.size() < slots) {
        cudaEvent_t e1, e2;
        CK(h, cudaEventCreateWithFlags(&e1, cudaEventDisableTiming)); h->ev_cnt.push_back(e1);
        CK(h, cudaEventCreateWithFlags(&e2, cudaEventDisableTiming)); h->ev_sorted.push_back(e2);
    }
    if (largest > h->d_ekeys[0].n) CK(h, cudaStreamSynchronize(h->st_back));  // a sort may still read them
    for (uint32_t i = 0; i < EAGER_BUFS; i++) CK(h, h->d_ekeys[i].ensure(largest));
    CK(h, h->d_q_a.ensure(largest)); CK(h, h->d_q_b.ensure(largest));
    const uint32_t long_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(max_mol / 16, 4096), 1u << 24);
    for (uint32_t i = 0; i < EAGER_BUFS; i++) { CK(h, h->d_elong_list[i].ensure(long_cap)); CK(h, h->d_elong_hi[i].ensure(long_cap)); }
    CK(h, h->d_warp_t.ensure(max_mol / 32 + 2));
    {   // Back half: the species table and the dedup scratch grow on demand (eager_back), which stalls the pipeline each time.
        // A first library is spared most of that by starting from a quarter of the bound (the bound itself is ~10x the
        // fragments a typical model keeps); later libraries find the buffers at the size the last one needed.
        const uint64_t need = h->eager_sp_ub + total_bound / 4, scratch = largest / 4 + 1;
        if (ns_need(h, need) || need > h->d_sp_count0.n || scratch > h->d_keys_sorted.n || scratch > h->d_uniq.n || scratch > h->d_rle_cnt.n) {
            CK(h, cudaStreamSynchronize(h->st_back));
            CK(h, cudaStreamSynchronize(h->st_pcr));
            h->pcr_in_flight = false;
            const uint64_t used = h->eager_sp_ub;
            CK(h, h->d_sp_count0.grow_preserve(need, used, h->st_back));
            CK(h, h->d_sp_tr.grow_preserve(need, used, h->st_back)); CK(h, h->d_sp_start.grow_preserve(need, used, h->st_back));
            CK(h, h->d_sp_len.grow_preserve(need, used, h->st_back)); CK(h, h->d_sp_count.grow_preserve(need, used, h->st_back));
            CK(h, h->d_sp_eff.grow_preserve(need, used, h->st_back)); CK(h, h->d_sp_gc.grow_preserve(need, used, h->st_back));
            CK(h, h->d_keys_sorted.ensure(scratch)); CK(h, h->d_uniq.ensure(scratch)); CK(h, h->d_rle_cnt.ensure(scratch));
        }
    }
    return ensure_len_eff(h);
}

// Front half of the transcripts [t0, t1) (ingest + prefix sums already queued on h->st).  Returns the chunk's slot, or -1.
int eager_front(rlsim_handle *h, uint32_t t0, uint32_t t1, uint64_t bound, int *slot_out) {
    int rc;
    *slot_out = -1;
    const uint64_t g0 = h->h_mol_off[t0], g1 = h->h_mol_off[t1];
    const char *fail_at = getenv("RLSIM_EAGER_TEST_FAIL");  // test hook: pretend the bound of chunk number N was too small
    const bool forced = fail_at && (uint32_t)atoi(fail_at) == h->eager_groups;
    if (g1 == g0) { h->eager_upto = t1; return 0; }
    h->eager_groups++;
    if (forced || bound >= (1ull << 31) || bound > h->d_ekeys[0].n || h->ecs_used >= h->d_ecs.n) { h->eager_failed = true; return 0; }
    const uint32_t slot = h->ecs_used++;
    rl::EagerChunk *cs = h->d_ecs.p + slot;
    uint64_t *keys = h->d_ekeys[slot % EAGER_BUFS].p;
    if (slot >= EAGER_BUFS) CK(h, cudaStreamWaitEvent(h->st, h->ev_sorted[slot - EAGER_BUFS], 0));  // that chunk's sort has read this buffer
    const uint32_t long_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((g1 - g0) / 16, 4096), h->d_elong_list[slot % EAGER_BUFS].n);
    DevModel m = make_model(h);
    DevTranscripts tr = make_tr(h);
    tr.n = t1;  // the metadata of later transcripts reaches the device chunk by chunk: searches over off / mol_off stop here
    const uint64_t q_cap = std::min<uint64_t>(bound, std::min(h->d_q_a.n, h->d_q_b.n));
    // molecules with more breakpoints than a thread sorts locally are listed (per key buffer: the back half serves them)
    FragBuffers fb{keys, bound, &cs->key_count, h->d_hist_frag.p, h->d_hist_polya.p, h->d_elong_list[slot % EAGER_BUFS].p, h->d_elong_hi[slot % EAGER_BUFS].p,
                   long_cap, &cs->long_count, h->d_error.p, h->d_q_a.p, h->d_q_b.p, q_cap, &cs->q_count};
    const bool tm = getenv("RLSIM_TRACE") && atoi(getenv("RLSIM_TRACE")) >= 2;
    if (tm) h->mark("front start", (int)slot, h->st);
    if ((rc = ensure_warp_table(h, tr, g0, g1))) return rc;
    if ((rc = launch_fragment(h->st, m, tr, g0, g1, fb, h->d_warp_t.p))) { h->eager_failed = true; return 0; }
    if (tm) h->mark("fragment end", (int)slot, h->st);
    launch_intervals(h->st, m, tr, fb, q_cap, 1, &cs->q_count);
    if (tm) h->mark("intervals end", (int)slot, h->st);
    h->launches[RLSIM_T_FRAGMENT] += 1; h->launches[RLSIM_T_INTERVALS] += 1;
    h->ecs_bound.resize(std::max<size_t>(h->ecs_bound.size(), slot + 1));  // host-side notes for the back half
    h->ecs_bound[slot] = {bound, q_cap, long_cap, t1};
    CK(h, cudaMemcpyAsync(&h->ecs_pin[slot], cs, sizeof(rl::EagerChunk), cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaEventRecord(h->ev_cnt[slot], h->st));
    CK(h, cudaGetLastError());
    h->eager_upto = t1;
    *slot_out = (int)slot;
    return 0;
}

// Back half of a chunk: sort + run-length encode its keys (transcript.go:156-215), place and decode its species, amplify.
int eager_back(rlsim_handle *h, int slot) {
    int rc;
    if (slot < 0 || h->eager_failed) return 0;
    CK(h, cudaEventSynchronize(h->ev_cnt[slot]));  // the one host wait per chunk; the next chunk's front half is queued behind it
    rl::EagerChunk c = h->ecs_pin[slot];
    const auto &b = h->ecs_bound[slot];
    if (c.key_count > b.bound || c.q_count > b.q_cap || c.long_count > b.long_cap || c.key_count >= (1ull << 31)) { h->eager_failed = true; return 0; }
    rl::EagerChunk *cs = h->d_ecs.p + slot;
    CK(h, cudaStreamWaitEvent(h->st_back, h->ev_cnt[slot], 0));
    if (c.long_count) {  // rare: deferred molecules add keys, so the count is read again
        DevModel ml = make_model(h);
        DevTranscripts trl = make_tr(h);
        trl.n = b.t1;
        FragBuffers fb{h->d_ekeys[slot % EAGER_BUFS].p, b.bound, &cs->key_count, h->d_hist_frag.p, h->d_hist_polya.p, h->d_elong_list[slot % EAGER_BUFS].p,
                       h->d_elong_hi[slot % EAGER_BUFS].p, b.long_cap, &cs->long_count, h->d_error.p, h->d_q_a.p, h->d_q_b.p, b.q_cap, &cs->q_count};
        launch_fragment_long(h->st_back, ml, trl, fb, c.long_count);
        CK(h, cudaMemcpyAsync(&h->ecs_pin[slot], cs, sizeof(rl::EagerChunk), cudaMemcpyDeviceToHost, h->st_back));
        CK(h, cudaStreamSynchronize(h->st_back));
        c = h->ecs_pin[slot];
        h->launches[RLSIM_T_FRAGMENT] += 1;
        if (c.key_count > b.bound || c.key_count >= (1ull << 31)) { h->eager_failed = true; return 0; }
    }
    const uint64_t nf = c.key_count;
    if (nf == 0) { CK(h, cudaEventRecord(h->ev_sorted[slot], h->st_back)); return 0; }
    const uint64_t need = h->eager_sp_ub + nf;
    if (ns_need(h, need) || need > h->d_sp_count0.n || nf > h->d_keys_sorted.n || nf > h->d_uniq.n || nf > h->d_rle_cnt.n) {
        CK(h, cudaStreamSynchronize(h->st_back));   // the arrays move: whatever reads them must be done
        CK(h, cudaStreamSynchronize(h->st_pcr));
        const uint64_t used = h->eager_sp_ub;
        CK(h, h->d_sp_count0.grow_preserve(need, used, h->st_back));
        CK(h, h->d_sp_tr.grow_preserve(need, used, h->st_back)); CK(h, h->d_sp_start.grow_preserve(need, used, h->st_back));
        CK(h, h->d_sp_len.grow_preserve(need, used, h->st_back)); CK(h, h->d_sp_count.grow_preserve(need, used, h->st_back));
        CK(h, h->d_sp_eff.grow_preserve(need, used, h->st_back)); CK(h, h->d_sp_gc.grow_preserve(need, used, h->st_back));
        CK(h, h->d_keys_sorted.ensure(nf)); CK(h, h->d_uniq.ensure(nf)); CK(h, h->d_rle_cnt.ensure(nf));
    }
    h->eager_sp_ub = need;
    DevModel m = make_model(h);
    DevTranscripts tr = make_tr(h);
    tr.n = b.t1;
    const uint64_t *keys = h->d_ekeys[slot % EAGER_BUFS].p;
    const int end_bit = (int)std::min<uint32_t>(64, key_len_bits((uint32_t)h->high) + bits_for(h->h_off[b.t1]));
    size_t tb = 0, tb2 = 0;
    CK(h, cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, h->d_keys_sorted.p, (int)nf, 0, end_bit, h->st_back));
    CK(h, cub::DeviceRunLengthEncode::Encode(nullptr, tb2, h->d_keys_sorted.p, h->d_uniq.p, h->d_rle_cnt.p, &cs->runs, (int)nf, h->st_back));
    if (std::max(tb, tb2) + 16 > h->d_cub.n) { CK(h, cudaStreamSynchronize(h->st_back)); if ((rc = cub_temp(h, std::max(tb, tb2) * 2))) return rc; }
    tb = h->d_cub.n;
    const bool tm = getenv("RLSIM_TRACE") && atoi(getenv("RLSIM_TRACE")) >= 2;
    if (tm) h->mark("back start", slot, h->st_back);
    CK(h, cub::DeviceRadixSort::SortKeys(h->d_cub.p, tb, keys, h->d_keys_sorted.p, (int)nf, 0, end_bit, h->st_back));
    CK(h, cudaEventRecord(h->ev_sorted[slot], h->st_back));
    if (tm) h->mark("sort end", slot, h->st_back);
    tb = h->d_cub.n;
    CK(h, cub::DeviceRunLengthEncode::Encode(h->d_cub.p, tb, h->d_keys_sorted.p, h->d_uniq.p, h->d_rle_cnt.p, &cs->runs, (int)nf, h->st_back));
    launch_species_begin(h->st_back, cs, h->d_eglobal.p, b.bound, b.q_cap, b.long_cap, h->d_sp_tr.n);
    DevSpecies sp{};
    sp.n = nf;
    sp.tr = h->d_sp_tr.p; sp.start = h->d_sp_start.p; sp.len = h->d_sp_len.p; sp.count0 = h->d_sp_count0.p; sp.count = h->d_sp_count.p;
    sp.eff = h->keep_debug ? h->d_sp_eff.p : nullptr; sp.gc = h->keep_debug ? h->d_sp_gc.p : nullptr;
    launch_decode_species_dyn(h->st_back, tr, h->d_uniq.p, h->d_rle_cnt.p, sp, (uint32_t)h->high, cs, nf);
    if (tm) h->mark("back end", slot, h->st_back);
    h->launches[RLSIM_T_DEDUP] += 4 + radix_launches(end_bit);
    CK(h, cudaEventRecord(h->ev_sp, h->st_back));
    CK(h, cudaStreamWaitEvent(h->st_pcr, h->ev_sp, 0));
    if (tm) h->mark("pcr start", slot, h->st_pcr);
    launch_pcr_dyn(h->st_pcr, m, tr, sp, h->d_error.p, cs, nf);  // thermocycler.go:106-147 for this chunk's species
    if (tm) h->mark("pcr end", slot, h->st_pcr);
    CK(h, cudaEventRecord(h->ev_pcr, h->st_pcr));
    h->pcr_in_flight = true;
    h->launches[RLSIM_T_PCR] += 1;
    CK(h, cudaGetLastError());
    return 0;
}

// End of a batch: totals and overflow flags (PCR of the last chunks may still be running; h->st waits for the back half).
int eager_async_finish(rlsim_handle *h) {
    if (!h->eager_async || !h->eager_started || !h->st_back) return 0;
    CK(h, cudaMemcpyAsync(h->pin, h->d_eglobal.p, sizeof(rl::EagerGlobal), cudaMemcpyDeviceToHost, h->st_back));
    CK(h, cudaStreamSynchronize(h->st_back));
    const rl::EagerGlobal g = *reinterpret_cast<const rl::EagerGlobal *>(h->pin);
    if (g.fail) { h->eager_failed = true; return 0; }
    if (h->eager_failed) return 0;
    h->eager_keys = g.frag_total;
    h->eager_sp = g.sp_cursor;
    return 0;
}

// everything that needs the complete transcript set: molecule offsets, remaining prefix sums
int finalize(rlsim_handle *h, bool with_prefix = true) {
    uint32_t n = (uint32_t)h->h_len.size();
    if (!h->mol_off_valid) {
        if (h->h_mol_off.empty()) h->h_mol_off.push_back(0);
        h->n_mol = h->h_mol_off[n];
        CK(h, h->d_mol_off.ensure(n + 1));
        CK(h, cudaMemcpyAsync(h->d_mol_off.p, h->h_mol_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        h->mol_off_valid = true;
    }
    if (with_prefix && needs_fwd(h->p.frag_method) && h->prefix_upto < n && !fused_range_ok(h, 0, n)) {
        int rcg = grow_prefix(h, h->total_padded, n, h->prefix_upto ? h->total_padded : 0, h->prefix_upto);
        if (rcg) return rcg;
        StageTimer tm(h, RLSIM_T_PREFIX, 0);
        int rc = launch_prefix_range(h, h->prefix_upto, n);
        if (rc) return rc;
        tm.stop();
        CK(h, cudaGetLastError());
        h->prefix_upto = n;
    }
    return 0;
}

int upload_tables(rlsim_handle *h) {
    CK(h, h->d_gc_raw.ensure(101));
    CK(h, cudaMemcpyAsync(h->d_gc_raw.p, h->p.gc_raw, 101 * 8, cudaMemcpyHostToDevice, h->st));
    if (!h->raw_len.empty()) {
        CK(h, h->d_raw_len.ensure(h->raw_len.size()));
        CK(h, h->d_raw_prob.ensure(h->raw_prob.size()));
        CK(h, cudaMemcpyAsync(h->d_raw_len.p, h->raw_len.data(), h->raw_len.size() * 4, cudaMemcpyHostToDevice, h->st));
        CK(h, cudaMemcpyAsync(h->d_raw_prob.p, h->raw_prob.data(), h->raw_prob.size() * 8, cudaMemcpyHostToDevice, h->st));
    }
    return 0;
}

int cub_temp(rlsim_handle *h, size_t bytes) { CK(h, h->d_cub.ensure(bytes + 16)); return 0; }

}  // namespace rl_api
using namespace rl_api;

extern "C" {

int rlsim_create(int device, rlsim_handle **out) {
    *out = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return RLSIM_E_CUDA; }
    rlsim_handle *h = new rlsim_handle();
    h->device = device;
    if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->st_pcr, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->st_aux, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_aux, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_sp, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_pcr, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_ingest, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&h->pin, 64, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess ||
        (e = h->d_error.ensure(1)) != cudaSuccess || (e = h->d_counters.ensure(4)) != cudaSuccess ||
        (e = h->d_long_count.ensure(1)) != cudaSuccess || (e = h->d_work.ensure(16)) != cudaSuccess || (e = h->d_kmax.ensure(1)) != cudaSuccess) {
        g_create_error = std::string("rlsim_create: ") + cudaGetErrorString(e);
        delete h;
        return RLSIM_E_CUDA;
    }
    cudaMemsetAsync(h->d_error.p, 0, sizeof(int), h->st);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = h;
    return 0;
}

void rlsim_destroy(rlsim_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    drain_timers(h);
    comm_destroy(h);
    for (auto e : h->ev_chunk) cudaEventDestroy(e);
    for (auto e : h->ev_free) cudaEventDestroy(e);
    if (h->st_copy) cudaStreamDestroy(h->st_copy);
    if (h->st_pcr) cudaStreamDestroy(h->st_pcr);
    if (h->st_back) cudaStreamDestroy(h->st_back);
    if (h->st_in) cudaStreamDestroy(h->st_in);
    for (auto e : h->ev_in) cudaEventDestroy(e);
    for (auto e : h->ev_cnt) cudaEventDestroy(e);
    for (auto e : h->ev_sorted) cudaEventDestroy(e);
    if (h->st_aux) cudaStreamDestroy(h->st_aux);
    if (h->ev_aux) cudaEventDestroy(h->ev_aux);
    if (h->ev_sp) cudaEventDestroy(h->ev_sp);
    if (h->ev_pcr) cudaEventDestroy(h->ev_pcr);
    if (h->ev_ingest) cudaEventDestroy(h->ev_ingest);
    if (h->pin) cudaFreeHost(h->pin);
    delete h->team;
    if (h->stage_ring) cudaFreeHost(h->stage_ring);
    for (auto e : h->ev_slot) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

const char *rlsim_last_error(const rlsim_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rlsim_set_model(rlsim_handle *h, const rlsim_params *p) {
    if (p->kmer_len < 1 || p->kmer_len > 12) FAIL(h, RLSIM_E_UNSUPPORTED, "primer length outside 1..12 is not supported by the GPU path");
    if (p->n_polya < 1 || p->n_polya > RLSIM_MAX_COMPS || p->n_target < 0 || p->n_target > RLSIM_MAX_COMPS)
        FAIL(h, RLSIM_E_ARG, "invalid number of mixture components");
    if (p->frag_method < 0 || p->frag_method > RLSIM_PRIM_JUMP) FAIL(h, RLSIM_E_ARG, "Invalid fragmentation method.");
    if (p->frag_method == RLSIM_PRE_PRIM && p->frag_param <= 0) FAIL(h, RLSIM_E_ARG, "pre_prim needs a positive parameter");
    if (p->nr_cycles < 0) FAIL(h, RLSIM_E_ARG, "negative number of PCR cycles");
    cudaSetDevice(h->device);
    h->p = *p;
    h->have_model = true;
    uint64_t mn, mx;
    mix_minmax(p->polya, p->n_polya, &mn, &mx);  // pool.go:84
    if (mx >= (1u << 24)) FAIL(h, RLSIM_E_UNSUPPORTED, "poly(A) maximum >= 2^24");
    if (!h->h_len.empty() && h->polya_max != (uint32_t)mx) FAIL(h, RLSIM_E_ARG, "poly(A) maximum cannot change after transcripts were added");
    h->polya_max = (uint32_t)mx;
    h->prefix_upto = 0;
    h->fused_disabled = false;
    h->lut_valid = h->lut_valid && h->lut_k == p->kmer_len && h->lut_T == p->priming_temp;
    h->want_mirror = getenv("RLSIM_NO_MIRROR") == nullptr;  // diagnostic switch: always build explicit reverse arrays
    h->len_eff_ready = false;
    eager_reset(h);
    if (p->n_target > 0) {
        mix_minmax(p->target, p->n_target, &h->low, &h->high);  // target.go:67
        if (h->high >= (1u << 24)) FAIL(h, RLSIM_E_UNSUPPORTED, "maximum fragment length >= 2^24");
    }
    h->fragmented = h->amplified = h->sampled = false;
    return upload_tables(h);
}

int rlsim_set_len_eff(rlsim_handle *h, const double *len_eff, uint64_t low, uint64_t high) {
    if (high < low) FAIL(h, RLSIM_E_ARG, "len_eff: high < low");
    h->h_len_eff.assign(len_eff, len_eff + (high - low + 1));
    h->len_eff_low = low; h->len_eff_high = high;
    h->len_eff_ready = false;
    h->eager_pcr = false;
    if (h->eager_upto) { h->eager_upto = 0; h->eager_failed = true; }  // species already amplified with another table
    return 0;
}

int rlsim_set_raw_target(rlsim_handle *h, const uint32_t *lengths, const double *probs, uint32_t n) {
    if (n == 0) FAIL(h, RLSIM_E_ARG, "empty raw target");
    cudaSetDevice(h->device);
    h->raw_len.assign(lengths, lengths + n);
    h->raw_prob.assign(probs, probs + n);
    for (uint32_t i = 0; i < n; i++)
        if (lengths[i] >= (1u << 24)) FAIL(h, RLSIM_E_UNSUPPORTED, "maximum fragment length >= 2^24");
    return upload_tables(h);
}

}  // extern "C"
// bring the sampled target histogram to the host (no-op when it is already there)
int rl_api::target_host(rlsim_handle *h) {
    if (!h->target_pending) return 0;
    h->target_pending = false;
    h->h_target.assign(h->target_hist_n, 0);
    CK(h, cudaMemcpyAsync(h->h_target.data(), h->d_hist_target.p, (size_t)h->target_hist_n * 8, cudaMemcpyDeviceToHost, h->st_aux));
    CK(h, cudaStreamSynchronize(h->st_aux));
    return 0;
}

extern "C" {
static int sample_target_range(rlsim_handle *h, uint64_t first, uint64_t count, bool whole) {
    if (!h->have_model) FAIL(h, RLSIM_E_ARG, "rlsim_set_model first");
    bool raw = h->p.n_target == 0;
    if (raw && h->raw_len.empty()) FAIL(h, RLSIM_E_ARG, "raw target requested but rlsim_set_raw_target was not called");
    cudaSetDevice(h->device);
    uint32_t hist_n;
    if (raw) hist_n = *std::max_element(h->raw_len.begin(), h->raw_len.end()) + 1;
    else { uint64_t mn, mx; mix_minmax(h->p.target, h->p.n_target, &mn, &mx); hist_n = (uint32_t)mx + 1; }
    CK(h, h->d_hist_target.ensure(hist_n));
    // The draws depend on nothing but the model: they run on their own stream beside the transcript pipeline, and the
    // histogram comes to the host when it is first needed (target_host).
    CK(h, cudaEventRecord(h->ev_aux, h->st));                 // model tables uploaded, earlier readers of the histogram done
    CK(h, cudaStreamWaitEvent(h->st_aux, h->ev_aux, 0));
    CK(h, cudaMemsetAsync(h->d_hist_target.p, 0, (size_t)hist_n * 8, h->st_aux));
    DevModel m = make_model(h);
    StageTimer tm(h, RLSIM_T_TARGET, 1, h->st_aux);
    launch_target(h->st_aux, m, first, count, h->d_hist_target.p, hist_n);
    tm.stop();
    CK(h, cudaGetLastError());
    h->target_pending = true;
    h->target_hist_n = hist_n;
    h->req_frags = count;
    h->have_target = whole;  // a partial histogram must be completed with rlsim_set_target_hist
    if (raw && whole) {  // the raw target's range and mean shape the library: needed now
        int rc = target_host(h);
        if (rc) return rc;
        raw_target_finish(h); h->len_eff_ready = false; if (h->eager_upto) { h->eager_upto = 0; h->eager_failed = true; }
    }
    return 0;
}

int rlsim_sample_target(rlsim_handle *h, int64_t req_frags) {
    if (req_frags < 0) FAIL(h, RLSIM_E_ARG, "Illegal fragment number!");
    return sample_target_range(h, 0, (uint64_t)req_frags, true);
}

int rlsim_sample_target_range(rlsim_handle *h, uint64_t first, uint64_t count) {
    return sample_target_range(h, first, count, false);
}

int rlsim_set_target_hist(rlsim_handle *h, const uint32_t *lengths, const uint64_t *counts, uint32_t n) {
    if (!h->have_model) FAIL(h, RLSIM_E_ARG, "rlsim_set_model first");
    uint32_t mx = 0;
    for (uint32_t i = 0; i < n; i++) mx = std::max(mx, lengths[i]);
    if (mx >= (1u << 24)) FAIL(h, RLSIM_E_UNSUPPORTED, "maximum fragment length >= 2^24");
    if (h->target_pending) { cudaSetDevice(h->device); CK(h, cudaStreamSynchronize(h->st_aux)); h->target_pending = false; }
    std::vector<uint64_t> hist((size_t)mx + 1, 0);
    for (uint32_t i = 0; i < n; i++) hist[lengths[i]] += counts[i];
    set_target_from_hist(h, hist.data(), mx + 1);
    return 0;
}
}  // extern "C"
// impose a complete target histogram indexed by length (the host's own Targeter, or the sum of the ranks' partial draws)
void rl_api::set_target_from_hist(rlsim_handle *h, const uint64_t *hist, uint32_t n) {
    h->target_pending = false;
    h->h_target.assign(hist, hist + n);
    h->req_frags = 0;
    for (uint32_t i = 0; i < n; i++) h->req_frags += hist[i];
    h->have_target = true;
    if (h->p.n_target == 0) { raw_target_finish(h); h->len_eff_ready = false; if (h->eager_upto) { h->eager_upto = 0; h->eager_failed = true; } }
}
extern "C" {

int rlsim_set_first_index(rlsim_handle *h, uint32_t first_index) {
    if (first_index != h->first_index && h->eager_upto) { h->eager_upto = 0; h->eager_failed = true; }  // streams are keyed by it
    h->first_index = first_index;
    return 0;
}

int rlsim_clear_transcripts(rlsim_handle *h) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    h->h_srclen.clear(); h->h_expr.clear(); h->h_off.clear(); h->h_len.clear();
    h->prefix_upto = 0; h->fragmented = h->amplified = h->sampled = false;
    h->fused_disabled = false; h->fitems_key = -1; h->max_len = 0; h->warp_t_n = 0;
    h->n_mol = h->total_padded = h->n_frag = h->n_sp = h->n_out = h->pool_total = 0;
    h->sum_expr = h->sum_expr_len = 0; h->mol_off_valid = false;
    h->h_mol_off.clear();
    eager_reset(h);
    drain_timers(h);
    memset(h->ms, 0, sizeof h->ms);
    memset(h->launches, 0, sizeof h->launches);
    return 0;
}

}  // extern "C"

// bases: host memory (copied in pipelined chunks), or -- dev_bases -- sequences already in device memory
// Per-transcript metadata of a chunk, written by the host into pinned arrays (device-accessible under unified addressing)
// while the chunk is on the wire, and pulled into device memory by this kernel on the compute stream.  A cudaMemcpyAsync
// would queue behind the chunk copies already handed to the copy engine (measured twice: the first ingest then waits for the
// whole transfer); a kernel reading host memory does not.
__global__ void k_copy_meta(const uint64_t *__restrict__ h_off, const uint32_t *__restrict__ h_len, const uint64_t *__restrict__ h_expr,
                            const uint64_t *__restrict__ h_mol_off, const uint64_t *__restrict__ h_rel, uint64_t *__restrict__ d_off,
                            uint32_t *__restrict__ d_len, uint64_t *__restrict__ d_expr, uint64_t *__restrict__ d_mol_off,
                            uint64_t *__restrict__ d_rel, uint32_t n) {  // all pointers at the chunk's first transcript
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    d_off[i] = h_off[i];           // n + 1 boundary values
    d_rel[i] = h_rel[i];
    if (d_mol_off) d_mol_off[i] = h_mol_off[i];
    if (i < n) { d_len[i] = h_len[i]; d_expr[i] = h_expr[i]; }
}

// bases: host memory (copied in pipelined chunks), or -- dev_bases -- sequences already in device memory.
// packed: `bases` holds 2-bit codes (4 bases per byte) and `pk_mask` (may be null) 1 bit per base, both indexed by the
// same base offsets as an ASCII batch (rlsim_add_transcripts_packed); 0.375 B per base cross the host link.
static int add_core(rlsim_handle *h, const uint8_t *bases, const uint8_t *dev_bases, const uint64_t *offsets, const uint64_t *expr, uint32_t n,
                    bool packed = false, const uint8_t *pk_mask = nullptr) {
    if (!h->have_model) FAIL(h, RLSIM_E_ARG, "rlsim_set_model must precede rlsim_add_transcripts (the poly(A) maximum shapes the layout)");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    const bool trace = getenv("RLSIM_TRACE") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_entry = now_ms();
    const uint32_t n0 = (uint32_t)h->h_len.size(), nn = n0 + n;
    const uint64_t o0 = offsets[0], nbytes = offsets[n] - o0, pa = h->polya_max;
    // ---- sizes.  The padded layout of the batch is only known after a pass over its transcripts; that pass runs chunk by
    // chunk while the chunks are on the wire, so the device arrays are sized from an upper bound (every transcript grows by
    // the poly(A) maximum and at most 31 bases of padding).
    const uint64_t old_total = h->total_padded;
    const uint64_t tot_ub = old_total + nbytes + (uint64_t)n * (pa + 31);
    if (tot_ub >= (1ull << 40)) FAIL(h, RLSIM_E_UNSUPPORTED, "more than 2^40 bases on one GPU");
    if (h->h_off.empty()) h->h_off.push_back(0);
    if (h->h_mol_off.empty()) h->h_mol_off.push_back(0);
    PinVec<uint64_t> &rel = h->h_rel;
    rel.resize_uninit(n + 1);
    h->h_srclen.resize(nn); h->h_expr.resize_uninit(nn); h->h_len.resize_uninit(nn); h->h_off.resize_uninit(nn + 1);
    h->h_mol_off.resize_uninit(nn + 1);
    h->mol_off_valid = false;
    h->eager_pcr = false;
    h->fragmented = h->amplified = h->sampled = false;
    // ---- device arrays: grow keeping what earlier batches wrote
    CK(h, h->d_bases.grow_preserve(tot_ub + 64, old_total, h->st));
    CK(h, h->d_packed.grow_preserve(tot_ub / 16 + 8, old_total / 16, h->st));
    CK(h, h->d_nmask.grow_preserve(tot_ub / 32 + 8, old_total / 32, h->st));
    CK(h, h->d_gc.grow_preserve(tot_ub / 32 + nn + 8, old_total / 32 + n0 + 1, h->st));
    // priming prefix sums: in shared memory inside the fused kernel (opt-in) when every transcript of the batch fits, else in HBM
    const bool fused_batch = fused_method(h);  // decided per batch below, once the lengths are known
    const bool want_prefix_any = needs_fwd(h->p.frag_method) && h->prefix_upto == n0;
    if (want_prefix_any) {
        int rcg = grow_prefix(h, tot_ub, nn, old_total, n0);
        if (rcg) return rcg;
    }
    CK(h, h->d_off.grow_preserve(nn + 1, n0 + 1, h->st)); CK(h, h->d_len.grow_preserve(nn, n0, h->st)); CK(h, h->d_expr.grow_preserve(nn, n0, h->st));
    CK(h, h->d_mol_off.grow_preserve(nn + 1, n0 + 1, h->st));
    CK(h, h->d_dirty.grow_preserve(nn + 1, n0, h->st));
    CK(h, h->staging.off.ensure(n + 1));
    // packed source: codes from byte (o0 >> 2) of the caller's array at staging byte 0, the mask from byte (o0 >> 3) at pk_m0
    const uint64_t pk_m0 = ((nbytes >> 2) + 64 + 15) & ~15ull;
    if (!dev_bases) CK(h, h->staging.bases.ensure(std::max(nbytes + 64, pk_m0 + (nbytes >> 3) + 80)));  // the ingest kernels read whole aligned words
    // ---- the chunks
    // (measured) a timing event recorded on the compute stream while the chunk copies are queued completes only after
    // all of them, and holds back the first ingest: the stage timer starts before the copies are queued
    StageTimer tm(h, RLSIM_T_LAYOUT, 0);
    CK(h, cudaEventRecord(h->ev_ingest, h->st));  // the previous batch's ingest kernels read the staging buffer
    CK(h, cudaStreamWaitEvent(h->st_copy, h->ev_ingest, 0));
    const uint8_t *src = dev_bases ? dev_bases + o0 : h->staging.bases.p;
    // ---- pipeline: H2D of chunk c+1 on the copy stream overlaps ingest (+ prefix sums) of chunk c
    const uint64_t chunk_bytes = getenv("RLSIM_CHUNK_BYTES") ? std::max(1ll, atoll(getenv("RLSIM_CHUNK_BYTES")))  // test knob
                               : (uint64_t)(getenv("RLSIM_CHUNK_MB") ? std::max(1, atoi(getenv("RLSIM_CHUNK_MB"))) : (packed ? 128 : 32)) << 20;
    // (bases per chunk.  A packed batch crosses the link in 3/8 of the time, so its pipeline is bound by the kernels, which
    // run more efficiently over larger chunks: measured 7.6 / 6.9 / 6.6 ms per C3 batch at 32 / 64 / 128 M bases per chunk)
    // The last chunks taper (1/2, 1/4, 1/8, 1/8 of a chunk): what remains to be computed after the last byte has arrived is
    // the work of the last group, so it should be small.
    std::vector<uint32_t> cuts{0};
    for (uint32_t t0 = 0; t0 < n;) {
        const uint64_t left = nbytes - (offsets[t0] - o0);
        uint64_t want = chunk_bytes;
        if (left <= chunk_bytes + chunk_bytes / 2) want = left > chunk_bytes / 4 ? left / 2 : left;
        want = std::max<uint64_t>(want, std::min<uint64_t>(chunk_bytes, 1u << 20));
        // first t1 in (t0, n] whose transcripts [t0, t1) hold at least `want` bytes
        const uint32_t t1 = (uint32_t)(std::lower_bound(offsets + t0, offsets + n, offsets[t0] + want) - offsets);
        cuts.push_back(std::max(t1, t0 + 1));
        t0 = cuts.back();
    }
    const uint32_t n_chunks = (uint32_t)cuts.size() - 1;
    while (h->ev_chunk.size() < n_chunks) { cudaEvent_t e; CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_chunk.push_back(e); }
    auto chunk_a = [&](uint32_t c) { return offsets[cuts[c]] - o0; };      // byte range of chunk c in the batch
    auto chunk_b = [&](uint32_t c) { return offsets[cuts[c + 1]] - o0; };
    // A pageable caller buffer (Go slice, numpy array) cannot be read by the copy engine: its chunks are staged through a
    // ring of pinned slots by a feeder thread + copy team while the previous slot is on the wire.  Pinned buffers are
    // copied directly.  `issued` = chunks whose copy and event have been queued (the main thread waits for it before it
    // touches a chunk's event).
    // (a pageable packed batch is left to the driver's own staging: a third of the bytes, and hosts that keep a packed
    // transcriptome can keep it pinned)
    const bool stage = !dev_bases && !packed && nbytes > 0 && !getenv("RLSIM_NO_STAGING") && !host_pinned(bases + o0);
    h->staged_pageable = stage;
    std::atomic<uint32_t> issued{0};
    std::atomic<int> feeder_err{0};
    std::thread feeder;
    // on every way out: the feeder is joined and the queued copies have finished reading the caller's buffer
    struct Joiner { std::thread &t; cudaStream_t sc; ~Joiner() { if (t.joinable()) t.join(); cudaStreamSynchronize(sc); } } joiner{feeder, h->st_copy};
    if (!stage) {
        for (uint32_t c = 0; c < n_chunks; c++) {  // all copies are queued first: the copy engine never waits for the host
            if (chunk_b(c) > chunk_a(c) && packed) {
                const uint64_t A = o0 + chunk_a(c), B = o0 + chunk_b(c);  // base range of the chunk in the caller's arrays
                CK(h, cudaMemcpyAsync(h->staging.bases.p + ((A >> 2) - (o0 >> 2)), bases + (A >> 2), ((B + 3) >> 2) - (A >> 2), cudaMemcpyHostToDevice, h->st_copy));
                if (pk_mask)
                    CK(h, cudaMemcpyAsync(h->staging.bases.p + pk_m0 + ((A >> 3) - (o0 >> 3)), pk_mask + (A >> 3), ((B + 7) >> 3) - (A >> 3), cudaMemcpyHostToDevice, h->st_copy));
            } else if (chunk_b(c) > chunk_a(c) && !dev_bases)
                CK(h, cudaMemcpyAsync(h->staging.bases.p + chunk_a(c), bases + o0 + chunk_a(c), chunk_b(c) - chunk_a(c), cudaMemcpyHostToDevice, h->st_copy));
            CK(h, cudaEventRecord(h->ev_chunk[c], h->st_copy));
        }
        issued.store(n_chunks);
    } else {
        constexpr uint32_t RING = 3;
        size_t slot = 0;
        for (uint32_t c = 0; c < n_chunks; c++) slot = std::max<size_t>(slot, chunk_b(c) - chunk_a(c));
        slot = (slot + 4095) & ~(size_t)4095;
        if (h->stage_ring_bytes < RING * slot) {
            if (h->stage_ring) cudaFreeHost(h->stage_ring);
            h->stage_ring = nullptr; h->stage_ring_bytes = 0;
            CK(h, cudaHostAlloc((void **)&h->stage_ring, RING * slot, cudaHostAllocDefault));
            h->stage_ring_bytes = RING * slot;
        }
        if (!h->team) h->team = new CopyTeam(stage_threads());
        const uint8_t *src_host = bases + o0;
        feeder = std::thread([&, slot, src_host] {
            cudaSetDevice(h->device);
            for (uint32_t c = 0; c < n_chunks; c++) {
                cudaError_t e = cudaSuccess;
                if (c >= RING) e = cudaEventSynchronize(h->ev_chunk[c - RING]);  // the slot's previous copy has left it
                const uint64_t a = chunk_a(c), b = chunk_b(c);
                uint8_t *sl = h->stage_ring + (size_t)(c % RING) * slot;
                if (e == cudaSuccess && b > a) {
                    h->team->copy(sl, src_host + a, b - a);
                    e = cudaMemcpyAsync(h->staging.bases.p + a, sl, b - a, cudaMemcpyHostToDevice, h->st_copy);
                }
                if (e == cudaSuccess) e = cudaEventRecord(h->ev_chunk[c], h->st_copy);
                if (e != cudaSuccess) { feeder_err.store((int)e); issued.store(n_chunks, std::memory_order_release); return; }
                issued.store(c + 1, std::memory_order_release);
            }
        });
    }
    auto wait_issued = [&](uint32_t c) { while (issued.load(std::memory_order_acquire) <= c) std::this_thread::yield(); };
    // ---- bookkeeping of the transcripts [t0, t1) of the batch (host arrays; the device copies follow through k_copy_meta)
    rel[0] = 0;
    auto host_meta = [&](uint32_t t0, uint32_t t1) -> int {
        uint64_t *srclen = h->h_srclen.data() + n0, *ex = h->h_expr.data() + n0, *off = h->h_off.data() + n0, *molo = h->h_mol_off.data() + n0;
        uint32_t *len = h->h_len.data() + n0;
        double se = 0, sel = 0;
        uint64_t too_long = 0;
        for (uint32_t i = t0; i < t1; i++) {
            const uint64_t r1 = offsets[i + 1] - o0, l = r1 - rel[i], L = l + pa;
            rel[i + 1] = r1;
            too_long |= L >> 32;
            srclen[i] = l; ex[i] = expr[i]; len[i] = (uint32_t)L;
            if (expr[i] && L > h->max_len && !(L >> 32)) h->max_len = (uint32_t)L;
            off[i + 1] = off[i] + ((L + 31) & ~31ull);
            molo[i + 1] = molo[i] + expr[i];
            se += (double)expr[i];
            sel += (double)expr[i] * (double)L;
        }
        if (too_long) return RLSIM_E_UNSUPPORTED;
        h->sum_expr += se; h->sum_expr_len += sel;
        h->total_padded = off[t1];
        return 0;
    };
    bool async_meta = false;  // asynchronous pipeline: the metadata follow the ingest stream
    auto meta_to_device = [&](uint32_t t0, uint32_t t1) {
        const uint32_t m = t1 - t0;
        k_copy_meta<<<(m + 1 + 255) / 256, 256, 0, async_meta ? h->st_in : h->st>>>(h->h_off.data() + n0 + t0, h->h_len.data() + n0 + t0, h->h_expr.data() + n0 + t0,
                                                         h->h_mol_off.data() + n0 + t0, rel.data() + t0, h->d_off.p + n0 + t0, h->d_len.p + n0 + t0,
                                                         h->d_expr.p + n0 + t0, h->d_mol_off.p + n0 + t0, h->staging.off.p + t0, m);
    };
    auto fail_batch = [&](int code, const char *msg) {  // forget the batch: the handle stays usable with what it held before
        h->h_srclen.resize(n0); h->h_expr.resize(n0); h->h_len.resize(n0); h->h_off.resize(n0 + 1); h->h_mol_off.resize(n0 + 1);
        h->total_padded = old_total;
        cudaStreamSynchronize(h->st);
        h->err = msg;
        return code;
    };
    // With the fused (opt-in) kernels the per-batch decisions need every length first: the whole pass runs up front.
    const bool lazy_meta = !fused_batch;
    if (!lazy_meta) {
        if (host_meta(0, n)) return fail_batch(RLSIM_E_UNSUPPORTED, "transcript longer than 2^32-1 bases");
        meta_to_device(0, n);
    }
    const bool range_fused = !lazy_meta && fused_range_ok(h, n0, nn);
    const bool want_prefix = want_prefix_any && !range_fused;
    const bool want_eager = (want_prefix || range_fused || !needs_fwd(h->p.frag_method)) && eager_possible(h, n0);
    // Eager groups are self-sizing: once a group is done, every chunk that has arrived meanwhile forms the next one, so
    // the per-group fixed cost (host round trips for counts, small sorts) never makes compute fall behind the copies.
    // The grouping does not influence the result (streams are addressed by global molecule / species identity).
    bool eager = want_eager;
    // asynchronous form (default): no host round trip inside the batch; RLSIM_EAGER_SYNC=1 keeps the grouped form above
    const bool async = want_eager && lazy_meta && !getenv("RLSIM_EAGER_SYNC") && (h->eager_async || !h->eager_started);
    std::vector<uint64_t> bound;
    int pending_slot = -1;
    if (async) {  // the chunks' fragment bounds from one pass over lengths and expression levels
        bound.resize(n_chunks);
        uint64_t largest = 1, max_mol = 0, total_bound = 0;
        for (uint32_t c = 0; c < n_chunks; c++) {
            double se = 0, sel = 0;
            uint64_t mol = 0;
            for (uint32_t i = cuts[c]; i < cuts[c + 1]; i++) {
                se += (double)expr[i];
                sel += (double)expr[i] * (double)(offsets[i + 1] - offsets[i] + pa);
                mol += expr[i];
            }
            bound[c] = mol ? eager_chunk_bound(h, sel, se) : 0;
            largest = std::max(largest, bound[c]); max_mol = std::max(max_mol, mol); total_bound += bound[c];
        }
        int rc = eager_async_reserve(h, n_chunks, largest, max_mol, total_bound);
        if (rc) return rc;
        // ingest + prefix sums of a chunk run on their own stream, beside the fragmentation of the chunk before
        if (want_prefix && (rc = ensure_lut(h))) return rc;
        CK(h, cudaEventRecord(h->ev_aux, h->st));
        CK(h, cudaStreamWaitEvent(h->st_in, h->ev_aux, 0));
    }
    cudaStream_t sa = async ? h->st_in : h->st;
    async_meta = async;
    const double t_begin = now_ms();
    if (trace) fprintf(stderr, "[rlsim trace] host prologue %.3f ms (%u chunks)\n", t_begin - t_entry, n_chunks);
    for (uint32_t c = 0; c < n_chunks;) {
        uint32_t c_end = c + 1;
        if (lazy_meta) {  // while chunk c is on the wire
            if (host_meta(cuts[c], cuts[c + 1])) return fail_batch(RLSIM_E_UNSUPPORTED, "transcript longer than 2^32-1 bases");
            meta_to_device(cuts[c], cuts[c + 1]);
        }
        wait_issued(c);
        if (feeder_err.load()) break;
        if (eager && !async) {
            CK(h, cudaEventSynchronize(h->ev_chunk[c]));
            while (c_end < n_chunks && issued.load(std::memory_order_acquire) > c_end && cudaEventQuery(h->ev_chunk[c_end]) == cudaSuccess) {
                if (lazy_meta) {
                    if (host_meta(cuts[c_end], cuts[c_end + 1])) return fail_batch(RLSIM_E_UNSUPPORTED, "transcript longer than 2^32-1 bases");
                    meta_to_device(cuts[c_end], cuts[c_end + 1]);
                }
                c_end++;
            }
            (void)cudaGetLastError();  // cudaErrorNotReady from the query is not an error
        }
        for (uint32_t cc = c; cc < c_end; cc++) {
            const uint32_t t0 = cuts[cc], t1 = cuts[cc + 1];
            CK(h, cudaStreamWaitEvent(sa, h->ev_chunk[cc], 0));
            const bool tm = async && trace && atoi(getenv("RLSIM_TRACE")) >= 2;
            if (tm) h->mark("arrived", (int)cc, sa);
            if (packed)
                launch_ingest_packed(sa, reinterpret_cast<const uint32_t *>(h->staging.bases.p),
                                     pk_mask ? reinterpret_cast<const uint32_t *>(h->staging.bases.p + pk_m0) : nullptr, (uint32_t)(o0 & 3), (uint32_t)(o0 & 7),
                                     h->staging.off.p + t0, t1 - t0, n0 + t0, (h->h_off[n0 + t1] - h->h_off[n0 + t0]) >> 5, h->polya_max,
                                     h->d_off.p + n0 + t0, h->d_bases.p, h->d_packed.p, h->d_nmask.p, h->d_gc.p, h->d_dirty.p);
            else
                launch_ingest(sa, src, h->staging.off.p + t0, t1 - t0, n0 + t0, (h->h_off[n0 + t1] - h->h_off[n0 + t0]) >> 5,
                              h->polya_max, h->d_off.p + n0 + t0, h->d_bases.p, h->d_packed.p, h->d_nmask.p, h->d_gc.p, h->d_dirty.p);
            h->launches[RLSIM_T_LAYOUT] += 3;
            if (tm) h->mark("ingest end", (int)cc, sa);
            if (want_prefix) { int rc = launch_prefix_range(h, n0 + t0, n0 + t1, sa); if (rc) return rc; }
            if (tm) h->mark("prefix end", (int)cc, sa);
        }
        if (async) {  // the fragmentation of this chunk (h->st) follows its ingest + prefix sums (h->st_in)
            CK(h, cudaEventRecord(h->ev_in[c], h->st_in));
            CK(h, cudaStreamWaitEvent(h->st, h->ev_in[c], 0));
        }
        if (eager && async) {
            int slot = -1, rc = eager_front(h, n0 + cuts[c], n0 + cuts[c_end], bound[c], &slot);
            if (rc) return rc;
            if ((rc = eager_back(h, pending_slot))) return rc;  // the chunk before: its count has had a whole front half to arrive
            pending_slot = slot;
            eager = !h->eager_failed;
            if (trace) fprintf(stderr, "[rlsim trace] chunk %u front queued at %.3f ms\n", c, now_ms() - t_begin);
        } else if (eager) {
            if (trace) { const double t_l = now_ms(); cudaStreamSynchronize(h->st); fprintf(stderr, "[rlsim trace]   prev pcr + ingest + prefix %.3f ms\n", now_ms() - t_l); }
            const double t_a = now_ms();
            int rc = eager_chunk(h, n0 + cuts[c], n0 + cuts[c_end]);
            if (rc) return rc;
            eager = !h->eager_failed;
            if (trace) fprintf(stderr, "[rlsim trace] chunks %u..%u arrived %.3f ms, group done %.3f ms\n", c, c_end, t_a - t_begin, now_ms() - t_begin);
        }
        c = c_end;
    }
    if (feeder.joinable()) feeder.join();
    if (feeder_err.load()) { h->err = std::string("staged host-to-device copy: ") + cudaGetErrorString((cudaError_t)feeder_err.load()); return RLSIM_E_CUDA; }
    h->n_mol = h->h_mol_off[nn];
    if (async) {
        CK(h, cudaEventRecord(h->ev_aux, h->st_in));  // chunks past a failed bound were only ingested: h->st follows them too
        CK(h, cudaStreamWaitEvent(h->st, h->ev_aux, 0));
        int rc = eager_back(h, pending_slot);
        if (rc) return rc;
        if ((rc = eager_async_finish(h))) return rc;
        if (trace) fprintf(stderr, "[rlsim trace] last chunk deduplicated %.3f ms\n", now_ms() - t_begin);
        if (!h->trace_marks.empty()) {  // RLSIM_TRACE=2: the device's own timeline, relative to the first mark
            cudaStreamSynchronize(h->st_pcr); cudaStreamSynchronize(h->st_in); cudaStreamSynchronize(h->st);
            for (auto &mk : h->trace_marks) {
                float ms = 0;
                cudaEventElapsedTime(&ms, h->trace_marks[0].ev, mk.ev);
                fprintf(stderr, "[rlsim timeline] %8.3f ms  chunk %2d  %s\n", ms, mk.chunk, mk.what);
                if (&mk != &h->trace_marks[0]) cudaEventDestroy(mk.ev);
            }
            cudaEventDestroy(h->trace_marks[0].ev);
            h->trace_marks.clear();
        }
    }
    CK(h, cudaEventRecord(h->ev_ingest, h->st));  // behind the last k_copy_meta (see below)
    {   // zero words behind the last transcript (k_prefix reads whole words past a transcript's end)
        const uint64_t tot = h->total_padded;
        CK(h, cudaMemsetAsync(h->d_packed.p + tot / 16, 0, 8 * 4, h->st));
        CK(h, cudaMemsetAsync(h->d_nmask.p + tot / 32, 0, 8 * 4, h->st));
    }
    if (h->pcr_in_flight) { CK(h, cudaStreamWaitEvent(h->st, h->ev_pcr, 0)); h->pcr_in_flight = false; }
    CK(h, cudaStreamSynchronize(h->st_copy));  // caller-owned buffers are only read during the call
    // k_copy_meta reads the pinned host arrays, which the next call may move when it grows them: all reads are done here
    CK(h, cudaEventSynchronize(h->ev_ingest));
    if (trace) fprintf(stderr, "[rlsim trace] copies done %.3f ms\n", now_ms() - t_begin);
    tm.stop();
    CK(h, cudaGetLastError());
    if (want_prefix) h->prefix_upto = nn;
    if (want_eager && !h->eager_failed) h->mol_off_valid = true;
    return 0;
}

extern "C" {

int rlsim_add_transcripts(rlsim_handle *h, const uint8_t *bases, const uint64_t *offsets, const uint64_t *expr, uint32_t n) {
    try {
        return add_core(h, bases, nullptr, offsets, expr, n);
    } catch (const std::bad_alloc &) {  // host bookkeeping arrays: no exception crosses the C ABI
        FAIL(h, RLSIM_E_NOMEM, "out of host memory while adding transcripts");
    }
}

// Device-side FASTA reader (k_fasta_in.cu): fasta.go:113-212 + input.go:82-124 in four kernels, then the usual ingest
int rlsim_add_transcripts_packed(rlsim_handle *h, const uint8_t *codes, const uint8_t *nmask, const uint64_t *offsets, const uint64_t *expr,
                                 uint32_t n) {
    if (!codes && n && offsets[n] > offsets[0]) FAIL(h, RLSIM_E_ARG, "rlsim_add_transcripts_packed: codes is NULL");
    try {
        return add_core(h, codes, nullptr, offsets, expr, n, true, nmask);
    } catch (const std::bad_alloc &) {
        FAIL(h, RLSIM_E_NOMEM, "out of host memory while adding transcripts");
    }
}

int rlsim_add_fasta(rlsim_handle *h, const uint8_t *text, uint64_t n_bytes, double expr_mul, uint32_t *n_records, uint32_t *n_added) {
    if (n_records) *n_records = 0;
    if (n_added) *n_added = 0;
    h->fa_n = 0;
    if (!h->have_model) FAIL(h, RLSIM_E_ARG, "rlsim_set_model must precede rlsim_add_fasta (the poly(A) maximum shapes the layout)");
    if (n_bytes == 0) return 0;
    if (text[0] != '>') {  // fasta.go:183-185
        std::string head((const char *)text, (size_t)std::min<uint64_t>(n_bytes, 80));
        FAIL(h, RLSIM_E_ARG, ("Illegal fasta record:\n" + head).c_str());
    }
    cudaSetDevice(h->device);
    CK(h, cudaStreamSynchronize(h->st));
    const uint32_t nt = fasta_tiles(n_bytes);
    CK(h, h->d_fa_text.ensure(n_bytes + 64));
    CK(h, h->d_fa_tiles.ensure((size_t)nt * fasta_tile_bytes() + 64));
    CK(h, h->d_fa_carry.ensure(nt + 1));
    CK(h, h->d_fa_tot.ensure(2));
    StageTimer tm(h, RLSIM_T_LAYOUT, 4);
    CK(h, cudaMemcpyAsync(h->d_fa_text.p, text, n_bytes, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemsetAsync(h->d_error.p, 0, 4, h->st));
    launch_fasta_parse(h->st, h->d_fa_text.p, n_bytes, h->d_fa_tiles.p, h->d_fa_carry.p, h->d_fa_tot.p, nullptr, nullptr, nullptr, nullptr,
                       h->d_error.p, 0);
    CK(h, cudaMemcpyAsync(h->pin + 4, h->d_fa_tot.p, 16, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    CK(h, cudaGetLastError());
    const uint64_t n_rec64 = h->pin[4], kept = h->pin[5];
    if (n_rec64 >= (1ull << 32)) FAIL(h, RLSIM_E_UNSUPPORTED, "more than 2^32-1 FASTA records");
    const uint32_t n_rec = (uint32_t)n_rec64;
    CK(h, h->d_fa_rec_start.ensure(n_rec + 1)); CK(h, h->d_fa_hdr_end.ensure(n_rec + 1)); CK(h, h->d_fa_seq_off.ensure(n_rec + 1));
    CK(h, h->d_fa_seq.ensure(kept + 64));
    CK(h, h->d_fa_hdr_len.ensure(n_rec + 1)); CK(h, h->d_fa_name_len.ensure(n_rec + 1)); CK(h, h->d_fa_level.ensure(n_rec + 1));
    CK(h, h->d_fa_seq_len.ensure(n_rec + 1)); CK(h, h->d_fa_valid.ensure(n_rec + 1));
    CK(h, cudaMemsetAsync(h->d_fa_hdr_end.p, 0xff, (size_t)(n_rec + 1) * 8, h->st));
    launch_fasta_parse(h->st, h->d_fa_text.p, n_bytes, h->d_fa_tiles.p, h->d_fa_carry.p, h->d_fa_tot.p, h->d_fa_rec_start.p,
                       h->d_fa_hdr_end.p, h->d_fa_seq_off.p, h->d_fa_seq.p, h->d_error.p, 1);
    launch_fasta_headers(h->st, h->d_fa_text.p, n_bytes, n_rec, h->d_fa_rec_start.p, h->d_fa_hdr_end.p, h->d_fa_seq_off.p, kept, expr_mul,
                         h->d_fa_hdr_len.p, h->d_fa_name_len.p, h->d_fa_level.p, h->d_fa_seq_len.p, h->d_fa_valid.p);
    h->fa_start.resize(n_rec); h->fa_hdr_len.resize(n_rec); h->fa_name_len.resize(n_rec); h->fa_level.resize(n_rec);
    h->fa_seq_len.resize(n_rec); h->fa_valid.resize(n_rec);
    std::vector<uint64_t> seq_off(n_rec + 1);
    if (n_rec) {
        D2H(h->fa_start.data(), h->d_fa_rec_start.p, (size_t)n_rec * 8); D2H(h->fa_hdr_len.data(), h->d_fa_hdr_len.p, (size_t)n_rec * 4);
        D2H(h->fa_name_len.data(), h->d_fa_name_len.p, (size_t)n_rec * 4); D2H(h->fa_level.data(), h->d_fa_level.p, (size_t)n_rec * 8);
        D2H(h->fa_seq_len.data(), h->d_fa_seq_len.p, (size_t)n_rec * 8); D2H(h->fa_valid.data(), h->d_fa_valid.p, (size_t)n_rec);
        D2H(seq_off.data(), h->d_fa_seq_off.p, (size_t)n_rec * 8);
    }
    CK(h, cudaStreamSynchronize(h->st));
    tm.stop();
    CK(h, cudaGetLastError());
    int rc;
    if ((rc = check_device_error(h))) {
        if (rc == RLSIM_E_UNSUPPORTED) h->err = "FASTA text with a long run of consecutive '>' characters: use the host reader";
        return rc;
    }
    seq_off[n_rec] = kept;
    h->fa_n = n_rec;
    for (uint32_t r = 0; r < n_rec; r++) h->fa_start[r] += 1;  // name line starts after the '>'
    // records with malformed names are skipped (input.go:108-110)
    std::vector<uint64_t> offsets, levels;
    offsets.reserve(n_rec + 1); levels.reserve(n_rec);
    uint32_t n_valid = 0;
    for (uint32_t r = 0; r < n_rec; r++) n_valid += h->fa_valid[r];
    const uint8_t *src = h->d_fa_seq.p;
    if (n_valid == n_rec) {
        offsets.assign(seq_off.begin(), seq_off.end());
        levels.assign(h->fa_level.begin(), h->fa_level.end());
    } else {
        std::vector<uint64_t> from;
        uint64_t pos = 0;
        for (uint32_t r = 0; r < n_rec; r++)
            if (h->fa_valid[r]) { from.push_back(seq_off[r]); offsets.push_back(pos); levels.push_back(h->fa_level[r]); pos += h->fa_seq_len[r]; }
        offsets.push_back(pos);
        if (n_valid) {
            CK(h, h->d_fa_seq2.ensure(pos + 64));
            DBuf<uint64_t> d_from, d_to;
            CK(h, d_from.ensure(n_valid)); CK(h, d_to.ensure(n_valid + 1));
            CK(h, cudaMemcpyAsync(d_from.p, from.data(), (size_t)n_valid * 8, cudaMemcpyHostToDevice, h->st));
            CK(h, cudaMemcpyAsync(d_to.p, offsets.data(), (size_t)(n_valid + 1) * 8, cudaMemcpyHostToDevice, h->st));
            launch_fasta_gather(h->st, h->d_fa_seq.p, d_from.p, d_to.p, n_valid, h->d_fa_seq2.p);
            CK(h, cudaStreamSynchronize(h->st));
            src = h->d_fa_seq2.p;
        }
    }
    if (n_records) *n_records = n_rec;
    if (n_added) *n_added = n_valid;
    if (!n_valid) return 0;
    try {
        return add_core(h, nullptr, src, offsets.data(), levels.data(), n_valid);
    } catch (const std::bad_alloc &) {
        FAIL(h, RLSIM_E_NOMEM, "out of host memory while adding transcripts");
    }
}

int rlsim_get_fasta_index(rlsim_handle *h, uint64_t *name_start, uint32_t *line_len, uint32_t *name_len, uint64_t *level,
                          uint64_t *seq_len, uint8_t *valid) {
    const size_t n = h->fa_n;
    if (name_start) memcpy(name_start, h->fa_start.data(), n * 8);
    if (line_len) memcpy(line_len, h->fa_hdr_len.data(), n * 4);
    if (name_len) memcpy(name_len, h->fa_name_len.data(), n * 4);
    if (level) memcpy(level, h->fa_level.data(), n * 8);
    if (seq_len) memcpy(seq_len, h->fa_seq_len.data(), n * 8);
    if (valid) memcpy(valid, h->fa_valid.data(), n);
    return 0;
}

int rlsim_fragment(rlsim_handle *h) {
    if (!h->have_model) FAIL(h, RLSIM_E_ARG, "rlsim_set_model first");
    if (h->p.n_target == 0 && !h->have_target) FAIL(h, RLSIM_E_ARG, "a raw target needs its sampled histogram before fragmentation");
    cudaSetDevice(h->device);
    int rc = finalize(h, false);  // molecule offsets; the prefix sums follow once the path is chosen
    if (rc) return rc;
    if (h->low == 0) FAIL(h, RLSIM_E_ARG, "Serious trouble: fragment length is zero!");  // transcript.go:194-196
    uint32_t n = (uint32_t)h->h_len.size();
    int method = h->p.frag_method;
    // Resident transcripts, nothing computed yet: run them through the same grouped pipeline rlsim_add_transcripts uses while
    // copying, so that the PCR of one group (FP64 issue bound, own stream) overlaps the memory-bound kernels of the next.
    // Groups hold equal numbers of molecules; any grouping gives the same result (global stream addressing).
    if (n && h->eager_upto == 0 && !h->eager_failed && !h->eager_started && h->n_mol >= (1u << 20) && eager_possible(h, 0) &&
        getenv("RLSIM_GROUPS") && !getenv("RLSIM_NO_GROUPS")) {  // opt-in: measured slower than one launch per stage (profiles/r2_experiments.md)
        const uint32_t G = (uint32_t)std::max(1, atoi(getenv("RLSIM_GROUPS")));
        uint32_t t0 = 0;
        for (uint32_t g = 1; g <= G && t0 < n && !h->eager_failed; g++) {
            uint32_t t1 = n;
            if (g < G) {
                const uint64_t want = h->n_mol / G * g;
                t1 = (uint32_t)(std::lower_bound(h->h_mol_off.begin(), h->h_mol_off.begin() + n + 1, want) - h->h_mol_off.begin());
                t1 = std::min(std::max(t1, t0), n);
            }
            if (t1 == t0) continue;
            if (needs_fwd(method) && !fused_range_ok(h, t0, t1) && h->prefix_upto < t1) {
                if ((rc = grow_prefix(h, h->total_padded, n, h->prefix_upto ? h->total_padded : 0, h->prefix_upto))) return rc;
                StageTimer tm(h, RLSIM_T_PREFIX, 0);
                if ((rc = launch_prefix_range(h, std::max(t0, h->prefix_upto), t1))) return rc;
                tm.stop();
            }
            if ((rc = eager_chunk(h, t0, t1, true))) return rc;
            t0 = t1;
        }
        if (h->pcr_in_flight) { CK(h, cudaStreamWaitEvent(h->st, h->ev_pcr, 0)); h->pcr_in_flight = false; }
        if (!h->eager_failed && needs_fwd(method) && !fused_range_ok(h, 0, n)) h->prefix_upto = n;
    }
    if (!(n && h->eager_upto == n && !h->eager_failed) && (rc = finalize(h))) return rc;  // monolithic path: all prefix sums now
    if (n && h->eager_upto == n && !h->eager_failed) {
        // every chunk was fragmented, deduplicated and amplified while the transcripts were still arriving
        if ((rc = check_device_error(h))) return rc;
        uint32_t hist_frag_n = (uint32_t)h->high + 1, hist_pa_n = h->polya_max + 1;
        h->h_hist_frag.assign(hist_frag_n, 0);
        h->h_hist_polya.assign(hist_pa_n, 0);
        CK(h, cudaMemcpyAsync(h->h_hist_frag.data(), h->d_hist_frag.p, (size_t)hist_frag_n * 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaMemcpyAsync(h->h_hist_polya.data(), h->d_hist_polya.p, (size_t)hist_pa_n * 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        h->n_frag = h->eager_keys;
        h->n_sp = h->eager_sp;
        h->eager_pcr = true;
        h->eager_used = true;
        h->eager_upto = 0; h->eager_failed = true;  // consumed: a repeated call recomputes from the resident transcripts
        h->fragmented = true; h->amplified = false; h->sampled = false;
        return 0;
    }
    h->eager_upto = 0; h->eager_failed = true; h->eager_pcr = false; h->eager_used = false; h->fused_used = false;
    // ---- capacity estimate for the fragment keys
    double loc = h->raw_location;
    if (h->p.n_target > 0) {
        double wsum = 0, lsum = 0, lmin = 1e300;
        for (int i = 0; i < h->p.n_target; i++) { wsum += h->p.target[i].weight; lsum += h->p.target[i].weight * h->p.target[i].location; lmin = std::min(lmin, h->p.target[i].location); }
        loc = lmin > 0 ? lmin : (wsum > 0 ? lsum / wsum : 1.0);
    }
    if (!(loc > 0)) loc = 1.0;
    // expected fragments <= sum over molecules of (Poisson rate + 1): (sum expr*L) / denominator + sum expr
    double denom = method == RLSIM_PRIM_JUMP ? (double)std::max<uint64_t>(h->low, 1)
                 : method == RLSIM_PRE_PRIM ? loc
                 : (h->p.frag_param == 0 ? 2.0 * loc : (double)h->p.frag_param);
    double est = h->sum_expr_len / denom + h->sum_expr;
    uint64_t cap = (uint64_t)(est * 1.15) + 65536;
    uint32_t hist_frag_n = (uint32_t)h->high + 1, hist_pa_n = h->polya_max + 1;
    CK(h, h->d_hist_frag.ensure(hist_frag_n)); CK(h, h->d_hist_polya.ensure(hist_pa_n));
    uint32_t long_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(h->n_mol / 64, 4096), 1u << 24);
    DevModel m = make_model(h);
    DevTranscripts tr = make_tr(h);
    bool queued = (method <= RLSIM_AFTER_NOPRIM_DOUBLE);  // after_* methods go through the interval queue
    uint64_t q_cap = queued ? cap : 1;
    for (int attempt = 0;; attempt++) {
        if (attempt > 8) FAIL(h, RLSIM_E_NOMEM, "fragment buffer kept overflowing");
        CK(h, h->d_keys.ensure(cap));
        CK(h, h->d_q_a.ensure(q_cap)); CK(h, h->d_q_b.ensure(q_cap));
        CK(h, h->d_long_list.ensure(long_cap)); CK(h, h->d_long_hi.ensure(long_cap));
        CK(h, cudaMemsetAsync(h->d_counters.p, 0, 4 * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_long_count.p, 0, 4, h->st));
        CK(h, cudaMemsetAsync(h->d_hist_frag.p, 0, (size_t)hist_frag_n * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_hist_polya.p, 0, (size_t)hist_pa_n * 8, h->st));
        CK(h, cudaMemsetAsync(h->d_error.p, 0, 4, h->st));
        FragBuffers fb{h->d_keys.p, cap, h->d_counters.p, h->d_hist_frag.p, h->d_hist_polya.p, h->d_long_list.p, h->d_long_hi.p,
                       long_cap, h->d_long_count.p, h->d_error.p, h->d_q_a.p, h->d_q_b.p, q_cap, h->d_counters.p + 2};
        const bool fused = fused_range_ok(h, 0, n);
        StageTimer tm(h, RLSIM_T_FRAGMENT, fused ? 0 : 1);
        unsigned int n_long = 0;
        unsigned long long n_q = 0, n_q_dirty = 0;
        int fell_back = 0;
        if (fused) {
            if ((rc = run_fused_range(h, m, tr, fb, 0, n, &n_long, &fell_back))) { tm.stop(); h->fused_disabled = true; attempt--; if ((rc = finalize(h))) return rc; tr = make_tr(h); continue; }
        } else {
            if ((rc = ensure_warp_table(h, tr, 0, h->n_mol))) return rc;
            rc = launch_fragment(h->st, m, tr, 0, h->n_mol, fb, h->d_warp_t.p);
            if (rc) FAIL(h, rc, "fragmentation launch configuration unsupported (poly(A) maximum too large)");
            CK(h, cudaMemcpyAsync(&n_long, h->d_long_count.p, 4, cudaMemcpyDeviceToHost, h->st));
            CK(h, cudaMemcpyAsync(&n_q, h->d_counters.p + 2, 8, cudaMemcpyDeviceToHost, h->st));
            CK(h, cudaMemcpyAsync(&n_q_dirty, h->d_counters.p + 3, 8, cudaMemcpyDeviceToHost, h->st));
            CK(h, cudaStreamSynchronize(h->st));
        }
        if (n_long > long_cap) { tm.stop(); long_cap = n_long + 1024; continue; }
        if (fused && n_long && !fell_back) {
            DevTranscripts trs = tr; trs.sym = h->lut_sym && h->want_mirror ? 1 : 0;
            h->launches[RLSIM_T_FRAGMENT]++;
            launch_fused_long(h->st, m, trs, fb, n_long, FUSED_CAP_LARGE, h->d_wq_f.p, h->d_wq_r.p, h->d_kmax.p, reinterpret_cast<int *>(h->d_work.p + 6));
            CK(h, cudaMemcpyAsync(&fell_back, h->d_work.p + 6, 4, cudaMemcpyDeviceToHost, h->st));
            CK(h, cudaStreamSynchronize(h->st));
        }
        if (fused && fell_back) {  // a transcript's sums do not fit shared memory: HBM prefix sums for this library
            tm.stop();
            h->fused_disabled = true;
            if ((rc = finalize(h))) return rc;
            tr = make_tr(h);
            continue;
        }
        if (fused) h->fused_used = true;
        if (n_q > q_cap) { tm.stop(); q_cap = n_q + n_q / 8 + 65536; continue; }
        if (!fused && n_long) { h->launches[RLSIM_T_FRAGMENT]++; launch_fragment_long(h->st, m, tr, fb, n_long); }
        tm.stop();
        if (n_q) {
            StageTimer tq(h, RLSIM_T_INTERVALS, 1);
            launch_intervals(h->st, m, tr, fb, n_q, n_q_dirty);
            tq.stop();
        }
        CK(h, cudaGetLastError());
        unsigned long long cnt = 0;
        CK(h, cudaMemcpyAsync(&cnt, h->d_counters.p, 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        if (cnt > cap) { cap = cnt + cnt / 8 + 65536; continue; }
        h->n_frag = cnt;
        break;
    }
    if ((rc = check_device_error(h))) return rc;
    h->h_hist_frag.assign(hist_frag_n, 0);
    h->h_hist_polya.assign(hist_pa_n, 0);
    CK(h, cudaMemcpyAsync(h->h_hist_frag.data(), h->d_hist_frag.p, (size_t)hist_frag_n * 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaMemcpyAsync(h->h_hist_polya.data(), h->d_hist_polya.p, (size_t)hist_pa_n * 8, cudaMemcpyDeviceToHost, h->st));
    // ---- species: sort + run-length encode (transcript.go:156-215)
    uint64_t nf = h->n_frag;
    CK(h, h->d_keys_sorted.ensure(nf)); CK(h, h->d_uniq.ensure(nf)); CK(h, h->d_sp_count0.ensure(nf));
    {
        StageTimer tm(h, RLSIM_T_DEDUP, 3 + radix_launches((int)std::min<uint32_t>(64, key_len_bits((uint32_t)h->high) + bits_for(h->total_padded))));
        h->n_sp = 0;
        if (nf) {
            if (nf >= (1ull << 31)) FAIL(h, RLSIM_E_UNSUPPORTED, "more than 2^31 fragments on one GPU");
            int end_bit = (int)std::min<uint32_t>(64, key_len_bits((uint32_t)h->high) + bits_for(h->total_padded));
            size_t tb = 0;
            CK(h, cub::DeviceRadixSort::SortKeys(nullptr, tb, h->d_keys.p, h->d_keys_sorted.p, (int)nf, 0, end_bit, h->st));
            if ((rc = cub_temp(h, tb))) return rc;
            CK(h, cub::DeviceRadixSort::SortKeys(h->d_cub.p, tb, h->d_keys.p, h->d_keys_sorted.p, (int)nf, 0, end_bit, h->st));
            tb = 0;
            CK(h, cub::DeviceRunLengthEncode::Encode(nullptr, tb, h->d_keys_sorted.p, h->d_uniq.p, h->d_sp_count0.p,
                                                     h->d_counters.p + 1, (int)nf, h->st));
            if ((rc = cub_temp(h, tb))) return rc;
            CK(h, cub::DeviceRunLengthEncode::Encode(h->d_cub.p, tb, h->d_keys_sorted.p, h->d_uniq.p, h->d_sp_count0.p,
                                                     h->d_counters.p + 1, (int)nf, h->st));
            unsigned long long runs = 0;
            CK(h, cudaMemcpyAsync(&runs, h->d_counters.p + 1, 8, cudaMemcpyDeviceToHost, h->st));
            CK(h, cudaStreamSynchronize(h->st));
            h->n_sp = runs;
        }
        CK(h, h->d_sp_tr.ensure(h->n_sp)); CK(h, h->d_sp_start.ensure(h->n_sp)); CK(h, h->d_sp_len.ensure(h->n_sp));
        CK(h, h->d_sp_count.ensure(h->n_sp)); CK(h, h->d_sp_eff.ensure(h->n_sp)); CK(h, h->d_sp_gc.ensure(h->n_sp));
        launch_decode_species(h->st, tr, h->d_uniq.p, make_sp(h), (uint32_t)h->high);
        tm.stop();
        CK(h, cudaGetLastError());
    }
    CK(h, cudaStreamSynchronize(h->st));
    h->fragmented = true; h->amplified = false; h->sampled = false;
    return 0;
}

int rlsim_pcr(rlsim_handle *h) {
    if (!h->fragmented) FAIL(h, RLSIM_E_ARG, "rlsim_fragment first");
    cudaSetDevice(h->device);
    int rc;
    const bool had_len_eff = h->len_eff_ready;
    if ((rc = ensure_len_eff(h))) return rc;
    DevModel m = make_model(h);
    DevTranscripts tr = make_tr(h);
    if (h->eager_pcr) {
        h->eager_pcr = false;  // counts were amplified chunk by chunk
    } else {
        CK(h, cudaMemsetAsync(h->d_error.p, 0, 4, h->st));
        StageTimer tm(h, RLSIM_T_PCR, had_len_eff ? 1 : 2);
        launch_pcr(h->st, m, tr, make_sp(h), h->d_error.p);
        tm.stop();
        CK(h, cudaGetLastError());
    }
    if ((rc = check_device_error(h))) return rc;
    // ---- per-length classes (pool.go:156-189): species view sorted by length, prefix sums of the amplified counts
    uint64_t ns = h->n_sp;
    h->n_len = (uint32_t)h->high + 1;
    CK(h, h->d_iota.ensure(ns)); CK(h, h->d_sorted_len.ensure(ns)); CK(h, h->d_order.ensure(ns));
    CK(h, h->d_cls_cnt.ensure(ns)); CK(h, h->d_cls_pre.ensure(ns + 1)); CK(h, h->d_bounds.ensure(h->n_len + 2));
    CK(h, h->d_totals.ensure(h->n_len + 1));
    {
        StageTimer tm(h, RLSIM_T_SELECT, 6 + radix_launches((int)bits_for(h->high)));
        CK(h, cudaMemsetAsync(h->d_cls_pre.p, 0, 8, h->st));
        if (ns) {
            k_iota<<<(unsigned)((ns + 255) / 256), 256, 0, h->st>>>(h->d_iota.p, ns);
            size_t tb = 0;
            int eb = (int)bits_for(h->high);
            CK(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, h->d_sp_len.p, h->d_sorted_len.p, h->d_iota.p, h->d_order.p, (int)ns, 0, eb, h->st));
            if ((rc = cub_temp(h, tb))) return rc;
            CK(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, h->d_sp_len.p, h->d_sorted_len.p, h->d_iota.p, h->d_order.p, (int)ns, 0, eb, h->st));
            launch_gather_counts(h->st, h->d_order.p, h->d_sp_count.p, h->d_cls_cnt.p, ns);
            tb = 0;
            CK(h, cub::DeviceScan::InclusiveSum(nullptr, tb, h->d_cls_cnt.p, h->d_cls_pre.p + 1, (int)ns, h->st));
            if ((rc = cub_temp(h, tb))) return rc;
            CK(h, cub::DeviceScan::InclusiveSum(h->d_cub.p, tb, h->d_cls_cnt.p, h->d_cls_pre.p + 1, (int)ns, h->st));
        }
        launch_class_bounds(h->st, h->d_sorted_len.p, ns, h->n_len, h->d_bounds.p);
        k_class_totals<<<(h->n_len + 127) / 128, 128, 0, h->st>>>(h->d_cls_pre.p, h->d_bounds.p, h->n_len, h->d_totals.p);
        tm.stop();
        CK(h, cudaGetLastError());
    }
    h->h_totals.assign(h->n_len, 0);
    CK(h, cudaMemcpyAsync(h->h_totals.data(), h->d_totals.p, (size_t)h->n_len * 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    // fragstats.go:84-95
    h->pool_total = 0;
    for (uint64_t c : h->h_totals) {
        uint64_t nt = h->pool_total + c;
        if (nt < h->pool_total) FAIL(h, RLSIM_E_POOL_OVERFLOW, "Integer overflow when updating total fragment count!");
        h->pool_total = nt;
    }
    h->h_after_pcr = h->h_totals;
    h->amplified = true; h->sampled = false;
    return 0;
}

int rlsim_build_library(rlsim_handle *h) {
    int rc = rlsim_fragment(h);
    if (rc) return rc;
    return rlsim_pcr(h);
}

int rlsim_class_totals(rlsim_handle *h, uint64_t *totals, uint32_t n_len) {
    if (!h->amplified) FAIL(h, RLSIM_E_ARG, "rlsim_pcr first");
    for (uint32_t l = 0; l < n_len; l++) totals[l] = l < h->h_totals.size() ? h->h_totals[l] : 0;
    return 0;
}

}  // extern "C" (selection helpers follow)

namespace rl_api {

// Size selection core (sampler.go:65-106 + spec 4.6) for the length classes l with l % nparts == part.
//   records == nullptr : single-rank form - drawn ranks are applied to this rank's species directly (rank_base form)
//   records != nullptr : drawn / excluded ranks are written as records grouped by owner rank (multi-GPU form)
int select_core(rlsim_handle *h, const std::vector<uint64_t> &T, const std::vector<uint64_t> &base, uint32_t nparts,
                uint32_t part, const uint64_t *d_all_base, uint4 *records, uint64_t cap, uint64_t *counts_out,
                uint64_t *needed_out, bool sync_records, const uint64_t *all_tot, uint64_t *seg_out) {
    int rc;
    if ((rc = target_host(h))) return rc;
    uint32_t n_len = h->n_len;
    std::vector<double> expect(nparts, 0.0);  // expected records per owner rank (one-sweep emit)
    // ---- sampler.go:65-106 bookkeeping per requested length (global view, identical on every rank)
    size_t nt = std::max<size_t>(h->h_target.size(), n_len);
    h->h_sampled.assign(nt, 0); h->h_missing.assign(nt, 0);
    h->h_after_sampling.assign(nt, 0);
    std::vector<uint8_t> mode(n_len + 1, 0);
    std::vector<uint32_t> c_len;
    std::vector<uint64_t> c_T, c_base, c_G, c_pick, c_cand;
    uint64_t G = 0, n_cand = 0;
    double slack = 1.06;
    for (size_t l = 0; l < nt; l++) {
        uint64_t req = l < h->h_target.size() ? h->h_target[l] : 0;
        uint64_t tot = l < n_len ? T[l] : 0;
        uint64_t take = std::min(req, tot);
        h->h_sampled[l] = take;
        h->h_missing[l] = req - take;
        h->h_after_sampling[l] = tot - take;
        if (!take || l >= n_len) continue;
        if (take == tot) { mode[l] = 1; continue; }
        bool include = take <= tot - take;
        uint64_t pick = include ? take : tot - take;
        mode[l] = include ? 2 : 3;
        if (l % nparts != part) continue;  // another rank draws this class
        c_len.push_back((uint32_t)l); c_T.push_back(tot); c_base.push_back(base[l]); c_G.push_back(G); c_pick.push_back(pick);
        G += tot;
        if (all_tot)  // a uniform draw without replacement splits over the owners in proportion to their copies
            for (uint32_t r = 0; r < nparts; r++) expect[r] += (double)pick * (double)all_tot[(size_t)r * n_len + l] / (double)tot;
        // expected draws until `pick` distinct values out of tot: tot * ln(tot / (tot - pick))
        double e = (double)tot * std::log((double)tot / (double)(tot - pick));
        c_cand.push_back((uint64_t)(e * slack + 6.0 * std::sqrt((double)pick) + 64.0));
    }
    uint64_t ns = h->n_sp;
    DevModel m = make_model(h);
    CK(h, h->d_hits.ensure(ns)); CK(h, h->d_taken.ensure(ns)); CK(h, h->d_out_off.ensure(ns + 1)); CK(h, h->d_mode.ensure(n_len + 1));
    CK(h, cudaMemcpyAsync(h->d_mode.p, mode.data(), n_len + 1, cudaMemcpyHostToDevice, h->st));
    CK(h, h->d_emit.ensure(3 * 64 + 2));
    uint32_t nc = (uint32_t)c_len.size();
    if (counts_out) for (uint32_t r = 0; r < nparts; r++) counts_out[r] = 0;
    if (needed_out) *needed_out = 0;
    for (int attempt = 0;; attempt++) {
        if (attempt > 6) FAIL(h, RLSIM_E_NOMEM, "selection kept running out of candidates");
        CK(h, cudaMemsetAsync(h->d_hits.p, 0, std::max<uint64_t>(ns, 1) * 8, h->st));
        if (!nc) break;
        std::vector<uint64_t> cand_off(nc + 1, 0);
        for (uint32_t c = 0; c < nc; c++) cand_off[c + 1] = cand_off[c] + c_cand[c];
        n_cand = cand_off[nc];
        if (n_cand >= (1ull << 31)) FAIL(h, RLSIM_E_UNSUPPORTED, "more than 2^31 selection candidates on one GPU");
        // plan arrays packed into one u64 buffer: T | base | G | cand_off | pick
        std::vector<uint64_t> plan;
        plan.insert(plan.end(), c_T.begin(), c_T.end());
        plan.insert(plan.end(), c_base.begin(), c_base.end());
        plan.insert(plan.end(), c_G.begin(), c_G.end());
        plan.insert(plan.end(), cand_off.begin(), cand_off.end());
        plan.insert(plan.end(), c_pick.begin(), c_pick.end());
        CK(h, h->d_plan_u64.ensure(plan.size())); CK(h, h->d_plan_len.ensure(nc)); CK(h, h->d_short.ensure(nc)); CK(h, h->d_plan_cut.ensure(nc));
        CK(h, cudaMemcpyAsync(h->d_plan_u64.p, plan.data(), plan.size() * 8, cudaMemcpyHostToDevice, h->st));
        CK(h, cudaMemcpyAsync(h->d_plan_len.p, c_len.data(), nc * 4, cudaMemcpyHostToDevice, h->st));
        CK(h, cudaMemsetAsync(h->d_short.p, 0, nc * 4, h->st));
        DevClassPlan plan_d{nc, h->d_plan_len.p, h->d_plan_u64.p, h->d_plan_u64.p + nc, h->d_plan_u64.p + 2 * nc,
                            h->d_plan_u64.p + 3 * nc, h->d_plan_u64.p + 4 * nc + 1};
        CK(h, h->d_cand_keys.ensure(n_cand)); CK(h, h->d_cand_keys_sorted.ensure(n_cand));
        CK(h, h->d_cand_idx.ensure(n_cand)); CK(h, h->d_cand_idx_sorted.ensure(n_cand));
        CK(h, h->d_flags.ensure(n_cand)); CK(h, h->d_fscan.ensure(n_cand + 1));
        StageTimer tm(h, RLSIM_T_SELECT, 5 + radix_launches((int)bits_for(G)));
        // every key is below G: when G fits 32 bits the sort runs on a narrow copy of the keys (a third less traffic)
        const bool narrow = G <= 0xffffffffull;
        if (narrow) { CK(h, h->d_cand_k32.ensure(n_cand)); CK(h, h->d_cand_k32_sorted.ensure(n_cand)); }
        launch_candidates(h->st, m, plan_d, n_cand, h->d_cand_keys.p, narrow ? h->d_cand_k32.p : nullptr, h->d_cand_idx.p);
        size_t tb = 0;
        // the sort only groups equal keys: when there are at most two candidates per 2^24 values it stops after three radix
        // passes and the few candidates that share their low 24 bits are compared directly (k_first_flags)
        int eb = (int)bits_for(G);
        if (eb > 24 && n_cand <= (1ull << 25) && !getenv("RLSIM_FULL_SORT")) eb = 24;
        if (narrow) {
            CK(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, h->d_cand_k32.p, h->d_cand_k32_sorted.p, h->d_cand_idx.p, h->d_cand_idx_sorted.p, (int)n_cand, 0, eb, h->st));
            if ((rc = cub_temp(h, tb))) return rc;
            CK(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, h->d_cand_k32.p, h->d_cand_k32_sorted.p, h->d_cand_idx.p, h->d_cand_idx_sorted.p, (int)n_cand, 0, eb, h->st));
            launch_first_flags32(h->st, h->d_cand_k32_sorted.p, h->d_cand_idx_sorted.p, n_cand, h->d_flags.p, eb);
        } else {
            CK(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, h->d_cand_keys.p, h->d_cand_keys_sorted.p, h->d_cand_idx.p, h->d_cand_idx_sorted.p, (int)n_cand, 0, eb, h->st));
            if ((rc = cub_temp(h, tb))) return rc;
            CK(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, h->d_cand_keys.p, h->d_cand_keys_sorted.p, h->d_cand_idx.p, h->d_cand_idx_sorted.p, (int)n_cand, 0, eb, h->st));
            launch_first_flags(h->st, h->d_cand_keys_sorted.p, h->d_cand_idx_sorted.p, n_cand, h->d_flags.p, eb);
        }
        CK(h, cudaMemsetAsync(h->d_fscan.p, 0, 8, h->st));
        tb = 0;
        CK(h, cub::DeviceScan::InclusiveSum(nullptr, tb, h->d_flags.p, h->d_fscan.p + 1, (int)n_cand, h->st));
        if ((rc = cub_temp(h, tb))) return rc;
        CK(h, cub::DeviceScan::InclusiveSum(h->d_cub.p, tb, h->d_flags.p, h->d_fscan.p + 1, (int)n_cand, h->st));
        unsigned long long h_counts[64] = {0};
        if (!records) {
            launch_apply_picks_sorted(h->st, plan_d, n_cand, narrow ? nullptr : h->d_cand_keys_sorted.p, narrow ? h->d_cand_k32_sorted.p : nullptr,
                                      h->d_cand_idx_sorted.p, h->d_fscan.p, h->d_plan_cut.p, h->d_cls_pre.p, h->d_bounds.p, h->d_order.p, h->d_hits.p,
                                      h->d_short.p, eb);
        } else if (seg_out && cap) {
            // one sweep: owner groups sized from the expected split (+ 6 sigma); an overflow falls back to count + scatter
            unsigned long long seg[128] = {0}, at = 0;
            for (uint32_t r = 0; r < nparts; r++) {
                seg[r] = at;
                seg[64 + r] = (unsigned long long)(expect[r] + 6.0 * std::sqrt(expect[r] + 1.0) + 256.0);
                at += seg[64 + r];
            }
            if (at <= cap) {
                int *overflow = reinterpret_cast<int *>(h->d_emit.p + 3 * 64);
                CK(h, cudaMemsetAsync(h->d_emit.p, 0, (3 * 64 + 1) * 8, h->st));
                CK(h, cudaMemcpyAsync(h->d_emit.p + 64, seg, 128 * 8, cudaMemcpyHostToDevice, h->st));  // [64,128) offsets, [128,192) capacities
                // cursors live in [0, 64): the kernel reads seg_off[r] and seg_off[64 + r]
                launch_emit_selected(h->st, plan_d, n_cand, h->d_cand_keys.p, h->d_flags.p, h->d_fscan.p, d_all_base, nparts, n_len, 2,
                                     nullptr, h->d_emit.p + 64, h->d_emit.p, records, h->d_short.p, h->rec8 ? 1 : 0, overflow);
                tm.stop();
                CK(h, cudaGetLastError());
                std::vector<int> shortf(nc, 0);
                int ovf = 0;
                CK(h, cudaMemcpyAsync(shortf.data(), h->d_short.p, nc * 4, cudaMemcpyDeviceToHost, h->st));
                CK(h, cudaMemcpyAsync(h_counts, h->d_emit.p, nparts * 8, cudaMemcpyDeviceToHost, h->st));
                CK(h, cudaMemcpyAsync(&ovf, overflow, 4, cudaMemcpyDeviceToHost, h->st));
                CK(h, cudaStreamSynchronize(h->st));
                bool again = false;
                for (uint32_t c = 0; c < nc; c++)
                    if (shortf[c]) { c_cand[c] = c_cand[c] * 2 + 1024; again = true; }
                if (again) continue;
                if (!ovf) {
                    unsigned long long total = 0;
                    for (uint32_t r = 0; r < nparts; r++) { counts_out[r] = h_counts[r]; seg_out[r] = seg[r]; total += h_counts[r]; }
                    *needed_out = total;
                    break;
                }
                // overflow: the split was unusually skewed - redo this attempt with the exact counts (below)
                for (uint32_t r = 0; r < nparts; r++) h_counts[r] = 0;
            }
            seg_out = nullptr;  // this call continues in the two-pass form; the caller sees contiguous groups (seg_out untouched => prefix of counts)
            continue;
        } else {  // count per owner, then scatter into the owner groups
            CK(h, cudaMemsetAsync(h->d_emit.p, 0, 3 * 64 * 8, h->st));
            launch_emit_selected(h->st, plan_d, n_cand, h->d_cand_keys.p, h->d_flags.p, h->d_fscan.p, d_all_base, nparts, n_len, 0,
                                 h->d_emit.p, nullptr, nullptr, nullptr, h->d_short.p);
            CK(h, cudaMemcpyAsync(h_counts, h->d_emit.p, nparts * 8, cudaMemcpyDeviceToHost, h->st));
        }
        tm.stop();
        CK(h, cudaGetLastError());
        std::vector<int> shortf(nc, 0);
        CK(h, cudaMemcpyAsync(shortf.data(), h->d_short.p, nc * 4, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
        bool again = false;
        for (uint32_t c = 0; c < nc; c++)
            if (shortf[c]) { c_cand[c] = c_cand[c] * 2 + 1024; again = true; }
        if (again) continue;
        if (records) {
            unsigned long long seg[64] = {0}, total = 0;
            for (uint32_t r = 0; r < nparts; r++) { seg[r] = total; total += h_counts[r]; counts_out[r] = h_counts[r]; }
            *needed_out = total;
            if (total > cap) return 0;  // caller grows its buffer and calls again
            CK(h, cudaMemcpyAsync(h->d_emit.p + 64, seg, nparts * 8, cudaMemcpyHostToDevice, h->st));
            StageTimer tm2(h, RLSIM_T_SELECT, 1);
            launch_emit_selected(h->st, plan_d, n_cand, h->d_cand_keys.p, h->d_flags.p, h->d_fscan.p, d_all_base, nparts, n_len, 1,
                                 h->d_emit.p, h->d_emit.p + 64, h->d_emit.p + 128, records, h->d_short.p, h->rec8 ? 1 : 0);
            tm2.stop();
            CK(h, cudaGetLastError());
            // h->st is a non-blocking stream: the caller's collective runs on another one, so the records must be
            // complete in memory when this call returns (include/rlsim_gpu.h documents that)
            if (sync_records) CK(h, cudaStreamSynchronize(h->st));
        }
        break;
    }
    return 0;
}

// copies drawn per species -> output records (sampler.go:187-227)
int finish_sample(rlsim_handle *h) {
    int rc;
    uint64_t ns = h->n_sp;
    DevModel m = make_model(h);
    StageTimer tm(h, RLSIM_T_EXPAND, 6);
    launch_taken(h->st, make_sp(h), h->d_mode.p, h->n_len, h->d_hits.p, h->d_taken.p);
    CK(h, cudaMemsetAsync(h->d_out_off.p, 0, 8, h->st));
    h->n_out = 0;
    if (ns) {
        size_t tb = 0;
        CK(h, cub::DeviceScan::InclusiveSum(nullptr, tb, h->d_taken.p, h->d_out_off.p + 1, (int)ns, h->st));
        if ((rc = cub_temp(h, tb))) return rc;
        CK(h, cub::DeviceScan::InclusiveSum(h->d_cub.p, tb, h->d_taken.p, h->d_out_off.p + 1, (int)ns, h->st));
        CK(h, cudaMemcpyAsync(&h->n_out, h->d_out_off.p + ns, 8, cudaMemcpyDeviceToHost, h->st));
        CK(h, cudaStreamSynchronize(h->st));
    }
    CK(h, h->d_o_tr.ensure(h->n_out)); CK(h, h->d_o_start.ensure(h->n_out)); CK(h, h->d_o_end.ensure(h->n_out));
    CK(h, h->d_o_minus.ensure(h->n_out)); CK(h, h->d_o_id.ensure(h->n_out));
    DevSampled out{h->d_o_tr.p, h->d_o_start.p, h->d_o_end.p, h->d_o_minus.p, h->d_o_id.p};
    CK(h, h->d_tr_out_base.ensure(h->h_len.size() + 1));
    if (h->n_out < (1ull << 32)) {  // packed host format (rlsim_get_sampled_packed)
        if (h->high < 32768) { CK(h, h->d_o_ls16.ensure(h->n_out)); out.ls16 = h->d_o_ls16.p; }
        else { CK(h, h->d_o_ls32.ensure(h->n_out)); out.ls32 = h->d_o_ls32.p; }
        CK(h, h->d_o_trcnt.ensure(h->h_len.size() + 1));
        CK(h, cudaMemsetAsync(h->d_o_trcnt.p, 0, (h->h_len.size() + 1) * 4, h->st));
        out.tr_counts = h->d_o_trcnt.p;
    }
    launch_expand(h->st, m, make_sp(h), h->d_out_off.p, h->d_tr_out_base.p, h->n_out, out);
    tm.stop();
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->st));
    h->sampled = true;
    return 0;
}

}  // namespace rl_api

extern "C" {

int rlsim_sample(rlsim_handle *h, const uint64_t *global_totals, const uint64_t *rank_base, uint32_t n_len_in) {
    if (!h->amplified) FAIL(h, RLSIM_E_ARG, "rlsim_pcr first");
    if (!h->have_target) FAIL(h, RLSIM_E_ARG, "no target histogram");
    cudaSetDevice(h->device);
    uint32_t n_len = h->n_len;
    std::vector<uint64_t> T(n_len, 0), base(n_len, 0);
    for (uint32_t l = 0; l < n_len; l++) {
        T[l] = global_totals ? (l < n_len_in ? global_totals[l] : 0) : h->h_totals[l];
        base[l] = rank_base ? (l < n_len_in ? rank_base[l] : 0) : 0;
    }
    int rc = select_core(h, T, base, 1, 0, nullptr, nullptr, 0, nullptr, nullptr);
    if (rc) return rc;
    return finish_sample(h);
}

int rlsim_select_classes(rlsim_handle *h, const uint64_t *all_totals, uint32_t nparts, uint32_t part, uint32_t n_len_in,
                         void *records_dev, uint64_t cap, uint64_t *counts_out, uint64_t *needed_out) {
    return select_partitioned(h, all_totals, nparts, part, n_len_in, records_dev, cap, counts_out, needed_out, true);
}
}  // extern "C"
int rl_api::select_partitioned(rlsim_handle *h, const uint64_t *all_totals, uint32_t nparts, uint32_t part, uint32_t n_len_in,
                               void *records_dev, uint64_t cap, uint64_t *counts_out, uint64_t *needed_out, bool sync_records,
                               uint64_t *seg_out) {
    if (!h->amplified) FAIL(h, RLSIM_E_ARG, "rlsim_pcr first");
    if (!h->have_target) FAIL(h, RLSIM_E_ARG, "no target histogram");
    if (nparts < 1 || nparts > 64 || part >= nparts) FAIL(h, RLSIM_E_ARG, "invalid rank partition");
    if (!records_dev && cap) FAIL(h, RLSIM_E_ARG, "record buffer missing");
    cudaSetDevice(h->device);
    uint32_t n_len = h->n_len;
    // global totals and every rank's first copy rank per class
    std::vector<uint64_t> T(n_len, 0), all_base((size_t)nparts * n_len, 0), mine(n_len, 0);
    for (uint32_t l = 0; l < n_len; l++) {
        uint64_t acc = 0;
        for (uint32_t r = 0; r < nparts; r++) {
            all_base[(size_t)r * n_len + l] = acc;
            uint64_t t = l < n_len_in ? all_totals[(size_t)r * n_len_in + l] : 0;
            uint64_t na = acc + t;
            if (na < acc) FAIL(h, RLSIM_E_POOL_OVERFLOW, "Integer overflow when updating total fragment count!");
            acc = na;
        }
        T[l] = acc;
        mine[l] = all_base[(size_t)part * n_len + l];
    }
    CK(h, h->d_all_base.ensure(all_base.size())); CK(h, h->d_my_base.ensure(n_len));
    CK(h, cudaMemcpyAsync(h->d_all_base.p, all_base.data(), all_base.size() * 8, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemcpyAsync(h->d_my_base.p, mine.data(), (size_t)n_len * 8, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemsetAsync(h->d_error.p, 0, 4, h->st));
    static uint4 dummy_sentinel;
    uint4 *rec = records_dev ? (uint4 *)records_dev : &dummy_sentinel;  // non-null => record form; cap 0 => sizes only
    std::vector<uint64_t> tot_mat;
    if (seg_out) {  // the expected split needs every rank's totals on this handle's length grid
        tot_mat.assign((size_t)nparts * n_len, 0);
        for (uint32_t r = 0; r < nparts; r++)
            for (uint32_t l = 0; l < n_len && l < n_len_in; l++) tot_mat[(size_t)r * n_len + l] = all_totals[(size_t)r * n_len_in + l];
        for (uint32_t r = 0; r < nparts; r++) seg_out[r] = ~0ull;  // stays so when the call ends in the two-pass form
    }
    return select_core(h, T, mine, nparts, part, h->d_all_base.p, rec, records_dev ? cap : 0, counts_out, needed_out, sync_records,
                       seg_out ? tot_mat.data() : nullptr, seg_out);
}
extern "C" {

int rlsim_apply_selected(rlsim_handle *h, const void *records_dev, uint64_t n) {
    if (!h->amplified) FAIL(h, RLSIM_E_ARG, "rlsim_pcr first");
    cudaSetDevice(h->device);
    return apply_records(h, records_dev, n);
}
}  // extern "C"
int rl_api::apply_records(rlsim_handle *h, const void *records_dev, uint64_t n) {
    StageTimer tm(h, RLSIM_T_SELECT, 1);
    launch_apply_records(h->st, (const uint4 *)records_dev, n, h->d_my_base.p, h->d_cls_pre.p, h->d_bounds.p, h->d_order.p,
                         h->d_hits.p, h->d_error.p, h->rec8 ? 1 : 0);
    tm.stop();
    CK(h, cudaGetLastError());
    int e = 0;
    CK(h, cudaMemcpyAsync(&e, h->d_error.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    if (e) FAIL(h, RLSIM_E_ARG, "a selection record was routed to a rank that does not own the copy");
    return 0;
}
extern "C" {

int rlsim_finish_sample(rlsim_handle *h) {
    if (!h->amplified) FAIL(h, RLSIM_E_ARG, "rlsim_pcr first");
    cudaSetDevice(h->device);
    return finish_sample(h);
}

int rlsim_get_stats(rlsim_handle *h, rlsim_stats *out) {
    memset(out, 0, sizeof *out);
    out->n_transcripts = h->h_srclen.size();
    out->n_molecules = h->h_mol_off.empty() ? 0 : h->h_mol_off[h->h_mol_off.size() - 1];  // running sum of the expression levels
    out->n_fragments = h->n_frag;
    out->n_species = h->n_sp;
    out->pool_total = h->pool_total;
    out->n_requested = h->req_frags;
    out->n_sampled = h->sampled ? h->n_out : 0;
    out->low = h->low; out->high = h->high; out->polya_max = h->polya_max;
    out->flags = (needs_rev(h->p.frag_method) && h->lut_valid && h->lut_sym && h->want_mirror ? RLSIM_F_MIRRORED_REVERSE : 0) |
                 (h->eager_used ? RLSIM_F_EAGER : 0) | (h->staged_pageable ? RLSIM_F_PAGEABLE_STAGED : 0) |
                 (h->fused_used ? RLSIM_F_FUSED : 0) |
                 (std::min<uint32_t>(h->eager_groups, 255u) << RLSIM_F_GROUPS_SHIFT);
    return 0;
}

int rlsim_get_hist(rlsim_handle *h, int kind, uint64_t *out, uint32_t n) {
    if (kind == RLSIM_H_TARGET && h->target_pending) { cudaSetDevice(h->device); int rc = target_host(h); if (rc) return rc; }
    const std::vector<uint64_t> *src = nullptr;
    switch (kind) {
    case RLSIM_H_TARGET: src = &h->h_target; break;
    case RLSIM_H_POLYA: src = &h->h_hist_polya; break;
    case RLSIM_H_AFTER_FRAG: src = &h->h_hist_frag; break;
    case RLSIM_H_AFTER_PCR: src = &h->h_after_pcr; break;
    case RLSIM_H_SAMPLED: src = &h->h_sampled; break;
    case RLSIM_H_MISSING: src = &h->h_missing; break;
    case RLSIM_H_AFTER_SAMPLING: src = &h->h_after_sampling; break;
    case RLSIM_H_POOL_GLOBAL: src = &h->h_global_totals; break;
    default: FAIL(h, RLSIM_E_ARG, "unknown histogram kind");
    }
    for (uint32_t i = 0; i < n; i++) out[i] = i < src->size() ? (*src)[i] : 0;
    return 0;
}


int rlsim_get_species(rlsim_handle *h, uint32_t *tr, uint32_t *start, uint32_t *end, uint32_t *count_frag, uint64_t *count_pcr,
                      double *eff, uint32_t *gc) {
    if (!h->fragmented) FAIL(h, RLSIM_E_ARG, "rlsim_fragment first");
    cudaSetDevice(h->device);
    uint64_t n = h->n_sp;
    if (!n) return 0;
    std::vector<uint32_t> len;
    D2H(tr, h->d_sp_tr.p, n * 4);
    D2H(start, h->d_sp_start.p, n * 4);
    D2H(count_frag, h->d_sp_count0.p, n * 4);
    if (end) { len.resize(n); CK(h, cudaMemcpyAsync(len.data(), h->d_sp_len.p, n * 4, cudaMemcpyDeviceToHost, h->st));
               if (!start) FAIL(h, RLSIM_E_ARG, "end needs start"); }
    if (h->amplified) { D2H(count_pcr, h->d_sp_count.p, n * 8); D2H(eff, h->d_sp_eff.p, n * 8); D2H(gc, h->d_sp_gc.p, n * 4); }
    CK(h, cudaStreamSynchronize(h->st));
    if (end) for (uint64_t i = 0; i < n; i++) end[i] = start[i] + len[i];
    if (tr) for (uint64_t i = 0; i < n; i++) tr[i] += h->first_index;
    return 0;
}

int rlsim_get_sampled(rlsim_handle *h, uint32_t *tr, uint32_t *start, uint32_t *end, uint8_t *minus_strand, uint64_t *id) {
    if (!h->sampled) FAIL(h, RLSIM_E_ARG, "rlsim_sample first");
    cudaSetDevice(h->device);
    uint64_t n = h->n_out;
    if (!n) return 0;
    int rc;
    if ((rc = d2h_any(h, tr, h->d_o_tr.p, n * 4)) || (rc = d2h_any(h, start, h->d_o_start.p, n * 4)) ||
        (rc = d2h_any(h, end, h->d_o_end.p, n * 4)) || (rc = d2h_any(h, minus_strand, h->d_o_minus.p, n)) ||
        (rc = d2h_any(h, id, h->d_o_id.p, n * 8)))
        return rc;
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_get_sampled_packed(rlsim_handle *h, uint32_t *tr_counts, uint32_t *start, void *len_strand, int elem_bytes) {
    if (!h->sampled) FAIL(h, RLSIM_E_ARG, "rlsim_sample first");
    if (elem_bytes != 2 && elem_bytes != 4) FAIL(h, RLSIM_E_ARG, "len_strand elements are 2 or 4 bytes wide");
    if (h->n_out >= (1ull << 32)) FAIL(h, RLSIM_E_UNSUPPORTED, "more than 2^32-1 records: use rlsim_get_sampled");
    const bool have16 = h->high < 32768;
    if (elem_bytes == 2 && !have16) FAIL(h, RLSIM_E_ARG, "fragment lengths need 4-byte len_strand elements (high >= 32768)");
    cudaSetDevice(h->device);
    const uint64_t n = h->n_out;
    int rc;
    const size_t ntr = h->h_len.size();
    if (tr_counts && !n) memset(tr_counts, 0, ntr * 4);  // nothing was expanded: the device table is stale
    if (n) {
        if ((rc = d2h_any(h, tr_counts, h->d_o_trcnt.p, ntr * 4)) || (rc = d2h_any(h, start, h->d_o_start.p, n * 4))) return rc;
        if (len_strand) {
            if (elem_bytes == 2) { if ((rc = d2h_any(h, len_strand, h->d_o_ls16.p, n * 2))) return rc; }
            else if (!have16) { if ((rc = d2h_any(h, len_strand, h->d_o_ls32.p, n * 4))) return rc; }
            else {  // 4-byte elements requested although the device holds the 2-byte form: widen on the host
                std::vector<uint16_t> tmp(n);
                CK(h, cudaMemcpyAsync(tmp.data(), h->d_o_ls16.p, n * 2, cudaMemcpyDeviceToHost, h->st));
                CK(h, cudaStreamSynchronize(h->st));
                uint32_t *o = (uint32_t *)len_strand;
                for (uint64_t i = 0; i < n; i++) o[i] = (tmp[i] & 0x7fffu) | ((uint32_t)(tmp[i] >> 15) << 31);
            }
        }
    }
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_format_fasta(rlsim_handle *h, const uint8_t *names, const uint64_t *name_off, uint8_t *buf, uint64_t cap,
                       uint64_t *size_out) {
    if (!h->sampled) FAIL(h, RLSIM_E_ARG, "rlsim_sample first");
    cudaSetDevice(h->device);
    int rc;
    uint64_t n = h->n_out;
    uint32_t ntr = (uint32_t)h->h_len.size();
    *size_out = 0;
    if (!n) return 0;
    CK(h, h->d_name_off.ensure(ntr + 1)); CK(h, h->d_names.ensure(name_off[ntr] + 1));
    CK(h, cudaMemcpyAsync(h->d_name_off.p, name_off, (ntr + 1) * 8, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemcpyAsync(h->d_names.p, names, name_off[ntr], cudaMemcpyHostToDevice, h->st));
    CK(h, h->d_rec_sizes.ensure(n)); CK(h, h->d_rec_off.ensure(n + 1));
    DevSampled s{h->d_o_tr.p, h->d_o_start.p, h->d_o_end.p, h->d_o_minus.p, h->d_o_id.p};
    StageTimer tm(h, RLSIM_T_FORMAT, 3);
    launch_fasta_sizes(h->st, s, n, h->d_name_off.p, h->first_index, h->d_rec_sizes.p);
    CK(h, cudaMemsetAsync(h->d_rec_off.p, 0, 8, h->st));
    size_t tb = 0;
    CK(h, cub::DeviceScan::InclusiveSum(nullptr, tb, h->d_rec_sizes.p, h->d_rec_off.p + 1, (int)n, h->st));
    if ((rc = cub_temp(h, tb))) return rc;
    CK(h, cub::DeviceScan::InclusiveSum(h->d_cub.p, tb, h->d_rec_sizes.p, h->d_rec_off.p + 1, (int)n, h->st));
    uint64_t total = 0;
    CK(h, cudaMemcpyAsync(&total, h->d_rec_off.p + n, 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    *size_out = total;
    if (!buf) { tm.stop(); return 0; }
    if (cap < total) { tm.stop(); FAIL(h, RLSIM_E_ARG, "FASTA buffer too small"); }
    CK(h, h->d_fasta.ensure(total));
    launch_fasta_write(h->st, make_tr(h), s, n, h->d_names.p, h->d_name_off.p, h->d_rec_off.p, h->first_index, h->d_fasta.p);
    tm.stop();
    CK(h, cudaGetLastError());
    if ((rc = d2h_any(h, buf, h->d_fasta.p, total))) return rc;
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

// ---- probes ----
int rlsim_probe_kmer_lut(rlsim_handle *h, double *K_out, uint32_t *w_out, uint32_t n) {
    if (!h->have_model) FAIL(h, RLSIM_E_ARG, "rlsim_set_model first");
    cudaSetDevice(h->device);
    uint32_t nk = 1u << (2 * h->p.kmer_len);
    if (n != nk) FAIL(h, RLSIM_E_ARG, "k-mer table size mismatch");
    CK(h, h->d_K.ensure(nk)); CK(h, h->d_wq_f.ensure(nk)); CK(h, h->d_wq_r.ensure(nk));
    launch_kmer_lut(h->st, h->p.kmer_len, h->p.priming_temp, h->d_K.p, h->d_kmax.p, h->d_wq_f.p, h->d_wq_r.p, h->d_work.p + 2);
    CK(h, cudaGetLastError());
    D2H(K_out, h->d_K.p, (size_t)nk * 8);
    D2H(w_out, h->d_wq_f.p, (size_t)nk * 4);
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_probe_prefix(rlsim_handle *h, uint32_t tr, int reverse, uint64_t *out, uint32_t n) {
    if (tr >= h->h_len.size()) FAIL(h, RLSIM_E_ARG, "transcript index out of range");
    if (!needs_fwd(h->p.frag_method)) FAIL(h, RLSIM_E_ARG, "no priming prefix sums for this fragmentation method");
    if (reverse && !needs_rev(h->p.frag_method)) FAIL(h, RLSIM_E_ARG, "no reverse profile for this method");
    cudaSetDevice(h->device);
    if (h->prefix_upto <= tr) {  // the fused path keeps the sums in shared memory: materialise them in HBM for the probe
        const uint32_t nt = (uint32_t)h->h_len.size();
        int rcg = grow_prefix(h, h->total_padded, nt, h->prefix_upto ? h->total_padded : 0, h->prefix_upto);
        if (rcg) return rcg;
        if ((rcg = launch_prefix_range(h, h->prefix_upto, nt))) return rcg;
        CK(h, cudaStreamSynchronize(h->st));
        CK(h, cudaGetLastError());
        h->prefix_upto = nt;
    }
    // P(0..proflen) evaluated on the device by the profile the searches use (record layout, two-level arrays, mirrored)
    DBuf<uint64_t> d_out;
    CK(h, d_out.ensure(std::max<uint32_t>(n, 1)));
    launch_probe_prefix(h->st, make_model(h), make_tr(h), tr, reverse, d_out.p, n);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(out, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_probe_gc_prefix(rlsim_handle *h, uint32_t tr, uint32_t *out, uint32_t n) {
    if (tr >= h->h_len.size()) FAIL(h, RLSIM_E_ARG, "transcript index out of range");
    cudaSetDevice(h->device);
    uint32_t m = std::min(n, h->h_len[tr] + 1);
    std::vector<uint2> g((size_t)(h->h_len[tr] >> 5) + 1);
    CK(h, cudaMemcpyAsync(g.data(), h->d_gc.p + (h->h_off[tr] >> 5) + tr, g.size() * sizeof(uint2), cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    for (uint32_t x = 0; x < m; x++) out[x] = g[x >> 5].x + (uint32_t)__builtin_popcount(g[x >> 5].y & ((1u << (x & 31)) - 1u));
    return 0;
}

int rlsim_probe_math(rlsim_handle *h, int fn, const double *x, const double *y, double *out, uint32_t n) {
    cudaSetDevice(h->device);
    DBuf<double> dx, dy, dout;
    CK(h, dx.ensure(n)); CK(h, dy.ensure(n)); CK(h, dout.ensure(n));
    CK(h, cudaMemcpyAsync(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, h->st));
    if (y) CK(h, cudaMemcpyAsync(dy.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, h->st));
    else CK(h, cudaMemsetAsync(dy.p, 0, (size_t)n * 8, h->st));
    launch_math_probe(h->st, fn, dx.p, dy.p, dout.p, n);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_probe_philox(rlsim_handle *h, const uint32_t *ctr_key6, uint32_t *out4, uint32_t n) {
    cudaSetDevice(h->device);
    DBuf<uint32_t> din, dout;
    CK(h, din.ensure((size_t)n * 6)); CK(h, dout.ensure((size_t)n * 4));
    CK(h, cudaMemcpyAsync(din.p, ctr_key6, (size_t)n * 24, cudaMemcpyHostToDevice, h->st));
    launch_philox_probe(h->st, din.p, dout.p, n);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(out4, dout.p, (size_t)n * 16, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_probe_amplify(rlsim_handle *h, int64_t seed_pcr, const uint32_t *tse, const uint64_t *n0, const double *p, int cycles,
                        uint64_t *out, uint32_t n) {
    cudaSetDevice(h->device);
    DBuf<uint32_t> dt;
    DBuf<uint64_t> dn, dout;
    DBuf<double> dp;
    CK(h, dt.ensure((size_t)n * 3)); CK(h, dn.ensure(n)); CK(h, dout.ensure(n)); CK(h, dp.ensure(n));
    CK(h, cudaMemcpyAsync(dt.p, tse, (size_t)n * 12, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemcpyAsync(dn.p, n0, (size_t)n * 8, cudaMemcpyHostToDevice, h->st));
    CK(h, cudaMemcpyAsync(dp.p, p, (size_t)n * 8, cudaMemcpyHostToDevice, h->st));
    launch_amplify_probe(h->st, px_derive_key(seed_pcr, TAG_PCR), dt.p, dn.p, dp.p, cycles, dout.p, n);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

int rlsim_get_timings(rlsim_handle *h, float *ms_out, uint64_t *launches_out) {
    cudaSetDevice(h->device);
    drain_timers(h);
    for (int i = 0; i < RLSIM_T_COUNT; i++) {
        if (ms_out) ms_out[i] = h->ms[i];
        if (launches_out) launches_out[i] = h->launches[i];
    }
    return 0;
}

}  // extern "C"

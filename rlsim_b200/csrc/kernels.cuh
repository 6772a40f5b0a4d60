// Launch prototypes of the kernel translation units (host side glue in api.cu).
#pragma once
#include "common.cuh"

namespace rl {

// k_transcripts.cu
void launch_ingest(cudaStream_t st, const uint8_t *src, const uint64_t *src_off, uint32_t n, uint32_t t_index0, uint64_t n_groups,
                   uint32_t polya_max, const uint64_t *off, uint8_t *bases, uint32_t *packed, uint32_t *nmask, uint2 *gc,
                   uint32_t *dirty);
void launch_ingest_packed(cudaStream_t st, const uint32_t *codes, const uint32_t *mask, uint32_t c_shift, uint32_t m_shift,
                          const uint64_t *src_off, uint32_t n, uint32_t t_index0, uint64_t n_groups, uint32_t polya_max,
                          const uint64_t *off, uint8_t *bases, uint32_t *packed, uint32_t *nmask, uint2 *gc, uint32_t *dirty);
void launch_kmer_lut(cudaStream_t st, int k, double T, double *K, unsigned long long *kmax_bits, uint32_t *wq_f,
                     uint32_t *wq_r, unsigned int *asym);
void launch_prefix(cudaStream_t st, const DevTranscripts &tr, uint32_t t_begin, uint32_t t_end, int k, double T, const uint32_t *wq_f, const uint32_t *wq_r,
                   const unsigned long long *kmax_bits, uint32_t *Ff, uint32_t *Fr, uint64_t *Cf, uint64_t *Cr, Rec32 *Rf, int strands,
                   unsigned int *work, uint32_t *long_list = nullptr, uint32_t long_cap = 0, bool bulk_stores = false);
void launch_probe_prefix(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, uint32_t t, int reverse, uint64_t *out, uint32_t n);

// k_fragment.cu
struct FragBuffers {
    uint64_t *keys; uint64_t cap; unsigned long long *count;       // fragment keys out
    unsigned long long *hist_frag, *hist_polya;                    // histograms by length
    uint2 *long_list; uint32_t *long_hi; uint32_t long_cap; unsigned int *long_count;  // deferred long molecules
    int *error;
    uint4 *q_a; uint2 *q_b; uint64_t q_cap; unsigned long long *q_count;               // interval work queue
};
// warp_t: optional table from launch_warp_transcripts over the same molecule range (FRAG_THREADS is a multiple of 32)
int launch_fragment(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, uint64_t g_begin, uint64_t g_end,
                    const FragBuffers &fb, const uint32_t *warp_t = nullptr);
void launch_warp_transcripts(cudaStream_t st, const DevTranscripts &tr, uint64_t g_begin, uint64_t g_end, uint32_t *warp_t);
// n_q_dev / n_long_dev: the counts stay on the device (asynchronous chunk pipeline); n_q / n_long then only bound the grid
int launch_intervals(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint64_t n_q, uint64_t n_q_dirty,
                     const unsigned long long *n_q_dev = nullptr);
int launch_fragment_long(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint32_t n_long,
                         const unsigned int *n_long_dev = nullptr);
// fused prefix + fragmentation (after_prim / after_prim_double): prefix sums in shared memory
struct FusedItem { uint32_t t0, t1; uint64_t g0, g1; };  // transcripts [t0, t1), global molecules [g0, g1)
constexpr uint32_t FUSED_CAP_SMALL = 20480;   // prefix entries a CTA holds, two CTAs per SM
constexpr uint32_t FUSED_CAP_LARGE = 49152;   // ... one CTA per SM (transcripts up to 49 k bases)
constexpr uint32_t FUSED_LONG_BREAKS = 1024;  // expected breakpoints per molecule the deferred-molecule kernel is trusted with
int launch_fused(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, const FusedItem *items,
                 uint32_t n_items, uint32_t cap_entries, unsigned int *work, const uint32_t *wq_f, const uint32_t *wq_r,
                 const unsigned long long *kmax_bits, int *fallback, int sm_count);
// warp-per-transcript variant with the compact shared-memory profile; unfit transcripts go to `redo` (served by launch_fused)
constexpr uint32_t FUSED_WARP_CAP_S = 4096, FUSED_WARP_CAP_M = 16384, FUSED_WARP_CAP_L = 49152;
int launch_fused_warp(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint64_t g_begin, uint64_t g_end,
                      uint32_t cap_lo, uint32_t cap_entries, int warps, unsigned int *work, const uint32_t *wq_f, FusedItem *redo,
                      unsigned int *redo_count, uint32_t redo_cap, int sm_count);
uint64_t fused_warp_units(uint64_t n_mol);
int launch_fused_long(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, const FragBuffers &fb, uint32_t n_long,
                      uint32_t cap_entries, const uint32_t *wq_f, const uint32_t *wq_r, const unsigned long long *kmax_bits, int *fallback);

// k_pcr.cu
struct DevSpecies {
    uint64_t n;
    uint32_t *tr, *start, *len;   // sorted by (tr, start, len)
    uint32_t *count0;             // copies after fragmentation
    uint64_t *count;              // copies after PCR
    double *eff;                  // optional (tests): per-species efficiency
    uint32_t *gc;                 // optional (tests): GC count
};
void launch_decode_species(cudaStream_t st, const DevTranscripts &tr, const uint64_t *keys, DevSpecies sp, uint32_t high);
// Asynchronous chunk pipeline (api.cu: eager_front / eager_back): only a chunk's fragment count visits the host.
struct EagerChunk {            // one per chunk, device memory, zeroed before use
    unsigned long long key_count;   // fragment keys written into the chunk's key buffer (may exceed its capacity: overflow)
    unsigned long long runs;        // run count of the run-length pass = species of the chunk
    unsigned long long q_count, q_dirty;  // interval work queue
    unsigned long long sp_base, sp_n;     // the chunk's species: [sp_base, sp_base + sp_n) of the species table
    unsigned int long_count, pad;
};
struct EagerGlobal { unsigned long long sp_cursor, frag_total, fail, pad; };
constexpr uint32_t EAGER_BUFS = 4;        // key buffers the chunks rotate through
void launch_species_begin(cudaStream_t st, EagerChunk *cs, EagerGlobal *g, uint64_t key_cap, uint64_t q_cap,
                          uint32_t long_cap, uint64_t sp_cap);
// sp: whole-table base pointers; the chunk's rows are cs->sp_base ...; n_bound threads are launched
void launch_decode_species_dyn(cudaStream_t st, const DevTranscripts &tr, const uint64_t *uniq, const uint32_t *run_len, DevSpecies sp,
                               uint32_t high, const EagerChunk *cs, uint64_t n_bound);
void launch_pcr_dyn(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, DevSpecies sp, int *error, const EagerChunk *cs, uint64_t n_bound);
void launch_len_eff(cudaStream_t st, const DevModel &m, double *len_eff);
void launch_pcr(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, DevSpecies sp, int *error);
void launch_amplify_probe(cudaStream_t st, PxKey key, const uint32_t *tse, const uint64_t *n0, const double *p, int cycles,
                          uint64_t *out, uint32_t n);
void launch_math_probe(cudaStream_t st, int fn, const double *x, const double *y, double *out, uint32_t n);
void launch_philox_probe(cudaStream_t st, const uint32_t *ck, uint32_t *out, uint32_t n);

// k_select.cu
void launch_target(cudaStream_t st, const DevModel &m, uint64_t first, uint64_t n, unsigned long long *hist, uint32_t hist_n);
void launch_gather_counts(cudaStream_t st, const uint32_t *order, const uint64_t *count, uint64_t *out, uint64_t n);
void launch_class_bounds(cudaStream_t st, const uint32_t *sorted_len, uint64_t n, uint32_t n_len, uint64_t *bounds);
struct DevClassPlan {          // per requested length class (host-built, device-resident)
    uint32_t n_classes;
    const uint32_t *len;       // class length
    const uint64_t *T;         // global total of the class
    const uint64_t *base;      // this rank's first global rank within the class
    const uint64_t *G;         // exclusive prefix of T over the classes (sort-key offset)
    const uint64_t *cand_off;  // [n_classes+1] candidate index range of the class
    const uint64_t *pick;      // number of distinct ranks to keep
};
void launch_candidates(cudaStream_t st, const DevModel &m, DevClassPlan plan, uint64_t n_cand, uint64_t *keys, uint32_t *keys32,
                       uint32_t *idx);
// sort_bits: the low key bits the stable sort covered (runs that agree in them are scanned for equal keys)
void launch_first_flags(cudaStream_t st, const uint64_t *sorted_keys, const uint32_t *sorted_idx, uint64_t n, uint32_t *flags, int sort_bits);
void launch_first_flags32(cudaStream_t st, const uint32_t *sorted_keys, const uint32_t *sorted_idx, uint64_t n, uint32_t *flags, int sort_bits);
void launch_apply_picks(cudaStream_t st, DevClassPlan plan, uint64_t n_cand, const uint64_t *keys_by_idx,
                        const uint32_t *flags, const uint64_t *flag_scan, const uint64_t *cls_pre,
                        const uint64_t *bounds, const uint32_t *order, unsigned long long *hits, int *short_flag);
void launch_apply_picks_sorted(cudaStream_t st, DevClassPlan plan, uint64_t n_cand, const uint64_t *sk64, const uint32_t *sk32,
                               const uint32_t *sidx, const uint64_t *fscan, uint64_t *cut, const uint64_t *cls_pre, const uint64_t *bounds,
                               const uint32_t *order, unsigned long long *hits, int *short_flag, int sort_bits);
void launch_emit_selected(cudaStream_t st, DevClassPlan plan, uint64_t n_cand, const uint64_t *keys, const uint32_t *flags,
                          const uint64_t *fscan, const uint64_t *all_base, uint32_t nparts, uint32_t n_len, int pass,
                          unsigned long long *counts, const unsigned long long *seg_off, unsigned long long *cursor,
                          uint4 *records, int *short_flag, int compact = 0, int *overflow = nullptr);
void launch_apply_records(cudaStream_t st, const uint4 *records, uint64_t n, const uint64_t *my_base, const uint64_t *cls_pre,
                          const uint64_t *bounds, const uint32_t *order, unsigned long long *hits, int *error, int compact = 0);
void launch_taken(cudaStream_t st, const DevSpecies &sp, const uint8_t *mode_by_len, uint32_t n_len,
                  const unsigned long long *hits, uint64_t *taken);
struct DevSampled {
    uint32_t *tr, *start, *end; uint8_t *minus; uint64_t *id;
    // packed host format (rlsim_get_sampled_packed); any of them may be null
    uint16_t *ls16 = nullptr; uint32_t *ls32 = nullptr; uint32_t *tr_counts = nullptr;  // tr_counts: zeroed by the caller
};
void launch_expand(cudaStream_t st, const DevModel &m, const DevSpecies &sp, const uint64_t *out_off, uint64_t *tr_out_base,
                   uint64_t n_out, DevSampled out);

// k_fasta_in.cu
uint32_t fasta_tiles(uint64_t n);
size_t fasta_tile_bytes();
void launch_fasta_parse(cudaStream_t st, const uint8_t *text, uint64_t n, void *tiles, uint4 *carry, unsigned long long *totals,
                        uint64_t *rec_start, uint64_t *hdr_end, uint64_t *seq_off, uint8_t *seq, int *err, int pass);
void launch_fasta_headers(cudaStream_t st, const uint8_t *text, uint64_t n, uint32_t n_rec, const uint64_t *rec_start,
                          const uint64_t *hdr_end, const uint64_t *seq_off, uint64_t total_kept, double expr_mul,
                          uint32_t *hdr_len, uint32_t *name_len, uint64_t *level, uint64_t *seq_len, uint8_t *valid);
void launch_fasta_gather(cudaStream_t st, const uint8_t *src, const uint64_t *from, const uint64_t *to, uint32_t n, uint8_t *dst);

// k_fasta.cu
void launch_fasta_sizes(cudaStream_t st, const DevSampled &s, uint64_t n, const uint64_t *name_off, uint32_t first_index,
                        uint64_t *sizes);
void launch_fasta_write(cudaStream_t st, const DevTranscripts &tr, const DevSampled &s, uint64_t n, const uint8_t *names,
                        const uint64_t *name_off, const uint64_t *rec_off, uint32_t first_index, uint8_t *out);

}  // namespace rl

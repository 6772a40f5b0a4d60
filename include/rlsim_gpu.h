/* include/rlsim_gpu.h - C ABI of librlsim_gpu.so, the B200 implementation of rlsim's
 * library-construction hot path.
 *
 * The reference (sbotond/rlsim, Go) has no FFI boundary; its seams are the Go interfaces wired in
 * main() (/root/reference/src/main.go:88-152).  The GPU path plugs in at the two *batch* methods
 *     Pooler.InitTranscripts    (/root/reference/src/pool.go:83-135)
 *     Sampler.SampleFragments   (/root/reference/src/sampler.go:108-154)
 * and this header declares exactly what a cgo shim implementing those two methods binds
 * (INTEGRATION.md shows the Go side).  Plain pointers and sizes only; every input buffer is
 * caller-owned and only read during the call (cgo pointer rules); outputs are copied into
 * caller-provided buffers.  All functions return 0 on success or an RLSIM_E_* code; the text of
 * the last error is available from rlsim_last_error() - the Go shim turns any non-zero code into
 * L.Fatalf, the reference's only error policy (/root/reference/src/logging.go:67-73).
 *
 * Threading: one handle = one GPU = one host thread at a time.  No callbacks.
 */
#ifndef RLSIM_GPU_H
#define RLSIM_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLSIM_MAX_COMPS 8

/* error codes; the overflow ones carry the reference's own messages */
enum {
    RLSIM_OK = 0,
    RLSIM_E_CUDA = 1,         /* CUDA runtime failure */
    RLSIM_E_ARG = 2,          /* invalid argument / call order */
    RLSIM_E_PRIMING = 3,      /* "Error when simulating priming!"                          nnthermo.go:213 */
    RLSIM_E_PCR_OVERFLOW = 4, /* "Integer overflow detected when amplifying fragments!"    thermocycler.go:142-144 */
    RLSIM_E_POOL_OVERFLOW = 5,/* "Integer overflow when updating total fragment count!"    fragstats.go:88-94 */
    RLSIM_E_UNSUPPORTED = 6,  /* outside the GPU path's limits (documented in DESIGN.md) */
    RLSIM_E_NOMEM = 7
};

/* fragmentation methods, /root/reference/src/fragmentor.go:39-60 */
enum { RLSIM_AFTER_PRIM = 0, RLSIM_AFTER_PRIM_DOUBLE = 1, RLSIM_AFTER_NOPRIM = 2, RLSIM_AFTER_NOPRIM_DOUBLE = 3,
       RLSIM_PRE_PRIM = 4, RLSIM_PRIM_JUMP = 5 };
/* mixture component types, /root/reference/src/target_mix.go:61-67 */
enum { RLSIM_COMP_N = 0, RLSIM_COMP_SN = 1, RLSIM_COMP_G = 2 };

/* one "w:type:(loc, scale[, shape], low, high)" component, target_mix.go:37-44 */
typedef struct {
    int32_t type;
    int32_t reserved;
    double weight, location, scale, shape;
    uint64_t low, high;
} rlsim_mixcomp;

/* Everything NewFragmentor / NewTechne / NewTarget / NewLenSampler receive in main.go:98-127. */
typedef struct {
    int32_t kmer_len;        /* -k   args.go:107 */
    int32_t frag_method;     /* -f   args.go:277-307 */
    int32_t frag_param;      /*      method:param */
    int32_t nr_cycles;       /* -c */
    double priming_temp;     /* -p */
    double frag_loss_prob;   /* -flg fragfilter.go:45 */
    double fixed_eff;        /* -e   thermocycler.go:134-137 */
    double strand_bias;      /* -b   sampler.go:219-227 (probability of "-") */
    int32_t gc_mode;         /* 0 = parametric -eg (thermocycler.go:173-177), 1 = raw table (:169-171) */
    int32_t n_target;        /* components of -d; 0 = raw target (-j), see rlsim_set_raw_target */
    int32_t n_polya;         /* components of -a */
    int32_t reserved;
    double gc_shape, gc_min, gc_max;
    double gc_raw[101];      /* after ProcessRawGcEffs, thermocycler.go:94-104 */
    double len_shape, len_min, len_max; /* -el; used only when no length-efficiency table is supplied */
    rlsim_mixcomp target[RLSIM_MAX_COMPS];
    rlsim_mixcomp polya[RLSIM_MAX_COMPS];
    int64_t seed_init, seed_pcr, seed_sampling; /* -si -sp -ss; 0 = not given (main.go:72-81,129-149) */
} rlsim_params;

/* per-run scalar results (fragstats.go:45-56 totals + bookkeeping) */
typedef struct {
    uint64_t n_transcripts;     /* records added (expression 0 included) */
    uint64_t n_molecules;       /* sum of expression levels */
    uint64_t n_fragments;       /* fragments kept by fragmentation (sum of "after fragmentation") */
    uint64_t n_species;         /* distinct (transcript,start,end) */
    uint64_t pool_total;        /* FragStats.TotalFrags */
    uint64_t n_requested;       /* Targeter.GetReqFrags */
    uint64_t n_sampled;         /* fragments in the result of rlsim_sample */
    uint64_t low, high;         /* Targeter.GetLow / GetHigh */
    uint32_t polya_max;         /* pool.go:84 */
    uint32_t flags;             /* RLSIM_F_*: how the library was built (informational) */
} rlsim_stats;
/* rlsim_stats.flags */
enum { RLSIM_F_MIRRORED_REVERSE = 1,  /* reverse priming reads the forward prefix sums mirrored (symmetric k-mer table) */
       RLSIM_F_EAGER = 2,             /* the last library was fragmented and amplified while its transcripts were copied */
       RLSIM_F_GROUPS_SHIFT = 8,      /* bits 8..15: number of chunk groups the eager pipeline ran (saturating at 255) */
       RLSIM_F_PAGEABLE_STAGED = 4,   /* the last rlsim_add_transcripts staged a pageable caller buffer through pinned memory */
       RLSIM_F_FUSED = 16 };          /* fragmentation kept the priming prefix sums in shared memory (fused kernel) */

/* histogram kinds (fragstats.go:45-56); every histogram is indexed by length */
enum { RLSIM_H_TARGET = 0, RLSIM_H_POLYA = 1, RLSIM_H_AFTER_FRAG = 2, RLSIM_H_AFTER_PCR = 3, RLSIM_H_SAMPLED = 4,
       RLSIM_H_MISSING = 5, RLSIM_H_AFTER_SAMPLING = 6,
       RLSIM_H_POOL_GLOBAL = 7 };  /* after PCR, summed over the ranks (rlsim_sample_distributed) */

typedef struct rlsim_handle rlsim_handle;

/* ---- lifecycle: NewPool (pool.go:66-81) / Pool.Cleanup (pool.go:210-221) ---- */
int rlsim_create(int device, rlsim_handle **out);
void rlsim_destroy(rlsim_handle *h);
const char *rlsim_last_error(const rlsim_handle *h);   /* h may be NULL: error of a failed rlsim_create */

/* ---- model: NewFragmentor + NewTechne + NewLenSampler (main.go:98-123) ----
 * len_eff: Techne.CalcLengthEff(l) for l = low..high (thermocycler.go:149-156), filled by the host with
 * the same code that fills the report so both agree bit for bit; NULL = computed on the device from
 * len_shape/min/max with the library's restatement of Go's math.Pow. */
int rlsim_set_model(rlsim_handle *h, const rlsim_params *p);
int rlsim_set_len_eff(rlsim_handle *h, const double *len_eff, uint64_t low, uint64_t high);
/* raw (-j) target, parse_raw.go:83-104: lengths and normalised probabilities */
int rlsim_set_raw_target(rlsim_handle *h, const uint32_t *lengths, const double *probs, uint32_t n);

/* ---- Targeter (target.go:102-132 / raw_target.go:104-138) ----
 * Either draw the req_frags target lengths on the device ... (the draws run on a stream of their own beside the
 * transcript pipeline; the histogram reaches the host when it is first needed: selection, rlsim_get_hist) */
int rlsim_sample_target(rlsim_handle *h, int64_t req_frags);
/* Multi-GPU: draw only the target lengths with draw index in [first, first+count) (the stream is addressed by the draw
 * index); read the partial histogram with rlsim_get_hist(RLSIM_H_TARGET), sum it over the ranks and impose the sum with
 * rlsim_set_target_hist. */
int rlsim_sample_target_range(rlsim_handle *h, uint64_t first, uint64_t count);
/* ... or impose the histogram the host's own Targeter sampled (NextLenCount pairs). */
int rlsim_set_target_hist(rlsim_handle *h, const uint32_t *lengths, const uint64_t *counts, uint32_t n);

/* ---- Pooler.InitTranscripts (pool.go:83-135) ----
 * add_transcripts replaces the `chan *Transcript` loop (input.go:103-124): `bases` = concatenated
 * upper-cased sequences (no poly-A), offsets[n+1] into it, expr[n] = expression levels with -m applied.
 * May be called repeatedly (batches).  first_index = global index of the first record of this process's
 * shard (0 on a single GPU); it keys the per-transcript random streams so any sharding reproduces the
 * same fragments.
 * When the model, the target (and, if the host supplies it, the length-efficiency table) and first_index are already set,
 * the library does not wait for rlsim_build_library: every group of 32 MB chunks that has reached the device is
 * fragmented, deduplicated and amplified while the following chunks are still being copied, and rlsim_build_library only
 * finishes the per-length view (rlsim_stats.flags & RLSIM_F_EAGER).  The result does not depend on that. */
int rlsim_add_transcripts(rlsim_handle *h, const uint8_t *bases, const uint64_t *offsets, const uint64_t *expr,
                          uint32_t n);
/* The same batch from a host that keeps its transcriptome packed (0.375 B per base on the host link instead of 1):
 * codes = 2 bits per base, base i of the concatenation in bits 2*(i&3) of byte i>>2 (A C G T = 0 1 2 3); nmask = 1 bit per
 * base, bit i&7 of byte i>>3, set where the letter is outside ACGT (its code is ignored), or NULL when there is none.
 * offsets[n+1] and expr[n] as for rlsim_add_transcripts (offsets count bases, offsets[0] need not be 0 or a multiple of 4).
 * The library built from it is the one rlsim_add_transcripts builds from the same sequences with every letter outside
 * ACGT written as 'N': exact for inputs whose only such letter is N.  (The reference itself tells other letters apart
 * from N in one place only, the palindrome test of KmerAffinity, nnthermo.go:154-161: their doublets miss the energy
 * maps alike, nnthermo.go:138-142, and none counts as G/C, thermocycler.go:161-166.)  FASTA output shows 'N' for them.
 * Pinned caller arrays are copied directly; pageable ones go through the driver's staging. */
int rlsim_add_transcripts_packed(rlsim_handle *h, const uint8_t *codes, const uint8_t *nmask, const uint64_t *offsets,
                                 const uint64_t *expr, uint32_t n);
/* Device-side FASTA reader (next row 8f-3; fasta.go:113-212 ReadFasta/FastaToSeq, input.go:82-124 ParseSeqName and the
 * -m multiplier): `text` = the raw bytes of the input files, concatenated as io.MultiReader would (fasta.go:98-104).
 * Record splitting, despacing, upper-casing and `name$level` parsing run on the GPU; records with malformed names are
 * skipped as the reference does; the remaining records are added exactly as rlsim_add_transcripts would add them.
 * *n_records = records found, *n_added = records added.  RLSIM_E_ARG with the reference's message when the text does
 * not start with '>' ("Illegal fasta record"); RLSIM_E_UNSUPPORTED for a run of more than 64 consecutive '>'. */
int rlsim_add_fasta(rlsim_handle *h, const uint8_t *text, uint64_t n_bytes, double expr_mul, uint32_t *n_records,
                    uint32_t *n_added);
/* Record table of the last rlsim_add_fasta, one entry per record found (any pointer may be NULL): name_start = offset of
 * the name line in `text`, line_len = its length, name_len = bytes before the '$' (0 when malformed), level = expression
 * level with the multiplier applied, seq_len = sequence length, valid = 0 for skipped records ("Skipping transcript
 * with malformed name: <line>", input.go:83). */
int rlsim_get_fasta_index(rlsim_handle *h, uint64_t *name_start, uint32_t *line_len, uint32_t *name_len, uint64_t *level,
                          uint64_t *seq_len, uint8_t *valid);
int rlsim_set_first_index(rlsim_handle *h, uint32_t first_index);
/* forget the transcripts and everything derived from them; device buffers are kept for the next library */
int rlsim_clear_transcripts(rlsim_handle *h);
/* fragment every molecule (transcript.go:139-150 + fragmentor.go), collapse identical fragments into species
 * (transcript.go:156-215), amplify (thermocycler.go:106-147) and build the per-length pool (pool.go:156-189). */
int rlsim_build_library(rlsim_handle *h);
/* the two halves of build_library, for tests and profiling */
int rlsim_fragment(rlsim_handle *h);
int rlsim_pcr(rlsim_handle *h);

/* ---- Sampler.SampleFragments (sampler.go:108-154) ----
 * class_totals: this GPU's pooled copies per length, T_r[l], l = 0..n_len-1 (n_len = high+1).
 * sample: global totals and this rank's base offset per length (single GPU: totals = class_totals, base = NULL).
 * Multi-GPU: the caller all-gathers class_totals (the only data that crosses NVLink) and passes
 * totals[l] = sum_r T_r[l], base[l] = sum_{r' < rank} T_r'[l]. */
int rlsim_class_totals(rlsim_handle *h, uint64_t *totals, uint32_t n_len);
int rlsim_sample(rlsim_handle *h, const uint64_t *global_totals, const uint64_t *rank_base, uint32_t n_len);

/* Multi-GPU form without redundant work: the length classes are partitioned over the ranks
 * (class l is drawn by rank l % nparts).  all_totals = T_r[l] of every rank, row-major [nparts][n_len] (the all-gathered
 * class_totals).  rlsim_select_classes draws this rank's classes and writes the copy ranks to take (or to leave, for
 * classes drawn more than half) as 16-byte records {u32 len, u32 0, u64 rank_in_class} into `records_dev` - the caller's
 * DEVICE buffer of `cap` records - grouped by owner rank; counts_out[r] = records for rank r, *needed_out = their sum
 * (if it exceeds cap nothing is written: grow the buffer and call again).  The records are complete in device memory
 * when the call returns (the library synchronises its own stream).  The caller routes the groups with one
 * all-to-all, hands what it receives to rlsim_apply_selected, and finishes with rlsim_finish_sample. */
int rlsim_select_classes(rlsim_handle *h, const uint64_t *all_totals, uint32_t nparts, uint32_t part, uint32_t n_len,
                         void *records_dev, uint64_t cap, uint64_t *counts_out, uint64_t *needed_out);
int rlsim_apply_selected(rlsim_handle *h, const void *records_dev, uint64_t n);
int rlsim_finish_sample(rlsim_handle *h);

/* ---- the same exchange inside the library: one process per GPU, an NCCL communicator of the library's own, everything
 * queued on the handle's stream (no host framework between the steps).  libnccl.so.2 is bound at run time.
 * Rank 0 makes an id with rlsim_comm_unique_id, the host passes it to every rank by any means, every rank calls
 * rlsim_comm_init (collective).  rlsim_sample_distributed = Sampler.SampleFragments on all ranks: all-gather of
 * [per-length pool totals | partial target histogram], class-partitioned draw, one grouped send/recv that routes the drawn
 * copy ranks to their owners, apply, expand.  Afterwards RLSIM_H_POOL_GLOBAL holds the pool totals summed over the ranks.
 * rlsim_comm_sum_target sums the partial target histograms early (a raw -j target shapes the library).
 * rlsim_exchange_plan is the host arithmetic of the routing: counts[s*n+d] = records of sender s for rank d. */
#define RLSIM_COMM_ID_BYTES 128
int rlsim_comm_unique_id(uint8_t *id_out);
int rlsim_comm_init(rlsim_handle *h, const uint8_t *id, int nranks, int rank);
int rlsim_comm_sum_target(rlsim_handle *h);
int rlsim_sample_distributed(rlsim_handle *h);
int rlsim_exchange_plan(const uint64_t *counts, uint32_t n, uint32_t rank, uint64_t *send_off, uint64_t *recv_off);

/* ---- results ---- */
int rlsim_get_stats(rlsim_handle *h, rlsim_stats *out);
/* histogram `kind` into out[0..n) (n = number of bins wanted; bins beyond the data are 0) */
int rlsim_get_hist(rlsim_handle *h, int kind, uint64_t *out, uint32_t n);
/* species table sorted by (transcript, start, end); any pointer may be NULL */
int rlsim_get_species(rlsim_handle *h, uint32_t *tr, uint32_t *start, uint32_t *end, uint32_t *count_frag,
                      uint64_t *count_pcr, double *eff, uint32_t *gc);
/* sampled fragments (Frag records, frag.go:46-52) in output order; any pointer may be NULL */
int rlsim_get_sampled(rlsim_handle *h, uint32_t *tr, uint32_t *start, uint32_t *end, uint8_t *minus_strand,
                      uint64_t *id);
/* The same records packed for the host link: 6 bytes per fragment instead of 21.  The records are ordered by
 * (transcript, start, end), so tr_counts[t] (t = 0 .. n_transcripts-1, this shard's indices) = number of records of
 * transcript t replaces the per-record transcript index AND the id (ids count 0,1,.. within a transcript's run,
 * sampler.go:189,204).  start[i] as above; len_strand[i] = (end - start) | strand << (8*elem_bytes - 1), elem_bytes = 2
 * (needs high < 32768) or 4.  Any pointer may be NULL.  Pageable destinations are staged through pinned memory. */
int rlsim_get_sampled_packed(rlsim_handle *h, uint32_t *tr_counts, uint32_t *start, void *len_strand, int elem_bytes);
/* Frag.String for every sampled fragment (frag.go:71-126, sampler.go:206), formatted on the device:
 * ">Fg_<id>_<name> (Strand <s> Offset <start> -- <end>)\n<seq>\n".  names/name_off: transcript names of
 * this shard.  Call with buf = NULL to get the size. */
int rlsim_format_fasta(rlsim_handle *h, const uint8_t *names, const uint64_t *name_off, uint8_t *buf,
                       uint64_t cap, uint64_t *size_out);

/* ---- probes used by the parity tests (deterministic building blocks on the device) ---- */
int rlsim_probe_kmer_lut(rlsim_handle *h, double *K_out, uint32_t *w_out, uint32_t n);       /* nnthermo.go:126-175 */
int rlsim_probe_prefix(rlsim_handle *h, uint32_t tr, int reverse, uint64_t *out, uint32_t n); /* nnthermo.go:177-201 */
int rlsim_probe_gc_prefix(rlsim_handle *h, uint32_t tr, uint32_t *out, uint32_t n);
int rlsim_probe_math(rlsim_handle *h, int fn, const double *x, const double *y, double *out, uint32_t n);
int rlsim_probe_philox(rlsim_handle *h, const uint32_t *ctr_key6, uint32_t *out4, uint32_t n);
int rlsim_probe_amplify(rlsim_handle *h, int64_t seed_pcr, const uint32_t *tse, const uint64_t *n0, const double *p,
                        int cycles, uint64_t *out, uint32_t n);
/* device time of the kernels of the last build_library / sample call, milliseconds, by stage */
enum { RLSIM_T_LAYOUT = 0, RLSIM_T_PREFIX = 1, RLSIM_T_FRAGMENT = 2, RLSIM_T_DEDUP = 3, RLSIM_T_PCR = 4,
       RLSIM_T_SELECT = 5, RLSIM_T_EXPAND = 6, RLSIM_T_FORMAT = 7, RLSIM_T_TARGET = 8, RLSIM_T_INTERVALS = 9,
       RLSIM_T_EXCHANGE = 10, RLSIM_T_COUNT = 11 };
int rlsim_get_timings(rlsim_handle *h, float *ms_out, uint64_t *launches_out);

#ifdef __cplusplus
}
#endif
#endif

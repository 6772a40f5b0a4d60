/* oracle/oracle.h - public interface of the CPU oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rlsim_b200/ includes, links or
 * calls this; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs do, and only as the checker (see oracle/README.md).
 *
 * The oracle restates the reference's library-construction hot path
 *  in plain C, in two modes:
 *   ORC_MODE_REF    - the reference's own algorithms driven by a bit-exact
 *                     replica of Go's math/rand; pinned against the release
 *                     binary (tests/test_oracle_pin.py).
 *   ORC_MODE_PHILOX - the normative "--rng=philox" spec;
 *                     the CUDA path must reproduce it bit for bit.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#define ORC_MAX_COMPS 8
#define ORC_MODE_REF 0
#define ORC_MODE_PHILOX 1

/* fragmentation methods, fragmentor.go:39-60 */
enum { ORC_AFTER_PRIM = 0, ORC_AFTER_PRIM_DOUBLE = 1, ORC_AFTER_NOPRIM = 2, ORC_AFTER_NOPRIM_DOUBLE = 3,
       ORC_PRE_PRIM = 4, ORC_PRIM_JUMP = 5 };
/* mixture component types, target_mix.go:61-67 */
enum { ORC_COMP_N = 0, ORC_COMP_SN = 1, ORC_COMP_G = 2 };

/* Same memory layout as rlsim_mixcomp / rlsim_params in include/rlsim_gpu.h
 * (tests build one ctypes structure for both). */
typedef struct {
    int32_t type;
    int32_t reserved;
    double weight, location, scale, shape;
    uint64_t low, high;
} orc_mixcomp;

typedef struct {
    int32_t kmer_len;      /* -k  args.go:107 */
    int32_t frag_method;   /* -f  args.go:277-307 */
    int32_t frag_param;
    int32_t nr_cycles;     /* -c */
    double priming_temp;   /* -p */
    double frag_loss_prob; /* -flg */
    double fixed_eff;      /* -e */
    double strand_bias;    /* -b */
    int32_t gc_mode;       /* 0 = parametric (-eg), 1 = raw table (default / -j) */
    int32_t n_target;      /* components of -d ; 0 when the target is raw (-j) */
    int32_t n_polya;       /* components of -a */
    int32_t reserved;
    double gc_shape, gc_min, gc_max;
    double gc_raw[101];    /* already clamped as thermocycler.go:94-104 */
    double len_shape, len_min, len_max; /* -el */
    orc_mixcomp target[ORC_MAX_COMPS];
    orc_mixcomp polya[ORC_MAX_COMPS];
    int64_t seed_init, seed_pcr, seed_sampling; /* -si -sp -ss (0 = not given), main.go:72-81,129-149 */
} orc_params;

typedef struct orc_sim orc_sim;

#ifdef __cplusplus
extern "C" {
#endif
/* ---- deterministic building blocks ---- */
double orc_kmer_affinity(const char *kmer, int k, double T, int detmath);              /* nnthermo.go:126-175 */
void orc_binding_profile(const char *plus, uint32_t len, int k, double T, int reverse, /* nnthermo.go:177-201 */
                         int detmath, double *out);
void orc_revcomp(const char *in, uint32_t n, char *out);                               /* utils.go:37-59 */
uint32_t orc_gc_count(const char *plus, uint32_t start, uint32_t end);                 /* thermocycler.go:161-166 */
double orc_gc_eff(const orc_params *p, uint32_t gc_count, uint32_t len, int detmath);  /* thermocycler.go:158-180 */
double orc_gc_eff_fixed(const orc_params *p, double gc, int detmath);                  /* thermocycler.go:182-193 */
void orc_len_eff_table(const orc_params *p, uint64_t low, uint64_t high, int detmath, double *out); /* :83-92,149-156 */
void orc_quantized_weights(const char *plus, uint32_t len, int k, double T, int reverse, uint64_t *w_out); /* spec 4.3 */
/* detmath / philox probes for kernel-vs-oracle unit parity */
double orc_det_log(double x);
double orc_det_exp(double x);
double orc_det_ndtri(double p);
double orc_det_logfact(double k);
double orc_det_go_pow(double x, double y);
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_derive_key(int64_t seed, uint32_t tag, uint32_t out[2]);
uint64_t orc_px_binomial(int64_t seed_pcr, uint32_t t, uint32_t start, uint32_t end, uint64_t n, double p, int ncalls);
int32_t orc_px_poisson(int64_t seed, uint32_t t, uint64_t m, double mean);
/* Go math/rand probes */
void orc_go_rand_floats(int64_t seed, int n, double *out);

/* ---- whole-path simulation ---- */
orc_sim *orc_sim_new(const orc_params *p, int mode);
void orc_sim_free(orc_sim *s);
const char *orc_sim_error(const orc_sim *s);
/* raw (-j) target: lengths + normalised probabilities in file order, parse_raw.go:83-104 */
int orc_sim_set_raw_target(orc_sim *s, const uint32_t *len, const double *prob, int n);
/* a19: draw req_frags target lengths -> histogram (target.go:102-123 / raw_target.go:104-129) */
int orc_sim_sample_target(orc_sim *s, int64_t req_frags);
/* or impose an externally sampled histogram */
int orc_sim_set_target_hist(orc_sim *s, const uint32_t *len, const uint64_t *count, int n);
/* transcripts: concatenated upper-cased bases (no poly-A appended), offsets[n+1], expression levels */
int orc_sim_add_transcripts(orc_sim *s, const char *bases, const uint64_t *offsets, const uint64_t *expr, int n);
/* pool.go:83-135: fragment (+dedup) every transcript, then PCR */
int orc_sim_fragment(orc_sim *s);
int orc_sim_pcr(orc_sim *s);
/* sampler.go:65-154: size selection + final draw */
int orc_sim_sample(orc_sim *s);

/* results (valid until orc_sim_free) */
uint64_t orc_sim_low(const orc_sim *s);
uint64_t orc_sim_high(const orc_sim *s);
uint32_t orc_sim_polya_max(const orc_sim *s);
/* histograms indexed by length, size *n_out; kinds: */
enum { ORC_H_TARGET = 0, ORC_H_POLYA = 1, ORC_H_AFTER_FRAG = 2, ORC_H_AFTER_PCR = 3, ORC_H_SAMPLED = 4,
       ORC_H_MISSING = 5, ORC_H_AFTER_SAMPLING = 6 };
const uint64_t *orc_sim_hist(const orc_sim *s, int kind, uint32_t *n_out);
uint64_t orc_sim_n_species(const orc_sim *s);
/* species sorted by (transcript, start, end) */
void orc_sim_species(const orc_sim *s, uint32_t *tr, uint32_t *start, uint32_t *end, uint64_t *count0,
                     uint64_t *count_pcr, double *eff, uint32_t *gc);
uint64_t orc_sim_pool_total(const orc_sim *s);
uint64_t orc_sim_n_sampled(const orc_sim *s);
/* sampled fragments in output order */
void orc_sim_sampled(const orc_sim *s, uint32_t *tr, uint32_t *start, uint32_t *end, uint8_t *minus_strand,
                     uint64_t *id);
/* frag.go:71-126: FASTA text of the sampled fragments; returns bytes needed (call with buf=NULL first) */
uint64_t orc_sim_fasta(const orc_sim *s, const char *names, const uint64_t *name_off, char *buf, uint64_t cap);
#ifdef __cplusplus
}
#endif
#endif

/* oracle/oracle_common.h - shared state of the CPU oracle.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "detmath.h"
#include "go_rand.h"
#include "oracle.h"
#include "philox.h"

#define ORC_NHIST 7

typedef struct { uint32_t t, start, end; } orc_frag;

typedef struct {
    uint32_t t, start, end, gc;
    uint64_t count0, count; /* after fragmentation / current (after PCR, decremented by ref-mode sampling) */
    uint64_t taken;         /* philox-mode selection: copies drawn */
    double eff;
} orc_species;

typedef struct {
    uint32_t t, start, end;
    uint8_t minus;
    uint64_t id;
} orc_sampled;

struct orc_sim {
    orc_params p;
    int mode;
    char err[256];
    /* raw target */
    int n_raw;
    uint32_t *raw_len;
    double *raw_prob;
    double raw_location; /* RawTarget pseudo component Location, raw_target.go:76-90 */
    int64_t req_frags;
    uint64_t low, high;
    uint32_t polya_max;
    /* histograms indexed by length */
    uint64_t *h[ORC_NHIST];
    uint32_t hn[ORC_NHIST];
    /* transcripts */
    int ntr;
    char **plus, **minus; /* with polyAmax 'A's appended, transcript.go:98-117 */
    uint32_t *trlen;      /* length including the maximal poly-A tail */
    uint64_t *expr;
    /* ref-mode generators, main.go:72-81,129-149 */
    go_rand_t rg, rg_pcr, rg_global;
    int have_pcr_rg;
    /* fragments -> species */
    orc_frag *frags;
    uint64_t nfrags, cfrags;
    orc_species *sp;
    uint64_t nsp;
    uint64_t pool_total;
    /* sampled fragments */
    orc_sampled *out;
    uint64_t nout, cout_;
    /* reusable scratch for windowed cumulative sums, random.go:369-401 */
    double *csum;
    uint64_t ccsum;
};

static inline int orc_fail(orc_sim *s, const char *msg) {
    snprintf(s->err, sizeof s->err, "%s", msg);
    return 1;
}

static inline void orc_hist_add(orc_sim *s, int kind, uint32_t len, uint64_t c) {
    if (len >= s->hn[kind]) {
        uint32_t nn = s->hn[kind] ? s->hn[kind] : 256;
        while (nn <= len) nn *= 2;
        s->h[kind] = (uint64_t *)realloc(s->h[kind], (size_t)nn * 8);
        memset(s->h[kind] + s->hn[kind], 0, (size_t)(nn - s->hn[kind]) * 8);
        s->hn[kind] = nn;
    }
    s->h[kind][len] += c;
}

static inline void orc_push_frag(orc_sim *s, uint32_t t, uint32_t start, uint32_t end) {
    if (s->nfrags == s->cfrags) {
        s->cfrags = s->cfrags ? s->cfrags * 2 : 4096;
        s->frags = (orc_frag *)realloc(s->frags, s->cfrags * sizeof(orc_frag));
    }
    orc_frag f = {t, start, end};
    s->frags[s->nfrags++] = f;
}

static inline void orc_push_out(orc_sim *s, uint32_t t, uint32_t start, uint32_t end, int minus, uint64_t id) {
    if (s->nout == s->cout_) {
        s->cout_ = s->cout_ ? s->cout_ * 2 : 4096;
        s->out = (orc_sampled *)realloc(s->out, s->cout_ * sizeof(orc_sampled));
    }
    orc_sampled o = {t, start, end, (uint8_t)minus, id};
    s->out[s->nout++] = o;
}

/* nnthermo.c */
double orc_kmer_affinity_n(const char *kmer, int k, double T, int detmath);
/* shared helpers implemented in oracle_model.c */
void orc_mix_minmax(const orc_mixcomp *c, int n, uint64_t *mn, uint64_t *mx);
double orc_len_eff(const orc_params *p, uint64_t low, uint64_t high, uint32_t l, int detmath);
/* mode-specific stages */
int orc_ref_sample_target(orc_sim *s);
int orc_ref_fragment(orc_sim *s);
int orc_ref_pcr(orc_sim *s);
int orc_ref_sample(orc_sim *s);
int orc_px_sample_target(orc_sim *s);
int orc_px_fragment(orc_sim *s);
int orc_px_pcr(orc_sim *s);
int orc_px_sample(orc_sim *s);
#endif

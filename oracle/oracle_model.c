/* oracle/oracle_model.c - deterministic model pieces + simulation container.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Plain-C restatement of
 *   /root/reference/src/nnthermo.go:63-201   (SantaLucia NN affinities, binding profiles)
 *   /root/reference/src/utils.go:37-59       (RevCompDNA)
 *   /root/reference/src/thermocycler.go:83-104,149-193 (efficiency functions)
 *   /root/reference/src/transcript.go:98-117,156-215   (transcript construction, species dedup)
 *   /root/reference/src/frag.go:71-126       (fragment sequence + FASTA record)
 *   /root/reference/src/target.go:72-87      (getGlobalMinMax)
 */
#include "oracle_common.h"

/* ---- nnthermo.go:68-104 doublet tables, index = 4*code(a)+code(b), A C G T = 0 1 2 3 ---- */
static int base_code(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}
/*                         AA     AC     AG     AT     CA     CC     CG     CT     GA     GC    GG     GT     TA     TC     TG     TT */
static const double NN_DH[16] = {-7.6, -8.4, -7.8, -7.2, -8.5, -8.0, -10.6, -7.8, -8.2, -9.8, -8.0, -8.4, -7.2, -8.2, -8.5, -7.6};
static const double NN_DS[16] = {-21.3, -22.4, -21.0, -20.4, -22.7, -19.9, -27.2, -21.0, -22.2, -24.4, -19.9, -22.4, -21.3, -22.2, -22.7, -21.3};

/* nnthermo.go:126-175 */
double orc_kmer_affinity_n(const char *kmer, int kl, double T, int detmath) {
    double dH = +0.2, dS = -5.7; /* initiation, :108-109 */
    for (int i = 0; i <= kl - 2; i++) {
        int a = base_code(kmer[i]), b = base_code(kmer[i + 1]);
        if (a >= 0 && b >= 0) { /* Go map miss on any other doublet adds 0 */
            dH += NN_DH[4 * a + b];
            dS += NN_DS[4 * a + b];
        }
    }
    if (kmer[0] == 'A' || kmer[0] == 'T') { dH += 2.2; dS += 6.9; }
    if (kmer[kl - 1] == 'A' || kmer[kl - 1] == 'T') { dH += 2.2; dS += 6.9; }
    int sym = 1; /* string palindrome over j <= kl/2, :155-161 */
    for (int j = 0; j <= kl / 2; j++)
        if (kmer[j] != kmer[kl - j - 1]) { sym = 0; break; }
    if (sym) dS += -1.4;
    double e = dH - (T * dS) / 1000.0;
    double x = -e / (1.9872 * T);
    return detmath ? det_exp(x) : exp(x);
}
double orc_kmer_affinity(const char *kmer, int k, double T, int detmath) { return orc_kmer_affinity_n(kmer, k, T, detmath); }

/* utils.go:37-59 */
void orc_revcomp(const char *in, uint32_t n, char *out) {
    for (uint32_t i = 0; i < n; i++) {
        char o;
        switch (in[i]) { case 'A': o = 'T'; break; case 'T': o = 'A'; break; case 'G': o = 'C'; break; case 'C': o = 'G'; break; default: o = 'N'; }
        out[n - 1 - i] = o;
    }
}

/* nnthermo.go:177-201: profile over `plus` (reverse=0) or over RevComp(plus) (reverse=1); len-k entries. */
void orc_binding_profile(const char *plus, uint32_t len, int k, double T, int reverse, int detmath, double *out) {
    const char *seq = plus;
    char *tmp = NULL;
    if (reverse) { tmp = (char *)malloc(len + 1); orc_revcomp(plus, len, tmp); seq = tmp; }
    if (len > (uint32_t)k)
        for (uint32_t i = 0; i < len - (uint32_t)k; i++) out[i] = orc_kmer_affinity_n(seq + i, k, T, detmath);
    free(tmp);
}

/* spec 4.3: fixed-point priming weights.  scale = 2^26 / Kmax, Kmax = max K over the 4^k ACGT k-mers
 * (det_exp), w = floor(K * scale + 0.5) clamped to [1, 2^26-1] (64 of them sum to less than 2^32: the GPU keeps
 * u32 sums inside 64-entry blocks and u64 sums per block). */
static double px_kmax(int k, double T) {
    static int cache_k = -1;          /* one (k, T) at a time: the table scan is 4^k evaluations */
    static double cache_T = 0.0, cache_v = 0.0;
    if (k == cache_k && T == cache_T) return cache_v;
    char buf[32];
    double mx = 0.0;
    uint64_t n = (uint64_t)1 << (2 * k);
    for (uint64_t c = 0; c < n; c++) {
        for (int i = 0; i < k; i++) buf[i] = "ACGT"[(c >> (2 * (k - 1 - i))) & 3];
        double v = orc_kmer_affinity_n(buf, k, T, 1);
        if (v > mx) mx = v;
    }
    cache_k = k; cache_T = T; cache_v = mx;
    return mx;
}
static uint64_t px_quant(double K, double scale) {
    double x = K * scale + 0.5;
    if (!(x < 67108863.0)) return 67108863ULL;
    if (x < 1.0) return 1ULL;  /* never 0: a window of weak k-mers (low -p, runs of N) keeps a positive total */
    return (uint64_t)x;
}
void orc_quantized_weights(const char *plus, uint32_t len, int k, double T, int reverse, uint64_t *w) {
    if (len <= (uint32_t)k) return;
    double *prof = (double *)malloc((size_t)(len - k) * 8);
    orc_binding_profile(plus, len, k, T, reverse, 1, prof);
    double scale = 67108864.0 / px_kmax(k, T);
    for (uint32_t i = 0; i < len - (uint32_t)k; i++) w[i] = px_quant(prof[i], scale);
    free(prof);
}

/* thermocycler.go:161-166 */
uint32_t orc_gc_count(const char *plus, uint32_t start, uint32_t end) {
    uint32_t gc = 0;
    for (uint32_t i = start; i < end; i++) gc += (plus[i] == 'G' || plus[i] == 'C');
    return gc;
}
/* Go's math.Pow (pow.go) with the platform libm standing in for Go's Exp/Log: exact for integer exponents
 * (square-and-multiply), within a few ulp of the release binary for fractional ones. */
static double ref_go_pow(double x, double y) {
    if (y == 0.0 || x == 1.0) return 1.0;
    if (y == 1.0) return x;
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1.0 / sqrt(x);
    if (x != x || y != y) return NAN;
    if (x == 0.0) return y < 0.0 ? INFINITY : 0.0;
    double absy = fabs(y);
    int flip = y < 0.0;
    double yi = floor(absy), yf = absy - yi;
    if (yf != 0.0 && x < 0.0) return NAN;
    if (yi >= 9.223372036854775808e18) return exp(y * log(x));
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0.0) {
        if (yf > 0.5) { yf -= 1.0; yi += 1.0; }
        a1 = exp(yf * log(x));
    }
    int xe;
    double x1 = frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < 0.5) { x1 += x1; xe--; }
    }
    if (flip) { a1 = 1.0 / a1; ae = -ae; }
    return ldexp(a1, ae);
}
static double pw(double x, double y, int detmath) { return detmath ? det_go_pow(x, y) : ref_go_pow(x, y); }
/* thermocycler.go:167-180 with gc = count/len */
double orc_gc_eff(const orc_params *p, uint32_t gc_count, uint32_t len, int detmath) {
    double gc = (double)gc_count / (double)len;
    return orc_gc_eff_fixed(p, gc, detmath);
}
/* thermocycler.go:182-193 */
double orc_gc_eff_fixed(const orc_params *p, double gc, int detmath) {
    if (p->gc_mode == 1) return p->gc_raw[(int)(gc * 100.0)];
    return p->gc_min + (p->gc_max - p->gc_min) * pw(1 - pw(gc, p->gc_shape, detmath), p->gc_shape, detmath);
}
/* thermocycler.go:83-92,149-156 */
double orc_len_eff(const orc_params *p, uint64_t low, uint64_t high, uint32_t l, int detmath) {
    if (p->len_shape == 0.0) return 1.0;
    double alpha = pw((double)low, -p->len_shape, detmath);
    double beta = pw((double)high, -p->len_shape, detmath);
    double A = (p->len_max - p->len_min) / (alpha - beta);
    double B = p->len_min - A * beta;
    return A * pw((double)l, -p->len_shape, detmath) + B;
}
void orc_len_eff_table(const orc_params *p, uint64_t low, uint64_t high, int detmath, double *out) {
    for (uint64_t l = low; l <= high; l++) out[l - low] = orc_len_eff(p, low, high, (uint32_t)l, detmath);
}

/* target.go:72-87 */
void orc_mix_minmax(const orc_mixcomp *c, int n, uint64_t *mn, uint64_t *mx) {
    uint64_t min = 0, max = 0;
    for (int i = 0; i < n; i++) {
        if (i == 0) min = c[i].low;
        if (max < c[i].high) max = c[i].high;
        if (min > c[i].low) min = c[i].low;
    }
    *mn = min; *mx = max;
}

/* ---- probes ---- */
double orc_det_log(double x) { return det_log(x); }
double orc_det_exp(double x) { return det_exp(x); }
double orc_det_ndtri(double p) { return det_ndtri(p); }
double orc_det_logfact(double k) { return det_logfact(k); }
double orc_det_go_pow(double x, double y) { return det_go_pow(x, y); }
void orc_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    px_block b = px_philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    memcpy(out, b.v, 16);
}
void orc_derive_key(int64_t seed, uint32_t tag, uint32_t out[2]) {
    px_key k = px_derive_key(seed, tag);
    out[0] = k.k0; out[1] = k.k1;
}
void orc_go_rand_floats(int64_t seed, int n, double *out) {
    go_rand_t g;
    go_rand_seed(&g, seed);
    for (int i = 0; i < n; i++) out[i] = go_float64(&g);
}

/* ---- simulation container ---- */
orc_sim *orc_sim_new(const orc_params *p, int mode) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof *s);
    s->p = *p;
    s->mode = mode;
    uint64_t mn;
    uint64_t mx;
    orc_mix_minmax(p->polya, p->n_polya, &mn, &mx); /* pool.go:84 */
    s->polya_max = (uint32_t)mx;
    if (p->n_target > 0) orc_mix_minmax(p->target, p->n_target, &s->low, &s->high); /* target.go:67 */
    /* main.go:72-81,129-136 (time-based seeding is not reproducible: seed 0 -> 1 here) */
    go_rand_seed(&s->rg, p->seed_init ? p->seed_init : 1);
    go_rand_seed(&s->rg_global, 1); /* package-global math/rand source, used by random.go:184,205,220 */
    if (p->seed_pcr) { go_rand_seed(&s->rg_pcr, p->seed_pcr); s->have_pcr_rg = 1; }
    return s;
}
void orc_sim_free(orc_sim *s) {
    if (!s) return;
    for (int i = 0; i < s->ntr; i++) { free(s->plus[i]); free(s->minus[i]); }
    free(s->plus); free(s->minus); free(s->trlen); free(s->expr);
    for (int i = 0; i < ORC_NHIST; i++) free(s->h[i]);
    free(s->raw_len); free(s->raw_prob); free(s->frags); free(s->sp); free(s->out); free(s->csum);
    free(s);
}
const char *orc_sim_error(const orc_sim *s) { return s->err; }

int orc_sim_set_raw_target(orc_sim *s, const uint32_t *len, const double *prob, int n) {
    s->n_raw = n;
    s->raw_len = (uint32_t *)malloc((size_t)n * 4);
    s->raw_prob = (double *)malloc((size_t)n * 8);
    memcpy(s->raw_len, len, (size_t)n * 4);
    memcpy(s->raw_prob, prob, (size_t)n * 8);
    return 0;
}

/* raw_target.go:54,59-90: Low/High = min/max of the *sampled* lengths, pseudo component mean */
static void raw_target_finish(orc_sim *s) {
    uint32_t mn = 0, mx = 0;
    int first = 1;
    double mean = 0.0, t = (double)s->req_frags;
    for (uint32_t l = 0; l < s->hn[ORC_H_TARGET]; l++) {
        uint64_t c = s->h[ORC_H_TARGET][l];
        if (!c) continue;
        if (first) { mn = mx = l; first = 0; }
        if (l < mn) mn = l;
        if (l > mx) mx = l;
    }
    /* raw_target.go:82-85 accumulates in Go map order; ascending order is used here (the value only
     * enters the Poisson rate, fragmentor.go:108) */
    for (uint32_t l = 0; l < s->hn[ORC_H_TARGET]; l++) {
        uint64_t c = s->h[ORC_H_TARGET][l];
        if (c) mean += ((double)c / t) * (double)l;
    }
    s->low = mn; s->high = mx; s->raw_location = mean;
}

int orc_sim_sample_target(orc_sim *s, int64_t req_frags) {
    s->req_frags = req_frags;
    int rc = s->mode == ORC_MODE_REF ? orc_ref_sample_target(s) : orc_px_sample_target(s);
    if (rc) return rc;
    if (s->n_raw) raw_target_finish(s);
    return 0;
}
int orc_sim_set_target_hist(orc_sim *s, const uint32_t *len, const uint64_t *count, int n) {
    s->req_frags = 0;
    for (int i = 0; i < n; i++) { orc_hist_add(s, ORC_H_TARGET, len[i], count[i]); s->req_frags += (int64_t)count[i]; }
    if (s->n_raw || s->p.n_target == 0) raw_target_finish(s);
    return 0;
}

/* transcript.go:98-117 */
int orc_sim_add_transcripts(orc_sim *s, const char *bases, const uint64_t *off, const uint64_t *expr, int n) {
    int base = s->ntr;
    s->ntr += n;
    s->plus = (char **)realloc(s->plus, sizeof(char *) * s->ntr);
    s->minus = (char **)realloc(s->minus, sizeof(char *) * s->ntr);
    s->trlen = (uint32_t *)realloc(s->trlen, 4 * (size_t)s->ntr);
    s->expr = (uint64_t *)realloc(s->expr, 8 * (size_t)s->ntr);
    for (int i = 0; i < n; i++) {
        uint32_t l = (uint32_t)(off[i + 1] - off[i]);
        uint32_t L = l + s->polya_max;
        char *pl = (char *)malloc((size_t)L + 1);
        memcpy(pl, bases + off[i], l);
        memset(pl + l, 'A', s->polya_max);
        pl[L] = 0;
        char *mi = (char *)malloc((size_t)L + 1);
        orc_revcomp(pl, L, mi);
        mi[L] = 0;
        s->plus[base + i] = pl; s->minus[base + i] = mi; s->trlen[base + i] = L; s->expr[base + i] = expr[i];
    }
    return 0;
}

static int frag_cmp(const void *a, const void *b) {
    const orc_frag *x = (const orc_frag *)a, *y = (const orc_frag *)b;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return 0;
}

/* transcript.go:156-170,189-215: identical (length,start,end) of one transcript collapse into a species with a count */
static void dedup(orc_sim *s) {
    qsort(s->frags, s->nfrags, sizeof(orc_frag), frag_cmp);
    free(s->sp);
    s->sp = (orc_species *)calloc(s->nfrags ? s->nfrags : 1, sizeof(orc_species));
    s->nsp = 0;
    for (uint64_t i = 0; i < s->nfrags; i++) {
        orc_frag *f = &s->frags[i];
        if (s->nsp && s->sp[s->nsp - 1].t == f->t && s->sp[s->nsp - 1].start == f->start && s->sp[s->nsp - 1].end == f->end) {
            s->sp[s->nsp - 1].count0++;
            s->sp[s->nsp - 1].count++;
            continue;
        }
        orc_species *q = &s->sp[s->nsp++];
        q->t = f->t; q->start = f->start; q->end = f->end; q->count0 = q->count = 1;
    }
}

int orc_sim_fragment(orc_sim *s) {
    if (s->hn[ORC_H_TARGET] == 0 && s->p.n_target == 0) return orc_fail(s, "raw target needs a target histogram before fragmentation");
    s->nfrags = 0;
    int rc = s->mode == ORC_MODE_REF ? orc_ref_fragment(s) : orc_px_fragment(s);
    if (rc) return rc;
    dedup(s);
    return 0;
}
int orc_sim_pcr(orc_sim *s) {
    int rc = s->mode == ORC_MODE_REF ? orc_ref_pcr(s) : orc_px_pcr(s);
    if (rc) return rc;
    /* thermocycler.go:125-129 + fragstats.go:84-95 */
    s->pool_total = 0;
    for (uint64_t i = 0; i < s->nsp; i++) {
        uint32_t len = s->sp[i].end - s->sp[i].start;
        orc_hist_add(s, ORC_H_AFTER_PCR, len, s->sp[i].count);
        orc_hist_add(s, ORC_H_AFTER_SAMPLING, len, s->sp[i].count);
        uint64_t nt = s->pool_total + s->sp[i].count;
        if (nt < s->pool_total) return orc_fail(s, "Integer overflow when updating total fragment count!");
        s->pool_total = nt;
    }
    return 0;
}
int orc_sim_sample(orc_sim *s) {
    s->nout = 0;
    return s->mode == ORC_MODE_REF ? orc_ref_sample(s) : orc_px_sample(s);
}

uint64_t orc_sim_low(const orc_sim *s) { return s->low; }
uint64_t orc_sim_high(const orc_sim *s) { return s->high; }
uint32_t orc_sim_polya_max(const orc_sim *s) { return s->polya_max; }
const uint64_t *orc_sim_hist(const orc_sim *s, int kind, uint32_t *n_out) { *n_out = s->hn[kind]; return s->h[kind]; }
uint64_t orc_sim_n_species(const orc_sim *s) { return s->nsp; }
void orc_sim_species(const orc_sim *s, uint32_t *tr, uint32_t *start, uint32_t *end, uint64_t *count0, uint64_t *count_pcr,
                     double *eff, uint32_t *gc) {
    for (uint64_t i = 0; i < s->nsp; i++) {
        tr[i] = s->sp[i].t; start[i] = s->sp[i].start; end[i] = s->sp[i].end;
        count0[i] = s->sp[i].count0; count_pcr[i] = s->sp[i].count; eff[i] = s->sp[i].eff; gc[i] = s->sp[i].gc;
    }
}
uint64_t orc_sim_pool_total(const orc_sim *s) { return s->pool_total; }
uint64_t orc_sim_n_sampled(const orc_sim *s) { return s->nout; }
void orc_sim_sampled(const orc_sim *s, uint32_t *tr, uint32_t *start, uint32_t *end, uint8_t *minus, uint64_t *id) {
    for (uint64_t i = 0; i < s->nout; i++) {
        tr[i] = s->out[i].t; start[i] = s->out[i].start; end[i] = s->out[i].end; minus[i] = s->out[i].minus; id[i] = s->out[i].id;
    }
}

/* frag.go:71-89,123-126 + sampler.go:206: ">Fg_<id>_<name> (Strand <s> Offset <start> -- <end>)\n<seq>\n" */
uint64_t orc_sim_fasta(const orc_sim *s, const char *names, const uint64_t *name_off, char *buf, uint64_t cap) {
    uint64_t pos = 0;
    char hdr[1024];
    for (uint64_t i = 0; i < s->nout; i++) {
        const orc_sampled *o = &s->out[i];
        int nl = (int)(name_off[o->t + 1] - name_off[o->t]);
        int hl = snprintf(hdr, sizeof hdr, ">Fg_%llu_%.*s (Strand %c Offset %u -- %u)\n", (unsigned long long)o->id, nl,
                          names + name_off[o->t], o->minus ? '-' : '+', o->start, o->end);
        uint32_t sl = o->end - o->start;
        if (buf && pos + hl + sl + 1 <= cap) {
            memcpy(buf + pos, hdr, hl);
            uint32_t final = s->trlen[o->t] - 1;
            if (!o->minus) memcpy(buf + pos + hl, s->plus[o->t] + o->start, sl);
            else memcpy(buf + pos + hl, s->minus[o->t] + (final - o->end + 1), sl); /* frag.go:84-86 */
            buf[pos + hl + sl] = '\n';
        }
        pos += (uint64_t)hl + sl + 1;
    }
    return pos;
}

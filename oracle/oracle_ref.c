/* oracle/oracle_ref.c - ORC_MODE_REF: the reference's own stochastic algorithms, driven by the Go
 * math/rand replica, in the reference's draw order.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Restates
 *   /root/reference/src/random.go:120-475      (Poisson, Binomial, truncated variates, weighted index)
 *   /root/reference/src/target.go:97-123, raw_target.go:104-138, target_mix.go:55-72,232-250
 *   /root/reference/src/transcript.go:139-150,279-284 (per-molecule loop, poly-A)
 *   /root/reference/src/fragmentor.go:95-351   (after_*, pre_prim, prim_jump)
 *   /root/reference/src/fragfilter.go:45-50
 *   /root/reference/src/thermocycler.go:106-147
 *   /root/reference/src/pool.go:156-204, sampler.go:65-227, transcript.go:172-187
 *
 * What is reproducible against the release binary with a fixed -si (and is pinned by
 * tests/test_oracle_pin.py): target lengths, poly-A tail lengths, fragments after fragmentation.
 * PCR and sampling are order-dependent in the reference (Go map iteration, goroutines), so they are
 * compared statistically.
 */
#include "oracle_common.h"

/* ---------- random.go variates over a go_rand_t ---------- */
static double go_lgamma(double x) { return lgamma(x); }

/* random.go:120-163 */
static int32_t ref_poisson(go_rand_t *rg, double xm) {
    double sq, alxm, g, em, t, y;
    if (xm < 12.0) {
        g = exp(-xm);
        em = -1;
        t = 1.0;
        for (;;) {
            em++;
            t *= go_float64(rg);
            if (!(t > g)) return (int32_t)em;
        }
    }
    sq = sqrt(2.0 * xm);
    alxm = log(xm);
    g = xm * alxm - go_lgamma(xm + 1.0);
    for (;;) {
        for (;;) {
            y = tan(M_PI * go_float64(rg));
            em = sq * y + xm;
            if (!(em < 0.0)) break;
        }
        em = floor(em);
        t = 0.9 * (1.0 + y * y) * exp(em * alxm - go_lgamma(em + 1.0) - g);
        if (!(go_float64(rg) > t)) return (int32_t)em;
    }
}

/* random.go:168-233; `gl` is the package-global source (random.go:184,205,220) */
static uint64_t ref_binomial(go_rand_t *gl, uint64_t n, double pp) {
    double p = pp;
    if (pp <= 0.5) p = 1.0 - pp;
    double am = (double)n * p;
    uint64_t bnl = 0;
    if (n < 25) {
        for (uint64_t j = 1; j <= n; j++)
            if (go_float64(gl) < p) bnl++;
    } else {
        double en = (double)n;
        double g = go_lgamma(en + 1.0);
        double pc = 1.0 - p;
        double plog = log(p), pclog = log(pc);
        double sq = sqrt(2.0 * am * pc);
        double em, y, t;
        for (;;) {
            for (;;) {
                double angle = M_PI * go_float64(gl);
                y = tan(angle);
                em = sq * y + am;
                if (em < 0.0 || em >= (en + 1.0)) continue;
                break;
            }
            em = floor(em);
            t = 1.2 * sq * (1.0 + y * y) * exp(g - go_lgamma(em + 1.0) - go_lgamma(en - em + 1.0) + em * plog + (en - em) * pclog);
            if (go_float64(gl) > t) continue; /* NaN t (p in {0,1}) compares false -> accept, SURVEY a15 */
            break;
        }
        bnl = (uint64_t)em;
    }
    if (p != pp) bnl = n - bnl;
    return bnl;
}

/* Go's float64 -> uint64 conversion on amd64 for |x| < 2^63: truncate as int64, reinterpret */
static uint64_t go_f2u64(double x) { return (uint64_t)(int64_t)x; }

/* random.go:235-247 */
static uint32_t ref_trunc_norm(go_rand_t *rg, double mean, double sd, uint64_t low, uint64_t high) {
    double h = (double)high, l = (double)low;
    double uf = go_norm_float64(rg) * sd + mean;
    while (uf < l || uf > h) uf = go_norm_float64(rg) * sd + mean;
    return (uint32_t)uf;
}
/* random.go:250-276 */
static uint64_t ref_snorm_u64(go_rand_t *rg, double loc, double scale, double shape) {
    double u1 = go_norm_float64(rg);
    double u2 = go_norm_float64(rg);
    if (u2 > shape * u1) u1 = -u1;
    return go_f2u64(loc + scale * u1);
}
static uint32_t ref_trunc_snorm(go_rand_t *rg, double loc, double scale, double shape, uint64_t low, uint64_t high) {
    uint64_t u = ref_snorm_u64(rg, loc, scale, shape);
    while (u < low || u > high) u = ref_snorm_u64(rg, loc, scale, shape);
    return (uint32_t)u;
}
/* random.go:282-303 (Knuth) */
static double ref_rgamma1(go_rand_t *rg, double shape) {
    double shm1 = shape - 1.0, sh2m1 = sqrt(2.0 * shape - 1.0);
    for (;;) {
        double y = tan(M_PI * go_float64(rg));
        double x = sh2m1 * y + shm1;
        if (x <= 0.0) continue;
        double v = go_float64(rg);
        if (v > (1.0 + y * y) * exp(shm1 * log(x / shm1) - sh2m1 * y)) continue;
        return x;
    }
}
/* random.go:308-342 (Kundu & Gupta 2007) */
static double ref_rgamma2(go_rand_t *rg, double shape) {
    double d = 1.0334 - 0.0766 * exp(2.2942 * shape);
    double a = exp2(shape) * pow(-expm1(-d / 2), shape);
    double pdsh = pow(d, shape - 1.0);
    double b = shape * pdsh * exp(-d);
    double c = a + b;
    for (;;) {
        double u = go_float64(rg), x;
        if (u <= a / c) x = -2.0 * log1p(-pow(c * u, 1.0 / shape) / 2.0);
        else x = -log(c * (1.0 - u) / (shape * pdsh));
        double v = go_float64(rg);
        if (x <= d) {
            double p = pow(x, shape - 1.0) * exp(-x / 2.0) / (exp2(shape - 1.0) * pow(-expm1(-x / 2.0), shape - 1.0));
            if (v > p) continue;
        } else if (v > pow(d / x, 1.0 - shape)) continue;
        return x;
    }
}
/* random.go:345-367 */
static double ref_rgamma(go_rand_t *rg, double shape, double loc) {
    if (shape == 1.0) return (loc / shape) * go_exp_float64(rg);
    if (shape > 1.0) return (loc / shape) * ref_rgamma1(rg, shape);
    return (loc / shape) * ref_rgamma2(rg, shape);
}
static uint32_t ref_trunc_gamma(go_rand_t *rg, double loc, double shape, uint64_t low, uint64_t high) {
    double l = (double)low, h = (double)high;
    double u = ref_rgamma(rg, shape, loc);
    while (u < l || u > h) u = ref_rgamma(rg, shape, loc);
    return (uint32_t)u;
}

/* random.go:369-437: sequential cumulative sum + binary search for the first csum > u */
static int ref_sample_index_f64(orc_sim *s, go_rand_t *rg, const double *p, uint64_t n, uint64_t *ind) {
    if (n > s->ccsum) { s->ccsum = n * 2; s->csum = (double *)realloc(s->csum, s->ccsum * 8); }
    double *cs = s->csum;
    cs[0] = p[0];
    for (uint64_t i = 1; i < n; i++) cs[i] = cs[i - 1] + p[i];
    double high = cs[n - 1];
    if (high == 0.0) return 0;
    double u = go_float64(rg) * high;
    uint64_t i = 0, j = n;
    while (i < j) {
        uint64_t h = i + (j - i) / 2;
        if (!(cs[h] > u)) i = h + 1; else j = h;
    }
    *ind = i;
    return 1;
}
/* random.go:439-475; returns 0 when the weights sum to zero */
static int ref_sample_index_u64(go_rand_t *rg, const uint64_t *p, uint64_t n, uint64_t *scratch, uint64_t *ind, int *ovf) {
    scratch[0] = p[0];
    for (uint64_t i = 1; i < n; i++) {
        scratch[i] = scratch[i - 1] + p[i];
        if (scratch[i] < scratch[i - 1]) { *ovf = 1; return 0; }
    }
    uint64_t high = scratch[n - 1];
    if (high == 0) return 0;
    uint64_t u = (uint64_t)go_int63n(rg, (int64_t)high);
    uint64_t i = 0, j = n;
    while (i < j) {
        uint64_t h = i + (j - i) / 2;
        if (!(scratch[h] > u)) i = h + 1; else j = h;
    }
    *ind = i;
    return 1;
}

/* target_mix.go:232-245 (component order = caller's order; Go map order in the reference) */
static const orc_mixcomp *ref_mix_comp(orc_sim *s, go_rand_t *rg, const orc_mixcomp *c, int n) {
    double p[ORC_MAX_COMPS];
    for (int i = 0; i < n; i++) p[i] = c[i].weight;
    uint64_t idx = 0;
    ref_sample_index_f64(s, rg, p, (uint64_t)n, &idx);
    return &c[idx];
}
/* target_mix.go:55-72 */
static uint32_t ref_comp_len(go_rand_t *rg, const orc_mixcomp *c) {
    if (c->low == c->high) return (uint32_t)c->low;
    switch (c->type) {
    case ORC_COMP_N: return ref_trunc_norm(rg, c->location, c->scale, c->low, c->high);
    case ORC_COMP_SN: return ref_trunc_snorm(rg, c->location, c->scale, c->shape, c->low, c->high);
    default: return ref_trunc_gamma(rg, c->location, c->scale, c->low, c->high);
    }
}

/* target.go:102-123 / raw_target.go:104-129 */
int orc_ref_sample_target(orc_sim *s) {
    for (int64_t i = 0; i < s->req_frags; i++) {
        uint32_t l;
        if (s->n_raw) {
            uint64_t pos;
            if (!ref_sample_index_f64(s, &s->rg, s->raw_prob, (uint64_t)s->n_raw, &pos))
                return orc_fail(s, "Error when sampling from raw fragment size distribution!");
            l = s->raw_len[pos];
        } else {
            const orc_mixcomp *c = ref_mix_comp(s, &s->rg, s->p.target, s->p.n_target);
            l = ref_comp_len(&s->rg, c);
        }
        orc_hist_add(s, ORC_H_TARGET, l, 1);
    }
    return 0;
}

/* nnthermo.go:203-217 */
static int ref_priming(orc_sim *s, go_rand_t *rg, const double *prof, uint32_t proflen, uint32_t start, uint32_t end, uint32_t *res) {
    if (end > proflen) end = proflen;
    if (start >= end) { *res = 0; return 0; }
    uint64_t ind;
    if (!ref_sample_index_f64(s, rg, prof + start, end - start, &ind)) return 1;
    *res = (uint32_t)ind + start;
    return 0;
}

static int int_cmp(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : x > y;
}

typedef struct {
    orc_sim *s;
    uint32_t t;
    const double *fwd, *rev;
    uint32_t proflen;
    int64_t *breaks;
    uint64_t cbreaks;
} ref_ctx;

static int64_t *ref_breaks(ref_ctx *c, uint64_t n) {
    if (n > c->cbreaks) { c->cbreaks = n * 2; c->breaks = (int64_t *)realloc(c->breaks, c->cbreaks * 8); }
    return c->breaks;
}

/* target.go:97-100 / raw_target.go:100-102: Location of the sampled component */
static double ref_target_location(orc_sim *s, go_rand_t *rg) {
    if (s->n_raw || s->p.n_target == 0) return s->raw_location;
    return ref_mix_comp(s, rg, s->p.target, s->p.n_target)->location;
}
/* fragfilter.go:45-50 */
static int ref_keep(orc_sim *s, go_rand_t *rg) { return !(go_float64(rg) < s->p.frag_loss_prob); }

static void ref_register(orc_sim *s, uint32_t t, uint32_t size, uint32_t start, uint32_t end) {
    orc_push_frag(s, t, start, end);          /* transcript.go:156-170 */
    orc_hist_add(s, ORC_H_AFTER_FRAG, size, 1); /* fragstats.go:80 */
}

/* fragmentor.go:95-176 */
static int ref_after_prim(ref_ctx *c, go_rand_t *rg, int polyAend, int simPriming, int doublePrime) {
    orc_sim *s = c->s;
    double loc = ref_target_location(s, rg);
    int length = polyAend;
    uint32_t low = (uint32_t)s->low, high = (uint32_t)s->high;
    double rate;
    if (s->p.frag_param == 0) rate = ((double)length / loc) / 2.0;
    else rate = (double)length / (double)(uint32_t)s->p.frag_param;
    int32_t nb = ref_poisson(rg, rate) - 1;
    if (nb <= 0) nb = 1;
    int64_t *br = ref_breaks(c, (uint64_t)nb + 2);
    br[0] = 0;
    br[1] = length;
    for (int32_t i = 0; i < nb; i++) br[i + 2] = go_int63n(rg, (int64_t)length);
    qsort(br, (size_t)nb + 2, 8, int_cmp);
    for (int i = 0; i < nb; i++) { /* i < len(breaks)-2: the last interval is never visited */
        uint32_t start = (uint32_t)br[i], end = (uint32_t)br[i + 1], new_start;
        if (end - start < low) continue;
        if (simPriming) {
            if (ref_priming(s, rg, c->fwd, c->proflen, start, end, &new_start)) return orc_fail(s, "Error when simulating priming!");
            if (doublePrime) {
                uint32_t final = s->trlen[c->t] - 1;
                uint32_t ss = final - end + 1, ee = final - new_start + 1, ns;
                if (ref_priming(s, rg, c->rev, c->proflen, ss, ee, &ns)) return orc_fail(s, "Error when simulating priming!");
                end = final - ns + 1;
            }
        } else {
            new_start = start + (uint32_t)go_int31n(rg, (int32_t)(end - start));
            if (doublePrime) end = new_start + (uint32_t)go_int31n(rg, (int32_t)(end - new_start));
        }
        uint32_t size = end - new_start;
        if (size < low || size > high) continue;
        if (!ref_keep(s, rg)) continue;
        ref_register(s, c->t, size, new_start, end);
    }
    return 0;
}

/* fragmentor.go:205-270 */
static int ref_pre_prim(ref_ctx *c, go_rand_t *rg, int polyAend) {
    orc_sim *s = c->s;
    double loc = ref_target_location(s, rg);
    int length = polyAend;
    uint32_t low = (uint32_t)s->low, high = (uint32_t)s->high;
    double fp = 1.0 / (double)(uint32_t)s->p.frag_param;
    int elong = (int)(go_exp_float64(rg) / fp);
    if (elong > length) elong = length;
    elong = length - elong;
    double rate = (double)length / loc;
    int32_t nb = ref_poisson(rg, rate) - 1;
    if (nb <= 0) nb = 1;
    int64_t *br = ref_breaks(c, (uint64_t)nb + 2);
    br[0] = elong;
    br[1] = length;
    for (int32_t i = 0; i < nb; i++) {
        int64_t len_tmp = (int64_t)length - (int64_t)elong;
        if (len_tmp < 1) return 0;
        br[i + 2] = elong + go_int63n(rg, len_tmp);
    }
    qsort(br, (size_t)nb + 2, 8, int_cmp);
    for (int i = 0; i < nb; i++) {
        uint32_t start = (uint32_t)br[i], end = (uint32_t)br[i + 1];
        uint32_t size = end - start;
        if (size < low || size > high) continue;
        if (!ref_keep(s, rg)) continue;
        ref_register(s, c->t, size, start, end);
    }
    return 0;
}

/* target.go:141-144 / raw_target.go:148-154 */
static int ref_target_mixlen(orc_sim *s, go_rand_t *rg, uint32_t *l) {
    if (s->n_raw) {
        uint64_t pos;
        if (!ref_sample_index_f64(s, rg, s->raw_prob, (uint64_t)s->n_raw, &pos)) return 1;
        *l = s->raw_len[pos];
        return 0;
    }
    *l = ref_comp_len(rg, ref_mix_comp(s, rg, s->p.target, s->p.n_target));
    return 0;
}

/* fragmentor.go:303-351 */
static int ref_prim_jump(ref_ctx *c, go_rand_t *rg, int polyAend) {
    orc_sim *s = c->s;
    uint32_t start = 0, end = 0, new_end = 0;
    int ltmp = (int)s->trlen[c->t] - (int)c->proflen;
    uint32_t final = (uint32_t)(polyAend - ltmp);
    uint32_t low = (uint32_t)s->low, high = (uint32_t)s->high;
    double fp = s->p.frag_param != 0 ? 1.0 / (double)(uint32_t)s->p.frag_param : 0.0;
    while (end < final) {
        if (ref_priming(s, rg, c->fwd, c->proflen, end, final, &start)) return orc_fail(s, "Error when simulating priming!");
        if (fp != 0.0) new_end = start + (uint32_t)(go_exp_float64(rg) / fp);
        else {
            uint32_t l;
            if (ref_target_mixlen(s, rg, &l)) return orc_fail(s, "Error when sampling from raw fragment size distribution!");
            new_end = start + l;
        }
        if (new_end > final) new_end = final;
        end = new_end;
        uint32_t size = end - start;
        if (size < low || size > high) continue;
        if (!ref_keep(s, rg)) continue;
        ref_register(s, c->t, size, start, end);
    }
    return 0;
}

/* pool.go:101-124 + transcript.go:139-150 */
int orc_ref_fragment(orc_sim *s) {
    int k = s->p.kmer_len, m = s->p.frag_method;
    int need_fwd = (m == ORC_AFTER_PRIM || m == ORC_AFTER_PRIM_DOUBLE || m == ORC_PRIM_JUMP);
    int need_rev = (m == ORC_AFTER_PRIM_DOUBLE);
    ref_ctx c;
    memset(&c, 0, sizeof c);
    c.s = s;
    int rc = 0;
    for (int t = 0; t < s->ntr && !rc; t++) {
        if (s->expr[t] == 0) continue; /* pool.go:104 */
        uint32_t L = s->trlen[t];
        uint32_t proflen = L - (uint32_t)k;
        double *fwd = NULL, *rev = NULL;
        if (need_fwd) { fwd = (double *)malloc((size_t)proflen * 8 + 8); orc_binding_profile(s->plus[t], L, k, s->p.priming_temp, 0, 0, fwd); }
        if (need_rev) { rev = (double *)malloc((size_t)proflen * 8 + 8); orc_binding_profile(s->plus[t], L, k, s->p.priming_temp, 1, 0, rev); }
        c.t = (uint32_t)t; c.fwd = fwd; c.rev = rev; c.proflen = proflen;
        for (uint64_t mol = 0; mol < s->expr[t] && !rc; mol++) {
            /* transcript.go:279-284 */
            const orc_mixcomp *pc = ref_mix_comp(s, &s->rg, s->p.polya, s->p.n_polya);
            uint32_t tail = ref_comp_len(&s->rg, pc);
            orc_hist_add(s, ORC_H_POLYA, tail, 1);
            int polyAend = (int)L - (int)s->polya_max + (int)tail;
            switch (m) {
            case ORC_AFTER_PRIM: rc = ref_after_prim(&c, &s->rg, polyAend, 1, 0); break;
            case ORC_AFTER_PRIM_DOUBLE: rc = ref_after_prim(&c, &s->rg, polyAend, 1, 1); break;
            case ORC_AFTER_NOPRIM: rc = ref_after_prim(&c, &s->rg, polyAend, 0, 0); break;
            case ORC_AFTER_NOPRIM_DOUBLE: rc = ref_after_prim(&c, &s->rg, polyAend, 0, 1); break;
            case ORC_PRE_PRIM: rc = ref_pre_prim(&c, &s->rg, polyAend); break;
            default: rc = ref_prim_jump(&c, &s->rg, polyAend); break;
            }
        }
        free(fwd); free(rev);
    }
    free(c.breaks);
    return rc;
}

/* thermocycler.go:106-147.  Species order here is (transcript, start, end); the reference iterates Go
 * maps, so its binomial draw order is not reproducible. */
int orc_ref_pcr(orc_sim *s) {
    for (uint64_t i = 0; i < s->nsp; i++) {
        orc_species *q = &s->sp[i];
        uint32_t len = q->end - q->start;
        q->gc = orc_gc_count(s->plus[q->t], q->start, q->end);
        double e = s->p.fixed_eff;
        if (e == 0.0) e = orc_len_eff(&s->p, s->low, s->high, len, 0) * orc_gc_eff(&s->p, q->gc, len, 0);
        q->eff = e;
        uint64_t c = q->count;
        for (int cyc = 0; cyc < s->p.nr_cycles; cyc++) {
            uint64_t old = c;
            c += ref_binomial(&s->rg_global, c, e);
            if (c < old) return orc_fail(s, "Integer overflow detected when amplifying fragments!");
        }
        q->count = c;
    }
    return 0;
}

static uint64_t *g_sort_key;
static int idx_cmp(const void *a, const void *b) {
    uint64_t x = g_sort_key[*(const uint64_t *)a], y = g_sort_key[*(const uint64_t *)b];
    if (x != y) return x < y ? -1 : 1;
    uint64_t i = *(const uint64_t *)a, j = *(const uint64_t *)b;
    return i < j ? -1 : i > j;
}

/* sampler.go:65-154 with GOMAXPROCS-1 = 1 sampler goroutine; lengths and transcripts are visited in
 * ascending order (Go map order in the reference). */
int orc_ref_sample(orc_sim *s) {
    /* main.go:142-149 */
    go_rand_t *rg = &s->rg;
    go_rand_t ss;
    if (s->p.seed_sampling) { go_rand_seed(&ss, s->p.seed_sampling); rg = &ss; }
    else if (s->have_pcr_rg) rg = &s->rg_pcr;
    uint64_t n = s->nsp;
    /* view of the species ordered by (length, transcript, start) */
    uint64_t *ord = (uint64_t *)malloc((n + 1) * 8);
    uint64_t *key = (uint64_t *)malloc((n + 1) * 8);
    for (uint64_t i = 0; i < n; i++) { ord[i] = i; key[i] = ((uint64_t)(s->sp[i].end - s->sp[i].start) << 32) | s->sp[i].t; }
    g_sort_key = key;
    qsort(ord, n, 8, idx_cmp);
    uint64_t *trcount = (uint64_t *)calloc((size_t)s->ntr + 1, 8);
    uint64_t *scratch = (uint64_t *)malloc(((size_t)(s->ntr > (int)1 ? s->ntr : 1) + n + 1) * 8);
    /* per transcript: requests per length, kept as (t,len,req) triples */
    uint64_t nreq = 0, creq = 1024;
    uint64_t *req_t = (uint64_t *)malloc(creq * 8), *req_len = (uint64_t *)malloc(creq * 8), *req_n = (uint64_t *)malloc(creq * 8);
    int ovf = 0;
    uint64_t pos = 0;
    for (uint32_t len = 0; len < s->hn[ORC_H_TARGET]; len++) {
        uint64_t count = s->h[ORC_H_TARGET][len];
        if (!count) continue;
        while (pos < n && (key[ord[pos]] >> 32) < len) pos++;
        uint64_t b = pos, e = pos;
        while (e < n && (key[ord[e]] >> 32) == len) e++;
        /* pool.go:156-189: per-length (transcript, total) table */
        memset(trcount, 0, (size_t)s->ntr * 8);
        for (uint64_t i = b; i < e; i++) trcount[s->sp[ord[i]].t] += s->sp[ord[i]].count;
        uint64_t *picked = (uint64_t *)calloc((size_t)s->ntr + 1, 8);
        uint64_t i = 0;
        while (i < count) {
            uint64_t idx;
            if (e == b || !ref_sample_index_u64(rg, trcount, (uint64_t)s->ntr, scratch, &idx, &ovf)) break; /* pool.go:191-204 */
            trcount[idx]--;
            picked[idx]++;
            i++;
        }
        if (ovf) return orc_fail(s, "Integer overflow detected when calculating cumulative sum!");
        orc_hist_add(s, ORC_H_SAMPLED, len, i);
        s->h[ORC_H_AFTER_SAMPLING][len] -= i; /* fragstats.go:97-102 */
        orc_hist_add(s, ORC_H_MISSING, len, count - i);
        for (int t = 0; t < s->ntr; t++)
            if (picked[t]) {
                if (nreq == creq) {
                    creq *= 2;
                    req_t = (uint64_t *)realloc(req_t, creq * 8); req_len = (uint64_t *)realloc(req_len, creq * 8); req_n = (uint64_t *)realloc(req_n, creq * 8);
                }
                req_t[nreq] = (uint64_t)t; req_len[nreq] = len; req_n[nreq] = picked[t]; nreq++;
            }
        free(picked);
        pos = e;
    }
    /* sampler.go:112-121: one split generator */
    go_rand_t split;
    go_rand_seed(&split, go_int63(rg));
    /* sampler.go:125-146: per transcript, per requested length, draw fragments without replacement */
    uint64_t *cnt = (uint64_t *)malloc((n + 1) * 8);
    for (int t = 0; t < s->ntr; t++) {
        uint64_t fid = 0;
        for (uint64_t r = 0; r < nreq; r++) {
            if (req_t[r] != (uint64_t)t) continue;
            uint32_t len = (uint32_t)req_len[r];
            /* species of (t,len): contiguous in ord */
            uint64_t lo = 0, hi = n, target = ((uint64_t)len << 32) | (uint64_t)t;
            while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (key[ord[mid]] < target) lo = mid + 1; else hi = mid; }
            uint64_t b = lo, e = lo;
            while (e < n && key[ord[e]] == target) e++;
            for (uint64_t i = b; i < e; i++) cnt[i - b] = s->sp[ord[i]].count;
            for (uint64_t j = 0; j < req_n[r]; j++) {
                uint64_t idx;
                if (!ref_sample_index_u64(&split, cnt, e - b, scratch, &idx, &ovf)) return orc_fail(s, "Transcript is out of fragments! Simulation is inconsistent!");
                cnt[idx]--; /* transcript.go:183 */
                s->sp[ord[b + idx]].count--;
                int minus = go_float64(&split) < s->p.strand_bias; /* sampler.go:219-227 */
                orc_push_out(s, (uint32_t)t, s->sp[ord[b + idx]].start, s->sp[ord[b + idx]].end, minus, fid++);
            }
        }
    }
    free(ord); free(key); free(trcount); free(scratch); free(req_t); free(req_len); free(req_n); free(cnt);
    return 0;
}

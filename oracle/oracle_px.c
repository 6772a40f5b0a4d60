/* oracle/oracle_px.c - ORC_MODE_PHILOX: the normative "--rng=philox" specification,
 * executed sequentially on the CPU.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The CUDA kernels in rlsim_b200/csrc/ must reproduce
 * every number produced here bit for bit; that is what the `-m gpu` parity tests check.
 *
 * Same model as the reference (citations per function), with the reference's sequential math/rand stream
 * replaced by counter-based Philox4x32-10 streams addressed by *what* is being drawn (molecule, species,
 * length class), so the result does not depend on iteration order, thread count or GPU count:
 *   target draw i          : stream(TARGET; i)                         sequential cursor
 *   molecule m of tr. t    : stream(FRAG; t, m, sub)                   header: cursor; breaks/intervals: random access
 *   species (t,start,end)  : stream(PCR; t, start, end)                sequential cursor over all cycles
 *   length class l         : stream(SELECT; l)                         candidate j = block j
 *   copy j of a species    : stream(STRAND; t, start, end)             draw j
 */
#include "oracle_common.h"

/* ---------- variates of the spec ---------- */
static double px_normal(px_stream *st) { return det_ndtri(px_u01_open(px_next64(st))); }
static double px_expo(px_stream *st) { return -det_log(px_u01_open(px_next64(st))); }

/* Marsaglia & Tsang (2000) for shape >= 1, boosted by U^(1/shape) below 1.  Gamma(shape, 1). */
static double px_gamma(px_stream *st, double shape) {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double res;
    for (;;) {
        double x, v;
        do { x = px_normal(st); v = 1.0 + c * x; } while (v <= 0.0);
        v = v * v * v;
        double u = px_u01_open(px_next64(st));
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) { res = d * v; break; }
        if (det_log(u) < 0.5 * x2 + d * (1.0 - v + det_log(v))) { res = d * v; break; }
    }
    if (shape < 1.0) res = res * det_exp(det_log(px_u01_open(px_next64(st))) / shape);
    return res;
}

/* target_mix.go:55-72 + random.go:235-367 (same truncation rules: n and g compare floats, sn compares the
 * truncated integer; gamma is parameterised (mean=Location, shape=Scale) as target_mix.go:66) */
static uint32_t px_comp_len(px_stream *st, const orc_mixcomp *c) {
    if (c->low == c->high) return (uint32_t)c->low;
    double l = (double)c->low, h = (double)c->high;
    for (;;) {
        if (c->type == ORC_COMP_N) {
            double uf = px_normal(st) * c->scale + c->location;
            if (uf < l || uf > h) continue;
            return (uint32_t)uf;
        } else if (c->type == ORC_COMP_SN) {
            double u1 = px_normal(st);
            double u2 = px_normal(st);
            if (u2 > c->shape * u1) u1 = -u1;
            double x = c->location + c->scale * u1;
            if (!(x > -1.0) || !(x < 4294967296.0)) continue;
            uint64_t u = (uint64_t)(int64_t)x; /* (-1,0) truncates to 0 as Go does */
            if (u < c->low || u > c->high) continue;
            return (uint32_t)u;
        } else {
            double shape = c->scale, loc = c->location;
            double g = shape == 1.0 ? px_expo(st) : px_gamma(st, shape);
            double u = (loc / shape) * g;
            if (u < l || u > h) continue;
            return (uint32_t)u;
        }
    }
}
/* target_mix.go:232-245: one draw, components in declaration order */
static const orc_mixcomp *px_mix_comp(px_stream *st, const orc_mixcomp *c, int n) {
    double tot = 0.0;
    for (int i = 0; i < n; i++) tot = tot + c[i].weight;
    double u = px_u01(px_next64(st)) * tot;
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        acc = acc + c[i].weight;
        if (acc > u) return &c[i];
    }
    return &c[n - 1];
}
/* raw target: cumulative probabilities in file order, first csum > u (random.go:388-400) */
static uint32_t px_raw_len(orc_sim *s, px_stream *st) {
    double tot = 0.0;
    for (int i = 0; i < s->n_raw; i++) tot = tot + s->raw_prob[i];
    double u = px_u01(px_next64(st)) * tot;
    double acc = 0.0;
    for (int i = 0; i < s->n_raw; i++) {
        acc = acc + s->raw_prob[i];
        if (acc > u) return s->raw_len[i];
    }
    return s->raw_len[s->n_raw - 1];
}

/* Poisson: multiplication method below 12 (as random.go:127-139), PTRS (Hoermann 1993) above. */
static int32_t px_poisson(px_stream *st, double mean) {
    if (mean < 12.0) {
        double g = det_exp(-mean), t = 1.0;
        int32_t em = -1;
        for (;;) {
            em++;
            t = t * px_u01(px_next64(st));
            if (!(t > g)) return em;
        }
    }
    double slam = sqrt(mean), loglam = det_log(mean);
    double b = 0.931 + 2.53 * slam;
    double a = -0.059 + 0.02483 * b;
    double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        double U = px_u01(px_next64(st)) - 0.5;
        double V = px_u01_open(px_next64(st));
        double us = 0.5 - fabs(U);
        double k = floor((2.0 * a / us + b) * U + mean + 0.43);
        if (us >= 0.07 && V <= vr) return (int32_t)k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (det_log(V) + det_log(invalpha) - det_log(a / (us * us) + b) <= -mean + k * loglam - det_logfact(k)) return (int32_t)k;
    }
}

/* Binomial(n, pp): BINV inversion below n*min(p,q) = 10, BTRS (Hoermann 1993) above; exact in distribution, same law as
 * random.go:168-233.  Every attempt (inversion or BTRS) consumes ONE whole Philox block of the species' PCR stream:
 * inversion uses its low half, BTRS the low half for u and the high half for v.  1/x in the inversion recurrence is the
 * correctly rounded reciprocal (a table on the GPU), 1/b is shared by the two BTRS constants that need it. */
static uint64_t px_binomial(px_stream *st, uint64_t n, double pp) {
    if (n == 0 || !(pp > 0.0)) return 0;
    if (pp >= 1.0) return n;
    double p = pp <= 0.5 ? pp : 1.0 - pp;
    double q = 1.0 - p;
    double nd = (double)n;
    double r = p / q;
    uint64_t res;
    if (nd * p < 10.0) {
        double qn = det_exp(nd * det_log(q));
        double bound = nd * p + 10.0 * sqrt(nd * p * q + 1.0);
        if (bound > nd) bound = nd;
        for (;;) {
            double x = 0.0, px = qn, u = px_u01(px_blk_lo(px_next_block(st)));
            int restart = 0;
            while (u > px) {
                x = x + 1.0;
                if (x > bound) { restart = 1; break; }
                u = u - px;
                px = (((nd - x + 1.0) * r) * px) * (1.0 / x);
            }
            if (!restart) { res = (uint64_t)x; break; }
        }
    } else {
        double spq = sqrt(nd * p * q);
        double b = 1.15 + 2.53 * spq;
        double invb = 1.0 / b;
        double a = -0.0873 + 0.0248 * b + 0.01 * p;
        double c = nd * p + 0.5;
        double vr = 0.92 - 4.2 * invb;
        double alpha = (2.83 + 5.1 * invb) * spq;
        double m = floor((nd + 1.0) * p);
        for (;;) {
            px_block blk = px_next_block(st);
            double u = px_u01(px_blk_lo(blk)) - 0.5;
            double v = px_u01_open(px_blk_hi(blk));
            double us = 0.5 - fabs(u);
            double k = floor((2.0 * a / us + b) * u + c);
            if (us >= 0.07 && v <= vr) { res = (uint64_t)k; break; }
            if (k < 0.0 || k > nd) continue;
            v = det_log(v * alpha / (a / (us * us) + b));
            double ub = (m + 0.5) * det_log((m + 1.0) / (r * (nd - m + 1.0))) +
                        (nd + 1.0) * det_log((nd - m + 1.0) / (nd - k + 1.0)) +
                        (k + 0.5) * det_log(r * (nd - k + 1.0) / (k + 1.0)) +
                        det_logfact_tail(m) + det_logfact_tail(nd - m) - det_logfact_tail(k) - det_logfact_tail(nd - k);
            if (v <= ub) { res = (uint64_t)k; break; }
        }
    }
    return pp <= 0.5 ? res : n - res;
}

/* ---------- stream keys: main.go:72-81,129-149 seed policy ---------- */
static px_key key_for(const orc_sim *s, uint32_t tag) {
    int64_t si = s->p.seed_init, sp = s->p.seed_pcr ? s->p.seed_pcr : si;
    int64_t ss = s->p.seed_sampling ? s->p.seed_sampling : sp;
    switch (tag) {
    case PX_TAG_TARGET: case PX_TAG_FRAG: return px_derive_key(si, tag);
    case PX_TAG_PCR: return px_derive_key(sp, tag);
    default: return px_derive_key(ss, tag);
    }
}

uint64_t orc_px_binomial(int64_t seed_pcr, uint32_t t, uint32_t start, uint32_t end, uint64_t n, double p, int ncalls) {
    px_stream st;
    px_stream_init(&st, px_derive_key(seed_pcr, PX_TAG_PCR), t, start, end);
    uint64_t c = n;
    for (int i = 0; i < ncalls; i++) c += px_binomial(&st, c, p);
    return c;
}
int32_t orc_px_poisson(int64_t seed, uint32_t t, uint64_t m, double mean) {
    px_stream st;
    px_stream_init(&st, px_derive_key(seed, PX_TAG_FRAG), t, (uint32_t)m, ((uint32_t)PX_SUB_HEADER << 24) | (uint32_t)((m >> 32) & 0xffffff));
    return px_poisson(&st, mean);
}

/* ---------- a19: target.go:102-123 / raw_target.go:104-129 ---------- */
int orc_px_sample_target(orc_sim *s) {
    px_key key = key_for(s, PX_TAG_TARGET);
    for (int64_t i = 0; i < s->req_frags; i++) {
        px_stream st;
        px_stream_init(&st, key, (uint32_t)((uint64_t)i & 0xffffffffu), (uint32_t)((uint64_t)i >> 32), 0);
        uint32_t l = s->n_raw ? px_raw_len(s, &st) : px_comp_len(&st, px_mix_comp(&st, s->p.target, s->p.n_target));
        orc_hist_add(s, ORC_H_TARGET, l, 1);
    }
    return 0;
}

/* ---------- a3/a4: integer prefix form of nnthermo.go:177-217 (SURVEY section 7 hard part 4) ---------- */
static uint64_t *px_prefix(const char *plus, uint32_t L, int k, double T, int reverse) {
    uint32_t proflen = L > (uint32_t)k ? L - (uint32_t)k : 0;
    uint64_t *P = (uint64_t *)malloc(((size_t)proflen + 1) * 8);
    uint64_t *w = (uint64_t *)malloc(((size_t)proflen + 1) * 8);
    orc_quantized_weights(plus, L, k, T, reverse, w);
    P[0] = 0;
    for (uint32_t i = 0; i < proflen; i++) P[i + 1] = P[i] + w[i];
    free(w);
    return P;
}
static int px_priming(const px_stream *iv, uint32_t block, const uint64_t *P, uint32_t proflen, uint32_t start, uint32_t end, uint32_t *res) {
    if (end > proflen) end = proflen;
    if (start >= end) { *res = 0; return 0; }
    uint64_t total = P[end] - P[start];
    if (total == 0) return 1;
    uint64_t tgt = P[start] + px_below(iv, block, total);
    uint32_t lo = start, hi = end; /* smallest i in [start,end) with P[i+1] > tgt */
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (!(P[mid + 1] > tgt)) lo = mid + 1; else hi = mid;
    }
    *res = lo;
    return 0;
}

static int u64_cmp(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

typedef struct {
    orc_sim *s;
    uint32_t t, L, proflen;
    const uint64_t *Pf, *Pr;
    uint64_t *br;
    uint64_t cbr;
    px_key key;
} px_ctx;

static void px_sub(px_stream *st, const px_ctx *c, uint64_t m, uint32_t sub) {
    px_stream_init(st, c->key, c->t, (uint32_t)(m & 0xffffffffu), (sub << 24) | (uint32_t)((m >> 32) & 0xffffff));
}
static double px_target_location(orc_sim *s, px_stream *hdr) {
    if (s->n_raw || s->p.n_target == 0) return s->raw_location;
    return px_mix_comp(hdr, s->p.target, s->p.n_target)->location;
}
static int px_keep(const orc_sim *s, const px_stream *iv, uint32_t i) { /* fragfilter.go:45-50 */
    px_block b = px_stream_block(iv, 4 * i + 2);
    double u = px_u01((uint64_t)b.v[0] | ((uint64_t)b.v[1] << 32));
    return !(u < s->p.frag_loss_prob);
}
static void px_register(orc_sim *s, uint32_t t, uint32_t size, uint32_t start, uint32_t end) {
    orc_push_frag(s, t, start, end);
    orc_hist_add(s, ORC_H_AFTER_FRAG, size, 1);
}
static uint64_t *px_breaks(px_ctx *c, uint64_t n) {
    if (n > c->cbr) { c->cbr = n * 2; c->br = (uint64_t *)realloc(c->br, c->cbr * 8); }
    return c->br;
}
static int32_t px_nbreaks(px_stream *hdr, double rate) { /* fragmentor.go:113-117 */
    int32_t nb = px_poisson(hdr, rate) - 1;
    return nb <= 0 ? 1 : nb;
}

/* fragmentor.go:95-176 */
static int px_after_prim(px_ctx *c, uint64_t m, px_stream *hdr, int polyAend, int simPriming, int doublePrime) {
    orc_sim *s = c->s;
    double loc = px_target_location(s, hdr);
    int length = polyAend;
    if (length < 1) return 0;
    uint32_t low = (uint32_t)s->low, high = (uint32_t)s->high;
    double rate = s->p.frag_param == 0 ? ((double)length / loc) / 2.0 : (double)length / (double)(uint32_t)s->p.frag_param;
    int32_t nb = px_nbreaks(hdr, rate);
    px_stream bs, iv;
    px_sub(&bs, c, m, PX_SUB_BREAKS);
    px_sub(&iv, c, m, PX_SUB_INTERVAL);
    uint64_t *br = px_breaks(c, (uint64_t)nb + 2);
    br[0] = 0;
    br[1] = (uint64_t)length;
    for (int32_t j = 0; j < nb; j++) br[j + 2] = px_below(&bs, (uint32_t)j, (uint64_t)length);
    qsort(br, (size_t)nb + 2, 8, u64_cmp);
    for (int32_t i = 0; i < nb; i++) {
        uint32_t start = (uint32_t)br[i], end = (uint32_t)br[i + 1], new_start;
        if (end - start < low) continue;
        if (simPriming) {
            if (px_priming(&iv, 4 * (uint32_t)i, c->Pf, c->proflen, start, end, &new_start)) return orc_fail(s, "Error when simulating priming!");
            if (doublePrime) {
                uint32_t final = c->L - 1;
                uint32_t ss = final - end + 1, ee = final - new_start + 1, ns;
                if (px_priming(&iv, 4 * (uint32_t)i + 1, c->Pr, c->proflen, ss, ee, &ns)) return orc_fail(s, "Error when simulating priming!");
                end = final - ns + 1;
            }
        } else {
            new_start = start + (uint32_t)px_below(&iv, 4 * (uint32_t)i, end - start);
            if (doublePrime) end = new_start + (uint32_t)px_below(&iv, 4 * (uint32_t)i + 1, end - new_start);
        }
        uint32_t size = end - new_start;
        if (size < low || size > high) continue;
        if (!px_keep(s, &iv, (uint32_t)i)) continue;
        px_register(s, c->t, size, new_start, end);
    }
    return 0;
}

/* fragmentor.go:205-270 */
static int px_pre_prim(px_ctx *c, uint64_t m, px_stream *hdr, int polyAend) {
    orc_sim *s = c->s;
    double loc = px_target_location(s, hdr);
    int length = polyAend;
    if (length < 1) return 0;
    uint32_t low = (uint32_t)s->low, high = (uint32_t)s->high;
    double fp = 1.0 / (double)(uint32_t)s->p.frag_param;
    double ex = px_expo(hdr) / fp;
    int elong = ex >= (double)length ? length : (int)ex;
    elong = length - elong;
    double rate = (double)length / loc;
    int32_t nb = px_nbreaks(hdr, rate);
    int64_t len_tmp = (int64_t)length - (int64_t)elong;
    if (len_tmp < 1) return 0;
    px_stream bs, iv;
    px_sub(&bs, c, m, PX_SUB_BREAKS);
    px_sub(&iv, c, m, PX_SUB_INTERVAL);
    uint64_t *br = px_breaks(c, (uint64_t)nb + 2);
    br[0] = (uint64_t)elong;
    br[1] = (uint64_t)length;
    for (int32_t j = 0; j < nb; j++) br[j + 2] = (uint64_t)elong + px_below(&bs, (uint32_t)j, (uint64_t)len_tmp);
    qsort(br, (size_t)nb + 2, 8, u64_cmp);
    for (int32_t i = 0; i < nb; i++) {
        uint32_t start = (uint32_t)br[i], end = (uint32_t)br[i + 1];
        uint32_t size = end - start;
        if (size < low || size > high) continue;
        if (!px_keep(s, &iv, (uint32_t)i)) continue;
        px_register(s, c->t, size, start, end);
    }
    return 0;
}

/* fragmentor.go:303-351 */
static int px_prim_jump(px_ctx *c, uint64_t m, int polyAend) {
    orc_sim *s = c->s;
    int ltmp = (int)c->L - (int)c->proflen;
    if (polyAend < ltmp) return 0;
    uint32_t final = (uint32_t)(polyAend - ltmp);
    uint32_t low = (uint32_t)s->low, high = (uint32_t)s->high;
    double fp = s->p.frag_param != 0 ? 1.0 / (double)(uint32_t)s->p.frag_param : 0.0;
    px_stream iv, jl;
    px_sub(&iv, c, m, PX_SUB_INTERVAL);
    px_sub(&jl, c, m, 3u);
    uint32_t start = 0, end = 0, j = 0;
    while (end < final) {
        if (px_priming(&iv, 4 * j, c->Pf, c->proflen, end, final, &start)) return orc_fail(s, "Error when simulating priming!");
        uint64_t new_end;
        if (fp != 0.0) {
            double ex = px_expo(&jl) / fp;
            new_end = (uint64_t)start + (ex >= 4294967296.0 ? 4294967295ULL : (uint64_t)ex);
        } else {
            uint32_t l = s->n_raw ? px_raw_len(s, &jl) : px_comp_len(&jl, px_mix_comp(&jl, s->p.target, s->p.n_target));
            new_end = (uint64_t)start + l;
        }
        if (new_end > final) new_end = final;
        end = (uint32_t)new_end;
        uint32_t size = end - start;
        uint32_t jj = j++;
        if (size < low || size > high) continue;
        if (!px_keep(s, &iv, jj)) continue;
        px_register(s, c->t, size, start, end);
    }
    return 0;
}

/* pool.go:101-124 + transcript.go:139-150,279-284 */
int orc_px_fragment(orc_sim *s) {
    int k = s->p.kmer_len, mth = s->p.frag_method;
    int need_fwd = (mth == ORC_AFTER_PRIM || mth == ORC_AFTER_PRIM_DOUBLE || mth == ORC_PRIM_JUMP);
    int need_rev = (mth == ORC_AFTER_PRIM_DOUBLE);
    px_ctx c;
    memset(&c, 0, sizeof c);
    c.s = s;
    c.key = key_for(s, PX_TAG_FRAG);
    int rc = 0;
    for (int t = 0; t < s->ntr && !rc; t++) {
        if (s->expr[t] == 0) continue;
        uint32_t L = s->trlen[t];
        uint64_t *Pf = need_fwd ? px_prefix(s->plus[t], L, k, s->p.priming_temp, 0) : NULL;
        uint64_t *Pr = need_rev ? px_prefix(s->plus[t], L, k, s->p.priming_temp, 1) : NULL;
        c.t = (uint32_t)t; c.L = L; c.proflen = L > (uint32_t)k ? L - (uint32_t)k : 0; c.Pf = Pf; c.Pr = Pr;
        for (uint64_t m = 0; m < s->expr[t] && !rc; m++) {
            px_stream hdr;
            px_sub(&hdr, &c, m, PX_SUB_HEADER);
            uint32_t tail = px_comp_len(&hdr, px_mix_comp(&hdr, s->p.polya, s->p.n_polya));
            orc_hist_add(s, ORC_H_POLYA, tail, 1);
            int polyAend = (int)L - (int)s->polya_max + (int)tail;
            switch (mth) {
            case ORC_AFTER_PRIM: rc = px_after_prim(&c, m, &hdr, polyAend, 1, 0); break;
            case ORC_AFTER_PRIM_DOUBLE: rc = px_after_prim(&c, m, &hdr, polyAend, 1, 1); break;
            case ORC_AFTER_NOPRIM: rc = px_after_prim(&c, m, &hdr, polyAend, 0, 0); break;
            case ORC_AFTER_NOPRIM_DOUBLE: rc = px_after_prim(&c, m, &hdr, polyAend, 0, 1); break;
            case ORC_PRE_PRIM: rc = px_pre_prim(&c, m, &hdr, polyAend); break;
            default: rc = px_prim_jump(&c, m, polyAend); break;
            }
        }
        free(Pf); free(Pr);
    }
    free(c.br);
    return rc;
}

/* thermocycler.go:106-147 */
int orc_px_pcr(orc_sim *s) {
    px_key key = key_for(s, PX_TAG_PCR);
    for (uint64_t i = 0; i < s->nsp; i++) {
        orc_species *q = &s->sp[i];
        uint32_t len = q->end - q->start;
        q->gc = orc_gc_count(s->plus[q->t], q->start, q->end);
        double e = s->p.fixed_eff;
        if (e == 0.0) e = orc_len_eff(&s->p, s->low, s->high, len, 1) * orc_gc_eff(&s->p, q->gc, len, 1);
        q->eff = e;
        px_stream st;
        px_stream_init(&st, key, q->t, q->start, q->end);
        uint64_t c = q->count;
        for (int cyc = 0; cyc < s->p.nr_cycles; cyc++) {
            uint64_t old = c;
            c += px_binomial(&st, c, e);
            if (c < old) return orc_fail(s, "Integer overflow detected when amplifying fragments!");
        }
        q->count = c;
    }
    return 0;
}

/* ---------- a20-a23: size selection + final draw ---------- */
static uint64_t *g_key;
static int ord_cmp(const void *a, const void *b) {
    uint64_t i = *(const uint64_t *)a, j = *(const uint64_t *)b;
    if (g_key[i] != g_key[j]) return g_key[i] < g_key[j] ? -1 : 1;
    return i < j ? -1 : i > j;
}
static uint64_t hash64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

/* sampler.go:65-106 + pool.go:191-204 + transcript.go:172-187: per target length, draw min(req, T) of the T pooled
 * fragment copies of that length uniformly without replacement (multivariate hypergeometric over species).
 * Spec 4.6: the drawn (or, when more than half is drawn, the excluded) ranks are the first `pick` distinct
 * values of below(block j, T), j = 0,1,...; ranks index the copies of the class's species in (t,start,end) order. */
int orc_px_sample(orc_sim *s) {
    px_key ksel = key_for(s, PX_TAG_SELECT), kstr = key_for(s, PX_TAG_STRAND);
    uint64_t n = s->nsp;
    uint64_t *ord = (uint64_t *)malloc((n + 1) * 8), *key = (uint64_t *)malloc((n + 1) * 8), *pre = (uint64_t *)malloc((n + 2) * 8);
    for (uint64_t i = 0; i < n; i++) { ord[i] = i; key[i] = s->sp[i].end - s->sp[i].start; s->sp[i].taken = 0; }
    g_key = key;
    qsort(ord, n, 8, ord_cmp);
    uint64_t pos = 0;
    for (uint32_t len = 0; len < s->hn[ORC_H_TARGET]; len++) {
        uint64_t req = s->h[ORC_H_TARGET][len];
        if (!req) continue;
        while (pos < n && key[ord[pos]] < len) pos++;
        uint64_t b = pos, e = pos;
        while (e < n && key[ord[e]] == len) e++;
        pos = e;
        uint64_t T = 0;
        pre[0] = 0;
        for (uint64_t i = b; i < e; i++) {
            uint64_t nt = T + s->sp[ord[i]].count;
            if (nt < T) return orc_fail(s, "Integer overflow detected when calculating cumulative sum!");
            T = nt;
            pre[i - b + 1] = T;
        }
        uint64_t take = req < T ? req : T;
        orc_hist_add(s, ORC_H_SAMPLED, len, take);
        orc_hist_add(s, ORC_H_MISSING, len, req - take);
        if (len < s->hn[ORC_H_AFTER_SAMPLING]) s->h[ORC_H_AFTER_SAMPLING][len] -= take;
        if (take == 0) continue;
        if (take == T) { for (uint64_t i = b; i < e; i++) s->sp[ord[i]].taken = s->sp[ord[i]].count; continue; }
        int include = (take <= T - take);
        uint64_t pick = include ? take : T - take;
        uint64_t cap = 16;
        while (cap < pick * 2 + 2) cap *= 2;
        uint64_t *set = (uint64_t *)malloc(cap * 8);
        memset(set, 0xff, cap * 8);
        px_stream st;
        px_stream_init(&st, ksel, len, 0, 0);
        uint64_t got = 0;
        for (uint32_t j = 0; got < pick; j++) {
            uint64_t x = px_below(&st, j, T);
            uint64_t hsh = hash64(x) & (cap - 1);
            int dup = 0;
            while (set[hsh] != UINT64_MAX) { if (set[hsh] == x) { dup = 1; break; } hsh = (hsh + 1) & (cap - 1); }
            if (dup) continue;
            set[hsh] = x;
            got++;
            uint64_t lo = 0, hi = e - b; /* species i with pre[i] <= x < pre[i+1] */
            while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (!(pre[mid + 1] > x)) lo = mid + 1; else hi = mid; }
            s->sp[ord[b + lo]].taken++;
        }
        free(set);
        if (!include) for (uint64_t i = b; i < e; i++) s->sp[ord[i]].taken = s->sp[ord[i]].count - s->sp[ord[i]].taken;
    }
    /* sampler.go:187-227 + frag.go: output in (transcript, start, end) order, per-transcript ids, strand per copy */
    uint64_t fid = 0;
    uint32_t cur_t = UINT32_MAX;
    for (uint64_t i = 0; i < n; i++) {
        orc_species *q = &s->sp[i];
        if (q->t != cur_t) { cur_t = q->t; fid = 0; }
        if (!q->taken) continue;
        px_stream st;
        px_stream_init(&st, kstr, q->t, q->start, q->end);
        for (uint64_t j = 0; j < q->taken; j++) {
            int minus = px_u01(px_draw64(&st, j)) < s->p.strand_bias;
            orc_push_out(s, q->t, q->start, q->end, minus, fid++);
        }
        q->count -= q->taken;
    }
    free(ord); free(key); free(pre);
    return 0;
}

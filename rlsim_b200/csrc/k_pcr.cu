// PCR kernel (kernel 2): one thread per fragment species.
//
// Restates thermocycler.go:106-147 (Techne.Pcr / AmplifyFragment), :149-156 (CalcLengthEff, via a per-length table),
// :158-180 (CalcGcEff: the byte scan becomes two reads of the GC prefix counts) and the exact per-cycle draw
// c <- c + Binomial(c, e) of random.go:168-233, from the species' own Philox stream (key_pcr; t, start, end).
// Algorithmic traffic: 12 B coordinates + 4 B initial count + 2 x 4 B GC prefix + 8 B amplified count = 32 B/species.
#include "kernels.cuh"
#include "variates.cuh"

namespace rl {

__global__ void k_decode_species(DevTranscripts tr, const uint64_t *__restrict__ keys, DevSpecies sp, uint32_t len_bits) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    uint64_t key = keys[i];
    uint64_t gstart = key >> len_bits;
    uint32_t t = upper_bound_u64(tr.off, tr.n + 1, gstart) - 1;
    sp.tr[i] = t;
    sp.start[i] = (uint32_t)(gstart - tr.off[t]);
    sp.len[i] = (uint32_t)(key & ((1ull << len_bits) - 1ull));
}
void launch_decode_species(cudaStream_t st, const DevTranscripts &tr, const uint64_t *keys, DevSpecies sp, uint32_t high) {
    if (sp.n) k_decode_species<<<(unsigned)((sp.n + 255) / 256), 256, 0, st>>>(tr, keys, sp, key_len_bits(high));
}

// ---- asynchronous chunk pipeline: a chunk's species land behind those of the chunks before it, without a host round trip
// Runs after the run-length pass of a chunk (one thread): where the chunk's species go, and whether anything overflowed.
__global__ void k_species_begin(EagerChunk *cs, EagerGlobal *g, uint64_t key_cap, uint64_t q_cap,
                                uint32_t long_cap, uint64_t sp_cap) {
    unsigned long long n = cs->runs;
    if (cs->key_count > key_cap || cs->q_count > q_cap || cs->long_count > long_cap) { g->fail |= 1ull; n = 0; }
    if (g->sp_cursor + n > sp_cap) { g->fail |= 2ull; n = 0; }
    cs->sp_base = g->sp_cursor;
    cs->sp_n = n;
    g->sp_cursor += n;
    g->frag_total += cs->key_count;
}
void launch_species_begin(cudaStream_t st, EagerChunk *cs, EagerGlobal *g, uint64_t key_cap, uint64_t q_cap,
                          uint32_t long_cap, uint64_t sp_cap) {
    k_species_begin<<<1, 1, 0, st>>>(cs, g, key_cap, q_cap, long_cap, sp_cap);
}

__global__ void k_decode_species_dyn(DevTranscripts tr, const uint64_t *__restrict__ uniq, const uint32_t *__restrict__ run_len, DevSpecies sp,
                                     uint32_t len_bits, const EagerChunk *__restrict__ cs) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cs->sp_n) return;
    const uint64_t j = cs->sp_base + i;
    const uint64_t key = uniq[i];
    const uint64_t gstart = key >> len_bits;
    const uint32_t t = upper_bound_u64(tr.off, tr.n + 1, gstart) - 1;
    sp.tr[j] = t;
    sp.start[j] = (uint32_t)(gstart - tr.off[t]);
    sp.len[j] = (uint32_t)(key & ((1ull << len_bits) - 1ull));
    sp.count0[j] = run_len[i];
}
void launch_decode_species_dyn(cudaStream_t st, const DevTranscripts &tr, const uint64_t *uniq, const uint32_t *run_len, DevSpecies sp,
                               uint32_t high, const EagerChunk *cs, uint64_t n_bound) {
    if (n_bound) k_decode_species_dyn<<<(unsigned)((n_bound + 255) / 256), 256, 0, st>>>(tr, uniq, run_len, sp, key_len_bits(high), cs);
}

// CalcLenScalers + CalcLengthEff, thermocycler.go:83-92,149-156, with the library's restatement of Go's math.Pow
__global__ void k_len_eff(const __grid_constant__ DevModel m, double *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m.high - m.low) return;
    if (m.len_shape == 0.0) { out[i] = 1.0; return; }
    double alpha = det_go_pow((double)m.low, -m.len_shape);
    double beta = det_go_pow((double)m.high, -m.len_shape);
    double A = (m.len_max - m.len_min) / (alpha - beta);
    double B = m.len_min - A * beta;
    out[i] = A * det_go_pow((double)(m.low + i), -m.len_shape) + B;
}
void launch_len_eff(cudaStream_t st, const DevModel &m, double *len_eff) {
    uint32_t n = m.high - m.low + 1;
    k_len_eff<<<(n + 127) / 128, 128, 0, st>>>(m, len_eff);
}

// CalcGcEff, thermocycler.go:158-180
__device__ __forceinline__ double gc_eff(const DevModel &m, uint32_t gc_count, uint32_t len) {
    double gc = (double)gc_count / (double)len;
    if (m.gc_mode == 1) return m.gc_raw[(int)(gc * 100.0)];
    return m.gc_min + (m.gc_max - m.gc_min) * det_go_pow(1 - det_go_pow(gc, m.gc_shape), m.gc_shape);
}

// 1/x for the inversion recurrence (x <= 44 while n*min(p,q) < 10): correctly rounded reciprocals, spec 4.4
__constant__ double RCP_TABLE[128] = {     0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 /
    8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19,
    1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0
    / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 /
    42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49, 1.0 / 50, 1.0 / 51, 1.0 / 52, 1.0 / 53,
    1.0 / 54, 1.0 / 55, 1.0 / 56, 1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61, 1.0 / 62, 1.0 / 63, 1.0 / 64, 1.0
    / 65, 1.0 / 66, 1.0 / 67, 1.0 / 68, 1.0 / 69, 1.0 / 70, 1.0 / 71, 1.0 / 72, 1.0 / 73, 1.0 / 74, 1.0 / 75, 1.0 /
    76, 1.0 / 77, 1.0 / 78, 1.0 / 79, 1.0 / 80, 1.0 / 81, 1.0 / 82, 1.0 / 83, 1.0 / 84, 1.0 / 85, 1.0 / 86, 1.0 / 87,
    1.0 / 88, 1.0 / 89, 1.0 / 90, 1.0 / 91, 1.0 / 92, 1.0 / 93, 1.0 / 94, 1.0 / 95, 1.0 / 96, 1.0 / 97, 1.0 / 98, 1.0
    / 99, 1.0 / 100, 1.0 / 101, 1.0 / 102, 1.0 / 103, 1.0 / 104, 1.0 / 105, 1.0 / 106, 1.0 / 107, 1.0 / 108, 1.0 /
    109, 1.0 / 110, 1.0 / 111, 1.0 / 112, 1.0 / 113, 1.0 / 114, 1.0 / 115, 1.0 / 116, 1.0 / 117, 1.0 / 118, 1.0 / 119,
    1.0 / 120, 1.0 / 121, 1.0 / 122, 1.0 / 123, 1.0 / 124, 1.0 / 125, 1.0 / 126, 1.0 / 127};

// AmplifyFragment, thermocycler.go:133-147: `cycles` times c <- c + Binomial(c, pp), every attempt from one whole
// Philox block of the species' own stream (spec 4.4: BINV inversion below c*min(p,q) = 10, BTRS above).
//
// Scheduled per WARP, in rounds.  In a round every runnable lane makes exactly one attempt of its current cycle -
// Philox block, (setup), inversion walk or BTRS squeeze test - so that code runs with (nearly) all lanes.  An attempt
// that needs the exact BTRS acceptance test (four logarithms, ~14% of the attempts) only marks its lane PENDING; the
// test runs when at least PCR_SLOW_BATCH lanes are pending (or nothing else can run), for all of them together.
// (Round 1 found 10 the best batch for the FP64 test; since the test is decided in FP32 with FP32 Stirling tails the
// waiting costs more than the lanes it gathers, and the best batch is 1: the test runs at the end of the same round.)
// Per-species constants (p, q, p/q, log q) are hoisted out of the cycle loop.  Must be called by all 32 lanes
// (lanes without a species pass live = false).  The draws of a species depend only on its own block counter, so the
// schedule does not change the result: it is the sequence the oracle produces cycle by cycle.
#ifndef PCR_SLOW_BATCH_N
#define PCR_SLOW_BATCH_N 1   // measured 1.667 / 1.689 / 1.694 / 1.706 / 1.723 / 1.750 ms for 1 / 4 / 6 / 8 / 10 / 12 (C3)
#endif
constexpr int PCR_SLOW_BATCH = PCR_SLOW_BATCH_N;
// log(k!) - Stirling, for the FP32 pre-decision of the acceptance test (the spec's FP64 det_logfact_tail decides close calls)
__constant__ float LFT_F[10] = {0.08106146679532726f, 0.04134069595540929f, 0.02767792568499834f, 0.02079067210376509f,
                                0.01664469118982119f, 0.01387612882307075f, 0.01189670994589177f, 0.01041126526197209f,
                                0.009255462182712733f, 0.008330563433362871f};
__device__ __forceinline__ float logfact_tail_f(double k) {
    if (k <= 9.0) return LFT_F[(int)k];
    const float inv = 1.0f / (float)(k + 1.0);
    const float inv2 = inv * inv;
    return inv * (0.083333333f - inv2 * (0.0027777778f - inv2 * 0.00079365079f));
}
#ifndef PCR_BINV_MAX
#define PCR_BINV_MAX 10.0
#endif
__device__ __forceinline__ uint64_t amplify(const PxStream &stream, uint64_t c, double pp, int cycles, bool live, bool &overflow) {
    if (!(pp > 0.0) || c == 0) live = false;  // Binomial(c, 0) = 0, Binomial(0, .) = 0: nothing to draw
    if (live && pp >= 1.0) {                  // Binomial(c, 1) = c
        for (int cyc = 0; cyc < cycles; cyc++) {
            uint64_t nc = c + c;
            if (nc < c) { overflow = true; break; }
            c = nc;
        }
        live = false;
    }
    const bool flip = !(pp <= 0.5);
    const double p = flip ? 1.0 - pp : pp;
    const double q = 1.0 - p;
    const double r = p / q;
    const double lq = live ? det_log(q) : 0.0;
    int cyc = live ? 0 : cycles;
    uint32_t blk = 0;
    bool pending = false, have_setup = false, inv_mode = false;
    double nd = 0, b = 0, a = 0, cc = 0, vr = 0, alpha = 0, mm = 0, qn = 0, bound = 0;
    double us = 0, v = 0, k = 0;
    for (;;) {
        const bool runnable = cyc < cycles && !pending;
        if (!__any_sync(0xffffffffu, cyc < cycles)) break;
        bool done = false;   // this round's attempt produced the cycle's variate (committed once, at the end of the round)
        uint64_t res = 0;
        if (runnable) {
            if (!have_setup) {
                nd = (double)c;
                inv_mode = nd * p < PCR_BINV_MAX;
                if (inv_mode) {
                    qn = det_exp(nd * lq);
                    bound = nd * p + 10.0 * sqrt(nd * p * q + 1.0);
                    if (bound > nd) bound = nd;
                } else {
                    double spq = sqrt(nd * p * q);
                    b = 1.15 + 2.53 * spq;
                    double invb = 1.0 / b;
                    a = -0.0873 + 0.0248 * b + 0.01 * p;
                    cc = nd * p + 0.5;
                    vr = 0.92 - 4.2 * invb;
                    alpha = (2.83 + 5.1 * invb) * spq;
                    mm = floor((nd + 1.0) * p);
                }
                have_setup = true;
            }
            const PxBlock B = stream.block(blk++);
            if (inv_mode) {  // BINV: walk the pmf from 0
                double x = 0.0, px = qn, u = px_u01(px_lo64(B));
                bool restart = false;
                while (u > px) {
                    x = x + 1.0;
                    if (x > bound) { restart = true; break; }
                    u = u - px;
                    px = (((nd - x + 1.0) * r) * px) * RCP_TABLE[(int)x];
                }
                if (!restart) { res = (uint64_t)x; done = true; }
            } else {         // BTRS: squeeze
                double u = px_u01(px_lo64(B)) - 0.5;
                v = px_u01_open(px_hi64(B));
                us = 0.5 - fabs(u);
                k = floor((2.0 * a / us + b) * u + cc);
                if (us >= 0.07 && v <= vr) { res = (uint64_t)k; done = true; }
                else if (!(k < 0.0 || k > nd)) pending = true;
            }
        }
        const unsigned pend = __ballot_sync(0xffffffffu, pending);
        const unsigned run = __ballot_sync(0xffffffffu, !pending && (done ? cyc + 1 < cycles : cyc < cycles));
        if (pend != 0u && (__popc(pend) >= PCR_SLOW_BATCH || run == 0u)) {
            if (pending) {
                // Acceptance test  log(v * alpha / (a/us^2 + b))  <=  ub(k).  Decision first in FP32: every logarithm of ub
                // has an argument 1 + d with small d, d is formed from an exactly cancelling FP64 numerator, so the FP32
                // evaluation errs by less than ~5e-7 of the sum of the term magnitudes; when the gap is inside a 4x wider
                // margin the exact FP64 test of the spec decides.  Same decisions as the spec, a fraction of the FP64 work.
                const double t1 = r * (nd - mm + 1.0), t3 = k + 1.0, t2 = nd - k + 1.0;
                const float d1 = (float)((mm + 1.0) - t1) / (float)t1;         // (m+1)/(r(n-m+1)) - 1
                const float d2 = (float)(k - mm) / (float)t2;                  // (n-m+1)/(n-k+1) - 1
                const float d3 = (float)(r * t2 - t3) / (float)t3;             // r(n-k+1)/(k+1) - 1
                const float f1 = (float)(mm + 0.5) * log1pf(d1), f2 = (float)(nd + 1.0) * log1pf(d2), f3 = (float)(k + 0.5) * log1pf(d3);
                const float ft = (logfact_tail_f(mm) + logfact_tail_f(nd - mm)) - (logfact_tail_f(k) + logfact_tail_f(nd - k));
                const float usf = (float)us;
                const float lvf = logf((float)(v * alpha) / ((float)a / (usf * usf) + (float)b));
                const float gap = lvf - (f1 + f2 + f3 + ft);
                // 2e-6 * magnitudes: FP32 evaluation error (<= ~5e-7 of them); 2e-15 * coefficients: the spec's own FP64 quotient
                // rounding (1.1e-16 per argument) amplified by the coefficients, which matters once counts pass ~1e9
                // (the four tails, each below 0.084, and the quotient under the logarithm are evaluated in FP32 as well: a few
                // 1e-7 absolute, inside the constant term of the margin)
                const float margin = 2e-6f * (fabsf(f1) + fabsf(f2) + fabsf(f3) + fabsf(lvf) + 1.0f) + (float)(2e-15 * (mm + nd + k));
                bool accept;
                if (fabsf(gap) > margin && fabsf(gap) < 3.0e38f) {
                    accept = gap < 0.0f;
                } else {  // exact acceptance test of the spec
                    double lv = det_log(v * alpha / (a / (us * us) + b));
                    double ub = (mm + 0.5) * det_log((mm + 1.0) / (r * (nd - mm + 1.0))) +
                                (nd + 1.0) * det_log((nd - mm + 1.0) / (nd - k + 1.0)) +
                                (k + 0.5) * det_log(r * (nd - k + 1.0) / (k + 1.0)) + det_logfact_tail(mm) +
                                det_logfact_tail(nd - mm) - det_logfact_tail(k) - det_logfact_tail(nd - k);
                    accept = lv <= ub;
                }
                if (accept) { res = (uint64_t)k; done = true; }
                pending = false;
            }
        }
        if (done) {
            const uint64_t nc = c + (flip ? c - res : res);
            if (nc < c) { overflow = true; cyc = cycles; }
            else { c = nc; cyc++; have_setup = false; }
        }
    }
    return c;
}

constexpr int PCR_THREADS = 128;
template <bool DYN>
__global__ void __launch_bounds__(PCR_THREADS, 9)
k_pcr(const __grid_constant__ DevModel m, DevTranscripts tr, DevSpecies sp, int *error, const EagerChunk *__restrict__ cs) {
    uint64_t i = (uint64_t)blockIdx.x * PCR_THREADS + threadIdx.x;
    uint64_t n = sp.n;
    if (DYN) {  // asynchronous chunk pipeline: this launch covers the chunk's species, wherever they landed
        n = cs->sp_n;
        if ((uint64_t)blockIdx.x * PCR_THREADS >= n) return;
        n += cs->sp_base;
        i += cs->sp_base;
    }
    const bool live = i < n;
    uint32_t t = 0, start = 0, len = 1, gcc = 0;
    uint64_t c0 = 0;
    double e = m.fixed_eff;
    if (live) {
        t = sp.tr[i]; start = sp.start[i]; len = sp.len[i];
        c0 = sp.count0[i];
        if (e == 0.0 || sp.gc) {
            const uint2 *gcp = tr.gc + (tr.off[t] >> 5) + t;
            gcc = gc_before(gcp, start + len) - gc_before(gcp, start);
        }
        if (e == 0.0) e = m.len_eff[len - m.low] * gc_eff(m, gcc, len);
    }
    const uint32_t end = start + len;
    const PxStream stream{m.key_pcr, m.first_index + t, start, end};
    bool ovf = false;
    uint64_t c = amplify(stream, c0, e, m.cycles, live, ovf);
    if (!live) return;
    if (ovf) atomicExch(error, RLSIM_E_PCR_OVERFLOW);
    sp.count[i] = c;
    if (sp.eff) sp.eff[i] = e;
    if (sp.gc) sp.gc[i] = gcc;
}
void launch_pcr(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, DevSpecies sp, int *error) {
    if (sp.n) k_pcr<false><<<(unsigned)((sp.n + PCR_THREADS - 1) / PCR_THREADS), PCR_THREADS, 0, st>>>(m, tr, sp, error, nullptr);
}
void launch_pcr_dyn(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, DevSpecies sp, int *error, const EagerChunk *cs, uint64_t n_bound) {
    if (n_bound) k_pcr<true><<<(unsigned)((n_bound + PCR_THREADS - 1) / PCR_THREADS), PCR_THREADS, 0, st>>>(m, tr, sp, error, cs);
}

// ---- probes (parity tests of the building blocks) ----
__global__ void k_amplify_probe(PxKey key, const uint32_t *tse, const uint64_t *n0, const double *p, int cycles,
                                uint64_t *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint32_t j = live ? i : 0;
    const PxStream stream{key, tse[3 * j], tse[3 * j + 1], tse[3 * j + 2]};
    bool ovf = false;
    uint64_t r = amplify(stream, n0[j], p[j], cycles, live, ovf);
    if (live) out[i] = r;
}
void launch_amplify_probe(cudaStream_t st, PxKey key, const uint32_t *tse, const uint64_t *n0, const double *p, int cycles,
                          uint64_t *out, uint32_t n) {
    if (n) k_amplify_probe<<<(n + 127) / 128, 128, 0, st>>>(key, tse, n0, p, cycles, out, n);
}

__global__ void k_math_probe(int fn, const double *x, const double *y, double *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    switch (fn) {
    case 0: r = det_log(x[i]); break;
    case 1: r = det_exp(x[i]); break;
    case 2: r = det_ndtri(x[i]); break;
    case 3: r = det_logfact(x[i]); break;
    default: r = det_go_pow(x[i], y[i]); break;
    }
    out[i] = r;
}
void launch_math_probe(cudaStream_t st, int fn, const double *x, const double *y, double *out, uint32_t n) {
    if (n) k_math_probe<<<(n + 127) / 128, 128, 0, st>>>(fn, x, y, out, n);
}

__global__ void k_philox_probe(const uint32_t *ck, uint32_t *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PxBlock b = philox4x32_10(ck[6 * i], ck[6 * i + 1], ck[6 * i + 2], ck[6 * i + 3], ck[6 * i + 4], ck[6 * i + 5]);
    out[4 * i] = b.v0; out[4 * i + 1] = b.v1; out[4 * i + 2] = b.v2; out[4 * i + 3] = b.v3;
}
void launch_philox_probe(cudaStream_t st, const uint32_t *ck, uint32_t *out, uint32_t n) {
    if (n) k_philox_probe<<<(n + 127) / 128, 128, 0, st>>>(ck, out, n);
}

}  // namespace rl

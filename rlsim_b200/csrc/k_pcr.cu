// PCR kernel (kernel 2): one thread per fragment species.
//
// Restates thermocycler.go:106-147 (Techne.Pcr / AmplifyFragment), :149-156 (CalcLengthEff, via a per-length table),
// :158-180 (CalcGcEff: the byte scan becomes two reads of the GC prefix counts) and the exact per-cycle draw
// c <- c + Binomial(c, e) of random.go:168-233, from the species' own Philox stream (key_pcr; t, start, end).
// Algorithmic traffic: 12 B coordinates + 4 B initial count + 2 x 4 B GC prefix + 8 B amplified count = 32 B/species.
#include "kernels.cuh"
#include "variates.cuh"

namespace rl {

__global__ void k_decode_species(DevTranscripts tr, const uint64_t *__restrict__ keys, DevSpecies sp) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    uint64_t key = keys[i];
    uint64_t gstart = key >> 24;
    uint32_t t = upper_bound_u64(tr.off, tr.n + 1, gstart) - 1;
    sp.tr[i] = t;
    sp.start[i] = (uint32_t)(gstart - tr.off[t]);
    sp.len[i] = (uint32_t)(key & 0xffffffu);
}
void launch_decode_species(cudaStream_t st, const DevTranscripts &tr, const uint64_t *keys, DevSpecies sp) {
    if (sp.n) k_decode_species<<<(unsigned)((sp.n + 255) / 256), 256, 0, st>>>(tr, keys, sp);
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

// AmplifyFragment, thermocycler.go:133-147: `cycles` times c <- c + Binomial(c, pp), from the species' own cursor.
//
// Same draws, in the same per-species order, as px_binomial() called once per cycle (variates.cuh; the oracle restates
// that form) - but scheduled for a warp: the BTRS acceptance test that needs four logarithms ("slow path", ~14% of the
// attempts) is not evaluated inside the cycle loop, where it would run for a handful of lanes in almost every
// iteration.  A lane that needs it PARKS; lanes keep running fast-path work (inversion, squeeze accepts) until each
// one is parked or done, then the slow test runs once for all parked lanes together.  Per-species constants
// (p, q, p/q, log q) are hoisted out of the cycle loop.
// Must be called by ALL 32 lanes of a warp (lanes without a species pass live = false): the fast and slow phases
// are separated by __syncwarp() so the slow test really runs once per round for all parked lanes together.
__device__ __forceinline__ uint64_t amplify(PxCursor &cur, uint64_t c, double pp, int cycles, bool live, bool &overflow) {
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
    const double lq = det_log(q);
    int cyc = live ? 0 : cycles;
    bool have_setup = false;
    double nd = 0, b = 0, a = 0, cc = 0, vr = 0, alpha = 0, mm = 0;  // BTRS constants of the current cycle
    double us = 0, v = 0, k = 0;                                    // the parked attempt
    while (__any_sync(0xffffffffu, cyc < cycles)) {
        bool parked = false;
        // ---- fast phase: run until this lane needs the slow test or has finished all cycles
        while (cyc < cycles) {
            uint64_t res;
            if (!have_setup) {
                nd = (double)c;
                if (nd * p < 10.0) {  // BINV inversion
                    double qn = det_exp(nd * lq);
                    double bound = nd * p + 10.0 * sqrt(nd * p * q + 1.0);
                    if (bound > nd) bound = nd;
                    for (;;) {
                        double x = 0.0, px = qn, u = cur.u01();
                        bool restart = false;
                        while (u > px) {
                            x = x + 1.0;
                            if (x > bound) { restart = true; break; }
                            u = u - px;
                            px = ((nd - x + 1.0) * r) * px / x;
                        }
                        if (!restart) { res = (uint64_t)x; break; }
                    }
                    uint64_t nc = c + (flip ? c - res : res);
                    if (nc < c) { overflow = true; cyc = cycles; break; }
                    c = nc;
                    cyc++;
                    continue;
                }
                double spq = sqrt(nd * p * q);
                b = 1.15 + 2.53 * spq;
                a = -0.0873 + 0.0248 * b + 0.01 * p;
                cc = nd * p + 0.5;
                vr = 0.92 - 4.2 / b;
                alpha = (2.83 + 5.1 / b) * spq;
                mm = floor((nd + 1.0) * p);
                have_setup = true;
            }
            double u = cur.u01() - 0.5;
            v = cur.u01_open();
            us = 0.5 - fabs(u);
            k = floor((2.0 * a / us + b) * u + cc);
            if (us >= 0.07 && v <= vr) {  // squeeze accept
                res = (uint64_t)k;
                uint64_t nc = c + (flip ? c - res : res);
                if (nc < c) { overflow = true; cyc = cycles; break; }
                c = nc;
                cyc++;
                have_setup = false;
                continue;
            }
            if (k < 0.0 || k > nd) continue;
            parked = true;
            break;
        }
        __syncwarp();
        // ---- slow phase: exact acceptance test of the parked attempts, all parked lanes together
        if (parked) {
            double lv = det_log(v * alpha / (a / (us * us) + b));
            double ub = (mm + 0.5) * det_log((mm + 1.0) / (r * (nd - mm + 1.0))) +
                        (nd + 1.0) * det_log((nd - mm + 1.0) / (nd - k + 1.0)) +
                        (k + 0.5) * det_log(r * (nd - k + 1.0) / (k + 1.0)) + det_logfact_tail(mm) +
                        det_logfact_tail(nd - mm) - det_logfact_tail(k) - det_logfact_tail(nd - k);
            if (lv <= ub) {
                uint64_t res = (uint64_t)k;
                uint64_t nc = c + (flip ? c - res : res);
                if (nc < c) { overflow = true; cyc = cycles; }
                else { c = nc; cyc++; have_setup = false; }
            }
        }
        __syncwarp();
    }
    return c;
}

constexpr int PCR_THREADS = 128;
__global__ void __launch_bounds__(PCR_THREADS, 5)
k_pcr(const __grid_constant__ DevModel m, DevTranscripts tr, DevSpecies sp, int *error) {
    uint64_t i = (uint64_t)blockIdx.x * PCR_THREADS + threadIdx.x;
    const bool live = i < sp.n;
    uint32_t t = 0, start = 0, len = 1, gcc = 0;
    uint64_t c0 = 0;
    double e = m.fixed_eff;
    if (live) {
        t = sp.tr[i]; start = sp.start[i]; len = sp.len[i];
        c0 = sp.count0[i];
        if (e == 0.0 || sp.gc) {
            const uint32_t *gcp = tr.gc + tr.off[t] + t;
            gcc = gcp[start + len] - gcp[start];
        }
        if (e == 0.0) e = m.len_eff[len - m.low] * gc_eff(m, gcc, len);
    }
    const uint32_t end = start + len;
    PxCursor cur(PxStream{m.key_pcr, m.first_index + t, start, end});
    bool ovf = false;
    uint64_t c = amplify(cur, c0, e, m.cycles, live, ovf);
    if (!live) return;
    if (ovf) atomicExch(error, RLSIM_E_PCR_OVERFLOW);
    sp.count[i] = c;
    if (sp.eff) sp.eff[i] = e;
    if (sp.gc) sp.gc[i] = gcc;
}
void launch_pcr(cudaStream_t st, const DevModel &m, const DevTranscripts &tr, DevSpecies sp, int *error) {
    if (sp.n) k_pcr<<<(unsigned)((sp.n + PCR_THREADS - 1) / PCR_THREADS), PCR_THREADS, 0, st>>>(m, tr, sp, error);
}

// ---- probes (parity tests of the building blocks) ----
__global__ void k_amplify_probe(PxKey key, const uint32_t *tse, const uint64_t *n0, const double *p, int cycles,
                                uint64_t *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint32_t j = live ? i : 0;
    PxCursor cur(PxStream{key, tse[3 * j], tse[3 * j + 1], tse[3 * j + 2]});
    bool ovf = false;
    uint64_t r = amplify(cur, n0[j], p[j], cycles, live, ovf);
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

// Random variates of the "--rng=philox" spec, device side.
// Same laws as /root/reference/src/random.go:120-367; algorithms chosen for the GPU:
//   Poisson  : multiplication method below 12 (random.go:127-139), PTRS (Hoermann 1993) above
//   Binomial : BINV inversion below n*min(p,q)=10, BTRS (Hoermann 1993) above  (random.go:168-233: NR rejection); k_pcr.cu
//   normal   : inverse CDF (AS241) of one open uniform          (Go: ziggurat)
//   gamma    : Marsaglia-Tsang (random.go:282-367: Knuth / Kundu-Gupta)
// Compiled with -fmad=false; the oracle restates them in plain C and must agree bit for bit.
#pragma once
#include "common.cuh"
#include "detmath.cuh"

namespace rl {

template <class Cur> __device__ __forceinline__ double px_normal(Cur &c) { return det_ndtri(c.u01_open()); }
template <class Cur> __device__ __forceinline__ double px_expo(Cur &c) { return -det_log(c.u01_open()); }

template <class Cur> static __device__ __noinline__ double px_gamma(Cur &cur, double shape) {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double res;
    for (;;) {
        double x, v;
        do { x = px_normal(cur); v = 1.0 + c * x; } while (v <= 0.0);
        v = v * v * v;
        double u = cur.u01_open();
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) { res = d * v; break; }
        if (det_log(u) < 0.5 * x2 + d * (1.0 - v + det_log(v))) { res = d * v; break; }
    }
    if (shape < 1.0) res = res * det_exp(det_log(cur.u01_open()) / shape);
    return res;
}

// MixComp.SampleLength, target_mix.go:55-72
template <class Cur> static __device__ __noinline__ uint32_t px_comp_len(Cur &cur, const rlsim_mixcomp &c) {
    if (c.low == c.high) return (uint32_t)c.low;
    double l = (double)c.low, h = (double)c.high;
    for (;;) {
        if (c.type == RLSIM_COMP_N) {
            double uf = px_normal(cur) * c.scale + c.location;
            if (uf < l || uf > h) continue;
            return (uint32_t)uf;
        } else if (c.type == RLSIM_COMP_SN) {
            double u1 = px_normal(cur);
            double u2 = px_normal(cur);
            if (u2 > c.shape * u1) u1 = -u1;
            double x = c.location + c.scale * u1;
            if (!(x > -1.0) || !(x < 4294967296.0)) continue;
            uint64_t u = (uint64_t)(long long)x;
            if (u < c.low || u > c.high) continue;
            return (uint32_t)u;
        } else {
            double shape = c.scale, loc = c.location;
            double g = shape == 1.0 ? px_expo(cur) : px_gamma(cur, shape);
            double u = (loc / shape) * g;
            if (u < l || u > h) continue;
            return (uint32_t)u;
        }
    }
}

// TargetMix.SampleMixComp, target_mix.go:232-245 (declaration order)
template <class Cur> __device__ __forceinline__ int px_mix_comp(Cur &cur, const rlsim_mixcomp *c, int n) {
    double tot = 0.0;
    for (int i = 0; i < n; i++) tot = tot + c[i].weight;
    double u = cur.u01() * tot;
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        acc = acc + c[i].weight;
        if (acc > u) return i;
    }
    return n - 1;
}

// RawTarget.SampleMixLen, raw_target.go:148-154: cumulative probabilities in file order, first csum > u
template <class Cur> static __device__ __noinline__ uint32_t px_raw_len(Cur &cur, const DevModel &m) {
    double tot = 0.0;
    for (int i = 0; i < m.n_raw; i++) tot = tot + m.raw_prob[i];
    double u = cur.u01() * tot;
    double acc = 0.0;
    for (int i = 0; i < m.n_raw; i++) {
        acc = acc + m.raw_prob[i];
        if (acc > u) return m.raw_len[i];
    }
    return m.raw_len[m.n_raw - 1];
}

template <class Cur> static __device__ __noinline__ int32_t px_poisson(Cur &cur, double mean) {
    if (mean < 12.0) {
        double g = det_exp(-mean), t = 1.0;
        int32_t em = -1;
        for (;;) {
            em++;
            t = t * cur.u01();
            if (!(t > g)) return em;
        }
    }
    double slam = sqrt(mean), loglam = det_log(mean);
    double b = 0.931 + 2.53 * slam;
    double a = -0.059 + 0.02483 * b;
    double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        double U = cur.u01() - 0.5;
        double V = cur.u01_open();
        double us = 0.5 - fabs(U);
        double k = floor((2.0 * a / us + b) * U + mean + 0.43);
        if (us >= 0.07 && V <= vr) return (int32_t)k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (det_log(V) + det_log(invalpha) - det_log(a / (us * us) + b) <= -mean + k * loglam - det_logfact(k))
            return (int32_t)k;
    }
}

// Binomial draws live in k_pcr.cu (amplify): they are scheduled per warp, see there.

}  // namespace rl

/* oracle/go_rand.h - replica of Go's math/rand (Go <= 1.3 era) value stream.
 *
 * TEST INFRASTRUCTURE ONLY: part of the CPU oracle (see oracle/README.md); the
 * product (rlsim_b200/) never includes or links this.
 *
 * Why it exists: the reference draws every variate through `math/rand`
 * (/root/reference/src/random.go:59-110: RandGen wraps rand.New(rand.NewSource(seed));
 * random.go:184,205,220: Binomial uses the *package-global* rand.Float64()).
 * math/rand is Go standard-library code, absent from /root/reference and
 * unpinned (no go.mod; release binary = Go <= 1.3 stdlib).  Restated here from
 * its published algorithm:
 *   - rngSource: additive lagged Fibonacci x[n] = x[n-607] + x[n-273] mod 2^63,
 *     seeded through the Park-Miller LCG (48271 mod 2^31-1) XORed with the
 *     607-entry rng_cooked table (Go src/pkg/math/rand/rng.go);
 *   - Int63/Uint32/Int31/Int63n/Int31n/Float64 (rand.go);
 *   - NormFloat64 (normal.go) and ExpFloat64 (exp.go): Marsaglia-Tsang
 *     ziggurat with 128 / 256 blocks.
 * Constant tables come from the reference's own release binary
 * (oracle/gen_go_tables.py -> go_rand_tables.h).
 *
 * Pinning: tests/test_oracle_pin.py replays `rlsim -si S` and requires the
 * "Target lenghts", "Poly-A tail lengths" and "Fragdist after fragmentation"
 * report panels of the release binary to be reproduced exactly.
 */
#ifndef ORACLE_GO_RAND_H
#define ORACLE_GO_RAND_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "go_rand_tables.h"

#define GO_RNG_LEN 607
#define GO_RNG_TAP 273
#define GO_RNG_MASK 0x7fffffffffffffffLL

typedef struct {
    int tap, feed;
    int64_t vec[GO_RNG_LEN];
} go_rand_t;

static inline int32_t go_seedrand(int32_t x) {
    const int32_t A = 48271, Q = 44488, R = 3399;
    int32_t hi = x / Q, lo = x % Q;
    x = A * lo - R * hi;
    if (x < 0) x += 2147483647;
    return x;
}

static inline void go_rand_seed(go_rand_t *g, int64_t seed) {
    g->tap = 0;
    g->feed = GO_RNG_LEN - GO_RNG_TAP;
    seed = seed % 2147483647LL;
    if (seed < 0) seed += 2147483647LL;
    if (seed == 0) seed = 89482311;
    int32_t x = (int32_t)seed;
    for (int i = -20; i < GO_RNG_LEN; i++) {
        x = go_seedrand(x);
        if (i >= 0) {
            int64_t u = (int64_t)((uint64_t)x << 40);
            x = go_seedrand(x);
            u ^= (int64_t)((uint64_t)x << 20);
            x = go_seedrand(x);
            u ^= (int64_t)x;
            u ^= go_rng_cooked[i];
            g->vec[i] = u & GO_RNG_MASK;
        }
    }
}

static inline int64_t go_int63(go_rand_t *g) {
    if (--g->tap < 0) g->tap += GO_RNG_LEN;
    if (--g->feed < 0) g->feed += GO_RNG_LEN;
    int64_t x = (int64_t)(((uint64_t)g->vec[g->feed] + (uint64_t)g->vec[g->tap]) & (uint64_t)GO_RNG_MASK);
    g->vec[g->feed] = x;
    return x;
}

static inline uint32_t go_uint32(go_rand_t *g) { return (uint32_t)(go_int63(g) >> 31); }
static inline int32_t go_int31(go_rand_t *g) { return (int32_t)(go_int63(g) >> 32); }

static inline int64_t go_int63n(go_rand_t *g, int64_t n) {
    int64_t max = (int64_t)(((uint64_t)1 << 63) - 1 - (((uint64_t)1 << 63) % (uint64_t)n));
    int64_t v = go_int63(g);
    while (v > max) v = go_int63(g);
    return v % n;
}

static inline int32_t go_int31n(go_rand_t *g, int32_t n) {
    int32_t max = (int32_t)(((uint32_t)1 << 31) - 1 - (((uint32_t)1 << 31) % (uint32_t)n));
    int32_t v = go_int31(g);
    while (v > max) v = go_int31(g);
    return v % n;
}

/* Float64 of the release binary's stdlib.  GO_FLOAT64_VARIANT selects between
 * the Go 1.0/1.1/1.3+ definition (0: Int63()/2^63) and the Go 1.2 definition
 * (1: Int63n(2^53)/2^53).  Established empirically: only variant 0 reproduces
 * the release binary's "Target lenghts" panel (10000/10000 draws, -si 42). */
#ifndef GO_FLOAT64_VARIANT
#define GO_FLOAT64_VARIANT 0
#endif
static inline double go_float64(go_rand_t *g) {
#if GO_FLOAT64_VARIANT == 0
    return (double)go_int63(g) / 9223372036854775808.0;
#else
    return (double)go_int63n(g, (int64_t)1 << 53) / 9007199254740992.0;
#endif
}

static inline float go_f32(uint32_t bits) {
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline double go_norm_float64(go_rand_t *g) {
    const double rn = 3.442619855899;
    for (;;) {
        int32_t j = (int32_t)go_uint32(g);
        int32_t i = j & 0x7F;
        double x = (double)j * (double)go_f32(go_wn_bits[i]);
        uint32_t aj = j < 0 ? (uint32_t)(-(int64_t)j) : (uint32_t)j;
        if (aj < go_kn[i]) return x;
        if (i == 0) {
            for (;;) {
                x = -log(go_float64(g)) * (1.0 / rn);
                double y = -log(go_float64(g));
                if (y + y >= x * x) break;
            }
            if (j > 0) return rn + x;
            return -rn - x;
        }
        float fi = go_f32(go_fn_bits[i]), fim1 = go_f32(go_fn_bits[i - 1]);
        if (fi + (float)go_float64(g) * (fim1 - fi) < (float)exp(-.5 * x * x)) return x;
    }
}

static inline double go_exp_float64(go_rand_t *g) {
    const double re = 7.69711747013104972;
    for (;;) {
        uint32_t j = go_uint32(g);
        uint32_t i = j & 0xFF;
        double x = (double)j * (double)go_f32(go_we_bits[i]);
        if (j < go_ke[i]) return x;
        if (i == 0) return re - log(go_float64(g));
        float fi = go_f32(go_fe_bits[i]), fim1 = go_f32(go_fe_bits[i - 1]);
        if (fi + (float)go_float64(g) * (fim1 - fi) < (float)exp(-x)) return x;
    }
}

#endif

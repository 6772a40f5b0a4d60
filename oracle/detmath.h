/* oracle/detmath.h - deterministic FP64 math kernels of the "--rng=philox" spec.
 *
 * TEST INFRASTRUCTURE ONLY: CPU oracle (see oracle/README.md).  The product has
 * its own implementation of the same spec in rlsim_b200/csrc/detmath.cuh.
 *
 * Every function here is defined purely in terms of IEEE-754 binary64
 * + - * / sqrt floor and integer bit manipulation, evaluated in the written
 * order with NO fused multiply-add (build with -ffp-contract=off; the CUDA side
 * builds with -fmad=false).  Two conforming implementations therefore agree
 * bit for bit, which is what lets a counter-based stream reproduce the exact
 * same fragment set on CPU and GPU (SURVEY.md section 7 hard part 1).
 *
 *   det_log / det_exp : the classic fdlibm argument-reduction + minimax
 *                       polynomial algorithms (e_log.c / e_exp.c), < 1 ulp.
 *   det_ndtri         : Wichura's AS241 PPND16 inverse normal CDF.
 *   det_logfact_tail  : Stirling tail fc(k) = log k! - [(k+1/2)log(k+1) - (k+1) + log(2pi)/2]
 *   det_logfact       : log k!
 *   det_go_pow        : restatement of the *structure* of Go's math.Pow
 *                       (frexp / square-and-multiply / Exp(yf*Log(x))) used by
 *                       the reference at thermocycler.go:85-86,155,177.
 */
#ifndef ORACLE_DETMATH_H
#define ORACLE_DETMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint64_t det_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double det_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* x = m * 2^e with m in [0.5, 1), for finite positive normal or subnormal x. */
static inline double det_frexp(double x, int *e) {
    uint64_t u = det_bits(x);
    int ex = (int)((u >> 52) & 0x7ff);
    int adj = 0;
    if (ex == 0) { /* subnormal: scale by 2^54 */
        x = x * 18014398509481984.0;
        u = det_bits(x);
        ex = (int)((u >> 52) & 0x7ff);
        adj = -54;
    }
    *e = ex - 1022 + adj;
    u = (u & 0x800fffffffffffffULL) | 0x3fe0000000000000ULL;
    return det_from_bits(u);
}

/* y * 2^k, y in [0.25, 4]; handles overflow / gradual underflow. */
static inline double det_ldexp(double y, int k) {
    const double two_p1023 = det_from_bits((uint64_t)2046 << 52);  /* 2^1023 */
    const double two_m969 = det_from_bits((uint64_t)54 << 52);     /* 2^-969 */
    if (k > 3000) k = 3000;
    if (k < -3000) k = -3000;
    while (k > 1023) { y = y * two_p1023; k -= 1023; }
    while (k < -1022) { y = y * two_m969; k += 969; }
    return y * det_from_bits((uint64_t)(k + 1023) << 52);
}

static inline double det_log(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01, L3 = 2.857142874366239149e-01,
                 L4 = 2.222219843214978396e-01, L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                 L7 = 1.479819860511658591e-01;
    if (x != x) return x;
    if (x < 0.0) return det_from_bits(0x7ff8000000000000ULL);
    if (x == 0.0) return det_from_bits(0xfff0000000000000ULL);
    if (det_bits(x) == 0x7ff0000000000000ULL) return x;
    int ki;
    double f1 = det_frexp(x, &ki);
    if (f1 < 0.70710678118654752440) { f1 = f1 * 2.0; ki--; }
    double f = f1 - 1.0;
    double k = (double)ki;
    double s = f / (2.0 + f);
    double s2 = s * s;
    double s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2;
    double hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

static inline double det_exp(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
                 Log2e = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return det_from_bits(0x7ff0000000000000ULL);
    if (x < -7.45133219101941108420e+02) return 0.0;
    if (-3.725290298461914e-09 < x && x < 3.725290298461914e-09) return 1.0 + x;
    int k = 0;
    if (x < 0.0) k = (int)(Log2e * x - 0.5);
    else if (x > 0.0) k = (int)(Log2e * x + 0.5);
    double hi = x - (double)k * Ln2Hi;
    double lo = (double)k * Ln2Lo;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    return det_ldexp(y, k);
}

/* Inverse standard normal CDF, Wichura (1988) algorithm AS241 PPND16, 0<p<1. */
static inline double det_ndtri(double p) {
    double q = p - 0.5, r, val;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        return q * (((((((r * 2.5090809287301226727e+3 + 3.3430575583588128105e+4) * r + 6.7265770927008700853e+4) * r +
                         4.5921953931549871457e+4) * r + 1.3731693765509461125e+4) * r + 1.9715909503065514427e+3) * r +
                      1.3314166789178437745e+2) * r + 3.3871328727963666080e0) /
               (((((((r * 5.2264952788528545610e+3 + 2.8729085735721942674e+4) * r + 3.9307895800092710610e+4) * r +
                    2.1213794301586595867e+4) * r + 5.3941960214247511077e+3) * r + 6.8718700749205790830e+2) * r +
                 4.2313330701600911252e+1) * r + 1.0);
    }
    r = q < 0.0 ? p : 1.0 - p;
    r = sqrt(-det_log(r));
    if (r <= 5.0) {
        r = r - 1.6;
        val = (((((((r * 7.74545014278341407640e-4 + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r +
                   1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r +
                4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
              (((((((r * 1.05075007164441684324e-9 + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r +
                   1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r +
                2.05319162663775882187e0) * r + 1.0);
    } else {
        r = r - 5.0;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r +
                   2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r +
                5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
              (((((((r * 2.04426310338993978564e-15 + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r +
                   7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r +
                5.99832206555887937690e-1) * r + 1.0);
    }
    return q < 0.0 ? -val : val;
}

/* Stirling tail correction fc(k), k >= 0 integer-valued. */
static inline double det_logfact_tail(double k) {
    static const double t[10] = {0.08106146679532726, 0.04134069595540929, 0.02767792568499834,
                                 0.02079067210376509, 0.01664469118982119, 0.01387612882307075,
                                 0.01189670994589177, 0.01041126526197209, 0.009255462182712733,
                                 0.008330563433362871};
    if (k <= 9.0) return t[(int)k];
    double inv = 1.0 / (k + 1.0);
    double inv2 = inv * inv;
    return inv * (1.0 / 12.0 - inv2 * (1.0 / 360.0 - inv2 * (1.0 / 1260.0)));
}

/* log k!, k >= 0 integer-valued. */
static inline double det_logfact(double k) {
    double kp1 = k + 1.0;
    return (k + 0.5) * det_log(kp1) - kp1 + 0.91893853320467274178 + det_logfact_tail(k);
}

/* Go math.Pow structure (special cases restricted to what the reference can
 * produce: x >= 0 finite, y finite). */
static inline double det_go_pow(double x, double y) {
    if (y == 0.0 || x == 1.0) return 1.0;
    if (y == 1.0) return x;
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1.0 / sqrt(x);
    if (x != x || y != y) return det_from_bits(0x7ff8000000000000ULL);
    if (x == 0.0) return y < 0.0 ? det_from_bits(0x7ff0000000000000ULL) : 0.0;
    double absy = y;
    int flip = 0;
    if (absy < 0.0) { absy = -absy; flip = 1; }
    double yi = floor(absy), yf = absy - yi;
    if (yf != 0.0 && x < 0.0) return det_from_bits(0x7ff8000000000000ULL);
    if (yi >= 9.223372036854775808e18) return det_exp(y * det_log(x));
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0.0) {
        if (yf > 0.5) { yf = yf - 1.0; yi = yi + 1.0; }
        a1 = det_exp(yf * det_log(x));
    }
    int xe;
    double x1 = det_frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (i & 1) { a1 = a1 * x1; ae += xe; }
        x1 = x1 * x1;
        xe <<= 1;
        if (x1 < 0.5) { x1 = x1 + x1; xe--; }
    }
    if (flip) { a1 = 1.0 / a1; ae = -ae; }
    if (a1 == 0.0 || a1 == det_from_bits(0x7ff0000000000000ULL) || a1 != a1) return a1;
    /* a1 may be outside [0.25,4] after the loop; renormalise before scaling */
    int e2;
    double m = det_frexp(a1, &e2);
    return det_ldexp(m, ae + e2);
}
#endif

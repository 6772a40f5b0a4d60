// detmath.go - deterministic FP64 math of the "--rng=philox" specification (DESIGN.md section 4.2), restating
// oracle/detmath.h.  Every function is defined by IEEE-754 binary64 + - * / sqrt floor in the written order with NO fused
// multiply-add, so a Go host, the C oracle (-ffp-contract=off) and the CUDA kernels (-fmad=false) agree bit for bit
// (tests/test_gpu_parity.py::test_detmath_bit_exact holds the CUDA side to the oracle on 40 000 arguments per function).
//
// The Go compiler may fuse x*y + z on arm64, ppc64le, s390x and riscv64 (never on amd64); the language specification
// guarantees that an explicit float64(...) conversion rounds and so prevents fusion across it - hence the float64(x*y)
// wrappers wherever a product feeds a sum.
//
// Not compiled in this repository (no Go toolchain in the build image).
package main

import "math"

func detFrexp(x float64) (float64, int) { // x = m * 2^e, m in [0.5, 1), x finite positive
	u := math.Float64bits(x)
	ex := int((u >> 52) & 0x7ff)
	adj := 0
	if ex == 0 { // subnormal: scale by 2^54
		x = x * 18014398509481984.0
		u = math.Float64bits(x)
		ex = int((u >> 52) & 0x7ff)
		adj = -54
	}
	u = (u & 0x800fffffffffffff) | 0x3fe0000000000000
	return math.Float64frombits(u), ex - 1022 + adj
}

func detLdexp(y float64, k int) float64 { // y * 2^k, y in [0.25, 4]
	twoP1023 := math.Float64frombits(uint64(2046) << 52)
	twoM969 := math.Float64frombits(uint64(54) << 52)
	if k > 3000 {
		k = 3000
	}
	if k < -3000 {
		k = -3000
	}
	for k > 1023 {
		y = y * twoP1023
		k -= 1023
	}
	for k < -1022 {
		y = y * twoM969
		k += 969
	}
	return y * math.Float64frombits(uint64(k+1023)<<52)
}

// DetLog: fdlibm e_log.c reduction + polynomial.
func DetLog(x float64) float64 {
	const (
		Ln2Hi = 6.93147180369123816490e-01
		Ln2Lo = 1.90821492927058770002e-10
		L1    = 6.666666666666735130e-01
		L2    = 3.999999999940941908e-01
		L3    = 2.857142874366239149e-01
		L4    = 2.222219843214978396e-01
		L5    = 1.818357216161805012e-01
		L6    = 1.531383769920937332e-01
		L7    = 1.479819860511658591e-01
	)
	if x != x {
		return x
	}
	if x < 0.0 {
		return math.NaN()
	}
	if x == 0.0 {
		return math.Inf(-1)
	}
	if math.IsInf(x, 1) {
		return x
	}
	f1, ki := detFrexp(x)
	if f1 < 0.70710678118654752440 {
		f1 = f1 * 2.0
		ki--
	}
	f := f1 - 1.0
	k := float64(ki)
	s := f / (2.0 + f)
	s2 := float64(s * s)
	s4 := float64(s2 * s2)
	t1 := s2 * (L1 + float64(s4*(L3+float64(s4*(L5+float64(s4*L7))))))
	t2 := s4 * (L2 + float64(s4*(L4+float64(s4*L6))))
	R := float64(t1) + float64(t2)
	hfsq := float64(float64(0.5*f) * f)
	return float64(k*Ln2Hi) - ((hfsq - (float64(s*(hfsq+R)) + float64(k*Ln2Lo))) - f)
}

// DetExp: fdlibm e_exp.c reduction + polynomial.
func DetExp(x float64) float64 {
	const (
		Ln2Hi = 6.93147180369123816490e-01
		Ln2Lo = 1.90821492927058770002e-10
		Log2e = 1.44269504088896338700e+00
		P1    = 1.66666666666666019037e-01
		P2    = -2.77777777770155933842e-03
		P3    = 6.61375632143793436117e-05
		P4    = -1.65339022054652515390e-06
		P5    = 4.13813679705723846039e-08
	)
	if x != x {
		return x
	}
	if x > 7.09782712893383973096e+02 {
		return math.Inf(1)
	}
	if x < -7.45133219101941108420e+02 {
		return 0.0
	}
	if -3.725290298461914e-09 < x && x < 3.725290298461914e-09 {
		return 1.0 + x
	}
	k := 0
	if x < 0.0 {
		k = int(float64(Log2e*x) - 0.5)
	} else if x > 0.0 {
		k = int(float64(Log2e*x) + 0.5)
	}
	hi := x - float64(float64(k)*Ln2Hi)
	lo := float64(float64(k) * Ln2Lo)
	r := hi - lo
	t := float64(r * r)
	c := r - float64(t*(P1+float64(t*(P2+float64(t*(P3+float64(t*(P4+float64(t*P5)))))))))
	y := 1.0 - ((lo - float64(r*c)/(2.0-c)) - hi)
	return detLdexp(y, k)
}

func detHorner8(r float64, c [8]float64) float64 { // ((((((c0 r + c1) r + c2) ... ) r + c7
	v := c[0]
	for i := 1; i < 8; i++ {
		v = float64(v*r) + c[i]
	}
	return v
}

// DetNdtri: inverse standard normal CDF, Wichura's AS241 (PPND16), 0 < p < 1.
func DetNdtri(p float64) float64 {
	q := p - 0.5
	if math.Abs(q) <= 0.425 {
		r := 0.180625 - float64(q*q)
		num := detHorner8(r, [8]float64{2.5090809287301226727e+3, 3.3430575583588128105e+4, 6.7265770927008700853e+4,
			4.5921953931549871457e+4, 1.3731693765509461125e+4, 1.9715909503065514427e+3, 1.3314166789178437745e+2, 3.3871328727963666080e0})
		den := detHorner8(r, [8]float64{5.2264952788528545610e+3, 2.8729085735721942674e+4, 3.9307895800092710610e+4,
			2.1213794301586595867e+4, 5.3941960214247511077e+3, 6.8718700749205790830e+2, 4.2313330701600911252e+1, 1.0})
		return float64(q*num) / den
	}
	r := 1.0 - p
	if q < 0.0 {
		r = p
	}
	r = math.Sqrt(-DetLog(r))
	var val float64
	if r <= 5.0 {
		r = r - 1.6
		num := detHorner8(r, [8]float64{7.74545014278341407640e-4, 2.27238449892691845833e-2, 2.41780725177450611770e-1,
			1.27045825245236838258e0, 3.64784832476320460504e0, 5.76949722146069140550e0, 4.63033784615654529590e0, 1.42343711074968357734e0})
		den := detHorner8(r, [8]float64{1.05075007164441684324e-9, 5.47593808499534494600e-4, 1.51986665636164571966e-2,
			1.48103976427480074590e-1, 6.89767334985100004550e-1, 1.67638483018380384940e0, 2.05319162663775882187e0, 1.0})
		val = num / den
	} else {
		r = r - 5.0
		num := detHorner8(r, [8]float64{2.01033439929228813265e-7, 2.71155556874348757815e-5, 1.24266094738807843860e-3,
			2.65321895265761230930e-2, 2.96560571828504891230e-1, 1.78482653991729133580e0, 5.46378491116411436990e0, 6.65790464350110377720e0})
		den := detHorner8(r, [8]float64{2.04426310338993978564e-15, 1.42151175831644588870e-7, 1.84631831751005468180e-5,
			7.86869131145613259100e-4, 1.48753612908506148525e-2, 1.36929880922735805310e-1, 5.99832206555887937690e-1, 1.0})
		val = num / den
	}
	if q < 0.0 {
		return -val
	}
	return val
}

var detFcTable = [10]float64{0.08106146679532726, 0.04134069595540929, 0.02767792568499834, 0.02079067210376509,
	0.01664469118982119, 0.01387612882307075, 0.01189670994589177, 0.01041126526197209, 0.009255462182712733, 0.008330563433362871}

// DetLogfactTail: Stirling tail fc(k) = log k! - [(k+1/2) log(k+1) - (k+1) + log(2 pi)/2], k >= 0 integer-valued.
func DetLogfactTail(k float64) float64 {
	if k <= 9.0 {
		return detFcTable[int(k)]
	}
	inv := 1.0 / (k + 1.0)
	inv2 := float64(inv * inv)
	return inv * (1.0/12.0 - float64(inv2*(1.0/360.0-float64(inv2*(1.0/1260.0)))))
}

// DetLogfact: log k!.
func DetLogfact(k float64) float64 {
	kp1 := k + 1.0
	return float64((k+0.5)*DetLog(kp1)) - kp1 + 0.91893853320467274178 + DetLogfactTail(k)
}

// DetGoPow restates the structure of Go's own math.Pow (frexp / square-and-multiply / Exp(yf * Log(x))) with DetExp /
// DetLog inside, for x >= 0 finite and y finite (all the reference can produce: thermocycler.go:85-86,155,177).
func DetGoPow(x, y float64) float64 {
	if y == 0.0 || x == 1.0 {
		return 1.0
	}
	if y == 1.0 {
		return x
	}
	if y == 0.5 {
		return math.Sqrt(x)
	}
	if y == -0.5 {
		return 1.0 / math.Sqrt(x)
	}
	if x != x || y != y {
		return math.NaN()
	}
	if x == 0.0 {
		if y < 0.0 {
			return math.Inf(1)
		}
		return 0.0
	}
	absy, flip := y, false
	if absy < 0.0 {
		absy, flip = -absy, true
	}
	yi := math.Floor(absy)
	yf := absy - yi
	if yf != 0.0 && x < 0.0 {
		return math.NaN()
	}
	if yi >= 9.223372036854775808e18 {
		return DetExp(float64(y * DetLog(x)))
	}
	a1, ae := 1.0, 0
	if yf != 0.0 {
		if yf > 0.5 {
			yf = yf - 1.0
			yi = yi + 1.0
		}
		a1 = DetExp(float64(yf * DetLog(x)))
	}
	x1, xe := detFrexp(x)
	for i := int64(yi); i != 0; i >>= 1 {
		if i&1 == 1 {
			a1 = a1 * x1
			ae += xe
		}
		x1 = x1 * x1
		xe <<= 1
		if x1 < 0.5 {
			x1 = x1 + x1
			xe--
		}
	}
	if flip {
		a1 = 1.0 / a1
		ae = -ae
	}
	if a1 == 0.0 || math.IsInf(a1, 0) || a1 != a1 {
		return a1
	}
	m, e2 := detFrexp(a1) // a1 may be outside [0.25, 4] after the loop: renormalise before scaling
	return detLdexp(m, ae+e2)
}

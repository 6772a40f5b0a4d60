"""Efficiency functions for the report panels (host side).

Mirror of /root/reference/src/thermocycler.go:83-92,149-156,182-221.  `go_pow` restates the structure of Go's
math.Pow (special cases, Frexp + square-and-multiply for the integer part, Exp(yf*Log(x)) for the fraction) so the
panels follow the reference's arithmetic rather than the platform libm's pow.
"""
import math


def go_pow(x, y):
    if y == 0 or x == 1:
        return 1.0
    if y == 1:
        return x
    if y == 0.5:
        return math.sqrt(x)
    if y == -0.5:
        return 1.0 / math.sqrt(x)
    if math.isnan(x) or math.isnan(y):
        return math.nan
    if x == 0:
        return math.inf if y < 0 else 0.0
    absy, flip = (-y, True) if y < 0 else (y, False)
    yi = math.floor(absy)
    yf = absy - yi
    if yf != 0 and x < 0:
        return math.nan
    if yi >= 2.0 ** 63:
        return math.exp(y * math.log(x))
    a1, ae = 1.0, 0
    if yf != 0:
        if yf > 0.5:
            yf -= 1
            yi += 1
        a1 = math.exp(yf * math.log(x))
    x1, xe = math.frexp(x)
    i = int(yi)
    while i != 0:
        if i & 1:
            a1 *= x1
            ae += xe
        x1 *= x1
        xe <<= 1
        if x1 < 0.5:
            x1 += x1
            xe -= 1
        i >>= 1
    if flip:
        a1 = 1 / a1
        ae = -ae
    try:
        return math.ldexp(a1, ae)
    except OverflowError:
        return math.inf


def gc_eff_fixed(params, gc):  # thermocycler.go:182-193
    if params.gc_mode == 1:
        return params.gc_raw[int(gc * 100)]
    mn, mx, sh = params.gc_min, params.gc_max, params.gc_shape
    return mn + (mx - mn) * go_pow(1 - go_pow(gc, sh), sh)


def len_eff(params, low, high, l):  # thermocycler.go:83-92,149-156
    sh = params.len_shape
    if sh == 0.0:
        return 1.0
    alpha, beta = go_pow(float(low), -sh), go_pow(float(high), -sh)
    A = (params.len_max - params.len_min) / (alpha - beta)
    B = params.len_min - A * beta
    return A * go_pow(float(l), -sh) + B


def gc_panel(params):  # thermocycler.go:199-209
    xs, ys, gc = [], [], 0.0
    for _ in range(201):
        xs.append(gc)
        ys.append(gc_eff_fixed(params, gc / 100.0))
        gc += 0.5
    return xs, ys


def len_panel(params, low, high):  # thermocycler.go:211-220
    return {l: len_eff(params, low, high, l) for l in range(int(low), int(high) + 1)}

"""Mixture mini-language of the `-d` / `-a` flags: "w:type:(params) + ...".

Host-side mirror of /root/reference/src/target_mix.go:37-214 (MixComp, TargetMix,
parseTargetMixString, parseTargetMixComponent, validateMixComp) and of
target.go:72-87 (getGlobalMinMax).  Components keep their declaration order (the
reference stores them in a Go map, i.e. in unspecified order).
"""
from dataclasses import dataclass
from typing import List

COMP_TYPES = {"n": 0, "sn": 1, "g": 2}
TYPE_NAMES = {v: k for k, v in COMP_TYPES.items()}
MAX_COMPS = 8


class ArgError(ValueError):
    """Fatal argument error (the reference calls L.Fatal -> exit status 1)."""


@dataclass
class MixComp:
    type: str
    location: float
    scale: float
    shape: float = 0.0
    low: int = 0
    high: int = 0
    weight: float = 1.0

    def __str__(self):  # target_mix.go:46-53
        if self.type == "sn":
            return "%s:(%g, %g, %g, %d, %d)" % (self.type, self.location, self.scale, self.shape, self.low, self.high)
        return "%s:(%g, %g, %d, %d)" % (self.type, self.location, self.scale, self.low, self.high)


def _scan_float(s, what, ctx):
    try:
        return float(s.strip())
    except ValueError:
        raise ArgError('Invalid %s "%s" in mixture: "%s"' % (what, s, ctx))


def _scan_uint(s, what, ctx):
    try:
        v = int(s.strip())
    except ValueError:
        raise ArgError('Invalid %s "%s" in mixture: "%s"' % (what, s, ctx))
    if v < 0:
        raise ArgError('Invalid %s "%s" in mixture: "%s"' % (what, s, ctx))
    return v


def parse_component(s: str) -> MixComp:  # target_mix.go:128-214
    spl1 = s.split(":")
    if len(spl1) != 3:
        raise ArgError("Something is missing from mixture component: " + s)
    try:
        weight = float(spl1[0].strip())
    except ValueError:
        raise ArgError("Invalid mixture weight: " + spl1[0])
    ctype = spl1[1].strip()
    if ctype not in COMP_TYPES:
        raise ArgError("Invalid mixture type: %s" % ctype)
    tmp = spl1[2]
    if not tmp or tmp[0] != "(" or tmp[-1] != ")":
        raise ArgError("Missing paranthesis in mixture component string: " + tmp)
    tmp = tmp.strip("()")
    spl2 = tmp.split(",")
    if ctype == "sn":
        if len(spl2) != 5:
            raise ArgError('Not enough parameters in skew normal mixture string: "%s"' % tmp)
    elif len(spl2) != 4:
        raise ArgError('Not enough parameters in mixture string: "%s"' % tmp)
    i = 0
    loc = _scan_float(spl2[i], "location", tmp); i += 1
    scale = _scan_float(spl2[i], "scale", tmp); i += 1
    shape = 0.0
    if ctype == "sn":
        shape = _scan_float(spl2[i], "shape", tmp); i += 1
    low = _scan_uint(spl2[i], "lower bound", tmp); i += 1
    high = _scan_uint(spl2[i], "upper bound", tmp)
    if loc < 0:
        raise ArgError("Location parameter is negative in mixture string: %s" % s)
    if scale < 0:
        raise ArgError("Scale parameter is negative in mixture string: %s" % s)
    return MixComp(ctype, loc, scale, shape, low, high, weight)


def parse_mixture(s: str) -> List[MixComp]:  # target_mix.go:90-113
    comps = []
    for part in s.split("+"):
        part = part.strip()
        if not part:
            raise ArgError("Empty target mixture string!")
        comp = parse_component(part)
        if comp.high < comp.low:
            raise ArgError('Maximum fragment size is smaller than minimum fragment size for component "' + part + '"!')
        comps.append(comp)
    if len(comps) > MAX_COMPS:
        raise ArgError("At most %d mixture components are supported" % MAX_COMPS)
    return comps


def global_min_max(comps: List[MixComp]):  # target.go:72-87
    mn = mx = 0
    for i, c in enumerate(comps):
        if i == 0:
            mn = c.low
        mx = max(mx, c.high)
        mn = min(mn, c.low)
    return mn, mx


def mixture_string(comps: List[MixComp]) -> str:  # target_mix.go:78-88
    return "\n".join(" \t%.2f:%s" % (c.weight, c) for c in comps)

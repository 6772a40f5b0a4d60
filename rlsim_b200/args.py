"""Command-line arguments of rlsim, unchanged.

Host-side mirror of /root/reference/src/args.go:52-343 (CmdArgs.Parse, usage text,
default raw GC efficiencies, parseFragMethodString, ParseEffParamString) and of
parse_raw.go:39-126 (DecodeRawParams).  Flags follow Go's `flag` package: single or
double dash, `-flag value` or `-flag=value`, boolean flags take no value, parsing
stops at the first non-flag argument.  One flag is new: `-rng` (accepted values
"philox"; the GPU path always uses the counter-based Philox streams).
"""
import ctypes as C
import json
import os
import sys
from dataclasses import dataclass, field
from typing import List, Optional

from .mixture import ArgError, MixComp, MAX_COMPS, COMP_TYPES, parse_mixture, global_min_max

VERSION = "1.3"  # version.go:34

FRAG_METHODS = {"after_prim": 0, "after_prim_double": 1, "after_noprim": 2, "after_noprim_double": 3,
                "pre_prim": 4, "prim_jump": 5}

# Raw GC efficiencies estimated from SRR521457: the reference's built-in default for -eg (args.go:194).
# These 101 numbers are part of the CLI contract (same defaults => same library).
EFF_SRR521457 = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.6214855142919264, 0.4159233164222049,
    0.44285355190022124, 0.5111795670132255, 0.5261218695690415, 0.5657937478213575, 0.6057227372065885,
    0.652512493251505, 0.6551266520303076, 0.6586493220034146, 0.6710292671335725, 0.6828989286775027,
    0.7103061475969692, 0.7145967526946446, 0.7125612213162864, 0.7158400735145201, 0.7361963080537015,
    0.7507700417579295, 0.7936248041169689, 0.8745263930903109, 0.8971210947107862, 0.8979381841721974,
    0.8723089598770069, 0.8768516147049334, 0.8892267818464259, 0.9064925780415922, 0.9077163532915999,
    0.8951126800219693, 0.9176033189485007, 0.9116016387064556, 0.8890492343084937, 0.8803727480781831,
    0.8898403139631976, 0.9348273906591118, 0.9387867073874407, 0.9213150498959719, 0.8888279850470155,
    0.8918981127741714, 0.8604755993712578, 0.8382051613193982, 0.8515837826695183, 0.858730482918872,
    0.8415828215289731, 0.8190334105144152, 0.8317518172797949, 0.8079382135754152, 0.7900809748209163,
    0.7800701153643972, 0.7698151649955114, 0.7549149228315972, 0.7393688207123867, 0.7085441619225441,
    0.7010892852241013, 0.6740704877921768, 0.6613179904098958, 0.6424878029890573, 0.6351387703085332,
    0.6218820987970053, 0.5928624580840685, 0.5646145379419862, 0.5097661577740067, 0.5432830324275337,
    0.5100248798209441, 0.47253575285403415, 0.3571200478171759, 0.31034679770879103, 0.23330181449341403,
    0.07433564353203437, 0, 0.27666333287124445, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]

TARGET_MIX_DEFAULT = "1.0:sn:(189, 24, -1.09975, 76, 294)"  # args.go:94
POLYA_MIX_DEFAULT = "1.0:g:(150,1,0,220)"  # args.go:95

USAGE = """Simulate RNA-seq library preparation with priming biases, PCR biases and size selection (version: %s).

Usage:
        rlsim [arguments] [transcriptome fasta files (optional)]

Optional arguments:
                argument                    type    default
        -n      requested fragments         int
        -d      fragment size distribution  string  "1.0:sn:(189, 24, -1.09975, 76, 294)"
        -f      fragmentation method        string  "after_prim_double"
        -b      strand bias                 float   0.5
        -c      PCR cycles                  int     11
        -p      priming bias parameter      float   5.0
        -k      primer length               int     6
        -a      poly(A) tail size dist.     string  "1.0:g:(150,1,0,220)"
        -flg    fragment loss probability   float   0.0
        -m      expression level multiplier float   1.0
        -e      fixed PCR efficiency        float   0.0
        -eg     GC efficiency parameters as "(shape, min, max)": raw from SRR521457
        -el     length efficiency parameters as "(shape, min, max)": (0.0, 1.0, 1.0)
        -j      raw parameter file          string  (superseeds -d, -c, -eg)
        -jm     minimum raw gc efficiency   float   0.0
        -r      report file                 string  "rlsim_report.json"
        -t      number of cores to use      int     4
        -g      keep fragments in memory    bool    false
        -si     initial random seed         int     from UTC time
        -sp     pcr random seed             int     auto
        -ss     sampling random seed        int     auto
        -gobdir fragment directory          string  "rlsim_gob_$PID"
        -v      toggle verbose mode         bool    false
        -h      print usage and exit        bool    false
        -V      print version and exit      bool    false
        -prof   write CPU profiling info    string  ""
        -gcfreq trigger garbage collection  int     100
        -randt  generate RNG test files     bool    false
        -rng    random stream               string  "philox" (GPU path)

Examples:
        rlsim -n 2000000 transcripts.fa
        cat transcripts.fa | rlsim -n 2000000
"""


# ---- ctypes mirror of include/rlsim_gpu.h ----
class CMixComp(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("weight", C.c_double), ("location", C.c_double),
                ("scale", C.c_double), ("shape", C.c_double), ("low", C.c_uint64), ("high", C.c_uint64)]


class CParams(C.Structure):
    _fields_ = [("kmer_len", C.c_int32), ("frag_method", C.c_int32), ("frag_param", C.c_int32),
                ("nr_cycles", C.c_int32), ("priming_temp", C.c_double), ("frag_loss_prob", C.c_double),
                ("fixed_eff", C.c_double), ("strand_bias", C.c_double), ("gc_mode", C.c_int32),
                ("n_target", C.c_int32), ("n_polya", C.c_int32), ("reserved", C.c_int32),
                ("gc_shape", C.c_double), ("gc_min", C.c_double), ("gc_max", C.c_double),
                ("gc_raw", C.c_double * 101), ("len_shape", C.c_double), ("len_min", C.c_double),
                ("len_max", C.c_double), ("target", CMixComp * MAX_COMPS), ("polya", CMixComp * MAX_COMPS),
                ("seed_init", C.c_int64), ("seed_pcr", C.c_int64), ("seed_sampling", C.c_int64)]


@dataclass
class FragMethod:  # args.go:38-41
    name: str = "after_prim_double"
    param: int = 0


@dataclass
class EffParam:  # args.go:43-47
    shape: float = 0.0
    min: float = 1.0
    max: float = 1.0


def parse_frag_method(s: str) -> FragMethod:  # args.go:277-307
    spl = s.split(":")
    err = "Malformed fragmentation method string: "
    if len(spl) > 2 or len(spl) < 1 or spl[0] not in FRAG_METHODS:
        raise ArgError(err + s)
    fm = FragMethod(spl[0], 0)
    if len(spl) == 2:
        try:
            fm.param = int(spl[1])
        except ValueError:
            raise ArgError(err + s)
    if fm.name == "pre_prim" and fm.param == 0:
        fm.param = 2000
    return fm


def parse_eff_param(s: str) -> EffParam:  # args.go:310-333
    s = s.strip()
    if not s or s[0] != "(" or s[-1] != ")":
        raise ArgError("Missing paranthesis in efficiency parameter string: " + s)
    body = s[1:-1]
    try:
        vals = [float(x) for x in body.split(",")]
        shape, mn, mx = vals
    except ValueError:
        raise ArgError("Failed to parse efficieny parameter string: " + body)
    if shape < 0:
        raise ArgError("Efficiency shape parameter %g is negative!" % shape)
    if mn < 0 or mn > 1.0:
        raise ArgError("The minimum efficiency %g is outside the range [0, 1]!" % mn)
    if mx < 0 or mx > 1.0:
        raise ArgError("The maximum efficiency %g is outside the range [0, 1]!" % mx)
    return EffParam(shape, mn, mx)


@dataclass
class RawParams:  # parse_raw.go:39-50
    req_frags: int
    nr_cycles: int
    lengths: List[int]
    probs: List[float]
    gc_effs: List[float]


def decode_raw_params(fname: str) -> RawParams:  # parse_raw.go:51-126
    try:
        with open(fname) as f:
            d = json.load(f)
    except (OSError, ValueError) as e:  # parse_raw.go:53-60: unreadable file or malformed JSON
        raise ArgError("Error when reading raw parameter file %s: %s" % (fname, e))
    if not isinstance(d, dict):
        raise ArgError("Error when reading raw parameter file %s: not a JSON object" % fname)
    for key in ("nr_frags", "nr_cycles", "frag_dist", "gc_eff"):
        if key not in d:
            raise ArgError("The %s key is missing from the raw parameter file!" % key)
    lengths, counts = [], []
    for k, v in d["frag_dist"].items():  # file order (the reference: Go map order)
        try:
            lengths.append(int(k))
            counts.append(float(v))
        except (TypeError, ValueError) as e:  # parse_raw.go:93-96
            raise ArgError("Failed to convert frag_dist key %s: %s" % (k, e))
    total = sum(counts)
    probs = [c / total for c in counts]
    gc = [0.0] * len(d["gc_eff"])
    for k, v in d["gc_eff"].items():
        try:
            gc[int(k)] = float(v)
        except (TypeError, ValueError, IndexError) as e:  # parse_raw.go:118-121
            raise ArgError("Failed to convert gc_eff key %s: %s" % (k, e))
    try:
        return RawParams(int(d["nr_frags"]), int(d["nr_cycles"]), lengths, probs, gc)
    except (TypeError, ValueError) as e:
        raise ArgError("Error when reading raw parameter file %s: %s" % (fname, e))


_FLAGS = {  # name -> (attribute, kind, default); args.go:98-125
    "n": ("req_frags", int, 0), "d": ("targ_mix", str, TARGET_MIX_DEFAULT), "c": ("nr_cycles", int, 11),
    "a": ("polya_params", str, POLYA_MIX_DEFAULT), "b": ("strand_bias", float, 0.5),
    "p": ("priming_temp", float, 5.0), "k": ("kmer_length", int, 6), "e": ("fixed_eff", float, 0.0),
    "eg": ("gc_eff_params", str, ""), "el": ("len_eff_params", str, "(0.0,1.0,1.0)"),
    "j": ("raw_params_file", str, ""), "jm": ("min_raw_gc_eff", float, 0.0),
    "r": ("report_file", str, "rlsim_report.json"), "t": ("max_procs", int, 4),
    "f": ("frag_method_str", str, "after_prim_double"), "flg": ("frag_loss_prob", float, 0.0),
    "m": ("expr_mul", float, 1.0), "g": ("gob", bool, False), "si": ("init_seed", int, 0),
    "sp": ("pcr_seed", int, 0), "ss": ("sampling_seed", int, 0), "gobdir": ("gob_dir", str, ""),
    "gcfreq": ("gc_freq", int, 100), "prof": ("prof_file", str, ""), "h": ("help", bool, False),
    "V": ("version", bool, False), "v": ("verbose", bool, False), "randt": ("randtest", bool, False),
    "rng": ("rng", str, "philox"),
}


class ExitRequest(Exception):
    """-h / -V: print and exit 0 (args.go:181-190)."""

    def __init__(self, text, code=0):
        super().__init__(text)
        self.text, self.code = text, code


@dataclass
class CmdArgs:  # args.go:52-80
    req_frags: int = 0
    nr_cycles: int = 11
    strand_bias: float = 0.5
    priming_temp: float = 5.0
    fixed_eff: float = 0.0
    gc_eff_param: Optional[EffParam] = None
    len_eff_param: EffParam = field(default_factory=EffParam)
    report_file: str = "rlsim_report.json"
    verbose: bool = False
    target_mix: List[MixComp] = field(default_factory=list)
    frag_method: FragMethod = field(default_factory=FragMethod)
    frag_loss_prob: float = 0.0
    input_files: List[str] = field(default_factory=list)
    max_procs: int = 4
    prof_file: str = ""
    kmer_length: int = 6
    gob_dir: str = ""
    gc_freq: int = 100
    polya_param: List[MixComp] = field(default_factory=list)
    init_seed: int = 0
    pcr_seed: int = 0
    sampling_seed: int = 0
    raw_params_file: str = ""
    raw_gc_effs: Optional[List[float]] = None
    raw_len_probs: Optional[RawParams] = None
    min_raw_gc_eff: float = 0.0
    expr_mul: float = 1.0
    rng: str = "philox"
    warnings: List[str] = field(default_factory=list)

    @classmethod
    def parse(cls, argv: List[str]) -> "CmdArgs":
        raw = {attr: default for attr, _, default in _FLAGS.values()}
        i = 0
        while i < len(argv):
            tok = argv[i]
            if len(tok) < 2 or tok[0] != "-":
                break
            if tok == "--":
                i += 1
                break
            name = tok.lstrip("-") if tok.startswith("--") else tok[1:]
            value = None
            if "=" in name:
                name, value = name.split("=", 1)
            if name not in _FLAGS:
                raise ArgError("flag provided but not defined: -%s" % name)
            attr, kind, _ = _FLAGS[name]
            if kind is bool:
                raw[attr] = True if value is None else value.lower() in ("1", "t", "true")
                i += 1
                continue
            if value is None:
                if i + 1 >= len(argv):
                    raise ArgError("flag needs an argument: -%s" % name)
                value = argv[i + 1]
                i += 1
            try:
                raw[attr] = kind(value)
            except ValueError:
                raise ArgError('invalid value "%s" for flag -%s' % (value, name))
            i += 1
        rest = argv[i:]
        if raw["help"]:
            raise ExitRequest(USAGE % VERSION)
        if raw["version"]:
            raise ExitRequest(VERSION + "\n")
        a = cls()
        for k in ("req_frags", "nr_cycles", "strand_bias", "priming_temp", "fixed_eff", "report_file", "verbose",
                  "frag_loss_prob", "max_procs", "prof_file", "gob_dir", "gc_freq", "init_seed", "pcr_seed",
                  "sampling_seed", "raw_params_file", "min_raw_gc_eff", "expr_mul", "rng"):
            setattr(a, k, raw[k])
        default_raw_gc = False
        if raw["gc_eff_params"] != "":
            a.gc_eff_param = parse_eff_param(raw["gc_eff_params"])
        else:
            a.raw_gc_effs = list(EFF_SRR521457)
            default_raw_gc = True
        a.len_eff_param = parse_eff_param(raw["len_eff_params"])
        if raw["kmer_length"] < 1:
            raise ArgError("Invalid primer length!")
        a.kmer_length = raw["kmer_length"]
        if a.frag_loss_prob < 0.0 or a.frag_loss_prob > 1:
            raise ArgError("The fragment loss probability must be in the interval [0, 1]!")
        a.polya_param = parse_mixture(raw["polya_params"])
        if a.req_frags < 0:
            raise ArgError("Illegal fragment number!")
        a.target_mix = parse_mixture(raw["targ_mix"])
        a.frag_method = parse_frag_method(raw["frag_method_str"])
        if not raw["gob"] and a.gob_dir == "":
            a.gob_dir = "rlsim_gob_%d" % os.getpid()
        a.input_files = rest
        if a.raw_params_file != "":  # args.go:247-255
            rp = decode_raw_params(a.raw_params_file)
            if a.req_frags == 0:
                a.req_frags = rp.req_frags
            a.nr_cycles = rp.nr_cycles
            a.raw_len_probs = rp
            a.raw_gc_effs = rp.gc_effs
        if a.rng != "philox":
            raise ArgError('Unsupported -rng "%s": the GPU path implements "philox" only' % a.rng)
        if default_raw_gc and a.req_frags > 0 and a.raw_params_file == "":
            a.warnings.append("WARNING: Using raw efficiencies estimated from SRR521457 by default")
            a.warnings.append("(see http://bit.ly/rlsim-pl and http://bit.ly/rlsim-pa for the details).")
        if a.req_frags == 0:
            raise ExitRequest("No fragments requested! Exiting.\n" + USAGE % VERSION)
        return a

    def processed_raw_gc_effs(self):  # thermocycler.go:94-104 (used when -e is 0 and raw effs are present)
        out = []
        for e in self.raw_gc_effs:
            e = min(e, 1.0)
            e = max(e, self.min_raw_gc_eff)
            out.append(e)
        return out

    def to_params(self) -> CParams:
        """POD block handed to rlsim_set_model (include/rlsim_gpu.h)."""
        p = CParams()
        p.kmer_len = self.kmer_length
        p.frag_method = FRAG_METHODS[self.frag_method.name]
        p.frag_param = self.frag_method.param
        p.nr_cycles = self.nr_cycles
        p.priming_temp = self.priming_temp
        p.frag_loss_prob = self.frag_loss_prob
        p.fixed_eff = self.fixed_eff
        p.strand_bias = self.strand_bias
        if self.raw_gc_effs is not None:  # thermocycler.go:62-66: raw effs win over -eg
            p.gc_mode = 1
            effs = self.processed_raw_gc_effs()
            if len(effs) != 101:
                raise ArgError("Raw GC efficiency table must have 101 entries, got %d" % len(effs))
            for i, e in enumerate(effs):
                p.gc_raw[i] = e
        else:
            p.gc_mode = 0
            p.gc_shape, p.gc_min, p.gc_max = self.gc_eff_param.shape, self.gc_eff_param.min, self.gc_eff_param.max
        p.len_shape, p.len_min, p.len_max = self.len_eff_param.shape, self.len_eff_param.min, self.len_eff_param.max
        raw_target = self.raw_len_probs is not None
        p.n_target = 0 if raw_target else len(self.target_mix)
        p.n_polya = len(self.polya_param)
        for dst, comps in ((p.target, [] if raw_target else self.target_mix), (p.polya, self.polya_param)):
            for i, c in enumerate(comps):
                dst[i].type = COMP_TYPES[c.type]
                dst[i].weight, dst[i].location, dst[i].scale, dst[i].shape = c.weight, c.location, c.scale, c.shape
                dst[i].low, dst[i].high = c.low, c.high
        p.seed_init, p.seed_pcr, p.seed_sampling = self.init_seed, self.pcr_seed, self.sampling_seed
        return p

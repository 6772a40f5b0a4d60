"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
CPU-baseline legs of bench.py - never by anything under rlsim_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_COMPS = 8
MODE_REF, MODE_PHILOX = 0, 1
H_TARGET, H_POLYA, H_AFTER_FRAG, H_AFTER_PCR, H_SAMPLED, H_MISSING, H_AFTER_SAMPLING = range(7)


class MixComp(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("weight", C.c_double), ("location", C.c_double),
                ("scale", C.c_double), ("shape", C.c_double), ("low", C.c_uint64), ("high", C.c_uint64)]


class Params(C.Structure):
    _fields_ = [("kmer_len", C.c_int32), ("frag_method", C.c_int32), ("frag_param", C.c_int32),
                ("nr_cycles", C.c_int32), ("priming_temp", C.c_double), ("frag_loss_prob", C.c_double),
                ("fixed_eff", C.c_double), ("strand_bias", C.c_double), ("gc_mode", C.c_int32),
                ("n_target", C.c_int32), ("n_polya", C.c_int32), ("reserved", C.c_int32),
                ("gc_shape", C.c_double), ("gc_min", C.c_double), ("gc_max", C.c_double),
                ("gc_raw", C.c_double * 101), ("len_shape", C.c_double), ("len_min", C.c_double),
                ("len_max", C.c_double), ("target", MixComp * MAX_COMPS), ("polya", MixComp * MAX_COMPS),
                ("seed_init", C.c_int64), ("seed_pcr", C.c_int64), ("seed_sampling", C.c_int64)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    d, u32, u64, i64, i32, vp = C.c_double, C.c_uint32, C.c_uint64, C.c_int64, C.c_int, C.c_void_p
    P = C.POINTER
    sig = {
        "orc_kmer_affinity": (d, [C.c_char_p, i32, d, i32]),
        "orc_binding_profile": (None, [C.c_char_p, u32, i32, d, i32, i32, vp]),
        "orc_revcomp": (None, [C.c_char_p, u32, vp]),
        "orc_gc_count": (u32, [C.c_char_p, u32, u32]),
        "orc_gc_eff": (d, [P(Params), u32, u32, i32]),
        "orc_gc_eff_fixed": (d, [P(Params), d, i32]),
        "orc_len_eff_table": (None, [P(Params), u64, u64, i32, vp]),
        "orc_quantized_weights": (None, [C.c_char_p, u32, i32, d, i32, vp]),
        "orc_det_log": (d, [d]), "orc_det_exp": (d, [d]), "orc_det_ndtri": (d, [d]),
        "orc_det_logfact": (d, [d]), "orc_det_go_pow": (d, [d, d]),
        "orc_philox": (None, [vp, vp, vp]), "orc_derive_key": (None, [i64, u32, vp]),
        "orc_px_binomial": (u64, [i64, u32, u32, u32, u64, d, i32]),
        "orc_px_poisson": (C.c_int32, [i64, u32, u64, d]),
        "orc_go_rand_floats": (None, [i64, i32, vp]),
        "orc_sim_new": (vp, [P(Params), i32]), "orc_sim_free": (None, [vp]),
        "orc_sim_error": (C.c_char_p, [vp]),
        "orc_sim_set_raw_target": (i32, [vp, vp, vp, i32]),
        "orc_sim_sample_target": (i32, [vp, i64]),
        "orc_sim_set_target_hist": (i32, [vp, vp, vp, i32]),
        "orc_sim_add_transcripts": (i32, [vp, vp, vp, vp, i32]),
        "orc_sim_fragment": (i32, [vp]), "orc_sim_pcr": (i32, [vp]), "orc_sim_sample": (i32, [vp]),
        "orc_sim_low": (u64, [vp]), "orc_sim_high": (u64, [vp]), "orc_sim_polya_max": (u32, [vp]),
        "orc_sim_hist": (P(u64), [vp, i32, P(u32)]),
        "orc_sim_n_species": (u64, [vp]),
        "orc_sim_species": (None, [vp] + [vp] * 7),
        "orc_sim_pool_total": (u64, [vp]), "orc_sim_n_sampled": (u64, [vp]),
        "orc_sim_sampled": (None, [vp] + [vp] * 5),
        "orc_sim_fasta": (u64, [vp, vp, vp, vp, u64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    pass


class Sim:
    """One run of the hot path through the oracle."""

    def __init__(self, params, mode):
        self.L = lib()
        if not isinstance(params, Params):
            params = Params.from_buffer_copy(bytes(params))
        self.params = params
        self.h = self.L.orc_sim_new(C.byref(params), mode)
        self.ntr = 0

    def close(self):
        if self.h:
            self.L.orc_sim_free(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            raise OracleError(self.L.orc_sim_error(self.h).decode())

    def set_raw_target(self, lengths, probs):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        self._chk(self.L.orc_sim_set_raw_target(self.h, _ptr(lengths), _ptr(probs), len(lengths)))

    def sample_target(self, n):
        self._chk(self.L.orc_sim_sample_target(self.h, int(n)))

    def set_target_hist(self, lengths, counts):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self._chk(self.L.orc_sim_set_target_hist(self.h, _ptr(lengths), _ptr(counts), len(lengths)))

    def add_transcripts(self, seqs, expr):
        """seqs: list of upper-case bytes objects; expr: expression levels."""
        bases = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, np.uint8)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        expr = np.ascontiguousarray(expr, dtype=np.uint64)
        bases = np.ascontiguousarray(bases)
        self._chk(self.L.orc_sim_add_transcripts(self.h, _ptr(bases), _ptr(off), _ptr(expr), len(seqs)))
        self.ntr += len(seqs)

    def fragment(self):
        self._chk(self.L.orc_sim_fragment(self.h))

    def pcr(self):
        self._chk(self.L.orc_sim_pcr(self.h))

    def sample(self):
        self._chk(self.L.orc_sim_sample(self.h))

    @property
    def low(self):
        return self.L.orc_sim_low(self.h)

    @property
    def high(self):
        return self.L.orc_sim_high(self.h)

    @property
    def polya_max(self):
        return self.L.orc_sim_polya_max(self.h)

    def hist(self, kind):
        n = C.c_uint32()
        p = self.L.orc_sim_hist(self.h, kind, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def hist_dict(self, kind):
        h = self.hist(kind)
        return {int(i): int(h[i]) for i in np.nonzero(h)[0]}

    def species(self):
        n = self.L.orc_sim_n_species(self.h)
        out = dict(tr=np.zeros(n, np.uint32), start=np.zeros(n, np.uint32), end=np.zeros(n, np.uint32),
                   count0=np.zeros(n, np.uint64), count=np.zeros(n, np.uint64), eff=np.zeros(n, np.float64),
                   gc=np.zeros(n, np.uint32))
        self.L.orc_sim_species(self.h, *[_ptr(out[k]) for k in ("tr", "start", "end", "count0", "count", "eff", "gc")])
        return out

    @property
    def pool_total(self):
        return self.L.orc_sim_pool_total(self.h)

    def sampled(self):
        n = self.L.orc_sim_n_sampled(self.h)
        out = dict(tr=np.zeros(n, np.uint32), start=np.zeros(n, np.uint32), end=np.zeros(n, np.uint32),
                   minus=np.zeros(n, np.uint8), id=np.zeros(n, np.uint64))
        self.L.orc_sim_sampled(self.h, *[_ptr(out[k]) for k in ("tr", "start", "end", "minus", "id")])
        return out

    def fasta(self, names):
        nb = np.frombuffer(b"".join(names), dtype=np.uint8)
        off = np.zeros(len(names) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(s) for s in names])
        nb = np.ascontiguousarray(nb)
        need = self.L.orc_sim_fasta(self.h, _ptr(nb), _ptr(off), None, 0)
        buf = np.zeros(need, dtype=np.uint8)
        self.L.orc_sim_fasta(self.h, _ptr(nb), _ptr(off), _ptr(buf), need)
        return buf.tobytes()

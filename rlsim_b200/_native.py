"""ctypes binding of rlsim_b200/librlsim_gpu.so (C ABI: include/rlsim_gpu.h).

This is the same binding a Go cgo shim would make.  There is no CPU fallback: if the shared
library is missing or no CUDA device is usable, every entry point raises.
"""
import ctypes as C
import os

import numpy as np

from .args import CParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLSIM_GPU_LIB") or os.path.join(HERE, "librlsim_gpu.so")  # override: A/B runs of two builds

H_TARGET, H_POLYA, H_AFTER_FRAG, H_AFTER_PCR, H_SAMPLED, H_MISSING, H_AFTER_SAMPLING = range(7)
T_NAMES = ["layout", "prefix", "fragment", "dedup", "pcr", "select", "expand", "format", "target", "intervals", "exchange"]
H_POOL_GLOBAL = 7
COMM_ID_BYTES = 128

E_CUDA, E_ARG, E_PRIMING, E_PCR_OVERFLOW, E_POOL_OVERFLOW, E_UNSUPPORTED, E_NOMEM = range(1, 8)  # include/rlsim_gpu.h

F_MIRRORED_REVERSE, F_EAGER, F_PAGEABLE_STAGED, F_GROUPS_SHIFT, F_FUSED = 1, 2, 4, 8, 16  # rlsim_stats.flags

EXPORTS = [
    "rlsim_create", "rlsim_destroy", "rlsim_last_error", "rlsim_set_model", "rlsim_set_len_eff", "rlsim_set_raw_target",
    "rlsim_sample_target", "rlsim_sample_target_range", "rlsim_set_target_hist", "rlsim_add_transcripts", "rlsim_add_transcripts_packed", "rlsim_add_fasta", "rlsim_get_fasta_index", "rlsim_set_first_index", "rlsim_clear_transcripts",
    "rlsim_build_library", "rlsim_fragment", "rlsim_pcr", "rlsim_class_totals", "rlsim_sample", "rlsim_select_classes", "rlsim_apply_selected", "rlsim_finish_sample", "rlsim_get_stats",
    "rlsim_get_hist", "rlsim_get_species", "rlsim_get_sampled", "rlsim_get_sampled_packed", "rlsim_format_fasta", "rlsim_probe_kmer_lut",
    "rlsim_probe_prefix", "rlsim_probe_gc_prefix", "rlsim_probe_math", "rlsim_probe_philox", "rlsim_probe_amplify",
    "rlsim_get_timings", "rlsim_comm_unique_id", "rlsim_comm_init", "rlsim_comm_sum_target", "rlsim_sample_distributed",
    "rlsim_exchange_plan",
]


class Stats(C.Structure):
    _fields_ = [("n_transcripts", C.c_uint64), ("n_molecules", C.c_uint64), ("n_fragments", C.c_uint64),
                ("n_species", C.c_uint64), ("pool_total", C.c_uint64), ("n_requested", C.c_uint64),
                ("n_sampled", C.c_uint64), ("low", C.c_uint64), ("high", C.c_uint64), ("polya_max", C.c_uint32),
                ("flags", C.c_uint32)]


class RlsimError(RuntimeError):
    """Non-zero return code of the C ABI (the reference would L.Fatalf, logging.go:67-73)."""

    def __init__(self, code, msg):
        super().__init__("rlsim_gpu error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m rlsim_b200.build` (nvcc, sm_100a)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, i64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_int64
    sig = {
        "rlsim_create": (i32, [i32, C.POINTER(vp)]),
        "rlsim_destroy": (None, [vp]),
        "rlsim_last_error": (C.c_char_p, [vp]),
        "rlsim_set_model": (i32, [vp, C.POINTER(CParams)]),
        "rlsim_set_len_eff": (i32, [vp, vp, u64, u64]),
        "rlsim_set_raw_target": (i32, [vp, vp, vp, u32]),
        "rlsim_sample_target": (i32, [vp, i64]),
        "rlsim_sample_target_range": (i32, [vp, u64, u64]),
        "rlsim_set_target_hist": (i32, [vp, vp, vp, u32]),
        "rlsim_add_transcripts": (i32, [vp, vp, vp, vp, u32]),
        "rlsim_add_transcripts_packed": (i32, [vp, vp, vp, vp, vp, u32]),
        "rlsim_add_fasta": (i32, [vp, vp, u64, C.c_double, C.POINTER(u32), C.POINTER(u32)]),
        "rlsim_get_fasta_index": (i32, [vp] + [vp] * 6),
        "rlsim_set_first_index": (i32, [vp, u32]),
        "rlsim_clear_transcripts": (i32, [vp]),
        "rlsim_build_library": (i32, [vp]),
        "rlsim_fragment": (i32, [vp]),
        "rlsim_pcr": (i32, [vp]),
        "rlsim_class_totals": (i32, [vp, vp, u32]),
        "rlsim_sample": (i32, [vp, vp, vp, u32]),
        "rlsim_select_classes": (i32, [vp, vp, u32, u32, u32, vp, u64, vp, C.POINTER(u64)]),
        "rlsim_apply_selected": (i32, [vp, vp, u64]),
        "rlsim_finish_sample": (i32, [vp]),
        "rlsim_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "rlsim_get_hist": (i32, [vp, i32, vp, u32]),
        "rlsim_get_species": (i32, [vp] + [vp] * 7),
        "rlsim_get_sampled": (i32, [vp] + [vp] * 5),
        "rlsim_get_sampled_packed": (i32, [vp, vp, vp, vp, i32]),
        "rlsim_format_fasta": (i32, [vp, vp, vp, vp, u64, C.POINTER(u64)]),
        "rlsim_probe_kmer_lut": (i32, [vp, vp, vp, u32]),
        "rlsim_probe_prefix": (i32, [vp, u32, i32, vp, u32]),
        "rlsim_probe_gc_prefix": (i32, [vp, u32, vp, u32]),
        "rlsim_probe_math": (i32, [vp, i32, vp, vp, vp, u32]),
        "rlsim_probe_philox": (i32, [vp, vp, vp, u32]),
        "rlsim_probe_amplify": (i32, [vp, i64, vp, vp, vp, i32, vp, u32]),
        "rlsim_get_timings": (i32, [vp, vp, vp]),
        "rlsim_comm_unique_id": (i32, [vp]),
        "rlsim_comm_init": (i32, [vp, vp, i32, i32]),
        "rlsim_comm_sum_target": (i32, [vp]),
        "rlsim_sample_distributed": (i32, [vp]),
        "rlsim_exchange_plan": (i32, [vp, u32, u32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def pack_bases(bases):
    """ASCII bases (uint8 array, upper case) -> (codes, nmask) in the layout of rlsim_add_transcripts_packed: 2 bits per base
    (A C G T = 0 1 2 3, base i in bits 2*(i&3) of byte i>>2) and 1 bit per base for letters outside ACGT (bit i&7 of byte
    i>>3), or nmask = None when every letter is one of ACGT.  What a host that keeps its transcriptome packed does once."""
    b = np.ascontiguousarray(bases, dtype=np.uint8)
    lut = np.full(256, 4, np.uint8)
    lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
    c = lut[b]
    bad = c == 4
    c[bad] = 0
    pad = (-len(c)) % 4
    c4 = np.concatenate([c, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    codes = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    nmask = np.packbits(bad, bitorder="little") if bad.any() else None
    return codes, nmask


def comm_unique_id():
    """ncclUniqueId made by the library (rank 0 calls this; the host broadcasts the bytes to every rank)."""
    buf = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
    rc = load().rlsim_comm_unique_id(buf.ctypes.data_as(C.c_void_p))
    if rc:
        raise RlsimError(rc, "NCCL is not available (libnccl.so.2)")
    return buf.tobytes()


def exchange_plan(counts, rank):
    """Host arithmetic of the record routing: counts[s, d] = records sender s holds for rank d -> (send_off, recv_off)."""
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    n = c.shape[0]
    so, ro = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    rc = load().rlsim_exchange_plan(c.ctypes.data_as(C.c_void_p), n, int(rank), so.ctypes.data_as(C.c_void_p), ro.ctypes.data_as(C.c_void_p))
    if rc:
        raise RlsimError(rc, "invalid exchange plan arguments")
    return so, ro


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _addr(x):
    """numpy array or torch tensor (host) -> raw pointer."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return x.ctypes.data_as(C.c_void_p)


class Handle:
    """One GPU's pool (rlsim_handle)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.rlsim_create(int(device), C.byref(h))
        if rc:
            raise RlsimError(rc, self.L.rlsim_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.rlsim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RlsimError(rc, self.L.rlsim_last_error(self.h).decode())

    # ---- model / target
    def set_model(self, params: CParams):
        self._chk(self.L.rlsim_set_model(self.h, C.byref(params)))

    def set_len_eff(self, table, low, high):
        table = np.ascontiguousarray(table, dtype=np.float64)
        self._chk(self.L.rlsim_set_len_eff(self.h, _p(table), int(low), int(high)))

    def set_raw_target(self, lengths, probs):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        self._chk(self.L.rlsim_set_raw_target(self.h, _p(lengths), _p(probs), len(lengths)))

    def sample_target(self, n):
        self._chk(self.L.rlsim_sample_target(self.h, int(n)))

    def sample_target_range(self, first, count):
        self._chk(self.L.rlsim_sample_target_range(self.h, int(first), int(count)))

    def set_target_hist(self, lengths, counts):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self._chk(self.L.rlsim_set_target_hist(self.h, _p(lengths), _p(counts), len(lengths)))

    # ---- transcripts
    def set_first_index(self, i):
        self.first = int(i)
        self._chk(self.L.rlsim_set_first_index(self.h, int(i)))

    def clear_transcripts(self):
        self._chk(self.L.rlsim_clear_transcripts(self.h))

    def add_transcripts_raw(self, bases, offsets, expr, n):
        """bases/offsets/expr: host numpy arrays or (pinned) torch tensors."""
        self._chk(self.L.rlsim_add_transcripts(self.h, _addr(bases), _addr(offsets), _addr(expr), int(n)))

    def add_transcripts_packed(self, codes, nmask, offsets, expr, n):
        """codes / nmask as pack_bases() returns them (nmask may be None); offsets count bases."""
        self._chk(self.L.rlsim_add_transcripts_packed(self.h, _addr(codes), _addr(nmask) if nmask is not None else None,
                                                      _addr(offsets), _addr(expr), int(n)))

    def add_transcripts(self, seqs, expr):
        bases = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, np.uint8)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        if seqs:
            off[1:] = np.cumsum([len(s) for s in seqs])
        expr = np.ascontiguousarray(expr, dtype=np.uint64)
        self.add_transcripts_raw(np.ascontiguousarray(bases), off, expr, len(seqs))

    def add_fasta(self, text, expr_mul=1.0):
        """Raw FASTA text (bytes / uint8 array / pinned tensor) -> parsed and added on the device.
        Returns the record table of rlsim_get_fasta_index as a dict of numpy arrays (one entry per record found)."""
        if isinstance(text, (bytes, bytearray, memoryview)):
            text = np.frombuffer(text, dtype=np.uint8)
        nbytes = text.numel() if hasattr(text, "numel") else text.size
        n_rec, n_add = C.c_uint32(), C.c_uint32()
        self._chk(self.L.rlsim_add_fasta(self.h, _addr(text), int(nbytes), float(expr_mul), C.byref(n_rec), C.byref(n_add)))
        n = n_rec.value
        idx = dict(name_start=np.zeros(n, np.uint64), line_len=np.zeros(n, np.uint32), name_len=np.zeros(n, np.uint32),
                   level=np.zeros(n, np.uint64), seq_len=np.zeros(n, np.uint64), valid=np.zeros(n, np.uint8))
        self._chk(self.L.rlsim_get_fasta_index(self.h, *[_p(idx[k]) for k in ("name_start", "line_len", "name_len", "level", "seq_len", "valid")]))
        idx["n_added"] = n_add.value
        return idx

    # ---- stages
    def fragment(self):
        self._chk(self.L.rlsim_fragment(self.h))

    def pcr(self):
        self._chk(self.L.rlsim_pcr(self.h))

    def build_library(self):
        self._chk(self.L.rlsim_build_library(self.h))

    def class_totals(self, n_len=None):
        n_len = int(self.stats().high) + 1 if n_len is None else n_len
        out = np.zeros(n_len, dtype=np.uint64)
        self._chk(self.L.rlsim_class_totals(self.h, _p(out), n_len))
        return out

    def sample(self, global_totals=None, rank_base=None):
        if global_totals is None:
            self._chk(self.L.rlsim_sample(self.h, None, None, 0))
            return
        gt = np.ascontiguousarray(global_totals, dtype=np.uint64)
        rb = np.ascontiguousarray(rank_base, dtype=np.uint64)
        self._chk(self.L.rlsim_sample(self.h, _p(gt), _p(rb), len(gt)))

    def select_classes(self, all_totals, part, records_ptr=None, cap=0):
        """Multi-GPU selection: draw the classes l % nparts == part.  all_totals: [nparts, n_len] u64 (all-gathered
        class_totals); records_ptr: device pointer to cap 16-byte records.  -> (counts per owner rank, records needed)"""
        at = np.ascontiguousarray(all_totals, dtype=np.uint64)
        nparts, n_len = at.shape
        counts = np.zeros(nparts, dtype=np.uint64)
        needed = C.c_uint64()
        self._chk(self.L.rlsim_select_classes(self.h, _p(at), nparts, int(part), n_len,
                                              C.c_void_p(records_ptr) if records_ptr else None, int(cap), _p(counts), C.byref(needed)))
        return counts, int(needed.value)

    def apply_selected(self, records_ptr, n):
        self._chk(self.L.rlsim_apply_selected(self.h, C.c_void_p(records_ptr) if records_ptr else None, int(n)))

    def finish_sample(self):
        self._chk(self.L.rlsim_finish_sample(self.h))

    # ---- the exchange inside the library (NCCL communicator of the library's own)
    def comm_init(self, id_bytes, nranks, rank):
        buf = np.frombuffer(bytes(id_bytes), dtype=np.uint8).copy()
        assert buf.size == COMM_ID_BYTES
        self._chk(self.L.rlsim_comm_init(self.h, _p(buf), int(nranks), int(rank)))
        self.comm_ready = True

    def comm_sum_target(self):
        self._chk(self.L.rlsim_comm_sum_target(self.h))

    def sample_distributed(self):
        self._chk(self.L.rlsim_sample_distributed(self.h))

    # ---- results
    def stats(self):
        s = Stats()
        self._chk(self.L.rlsim_get_stats(self.h, C.byref(s)))
        return s

    def hist(self, kind, n=None):
        if n is None:
            st = self.stats()
            n = max(int(st.high), int(st.polya_max)) + 2
        out = np.zeros(n, dtype=np.uint64)
        self._chk(self.L.rlsim_get_hist(self.h, kind, _p(out), n))
        return out

    def hist_dict(self, kind, n=None):
        h = self.hist(kind, n)
        return {int(i): int(h[i]) for i in np.nonzero(h)[0]}

    def species(self):
        n = int(self.stats().n_species)
        out = dict(tr=np.zeros(n, np.uint32), start=np.zeros(n, np.uint32), end=np.zeros(n, np.uint32),
                   count0=np.zeros(n, np.uint32), count=np.zeros(n, np.uint64), eff=np.zeros(n, np.float64),
                   gc=np.zeros(n, np.uint32))
        self._chk(self.L.rlsim_get_species(self.h, *[_p(out[k]) for k in ("tr", "start", "end", "count0", "count", "eff", "gc")]))
        return out

    def sampled_unpacked(self, into=None):
        """rlsim_get_sampled: five arrays, 21 bytes per record."""
        n = int(self.stats().n_sampled)
        out = into or dict(tr=np.zeros(n, np.uint32), start=np.zeros(n, np.uint32), end=np.zeros(n, np.uint32),
                           minus=np.zeros(n, np.uint8), id=np.zeros(n, np.uint64))
        self._chk(self.L.rlsim_get_sampled(self.h, *[_addr(out[k]) for k in ("tr", "start", "end", "minus", "id")]))
        return out

    def sampled_packed(self, into=None, elem_bytes=None):
        """rlsim_get_sampled_packed: records per transcript + (start, length|strand) per record, 6 bytes per record.
        into: dict(tr_counts, start, len_strand) of host arrays / pinned tensors.  -> (dict, elem_bytes)"""
        st = self.stats()
        n, ntr = int(st.n_sampled), int(st.n_transcripts)
        if elem_bytes is None:
            elem_bytes = 2 if int(st.high) < 32768 else 4
        out = into or dict(tr_counts=np.zeros(ntr, np.uint32), start=np.zeros(n, np.uint32),
                           len_strand=np.zeros(n, np.uint16 if elem_bytes == 2 else np.uint32))
        self._chk(self.L.rlsim_get_sampled_packed(self.h, _addr(out["tr_counts"]), _addr(out["start"]), _addr(out["len_strand"]),
                                                  int(elem_bytes)))
        return out, elem_bytes

    @staticmethod
    def unpack_records(packed, elem_bytes, first=0, n=None):
        """Packed records -> the five Frag arrays (frag.go:46-52): what a host does while it prints the records."""
        cnt = np.asarray(packed["tr_counts"]).astype(np.int64)
        n = int(cnt.sum()) if n is None else n
        start = np.asarray(packed["start"])[:n].astype(np.uint32)
        ls = np.asarray(packed["len_strand"])[:n]
        bit = 8 * elem_bytes - 1
        length = (ls & ((1 << bit) - 1)).astype(np.uint32)
        tr = np.repeat(np.arange(len(cnt), dtype=np.int64) + first, cnt).astype(np.uint32)
        base = np.cumsum(cnt) - cnt
        ids = (np.arange(n, dtype=np.int64) - np.repeat(base, cnt)).astype(np.uint64)
        return dict(tr=tr, start=start, end=start + length, minus=(ls >> bit).astype(np.uint8), id=ids)

    def sampled(self, into=None):
        """The sampled fragments (Frag records) in output order.  Travels over the host link in the packed form and is
        expanded here; `into` (pre-allocated five-array buffers, e.g. pinned tensors) selects the unpacked copy."""
        if into is not None:
            return self.sampled_unpacked(into)
        st = self.stats()
        if int(st.n_sampled) >= 1 << 32:
            return self.sampled_unpacked()
        packed, eb = self.sampled_packed()
        return self.unpack_records(packed, eb, getattr(self, "first", 0), int(st.n_sampled))

    def fasta(self, names):
        nb = np.frombuffer(b"".join(names), dtype=np.uint8) if names else np.zeros(0, np.uint8)
        off = np.zeros(len(names) + 1, dtype=np.uint64)
        if names:
            off[1:] = np.cumsum([len(s) for s in names])
        nb = np.ascontiguousarray(nb)
        size = C.c_uint64()
        self._chk(self.L.rlsim_format_fasta(self.h, _p(nb), _p(off), None, 0, C.byref(size)))
        buf = np.zeros(max(size.value, 1), dtype=np.uint8)
        self._chk(self.L.rlsim_format_fasta(self.h, _p(nb), _p(off), _p(buf), size.value, C.byref(size)))
        return buf[:size.value].tobytes()

    def fasta_into(self, names_u8, names_off, buf):
        """Format the sampled fragments as FASTA text into a caller-owned (ideally pinned) uint8 buffer; returns the size.
        names_u8 / names_off: concatenated transcript names and their u64 offsets.  buf=None only sizes the text."""
        size = C.c_uint64()
        if buf is None:
            self._chk(self.L.rlsim_format_fasta(self.h, _addr(names_u8), _addr(names_off), None, 0, C.byref(size)))
        else:
            cap = buf.numel() if hasattr(buf, "numel") else buf.size
            self._chk(self.L.rlsim_format_fasta(self.h, _addr(names_u8), _addr(names_off), _addr(buf), cap, C.byref(size)))
        return size.value

    def timings(self):
        ms = np.zeros(len(T_NAMES), dtype=np.float32)
        nl = np.zeros(len(T_NAMES), dtype=np.uint64)
        self._chk(self.L.rlsim_get_timings(self.h, _p(ms), _p(nl)))
        return {k: float(v) for k, v in zip(T_NAMES, ms)}, {k: int(v) for k, v in zip(T_NAMES, nl)}

    # ---- probes
    def probe_kmer_lut(self, k):
        n = 1 << (2 * k)
        K = np.zeros(n, dtype=np.float64)
        w = np.zeros(n, dtype=np.uint32)
        self._chk(self.L.rlsim_probe_kmer_lut(self.h, _p(K), _p(w), n))
        return K, w

    def probe_prefix(self, tr, reverse, n):
        out = np.zeros(n, dtype=np.uint64)
        self._chk(self.L.rlsim_probe_prefix(self.h, tr, int(reverse), _p(out), n))
        return out

    def probe_gc_prefix(self, tr, n):
        out = np.zeros(n, dtype=np.uint32)
        self._chk(self.L.rlsim_probe_gc_prefix(self.h, tr, _p(out), n))
        return out

    def probe_math(self, fn, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        out = np.zeros_like(x)
        self._chk(self.L.rlsim_probe_math(self.h, fn, _p(x), _p(y), _p(out), len(x)))
        return out

    def probe_philox(self, ctr_key):
        ck = np.ascontiguousarray(ctr_key, dtype=np.uint32).reshape(-1, 6)
        out = np.zeros((len(ck), 4), dtype=np.uint32)
        self._chk(self.L.rlsim_probe_philox(self.h, _p(ck), _p(out), len(ck)))
        return out

    def probe_amplify(self, seed_pcr, tse, n0, p, cycles):
        tse = np.ascontiguousarray(tse, dtype=np.uint32).reshape(-1, 3)
        n0 = np.ascontiguousarray(n0, dtype=np.uint64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.zeros(len(n0), dtype=np.uint64)
        self._chk(self.L.rlsim_probe_amplify(self.h, int(seed_pcr), _p(tse), _p(n0), _p(p), int(cycles), _p(out), len(n0)))
        return out

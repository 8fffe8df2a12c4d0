"""ctypes front end of the seeded synthetic-alignment generator (libcsvsynth.so, host only)."""
import ctypes as C
import os

import numpy as np

from . import build as _build


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("profile", C.c_int32), ("coverage", C.c_double),
        ("read_len_mean", C.c_double), ("read_len_sd", C.c_double), ("indel_rate", C.c_double),
        ("indel_len_max", C.c_uint32), ("n_sv", C.c_uint64), ("sv_len_max", C.c_uint32),
        ("sv_jitter_sd", C.c_double), ("frac_len50", C.c_double), ("frac_softclip", C.c_double),
        ("frac_supplementary", C.c_double), ("frac_secondary", C.c_double), ("frac_dup", C.c_double),
        ("frac_qcfail", C.c_double), ("frac_lowmapq", C.c_double), ("use_eqx", C.c_int32), ("threads", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB_SYNTH
        if not os.path.exists(path):
            _build.build_synth()
        _lib = C.CDLL(path)
        _lib.csv_synth_num_reads.restype = C.c_uint64
    return _lib


def default_params(**kw):
    p = SynthParams()
    lib().csv_synth_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def generate(contig_len, params=None, alloc=None, count_gaps=True, **kw):
    """Returns a dict of numpy arrays in the csv_reads layout.

    alloc(shape, dtype) -> ndarray lets the caller place the arrays in pinned memory.
    """
    p = params if params is not None else default_params(**kw)
    alloc = alloc or (lambda n, dt: np.empty(n, dtype=dt))
    cl = np.ascontiguousarray(contig_len, dtype=np.uint32)
    L = lib()
    n = int(L.csv_synth_num_reads(C.byref(p), C.c_uint32(len(cl)), _ptr(cl)))
    tid = alloc(n, np.int32); pos0 = alloc(n, np.int32); flag = alloc(n, np.uint16)
    mapq = alloc(n, np.uint8); cig_off = alloc(n + 1, np.uint64)
    n_ops = C.c_uint64(0)
    L.csv_synth_reads(C.byref(p), C.c_uint32(len(cl)), _ptr(cl), _ptr(tid), _ptr(pos0), _ptr(flag), _ptr(mapq),
                      _ptr(cig_off), C.byref(n_ops))
    cigar = alloc(max(int(n_ops.value), 1), np.uint32)
    L.csv_synth_cigar(C.byref(p), C.c_uint32(len(cl)), _ptr(cl), _ptr(cig_off), _ptr(cigar))
    r = {"n_reads": n, "n_ops": int(n_ops.value), "tid": tid, "pos0": pos0, "flag": flag, "mapq": mapq,
         "cig_off": cig_off, "cigar": cigar[: int(n_ops.value)], "contig_len": cl}
    if count_gaps:
        # what a packer leaves in csv_reads::n_gap / ref_len while it copies the CIGAR words: D / N ops and reference bases per record
        from . import _capi
        r["n_gap"], r["ref_len"] = _capi.record_stats(r, alloc)
    return r

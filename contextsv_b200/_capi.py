"""ctypes bindings of the C ABI declared in include/contextsv_b200.h.

The CUDA library is the product: if libcontextsv_b200.so is missing or no B200 is
visible every compute call raises -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

CSV_OK = 0
CSV_ERR_CAPACITY = 3
STATUS_NAMES = {0: "CSV_OK", 1: "CSV_ERR_CUDA", 2: "CSV_ERR_ARG", 3: "CSV_ERR_CAPACITY", 4: "CSV_ERR_LIMIT", 5: "CSV_ERR_STATE"}


class CsvError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), msg))
        self.status = status


class CsvReads(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_ops", C.c_uint64), ("tid", C.c_void_p), ("pos0", C.c_void_p),
                ("flag", C.c_void_p), ("mapq", C.c_void_p), ("cig_off", C.c_void_p), ("cigar", C.c_void_p), ("n_gap", C.c_void_p), ("ref_len", C.c_void_p)]


class CsvRegion(C.Structure):
    _fields_ = [("tid", C.c_int32), ("beg", C.c_uint32), ("end", C.c_uint32), ("map_size", C.c_uint32)]


class CsvScanParams(C.Structure):
    _fields_ = [("min_len", C.c_uint32), ("min_mapq", C.c_uint8), ("want_depth", C.c_uint8), ("want_sigs", C.c_uint8),
                ("reserved", C.c_uint8)]


class CsvSigs(C.Structure):
    _fields_ = [("start", C.c_void_p), ("end", C.c_void_p), ("kind", C.c_void_p), ("read_idx", C.c_void_p),
                ("op_idx", C.c_void_p), ("query_pos", C.c_void_p)]


# every entry point include/contextsv_b200.h declares for libcontextsv_b200.so
EXPORTS = [
    "csv_ctx_create", "csv_ctx_destroy", "csv_ctx_sync", "csv_last_error", "csv_version", "csv_host_alloc", "csv_host_free",
    "csv_timer_begin", "csv_timer_end", "csv_ctx_launch_count", "csv_ctx_set_pipeline_chunks", "csv_profile_enable", "csv_profile_read", "csv_batch_upload", "csv_batch_free", "csv_scan_run",
    "csv_depth_stats", "csv_depth_fetch", "csv_depth_fetch_all", "csv_ctx_set_fetch", "csv_ctx_fetch_stats", "csv_host_widen_u8", "csv_depth_device_ptr", "csv_sigs_count", "csv_sigs_fetch", "csv_sigs_dbscan1d",
    "csv_depth", "csv_cigar_scan", "csv_dbscan1d", "csv_dbscan1d_seg", "csv_dbscan2d", "csv_largest_cluster", "csv_window_sums", "csv_depth_at", "csv_record_summary", "csv_host_count_gaps",
    "csv_depth_at_tid", "csv_sigs_depth", "csv_depth_checksum", "csv_batch_reserve_sigs", "csv_batch_release_inputs", "csv_host_record_stats", "csv_debug_fetch",
]
SYNTH_EXPORTS = ["csv_synth_default_params", "csv_synth_num_reads", "csv_synth_reads", "csv_synth_cigar"]

_lib = None


def lib():
    """Loads libcontextsv_b200.so (built in-tree by contextsv_b200.build)."""
    global _lib
    if _lib is None:
        path = _build.LIB_CUDA
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run `python -m contextsv_b200.build` (needs nvcc); there is no CPU fallback" % path)
        L = C.CDLL(path)
        L.csv_last_error.restype = C.c_char_p
        L.csv_version.restype = C.c_char_p
        L.csv_host_alloc.restype = C.c_void_p
        L.csv_host_alloc.argtypes = [C.c_size_t]
        L.csv_host_free.argtypes = [C.c_void_p]
        L.csv_ctx_launch_count.restype = C.c_uint64
        L.csv_ctx_launch_count.argtypes = [C.c_void_p]
        L.csv_largest_cluster.restype = C.c_uint64
        L.csv_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.csv_ctx_destroy.argtypes = [C.c_void_p]
        L.csv_ctx_sync.argtypes = [C.c_void_p]
        L.csv_timer_begin.argtypes = [C.c_void_p]
        L.csv_ctx_set_pipeline_chunks.argtypes = [C.c_void_p, C.c_int]
        L.csv_timer_end.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.csv_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.csv_profile_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.csv_batch_upload.argtypes = [C.c_void_p, C.POINTER(CsvReads), C.c_uint32, C.POINTER(CsvRegion), C.POINTER(C.c_void_p)]
        L.csv_batch_free.argtypes = [C.c_void_p, C.c_void_p]
        L.csv_scan_run.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CsvScanParams)]
        L.csv_depth_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.csv_depth_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.csv_depth_fetch_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.csv_ctx_set_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64]
        L.csv_ctx_fetch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.csv_host_widen_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.csv_host_widen_u8.restype = None
        L.csv_host_count_gaps.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.csv_host_count_gaps.restype = None
        L.csv_host_record_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.csv_host_record_stats.restype = None
        L.csv_sigs_count.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.csv_sigs_fetch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CsvSigs), C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]
        L.csv_sigs_dbscan1d.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_uint64]
        L.csv_depth.argtypes = [C.c_void_p, C.POINTER(CsvReads), C.POINTER(CsvRegion), C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.csv_cigar_scan.argtypes = [C.c_void_p, C.POINTER(CsvReads), C.POINTER(CsvRegion), C.c_uint32, C.c_uint8, C.POINTER(CsvSigs),
                                     C.c_uint64, C.POINTER(C.c_uint64)]
        L.csv_dbscan1d.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.csv_record_summary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.csv_depth_at.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
        L.csv_depth_at_tid.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.csv_sigs_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.csv_depth_checksum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.csv_debug_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.csv_batch_reserve_sigs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.csv_batch_release_inputs.argtypes = [C.c_void_p, C.c_void_p]
        L.csv_dbscan2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_void_p]
        L.csv_dbscan1d_seg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.csv_largest_cluster.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.csv_window_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(status):
    if status != CSV_OK:
        raise CsvError(status, lib().csv_last_error().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pinned_empty(n, dtype):
    """numpy array backed by cudaHostAlloc memory (freed when the array is collected)."""
    dt = np.dtype(dtype)
    nbytes = max(int(n) * dt.itemsize, 1)
    p = lib().csv_host_alloc(nbytes)
    if not p:
        raise CsvError(1, lib().csv_last_error().decode())
    buf = (C.c_uint8 * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    _PINNED[id(buf)] = (buf, p)
    import weakref
    weakref.finalize(arr, _free_pinned, id(buf))
    return arr


_PINNED = {}


def _free_pinned(key):
    ent = _PINNED.pop(key, None)
    if ent is not None and _lib is not None:
        _lib.csv_host_free(ent[1])


def reads_struct(r):
    """CsvReads over a dict of numpy arrays (kept alive by the caller)."""
    n = int(r["n_reads"])
    tid = r.get("tid")
    arrs = {
        "tid": None if tid is None else np.ascontiguousarray(tid, np.int32),
        "pos0": np.ascontiguousarray(r["pos0"], np.int32),
        "flag": np.ascontiguousarray(r["flag"], np.uint16),
        "mapq": np.ascontiguousarray(r["mapq"], np.uint8),
        "cig_off": np.ascontiguousarray(r["cig_off"], np.uint64),
        "cigar": np.ascontiguousarray(r["cigar"], np.uint32),
        "n_gap": None if r.get("n_gap") is None else np.ascontiguousarray(r["n_gap"], np.uint32),
        "ref_len": None if r.get("ref_len") is None else np.ascontiguousarray(r["ref_len"], np.uint32),
    }
    n_ops = int(arrs["cig_off"][n]) if n else 0
    s = CsvReads(n, n_ops, ptr(arrs["tid"]), ptr(arrs["pos0"]), ptr(arrs["flag"]), ptr(arrs["mapq"]), ptr(arrs["cig_off"]),
                 ptr(arrs["cigar"]), ptr(arrs["n_gap"]), ptr(arrs["ref_len"]))
    return s, arrs


def count_gaps(r, alloc=None, threads=0):
    """What a packer hands over as csv_reads::n_gap: the number of D / N ops of every record (host helper of the
    library, no device needed).  Returns the array; callers store it as r["n_gap"]."""
    n = int(r["n_reads"])
    out = (alloc or (lambda k, dt: np.empty(k, dtype=dt)))(max(n, 1), np.uint32)
    if n:
        cig = np.ascontiguousarray(r["cigar"], np.uint32); off = np.ascontiguousarray(r["cig_off"], np.uint64)
        lib().csv_host_count_gaps(ptr(cig), ptr(off), n, ptr(out), int(threads))
    return out[:n]


def record_stats(r, alloc=None, threads=0):
    """csv_reads::n_gap and csv_reads::ref_len of every record (csv_host_record_stats): what a packer has at hand while it
    copies the CIGAR words -- D / N ops and reference bases consumed.  Returns (n_gap, ref_len)."""
    n = int(r["n_reads"])
    mk = alloc or (lambda k, dt: np.empty(k, dtype=dt))
    g, l = mk(max(n, 1), np.uint32), mk(max(n, 1), np.uint32)
    if n:
        cig = np.ascontiguousarray(r["cigar"], np.uint32); off = np.ascontiguousarray(r["cig_off"], np.uint64)
        lib().csv_host_record_stats(ptr(cig), ptr(off), n, ptr(g), ptr(l), int(threads))
    return g[:n], l[:n]

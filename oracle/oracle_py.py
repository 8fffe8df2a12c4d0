"""ctypes bindings of the checker libraries.  TEST INFRASTRUCTURE ONLY.

  liboracle.so                 plain-C restatement (oracle/csv_oracle.c)
  _ref/libcontextsv_ref.so     the reference's own sources, unmodified, + htslib shim + harness

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libcontextsv_ref.so")
LIB_REF_O0 = os.path.join(HERE, "_ref", "libcontextsv_ref_O0.so")


class OrcReads(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_ops", C.c_uint64), ("tid", C.c_void_p), ("pos0", C.c_void_p),
                ("flag", C.c_void_p), ("mapq", C.c_void_p), ("cig_off", C.c_void_p), ("cigar", C.c_void_p)]


SIG_DTYPE = np.dtype([("start", np.uint32), ("end", np.uint32), ("read_idx", np.uint32), ("op_idx", np.uint32),
                      ("query_pos", np.uint32), ("kind", np.uint8)], align=True)


class ShimMem(C.Structure):
    _fields_ = [("n_targets", C.c_int32), ("target_name", C.POINTER(C.c_char_p)), ("target_len", C.c_void_p),
                ("n_reads", C.c_uint64), ("tid", C.c_void_p), ("pos0", C.c_void_p), ("flag", C.c_void_p),
                ("mapq", C.c_void_p), ("cig_off", C.c_void_p), ("cigar", C.c_void_p), ("seq4", C.c_void_p),
                ("seq_off", C.c_void_p)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def norm_reads(r):
    """Contiguous arrays of the exact dtypes of the SoA layout."""
    out = dict(r)
    n = int(r["n_reads"])
    out["tid"] = None if r.get("tid") is None else np.ascontiguousarray(r["tid"], np.int32)
    out["pos0"] = np.ascontiguousarray(r["pos0"], np.int32)
    out["flag"] = np.ascontiguousarray(r["flag"], np.uint16)
    out["mapq"] = np.ascontiguousarray(r["mapq"], np.uint8)
    out["cig_off"] = np.ascontiguousarray(r["cig_off"], np.uint64)
    out["cigar"] = np.ascontiguousarray(r["cigar"], np.uint32)
    out["n_ops"] = int(out["cig_off"][n])
    return out


def make_reads(pos0, cigars, tid=None, flag=None, mapq=None):
    """Build the SoA from python lists: cigars = list of lists of (len, op)."""
    n = len(pos0)
    off = np.zeros(n + 1, np.uint64)
    words = []
    for i, c in enumerate(cigars):
        off[i + 1] = off[i] + len(c)
        words += [(int(l) << 4) | int(o) for l, o in c]
    return norm_reads({
        "n_reads": n, "tid": None if tid is None else np.asarray(tid, np.int32),
        "pos0": np.asarray(pos0, np.int32),
        "flag": np.zeros(n, np.uint16) if flag is None else np.asarray(flag, np.uint16),
        "mapq": np.full(n, 60, np.uint8) if mapq is None else np.asarray(mapq, np.uint8),
        "cig_off": off, "cigar": np.asarray(words, np.uint32)})


def _orc_struct(r):
    s = OrcReads(int(r["n_reads"]), int(r["n_ops"]), _p(r["tid"]), _p(r["pos0"]), _p(r["flag"]), _p(r["mapq"]),
                 _p(r["cig_off"]), _p(r["cigar"]))
    return s


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        if not os.path.exists(LIB_ORACLE) or os.path.getmtime(LIB_ORACLE) < os.path.getmtime(os.path.join(HERE, "csv_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
        self.lib = C.CDLL(LIB_ORACLE)
        L = self.lib
        L.orc_cigar_scan.restype = C.c_uint64
        L.orc_cigar_scan_fast.restype = C.c_uint64
        L.orc_largest_cluster.restype = C.c_uint64
        L.orc_mean_cov.restype = C.c_double
        L.orc_mean_cov.argtypes = [C.c_uint64, C.c_uint32]

    def depth(self, r, tid, map_size):
        r = norm_reads(r)
        d = np.zeros(map_size, np.uint32)
        s = C.c_uint64(0); nz = C.c_uint32(0)
        st = _orc_struct(r)
        self.lib.orc_depth(C.byref(st), C.c_int32(tid), C.c_uint32(map_size), _p(d), C.byref(s), C.byref(nz))
        return d, int(s.value), int(nz.value)

    def mean_cov(self, s, nz):
        return float(self.lib.orc_mean_cov(s, nz))

    def cigar_scan(self, r, tid, depth_map_size, min_len=50, min_mapq=20, fast=True):
        r = norm_reads(r)
        st = _orc_struct(r)
        fn = self.lib.orc_cigar_scan_fast if fast else self.lib.orc_cigar_scan
        args = [C.byref(st), C.c_int32(tid), C.c_uint32(min_len), C.c_uint8(min_mapq), C.c_uint32(depth_map_size)]
        n = int(fn(*args, None, C.c_uint64(0)))
        out = np.zeros(max(n, 1), SIG_DTYPE)
        fn(*args, _p(out), C.c_uint64(n))
        return out[:n]

    def record_summary(self, r):
        r = norm_reads(r)
        st = _orc_struct(r)
        n = int(r["n_reads"])
        e = np.zeros(n, np.int32); s = np.zeros(n, np.int32); q = np.zeros(n, np.int32)
        self.lib.orc_record_summary(C.byref(st), _p(e), _p(s), _p(q))
        return e, s, q

    def dbscan1d(self, pts, eps, min_pts, fast=False):
        pts = np.ascontiguousarray(pts, np.int32)
        lab = np.zeros(len(pts), np.int32)
        fn = self.lib.orc_dbscan1d_fast if fast else self.lib.orc_dbscan1d
        fn(_p(pts), C.c_uint64(len(pts)), C.c_double(eps), C.c_int(min_pts), _p(lab))
        return lab

    def dbscan2d(self, start, end, eps, min_pts):
        start = np.ascontiguousarray(start, np.uint32); end = np.ascontiguousarray(end, np.uint32)
        lab = np.zeros(len(start), np.int32)
        self.lib.orc_dbscan2d(_p(start), _p(end), C.c_uint64(len(start)), C.c_double(eps), C.c_int(min_pts), _p(lab))
        return lab

    def largest_cluster(self, pts, labels):
        pts = np.ascontiguousarray(pts, np.int32); labels = np.ascontiguousarray(labels, np.int32)
        out = np.zeros(max(len(pts), 1), np.int32)
        m = int(self.lib.orc_largest_cluster(_p(pts), _p(labels), C.c_uint64(len(pts)), _p(out)))
        return out[:m]

    def log2_windows(self, depth, start, end, sample_size, mean_cov):
        depth = np.ascontiguousarray(depth, np.uint32)
        ws = np.zeros(sample_size, np.uint32); we = np.zeros(sample_size, np.uint32)
        su = np.zeros(sample_size, np.uint64); cn = np.zeros(sample_size, np.uint32); lg = np.zeros(sample_size, np.float64)
        self.lib.orc_log2_windows(_p(depth), C.c_uint64(len(depth)), C.c_uint32(start), C.c_uint32(end), C.c_int(sample_size),
                                  C.c_double(mean_cov), _p(ws), _p(we), _p(su), _p(cn), _p(lg))
        return ws, we, su, cn, lg


def ref_available(o0=False):
    return os.path.exists(LIB_REF_O0 if o0 else LIB_REF)


class Reference:
    """The reference's own code (oracle/_ref), reached through ref_harness.cpp."""

    def __init__(self, o0=False):
        path = LIB_REF_O0 if o0 else LIB_REF
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
        self.lib = C.CDLL(path)
        self.lib.ref_cigar_scan.restype = C.c_int64
        self.lib.ref_largest_cluster.restype = C.c_uint64
        if hasattr(self.lib, "ref_split_dump"):
            self.lib.ref_split_dump.restype = C.c_int64
            self.lib.ref_split_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p]

    def split_dump(self, bam_path, out_path, chrom="", threads=1):
        """SVCaller::findSplitSVSignatures (sv_caller.cpp:68-504) on a BAM file; candidates appended to out_path, one
        line each (chr, start, end, SVType, ALT, evidence bits, aln_offset, cluster_size).  Returns their number."""
        n = int(self.lib.ref_split_dump(bam_path.encode(), chrom.encode(), int(threads), out_path.encode()))
        if n < 0:
            raise RuntimeError("ref_split_dump failed")
        return n

    def _mem(self, r, contig_len, seq4=None, seq_off=None):
        r = norm_reads(r)
        names = [("chr%d" % i).encode() for i in range(len(contig_len))]
        arr = (C.c_char_p * len(names))(*names)
        tl = np.ascontiguousarray(contig_len, np.uint32)
        m = ShimMem(len(names), arr, _p(tl), int(r["n_reads"]), _p(r["tid"]), _p(r["pos0"]), _p(r["flag"]), _p(r["mapq"]),
                    _p(r["cig_off"]), _p(r["cigar"]), _p(seq4), _p(seq_off))
        keep = (r, names, arr, tl, seq4, seq_off)
        return m, keep

    def depth(self, r, tid, contig_len, alloc_size=None):
        m, keep = self._mem(r, contig_len)
        size = int(contig_len[tid]) + 1
        d = np.zeros(size, np.uint32)
        s = C.c_uint64(0); nz = C.c_uint32(0); mean = C.c_double(0)
        rc = self.lib.ref_depth(C.byref(m), C.c_int32(tid), C.c_uint32(size if alloc_size is None else alloc_size), _p(d),
                                C.byref(s), C.byref(nz), C.byref(mean))
        assert rc == 0
        return d, int(s.value), int(nz.value), float(mean.value)

    def cigar_scan(self, r, tid, contig_len, depth_map_size=None, min_mapq=20, seq4=None, seq_off=None):
        m, keep = self._mem(r, contig_len, seq4, seq_off)
        dms = int(contig_len[tid]) + 1 if depth_map_size is None else depth_map_size
        cap = 1 << 16
        while True:
            st = np.zeros(cap, np.uint32); en = np.zeros(cap, np.uint32); ty = np.zeros(cap, np.int32)
            ev = np.zeros(cap, np.uint32); alt = np.zeros(cap * 64, np.uint8)
            n = int(self.lib.ref_cigar_scan(C.byref(m), C.c_int32(tid), C.c_uint32(dms), C.c_int(min_mapq), _p(st), _p(en), _p(ty),
                                            _p(ev), _p(alt), C.c_uint64(cap)))
            if n <= cap:
                break
            cap = n
        alts = [bytes(alt[64 * i: 64 * i + 64]).split(b"\0")[0].decode() for i in range(n)]
        return st[:n], en[:n], ty[:n], ev[:n], alts

    def record_summary(self, r, tid, contig_len):
        """(bam_endpos, query_start, query_end) of the records sam_itr_querys yields for contig tid, in file order."""
        m, keep = self._mem(r, contig_len)
        cap = int(norm_reads(r)["n_reads"]) + 1
        e = np.zeros(cap, np.int32); s = np.zeros(cap, np.int32); q = np.zeros(cap, np.int32)
        self.lib.ref_record_summary.restype = C.c_int64
        n = int(self.lib.ref_record_summary(C.byref(m), C.c_int32(tid), _p(e), _p(s), _p(q), C.c_uint64(cap)))
        return e[:n], s[:n], q[:n]

    def dbscan1d(self, pts, eps, min_pts):
        pts = np.ascontiguousarray(pts, np.int32)
        lab = np.zeros(len(pts), np.int32)
        self.lib.ref_dbscan1d(_p(pts), C.c_uint64(len(pts)), C.c_double(eps), C.c_int(min_pts), _p(lab))
        return lab

    def largest_cluster(self, pts, eps, min_pts):
        pts = np.ascontiguousarray(pts, np.int32)
        out = np.zeros(max(len(pts), 1), np.int32)
        m = int(self.lib.ref_largest_cluster(_p(pts), C.c_uint64(len(pts)), C.c_double(eps), C.c_int(min_pts), _p(out)))
        return out[:m]

    def log2_windows(self, depth, start, end, sample_size, mean_cov):
        depth = np.ascontiguousarray(depth, np.uint32)
        cap = sample_size + 8
        pos = np.zeros(cap, np.uint32); lg = np.zeros(cap, np.float64)
        n = int(self.lib.ref_log2_windows(_p(depth), C.c_uint64(len(depth)), C.c_uint32(start), C.c_uint32(end), C.c_int(sample_size),
                                          C.c_double(mean_cov), _p(pos), _p(lg), C.c_int(cap)))
        return pos[:n], lg[:n]

    def dbscan2d(self, start, end, eps, min_pts):
        start = np.ascontiguousarray(start, np.uint32); end = np.ascontiguousarray(end, np.uint32)
        lab = np.zeros(len(start), np.int32)
        self.lib.ref_dbscan2d(_p(start), _p(end), C.c_uint64(len(start)), C.c_double(eps), C.c_int(min_pts), _p(lab))
        return lab

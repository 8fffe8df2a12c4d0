"""Host-side mirror of the reference interfaces for the alignment-scan path.

Same names, argument meaning and error behaviour as the reference's C++ classes
(the reference has no Python or FFI surface for this path):

  DBSCAN1D(epsilon, minPts).fit / getClusters / getLargestCluster   include/dbscan1d.h:11-32
  CNVCaller.calculateMeanChromosomeCoverage                         include/cnv_caller.h:104
  CNVCaller.queryLog2Windows  (the window part of querySNPRegion)    src/cnv_caller.cpp:76-113
  SVCaller.findCIGARSVs / processChromosome                         include/sv_caller.h:84-86

The one deliberate difference: where the reference takes a BAM path and decodes it
with htslib, these take an `Alignments` object -- the packed SoA the host packer
produces from the same records (csv_reads in include/contextsv_b200.h).
Everything computes on the GPU through the C ABI; nothing here falls back to the CPU.
"""
import bisect
import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import CsvError, CsvRegion, CsvScanParams, CsvSigs, check, lib, ptr

KIND_CIGARINS, KIND_CIGARDEL, KIND_CIGARCLIP = 0, 1, 2
SEQ_NT16 = "=ACMGRSVTWYHKDBN"
_AMBIGUOUS = set("RYKMSWBDHV")


class Context:
    """csv_ctx: one CUDA stream + scratch on one device.  One per host thread."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().csv_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib().csv_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().csv_ctx_sync(self.h))

    def timer_begin(self):
        check(lib().csv_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        check(lib().csv_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def set_pipeline_chunks(self, n):
        """Batches uploaded from now on are scanned in up to n pipelined chunks of whole contigs."""
        check(lib().csv_ctx_set_pipeline_chunks(self.h, int(n)))

    def set_fetch(self, threads=-1, chunk_positions=-1, exception_slots=-1, min_positions=-1):
        """How depth maps come back (csv_ctx_set_fetch): `threads` host threads widen a byte-wide transfer into the
        caller's uint32 array; 0 = plain 32-bit DMA.  -1 keeps a value."""
        check(lib().csv_ctx_set_fetch(self.h, int(threads), int(chunk_positions), int(exception_slots), int(min_positions)))

    def fetch_stats(self):
        """(chunks fetched narrow, chunks re-fetched as plain words) since the context was created."""
        a = C.c_uint64(0); b = C.c_uint64(0)
        check(lib().csv_ctx_fetch_stats(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def profile_enable(self, on=True):
        check(lib().csv_profile_enable(self.h, int(on)))

    def profile_read(self, reset=True):
        """{stage: (total_ms, calls)} accumulated while profiling was enabled."""
        names = (C.c_char_p * 16)(); ms = (C.c_double * 16)(); calls = (C.c_uint32 * 16)()
        n = lib().csv_profile_read(self.h, 16, names, ms, calls, int(reset))
        if n < 0:
            check(-n)
        return {names[i].decode(): (float(ms[i]), int(calls[i])) for i in range(n)}

    @property
    def launches(self):
        return int(lib().csv_ctx_launch_count(self.h))


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


@dataclass
class Alignments:
    """Packed records of one BAM (csv_reads) + the header fields the path needs."""
    reads: dict
    contig_names: list
    contig_len: list
    seq4: np.ndarray = None      # optional 4-bit packed bases (BAM encoding), concatenated per record
    seq_off: np.ndarray = None   # optional [n_reads] byte offsets into seq4

    def tid_of(self, chrom):
        return self.contig_names.index(chrom) if chrom in self.contig_names else -1

    def base(self, read_idx, qpos):
        """seq_nt16_str[bam_seqi(seq, qpos)] of a record (sv_caller.cpp:575)."""
        if self.seq4 is None:
            raise ValueError("Alignments carries no sequences")
        b = int(self.seq4[int(self.seq_off[read_idx]) + (qpos >> 1)])
        return SEQ_NT16[(b >> ((~qpos & 1) << 2)) & 0xF]


def whole_contig_regions(contig_len, tids=None):
    tids = range(len(contig_len)) if tids is None else tids
    return [(int(t), 0, int(contig_len[t]) + 1, int(contig_len[t]) + 1) for t in tids]


# The packer's share of the pre-pass: when the records come without csv_reads::n_gap (D / N ops per record), Batch
# counts them on the host before the upload (csv_host_count_gaps).  False leaves the field NULL, and the device falls
# back to its op-level pre-pass (a second read of every CIGAR word).
COUNT_GAPS = True


class Batch:
    """csv_batch: reads resident in HBM + the regions they are scanned against."""

    def __init__(self, ctx, reads, regions):
        self.ctx = ctx
        self.regions = [tuple(int(x) for x in r) for r in regions]
        if COUNT_GAPS and reads.get("n_gap") is None and int(reads["n_reads"]):
            reads = dict(reads)
            reads["n_gap"], reads["ref_len"] = _capi.record_stats(reads)
        rs, self._keep = _capi.reads_struct(reads)
        arr = (CsvRegion * len(self.regions))(*[CsvRegion(*r) for r in self.regions])
        h = C.c_void_p()
        check(lib().csv_batch_upload(ctx.h, C.byref(rs), len(self.regions), arr, C.byref(h)))
        self.h = h
        self.n_reads = int(reads["n_reads"])

    def free(self):
        if self.h:
            lib().csv_batch_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def scan(self, want_depth=True, want_sigs=True, min_len=50, min_mapq=20):
        p = CsvScanParams(min_len, min_mapq, int(want_depth), int(want_sigs), 0)
        self._last_scan = p
        check(lib().csv_scan_run(self.ctx.h, self.h, C.byref(p)))

    def release_inputs(self):
        """Keeps only the results (depth slabs, signature columns, labels): csv_batch_release_inputs."""
        check(lib().csv_batch_release_inputs(self.ctx.h, self.h))

    def depth_stats(self):
        n = len(self.regions)
        s = np.zeros(n, np.uint64); nz = np.zeros(n, np.uint32)
        check(lib().csv_depth_stats(self.ctx.h, self.h, ptr(s), ptr(nz)))
        return s, nz

    def depth(self, region, out=None):
        _, beg, end, _ = self.regions[region]
        if out is None:
            out = np.empty(end - beg, np.uint32)
        assert out.dtype == np.uint32 and out.size == end - beg and out.flags.c_contiguous
        check(lib().csv_depth_fetch(self.ctx.h, self.h, region, ptr(out)))
        return out

    def depth_all(self, out=None):
        """Every region's slice through one fetch pipeline (csv_depth_fetch_all).  out: list of uint32 arrays or None."""
        if out is None:
            out = [np.empty(e - b, np.uint32) for (_, b, e, _) in self.regions]
        assert len(out) == len(self.regions)
        for o, (_, b, e, _) in zip(out, self.regions):
            assert o.dtype == np.uint32 and o.size == e - b and o.flags.c_contiguous
        ptrs = (C.c_void_p * len(out))(*[o.ctypes.data for o in out])
        check(lib().csv_depth_fetch_all(self.ctx.h, self.h, ptrs))
        return out

    def sigs_count(self):
        """Signatures of the last pass.  A pass that emitted more than the batch had reserved is run again with the
        reported capacity (csv_batch_reserve_sigs): the reference's vector has no limit."""
        n = C.c_uint64(0)
        rc = lib().csv_sigs_count(self.ctx.h, self.h, C.byref(n))
        if rc == _capi.CSV_ERR_CAPACITY and int(n.value) > 0:
            check(lib().csv_batch_reserve_sigs(self.ctx.h, self.h, int(n.value)))
            check(lib().csv_scan_run(self.ctx.h, self.h, C.byref(self._last_scan)))
            rc = lib().csv_sigs_count(self.ctx.h, self.h, C.byref(n))
        check(rc)
        return int(n.value)

    def sigs(self, out=None, n=None):
        """dict of arrays in the reference's vector order + 'region_off' [n_regions + 1].  `out` may hold preallocated
        arrays (start, end, kind, read_idx, op_idx, query_pos) of enough capacity: results land there (views returned)."""
        n = self.sigs_count() if n is None else n
        cap = max(n, 1)
        if out is None:
            o = {"start": np.zeros(cap, np.uint32), "end": np.zeros(cap, np.uint32), "kind": np.zeros(cap, np.uint8),
                 "read_idx": np.zeros(cap, np.uint32), "op_idx": np.zeros(cap, np.uint32), "query_pos": np.zeros(cap, np.uint32)}
        else:
            o = {k: out[k] for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")}
            cap = min(len(v) for v in o.values())
        st = CsvSigs(ptr(o["start"]), ptr(o["end"]), ptr(o["kind"]), ptr(o["read_idx"]), ptr(o["op_idx"]), ptr(o["query_pos"]))
        got = C.c_uint64(0)
        off = np.zeros(len(self.regions) + 1, np.uint64)
        check(lib().csv_sigs_fetch(self.ctx.h, self.h, C.byref(st), cap, C.byref(got), ptr(off)))
        o = {k: v[:n] for k, v in o.items()}
        o["region_off"] = off
        return o

    def sigs_dbscan1d(self, eps, min_pts, fetch=True, out=None, n=None):
        if not fetch:
            check(lib().csv_sigs_dbscan1d(self.ctx.h, self.h, float(eps), int(min_pts), None, 0))
            return None
        n = self.sigs_count() if n is None else n
        lab = np.zeros(max(n, 1), np.int32) if out is None else out
        check(lib().csv_sigs_dbscan1d(self.ctx.h, self.h, float(eps), int(min_pts), ptr(lab), len(lab)))
        return lab[:n]

    def record_summary(self):
        """(bam_endpos, query_start, query_end) of every record: what the split-read pass reads off each alignment
        (sv_caller.cpp:150-162, 663-690)."""
        e = np.zeros(self.n_reads, np.int32); s = np.zeros(self.n_reads, np.int32); q = np.zeros(self.n_reads, np.int32)
        check(lib().csv_record_summary(self.ctx.h, self.h, ptr(e), ptr(s), ptr(q)))
        return e, s, q

    def depth_at(self, region, positions):
        """SVCaller::getReadDepth for many positions, from the device-resident map (0 beyond it)."""
        pos = np.ascontiguousarray(positions, np.uint32)
        out = np.zeros(len(pos), np.uint32)
        check(lib().csv_depth_at(self.ctx.h, self.h, region, len(pos), ptr(pos), ptr(out)))
        return out

    def depth_at_tid(self, tid, positions):
        """getReadDepth for (contig, position) pairs anywhere in the batch; 0xffffffff = not covered by this batch's regions."""
        t = np.ascontiguousarray(tid, np.int32); pos = np.ascontiguousarray(positions, np.uint32)
        out = np.zeros(len(pos), np.uint32)
        check(lib().csv_depth_at_tid(self.ctx.h, self.h, len(pos), ptr(t), ptr(pos), ptr(out)))
        return out

    def sigs_depth(self, n=None, out=None):
        """Depth at the start of every signature, in sigs() order (sv_caller.cpp:1306)."""
        n = self.sigs_count() if n is None else n
        out = np.zeros(max(n, 1), np.uint32) if out is None else out
        check(lib().csv_sigs_depth(self.ctx.h, self.h, ptr(out), len(out)))
        return out[:n]

    def depth_checksum(self):
        """Additive position-weighted checksum of every region's depth slice (csv_depth_checksum)."""
        out = np.zeros(len(self.regions), np.uint64)
        check(lib().csv_depth_checksum(self.ctx.h, self.h, ptr(out)))
        return out

    def debug_array(self, name, dtype=np.uint32):
        """Diagnostics: one of the batch's intermediate device arrays (csv_debug_fetch)."""
        import ctypes as C
        size = C.c_uint64(0)
        check(lib().csv_debug_fetch(self.ctx.h, self.h, name.encode(), 0, 0, None, C.byref(size)))
        out = np.empty(size.value // np.dtype(dtype).itemsize, dtype)
        check(lib().csv_debug_fetch(self.ctx.h, self.h, name.encode(), 0, out.nbytes, ptr(out), None))
        return out

    def window_sums(self, region, start_pos, end_pos, sample_size):
        s = np.ascontiguousarray(start_pos, np.uint32); e = np.ascontiguousarray(end_pos, np.uint32)
        su = np.zeros(len(s) * sample_size, np.uint64); cn = np.zeros(len(s) * sample_size, np.uint32)
        check(lib().csv_window_sums(self.ctx.h, self.h, region, len(s), ptr(s), ptr(e), int(sample_size), ptr(su), ptr(cn)))
        return su.reshape(len(s), sample_size), cn.reshape(len(s), sample_size)


def scan_streamed(ctx, reads, contig_len, max_ops=(1 << 31) - (1 << 20), eps=None, min_pts=5, min_len=50, min_mapq=20):
    """The whole hot path over records that do not fit one batch: shards planned by CIGAR-op budget
    (shard.plan_by_ops), scanned one after the other on one device, results merged on the host exactly like
    the shards of a multi-GPU run.  Returns (depth per contig, (sum, nonzero) per contig,
    signatures per contig in the reference's vector order, DBSCAN1D labels per contig or None)."""
    from . import shard
    ends = shard.ref_end(reads)
    plans = shard.plan_by_ops(reads, contig_len, max_ops, ends)
    n_contigs = len(contig_len)
    depth = [np.zeros(int(l) + 1, np.uint32) for l in contig_len]
    sums = np.zeros(n_contigs, np.uint64); nzs = np.zeros(n_contigs, np.uint64)
    parts = []
    for regions in plans:
        sub, base = shard.select_reads(reads, regions, ends)
        b = Batch(ctx, sub, regions)
        b.scan(want_depth=True, want_sigs=True, min_len=min_len, min_mapq=min_mapq)
        s, nz = b.depth_stats()
        for i, (t, beg, end, _) in enumerate(regions):
            b.depth(i, out=depth[t][beg:end])
            sums[t] += s[i]; nzs[t] += nz[i]
        parts.append((b.sigs(), regions, base))
        b.free()
    sigs = shard.merge_signatures(parts)
    labels = None
    if eps is not None:
        labels = {}
        for t, sg in sigs.items():
            seg = (sg["kind"] != 1).astype(np.uint32)                  # DEL | INS groups (sv_object.cpp:61-83)
            labels[t] = dbscan1d_segments(sg["start"].astype(np.int32), seg, 2, eps, min_pts, ctx)[0] if len(seg) else np.zeros(0, np.int32)
    return depth, (sums, nzs), sigs, labels


# --------------------------------------------------------------------------- DBSCAN1D

class DBSCAN1D:
    """include/dbscan1d.h:11-32.  Labels: cluster id >= 0, -2 noise."""

    def __init__(self, epsilon, minPts, ctx=None):
        self.epsilon = float(epsilon)
        self.minPts = int(minPts)
        self.clusters = np.zeros(0, np.int32)
        self._ctx = ctx

    def fit(self, points):
        ctx = self._ctx or default_context()
        pts = np.ascontiguousarray(points, np.int32)
        lab = np.zeros(len(pts), np.int32)
        check(lib().csv_dbscan1d(ctx.h, ptr(pts), len(pts), self.epsilon, self.minPts, ptr(lab), None))
        self.clusters = lab

    def getClusters(self):
        return self.clusters

    def getLargestCluster(self, points):
        pts = np.ascontiguousarray(points, np.int32)
        lab = np.ascontiguousarray(self.clusters, np.int32)
        out = np.zeros(max(len(pts), 1), np.int32)
        m = int(lib().csv_largest_cluster(ptr(pts), ptr(lab), len(pts), ptr(out)))
        return out[:m]


class DBSCAN:
    """include/dbscan.h:11-33: the 2-D DBSCAN mergeSVs runs on SV calls (intervals, reciprocal-overlap distance).
    fit() takes the calls' (start, end) columns instead of a vector<SVCall>."""

    def __init__(self, epsilon, minPts, ctx=None):
        self.epsilon = float(epsilon)
        self.minPts = int(minPts)
        self.clusters = np.zeros(0, np.int32)
        self._ctx = ctx

    def fit(self, start, end):
        ctx = self._ctx or default_context()
        s = np.ascontiguousarray(start, np.uint32); e = np.ascontiguousarray(end, np.uint32)
        assert len(s) == len(e)
        lab = np.zeros(len(s), np.int32)
        check(lib().csv_dbscan2d(ctx.h, ptr(s), ptr(e), len(s), self.epsilon, self.minPts, ptr(lab)))
        self.clusters = lab

    def getClusters(self):
        return self.clusters


def dbscan1d_segments(points, seg_id, n_seg, eps, min_pts, ctx=None):
    """Many independent DBSCAN1D fits in one launch sequence (csv_dbscan1d_seg)."""
    ctx = ctx or default_context()
    pts = np.ascontiguousarray(points, np.int32)
    seg = None if seg_id is None else np.ascontiguousarray(seg_id, np.uint32)
    lab = np.zeros(len(pts), np.int32); nc = np.zeros(max(n_seg, 1), np.int32)
    check(lib().csv_dbscan1d_seg(ctx.h, ptr(pts), ptr(seg), len(pts), n_seg, float(eps), int(min_pts), ptr(lab), ptr(nc)))
    return lab, nc


# --------------------------------------------------------------------------- CNVCaller

class CNVCaller:
    """The depth part of include/cnv_caller.h."""

    def __init__(self, ctx=None):
        self._ctx = ctx
        self.last_batch = None

    def calculateMeanChromosomeCoverage(self, chromosomes, chr_pos_depth_map, chr_mean_cov_map, alignments, thread_count=1,
                                        printError=None):
        """cnv_caller.cpp:415-556.  chr_pos_depth_map[chr] must already hold the caller-allocated
        uint32 array (sv_caller.cpp:801); it is replaced by an array of target_len+1 entries when
        the size differs (cnv_caller.cpp:482-487).  Means are stored only when non-zero (:540)."""
        ctx = self._ctx or default_context()
        err = printError or (lambda m: None)
        tids, chroms = [], []
        for chrom in chromosomes:
            tid = alignments.tid_of(chrom)
            if tid < 0:
                err("ERROR: Could not create iterator for chromosome: " + chrom + ", check if the chromosome exists in the BAM file.")
                continue
            tids.append(tid); chroms.append(chrom)
        if not tids:
            return
        regions = whole_contig_regions(alignments.contig_len, tids)
        batch = Batch(ctx, alignments.reads, regions)
        batch.scan(want_depth=True, want_sigs=False)
        sums, nzs = batch.depth_stats()
        outs = []
        for i, chrom in enumerate(chroms):
            size = regions[i][3]
            cur = chr_pos_depth_map.get(chrom)
            if cur is None or len(cur) != size:
                err("ERROR: Chromosome length mismatch for %s: expected %d, found %d, resizing to %d"
                    % (chrom, size, 0 if cur is None else len(cur), size))
                cur = np.zeros(size, np.uint32)
            outs.append(np.ascontiguousarray(cur, np.uint32))
        batch.depth_all(outs)                                                 # one fetch pipeline over all contigs
        for i, chrom in enumerate(chroms):
            chr_pos_depth_map[chrom] = outs[i]
            mean = float(sums[i]) / float(nzs[i]) if nzs[i] > 0 else 0.0     # :538  (uint64 -> double, uint32 -> double)
            if mean != 0.0:
                chr_mean_cov_map[chrom] = mean
        self.last_batch = batch

    @staticmethod
    def log2_from_sums(win_sum, win_count, mean_chr_cov):
        """cnv_caller.cpp:100-108 from the integer window sums."""
        out = np.zeros(len(win_sum), np.float64)
        for i in range(len(win_sum)):
            if win_count[i] > 0:
                cov = float(win_sum[i])
                if cov == 0:
                    cov = 1e-9
                out[i] = math.log2((cov / float(int(win_count[i]))) / mean_chr_cov)
        return out


# --------------------------------------------------------------------------- SVCaller

@dataclass
class SVCall:
    """include/sv_object.h:16-35 (fields the CIGAR path sets)."""
    start: int
    end: int
    sv_type: str
    alt_allele: str = "."
    aln_type: frozenset = field(default_factory=frozenset)
    genotype: str = "./."
    hmm_likelihood: float = 0.0
    cn_state: int = 0
    aln_offset: int = 0
    cluster_size: int = 0

    def key(self):
        return (self.start, self.end)


_KIND_EVIDENCE = {0: "CIGARINS", 1: "CIGARDEL", 2: "CIGARCLIP"}


class SVCaller:
    """The CIGAR part of include/sv_caller.h."""

    min_mapq = 20      # sv_caller.h:72

    def __init__(self, ctx=None):
        self._ctx = ctx

    def _alt(self, alignments, kind, start, end, read_idx, query_pos):
        if kind == KIND_CIGARDEL:
            return "<DEL>"
        op_len = end - start + 1
        if op_len <= 50 and alignments.seq4 is not None:            # sv_caller.cpp:587-591
            s = []
            for j in range(op_len):
                b = alignments.base(read_idx, query_pos + j)
                s.append("N" if b in _AMBIGUOUS else b)             # :572-581
            return "".join(s)
        return "<INS>"

    def findCIGARSVs(self, alignments, region, sv_calls, pos_depth_map):
        """sv_caller.cpp:506-537: appends this contig's CIGAR signatures to sv_calls through the
        reference's sorted insert (sv_object.cpp:22-33).  pos_depth_map is only asked for its size."""
        ctx = self._ctx or default_context()
        tid = alignments.tid_of(region)
        if tid < 0:
            return
        size = len(pos_depth_map)
        batch = Batch(ctx, alignments.reads, [(tid, 0, size, size)])
        batch.scan(want_depth=False, want_sigs=True, min_len=50, min_mapq=self.min_mapq)
        s = batch.sigs()
        batch.free()
        new = []
        for i in range(len(s["start"])):
            kind = int(s["kind"][i]); st = int(s["start"][i]); en = int(s["end"][i])
            alt = self._alt(alignments, kind, st, en, int(s["read_idx"][i]), int(s["query_pos"][i]))
            new.append(SVCall(st, en, "DEL" if kind == KIND_CIGARDEL else "INS", alt, frozenset([_KIND_EVIDENCE[kind]])))
        if not sv_calls:
            sv_calls.extend(new)
            return
        # non-empty target: replay addSVCall's lower_bound insert in insertion order
        order = sorted(range(len(new)), key=lambda i: (int(s["read_idx"][i]), int(s["op_idx"][i])))
        keys = [c.key() for c in sv_calls]
        for i in order:
            k = new[i].key()
            p = bisect.bisect_left(keys, k)
            keys.insert(p, k); sv_calls.insert(p, new[i])

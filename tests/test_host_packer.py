"""CPU: the C++ host side above the C ABI (contextsv_b200/csrc/host/packed_reads.*): htslib iterator -> packed SoA,
halo selection for streamed shards, and the depth-pass -> CIGAR-pass packing cache.  Built against the htslib shim
(no CUDA, no reference sources) and driven through ctypes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from contextsv_b200 import bamio, build, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(build.CSRC, "host")
SHIM = os.path.join(ROOT, "oracle", "htslib_shim")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("hp")
    so = str(d / "libhp.so")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + SHIM, "-I" + os.path.join(ROOT, "oracle"), "-I" + build.INC, "-I" + HOST,
           os.path.join(ROOT, "tests", "native", "host_packer_harness.cpp"), os.path.join(HOST, "packed_reads.cpp"),
           os.path.join(SHIM, "shim.cpp"), "-o", so, "-lz", "-lpthread"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    old = os.environ.get("CONTEXTSV_CACHE_OPS")
    os.environ["CONTEXTSV_CACHE_OPS"] = "60000"          # read once, when the library's cache is constructed (at load)
    try:
        L = C.CDLL(so)
    finally:
        if old is None:
            del os.environ["CONTEXTSV_CACHE_OPS"]
        else:
            os.environ["CONTEXTSV_CACHE_OPS"] = old
    L.hp_pack.restype = C.c_void_p; L.hp_pack.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.hp_free.argtypes = [C.c_void_p]
    L.hp_size.restype = C.c_uint64; L.hp_size.argtypes = [C.c_void_p]
    L.hp_ops.restype = C.c_uint64; L.hp_ops.argtypes = [C.c_void_p]
    L.hp_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.hp_seq_count.restype = C.c_uint64; L.hp_seq_count.argtypes = [C.c_void_p]
    L.hp_bases.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p]
    L.hp_keep_reaching.argtypes = [C.c_void_p, C.c_uint32]
    L.hp_cache_put.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    L.hp_cache_take.restype = C.c_void_p; L.hp_cache_take.argtypes = [C.c_char_p, C.c_int]
    L.hp_file_name.restype = C.c_char_p; L.hp_file_name.argtypes = [C.c_char_p]
    L.hp_max_ops.restype = C.c_uint64
    return L


@pytest.fixture(scope="module")
def bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("bam")
    clen, names = [120_000, 60_000], ["chr21", "chr22"]
    r = synth.generate(clen, seed=7, n_sv=40, coverage=15.0, frac_len50=0.3)
    path = str(d / "x.bam")
    bamio.write_bam(path, r, names, clen, seed=3)
    return path, r, names, clen


def unpack(L, p):
    n, no = int(L.hp_size(p)), int(L.hp_ops(p))
    a = {"tid": np.zeros(n, np.int32), "pos0": np.zeros(n, np.int32), "flag": np.zeros(n, np.uint16), "mapq": np.zeros(n, np.uint8),
         "cig_off": np.zeros(n + 1, np.uint64), "cigar": np.zeros(max(no, 1), np.uint32), "ref_end": np.zeros(n, np.uint32)}
    L.hp_copy(p, *[a[k].ctypes.data_as(C.c_void_p) for k in ("tid", "pos0", "flag", "mapq", "cig_off", "cigar", "ref_end")])
    a["cigar"] = a["cigar"][:no]
    return a


def contig_slice(r, t):
    tid = np.asarray(r["tid"])
    i0, i1 = int(np.searchsorted(tid, t, "left")), int(np.searchsorted(tid, t, "right"))
    off = np.asarray(r["cig_off"]).astype(np.int64)
    return i0, i1, off


def test_packer_reproduces_the_soa(harness, bam):
    """sam_itr_querys(chrom) + PackedReads::append give back, record for record, the SoA the BAM was written from."""
    L = harness
    path, r, names, clen = bam
    ends = shard.ref_end(r)
    for t, name in enumerate(names):
        p = L.hp_pack(path.encode(), name.encode(), 1)
        assert p
        a = unpack(L, p)
        i0, i1, off = contig_slice(r, t)
        assert len(a["pos0"]) == i1 - i0 > 30
        for k in ("tid", "pos0", "flag", "mapq"):
            assert np.array_equal(a[k], np.asarray(r[k])[i0:i1]), k
        assert np.array_equal(a["cig_off"].astype(np.int64), off[i0:i1 + 1] - off[i0])
        assert np.array_equal(a["cigar"], np.asarray(r["cigar"])[off[i0]:off[i1]])
        assert np.array_equal(a["ref_end"].astype(np.int64), ends[i0:i1])
        # sequences are kept exactly for the records that carry an I / S op of 50 bases (sv_caller.cpp:587-591)
        cig = a["cigar"]; co = a["cig_off"].astype(np.int64)
        want = [i for i in range(len(co) - 1) if np.any(((cig[co[i]:co[i + 1]] >> 4) == 50) & np.isin(cig[co[i]:co[i + 1]] & 15, (1, 4)))]
        assert int(L.hp_seq_count(p)) == len(want) > 0
        buf = C.create_string_buffer(8)
        assert L.hp_bases(p, want[0], 0, 8, buf) == 1 and set(buf.raw[:8].decode()) <= set("ACGTN")
        not_kept = next(i for i in range(len(co) - 1) if i not in set(want))
        assert L.hp_bases(p, not_kept, 0, 1, buf) == 0
        L.hp_free(p)


def test_threaded_bgzf_read_ahead_yields_the_same_records(harness, tmp_path):
    """hts_set_threads(n > 1) in the shim: blocks are read and inflated ahead of the consumer by n threads, seeks restart
    the read-ahead; the packed records must not change (many blocks: ~4 MB of records)."""
    L = harness
    clen, names = [400_000, 250_000, 90_000], ["chr20", "chr21", "chr22"]
    r = synth.generate(clen, seed=5, n_sv=60, coverage=12.0)
    path = str(tmp_path / "t.bam")
    bamio.write_bam(path, r, names, clen, seed=2)
    want = {}
    for name in names:
        p = L.hp_pack(path.encode(), name.encode(), 0)
        want[name] = unpack(L, p); L.hp_free(p)
    try:
        for threads in (2, 5):
            L.hp_set_threads(threads)
            for name in reversed(names):                       # out of file order: every query seeks
                p = L.hp_pack(path.encode(), name.encode(), 0)
                got = unpack(L, p); L.hp_free(p)
                assert len(got["pos0"]) == len(want[name]["pos0"]) > 50
                for k in want[name]:
                    assert np.array_equal(got[k], want[name][k]), (threads, name, k)
    finally:
        L.hp_set_threads(0)


def test_threaded_read_ahead_is_clean_under_thread_sanitizer(tmp_path):
    """The same read-ahead (reader thread + inflating workers + the consumer that packs records) built with
    -fsanitize=thread as a stand-alone program: no data race reported, and the packed records hash to the same value
    with 0, 2 and 6 threads."""
    exe = str(tmp_path / "tsan_packer")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-w", "-I" + SHIM, "-I" + os.path.join(ROOT, "oracle"), "-I" + build.INC, "-I" + HOST,
           os.path.join(ROOT, "tests", "native", "tsan_packer_main.cpp"), os.path.join(ROOT, "tests", "native", "host_packer_harness.cpp"),
           os.path.join(HOST, "packed_reads.cpp"), os.path.join(SHIM, "shim.cpp"), "-o", exe, "-lz", "-lpthread"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0 and "sanitize" in p.stdout:
        pytest.skip("no ThreadSanitizer runtime here")
    assert p.returncode == 0, p.stdout[-3000:]
    clen, names = [300_000, 200_000, 90_000], ["chr20", "chr21", "chr22"]
    r = synth.generate(clen, seed=6, n_sv=50, coverage=12.0)
    path = str(tmp_path / "t.bam")
    bamio.write_bam(path, r, names, clen, seed=2)
    outs = []
    for threads in (0, 2, 6):
        q = subprocess.run([exe, path, str(threads)] + names, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                           env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0"))
        assert q.returncode == 0, (threads, q.stderr[-3000:])
        assert "ThreadSanitizer" not in q.stderr, (threads, q.stderr[-4000:])
        outs.append(q.stdout)
    assert outs[0] == outs[1] == outs[2] and len(outs[0].splitlines()) == len(names)


def test_keep_reaching_is_the_halo_of_the_next_shard(harness, bam):
    L = harness
    path, r, names, clen = bam
    p = L.hp_pack(path.encode(), b"chr21", 0)
    a = unpack(L, p)
    cut = 50_001
    L.hp_keep_reaching(p, cut)
    b = unpack(L, p)
    keep = np.nonzero(a["ref_end"] > cut)[0]
    assert 0 < len(keep) < len(a["pos0"])
    assert np.array_equal(b["pos0"], a["pos0"][keep]) and np.array_equal(b["ref_end"], a["ref_end"][keep])
    co = a["cig_off"].astype(np.int64)
    assert np.array_equal(b["cigar"], np.concatenate([a["cigar"][co[i]:co[i + 1]] for i in keep]))
    assert int(b["cig_off"][-1]) == len(b["cigar"])
    L.hp_free(p)


def test_packing_cache_between_passes(harness, bam):
    """cache_put / cache_take: what the depth pass parks is what the CIGAR pass of the same (file, contig) takes, once;
    the total is bounded by CONTEXTSV_CACHE_OPS (60 000 here)."""
    L = harness
    path, r, names, clen = bam
    assert L.hp_file_name(path.encode()).decode() == path
    p21 = L.hp_pack(path.encode(), b"chr21", 1)
    want = unpack(L, p21)
    n21 = int(L.hp_size(p21)) + int(L.hp_ops(p21))
    assert n21 < 60_000
    assert not L.hp_cache_take(path.encode(), 0)                   # nothing parked yet
    L.hp_cache_put(path.encode(), 0, p21); L.hp_free(p21)          # moved from
    assert not L.hp_cache_take((path + "x").encode(), 0)           # another file
    assert not L.hp_cache_take(path.encode(), 1)                   # another contig
    got = L.hp_cache_take(path.encode(), 0)
    assert got
    a = unpack(L, got)
    for k in want:
        assert np.array_equal(a[k], want[k]), k
    assert not L.hp_cache_take(path.encode(), 0)                   # taken once
    # budget: park chr21 again, then chr22 does not fit beside it any more if the sum exceeds the budget
    L.hp_cache_put(path.encode(), 0, got); L.hp_free(got)
    p22 = L.hp_pack(path.encode(), b"chr22", 1)
    n22 = int(L.hp_size(p22)) + int(L.hp_ops(p22))
    L.hp_cache_put(path.encode(), 1, p22); L.hp_free(p22)
    t22 = L.hp_cache_take(path.encode(), 1)
    assert bool(t22) == (n21 + n22 <= 60_000)
    if t22:
        L.hp_free(t22)
    t21 = L.hp_cache_take(path.encode(), 0)
    assert t21
    L.hp_free(t21)
    assert L.hp_max_ops() == (1 << 31) - (1 << 20) or "CONTEXTSV_MAX_OPS" in os.environ

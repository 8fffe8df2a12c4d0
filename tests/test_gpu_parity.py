"""GPU: the CUDA path, called through the C ABI, against the committed golden vectors (reference's
own outputs) and against the oracle on seeded inputs.  Bit-exact for every integer output."""
import ctypes as C

import os

import numpy as np
import pytest

import util
from util import M, I, D, N, S, H, P, EQ, X
from contextsv_b200 import api, shard
from contextsv_b200._capi import CsvReads, CsvRegion, CsvSigs, check, lib, ptr, reads_struct

pytestmark = pytest.mark.gpu


def run_batch(ctx, r, regions, **kw):
    b = api.Batch(ctx, r, regions)
    b.scan(**kw)
    return b


def check_contigs(ctx, oracle, r, clen, seq=None):
    regions = api.whole_contig_regions(clen)
    b = run_batch(ctx, r, regions)
    sums, nzs = b.depth_stats()
    sg = b.sigs()
    for tid in range(len(clen)):
        d, s, nz = oracle.depth(r, tid, clen[tid] + 1)
        got = b.depth(tid)
        assert np.array_equal(got, d), "depth tid %d: first diff at %s" % (tid, np.nonzero(got != d)[0][:5])
        assert int(sums[tid]) == s and int(nzs[tid]) == nz
        o = oracle.cigar_scan(r, tid, clen[tid] + 1)
        lo, hi = int(sg["region_off"][tid]), int(sg["region_off"][tid + 1])
        assert hi - lo == len(o), "tid %d: %d signatures, oracle %d" % (tid, hi - lo, len(o))
        for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
            assert np.array_equal(sg[f][lo:hi], o[f]), "signature field %s tid %d" % (f, tid)
    b.free()


def test_golden_dbscan1d(ctx):
    for i, pts, eps, mp, labels, largest in util.golden_db_cases():
        db = api.DBSCAN1D(eps, mp, ctx)
        db.fit(pts)
        assert np.array_equal(db.getClusters(), labels), "case %d eps=%g minPts=%d" % (i, eps, mp)
        assert np.array_equal(db.getLargestCluster(pts), largest), "largest cluster case %d" % i


def test_golden_depth_and_signatures(ctx):
    for i, r, clen, seq4, seq_off in util.golden_cg_cases():
        regions = api.whole_contig_regions(clen)
        b = run_batch(ctx, r, regions)
        sums, nzs = b.depth_stats()
        sg = b.sigs()
        for tid in range(len(clen)):
            ans = util.golden_cg_answer(i, tid)
            assert np.array_equal(b.depth(tid), ans["depth"]), "depth case %d tid %d" % (i, tid)
            assert int(sums[tid]) == ans["sum"] and int(nzs[tid]) == ans["nonzero"]
            mean = float(sums[tid]) / float(nzs[tid]) if nzs[tid] else 0.0
            assert mean == ans["mean"]
            lo, hi = int(sg["region_off"][tid]), int(sg["region_off"][tid + 1])
            assert np.array_equal(sg["start"][lo:hi], ans["start"]) and np.array_equal(sg["end"][lo:hi], ans["end"]), "case %d tid %d" % (i, tid)
            assert np.array_equal(sg["kind"][lo:hi], ans["kind"])
            alts = [util.oracle_alt(seq4, seq_off, {k: sg[k][j] for k in ("start", "end", "kind", "read_idx", "query_pos")}) for j in range(lo, hi)]
            assert alts == ans["alt"], "ALT allele case %d tid %d" % (i, tid)
            if "l2" in ans:
                a, e, ss = ans["l2par"]
                su, cn = b.window_sums(tid, [a], [e], ss)
                lg = api.CNVCaller.log2_from_sums(su[0], cn[0], ans["mean"])
                # golden is in the reference's hash order keyed by window centre: compare as multisets per centre
                step = float(np.uint32(e - a + 1)) / float(ss)
                cent = np.array([(int(np.uint32(a + k * step)) + int(np.uint32(a + (k + 1) * step))) // 2 for k in range(ss)], np.uint64)
                o = np.argsort(ans["l2pos"], kind="stable"); oo = np.argsort(cent, kind="stable")
                assert np.array_equal(cent[oo], ans["l2pos"][o].astype(np.uint64))
                ref = ans["l2"][o]; got = lg[oo]
                assert np.all(np.abs(got - ref) <= 1e-6 * np.maximum(np.abs(ref), 1e-3)), "log2 ratio beyond 1e-6 relative"
        b.free()


def test_random_adversarial_cigars(ctx, oracle):
    rng = np.random.default_rng(11)
    for it in range(30):
        clen = [int(rng.choice([300, 2500, 12000, 70000])) for _ in range(int(rng.integers(1, 4)))]
        r = util.random_cigar_reads(rng, int(rng.integers(0, 400)), clen, n_tids=len(clen), weird=bool(it % 2),
                                    max_ops=int(rng.choice([3, 12, 40])))
        check_contigs(ctx, oracle, r, clen)


def test_long_reads_spanning_many_spans(ctx, oracle):
    """ONT-like records: thousands of ops per record, so records straddle many spans of the walk (1024 ops each)."""
    r = util.synth_reads([3_000_000], seed=21, profile=1, coverage=6.0, read_len_mean=50000, indel_rate=0.1, indel_len_max=4, n_sv=60)
    assert (np.diff(r["cig_off"]).max()) > 5000
    check_contigs(ctx, oracle, r, [3_000_000])


def test_full_size_whole_genome_properties(ctx, oracle):
    """BASELINE configs[1] at full size (6.2 M reads, 375 M CIGAR ops, 3.1 G depth positions, one batch): exact parity with
    the oracle on every one of the 24 contigs, and size-independent properties on top --
    sum(depth) == covered bases counted from the CIGARs, INS / DEL signature counts == qualifying ops counted from the CIGARs, every
    region's signature list sorted the way addSVCall leaves it, DBSCAN1D labels dense per group, and a second pass over the
    resident batch reproducing the first bit for bit."""
    from contextsv_b200 import shard
    clen = [l for _, l in shard.GRCH38]
    r = util.synth_reads(clen, seed=20261019, n_sv=25000)
    regions = api.whole_contig_regions(clen)
    b = run_batch(ctx, r, regions)
    sums, nzs = b.depth_stats()
    sg = b.sigs()
    lab = b.sigs_dbscan1d(100.0, 5)
    # ---- properties from the CIGARs alone (numpy, no oracle)
    cig = r["cigar"]; off = r["cig_off"].astype(np.int64); n = int(r["n_reads"])
    op = cig & 15; ln = (cig >> 4).astype(np.int64)
    rec = np.repeat(np.arange(n), np.diff(off))
    flag = r["flag"]; tid = r["tid"]
    depth_ok = (flag & (0x4 | 0x100 | 0x200 | 0x400)) == 0
    sig_ok = ((flag & (0x4 | 0x100 | 0x200 | 0x400 | 0x800)) == 0) & (r["mapq"] >= 20)
    ends = shard.ref_end(r)
    inside = ends <= np.asarray(clen, np.int64)[tid] + 1
    assert inside.all(), "the generator keeps reads inside their contig; the property below relies on it"
    covered = np.bincount(tid[rec], weights=(ln * np.isin(op, (0, 7, 8)) * depth_ok[rec]).astype(np.float64), minlength=len(clen))
    assert np.array_equal(sums.astype(np.int64), covered.astype(np.int64))
    # I and D of >= 50 bases from records that pass the filter (soft clips also depend on the position: sv_caller.cpp:602)
    for kind, opv in ((0, 1), (1, 2)):
        want = np.bincount(tid[rec], weights=((op == opv) & (ln >= 50) & sig_ok[rec]).astype(np.float64), minlength=len(clen)).astype(np.int64)
        seg = np.repeat(np.arange(len(clen)), np.diff(sg["region_off"].astype(np.int64)))
        got = np.bincount(seg[sg["kind"] == kind], minlength=len(clen))
        assert np.array_equal(got, want), kind
    clips = np.bincount(tid[rec], weights=((op == 4) & (ln >= 50) & sig_ok[rec]).astype(np.float64), minlength=len(clen)).astype(np.int64)
    seg = np.repeat(np.arange(len(clen)), np.diff(sg["region_off"].astype(np.int64)))
    assert np.all(np.bincount(seg[sg["kind"] == 2], minlength=len(clen)) <= clips)
    for t in range(len(clen)):
        lo, hi = int(sg["region_off"][t]), int(sg["region_off"][t + 1])
        key = sg["start"][lo:hi].astype(np.int64) * (1 << 32) + sg["end"][lo:hi]
        assert np.all(np.diff(key) >= 0)
        seq = sg["read_idx"][lo:hi].astype(np.int64) * (1 << 20) + sg["op_idx"][lo:hi]
        tie = np.diff(key) == 0
        assert np.all(np.diff(seq)[tie] < 0)                       # equal keys: reverse insertion order
        for is_del in (True, False):
            l = lab[lo:hi][(sg["kind"][lo:hi] == 1) == is_del]
            ids = np.unique(l[l >= 0])
            assert np.all((l >= 0) | (l == -2)) and np.array_equal(ids, np.arange(len(ids)))
    # ---- exact parity on ALL 24 contigs of the full-size run (chr1-sized ones included: 5x the records per region of
    # chr21, equal-start runs in the ranking kernel, tile offsets near 2^31 bytes): the oracle side -- per-base depth loop,
    # O(N log N) signature order, closed-form DBSCAN1D -- runs one contig per host thread beside the fetches
    import concurrent.futures as cf
    import os

    def oracle_side(t, got):
        d, s_, nz_ = oracle.depth(r, t, clen[t] + 1)
        assert np.array_equal(got, d), "depth of contig %d" % t
        assert int(sums[t]) == s_ and int(nzs[t]) == nz_, "stats of contig %d" % t
        del d, got
        o = oracle.cigar_scan(r, t, clen[t] + 1)
        lo, hi = int(sg["region_off"][t]), int(sg["region_off"][t + 1])
        assert hi - lo == len(o)
        for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
            assert np.array_equal(sg[f][lo:hi], o[f]), (t, f)
        for is_del in (True, False):
            m = (sg["kind"][lo:hi] == 1) == is_del
            assert np.array_equal(lab[lo:hi][m], oracle.dbscan1d(sg["start"][lo:hi][m].astype(np.int32), 100.0, 5, fast=True)), (t, is_del)
        return t

    with cf.ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1))) as ex:
        futs = [ex.submit(oracle_side, t, b.depth(t)) for t in sorted(range(len(clen)), key=lambda t: -clen[t])]
        assert sorted(f.result() for f in futs) == list(range(len(clen)))
    # ---- idempotence: the second pass over the resident batch gives the same bits
    d21 = b.depth(21).copy()
    b.scan(want_depth=True, want_sigs=True)
    sums2, nzs2 = b.depth_stats()
    sg2 = b.sigs()
    assert np.array_equal(sums, sums2) and np.array_equal(nzs, nzs2) and np.array_equal(b.depth(21), d21)
    for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
        assert np.array_equal(sg[f], sg2[f])
    # ---- ... and so does every one of 60 more, with the tile kernel under full HBM back-pressure (regression: a warp
    # running a whole tile ahead of a stalled one once overwrote its share of the carry-in: +-k over 1024 positions in
    # about one tile in 10^5, i.e. every third whole-genome pass)
    cks = b.depth_checksum()
    for _ in range(60):
        b.scan(want_depth=True, want_sigs=True)
        b.sigs_dbscan1d(100.0, 5, fetch=False)
        s3, z3 = b.depth_stats()
        assert np.array_equal(s3, sums) and np.array_equal(z3, nzs) and np.array_equal(b.depth_checksum(), cks)
    b.free()


def test_op_level_prepass_still_serves_records_without_gap_counts(ctx, oracle, monkeypatch):
    """csv_reads::n_gap is optional.  Every other test hands it over (api.Batch counts the D / N ops like a packer
    would) and takes the record-level pre-pass; here the field stays NULL and the op-level pre-pass (k_span_agg + span
    scan) has to give the same bits -- adversarial CIGARs, empty records, multi-contig HiFi, pile-ups."""
    monkeypatch.setattr(api, "COUNT_GAPS", False)
    rng = np.random.default_rng(12)
    for it in range(12):
        clen = [int(rng.choice([300, 2500, 12000, 70000])) for _ in range(int(rng.integers(1, 4)))]
        r = util.random_cigar_reads(rng, int(rng.integers(0, 400)), clen, n_tids=len(clen), weird=bool(it % 2), max_ops=int(rng.choice([3, 12, 40])))
        assert r.get("n_gap") is None
        check_contigs(ctx, oracle, r, clen)
    clen = [1_500_000, 700_000, 50_000]
    r = dict(util.synth_reads(clen, seed=5, n_sv=300, coverage=30.0, frac_len50=0.1))
    r.pop("n_gap")
    check_contigs(ctx, oracle, r, clen)


def test_record_level_prepass(ctx, oracle):
    """The record-level pre-pass (csv_reads::n_gap given): span starts inside, at the head of and right behind records,
    records longer than several spans among short ones, empty CIGARs between real ones, records of gaps only, a batch of
    exactly k * 1024 ops (the sentinel span), and a wrong count, which must be refused and never trusted."""
    from oracle.oracle_py import make_reads
    from contextsv_b200._capi import CsvError, count_gaps
    rng = np.random.default_rng(5)
    # (a) total ops an exact multiple of the span, records ending exactly at span starts
    for n_ops_total, sizes in ((2048, (1024, 512, 512)), (3072, (1000, 24, 1024, 1, 1023)), (1024, (1024,)), (4096, (3000, 1096))):
        pos, cig = [], []
        p = 5
        for sz in sizes:
            ops = []
            for j in range(sz):
                ops.append((int(rng.integers(1, 4)), [M, D, I, N, EQ][int(rng.integers(0, 5))]))
            pos.append(p); cig.append(ops); p += int(rng.integers(0, 50))
        r = make_reads(np.array(pos, np.int32), cig)
        assert int(r["n_ops"]) == n_ops_total
        check_contigs(ctx, oracle, r, [20000])
    # (b) long records among short ones, empty records, gap-only records
    pos, cig = [], []
    p = 0
    for i in range(600):
        kind = i % 7
        if kind == 0:
            ops = [(int(rng.integers(1, 30)), [M, D, I, S, N, X][int(rng.integers(0, 6))]) for _ in range(int(rng.integers(1500, 4000)))]
        elif kind == 1:
            ops = []
        elif kind == 2:
            ops = [(int(rng.integers(1, 9)), D), (3, N)]
        else:
            ops = [(int(rng.integers(1, 70)), [M, D, I, S][int(rng.integers(0, 4))]) for _ in range(int(rng.integers(1, 90)))]
        pos.append(p); cig.append(ops); p += int(rng.integers(0, 400))
    r = make_reads(np.array(pos, np.int32), cig, flag=rng.choice([0, 16, 0x800, 0x100], 600).astype(np.uint16), mapq=rng.choice([0, 60], 600).astype(np.uint8))
    assert int(r["n_ops"]) <= 1024 * 600
    check_contigs(ctx, oracle, r, [p + 50000])
    # (c) a wrong count is reported, whichever way it is wrong
    r = dict(util.synth_reads([400_000], seed=9, n_sv=40, coverage=20.0))
    good = count_gaps(r)
    assert np.array_equal(good, r["n_gap"]) and good.sum() > 0
    for delta in (1, -1, 1000):
        bad = dict(r); g = good.astype(np.int64).copy()
        nzi = np.nonzero(g > 0)[0]
        victim = int(nzi[len(nzi) // 3])
        g[victim] = max(0, g[victim] + delta); bad["n_gap"] = g.astype(np.uint32)
        b = run_batch(ctx, bad, api.whole_contig_regions([400_000]))
        with pytest.raises(CsvError, match="n_gap"):
            b.depth_stats()
        b.free()
    # (d) ... and so is a wrong reference length (csv_reads::ref_len: the tile ranges are computed from it BEFORE the walk,
    # the walk compares it with what the CIGAR says)
    from contextsv_b200._capi import record_stats
    g2, l2 = record_stats(r)
    assert np.array_equal(g2, good) and np.array_equal(l2, r["ref_len"]) and np.array_equal(l2.astype(np.int64), shard.ref_end(r) - r["pos0"].astype(np.int64) - 1)
    plain = np.nonzero((r["flag"] == 0) & (r["pos0"] > 1000))[0]
    claim_on = os.environ.get("CSV_CLAIM_REFLEN", "0") not in ("", "0")     # opt-in (capi.cu: measured slower on B200); ignored otherwise
    for delta in (1, -1, 5000):
        bad = dict(r); l = l2.astype(np.int64).copy()
        victim = int(plain[len(plain) // 3])
        l[victim] = max(0, l[victim] + delta); bad["ref_len"] = l.astype(np.uint32)
        b = run_batch(ctx, bad, api.whole_contig_regions([400_000]))
        if claim_on:
            with pytest.raises(CsvError, match="ref_len"):
                b.depth_stats()
        else:
            b.depth_stats()                                                  # the claim is not used: a wrong one cannot hurt
        b.free()
    # without the lengths (n_gap only) the ranges are computed after the walk, as before
    only_gaps = dict(r); only_gaps.pop("ref_len")
    check_contigs(ctx, oracle, only_gaps, [400_000])
    check_contigs(ctx, oracle, r, [400_000])


def test_inputs_the_path_refuses(ctx):
    """Not silently wrong: unsorted records (the reference needs an indexed = sorted BAM) and a record that consumes
    2^31 reference bases (BAM positions are int32) are reported when results are fetched."""
    from oracle.oracle_py import make_reads
    from contextsv_b200._capi import CsvError
    r = make_reads([500, 100, 900], [[(50, M)], [(60, M)], [(10, M)]])
    b = run_batch(ctx, r, api.whole_contig_regions([2000]))
    with pytest.raises(CsvError, match="sorted"):
        b.depth_stats()
    b.free()
    big = (1 << 28) - 1
    r = make_reads([10, 20], [[(big, M)] * 9, [(30, M)]])
    b = run_batch(ctx, r, api.whole_contig_regions([5000]))
    with pytest.raises(CsvError, match="2\\^31"):
        b.depth_stats()
    b.free()
    # the same record one op shorter is fine (and clipped to the map like any read running off the contig end)
    r = make_reads([10, 20], [[(big, M)] * 7, [(30, M)]])
    b = run_batch(ctx, r, api.whole_contig_regions([5000]))
    d = b.depth(0)
    assert d[10] == 0 and d[11] == 1 and d[21] == 2 and d[50] == 2 and d[51] == 1 and d[5000] == 1
    b.free()


def test_pileups_deep_coverage(ctx, oracle):
    """Thousands of records starting / ending / deleting at the same base: tiles whose slice holds more than
    32767 pairs take the 32-bit-counter kernel, the others run the 16-bit one close to its bound."""
    from oracle.oracle_py import make_reads
    rng = np.random.default_rng(77)
    for n, L in ((70_000, 30_000), (20_000, 30_000), (33_000, 9_000)):
        pos0 = np.sort(np.concatenate([np.full(n // 2, 4000), rng.integers(0, L - 3000, n - n // 2)])).astype(np.int32)
        cig = []
        for p in pos0:
            if p == 4000:
                cig.append([(100, M), (7, D), (300, M)])                      # same start, same gap, same end
            else:
                a = int(rng.integers(1, 800))
                cig.append([(a, M), (int(rng.integers(1, 30)), D), (int(rng.integers(1, 900)), EQ)] if rng.random() < 0.5 else [(a, M)])
        r = make_reads(pos0, cig)
        check_contigs(ctx, oracle, r, [L])


def test_narrow_depth_fetch(oracle):
    """csv_depth_fetch / csv_depth_fetch_all bring the map back as bytes + a list of the values >= 255 and widen it on
    host threads.  Must equal the plain 32-bit DMA and the oracle for: many chunks through a short ring, values that
    need the list (a pile-up in the thousands), a list that overflows (chunk re-fetched plain), odd lengths,
    unaligned and pageable destinations."""
    from oracle.oracle_py import make_reads
    c = api.Context(0)
    rng = np.random.default_rng(99)
    clen = [40_001, 9_999, 513]
    pos0, tid, cig = [], [], []
    for t, L in enumerate(clen):
        n = {0: 3000, 1: 600, 2: 400}[t]
        p = np.sort(rng.integers(0, max(L - 300, 1), n))
        if t == 0:
            p = np.sort(np.concatenate([p, np.full(2500, 20_000), np.full(300, 33_000)]))    # depth > 2500 around 20 000, ~300 at 33 000
        for x in p:
            pos0.append(int(x)); tid.append(t)
            cig.append([(int(rng.integers(20, 200)), M), (int(rng.integers(1, 9)), D), (int(rng.integers(1, 60)), EQ)])
    r = make_reads(np.array(pos0, np.int32), cig, tid=np.array(tid, np.int32))
    regions = api.whole_contig_regions(clen)
    b = run_batch(c, r, regions, want_sigs=False)
    want = [oracle.depth(r, t, clen[t] + 1)[0] for t in range(len(clen))]
    assert want[0].max() > 2500
    c.set_fetch(threads=0)
    for t in range(len(clen)):
        assert np.array_equal(b.depth(t), want[t])
    n0 = c.fetch_stats()
    assert n0 == (0, 0)
    for threads, chunk, slots in ((3, 512, 64), (16, 1024, 2048), (2, 512, 4), (1, 4096, 0), (5, 1 << 20, 2048)):
        c.set_fetch(threads=threads, chunk_positions=chunk, exception_slots=slots, min_positions=0)
        before = c.fetch_stats()
        for t in range(len(clen)):
            buf = np.full(clen[t] + 1 + 3, 0xabababab, np.uint32)           # unaligned view into a pageable array
            got = b.depth(t, out=buf[3:])
            assert np.array_equal(got, want[t]), "narrow fetch tid %d (threads %d chunk %d slots %d): first diff at %s" % (
                t, threads, chunk, slots, np.nonzero(got != want[t])[0][:5])
            assert (buf[:3] == 0xabababab).all()
        outs = b.depth_all()
        for t in range(len(clen)):
            assert np.array_equal(outs[t], want[t])
        after = c.fetch_stats()
        assert after[0] > before[0]
        if slots <= 4:
            assert after[1] > before[1], "the pile-up chunks must have overflowed the exception list"
    for bad in (dict(chunk_positions=100), dict(chunk_positions=513), dict(chunk_positions=1 << 29), dict(exception_slots=1 << 25)):
        with pytest.raises(api.CsvError):
            c.set_fetch(**bad)
    # the one-shot entry point takes the same path
    c.set_fetch(threads=4, chunk_positions=2048, exception_slots=16, min_positions=0)
    rs, keep = reads_struct(r)
    out = np.zeros(clen[0] + 1, np.uint32); s = C.c_uint64(0); nz = C.c_uint32(0)
    reg = CsvRegion(0, 0, clen[0] + 1, clen[0] + 1)
    check(lib().csv_depth(c.h, C.byref(rs), C.byref(reg), ptr(out), C.byref(s), C.byref(nz)))
    assert np.array_equal(out, want[0]) and int(s.value) == int(want[0].astype(np.uint64).sum())
    b.free()
    c.close()


def test_synthetic_hifi_multi_contig(ctx, oracle):
    clen = [1_500_000, 700_000, 50_000]
    r = util.synth_reads(clen, seed=5, n_sv=300, coverage=30.0, frac_len50=0.1)
    check_contigs(ctx, oracle, r, clen)


def test_pipelined_chunks_of_contigs(ctx, oracle):
    """Three contigs big enough for three pipeline chunks (cuts between contigs, walk ranges aligned to 2048 spans):
    the chunked two-stream pass must give what the serial pass gives."""
    clen = [45_000_000, 40_000_000, 42_000_000]
    r = util.synth_reads(clen, seed=31, n_sv=200, coverage=30.0)
    # cuts fall on multiples of 2048 spans of 1024 ops: each contig must start in a later 2 M-op block than the one before
    block = [int(r["cig_off"][int(np.searchsorted(r["tid"], t))]) // (1024 * 2048) for t in (1, 2)]
    assert 0 < block[0] < block[1] and int(r["n_ops"]) // (1024 * 2048) >= block[1] + 1
    ctx.set_pipeline_chunks(4)
    try:
        check_contigs(ctx, oracle, r, clen)
    finally:
        ctx.set_pipeline_chunks(1)


def test_empty_and_degenerate_inputs(ctx, oracle):
    from oracle.oracle_py import make_reads
    # no reads at all
    r = make_reads([], [])
    check_contigs(ctx, oracle, r, [1000])
    # only empty CIGARs, and empty ones between real ones
    r = make_reads([5, 7, 9, 9, 30], [[], [(10, 0)], [], [], [(60, 1), (5, 0)]])
    check_contigs(ctx, oracle, r, [100])
    # one op exactly at the end of the contig; soft clip beyond the end (sv_caller.cpp:602-604)
    r = make_reads([90, 95], [[(10, 0), (60, 4), (50, 1)], [(5, 0), (50, 4), (50, 1), (50, 2)]])
    check_contigs(ctx, oracle, r, [100])
    # contig exactly one tile and one tile + 1
    for L in (8191, 8192, 8193):
        r = util.synth_reads([L], seed=L, coverage=40.0, read_len_mean=900, read_len_sd=100, n_sv=3)
        check_contigs(ctx, oracle, r, [L])


def test_one_shot_entry_points(ctx, oracle):
    clen = [60_000]
    r = util.synth_reads(clen, seed=9, n_sv=30, coverage=20.0)
    rs, keep = reads_struct(r)
    reg = CsvRegion(0, 0, clen[0] + 1, clen[0] + 1)
    depth = np.zeros(clen[0] + 1, np.uint32); s = C.c_uint64(0); nz = C.c_uint32(0)
    check(lib().csv_depth(ctx.h, C.byref(rs), C.byref(reg), ptr(depth), C.byref(s), C.byref(nz)))
    d, s0, nz0 = oracle.depth(r, 0, clen[0] + 1)
    assert np.array_equal(depth, d) and s.value == s0 and nz.value == nz0
    o = oracle.cigar_scan(r, 0, clen[0] + 1)
    cap = len(o) + 8
    arr = {k: np.zeros(cap, np.uint8 if k == "kind" else np.uint32) for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")}
    st = CsvSigs(*[ptr(arr[k]) for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")])
    n = C.c_uint64(0)
    check(lib().csv_cigar_scan(ctx.h, C.byref(rs), C.byref(reg), 50, 20, C.byref(st), cap, C.byref(n)))
    assert n.value == len(o)
    for k in arr:
        assert np.array_equal(arr[k][: len(o)], o[k])
    # too small a caller buffer is reported, not overrun
    rc = lib().csv_cigar_scan(ctx.h, C.byref(rs), C.byref(reg), 50, 20, C.byref(st), 1, C.byref(n))
    assert rc == 3 and n.value == len(o)


def test_region_sharding_with_halo_reads(ctx, oracle):
    """shards of one contig give disjoint depth slices whose concatenation is the whole map,
    stats add up, and every signature is emitted by exactly one shard."""
    L = 400_000
    r = util.synth_reads([L], seed=13, n_sv=120, coverage=25.0)
    d, s, nz = oracle.depth(r, 0, L + 1)
    o = oracle.cigar_scan(r, 0, L + 1)
    cuts = [0, 100_003, 100_004, 250_000, L + 1]
    regions = [(0, cuts[i], cuts[i + 1], L + 1) for i in range(len(cuts) - 1)]
    b = run_batch(ctx, r, regions)
    got = np.concatenate([b.depth(i) for i in range(len(regions))])
    assert np.array_equal(got, d)
    sums, nzs = b.depth_stats()
    assert int(sums.sum()) == s and int(nzs.sum()) == nz
    sg = b.sigs()
    assert len(sg["start"]) == len(o)
    # host merge with the addSVCall comparator: (start, end, reverse insertion order)
    order = np.lexsort((-(sg["read_idx"].astype(np.int64) * (1 << 20) + sg["op_idx"]), sg["end"], sg["start"]))
    for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
        assert np.array_equal(sg[f][order], o[f])
    # ownership: region of a signature == region containing pos0 + 1 of its read
    for i in range(len(regions)):
        lo, hi = int(sg["region_off"][i]), int(sg["region_off"][i + 1])
        idx = r["pos0"][sg["read_idx"][lo:hi]].astype(np.int64) + 1
        assert np.all((idx >= cuts[i]) & (idx < cuts[i + 1]))
    b.free()
    # regions in scrambled caller order
    perm = [2, 0, 3, 1]
    b = run_batch(ctx, r, [regions[p] for p in perm])
    for j, p in enumerate(perm):
        assert np.array_equal(b.depth(j), d[cuts[p]:cuts[p + 1]])
    b.free()


def test_streamed_shards_by_op_budget(ctx, oracle):
    """BASELINE config 3 in small: ONT-like dense CIGARs scanned as a sequence of shards whose records hold at
    most max_ops ops (a batch takes < 2^31), merged on the host -- identical to one whole-contig scan."""
    clen = [1_200_000, 500_000]
    r = util.synth_reads(clen, seed=41, profile=1, coverage=12.0, read_len_mean=30000, indel_rate=0.08, indel_len_max=4, n_sv=80, sv_jitter_sd=5.0)
    from contextsv_b200 import shard
    plans = shard.plan_by_ops(r, clen, 600_000)
    assert len(plans) >= 4
    depth, (sums, nzs), sigs, labels = api.scan_streamed(ctx, r, clen, max_ops=600_000, eps=100.0, min_pts=5)
    for tid in range(len(clen)):
        d, s, nz = oracle.depth(r, tid, clen[tid] + 1)
        assert np.array_equal(depth[tid], d) and int(sums[tid]) == s and int(nzs[tid]) == nz
        o = oracle.cigar_scan(r, tid, clen[tid] + 1)
        sg = sigs.get(tid, {k: np.zeros(0) for k in ("start",)})
        assert len(sg["start"]) == len(o)
        for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
            assert np.array_equal(sg[f].astype(np.int64), o[f].astype(np.int64)), f
        for is_del in (True, False):
            m = (sg["kind"] == 1) if is_del else (sg["kind"] != 1)
            assert np.array_equal(labels[tid][m], oracle.dbscan1d(sg["start"][m].astype(np.int32), 100.0, 5, fast=True))
    with pytest.raises(ValueError):
        shard.plan_by_ops(r, clen, 2000)      # fewer ops than the records over one position hold


def test_ont60x_full_contigs_streamed(ctx, oracle):
    """BASELINE configs[2] at the contig sizes it names: 60x ONT ultra-long reads (N50 50 kb, ~10 % indel rate, ~0.2 CIGAR
    ops per aligned base) over two whole GRCh38-sized contigs (chr21 + chr22: 0.96e9 CIGAR ops), scanned as a stream of
    shards of at most 2^28 ops + records -- what the whole genome at this depth (~3.7e10 ops) needs on one device -- and
    merged on the host.  Exact parity with the oracle: depth of every base, signatures in addSVCall order, DBSCAN1D labels."""
    from contextsv_b200 import shard
    import bench
    clen = [shard.GRCH38[20][1], shard.GRCH38[21][1]]
    r = util.synth_reads(clen, seed=20261018 + 3, n_sv=400, **bench.SYNTH_KW["ont60x_chr20"])
    assert int(r["n_ops"]) > 900_000_000
    max_ops = 1 << 28
    assert len(shard.plan_by_ops(r, clen, max_ops)) >= 4
    depth, (sums, nzs), sigs, labels = api.scan_streamed(ctx, r, clen, max_ops=max_ops, eps=100.0, min_pts=5)
    import concurrent.futures as cf

    def oracle_side(tid):
        d, s, nz = oracle.depth(r, tid, clen[tid] + 1)
        assert np.array_equal(depth[tid], d) and int(sums[tid]) == s and int(nzs[tid]) == nz
        o = oracle.cigar_scan(r, tid, clen[tid] + 1)
        sg = sigs[tid]
        assert len(sg["start"]) == len(o) > 1000
        for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
            assert np.array_equal(sg[f].astype(np.int64), o[f].astype(np.int64)), f
        for is_del in (True, False):
            m = (sg["kind"] == 1) if is_del else (sg["kind"] != 1)
            assert np.array_equal(labels[tid][m], oracle.dbscan1d(sg["start"][m].astype(np.int32), 100.0, 5, fast=True))
        return True
    with cf.ThreadPoolExecutor(max_workers=2) as ex:
        assert all(ex.map(oracle_side, range(len(clen))))


def test_long_cigar_cg_tag_through_the_packer(ctx, oracle, tmp_path):
    """A record with more than 65 535 CIGAR ops (reachable at config 3's read lengths): in the BAM its CIGAR lives in a CG:B,I tag
    behind the placeholder <l_seq>S<ref_len>N (SAM spec 4.2.2).  BAM -> htslib iterator (the shim promotes the tag like
    bam_read1 does) -> the C++ host packer -> GPU must give the depth and signatures of the real CIGAR."""
    import ctypes
    import subprocess
    from contextsv_b200 import bamio, build
    from oracle.oracle_py import make_reads
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host, shim = os.path.join(build.CSRC, "host"), os.path.join(root, "oracle", "htslib_shim")
    so = str(tmp_path / "libhp.so")
    p = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + shim, "-I" + os.path.join(root, "oracle"), "-I" + build.INC, "-I" + host,
                        os.path.join(root, "tests", "native", "host_packer_harness.cpp"), os.path.join(host, "packed_reads.cpp"), os.path.join(shim, "shim.cpp"),
                        "-o", so, "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    L = ctypes.CDLL(so)
    L.hp_pack.restype = ctypes.c_void_p; L.hp_pack.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    L.hp_size.restype = ctypes.c_uint64; L.hp_size.argtypes = [ctypes.c_void_p]
    L.hp_ops.restype = ctypes.c_uint64; L.hp_ops.argtypes = [ctypes.c_void_p]
    L.hp_copy.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 7
    L.hp_free.argtypes = [ctypes.c_void_p]
    rng = np.random.default_rng(77)
    clen = [900_000]
    pos, cig = [], []
    for i in range(40):
        n_ops = 70_000 if i in (7, 23) else int(rng.integers(20, 3000))          # two records beyond the 16-bit op count
        ops = []
        for j in range(n_ops):
            if j % 2 == 0:
                ops.append((int(rng.integers(1, 9)), M))
            else:
                ops.append((60 if rng.random() < 0.002 else int(rng.integers(1, 3)), [I, D][int(rng.integers(0, 2))]))
        pos.append(int(rng.integers(0, 200_000))); cig.append(ops)
    order = np.argsort(pos, kind="stable")
    r = make_reads(np.array([pos[i] for i in order], np.int32), [cig[i] for i in order])
    assert int(np.diff(r["cig_off"].astype(np.int64)).max()) > 65_535
    path = str(tmp_path / "long.bam")
    bamio.write_bam(path, r, ["chrL"], clen, seed=5)
    h = L.hp_pack(path.encode(), b"chrL", 0)
    n, no = int(L.hp_size(h)), int(L.hp_ops(h))
    assert n == int(r["n_reads"]) and no == int(r["n_ops"])                     # the placeholder CIGARs were replaced by the tags' contents
    a = {"tid": np.zeros(n, np.int32), "pos0": np.zeros(n, np.int32), "flag": np.zeros(n, np.uint16), "mapq": np.zeros(n, np.uint8),
         "cig_off": np.zeros(n + 1, np.uint64), "cigar": np.zeros(no, np.uint32), "ref_end": np.zeros(n, np.uint32)}
    L.hp_copy(h, *[a[k].ctypes.data_as(ctypes.c_void_p) for k in ("tid", "pos0", "flag", "mapq", "cig_off", "cigar", "ref_end")])
    L.hp_free(h)
    assert np.array_equal(a["cigar"], r["cigar"]) and np.array_equal(a["cig_off"], r["cig_off"])
    packed = {"n_reads": n, "n_ops": no, "tid": a["tid"], "pos0": a["pos0"], "flag": a["flag"], "mapq": a["mapq"], "cig_off": a["cig_off"], "cigar": a["cigar"]}
    check_contigs(ctx, oracle, packed, clen)


def test_against_the_compiled_reference_directly(ctx, reference):
    """Closes the loop on the GPU box: the CUDA path against oracle/_ref -- the reference's own sources compiled unmodified
    -- without the C restatement in between: depth, mean coverage, SVCall vectors (start, end, type, evidence, ALT) and
    DBSCAN1D labels on synthetic HiFi records and on adversarial CIGARs."""
    rng = np.random.default_rng(31)
    cases = [(util.synth_reads([260_000, 90_000], seed=17, n_sv=80, coverage=22.0, frac_len50=0.3), [260_000, 90_000])]
    for it in range(4):
        clen = [int(rng.choice([2500, 12000, 70000])) for _ in range(int(rng.integers(1, 3)))]
        cases.append((util.random_cigar_reads(rng, int(rng.integers(50, 400)), clen, n_tids=len(clen), weird=bool(it % 2)), clen))
    for r, clen in cases:
        seq4, seq_off = util.random_seq4(rng, r)
        b = run_batch(ctx, r, api.whole_contig_regions(clen))
        sums, nzs = b.depth_stats()
        sg = b.sigs()
        for tid in range(len(clen)):
            d, s, nz, mean = reference.depth(r, tid, clen)
            assert np.array_equal(b.depth(tid), d) and int(sums[tid]) == s and int(nzs[tid]) == nz
            assert (float(s) / float(nz) if nz else 0.0) == mean
            st, en, ty, ev, alts = reference.cigar_scan(r, tid, clen, seq4=seq4, seq_off=seq_off)
            lo, hi = int(sg["region_off"][tid]), int(sg["region_off"][tid + 1])
            assert hi - lo == len(st)
            assert np.array_equal(sg["start"][lo:hi], st) and np.array_equal(sg["end"][lo:hi], en)
            kind = sg["kind"][lo:hi]
            assert np.array_equal(np.where(kind == 1, 0, 3), ty)                                 # SVType: DEL 0, INS 3
            assert np.array_equal(1 << np.where(kind == 0, 0, np.where(kind == 1, 1, 2)), ev)     # CIGARINS / CIGARDEL / CIGARCLIP bit
            for i in range(lo, hi):
                one = {k: sg[k][i] for k in ("start", "end", "kind", "read_idx", "query_pos")}
                assert util.oracle_alt(seq4, seq_off, one) == alts[i - lo]
            for t_sv in (0, 3):
                pts = st[ty == t_sv].astype(np.int32)
                if len(pts) and len(pts) < 4000:
                    lab = np.zeros(len(pts), np.int32)
                    check(lib().csv_dbscan1d(ctx.h, ptr(pts), len(pts), 100.0, 5, ptr(lab), None))
                    assert np.array_equal(lab, reference.dbscan1d(pts, 100.0, 5))
        b.free()


def test_signature_pileup_at_one_start(ctx, oracle):
    """2e5 signatures that share one (region, start): amplicon-like data, every read carrying an insertion at the same
    coordinate.  The radix sort only orders (region, start); the order inside the run -- (end, reverse insertion order) --
    comes from the ranking kernel, which must stay exact and affordable for a run of this length."""
    import time
    from oracle.oracle_py import make_reads
    rng = np.random.default_rng(3)
    n = 200_000
    lens = rng.integers(50, 90, n)
    cig = [[(10, M), (int(l), I), (20, M)] for l in lens]
    r = make_reads(np.full(n, 1000, np.int32), cig)
    clen = [5000]
    b = run_batch(ctx, r, api.whole_contig_regions(clen))
    t0 = time.perf_counter()
    sg = b.sigs()
    dt = time.perf_counter() - t0
    o = oracle.cigar_scan(r, 0, clen[0] + 1)
    assert len(o) == n == len(sg["start"]) and np.all(sg["start"] == 1011)
    for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
        assert np.array_equal(sg[f], o[f]), f
    assert dt < 5.0, "ranking a run of %d took %.1f s" % (n, dt)
    d, s_, nz_ = oracle.depth(r, 0, clen[0] + 1)
    sums, nzs = b.depth_stats()
    assert np.array_equal(b.depth(0), d) and int(sums[0]) == s_ and int(nzs[0]) == nz_       # 2e5-deep pile: the 32-bit tile kernel
    b.free()


def test_more_signatures_than_the_batch_reserved(ctx, oracle):
    """A batch reserves room for max(2^20, ops / 16) signatures; clip- and insertion-heavy records can emit more (the
    reference's vector has no limit).  4.8 M ops, every second one a 60-base insertion: 2.4 M signatures against room for
    1.05 M, in a batch large enough for the 32 slot counters.  The pass must report the overflow with the size to reserve
    (csv_sigs_fetch -> CSV_ERR_CAPACITY, csv_batch_reserve_sigs), the rescan (dense slots) must give the reference's vector
    and the depth must be untouched by any of it."""
    n, per = 300_000, 16
    rng = np.random.default_rng(11)
    pos0 = np.sort(rng.integers(0, 2_000_000, n)).astype(np.int32)
    one = np.array([(60 << 4) | I, (10 << 4) | M] * (per // 2), np.uint32)
    r = {"n_reads": n, "n_ops": n * per, "tid": np.zeros(n, np.int32), "pos0": pos0, "flag": np.zeros(n, np.uint16),
         "mapq": np.full(n, 60, np.uint8), "cig_off": (np.arange(n + 1, dtype=np.uint64) * per), "cigar": np.tile(one, n)}
    clen = [2_100_000]
    assert r["n_ops"] >= (1 << 22)
    b = run_batch(ctx, r, api.whole_contig_regions(clen))
    out = {k: np.zeros(8, dt) for k, dt in (("start", np.uint32), ("end", np.uint32), ("kind", np.uint8), ("read_idx", np.uint32), ("op_idx", np.uint32), ("query_pos", np.uint32))}
    sig = CsvSigs(*[ptr(out[k]) for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")])
    n_out = C.c_uint64(0)
    rc = lib().csv_sigs_fetch(ctx.h, b.h, C.byref(sig), 8, C.byref(n_out), None)
    assert rc == 3 and n_out.value > n * per // 2, "the overflow must be reported with what to reserve"      # CSV_ERR_CAPACITY
    sg = b.sigs()                                            # reserves what was reported and scans again
    assert len(sg["start"]) == n * per // 2
    o = oracle.cigar_scan(r, 0, clen[0] + 1)
    for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
        assert np.array_equal(sg[f], o[f]), f
    d, s_, nz_ = oracle.depth(r, 0, clen[0] + 1)
    sums, nzs = b.depth_stats()
    assert np.array_equal(b.depth(0), d) and int(sums[0]) == s_ and int(nzs[0]) == nz_
    lab = b.sigs_dbscan1d(100.0, 5)
    assert np.array_equal(lab, oracle.dbscan1d(sg["start"].astype(np.int32), 100.0, 5, fast=True))       # one group: every signature is an insertion
    b.free()


def test_sv_rich_sweep_eps_minpts(ctx, oracle):
    """BASELINE config 5 in small: SV-rich reads with per-read breakpoint jitter, DBSCAN1D over the signature starts
    of every (contig, SVType) group for the whole eps x minPts grid."""
    clen = [2_000_000]
    r = util.synth_reads(clen, seed=55, n_sv=4000, coverage=30.0, sv_jitter_sd=10.0)
    b = run_batch(ctx, r, api.whole_contig_regions(clen))
    sg = b.sigs()
    assert len(sg["start"]) > 20_000
    for eps in (0, 1, 10, 50, 100, 500, 1000):
        for mp in (1, 2, 3, 5, 10, 20):
            lab = b.sigs_dbscan1d(float(eps), mp)
            for is_del in (True, False):
                m = (sg["kind"] == 1) if is_del else (sg["kind"] != 1)
                want = oracle.dbscan1d(sg["start"][m].astype(np.int32), float(eps), mp, fast=True)
                assert np.array_equal(lab[m], want), (eps, mp, is_del)
    b.free()


def test_dbscan1d_fuzz_and_large(ctx, oracle):
    rng = np.random.default_rng(3)
    for it in range(200):
        n = int(rng.integers(0, 300))
        span = int(rng.choice([5, 30, 200, 5000, 2_000_000_000]))
        pts = rng.integers(-span, span, n).astype(np.int32)
        eps = float(rng.choice([-1, 0, 0.5, 1, 2, 3.7, 10, 50, 100, 1e12]))
        mp = int(rng.choice([-1, 0, 1, 2, 3, 5, 8]))
        db = api.DBSCAN1D(eps, mp, ctx); db.fit(pts)
        assert np.array_equal(db.getClusters(), oracle.dbscan1d(pts, eps, mp)), (pts.tolist(), eps, mp)
    # sizes the O(N^2) reference cannot reach: checked against the closed-form oracle
    for n, eps, mp in ((20_000, 100, 5), (300_000, 10, 3), (1_000_000, 50, 5)):
        centers = rng.integers(0, 46_000_000, n // 25)
        pts = (rng.choice(centers, n) + np.rint(rng.normal(0, 10, n)).astype(np.int64)).astype(np.int32)
        db = api.DBSCAN1D(eps, mp, ctx); db.fit(pts)
        want = oracle.dbscan1d(pts, eps, mp, fast=True) if n > 20_000 else oracle.dbscan1d(pts, eps, mp)
        assert np.array_equal(db.getClusters(), want)


def test_dbscan2d_golden_fuzz_and_large(ctx, oracle):
    """DBSCAN::fit over intervals: golden vectors of the compiled reference (closed-form AND literal
    path: eps >= 1, eps < 0, zero lengths), a fuzz against the oracle, and signature-shaped sets the closed form
    handles at sizes where the O(N^2) oracle still finishes."""
    for i, st, en, eps, mp, want in util.golden_db2_cases():
        db = api.DBSCAN(eps, mp, ctx); db.fit(st, en)
        assert np.array_equal(db.getClusters(), want), (i, eps, mp)
    rng = np.random.default_rng(12)
    for it in range(150):
        n = int(rng.integers(0, 400))
        centers = rng.integers(0, 20000, max(1, n // 8 + 1))
        st = (rng.choice(centers, n) + rng.integers(0, 30, n)).astype(np.uint32)
        en = (st + rng.choice([1, 49, 50, 80, 300, 5000], n) + rng.integers(0, 5, n)).astype(np.uint32)
        eps = float(rng.choice([0, 0.02, 0.1, 0.25, 0.5, 0.8, 0.999])); mp = int(rng.choice([0, 1, 2, 3, 5, 10]))
        db = api.DBSCAN(eps, mp, ctx); db.fit(st, en)
        assert np.array_equal(db.getClusters(), oracle.dbscan2d(st, en, eps, mp)), (it, n, eps, mp)
    for n, eps, mp in ((6000, 0.1, 2), (20000, 0.1, 3), (20000, 0.3, 5)):
        centers = rng.integers(0, 40_000_000, n // 20)
        lens = rng.integers(50, 5000, n // 20)
        pick = rng.integers(0, n // 20, n)
        st = (centers[pick] + np.rint(rng.normal(0, 8, n)).astype(np.int64)).astype(np.uint32)
        en = (st + lens[pick] + np.rint(rng.normal(0, 6, n)).astype(np.int64) + 1).astype(np.uint32)
        order = np.lexsort((en, st))                       # mergeSVs sees the calls in vector order
        st, en = st[order], en[order]
        db = api.DBSCAN(eps, mp, ctx); db.fit(st, en)
        assert np.array_equal(db.getClusters(), oracle.dbscan2d(st, en, eps, mp)), (n, eps, mp)


def test_dbscan1d_one_launch_small_path(oracle):
    """csv_dbscan1d (default; CSV_DB_SMALL=0 turns it off): fits of at most 1024 points run as ONE launch (k_db_small, points and labels
    through mapped pinned memory) instead of the general pipeline's ~20 -- what the split-read pass's per-cluster fits
    (sv_caller.cpp:270) need.  Same labels as the golden vectors and the oracle; larger inputs and eps < 0 still take
    the general path."""
    import os
    import time
    old = os.environ.get("CSV_DB_SMALL")
    os.environ["CSV_DB_SMALL"] = "1"
    try:
        c = api.Context(0)
    finally:
        if old is None:
            del os.environ["CSV_DB_SMALL"]
        else:
            os.environ["CSV_DB_SMALL"] = old
    n_small = 0
    for i, pts, eps, mp, labels, largest in util.golden_db_cases():
        l0 = c.launches
        db = api.DBSCAN1D(eps, mp, c)
        db.fit(pts)
        assert np.array_equal(db.getClusters(), labels), "case %d eps=%g minPts=%d" % (i, eps, mp)
        if 0 < len(pts) <= 1024 and eps >= 0:
            assert c.launches - l0 == 1
            n_small += 1
    assert n_small >= 80
    rng = np.random.default_rng(31)
    for trial in range(300):
        n = int(rng.choice([1, 2, 3, 5, 17, 33, 64, 200, 1000, 1024, 1025, 3000]))
        span = int(rng.choice([5, 50, 1000, 100000]))
        pts = rng.integers(-span, span, n).astype(np.int32)
        if trial % 7 == 0:
            pts[: n // 2] = pts[0]
        if trial % 11 == 0:
            pts[0] = np.iinfo(np.int32).max; pts[-1] = np.iinfo(np.int32).min
        eps = float(rng.choice([0, 0.5, 1, 2.9, 10, 100, 1e5, 5e9, -1.0]))
        mp = int(rng.choice([-1, 0, 1, 2, 3, 5, 10, 2000]))
        lab = np.zeros(n, np.int32); nc = np.zeros(1, np.int32)
        l0 = c.launches
        check(lib().csv_dbscan1d(c.h, ptr(pts), n, eps, mp, ptr(lab), ptr(nc)))
        want = oracle.dbscan1d(pts, eps, mp, fast=n > 1024)
        assert np.array_equal(lab, want), "trial %d n=%d eps=%g minPts=%d" % (trial, n, eps, mp)
        if eps >= 0:
            assert (c.launches - l0 == 1) == (n <= 1024)
            assert int(nc[0]) == (int(want.max()) + 1 if (want >= 0).any() else 0)
    pts = rng.integers(0, 5000, 40).astype(np.int32)
    lab = np.zeros(40, np.int32)
    t0 = time.perf_counter()
    for _ in range(2000):
        lib().csv_dbscan1d(c.h, ptr(pts), 40, 100.0, 5, ptr(lab), None)
    us_small = 1e6 * (time.perf_counter() - t0) / 2000
    print("one-launch DBSCAN1D fit of 40 points: %.1f us per call" % us_small)
    c.close()


def test_dbscan1d_segments(ctx, oracle):
    rng = np.random.default_rng(4)
    n, n_seg = 5000, 7
    seg = rng.integers(0, n_seg, n).astype(np.uint32)
    pts = (rng.integers(0, 40, n) * 500 + rng.integers(-20, 20, n)).astype(np.int32)
    lab, nc = api.dbscan1d_segments(pts, seg, n_seg, 15.0, 4, ctx)
    for sgm in range(n_seg):
        m = seg == sgm
        want = oracle.dbscan1d(pts[m], 15.0, 4)
        assert np.array_equal(lab[m], want)
        assert nc[sgm] == (want.max() + 1 if len(want) and want.max() >= 0 else 0)


def test_signature_clustering_on_device(ctx, oracle):
    clen = [900_000, 300_000]
    r = util.synth_reads(clen, seed=17, n_sv=250, coverage=30.0, sv_jitter_sd=10.0)
    b = run_batch(ctx, r, api.whole_contig_regions(clen))
    sg = b.sigs()
    lab = b.sigs_dbscan1d(100.0, 5)
    for tid in range(len(clen)):
        lo, hi = int(sg["region_off"][tid]), int(sg["region_off"][tid + 1])
        for is_del in (True, False):
            m = np.zeros(len(lab), bool); m[lo:hi] = True
            m &= (sg["kind"] == 1) if is_del else (sg["kind"] != 1)
            want = oracle.dbscan1d(sg["start"][m].astype(np.int32), 100.0, 5, fast=True)
            assert np.array_equal(lab[m], want)
    b.free()


def test_record_summary(ctx, oracle):
    """bam_endpos + getAlignmentReadPositions of every record (split-read pass): golden vectors of the compiled
    reference, adversarial CIGARs (all ops, empty records, unmapped flags), HiFi- and ONT-shaped records."""
    for i, r, keep, e, s, q in util.golden_record_summary_cases():
        b = api.Batch(ctx, r, api.whole_contig_regions([50000]))
        ge, gs, gq = b.record_summary()
        assert np.array_equal(ge[keep], e) and np.array_equal(gs[keep], s) and np.array_equal(gq[keep], q), i
        b.free()
    rng = np.random.default_rng(61)
    cases = [util.random_cigar_reads(rng, 700, [90000, 5000], n_tids=2, weird=True, max_ops=mo) for mo in (1, 12, 40, 300)]
    cases.append(util.synth_reads([400_000], seed=3, n_sv=60, frac_softclip=0.3))
    cases.append(util.synth_reads([900_000], seed=4, profile=1, coverage=5.0, read_len_mean=60000, indel_rate=0.1, n_sv=20))
    for r in cases:
        clen = [int(x) for x in r.get("contig_len", [90000, 5000])]
        b = api.Batch(ctx, r, api.whole_contig_regions(clen))
        got = b.record_summary()
        want = oracle.record_summary(r)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        b.free()


def test_depth_at_positions(ctx, oracle):
    """SVCaller::getReadDepth served from the device-resident map (sv_caller.cpp:1332-1344): 0 beyond the map."""
    clen = [120_000, 40_000]
    r = util.synth_reads(clen, seed=29, n_sv=30, coverage=25.0)
    b = run_batch(ctx, r, api.whole_contig_regions(clen))
    rng = np.random.default_rng(29)
    for t in range(2):
        d, _, _ = oracle.depth(r, t, clen[t] + 1)
        pos = np.concatenate([rng.integers(0, clen[t] + 1, 5000), [0, clen[t], clen[t] + 1, clen[t] + 500, 2**32 - 1]]).astype(np.uint32)
        want = np.where(pos <= clen[t], d[np.minimum(pos, clen[t])], 0)
        assert np.array_equal(b.depth_at(t, pos), want)
    b.free()


def test_mirror_classes(ctx, oracle):
    """The reference-shaped host interface (CNVCaller / SVCaller mirrors in contextsv_b200.api)."""
    clen = [80_000, 20_000]
    r = util.synth_reads(clen, seed=23, n_sv=40, coverage=15.0, frac_len50=0.3)
    rng = np.random.default_rng(23)
    seq4, seq_off = util.random_seq4(rng, r)
    aln = api.Alignments(r, ["chr21", "chrM"], clen, seq4, seq_off)
    depth_map = {"chr21": np.zeros(clen[0] + 1, np.uint32), "chrM": np.zeros(5, np.uint32)}   # chrM: wrong size -> resized
    mean_map = {}
    errors = []
    api.CNVCaller(ctx).calculateMeanChromosomeCoverage(["chr21", "chrM", "chrNope"], depth_map, mean_map, aln, 1, errors.append)
    for tid, chrom in enumerate(["chr21", "chrM"]):
        d, s, nz = oracle.depth(r, tid, clen[tid] + 1)
        assert np.array_equal(depth_map[chrom], d)
        assert mean_map[chrom] == oracle.mean_cov(s, nz)
    assert len(errors) == 2 and "chrNope" in errors[0] and "mismatch" in errors[1]
    calls = []
    api.SVCaller(ctx).findCIGARSVs(aln, "chr21", calls, depth_map["chr21"])
    o = oracle.cigar_scan(r, 0, clen[0] + 1)
    assert [(c.start, c.end) for c in calls] == [(int(x["start"]), int(x["end"])) for x in o]
    assert [c.alt_allele for c in calls] == [util.oracle_alt(seq4, seq_off, x) for x in o]
    assert any(len(c.alt_allele) == 50 for c in calls)

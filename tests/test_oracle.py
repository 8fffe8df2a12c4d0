"""CPU: the oracle (oracle/csv_oracle.c) against the committed golden vectors (generated from the
reference's own code by tests/golden/make_golden.py) and, where oracle/_ref exists, against the
compiled reference live."""
import numpy as np
import pytest

import util


def test_golden_dbscan1d(oracle):
    n = 0
    for i, pts, eps, mp, labels, largest in util.golden_db_cases():
        assert np.array_equal(oracle.dbscan1d(pts, eps, mp), labels), "literal restatement, case %d" % i
        assert np.array_equal(oracle.dbscan1d(pts, eps, mp, fast=True), labels), "closed form, case %d" % i
        assert np.array_equal(oracle.largest_cluster(pts, labels), largest), "largest cluster, case %d" % i
        n += 1
    assert n >= 100


def test_golden_depth_and_signatures(oracle):
    n = 0
    for i, r, clen, seq4, seq_off in util.golden_cg_cases():
        for tid in range(len(clen)):
            ans = util.golden_cg_answer(i, tid)
            d, s, nz = oracle.depth(r, tid, clen[tid] + 1)
            assert np.array_equal(d, ans["depth"]) and s == ans["sum"] and nz == ans["nonzero"], "depth case %d tid %d" % (i, tid)
            assert oracle.mean_cov(s, nz) == ans["mean"]
            for fast in (False, True):
                sg = oracle.cigar_scan(r, tid, clen[tid] + 1, fast=fast)
                assert np.array_equal(sg["start"], ans["start"]) and np.array_equal(sg["end"], ans["end"]), "sig case %d tid %d" % (i, tid)
                assert np.array_equal(sg["kind"], ans["kind"])
            assert [util.oracle_alt(seq4, seq_off, x) for x in sg] == ans["alt"], "ALT case %d tid %d" % (i, tid)
            n += len(sg)
    assert n > 500


def test_golden_log2_windows(oracle):
    seen = 0
    for i, r, clen, _, _ in util.golden_cg_cases():
        for tid in range(len(clen)):
            ans = util.golden_cg_answer(i, tid)
            if "l2" not in ans:
                continue
            a, b, ss = ans["l2par"]
            ws, we, su, cn, lg = oracle.log2_windows(ans["depth"], a, b, ss, ans["mean"])
            cent = (ws.astype(np.uint64) + we) // 2
            o = np.argsort(ans["l2pos"], kind="stable"); oo = np.argsort(cent, kind="stable")
            assert np.array_equal(cent[oo], ans["l2pos"][o].astype(np.uint64))
            assert np.array_equal(lg[oo], ans["l2"][o])            # same libm, same operations: bit-identical
            seen += 1
    assert seen >= 3


def test_oracle_vs_reference_dbscan_fuzz(oracle, reference):
    rng = np.random.default_rng(1)
    for it in range(1500):
        n = int(rng.integers(0, 70))
        span = int(rng.choice([5, 30, 200, 5000]))
        pts = rng.integers(-span, span, n).astype(np.int32)
        eps = float(rng.choice([-1, 0, 0.5, 1, 2, 3.7, 10, 50, 100, 1e12]))
        mp = int(rng.choice([-1, 0, 1, 2, 3, 5, 8]))
        a = reference.dbscan1d(pts, eps, mp)
        assert np.array_equal(a, oracle.dbscan1d(pts, eps, mp))
        assert np.array_equal(a, oracle.dbscan1d(pts, eps, mp, fast=True))
        assert np.array_equal(reference.largest_cluster(pts, eps, mp), oracle.largest_cluster(pts, a))


def test_oracle_vs_reference_reads_fuzz(oracle, reference):
    rng = np.random.default_rng(2)
    for it in range(25):
        clen = [int(rng.choice([300, 2500, 12000])) for _ in range(int(rng.integers(1, 4)))]
        r = util.random_cigar_reads(rng, int(rng.integers(0, 150)), clen, n_tids=len(clen), weird=bool(it % 2))
        seq4, seq_off = util.random_seq4(rng, r)
        for tid in range(len(clen)):
            d, s, nz, mean = reference.depth(r, tid, clen)
            d2, s2, nz2 = oracle.depth(r, tid, clen[tid] + 1)
            assert np.array_equal(d, d2) and (s, nz) == (s2, nz2) and mean == oracle.mean_cov(s2, nz2)
            st, en, ty, ev, alts = reference.cigar_scan(r, tid, clen, seq4=seq4, seq_off=seq_off)
            sg = oracle.cigar_scan(r, tid, clen[tid] + 1, fast=bool(it % 2))
            assert np.array_equal(st, sg["start"]) and np.array_equal(en, sg["end"])
            assert np.array_equal(np.where(ev == 1, 0, np.where(ev == 2, 1, 2)), sg["kind"])
            assert alts == [util.oracle_alt(seq4, seq_off, x) for x in sg]


def test_reference_depth_map_resize(oracle, reference):
    """cnv_caller.cpp:482-487: a caller-side size that disagrees with the BAM header is resized."""
    r = util.synth_reads([30000], seed=3, n_sv=5, coverage=5.0)
    d, s, nz, mean = reference.depth(r, 0, [30000], alloc_size=1234)
    d2, s2, nz2 = oracle.depth(r, 0, 30001)
    assert np.array_equal(d, d2) and s == s2 and nz == nz2


def test_dbscan2d_restatement_against_golden_and_reference():
    """orc_dbscan2d (dbscan.cpp:9-81 restated) vs the committed outputs of the compiled reference, and live where
    oracle/_ref exists -- degenerate eps / min_pts / zero lengths included."""
    from oracle.oracle_py import Oracle, ref_available, Reference
    O = Oracle()
    for i, st, en, eps, mp, want in util.golden_db2_cases():
        assert np.array_equal(O.dbscan2d(st, en, eps, mp), want), i
    if not ref_available():
        return
    R = Reference()
    rng = np.random.default_rng(8)
    for it in range(150):
        n = int(rng.integers(0, 100))
        st = rng.integers(0, 800, n).astype(np.uint32)
        en = (st + rng.choice([0, 1, 20, 50, 200], n)).astype(np.uint32)
        eps = float(rng.choice([-0.5, 0, 0.1, 0.4, 0.99, 1.0, 2.0])); mp = int(rng.choice([-2, 0, 1, 2, 4]))
        assert np.array_equal(O.dbscan2d(st, en, eps, mp), R.dbscan2d(st, en, eps, mp)), (it, eps, mp)


def test_record_summary_restatement_against_golden():
    """orc_record_summary (getAlignmentReadPositions, sv_caller.cpp:663-690, + bam_endpos) vs the compiled reference's
    committed outputs."""
    from oracle.oracle_py import Oracle
    O = Oracle()
    n = 0
    for i, r, keep, e, s, q in util.golden_record_summary_cases():
        ge, gs, gq = O.record_summary(r)
        assert np.array_equal(ge[keep], e) and np.array_equal(gs[keep], s) and np.array_equal(gq[keep], q), i
        n += 1
    assert n >= 10

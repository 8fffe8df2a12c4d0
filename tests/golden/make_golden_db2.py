"""Golden vectors for the 2-D DBSCAN::fit (src/dbscan.cpp:9-81), generated from the UNMODIFIED reference compiled into
oracle/_ref (make -C oracle ref).  Run in the build container:  python tests/golden/make_golden_db2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle_py import Reference  # noqa: E402


def cases(rng):
    for it in range(80):
        n = int(rng.integers(0, 160))
        centers = rng.integers(100, 6000, max(1, n // 7 + 1))
        st = (rng.choice(centers, n) + rng.integers(-20, 20, n)).astype(np.int64)
        ln = rng.choice([0, 1, 30, 50, 60, 100, 400, 3000], n) + rng.integers(0, 9, n)
        if it % 4:
            ln = np.maximum(ln, 1)                      # mergeSVs-shaped: positive lengths
        eps = float(rng.choice([-1, 0, 0.05, 0.1, 0.1, 0.3, 0.5, 0.9, 1.0, 1.5]))
        mp = int(rng.choice([-1, 0, 1, 2, 2, 3, 5, 8]))
        yield st.astype(np.uint32), (st + ln).astype(np.uint32), eps, mp


def main():
    R = Reference()
    rng = np.random.default_rng(20261018)
    out = {}
    k = 0
    for st, en, eps, mp in cases(rng):
        out["c%d_start" % k] = st; out["c%d_end" % k] = en
        out["c%d_par" % k] = np.array([eps, mp], np.float64)
        out["c%d_labels" % k] = R.dbscan2d(st, en, eps, mp)
        k += 1
    out["n"] = np.array([k])
    # record summaries of the split-read pass (bam_endpos + getAlignmentReadPositions, sv_caller.cpp:150-162, 663-690)
    sys.path.insert(0, os.path.dirname(HERE))
    import util
    nrs = 0
    for it in range(12):
        clen = [int(rng.choice([3000, 40000]))]
        r = util.random_cigar_reads(rng, int(rng.integers(1, 150)), clen, n_tids=1, weird=False, max_ops=int(rng.choice([1, 4, 12, 40])))
        keep = np.nonzero(r["pos0"] < clen[0])[0]                    # what sam_itr_querys(contig) yields
        e, s_, q = R.record_summary(r, 0, clen)
        assert len(e) == len(keep)
        p = "rs%d_" % nrs
        for key in ("pos0", "flag", "mapq", "cig_off", "cigar"):
            out[p + key] = np.asarray(r[key])
        out[p + "keep"] = keep; out[p + "endpos"] = e; out[p + "qstart"] = s_; out[p + "qend"] = q
        nrs += 1
    out["n_rs"] = np.array([nrs])
    np.savez_compressed(os.path.join(HERE, "golden_db2_v1.npz"), **out)
    print("wrote", k, "cases")


if __name__ == "__main__":
    main()

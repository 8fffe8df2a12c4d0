"""Generates tests/golden/golden_v1.npz from the reference's OWN code (oracle/_ref:
ContextSV sources compiled unmodified + htslib shim, see oracle/Makefile).

Run in the build container only (needs /root/reference to have been compiled):
    python tests/golden/make_golden.py
The vectors are committed; the GPU box never needs the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle_py import Reference  # noqa: E402
import util  # noqa: E402


def reads_to_npz(prefix, r, out):
    out[prefix + "_pos0"] = r["pos0"]; out[prefix + "_flag"] = r["flag"]; out[prefix + "_mapq"] = r["mapq"]
    out[prefix + "_cig_off"] = r["cig_off"]; out[prefix + "_cigar"] = r["cigar"]
    out[prefix + "_tid"] = r["tid"] if r.get("tid") is not None else np.zeros(0, np.int32)


def main():
    R = Reference()
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- DBSCAN1D known answers
    n_db = 0
    for it in range(120):
        n = int(rng.integers(0, 80)) if it % 10 else int(rng.integers(200, 600))
        span = int(rng.choice([5, 40, 300, 5000, 100000]))
        if it % 3 == 0:   # clustered
            centers = rng.integers(-span, span, max(1, n // 12))
            pts = (rng.choice(centers, n) + rng.integers(-30, 30, n)).astype(np.int32)
        else:
            pts = rng.integers(-span, span, n).astype(np.int32)
        eps = float(rng.choice([-1.0, 0.0, 0.5, 1.0, 2.0, 3.7, 10.0, 50.0, 100.0, 1000.0, 1e12]))
        mp = int(rng.choice([-1, 0, 1, 2, 3, 5, 5, 8, 20]))
        out["db%d_pts" % n_db] = pts
        out["db%d_par" % n_db] = np.array([eps, mp], np.float64)
        out["db%d_labels" % n_db] = R.dbscan1d(pts, eps, mp)
        out["db%d_largest" % n_db] = R.largest_cluster(pts, eps, mp)
        n_db += 1
    out["n_db"] = np.array([n_db])
    # ---- CIGAR / depth known answers
    n_cg = 0
    cases = []
    for it in range(10):
        L = int(rng.choice([400, 3000, 20000]))
        cases.append(("adv", util.random_cigar_reads(rng, int(rng.integers(1, 120)), [L], weird=(it % 2 == 1)), [L]))
    cases.append(("adv2", util.random_cigar_reads(rng, 150, [5000, 9000, 700], n_tids=3, weird=True), [5000, 9000, 700]))
    cases.append(("syn", util.synth_reads([120000], seed=7, n_sv=40, coverage=20.0, frac_len50=0.3, frac_softclip=0.1), [120000]))
    cases.append(("syn_eqx", util.synth_reads([60000, 45000], seed=8, n_sv=30, coverage=15.0, use_eqx=1, frac_len50=0.2), [60000, 45000]))
    for name, r, clen in cases:
        from oracle.oracle_py import norm_reads
        r = norm_reads(r)
        seq4, seq_off = util.random_seq4(rng, r)
        p = "cg%d" % n_cg
        reads_to_npz(p, r, out)
        out[p + "_clen"] = np.array(clen, np.uint32)
        out[p + "_seq4"] = seq4; out[p + "_seq_off"] = seq_off
        for tid in range(len(clen)):
            d, s, nz, mean = R.depth(r, tid, clen)
            st, en, ty, ev, alts = R.cigar_scan(r, tid, clen, seq4=seq4, seq_off=seq_off)
            q = "%s_t%d" % (p, tid)
            out[q + "_depth"] = d; out[q + "_stats"] = np.array([s, nz], np.uint64); out[q + "_mean"] = np.array([mean])
            out[q + "_start"] = st; out[q + "_end"] = en; out[q + "_svtype"] = ty; out[q + "_evidence"] = ev
            out[q + "_alt"] = np.array(alts, dtype="U64") if alts else np.zeros(0, "U64")
            if mean > 0 and clen[tid] > 2500:
                a, b = 200, min(clen[tid] - 10, 200 + int(rng.integers(2000, clen[tid] - 300)))
                pos, lg = R.log2_windows(d, a, b, 20, mean)
                out[q + "_l2par"] = np.array([a, b, 20], np.int64); out[q + "_l2pos"] = pos; out[q + "_l2"] = lg
        n_cg += 1
    out["n_cg"] = np.array([n_cg])
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", n_db, "dbscan cases,", n_cg, "cigar cases")


if __name__ == "__main__":
    main()

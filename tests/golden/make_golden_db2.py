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
    np.savez_compressed(os.path.join(HERE, "golden_db2_v1.npz"), **out)
    print("wrote", k, "cases")


if __name__ == "__main__":
    main()

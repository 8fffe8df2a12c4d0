"""DBSCAN::fit (2-D, mergeSVs' clustering) on signature-shaped interval sets: GPU through the C ABI (host buffers in and
out) vs the unmodified reference (oracle/_ref, O(N^2)).  SURVEY.md 8f-1 / BASELINE config 5."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from contextsv_b200 import api
from oracle.oracle_py import Reference, ref_available

def make(n, rng):
    k = max(1, n // 25)
    centers = np.sort(rng.integers(0, 240_000_000, k)); lens = rng.integers(50, 10000, k)
    pick = rng.integers(0, k, n)
    st = (centers[pick] + np.rint(rng.normal(0, 10, n)).astype(np.int64)).clip(1).astype(np.uint32)
    en = (st + lens[pick] + np.rint(rng.normal(0, 8, n)).astype(np.int64).clip(-40) + 1).astype(np.uint32)
    o = np.lexsort((en, st))
    return st[o], en[o]

ctx = api.Context(0)
rng = np.random.default_rng(5)
R = Reference() if ref_available() else None
for n in (20_000, 40_000, 100_000, 1_000_000, 3_000_000):
    st, en = make(n, rng)
    db = api.DBSCAN(0.1, 2, ctx)
    db.fit(st, en)
    t0 = time.perf_counter()
    for _ in range(3): db.fit(st, en)
    ctx.sync(); t_gpu = (time.perf_counter() - t0) / 3
    line = "n=%8d  gpu %8.2f ms  clusters=%d noise=%d" % (n, 1e3 * t_gpu, int(db.getClusters().max()) + 1, int((db.getClusters() == -2).sum()))
    if R is not None and n <= 40_000:
        t0 = time.perf_counter(); want = R.dbscan2d(st, en, 0.1, 2); t_ref = time.perf_counter() - t0
        line += "  reference %8.1f ms (x%.0f)  equal=%s" % (1e3 * t_ref, t_ref / t_gpu, bool(np.array_equal(want, db.getClusters())))
    print(line, flush=True)

import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, util
from contextsv_b200 import api
from oracle.oracle_py import Oracle
ctx = api.Context(0); O = Oracle()
rng = np.random.default_rng(int(os.environ.get("SEED", "1234")))
bad = 0
for it in range(400):
    clen = [int(rng.choice([300, 2500, 8192, 8193, 12000, 70000, 200000])) for _ in range(int(rng.integers(1, 4)))]
    n = int(rng.integers(0, 1500))
    r = util.random_cigar_reads(rng, n, clen, n_tids=len(clen), weird=bool(it % 2), max_ops=int(rng.choice([1, 3, 12, 40, 300])), big_p=float(rng.choice([0.05, 0.25, 0.6])))
    regions = api.whole_contig_regions(clen)
    b = api.Batch(ctx, r, regions); b.scan()
    sums, nzs = b.depth_stats(); sg = b.sigs()
    for t in range(len(clen)):
        d, s, nz = O.depth(r, t, clen[t] + 1)
        o = O.cigar_scan(r, t, clen[t] + 1)
        lo, hi = int(sg["region_off"][t]), int(sg["region_off"][t + 1])
        ok = np.array_equal(b.depth(t), d) and int(sums[t]) == s and int(nzs[t]) == nz and hi - lo == len(o) and all(np.array_equal(sg[f][lo:hi], o[f]) for f in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"))
        if not ok: bad += 1; print("MISMATCH it", it, "tid", t, clen, n)
    b.free()
print("fuzz done, mismatches:", bad)

"""Region shards in scrambled order x long reads (searched tile path) x pipelined chunks: depth slices must concatenate to
the oracle's map, stats must add up, every signature must come from the region that owns its read."""
import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, util
from contextsv_b200 import api
from oracle.oracle_py import Oracle
ctx = api.Context(0); O = Oracle()
rng = np.random.default_rng(int(os.environ.get("SEED", "7")))
bad = 0
for it in range(24):
    clen = [int(rng.integers(200_000, 900_000)) for _ in range(int(rng.integers(1, 4)))]
    ont = it % 2 == 0
    kw = dict(profile=1, coverage=float(rng.choice([8, 25])), read_len_mean=float(rng.choice([20000, 60000])), indel_rate=0.08, indel_len_max=4) if ont else dict(coverage=float(rng.choice([20, 45])))
    r = util.synth_reads(clen, seed=1000 + it, n_sv=int(rng.integers(10, 120)), **kw)
    regions = []
    for t, L in enumerate(clen):
        cuts = sorted(set([0, L + 1] + [int(x) for x in rng.integers(1, L, int(rng.integers(0, 5)))]))
        regions += [(t, cuts[i], cuts[i + 1], L + 1) for i in range(len(cuts) - 1)]
    perm = rng.permutation(len(regions))
    regs = [regions[p] for p in perm]
    ctx.set_pipeline_chunks(int(rng.choice([1, 3])))
    b = api.Batch(ctx, r, regs); b.scan()
    sums, nzs = b.depth_stats(); sg = b.sigs()
    for t, L in enumerate(clen):
        d, s, nz = O.depth(r, t, L + 1)
        o = O.cigar_scan(r, t, L + 1)
        got = np.zeros(L + 1, np.uint32); ssum = 0; snz = 0; parts = []
        for j, (tt, beg, end, _) in enumerate(regs):
            if tt != t: continue
            got[beg:end] = b.depth(j); ssum += int(sums[j]); snz += int(nzs[j])
            lo, hi = int(sg["region_off"][j]), int(sg["region_off"][j + 1])
            idx = r["pos0"][sg["read_idx"][lo:hi]].astype(np.int64) + 1
            if not np.all((idx >= beg) & ((idx < end) | (end == L + 1))): bad += 1; print("OWNERSHIP", it, t, j)
            parts.append({k: sg[k][lo:hi] for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")})
        m = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]} if parts else None
        ok = np.array_equal(got, d) and ssum == s and snz == nz
        if m is not None:
            order = np.lexsort((-(m["read_idx"].astype(np.int64) * (1 << 24) + m["op_idx"]), m["end"], m["start"]))
            ok = ok and len(order) == len(o) and all(np.array_equal(m[f][order], o[f]) for f in m)
        if not ok: bad += 1; print("MISMATCH", it, t, ont, len(regs))
    b.free()
ctx.set_pipeline_chunks(1)
print("region fuzz done, mismatches:", bad)

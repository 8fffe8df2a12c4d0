"""Diagnostic: the region shards of a strong-scaling run, scanned one after the other on ONE device, against the
whole-genome scan: per-contig depth checksum, sum, non-zero count.  python scripts/diag_strong.py [world] [workload]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from contextsv_b200 import api, shard, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
workload = sys.argv[2] if len(sys.argv) > 2 else "wgs30x"
contig_len, n_sv = bench.workload_contigs(workload)
full = synth.generate(contig_len, seed=20261020, n_sv=n_sv, **bench.SYNTH_KW.get(workload, {}))
ctx = api.Context(0)
bt = api.Batch(ctx, full, api.whole_contig_regions(contig_len))
bt.scan(want_depth=True, want_sigs=True)
c1 = bt.depth_checksum().astype(np.int64); s1, z1 = bt.depth_stats()
s1 = s1.astype(np.int64); z1 = z1.astype(np.int64)
keep = {}
plan = shard.plan_regions(contig_len, world, full)
for regs in plan:
    for (t, b, e, m) in regs:
        if (b, e) != (0, m):
            keep[(t, b, e)] = None
for i, (t, _, _, _) in enumerate(api.whole_contig_regions(contig_len)):
    for key in keep:
        if key[0] == t:
            keep[key] = bt.depth(i)[key[1]:key[2]].copy()
bt.free()
cs = np.zeros(len(contig_len), np.int64); ss = np.zeros(len(contig_len), np.int64); zs = np.zeros(len(contig_len), np.int64)
for rank, regs in enumerate(plan):
    sub, base = shard.select_reads(full, regs)
    if os.environ.get("DIAG_PINNED"):
        from contextsv_b200 import _capi
        sub = {k: (lambda p, v: (p.__setitem__(slice(None), v), p)[1])(_capi.pinned_empty(len(v), v.dtype), v) if isinstance(v, np.ndarray) else v for k, v in sub.items()}
    b = api.Batch(ctx, sub, regs)
    for _ in range(int(os.environ.get("DIAG_REPEAT", "0"))):          # what bench.py does before its e2e step: scans, free, same geometry again from the pool
        b.scan(want_depth=True, want_sigs=True); b.sigs_dbscan1d(100.0, 5, fetch=False)
    if os.environ.get("DIAG_REPEAT"):
        ctx.sync(); b.free(); b = api.Batch(ctx, sub, regs)
    b.scan(want_depth=True, want_sigs=True)
    ck = b.depth_checksum().astype(np.int64); s, z = b.depth_stats()
    for i, (t, bb, ee, m) in enumerate(regs):
        cs[t] += ck[i]; ss[t] += int(s[i]); zs[t] += int(z[i])
        if (bb, ee) == (0, m) and (int(s[i]) != int(s1[t]) or int(z[i]) != int(z1[t])):
            print("shard %d region %s: sum %d / %d nz %d / %d   <-- DIFFERS" % (rank, (t, bb, ee, m), int(s[i]), int(s1[t]), int(z[i]), int(z1[t])))
        if (t, bb, ee) in keep:
            d = b.depth(i)
            bad = np.nonzero(d != keep[(t, bb, ee)])[0]
            print("shard %d region %s: %d positions differ%s" % (rank, (t, bb, ee, m), len(bad), "" if not len(bad) else " first at +%d (abs %d): %d vs %d" % (bad[0], bb + bad[0], d[bad[0]], keep[(t, bb, ee)][bad[0]])))
    b.free()
for t in range(len(contig_len)):
    flag = "" if (cs[t] == c1[t] and ss[t] == s1[t] and zs[t] == z1[t]) else "   <-- DIFFERS"
    print("contig %2d checksum %016x / %016x  sum %d / %d  nz %d / %d%s" % (t, int(cs[t]) & (2**64 - 1), int(c1[t]) & (2**64 - 1), ss[t], s1[t], zs[t], z1[t], flag))

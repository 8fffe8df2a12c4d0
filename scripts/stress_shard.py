"""Diagnostic: scans ONE region shard of a strong-scaling plan over and over the way bench.py's e2e step does (fresh batch
from the pool, scan, DBSCAN1D, fetches) and compares the per-region depth checksums with the first pass.  On a mismatch
the intermediate arrays of the bad pass (events, ev_start, ref_end, span_desc, pmax, tile_q) are compared with a re-scan.

python scripts/stress_shard.py <world> <rank> [iterations] [workload]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from contextsv_b200 import api, shard, synth, _capi

world = int(sys.argv[1]); rank = int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
workload = sys.argv[4] if len(sys.argv) > 4 else "wgs30x"
contig_len, n_sv = bench.workload_contigs(workload)
full = synth.generate(contig_len, seed=20261020, n_sv=n_sv, **bench.SYNTH_KW.get(workload, {}))
if world > 1:
    plan = shard.plan_regions(contig_len, world, full)
    regs = plan[rank]
    sub, base = shard.select_reads(full, regs)
else:
    regs = api.whole_contig_regions(contig_len); sub = full
reads = {}
for k, v in sub.items():
    if isinstance(v, np.ndarray):
        p = _capi.pinned_empty(len(v), v.dtype); p[:] = v; reads[k] = p
    else:
        reads[k] = v
del sub, full
ctx = api.Context(int(os.environ.get("LOCAL_RANK", "0")))
NAMES = [("events", np.uint32), ("ev_start", np.uint32), ("ref_end", np.uint32), ("span_desc", np.uint32), ("pmax", np.uint64), ("tile_q", np.uint32)]


def one_pass(b, dbscan=True):
    b.scan(want_depth=True, want_sigs=True)
    if dbscan:
        b.sigs_dbscan1d(100.0, 5, fetch=False)
    s, z = b.depth_stats()
    return b.depth_checksum().copy(), s.copy(), z.copy()


b = api.Batch(ctx, reads, regs)
c0, s0, z0 = one_pass(b)
for _ in range(3):
    c, s, z = one_pass(b)
    if not np.array_equal(c, c0) or not np.array_equal(s, s0):
        w = np.nonzero((c != c0) | (s != s0))[0]
        print("warm-up pass differs from the first pass: regions %s sum delta %s nz delta %s" % ([regs[i] for i in w], [int(s[i]) - int(s0[i]) for i in w], [int(z[i]) - int(z0[i]) for i in w]), flush=True)
print("reference pass: %d regions, sum %d" % (len(regs), int(s0.astype(np.uint64).sum())), flush=True)
if os.environ.get("STRESS_TIME"):
    # resident passes of this shard alone, timed like bench.py's value (what one rank of an N-way run does per step)
    for _ in range(3):
        b.scan(want_depth=True, want_sigs=True); b.sigs_dbscan1d(100.0, 5, fetch=False)
    ctx.sync(); ctx.profile_read(reset=True); ctx.profile_enable(True)
    n_t = 20
    ctx.timer_begin()
    for _ in range(n_t):
        b.scan(want_depth=True, want_sigs=True); b.sigs_dbscan1d(100.0, 5, fetch=False)
    ms = ctx.timer_end()
    ctx.profile_enable(False)
    st = ctx.profile_read(reset=True)
    print("shard %d/%d: %.4f ms/step  %s" % (rank, world, ms / n_t, {k: round(v[0] / n_t, 4) for k, v in st.items()}), flush=True)
    sys.exit(0)
bad_total = 0
t0 = time.time()
for it in range(iters):
    mode = it % 4
    if mode == 0:                       # what bench.py does: resident scans, free, same geometry again from the pool
        for _ in range(5):
            b.scan(want_depth=True, want_sigs=True); b.sigs_dbscan1d(100.0, 5, fetch=False)
        ctx.sync(); b.free(); b = api.Batch(ctx, reads, regs)
    elif mode == 1:
        b.free(); b = api.Batch(ctx, reads, regs)
    elif mode == 2 and os.environ.get("STRESS_SLEEP"):
        time.sleep(0.002)
    c, s, z = one_pass(b, dbscan=(mode != 3))
    if np.array_equal(c, c0) and np.array_equal(s, s0):
        continue
    bad_total += 1
    bad_regs = np.nonzero((c != c0) | (s != s0))[0]
    print("iteration %d (mode %d): regions %s differ: sum delta %s" % (it, mode, [regs[i] for i in bad_regs], [int(s[i]) - int(s0[i]) for i in bad_regs]), flush=True)
    if bad_total > 3:
        continue
    snap = {n: b.debug_array(n, dt) for n, dt in NAMES}
    dep_bad = {int(i): b.depth(int(i)).copy() for i in bad_regs}
    c2, s2, z2 = one_pass(b)
    print("  re-scan of the same batch: %s" % ("good" if np.array_equal(c2, c0) else "still differs"), flush=True)
    for i, d in dep_bad.items():
        g = b.depth(i)
        w = np.nonzero(g != d)[0]
        if len(w):
            print("  region %s: %d positions differ, first +%d last +%d (abs %d..%d), bad-good values at first: %d vs %d, distinct deltas %s" % (
                regs[i], len(w), w[0], w[-1], regs[i][1] + w[0], regs[i][1] + w[-1], d[w[0]], g[w[0]], np.unique(d[w].astype(np.int64) - g[w].astype(np.int64))[:8]), flush=True)
            runs = np.nonzero(np.diff(w) != 1)[0]
            print("  runs: %d; first run [%d, %d)" % (len(runs) + 1, w[0], (w[runs[0]] + 1) if len(runs) else w[-1] + 1), flush=True)
    for n, dt in NAMES:
        g = b.debug_array(n, dt)
        w = np.nonzero(g != snap[n])[0]
        if len(w):
            print("  %s: %d entries differ, first %d last %d; bad %s good %s" % (n, len(w), w[0], w[-1], snap[n][w[:6]], g[w[:6]]), flush=True)
            if n == "events":
                es = snap["ev_start"]
                k = np.searchsorted(es, w[0], side="right") - 1
                print("    first differing event slot %d belongs to compact record %d (slots %d..%d); span_desc of that neighbourhood not decoded here" % (w[0], k, es[k], es[k + 1]), flush=True)
        else:
            print("  %s: identical" % n, flush=True)
print("%d iterations, %d bad, %.1f s" % (iters, bad_total, time.time() - t0), flush=True)

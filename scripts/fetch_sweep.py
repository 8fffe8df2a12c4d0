"""Depth fetch of the whole-genome batch (12.4 GB of uint32 in host memory) for several thread counts / chunk sizes,
interleaved so that box-to-box noise hits every setting alike.  GPU box only."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contextsv_b200 import _capi, api, shard, synth

contig_len = [l for _, l in shard.GRCH38]
reads = synth.generate(contig_len, alloc=_capi.pinned_empty, seed=1, n_sv=25000)
regions = api.whole_contig_regions(contig_len)
ctx = api.Context(0)
bt = api.Batch(ctx, reads, regions)
bt.scan(want_depth=True, want_sigs=False)
pinned = [_capi.pinned_empty(e - b, np.uint32) for (_, b, e, _) in regions]
pageable = [np.zeros(e - b, np.uint32) for (_, b, e, _) in regions]
settings = [(0, 1 << 20), (8, 1 << 20), (12, 1 << 20), (14, 1 << 20), (16, 1 << 20), (14, 512 << 10), (14, 256 << 10), (14, 2 << 20), (14, 4 << 20), (12, 512 << 10)]
best = {}
for rep in range(4):
    for dst_name, dst in (("pinned", pinned), ("pageable", pageable)):
        for th, ch in settings:
            ctx.set_fetch(threads=th, chunk_positions=ch)
            t0 = time.perf_counter(); bt.depth_all(dst); dt = time.perf_counter() - t0
            k = (dst_name, th, ch)
            best.setdefault(k, []).append(dt)
for k, v in best.items():
    print("%-8s threads %2d chunk %4d Ki: min %6.1f ms  median %6.1f ms   (%.1f GB/s into host memory)" % (k[0], k[1], k[2] >> 10, 1e3 * min(v), 1e3 * sorted(v)[len(v) // 2], 12.35 / min(v)))
assert all(np.array_equal(a, b) for a, b in zip(pinned, pageable))

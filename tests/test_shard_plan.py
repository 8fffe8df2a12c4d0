"""Host logic of the op-budget shard planner (shard.plan_by_ops): no GPU needed."""
import numpy as np
import pytest

import util
from contextsv_b200 import shard


def _check_plan(r, clen, plans, max_ops):
    ends = shard.ref_end(r)
    cov = {t: [] for t in range(len(clen))}
    owned = np.zeros(int(r["n_reads"]), np.int32)
    idx = r["pos0"].astype(np.int64) + 1
    tid = r["tid"].astype(np.int64)
    for regions in plans:
        sub, base = shard.select_reads(r, regions, ends)
        assert int(sub["n_ops"]) + int(sub["n_reads"]) <= max_ops
        for (t, b, e, m) in regions:
            assert m == clen[t] + 1 and 0 <= b < e <= m
            cov[t].append((b, e))
            own = (tid == t) & (idx >= b) & ((idx < e) | (e == m))
            owned += own
            # every record overlapping [b, e) is inside the selected slice
            ov = np.nonzero((tid == t) & (idx < e) & (ends > b))[0]
            if len(ov):
                assert ov[0] >= base and ov[-1] < base + int(sub["n_reads"])
    for t, c in cov.items():
        c.sort()
        assert c[0][0] == 0 and c[-1][1] == clen[t] + 1 and all(c[i][1] == c[i + 1][0] for i in range(len(c) - 1))
    assert np.all(owned == 1)


def test_plan_by_ops_tiles_the_genome_and_respects_the_budget():
    clen = [700_000, 90_000, 400_000]
    r = util.synth_reads(clen, seed=3, profile=1, coverage=10.0, read_len_mean=20000, indel_rate=0.05, n_sv=30)
    for max_ops in (int(r["n_ops"]) + int(r["n_reads"]) + 1, 400_000, 150_000):
        plans = shard.plan_by_ops(r, clen, max_ops)
        _check_plan(r, clen, plans, max_ops)
    assert len(shard.plan_by_ops(r, clen, int(r["n_ops"]) + int(r["n_reads"]) + 1)) == 1
    assert len(shard.plan_by_ops(r, clen, 150_000)) > 5


def test_plan_by_ops_pileup_cannot_be_split():
    from oracle.oracle_py import make_reads
    pos0 = np.full(100, 500, np.int32)
    r = make_reads(pos0, [[(10, 0), (2, 2), (10, 0)]] * 100, tid=np.zeros(100, np.int32))
    assert len(shard.plan_by_ops(r, [5000], 1000)) == 1          # one position group: never split
    r2 = make_reads(np.arange(100, dtype=np.int32) * 3, [[(10, 0), (2, 2), (10, 0)]] * 100, tid=np.zeros(100, np.int32))
    with pytest.raises(ValueError):
        shard.plan_by_ops(r2, [5000], 10)

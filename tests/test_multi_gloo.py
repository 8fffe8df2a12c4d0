"""CPU, world_size 2, gloo: the host side of the multi-GPU path -- region planning, halo-read
selection, per-rank scan, gather and merge -- with the oracle standing in for the GPU scan of each rank.
The path has no data-path collective: the only communication is the final gather of results on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401  (sys.path)
from contextsv_b200 import shard, synth
from oracle.oracle_py import Oracle

CLEN = [260_000, 90_000, 40_000]


def scan_shard_with_oracle(O, sub, regions):
    """What api.Batch(...).scan() returns for a rank's regions: depth slices, stats, owned signatures."""
    depth, sums, nzs = [], [], []
    sg = {k: [] for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")}
    region_off = [0]
    for (tid, beg, end, ms) in regions:
        d, _, _ = O.depth(sub, tid, ms)
        sl = d[beg:end]
        depth.append(sl); sums.append(int(sl.sum(dtype=np.uint64))); nzs.append(int((sl > 0).sum()))
        o = O.cigar_scan(sub, tid, ms)
        idx = sub["pos0"][o["read_idx"]].astype(np.int64) + 1
        own = (idx >= beg) & ((idx < end) | (end == ms))
        for k in sg:
            sg[k].append(o[k][own])
        region_off.append(region_off[-1] + int(own.sum()))
    out = {k: np.concatenate(v) if v else np.zeros(0, np.uint32) for k, v in sg.items()}
    out["region_off"] = np.array(region_off, np.uint64)
    return depth, sums, nzs, out


def worker(rank, world, port, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = Oracle()
    reads = synth.generate(CLEN, seed=seed, n_sv=90, coverage=15.0, threads=1)
    regions = shard.plan_regions(CLEN, world)[rank]
    sub, base = shard.select_reads(reads, regions)
    depth, sums, nzs, sg = scan_shard_with_oracle(O, sub, regions)
    gathered = [None] * world
    dist.gather_object((regions, depth, sums, nzs, sg, base), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        # host merge: depth slices concatenate, stats add, signatures merge with the addSVCall comparator
        full_depth = {t: np.zeros(CLEN[t] + 1, np.uint32) for t in range(len(CLEN))}
        tot = {t: [0, 0] for t in range(len(CLEN))}
        parts = []
        for regs, dep, su, nz, s, b in gathered:
            for (tid, beg, end, ms), d, a, c in zip(regs, dep, su, nz):
                full_depth[tid][beg:end] = d; tot[tid][0] += a; tot[tid][1] += c
            parts.append((s, regs, b))
        merged = shard.merge_signatures(parts)
        ok = True
        for t in range(len(CLEN)):
            d, s, nz = O.depth(reads, t, CLEN[t] + 1)
            o = O.cigar_scan(reads, t, CLEN[t] + 1)
            ok &= np.array_equal(full_depth[t], d) and tot[t] == [s, nz]
            m = merged.get(t, {k: np.zeros(0) for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos")})
            for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
                ok &= np.array_equal(np.asarray(m[k]).astype(np.int64), o[k].astype(np.int64))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def worker_owner_merge(rank, world, port, seed, q):
    """The sharded bench's flow: no rank 0 bottleneck.  Every contig is finalised by the shard that holds its FIRST region:
    contigs it holds whole are its own run; a contig the plan cut is merged there from the holders' runs."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = Oracle()
    reads = synth.generate(CLEN, seed=seed, n_sv=90, coverage=15.0, threads=1)
    plan = shard.plan_regions(CLEN, world, reads)                 # cost-balanced cuts
    regions = plan[rank]
    sub, base = shard.select_reads(reads, regions)
    _, _, _, sg = scan_shard_with_oracle(O, sub, regions)
    board = [None] * world                                        # what the shared-memory board of bench.py holds
    dist.all_gather_object(board, (sg, base))
    holders = shard.contig_holders(plan)
    final = {}
    for t, ranks in holders.items():
        if ranks[0] != rank:
            continue
        parts = [p for r in ranks for p in shard.contig_part(board[r][0], plan[r], board[r][1], t)]
        final[t] = shard.merge_signatures(parts)[t]
    collected = [None] * world
    dist.gather_object(final, collected if rank == 0 else None, dst=0)
    if rank == 0:
        ok = any(len(r) > 1 for r in holders.values())            # the plan must cut a contig, or the test shows nothing
        owned = {}
        for f in collected:
            assert not (set(f) & set(owned))
            owned.update(f)
        for t in range(len(CLEN)):
            o = O.cigar_scan(reads, t, CLEN[t] + 1)
            ok &= t in owned
            for k in ("start", "end", "kind", "read_idx", "op_idx", "query_pos"):
                ok &= np.array_equal(np.asarray(owned[t][k]).astype(np.int64), o[k].astype(np.int64))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_owner_finalises_cut_contigs_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker_owner_merge, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_two_rank_region_sharding_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, 31, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_plan_regions_properties():
    for n in (1, 2, 3, 4, 8, 13):
        plan = shard.plan_regions([l for _, l in shard.GRCH38], n)
        assert len(plan) == n
        total = sum(l + 1 for _, l in shard.GRCH38)
        sizes = [sum(e - b for (_, b, e, _) in regs) for regs in plan]
        assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
        flat = [r for regs in plan for r in regs]
        assert flat == sorted(flat)                      # genome order, disjoint
        for a, b in zip(flat, flat[1:]):
            assert a[0] < b[0] or a[2] <= b[1]

#!/usr/bin/env python
"""bench.py -- aligned reads/sec through the alignment-scan hot path (CIGAR scan + depth + DBSCAN1D).

  python bench.py --gpus N --steps K --warmup W [--workload wgs30x|chr21|small] [--scaling weak|strong]
  python bench.py --impl reference ...     the reference's own CPU code (oracle/_ref) on the host cores

One "step" = one pass of the hot path over one batch: prep + CIGAR walk (signatures, depth events) +
tile event ranges + depth tiles + signature sort + DBSCAN1D over the signature starts.
  value  : device-resident inputs, CUDA-event time on the library's own stream, max over ranks
  e2e    : same work through the host-facing C-ABI calls with HOST buffers: H2D of the packed SoA and
           D2H of depth maps, signatures and labels inside the timed region (wall clock, max over ranks)
  roofline: dominant kernel (depth tiles), algorithmic bytes / event time vs MEASURED_PEAKS.json
Multi-GPU: one process per GPU (torchrun), no data-path collective.  weak = every rank scans its own
whole batch (a cohort of samples); strong = one genome region-sharded with halo reads (config 4).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from contextsv_b200 import shard, synth  # noqa: E402

METRIC = "aligned reads/sec (CIGAR scan+depth+DBSCAN1D)"
DB_EPS, DB_MIN_PTS = 100.0, 5        # the reference's own DBSCAN1D parameters (sv_caller.cpp:270)


SYNTH_KW = {
    # BASELINE configs[2] in small: one chromosome of 60x ONT ultra-long reads (N50 50 kb, ~10 % indel rate => ~0.2 CIGAR
    # ops per aligned base; the whole genome at this depth is ~3.7e10 ops and is scanned as a stream of such shards)
    "ont60x_chr20": dict(profile=1, coverage=60.0, read_len_mean=50000.0, indel_rate=0.10, indel_len_max=4),
    # BASELINE configs[4] in small: SV-rich set with per-read breakpoint jitter (100k SVs over the genome ~ 8.5k on chr1)
    "svrich_chr1": dict(sv_jitter_sd=10.0),
}


def workload_contigs(name):
    if name == "ont60x_chr20":
        return [shard.GRCH38[19][1]], 500
    if name == "svrich_chr1":
        return [shard.GRCH38[0][1]], 8500
    if name == "wgs30x":
        return [l for _, l in shard.GRCH38], 25000
    if name == "chr1_5":
        return [l for _, l in shard.GRCH38[:5]], 8500
    if name == "chr21":
        return [46709983], 400
    if name == "small":
        return [5_000_000, 3_000_000], 200
    raise SystemExit("unknown workload " + name)


def workload_name(name):
    return {"ont60x_chr20": "synthetic 60x ONT ultra-long chr20 (N50 50 kb, dense CIGAR) [BASELINE configs[2], one shard of the stream]",
            "svrich_chr1": "synthetic 30x HiFi chr1, SV-rich (8.5k SVs, breakpoint jitter) [BASELINE configs[4], one chromosome]",
            "wgs30x": "synthetic whole-genome 30x HiFi GRCh38-shaped (24 contigs, 15 kb reads, 25k SVs) [BASELINE configs[1]]",
            "chr1_5": "synthetic 30x HiFi chr1-chr5 (1.06 Gb; profiling workload)",
            "chr21": "synthetic 30x HiFi chr21 [BASELINE configs[0]]",
            "small": "synthetic 30x HiFi, 2 contigs of 5+3 Mb (debug)"}[name]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.11]
        if not inside and self.rows:          # timed region shorter than the sampling period: nearest sample
            inside = [min(self.rows, key=lambda x: abs(x[0] - t1))[1]]
        for row in inside:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------- reference arm

def reference_sample(cores, seed):
    """Bounded sample of the same workload: `cores` chr21-sized contigs, one per host thread, the way the
    reference parallelises (one chromosome per ThreadPool worker, sv_caller.cpp:828-851)."""
    L = 46709983 // 4          # quarter-chr21 contigs keep one step at a few seconds of CPU work per core
    clen = [L] * cores
    r = synth.generate(clen, seed=seed, n_sv=max(1, int(25000 * L * cores / 3.1e9)))
    return r, clen


def run_reference_steps(r, clen, steps, warmup):
    from oracle.oracle_py import Oracle, Reference, ref_available
    import concurrent.futures as cf
    kind = "reference" if ref_available() else "port"
    eng = Reference() if kind == "reference" else Oracle()
    cores = len(clen)
    subs = []
    ends = shard.ref_end(r)
    for tid in range(cores):
        sub, _ = shard.select_reads(r, [(tid, 0, clen[tid] + 1, clen[tid] + 1)], ends)
        sub = dict(sub); sub["tid"] = None          # one contig per worker, like one region string per task
        subs.append(sub)

    def one(tid):
        sub = subs[tid]
        if kind == "reference":
            d, s, nz, mean = eng.depth(sub, 0, [clen[tid]])
            st, en, ty, ev, alts = eng.cigar_scan(sub, 0, [clen[tid]])
            for t in (0, 3):     # DEL / INS, the grouping of mergeSVs (sv_object.cpp:61-83)
                eng.dbscan1d(st[ty == t].astype(np.int32), DB_EPS, DB_MIN_PTS)
        else:
            d, s, nz = eng.depth(sub, 0, clen[tid] + 1)
            sg = eng.cigar_scan(sub, 0, clen[tid] + 1, fast=False)
            for isdel in (True, False):
                eng.dbscan1d(sg["start"][(sg["kind"] == 1) == isdel].astype(np.int32), DB_EPS, DB_MIN_PTS)
        return len(d)

    times = []
    with cf.ThreadPoolExecutor(max_workers=cores) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            list(ex.map(one, range(cores)))
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return kind, sum(times)


def main_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    r, clen = reference_sample(cores, args.seed)
    kind, total = run_reference_steps(r, clen, args.steps, args.warmup)
    n_reads = int(r["n_reads"])
    value = n_reads * args.steps / total
    sample = "%d contigs of %d bp (quarter chr21) at 30x = %d reads/step, one contig per thread; %s build -O2" % (
        len(clen), clen[0], n_reads, "oracle/_ref (unmodified reference sources + htslib shim)" if kind == "reference" else "oracle port")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": {"workload": workload_name(args.workload), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ our arm

def main_ours(args):
    import torch
    import torch.distributed as dist
    from contextsv_b200 import _capi, api

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = api.Context(local_rank)
    contig_len, n_sv = workload_contigs(args.workload)
    t_gen = time.perf_counter()
    if args.scaling == "weak" or world == 1:
        # every rank scans its own sample of the workload (different seed per rank)
        reads = synth.generate(contig_len, alloc=_capi.pinned_empty, seed=args.seed + rank, n_sv=n_sv, **SYNTH_KW.get(args.workload, {}))
        regions = api.whole_contig_regions(contig_len)
    else:
        # config 4: one genome, region-sharded by cumulative length, halo reads included
        full = synth.generate(contig_len, seed=args.seed, n_sv=n_sv, **SYNTH_KW.get(args.workload, {}))
        regions = shard.plan_regions(contig_len, world)[rank]
        sub, _ = shard.select_reads(full, regions)
        reads = {}
        for k, v in sub.items():
            if isinstance(v, np.ndarray):
                p = _capi.pinned_empty(len(v), v.dtype); p[:] = v; reads[k] = p
            else:
                reads[k] = v
        del full, sub
    t_gen = time.perf_counter() - t_gen
    n_reads, n_ops = int(reads["n_reads"]), int(reads["n_ops"])
    depth_words = sum(e - b for (_, b, e, _) in regions)

    # ---- value: inputs resident in HBM, device time on the library's stream
    batch = api.Batch(ctx, reads, regions)

    def step_resident():
        if args.depth_only:          # diagnostic: the depth stages without the signature side stream beside them
            batch.scan(want_depth=True, want_sigs=False)
            return
        batch.scan(want_depth=True, want_sigs=True)
        batch.sigs_dbscan1d(DB_EPS, DB_MIN_PTS, fetch=False)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML init) can stall kernel launches for
    # hundreds of milliseconds; only samples taken inside the timed region are reported
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(1.0)
    for _ in range(args.warmup):
        step_resident()
    ctx.sync()
    n_sig = 0 if args.depth_only else batch.sigs_count()
    ctx.profile_read(reset=True)
    ctx.profile_enable(True)
    barrier()
    sampler.mark_begin()
    launches0 = ctx.launches
    ctx.timer_begin()
    for _ in range(args.steps):
        step_resident()
    ms = ctx.timer_end()
    sampler.mark_end()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    ctx.profile_enable(False)
    stages = ctx.profile_read(reset=True)
    ms_max = max_over_ranks(ms)
    total_reads = sum_over_ranks(n_reads)
    value = total_reads * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel and of the whole path (algorithmic bytes)
    peak, peak_src = measured_peak_gbs()
    b_alg = 15 * n_reads + 4 * n_ops + 4 * depth_words + 21 * n_sig + 8 * n_sig
    tile_ms = stages["k_depth_tiles16"][0] / args.steps  # the dominant kernel alone: CUDA events around its launch(es) on its stream
    tile_bytes = 4 * depth_words
    achieved = tile_bytes / (tile_ms * 1e-3) / 1e9 if tile_ms > 0 else 0.0
    path_gbs = b_alg / (ms / args.steps * 1e-3) / 1e9
    traffic = None          # dram bytes read + written by one launch of the dominant kernel, from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == args.workload and world == 1:
            traffic = tj["traffic_bytes_per_launch"]
    roofline = {
        "bound": "hbm", "kernel": "k_depth_tiles16", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": tile_bytes, "ms_per_launch": tile_ms,
        "path": {"algorithmic_bytes_per_step": b_alg, "achieved": path_gbs, "frac": path_gbs / peak},
        "stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()},
    }

    # ---- e2e: the host-facing calls with host buffers; H2D and D2H inside the timed region
    batch.free()
    out_depth = []
    for (_, b, e, _) in regions:
        try:
            out_depth.append(_capi.pinned_empty(e - b, np.uint32))
        except _capi.CsvError:
            out_depth.append(np.empty(e - b, np.uint32))

    # The depth maps come back through csv_depth_fetch_all: bytes over PCIe + host threads that widen them into the
    # caller's uint32 arrays (fetch.cu).  The host cores are shared by the ranks of one box.
    fetch_threads = max(1, min(16, (os.cpu_count() or 1) // max(world, 1) - (2 if world == 1 else 0)))
    ctx.set_fetch(threads=fetch_threads)

    def step_e2e():
        """One batch: upload everything, scan, fetch everything."""
        bt = api.Batch(ctx, reads, regions)                       # H2D of the packed SoA
        bt.scan(want_depth=True, want_sigs=True)
        lab = bt.sigs_dbscan1d(DB_EPS, DB_MIN_PTS)                # D2H labels
        sums, nzs = bt.depth_stats()
        bt.depth_all(out_depth)                                   # D2H depth maps (uint32 per base in host memory)
        sg = bt.sigs()                                            # D2H signatures
        bt.free()
        return len(lab), len(sg["start"]), int(sums.sum()), int(nzs.sum())

    def timed(step, n_steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            r = step()
        ctx.sync()
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt), r

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.skip_e2e:        # profiling runs only (ncu): never used for a reported number
        e2e_steps, dt_max, e2e_value = 0, 0.0, None
    else:
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        st0 = ctx.fetch_stats()
        dt_max, _ = timed(step_e2e, e2e_steps)
        st1 = ctx.fetch_stats()
        narrow_chunks, fallback_chunks = (st1[0] - st0[0]) // e2e_steps, (st1[1] - st0[1]) // e2e_steps     # per step
        e2e_value = total_reads * e2e_steps / dt_max
        # the same step with the plain 32-bit DMA of the map (csv_ctx_set_fetch threads = 0), for comparison
        ctx.set_fetch(threads=0)
        step_e2e()
        dt_plain, _ = timed(step_e2e, e2e_steps)
        ctx.set_fetch(threads=fetch_threads)
        # double-buffered across steps: the SoA of step s + 1 is uploaded (second context, helper thread) while the
        # maps of step s come back; every step still uploads its own inputs and returns its own results
        import concurrent.futures as cf
        ctx2 = api.Context(local_rank)
        ctx2.set_fetch(threads=fetch_threads)
        ex = cf.ThreadPoolExecutor(1)

        def run_double_buffered(n_steps):
            lanes = (ctx, ctx2)
            fut = ex.submit(api.Batch, lanes[0], reads, regions)
            r = None
            for s_ in range(n_steps):
                bt = fut.result()
                if s_ + 1 < n_steps:
                    fut = ex.submit(api.Batch, lanes[(s_ + 1) & 1], reads, regions)
                bt.scan(want_depth=True, want_sigs=True)
                lab = bt.sigs_dbscan1d(DB_EPS, DB_MIN_PTS)
                sums, nzs = bt.depth_stats()
                bt.depth_all(out_depth)
                sg = bt.sigs()
                bt.free()
                r = (len(lab), len(sg["start"]), int(sums.sum()), int(nzs.sum()))
            return r
        run_double_buffered(2)
        db_steps = 2 * e2e_steps
        barrier()
        t0 = time.perf_counter()
        r_db = run_double_buffered(db_steps)
        dt_db = max_over_ranks(time.perf_counter() - t0)
        barrier()
        ex.shutdown()
        ctx2.close()
    # ---- the same step when the consumers of the depth map query the device (csv_depth_at = getReadDepth for every
    # signature start, csv_window_sums for the log2 windows) instead of the 12 GB map crossing PCIe.  Extra information:
    # the contract's e2e above returns the whole map in host memory, as the reference's interface does.
    def step_e2e_device_consumers():
        bt = api.Batch(ctx, reads, regions)
        bt.scan(want_depth=True, want_sigs=True)
        lab = bt.sigs_dbscan1d(DB_EPS, DB_MIN_PTS)
        sums, nzs = bt.depth_stats()
        sg = bt.sigs()
        nq = 0
        for i in range(len(regions)):
            lo, hi = int(sg["region_off"][i]), int(sg["region_off"][i + 1])
            if hi > lo and regions[i][1] == 0 and regions[i][2] == regions[i][3]:
                bt.depth_at(i, sg["start"][lo:hi]); nq += hi - lo
        bt.free()
        return nq

    edc = None
    if not args.skip_e2e:
        step_e2e_device_consumers()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nq = step_e2e_device_consumers()
        ctx.sync()
        dt2 = max_over_ranks(time.perf_counter() - t0)
        barrier()
        edc = {"value": total_reads * e2e_steps / dt2, "unit": "reads/s", "ms_per_step": 1e3 * dt2 / e2e_steps,
               "h2d_bytes_per_step": 15 * n_reads + 8 + 4 * n_ops + 4 * nq, "d2h_bytes_per_step": 25 * n_sig + 12 * len(regions) + 4 * nq,
               "note": "depth map stays in HBM; getReadDepth for every signature start served by csv_depth_at"}
    h2d = 15 * n_reads + 8 + 4 * n_ops
    d2h_results = 21 * n_sig + 4 * n_sig + 12 * len(regions)
    d2h_plain = 4 * depth_words + d2h_results
    e2e_extra = {}
    if not args.skip_e2e:
        # bytes that really cross PCIe on the narrow path: one byte per base + header and exception list per chunk,
        # plus the chunks whose list overflowed, again as 32-bit words
        chunk, exc = 2 << 20, 2048
        n_chunks = sum((e - b + chunk - 1) // chunk for (_, b, e, _) in regions)
        d2h = depth_words + n_chunks * (16 + 8 * exc) + fallback_chunks * 4 * chunk + d2h_results
        e2e_extra = {"result_bytes_in_host_memory": d2h_plain, "depth_fetch": "narrow: u8 over PCIe + %d host threads widen to uint32 (fetch.cu)" % fetch_threads,
                     "fetch_chunks_narrow_per_step": narrow_chunks, "fetch_chunks_refetched_plain_per_step": fallback_chunks,
                     "double_buffered": {"value": total_reads * db_steps / dt_db, "ms_per_step": 1e3 * dt_db / db_steps, "steps": db_steps,
                                         "note": "upload of step s+1 (second context, helper thread) beside the fetch of step s; pipeline fill included"},
                     "plain_dma": {"value": total_reads * e2e_steps / dt_plain, "ms_per_step": 1e3 * dt_plain / e2e_steps, "d2h_bytes_per_step": d2h_plain}}
    else:
        d2h = d2h_plain

    # ---- CPU baseline (rank 0, N == 1): the reference's own code on a bounded sample of the workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r2, clen2 = reference_sample(cores, args.seed)
        kind, total = run_reference_steps(r2, clen2, 1, 0)
        cpu = {"value": int(r2["n_reads"]) / total, "unit": "reads/s", "cores": cores, "kind": kind,
               "sample": "%d contigs of %d bp at 30x (%d reads), one contig per host thread, depth + CIGAR signatures + DBSCAN1D, -O2"
                         % (len(clen2), clen2[0], int(r2["n_reads"]))}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "reads_per_rank": n_reads, "cigar_ops_per_rank": n_ops,
                       "depth_positions_per_rank": depth_words, "signatures_per_rank": n_sig, "regions_per_rank": len(regions),
                       "dbscan1d": {"eps": DB_EPS, "min_pts": DB_MIN_PTS, "groups": "per (region, SVType) over signature starts"},
                       "l2": "inputs_larger_than_l2 (CIGAR %.2f GB read + depth %.2f GB written per step)" % (4 * n_ops / 1e9, 4 * depth_words / 1e9),
                       "parallelism": "1 process/GPU, %s, no collective" % ("independent samples" if args.scaling == "weak" or world == 1 else "region shards + halo reads"),
                       "host_generation_s": round(t_gen, 2)},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": 1e3 * dt_max / max(e2e_steps, 1), **e2e_extra},
            "e2e_device_consumers": edc,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wgs30x", choices=["wgs30x", "chr1_5", "chr21", "small", "ont60x_chr20", "svrich_chr1"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--seed", type=int, default=20261018 + 2)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--depth-only", action="store_true", help="diagnostic: resident step without signatures / DBSCAN1D (never a reported number)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3        # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())

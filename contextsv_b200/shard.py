"""Host-side region planner and result merging for multi-GPU runs.

The path shards by chromosome/region with no exchange step: every rank scans the reads that overlap
its regions (its own reads plus the halo reads that start before a region and reach into it), depth
slices are disjoint and simply concatenated, per-region (sum, nonzero) add up, and the signatures of
each read come from the one region that owns it.  The reference only shards by whole chromosome
(sv_caller.cpp:847-851).
"""
import numpy as np

GRCH38 = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
    ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
    ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
    ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415),
]


OP_COST = 4.7     # cost of one CIGAR op in units of one depth position (B200: walk 3.5 ps/op, depth tiles 0.73 ps/position)


def plan_regions(contig_len, n_shards, reads=None, op_cost=OP_COST):
    """Cuts the concatenated depth-map index space (sum of L+1) into n_shards contiguous pieces of (almost) equal COST
    = positions + op_cost * CIGAR ops of the records that start there (sum(bases + c * ops)); without
    `reads` the pieces have equal length.  Uniform HiFi coverage makes the two the same; ONT-dense or SV-rich stretches
    move the cuts.  Cuts may fall anywhere inside a contig.  Returns a list (one entry per shard) of lists of
    (tid, beg, end, map_size)."""
    sizes = [int(l) + 1 for l in contig_len]
    total = sum(sizes)
    base_of = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    if reads is None or int(reads["n_reads"]) == 0 or n_shards == 1:
        bounds = [(total * k) // n_shards for k in range(n_shards + 1)]
    else:
        n = int(reads["n_reads"])
        tid = np.zeros(n, np.int64) if reads.get("tid") is None else np.asarray(reads["tid"]).astype(np.int64)
        ok = (tid >= 0) & (tid < len(sizes))
        t_ok = np.where(ok, tid, 0)
        # concatenated depth index of every record start (clipped into its contig), non-decreasing for sorted records
        x = base_of[t_ok] + np.clip(np.asarray(reads["pos0"]).astype(np.int64) + 1, 0, np.asarray(sizes, np.int64)[t_ok] - 1)
        x = np.maximum.accumulate(np.where(ok, x, 0))
        ops_before = np.asarray(reads["cig_off"]).astype(np.int64)[:n]          # ops of the records before record j
        f = x.astype(np.float64) + op_cost * ops_before                            # cost of everything left of record j's start
        f_total = float(total) + op_cost * float(int(reads["n_ops"]))
        bounds = [0]
        for k in range(1, n_shards):
            target = f_total * k / n_shards
            j = int(np.searchsorted(f, target, side="left"))
            if j >= n:
                cut = int(min(total, max(bounds[-1], total - (f_total - target))))   # past the last record: positions only
            else:
                # between record j-1 and record j only positions add cost: step back from record j's start
                cut = int(x[j] - min(f[j] - target, float(x[j] - (x[j - 1] if j else 0))))
            bounds.append(min(max(cut, bounds[-1]), total))
        bounds.append(total)
    shards = [[] for _ in range(n_shards)]
    base = 0
    for tid_, size in enumerate(sizes):
        for k in range(n_shards):
            lo, hi = max(bounds[k], base), min(bounds[k + 1], base + size)
            if lo < hi:
                shards[k].append((tid_, lo - base, hi - base, size))
        base += size
    return shards


def contig_holders(plan):
    """plan (plan_regions): contig id -> the shards that hold a region of it, in genome order.  A contig with more than one
    holder was cut; the FIRST holder finalises it (merges the holders' signature runs, re-fits its clusters)."""
    holders = {}
    for k, regs in enumerate(plan):
        for (tid, _, _, _) in regs:
            if k not in holders.setdefault(tid, []):
                holders[tid].append(k)
    return holders


def contig_part(sigs, regions, base, tid, extra=()):
    """The runs of one shard that belong to contig `tid`, as merge_signatures parts [(sigs restricted, [region], base), ...]."""
    parts = []
    for ri, g in enumerate(regions):
        if g[0] != tid:
            continue
        lo, hi = int(sigs["region_off"][ri]), int(sigs["region_off"][ri + 1])
        d = {k: sigs[k][lo:hi] for k in SIG_FIELDS + tuple(extra)}
        d["region_off"] = np.array([0, hi - lo], np.uint64)
        parts.append((d, [g], base))
    return parts


REF_MASK = (1 << 0) | (1 << 2) | (1 << 3) | (1 << 7) | (1 << 8)


def ref_end(reads):
    """pos0 + 1 + reference length of every record (depth-map index one past the last covered base)."""
    n = int(reads["n_reads"])
    cig = np.asarray(reads["cigar"]); off = np.asarray(reads["cig_off"]).astype(np.int64)
    rl = (cig >> 4).astype(np.int64) * ((REF_MASK >> (cig & 15)) & 1)
    csum = np.concatenate([[0], np.cumsum(rl)])
    return np.asarray(reads["pos0"]).astype(np.int64) + 1 + (csum[off[1:n + 1]] - csum[off[:n]])


def select_reads(reads, regions, ends=None):
    """Contiguous slice of the (tid, pos0)-sorted records that contains every record overlapping or owned
    by `regions` (a shard's region list, in genome order).  Views, no copy of the CIGAR words."""
    n = int(reads["n_reads"])
    if n == 0 or not regions:
        return dict(reads), 0
    tid = np.zeros(n, np.int64) if reads.get("tid") is None else np.asarray(reads["tid"]).astype(np.int64)
    idx = np.asarray(reads["pos0"]).astype(np.int64) + 1
    t0, b0, _, _ = regions[0]
    t1, _, e1, m1 = regions[-1]
    key = tid * (1 << 33) + idx
    i_own = int(np.searchsorted(key, t0 * (1 << 33) + b0, side="left"))
    # halo: records of contig t0 that start before b0 and reach past it (none when the region starts the contig)
    lo_t = int(np.searchsorted(key, t0 * (1 << 33), side="left"))
    i0 = i_own
    if lo_t < i_own:
        ends = ref_end(reads) if ends is None else ends
        halo = np.nonzero(ends[lo_t:i_own] > b0)[0]
        if len(halo):
            i0 = lo_t + int(halo[0])
    if e1 == m1:
        i1 = int(np.searchsorted(key, (t1 + 1) * (1 << 33), side="left"))     # last region also owns records beyond the contig end
    else:
        i1 = int(np.searchsorted(key, t1 * (1 << 33) + e1, side="left"))
    i1 = max(i1, i0)
    off = np.asarray(reads["cig_off"])
    o0, o1 = int(off[i0]), int(off[i1])
    sub = {
        "n_reads": i1 - i0, "n_ops": o1 - o0,
        "tid": None if reads.get("tid") is None else reads["tid"][i0:i1],
        "pos0": reads["pos0"][i0:i1], "flag": reads["flag"][i0:i1], "mapq": reads["mapq"][i0:i1],
        "cig_off": (off[i0:i1 + 1] - np.uint64(o0)).astype(np.uint64), "cigar": reads["cigar"][o0:o1],
    }
    for k in ("n_gap", "ref_len"):
        if reads.get(k) is not None:
            sub[k] = reads[k][i0:i1]
    return sub, i0


def plan_by_ops(reads, contig_len, max_ops, ends=None):
    """Shards for ONE device when the records hold more CIGAR ops than a batch takes (csv_batch_upload: ops + records
    < 2^31; BASELINE config 3, 60x ONT, is ~3.7e10 ops): the same region sharding as across GPUs, in time instead of
    space.  A shard is the slice [first halo record, last own record]; it is closed right before the first record that
    would push its ops + records past max_ops.  Cuts fall between records with different pos0, so every record is owned
    by exactly one shard.  One searchsorted per shard, no loop over records.
    Returns a list of region lists [(tid, beg, end, map_size), ...] in genome order."""
    n = int(reads["n_reads"])
    sizes = [int(l) + 1 for l in contig_len]
    if n == 0:
        return [[(t, 0, sz, sz) for t, sz in enumerate(sizes)]]
    tid = np.zeros(n, np.int64) if reads.get("tid") is None else np.asarray(reads["tid"]).astype(np.int64)
    idx = np.asarray(reads["pos0"]).astype(np.int64) + 1
    off = np.asarray(reads["cig_off"]).astype(np.int64)
    ends = ref_end(reads) if ends is None else ends
    weight = off + np.arange(n + 1)                              # ops + records before record j
    key = tid * (1 << 33) + idx
    new_group = np.concatenate([[True], key[1:] != key[:-1]])    # records sharing (tid, pos0) are never split
    group_start = np.maximum.accumulate(np.where(new_group, np.arange(n), 0))
    shards, cur = [], []
    cur_tid, cur_beg = 0, 0                   # the open region starts at (cur_tid, cur_beg)
    i_first = 0                               # first record (halo included) of the open shard
    while True:
        # first record that does not fit any more (its whole group must fit)
        j = int(np.searchsorted(weight, weight[i_first] + max_ops, side="right")) - 1      # records [i_first, j) fit
        if j >= n:
            break
        i = int(group_start[j])               # cut before the group that contains record j
        if i <= i_first or not (tid[i] > cur_tid or idx[i] > cur_beg):
            # the group at the start of the shard alone is too much, or the cut would not advance the open region
            g_end = int(np.searchsorted(key, key[j], side="right"))
            raise ValueError("max_ops=%d is too small: the records overlapping index %d of contig %d alone hold %d ops + records"
                             % (max_ops, int(idx[j]), int(tid[j]), int(weight[g_end] - weight[i_first])))
        t, cut = int(tid[i]), int(min(idx[i], sizes[int(tid[i])]))
        # close the open shard at (t, cut): whole contigs up to t, then [.., cut) of t
        while cur_tid < t:
            if cur_beg < sizes[cur_tid]:
                cur.append((cur_tid, cur_beg, sizes[cur_tid], sizes[cur_tid]))
            cur_tid += 1; cur_beg = 0
        if cut > cur_beg:
            cur.append((t, cur_beg, cut, sizes[t]))
            cur_beg = cut
        if cur:
            shards.append(cur); cur = []
        # halo of the next shard: records of contig t before i that reach past cut
        lo_t = int(np.searchsorted(tid, t, side="left"))
        h = np.nonzero(ends[lo_t:i] > cut)[0]
        i_first = lo_t + int(h[0]) if len(h) else i
    while cur_tid < len(sizes):
        if cur_beg < sizes[cur_tid]:
            cur.append((cur_tid, cur_beg, sizes[cur_tid], sizes[cur_tid]))
        cur_tid += 1; cur_beg = 0
    if cur:
        shards.append(cur)
    return shards


SIG_FIELDS = ("start", "end", "kind", "read_idx", "op_idx", "query_pos")


def merge_signatures(parts, extra=()):
    """parts: list of (sigs dict from Batch.sigs(), regions, read index offset) per shard, in genome order.
    Returns one dict per contig id in the reference's vector order (start, end, reverse insertion order).
    Every shard's run is already in that order and record indices grow with the shard, so the merge is one stable
    sort by (start, end) over the runs concatenated in REVERSE shard order: equal keys keep "later record first".
    Contigs that live in one region are passed through untouched (out[tid]["n_parts"] == 1).  `extra` names further
    per-signature arrays of the sigs dicts (labels, depths) to carry along."""
    fields = SIG_FIELDS + tuple(extra)
    per_tid = {}
    for sg, regions, base in parts:
        for r, (tid, _, _, _) in enumerate(regions):
            lo, hi = int(sg["region_off"][r]), int(sg["region_off"][r + 1])
            per_tid.setdefault(tid, []).append({k: (sg[k][lo:hi].astype(np.int64) + base if k == "read_idx" else sg[k][lo:hi]) for k in fields})
    out = {}
    for tid, runs in per_tid.items():
        runs = [x for x in runs if len(x["start"])] or runs[:1]
        if len(runs) == 1:
            m = dict(runs[0])
        else:
            in_order = all(int(a["read_idx"].max()) < int(b["read_idx"].min()) for a, b in zip(runs, runs[1:]))
            if in_order:
                m = {k: np.concatenate([x[k] for x in reversed(runs)]) for k in fields}
                key = (m["start"].astype(np.uint64) << np.uint64(32)) | m["end"].astype(np.uint64)
                order = np.argsort(key, kind="stable")
            else:                 # runs that interleave (regions handed over out of genome order): the full comparator
                m = {k: np.concatenate([x[k] for x in runs]) for k in fields}
                seq = m["read_idx"].astype(np.int64) * (1 << 31) + m["op_idx"].astype(np.int64)
                order = np.lexsort((-seq, m["end"], m["start"]))
            m = {k: v[order] for k, v in m.items()}
        m["n_parts"] = len(runs)
        out[tid] = m
    return out

"""Shared helpers of the test-suite: input builders and oracle-side conveniences."""
import numpy as np

from contextsv_b200 import synth
from oracle.oracle_py import make_reads, norm_reads

M, I, D, N, S, H, P, EQ, X, B = range(10)


def synth_reads(contig_len, **kw):
    return synth.generate(contig_len, **kw)


def random_cigar_reads(rng, n_reads, contig_len, max_ops=12, ops=(M, I, D, N, S, H, P, EQ, X), big_p=0.25, n_tids=1,
                       weird=False):
    """Adversarial records: arbitrary op alphabets, many ops >= 50, empty CIGARs, zero-length ops,
    reads running off the contig end, all flag/MAPQ combinations.  Sorted by (tid, pos0)."""
    tid = np.sort(rng.integers(0, n_tids, n_reads)).astype(np.int32)
    pos0 = np.zeros(n_reads, np.int64)
    for t in range(n_tids):
        m = tid == t
        pos0[m] = np.sort(rng.integers(0, contig_len[t] + (60 if weird else 0), int(m.sum())))
    if weird and n_reads > 3:
        first = np.nonzero(tid == 0)[0]
        if len(first):
            pos0[first[0]] = -1          # (uint32)(-1) + 1 == 0: depth index 0, sv pos wraps (cnv_caller.cpp:499)
    cigars = []
    for i in range(n_reads):
        k = int(rng.integers(0, max_ops + 1))
        c = []
        for _ in range(k):
            op = int(rng.choice(ops))
            u = rng.random()
            if u < big_p:
                ln = int(rng.choice([49, 50, 51, 60, 100, 500]))
            elif u < big_p + 0.05:
                ln = 0
            else:
                ln = int(rng.integers(1, 40))
            c.append((ln, op))
        cigars.append(c)
    flag = rng.choice([0, 16, 4, 0x100, 0x200, 0x400, 0x800, 0x810, 0], n_reads, p=[.4, .3, .03, .03, .03, .03, .06, .06, .06]).astype(np.uint16)
    mapq = rng.choice([0, 19, 20, 21, 60], n_reads).astype(np.uint8)
    r = make_reads(pos0.astype(np.int32), cigars, tid=tid if n_tids > 1 else None, flag=flag, mapq=mapq)
    return r


def random_seq4(rng, reads):
    """4-bit packed bases per record (BAM encoding); includes IUPAC codes to exercise the ->N mapping."""
    r = norm_reads(reads)
    n = int(r["n_reads"])
    qlen = np.zeros(n, np.int64)
    cig = r["cigar"]; off = r["cig_off"]
    qmask = (1 << 0) | (1 << 1) | (1 << 4) | (1 << 7) | (1 << 8)
    for i in range(n):
        c = cig[int(off[i]):int(off[i + 1])]
        qlen[i] = int(((c >> 4) * ((qmask >> (c & 15)) & 1)).sum())
    nbytes = (qlen + 1) // 2
    seq_off = np.zeros(n, np.uint64)
    seq_off[1:] = np.cumsum(nbytes)[:-1]
    total = int(nbytes.sum())
    codes = rng.choice(np.arange(16, dtype=np.uint8), size=2 * total + 2, p=[.01] + [.20, .20, .01, .20, .01, .01, .01, .20, .01, .01, .01, .01, .01, .01, .09])
    seq4 = ((codes[0::2][:total] << 4) | codes[1::2][:total]).astype(np.uint8)
    return np.ascontiguousarray(seq4), seq_off


def oracle_alt(seq4, seq_off, sig):
    """ALT allele the reference builds for a signature (sv_caller.cpp:572-591)."""
    NT16 = "=ACMGRSVTWYHKDBN"
    if sig["kind"] == 1:
        return "<DEL>"
    ln = int(sig["end"]) - int(sig["start"]) + 1
    if ln > 50:
        return "<INS>"
    out = []
    base = int(seq_off[int(sig["read_idx"])])
    for j in range(ln):
        q = int(sig["query_pos"]) + j
        b = int(seq4[base + (q >> 1)])
        ch = NT16[(b >> ((~q & 1) << 2)) & 0xF]
        out.append("N" if ch in "RYKMSWBDHV" else ch)
    return "".join(out)


# ------------------------------------------------------------------ golden vectors

import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = np.load(GOLDEN)
    return _golden


def golden_db_cases():
    g = golden()
    for i in range(int(g["n_db"][0])):
        eps, mp = g["db%d_par" % i]
        yield i, g["db%d_pts" % i], float(eps), int(mp), g["db%d_labels" % i], g["db%d_largest" % i]


def golden_cg_cases():
    g = golden()
    for i in range(int(g["n_cg"][0])):
        p = "cg%d" % i
        tid = g[p + "_tid"]
        r = norm_reads({"n_reads": len(g[p + "_pos0"]), "tid": tid if len(tid) else None, "pos0": g[p + "_pos0"], "flag": g[p + "_flag"],
                        "mapq": g[p + "_mapq"], "cig_off": g[p + "_cig_off"], "cigar": g[p + "_cigar"]})
        yield i, r, [int(x) for x in g[p + "_clen"]], g[p + "_seq4"], g[p + "_seq_off"]


def golden_cg_answer(i, tid):
    g = golden()
    q = "cg%d_t%d" % (i, tid)
    ev = g[q + "_evidence"]
    kind = np.where(ev == 1, 0, np.where(ev == 2, 1, 2)).astype(np.uint8)   # bitset<10>: bit0 CIGARINS, bit1 CIGARDEL, bit2 CIGARCLIP
    ans = {"depth": g[q + "_depth"], "sum": int(g[q + "_stats"][0]), "nonzero": int(g[q + "_stats"][1]), "mean": float(g[q + "_mean"][0]),
           "start": g[q + "_start"], "end": g[q + "_end"], "kind": kind, "svtype": g[q + "_svtype"], "alt": [str(a) for a in g[q + "_alt"]]}
    if q + "_l2" in g.files:
        ans["l2par"] = [int(x) for x in g[q + "_l2par"]]; ans["l2pos"] = g[q + "_l2pos"]; ans["l2"] = g[q + "_l2"]
    return ans


GOLDEN_DB2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_db2_v1.npz")


def golden_db2_cases():
    """2-D DBSCAN::fit cases: (start, end, eps, min_pts, labels of the compiled reference)."""
    g = np.load(GOLDEN_DB2)
    for i in range(int(g["n"][0])):
        eps, mp = g["c%d_par" % i]
        yield i, g["c%d_start" % i], g["c%d_end" % i], float(eps), int(mp), g["c%d_labels" % i]


def golden_record_summary_cases():
    """(reads, indices the reference's iterator yields, endpos, query_start, query_end of the compiled reference)."""
    g = np.load(GOLDEN_DB2)
    for i in range(int(g["n_rs"][0])):
        p = "rs%d_" % i
        r = norm_reads({"n_reads": len(g[p + "pos0"]), "tid": None, "pos0": g[p + "pos0"], "flag": g[p + "flag"], "mapq": g[p + "mapq"],
                        "cig_off": g[p + "cig_off"], "cigar": g[p + "cigar"]})
        yield i, r, g[p + "keep"], g[p + "endpos"], g[p + "qstart"], g[p + "qend"]


# ------------------------------------------------------------------ split-read events

def add_split_events(r, contig_len, rng, n_events=12, reads_per_event=(6, 14)):
    """Adds split alignments to a packed SoA: groups of reads whose primary alignment ends at one breakpoint and whose
    supplementary alignment (same query name, flag 0x800) starts at another one on the same contig -- deletion-like
    (far apart on the reference), insertion-like (an unaligned stretch of the read between the two) and inverted
    (supplementary on the other strand).  Returns (reads sorted by (tid, pos0), query name per record)."""
    n0 = int(r["n_reads"])
    off = np.asarray(r["cig_off"]).astype(np.int64)
    cig = np.asarray(r["cigar"])
    recs = [(int(r["tid"][i]) if r.get("tid") is not None else 0, int(r["pos0"][i]), int(r["flag"][i]), int(r["mapq"][i]),
             cig[off[i]:off[i + 1]], "r%d" % i) for i in range(n0)]
    for e in range(n_events):
        t = int(rng.integers(0, len(contig_len)))
        L = int(contig_len[t])
        kind = ("del", "ins", "inv")[e % 3]
        bp1 = int(rng.integers(20_000, L - 80_000))
        bp2 = bp1 + (int(rng.integers(3_000, 40_000)) if kind != "ins" else int(rng.integers(0, 30)))
        ins = int(rng.integers(2_500, 6_000)) if kind == "ins" else 0
        ls_event = int(rng.integers(3_000, 8_000)) if e % 2 == 0 else 0      # every other event: supplementary parts of one length (their ends cluster)
        for k in range(int(rng.integers(*reads_per_event))):
            lp = int(rng.integers(4_000, 9_000))
            ls = ls_event + int(rng.integers(-20, 21)) if ls_event else int(rng.integers(3_000, 8_000))
            j1, j2 = int(rng.integers(-15, 16)), int(rng.integers(-15, 16))
            name = "split%d_%d" % (e, k)
            rev = 16 if rng.random() < 0.3 else 0
            # primary: lp bases aligned up to the first breakpoint, the rest of the read hard-clipped (a trailing SOFT clip
            # would count into the reference's query_end, sv_caller.cpp:681, and hide the distance on the read)
            recs.append((t, bp1 + j1 - lp, rev, 60, np.array([(lp << 4) | 0, ((ins + ls) << 4) | 5], np.uint32), name))
            # supplementary: the clipped part, aligned from the second breakpoint on
            sflag = 0x800 | (rev ^ 16 if kind == "inv" else rev)
            recs.append((t, bp2 + j2, sflag, 60, np.array([((lp + ins) << 4) | 4, (ls << 4) | 0], np.uint32), name))
    recs.sort(key=lambda x: (x[0], x[1]))
    cigs = [x[4] for x in recs]
    cig_off = np.zeros(len(recs) + 1, np.uint64)
    cig_off[1:] = np.cumsum([len(c) for c in cigs])
    out = {"n_reads": len(recs), "n_ops": int(cig_off[-1]), "tid": np.array([x[0] for x in recs], np.int32), "pos0": np.array([x[1] for x in recs], np.int32),
           "flag": np.array([x[2] for x in recs], np.uint16), "mapq": np.array([x[3] for x in recs], np.uint8), "cig_off": cig_off,
           "cigar": np.concatenate(cigs).astype(np.uint32)}
    return out, [x[5] for x in recs]

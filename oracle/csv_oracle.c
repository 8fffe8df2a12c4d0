/*
 * csv_oracle.c -- CPU restatement of ContextSV's alignment-scan hot path.
 * TEST INFRASTRUCTURE ONLY (see csv_oracle.h).  Plain C99, no dependencies.
 */
#include "csv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* htslib constants used by the reference (sam.h; values fixed by the SAM spec) */
#define F_UNMAP 0x4
#define F_SECONDARY 0x100
#define F_QCFAIL 0x200
#define F_DUP 0x400
#define F_SUPPLEMENTARY 0x800
enum { C_MATCH = 0, C_INS = 1, C_DEL = 2, C_REF_SKIP = 3, C_SOFT_CLIP = 4,
       C_HARD_CLIP = 5, C_PAD = 6, C_EQUAL = 7, C_DIFF = 8 };

static int read_tid(const orc_reads* r, uint32_t i) { return r->tid ? r->tid[i] : 0; }

/* ------------------------------------------------------------------ depth */

/* cnv_caller.cpp:488-542 */
void orc_depth(const orc_reads* r, int32_t tid, uint32_t map_size,
               uint32_t* depth, uint64_t* sum_out, uint32_t* nonzero_out)
{
    for (uint32_t i = 0; i < r->n_reads; i++) {
        if (read_tid(r, i) != tid) continue;                       /* :466 region = contig */
        uint16_t flag = r->flag[i];
        if (flag & (F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP)) continue;   /* :491-495 */
        uint32_t ref_pos = (uint32_t)r->pos0[i] + 1u;              /* :499-500 */
        for (uint64_t k = r->cig_off[i]; k < r->cig_off[i + 1]; k++) {
            uint32_t op = r->cigar[k] & 0xf, op_len = r->cigar[k] >> 4;
            if (op == C_MATCH || op == C_EQUAL || op == C_DIFF) {  /* :507 */
                for (uint32_t j = 0; j < op_len; j++) {
                    uint32_t idx = ref_pos + j;                    /* uint32 arithmetic, :512 */
                    if ((size_t)idx >= (size_t)map_size) continue; /* :512-516 */
                    depth[idx]++;                                  /* :517 */
                }
            }
            if (op == C_MATCH || op == C_DEL || op == C_REF_SKIP || op == C_EQUAL || op == C_DIFF)
                ref_pos += op_len;                                 /* :523-524 */
            /* I,S,H,P: nothing; op >= 9 only logs (:525-529) */
        }
    }
    uint64_t cum = 0; uint32_t cnt = 0;                            /* :534-535 */
    for (uint32_t p = 0; p < map_size; p++) { cum += depth[p]; cnt += depth[p] > 0; }
    *sum_out = cum; *nonzero_out = cnt;
}

double orc_mean_cov(uint64_t sum, uint32_t nonzero)
{
    return nonzero > 0 ? (double)sum / (double)nonzero : 0.0;     /* :538 */
}

/* ------------------------------------------------------------- CIGAR scan */

static int sig_less(const orc_sig* a, const orc_sig* b)            /* sv_object.cpp:17-20 */
{
    return a->start < b->start || (a->start == b->start && a->end < b->end);
}

/* Collect the signatures of one record in op order (sv_caller.cpp:539-655).
 * Returns how many were appended to buf (capacity checked by caller: at most
 * one per op). */
static uint64_t scan_record(const orc_reads* r, uint32_t i, uint32_t min_len,
                            uint32_t depth_map_size, orc_sig* buf)
{
    uint64_t n = 0;
    uint32_t pos = (uint32_t)r->pos0[i];                           /* :542-543 */
    uint32_t query_pos = 0;                                        /* :547 */
    uint64_t beg = r->cig_off[i], end = r->cig_off[i + 1];
    for (uint64_t k = beg; k < end; k++) {
        int op_len = (int)(r->cigar[k] >> 4);                      /* :564 */
        int op = (int)(r->cigar[k] & 0xf);                         /* :565 */
        if (op_len >= (int)min_len) {                              /* :566 */
            if (op == C_INS) {                                     /* :569-596 */
                orc_sig s; s.start = pos + 1u; s.end = s.start + (uint32_t)op_len - 1u;
                s.read_idx = i; s.op_idx = (uint32_t)(k - beg); s.query_pos = query_pos; s.kind = 0;
                buf[n++] = s;
            } else if (op == C_SOFT_CLIP) {                        /* :599-631 */
                if ((size_t)(uint32_t)(pos + 1u) >= (size_t)depth_map_size)
                    continue;                                      /* :602-604 skips the advances */
                orc_sig s; s.start = pos + 1u; s.end = s.start + (uint32_t)op_len - 1u;
                s.read_idx = i; s.op_idx = (uint32_t)(k - beg); s.query_pos = query_pos; s.kind = 2;
                buf[n++] = s;
            } else if (op == C_DEL) {                              /* :634-643 */
                orc_sig s; s.start = pos + 1u; s.end = s.start + (uint32_t)op_len - 1u;
                s.read_idx = i; s.op_idx = (uint32_t)(k - beg); s.query_pos = query_pos; s.kind = 1;
                buf[n++] = s;
            }
        }
        if (op == C_MATCH || op == C_DEL || op == C_REF_SKIP || op == C_EQUAL || op == C_DIFF)
            pos += (uint32_t)op_len;                               /* :648-650 */
        if (op == C_MATCH || op == C_INS || op == C_SOFT_CLIP || op == C_EQUAL || op == C_DIFF)
            query_pos += (uint32_t)op_len;                         /* :653-655 */
    }
    return n;
}

static int sig_filter(const orc_reads* r, uint32_t i, uint8_t min_mapq)   /* :526 */
{
    uint16_t f = r->flag[i];
    return (f & F_SECONDARY) || (f & F_UNMAP) || (f & F_DUP) || (f & F_QCFAIL) ||
           ((int)r->mapq[i] < (int)min_mapq) || (f & F_SUPPLEMENTARY);
}

typedef struct { orc_sig* v; uint64_t n, cap; } sigvec;
static void sv_reserve(sigvec* s, uint64_t want)
{
    if (want <= s->cap) return;
    uint64_t c = s->cap ? s->cap : 1024;
    while (c < want) c *= 2;
    s->v = (orc_sig*)realloc(s->v, c * sizeof(orc_sig)); s->cap = c;
}

static uint64_t max_ops(const orc_reads* r)
{
    uint64_t m = 0;
    for (uint32_t i = 0; i < r->n_reads; i++) {
        uint64_t c = r->cig_off[i + 1] - r->cig_off[i];
        if (c > m) m = c;
    }
    return m;
}

uint64_t orc_cigar_scan(const orc_reads* r, int32_t tid, uint32_t min_len,
                        uint8_t min_mapq, uint32_t depth_map_size,
                        orc_sig* out, uint64_t cap)
{
    sigvec calls = {0, 0, 0};
    orc_sig* rec = (orc_sig*)malloc((max_ops(r) + 1) * sizeof(orc_sig));
    for (uint32_t i = 0; i < r->n_reads; i++) {
        if (read_tid(r, i) != tid) continue;
        if (sig_filter(r, i, min_mapq)) continue;
        uint64_t m = scan_record(r, i, min_len, depth_map_size, rec);
        for (uint64_t j = 0; j < m; j++) {                         /* :658-660 -> addSVCall */
            if (rec[j].start > rec[j].end) continue;               /* sv_object.cpp:25-28 */
            /* std::lower_bound on operator< (sv_object.cpp:31) */
            uint64_t lo = 0, hi = calls.n;
            while (lo < hi) {
                uint64_t mid = lo + (hi - lo) / 2;
                if (sig_less(&calls.v[mid], &rec[j])) lo = mid + 1; else hi = mid;
            }
            sv_reserve(&calls, calls.n + 1);
            memmove(&calls.v[lo + 1], &calls.v[lo], (calls.n - lo) * sizeof(orc_sig));  /* :32 insert */
            calls.v[lo] = rec[j]; calls.n++;
        }
    }
    uint64_t n = calls.n;
    if (out) memcpy(out, calls.v, (n < cap ? n : cap) * sizeof(orc_sig));
    free(calls.v); free(rec);
    return n;
}

static void merge_sort_stable(orc_sig* a, orc_sig* tmp, uint64_t n)
{
    if (n < 2) return;
    uint64_t h = n / 2;
    merge_sort_stable(a, tmp, h); merge_sort_stable(a + h, tmp, n - h);
    uint64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = sig_less(&a[j], &a[i]) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(orc_sig));
}

uint64_t orc_cigar_scan_fast(const orc_reads* r, int32_t tid, uint32_t min_len,
                             uint8_t min_mapq, uint32_t depth_map_size,
                             orc_sig* out, uint64_t cap)
{
    sigvec calls = {0, 0, 0};
    orc_sig* rec = (orc_sig*)malloc((max_ops(r) + 1) * sizeof(orc_sig));
    for (uint32_t i = 0; i < r->n_reads; i++) {
        if (read_tid(r, i) != tid) continue;
        if (sig_filter(r, i, min_mapq)) continue;
        uint64_t m = scan_record(r, i, min_len, depth_map_size, rec);
        for (uint64_t j = 0; j < m; j++) {
            if (rec[j].start > rec[j].end) continue;
            sv_reserve(&calls, calls.n + 1);
            calls.v[calls.n++] = rec[j];
        }
    }
    uint64_t n = calls.n;
    /* lower_bound insertion puts a new element BEFORE equal keys: equal keys
     * end up in reverse insertion order == stable sort of the reversed list. */
    for (uint64_t i = 0; i < n / 2; i++) { orc_sig t = calls.v[i]; calls.v[i] = calls.v[n - 1 - i]; calls.v[n - 1 - i] = t; }
    orc_sig* tmp = (orc_sig*)malloc((n + 1) * sizeof(orc_sig));
    merge_sort_stable(calls.v, tmp, n);
    if (out) memcpy(out, calls.v, (n < cap ? n : cap) * sizeof(orc_sig));
    free(tmp); free(calls.v); free(rec);
    return n;
}

/* ---------------------------------------------------------------- DBSCAN1D */

typedef struct { size_t* v; size_t n, cap; } idxvec;
static void iv_push(idxvec* s, size_t x)
{
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->v = (size_t*)realloc(s->v, s->cap * sizeof(size_t)); }
    s->v[s->n++] = x;
}

/* dbscan1d.cpp:58-70 */
static void region_query(const int32_t* pts, uint64_t n, size_t idx, double eps, idxvec* out)
{
    out->n = 0;
    for (size_t i = 0; i < n; i++) {
        /* std::abs(int - int) -> double; overflow of the int subtraction is UB
         * in the reference, here it is evaluated in 64-bit. */
        int64_t d = (int64_t)pts[idx] - (int64_t)pts[i];
        if (d < 0) d = -d;
        if ((double)d <= eps) iv_push(out, i);
    }
}

void orc_dbscan1d(const int32_t* pts, uint64_t n, double eps, int min_pts, int32_t* labels)
{
    int cluster_id = 0;
    for (uint64_t i = 0; i < n; i++) labels[i] = -1;               /* :10 */
    idxvec seeds = {0, 0, 0}, result = {0, 0, 0};
    for (size_t i = 0; i < n; i++) {                               /* :12 */
        if (labels[i] != -1) continue;
        /* expandCluster :25-56 */
        region_query(pts, n, i, eps, &seeds);
        if ((int)seeds.n < min_pts) { labels[i] = -2; continue; }  /* :27-30 */
        for (size_t s = 0; s < seeds.n; s++) labels[seeds.v[s]] = cluster_id;   /* :32-34 */
        size_t w = 0;                                              /* :36 erase(remove(i)) */
        for (size_t s = 0; s < seeds.n; s++) if (seeds.v[s] != i) seeds.v[w++] = seeds.v[s];
        seeds.n = w;
        while (seeds.n) {                                          /* :38 */
            size_t cur = seeds.v[--seeds.n];                       /* :39-40 */
            region_query(pts, n, cur, eps, &result);
            if ((int)result.n >= min_pts) {                        /* :43 */
                for (size_t q = 0; q < result.n; q++) {
                    size_t p = result.v[q];
                    if (labels[p] == -1 || labels[p] == -2) {      /* :45 */
                        if (labels[p] == -1) iv_push(&seeds, p);   /* :46-48 */
                        labels[p] = cluster_id;                    /* :49 */
                    }
                }
            }
        }
        ++cluster_id;                                              /* :15 */
    }
    free(seeds.v); free(result.v);
}

/* ---- per-record summaries of the split-read pass ("next" row 8f-3): SVCaller::getAlignmentReadPositions
 * (src/sv_caller.cpp:663-690) and htslib's bam_endpos (pos + max(1, reference length), 0 reference bases if unmapped) */
void orc_record_summary(const orc_reads* r, int32_t* endpos, int32_t* qstart, int32_t* qend)
{
    for (uint64_t i = 0; i < r->n_reads; i++) {
        int query_start = -1, query_end = 0;                       /* :665-666 */
        uint32_t rlen = 0;
        for (uint64_t o = r->cig_off[i]; o < r->cig_off[i + 1]; o++) {
            int op_len = (int)(r->cigar[o] >> 4), op = (int)(r->cigar[o] & 15u);
            if (query_start == -1 && (op == 0 || op == 1 || op == 7 || op == 8)) query_start = query_end;   /* :674-676 */
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) query_end += op_len;                    /* :680-682 */
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += (uint32_t)op_len;
        }
        if (query_start == -1) query_start = 0;                    /* :685-687 */
        if (r->flag[i] & 0x4) rlen = 0;
        endpos[i] = (int32_t)((uint32_t)r->pos0[i] + (rlen ? rlen : 1u));
        qstart[i] = query_start; qend[i] = query_end;
    }
}

/* ---- DBSCAN::fit, 2-D (src/dbscan.cpp:9-81): the clustering mergeSVs runs on the signatures ("next" row 8f-1) */
static double db2_distance(uint32_t s1, uint32_t e1, uint32_t s2, uint32_t e2)   /* :69-81 */
{
    int me = (int)e1 < (int)e2 ? (int)e1 : (int)e2, ms = (int)s1 > (int)s2 ? (int)s1 : (int)s2;
    int overlap = me - ms > 0 ? me - ms : 0;
    int length1 = (int)(e1 - s1), length2 = (int)(e2 - s2);
    double a = (double)overlap / (double)length1, b = (double)overlap / (double)length2;
    double m = (b < a) ? b : a;                                    /* std::min(a, b) */
    return 1.0 - m;
}
static void region_query2(const uint32_t* st, const uint32_t* en, uint64_t n, size_t q, double eps, idxvec* out)   /* :59-67 */
{
    out->n = 0;
    for (size_t i = 0; i < n; i++) if (db2_distance(st[q], en[q], st[i], en[i]) <= eps) iv_push(out, i);
}
void orc_dbscan2d(const uint32_t* start, const uint32_t* end, uint64_t n, double eps, int min_pts, int32_t* labels)
{
    int cluster_id = 0;
    for (uint64_t i = 0; i < n; i++) labels[i] = -1;               /* :11 */
    idxvec seeds = {0, 0, 0}, result = {0, 0, 0};
    for (size_t i = 0; i < n; i++) {                               /* :13 */
        if (labels[i] != -1) continue;
        region_query2(start, end, n, i, eps, &seeds);              /* expandCluster :27-57 */
        if ((int)seeds.n < min_pts) { labels[i] = -2; continue; }
        for (size_t s = 0; s < seeds.n; s++) labels[seeds.v[s]] = cluster_id;   /* :33-35 */
        size_t w = 0;                                              /* :37 erase(remove(i)) */
        for (size_t s = 0; s < seeds.n; s++) if (seeds.v[s] != i) seeds.v[w++] = seeds.v[s];
        seeds.n = w;
        while (seeds.n) {
            size_t cur = seeds.v[--seeds.n];                       /* :40-41 back(), pop_back() */
            region_query2(start, end, n, cur, eps, &result);
            if ((int)result.n >= min_pts) {
                for (size_t q = 0; q < result.n; q++) {
                    size_t p = result.v[q];
                    if (labels[p] == -1 || labels[p] == -2) {
                        if (labels[p] == -1) iv_push(&seeds, p);
                        labels[p] = cluster_id;
                    }
                }
            }
        }
        ++cluster_id;
    }
    free(seeds.v); free(result.v);
}

typedef struct { int32_t v; uint64_t i; } vi_t;
static int vi_cmp(const void* a, const void* b)
{
    const vi_t* x = (const vi_t*)a; const vi_t* y = (const vi_t*)b;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    return x->i < y->i ? -1 : (x->i > y->i);
}

void orc_dbscan1d_fast(const int32_t* pts, uint64_t n, double eps, int min_pts, int32_t* labels)
{
    if (n == 0) return;
    if (!(eps >= 0.0)) {
        /* no neighbourhood contains anything, not even the point itself */
        for (uint64_t i = 0; i < n; i++) labels[i] = (0 < min_pts) ? -2 : -1;
        return;
    }
    int64_t E = eps >= 4294967296.0 ? (int64_t)4294967296LL : (int64_t)floor(eps);
    vi_t* s = (vi_t*)malloc(n * sizeof(vi_t));
    for (uint64_t i = 0; i < n; i++) { s[i].v = pts[i]; s[i].i = i; }
    qsort(s, n, sizeof(vi_t), vi_cmp);
    uint8_t* core = (uint8_t*)calloc(n, 1);
    int64_t* run = (int64_t*)malloc(n * sizeof(int64_t));          /* run id of core point (sorted order) */
    uint64_t lo = 0, hi = 0;
    for (uint64_t k = 0; k < n; k++) {
        while ((int64_t)s[k].v - (int64_t)s[lo].v > E) lo++;
        if (hi < k) hi = k;
        while (hi + 1 < n && (int64_t)s[hi + 1].v - (int64_t)s[k].v <= E) hi++;
        int64_t nbr = (int64_t)(hi - lo + 1);
        core[k] = nbr >= (int64_t)min_pts;
    }
    /* runs of core points with consecutive gaps <= E */
    int64_t n_runs = 0; int64_t prev_core = -1;
    for (uint64_t k = 0; k < n; k++) {
        run[k] = -1;
        if (!core[k]) continue;
        if (prev_core < 0 || (int64_t)s[k].v - (int64_t)s[prev_core].v > E) n_runs++;
        run[k] = n_runs - 1; prev_core = (int64_t)k;
    }
    uint64_t* run_min = (uint64_t*)malloc((n_runs + 1) * sizeof(uint64_t));
    int32_t* run_pv = (int32_t*)malloc((n_runs + 1) * sizeof(int32_t));
    int32_t* run_id = (int32_t*)malloc((n_runs + 1) * sizeof(int32_t));
    for (int64_t c = 0; c < n_runs; c++) run_min[c] = UINT64_MAX;
    for (uint64_t k = 0; k < n; k++)
        if (core[k] && s[k].i < run_min[run[k]]) { run_min[run[k]] = s[k].i; run_pv[run[k]] = s[k].v; }
    /* cluster id = rank of the run by its minimum input index */
    uint8_t* is_init = (uint8_t*)calloc(n, 1);
    for (int64_t c = 0; c < n_runs; c++) is_init[run_min[c]] = 1;
    int32_t* rank_at = (int32_t*)malloc(n * sizeof(int32_t));
    int32_t acc = 0;
    for (uint64_t i = 0; i < n; i++) { rank_at[i] = acc; acc += is_init[i]; }
    for (int64_t c = 0; c < n_runs; c++) run_id[c] = rank_at[run_min[c]];
    /* labels */
    int64_t* left = (int64_t*)malloc(n * sizeof(int64_t));
    int64_t last = -1;
    for (uint64_t k = 0; k < n; k++) { if (core[k]) last = (int64_t)k; left[k] = last; }
    int64_t nxt = -1;
    for (uint64_t kk = n; kk-- > 0;) {
        uint64_t k = kk;
        if (core[k]) { nxt = (int64_t)k; labels[s[k].i] = run_id[run[k]]; continue; }
        int64_t cand[2]; int nc = 0;
        if (left[k] >= 0 && (int64_t)s[k].v - (int64_t)s[left[k]].v <= E) cand[nc++] = run[left[k]];
        if (nxt >= 0 && (int64_t)s[nxt].v - (int64_t)s[k].v <= E) cand[nc++] = run[nxt];
        if (nc == 0) { labels[s[k].i] = -2; continue; }
        int32_t mn = INT32_MAX, steal = -1;
        for (int c = 0; c < nc; c++) {
            int32_t id = run_id[cand[c]];
            if (id < mn) mn = id;
            int64_t d = (int64_t)s[k].v - (int64_t)run_pv[cand[c]];
            if (d < 0) d = -d;
            if (d <= E && id > steal) steal = id;
        }
        labels[s[k].i] = steal > mn ? steal : mn;
    }
    free(s); free(core); free(run); free(run_min); free(run_pv); free(run_id);
    free(is_init); free(rank_at); free(left);
}

/* dbscan1d.cpp:72-90 */
uint64_t orc_largest_cluster(const int32_t* pts, const int32_t* labels, uint64_t n, int32_t* out)
{
    /* std::map iterates ids in increasing order; strictly-greater keeps the first.
     * When no id >= 0 exists, largest_cluster_id stays -1 and the reference
     * returns cluster_map[-1] (:89): the points still labelled -1, which only
     * happens for eps < 0 with minPts <= 0. */
    int32_t max_id = -1;
    for (uint64_t i = 0; i < n; i++) if (labels[i] > max_id) max_id = labels[i];
    int32_t best = -1;
    if (max_id >= 0) {
        uint64_t* cnt = (uint64_t*)calloc((size_t)max_id + 1, sizeof(uint64_t));
        for (uint64_t i = 0; i < n; i++) if (labels[i] >= 0) cnt[labels[i]]++;
        uint64_t best_n = 0;
        for (int32_t c = 0; c <= max_id; c++) if (cnt[c] > best_n) { best_n = cnt[c]; best = c; }
        free(cnt);
    }
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++) if (labels[i] == best) out[m++] = pts[i];
    return m;
}

/* ---------------------------------------------------------- log2 windows */

/* cnv_caller.cpp:76-113 */
void orc_log2_windows(const uint32_t* depth, uint64_t map_size,
                      uint32_t start_pos, uint32_t end_pos, int sample_size,
                      double mean_chr_cov, uint32_t* win_start,
                      uint32_t* win_end, uint64_t* win_sum,
                      uint32_t* win_count, double* log2_out)
{
    double pos_step = (double)(uint32_t)(end_pos - start_pos + 1u) / (double)sample_size;   /* :76 */
    for (int i = 0; i < sample_size; i++) {
        win_start[i] = (uint32_t)(start_pos + i * pos_step);       /* :80 */
        win_end[i] = (uint32_t)(start_pos + (i + 1) * pos_step);   /* :81 */
        double cov_sum = 0.0; int pos_count = 0; uint64_t isum = 0;
        for (int j = 0; j < pos_step; j++) {                       /* :86 */
            uint32_t pos = (uint32_t)(start_pos + i * pos_step + j);
            if (pos > end_pos) break;                              /* :89-92 */
            if ((uint64_t)pos < map_size) { cov_sum += depth[pos]; isum += depth[pos]; pos_count++; }
        }
        double log2_cov = 0.0;
        if (pos_count > 0) {
            if (cov_sum == 0) cov_sum = 1e-9;                      /* :102-106 */
            log2_cov = log2((cov_sum / (double)pos_count) / mean_chr_cov);   /* :107 */
        }
        win_sum[i] = isum; win_count[i] = (uint32_t)pos_count; log2_out[i] = log2_cov;
    }
}

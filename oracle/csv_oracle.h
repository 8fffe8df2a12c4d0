/*
 * csv_oracle.h -- CPU restatement of ContextSV's alignment-scan hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Parity pin: the reference ships no golden vectors for this path
 * (SURVEY.md F6).  The restatement is therefore pinned against the
 * reference's own code compiled unmodified into oracle/_ref/ (see
 * oracle/Makefile, oracle/ref_harness.cpp); tests/golden/ holds vectors
 * generated from that build by tests/golden/make_golden.py.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the ContextSV source tree).
 */
#ifndef CSV_ORACLE_H
#define CSV_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Packed alignment records, structure-of-arrays, BAM (file) order. */
typedef struct {
    uint32_t        n_reads;
    uint64_t        n_ops;
    const int32_t*  tid;      /* [n_reads] contig id; NULL = all contig 0 */
    const int32_t*  pos0;     /* [n_reads] bam1_core_t.pos (0-based)      */
    const uint16_t* flag;     /* [n_reads] BAM FLAG                       */
    const uint8_t*  mapq;     /* [n_reads] bam1_core_t.qual               */
    const uint64_t* cig_off;  /* [n_reads+1] offsets into cigar[]          */
    const uint32_t* cigar;    /* [n_ops] raw BAM words, len<<4 | op        */
} orc_reads;

typedef struct {
    uint32_t start;      /* SVCall.start (1-based)                          */
    uint32_t end;        /* SVCall.end                                      */
    uint32_t read_idx;   /* index of the record in orc_reads                */
    uint32_t op_idx;     /* index of the CIGAR op inside the record         */
    uint32_t query_pos;  /* reference's query_pos when the op was visited   */
    uint8_t  kind;       /* 0 CIGARINS, 1 CIGARDEL, 2 CIGARCLIP             */
} orc_sig;

/* cnv_caller.cpp:488-542 for one contig.  depth[map_size] must be
 * zero-initialised by the caller (sv_caller.cpp:801).  Only reads whose tid
 * equals `tid` contribute (sam_itr_querys(chr), cnv_caller.cpp:466). */
void orc_depth(const orc_reads* r, int32_t tid, uint32_t map_size,
               uint32_t* depth, uint64_t* sum_out, uint32_t* nonzero_out);

/* mean = double(sum)/double(nonzero), 0.0 when nonzero == 0
 * (cnv_caller.cpp:538). */
double orc_mean_cov(uint64_t sum, uint32_t nonzero);

/* sv_caller.cpp:526,539-661 + sv_object.cpp:17-33 for one contig, literal
 * lower_bound + insert.  Returns the number of signatures; writes at most
 * cap of them, in the reference's vector order. */
uint64_t orc_cigar_scan(const orc_reads* r, int32_t tid, uint32_t min_len,
                        uint8_t min_mapq, uint32_t depth_map_size,
                        orc_sig* out, uint64_t cap);

/* Same result as orc_cigar_scan, O(N log N): stable merge sort by
 * (start,end) of the signatures taken in reverse insertion order. */
uint64_t orc_cigar_scan_fast(const orc_reads* r, int32_t tid, uint32_t min_len,
                             uint8_t min_mapq, uint32_t depth_map_size,
                             orc_sig* out, uint64_t cap);

/* dbscan1d.cpp:8-66, literal sequential O(N^2). */
void orc_dbscan1d(const int32_t* pts, uint64_t n, double eps, int min_pts,
                  int32_t* labels);

/* Closed-form labelling, O(N log N).  Must agree
 * with orc_dbscan1d wherever |a-b| does not overflow int. */
void orc_dbscan1d_fast(const int32_t* pts, uint64_t n, double eps, int min_pts,
                       int32_t* labels);

/* sv_caller.cpp:663-690 (getAlignmentReadPositions) + htslib bam_endpos, per record (split-read pass, sv_caller.cpp:150-162). */
void orc_record_summary(const orc_reads* r, int32_t* endpos, int32_t* qstart, int32_t* qend);

/* dbscan.cpp:9-81 (2-D DBSCAN::fit over intervals, reciprocal-overlap distance), literal sequential O(N^2). */
void orc_dbscan2d(const uint32_t* start, const uint32_t* end, uint64_t n, double eps, int min_pts, int32_t* labels);

/* dbscan1d.cpp:72-90: points of the first cluster id >= 0 with the strictly
 * largest size, in input order.  Returns their count; with no cluster id >= 0 the reference returns cluster_map[-1]. */
uint64_t orc_largest_cluster(const int32_t* pts, const int32_t* labels,
                             uint64_t n, int32_t* out);

/* cnv_caller.cpp:76-113: per-window integer depth sum, position count and
 * log2 ratio.  All arrays have sample_size entries. */
void orc_log2_windows(const uint32_t* depth, uint64_t map_size,
                      uint32_t start_pos, uint32_t end_pos, int sample_size,
                      double mean_chr_cov, uint32_t* win_start,
                      uint32_t* win_end, uint64_t* win_sum,
                      uint32_t* win_count, double* log2_out);

#ifdef __cplusplus
}
#endif
#endif

/*
 * contextsv_b200.h -- C ABI of the B200-native alignment-scan hot path.
 *
 * This is the only door into the CUDA code: plain pointers and sizes, no C++
 * or torch types.  ContextSV itself has no FFI for this path;
 * each entry point therefore cites the reference C++ interface it stands
 * behind.  File:line citations are relative to the ContextSV source tree.
 *
 * Conventions: no exception ever crosses this boundary; every
 * call returns a csv_status (0 = ok) and leaves a message retrievable with
 * csv_last_error() (thread-local).  A csv_ctx owns one CUDA stream and its
 * scratch buffers on one device; it is NOT thread-safe -- create one per host
 * thread (the reference runs one processChromosome task per ThreadPool
 * worker, sv_caller.cpp:828-851).  Different contexts never share state.
 * There is no CPU fallback: without a usable GPU every compute entry point
 * fails with CSV_ERR_CUDA.
 */
#ifndef CONTEXTSV_B200_H
#define CONTEXTSV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CSV_OK = 0,
    CSV_ERR_CUDA = 1,       /* CUDA runtime error / no device                      */
    CSV_ERR_ARG = 2,        /* invalid argument                                    */
    CSV_ERR_CAPACITY = 3,   /* caller buffer too small (required size is reported) */
    CSV_ERR_LIMIT = 4,      /* batch exceeds an implementation limit               */
    CSV_ERR_STATE = 5       /* call out of order (e.g. fetch before run)           */
} csv_status;

/* ------------------------------------------------------------------ context */

typedef struct csv_ctx csv_ctx;

int  csv_ctx_create(int device, csv_ctx** out);
void csv_ctx_destroy(csv_ctx* ctx);
int  csv_ctx_sync(csv_ctx* ctx);                 /* wait for everything enqueued so far  */
const char* csv_last_error(void);                /* thread-local, never NULL             */
const char* csv_version(void);

/* Pinned host memory for SoA buffers and results (cudaHostAlloc). */
void* csv_host_alloc(size_t bytes);
void  csv_host_free(void* p);

/* Device timing on the context's own stream (CUDA events).  bench.py uses
 * these because torch.cuda.Event only sees torch's current stream. */
int csv_timer_begin(csv_ctx* ctx);
int csv_timer_end(csv_ctx* ctx, float* ms_out);  /* synchronises the end event           */
/* Batches uploaded after this call are scanned in up to n_chunks pipelined chunks of whole contigs: the CIGAR walk
 * of chunk c+2 runs beside the depth tiles of chunk c on a second stream.  Results are identical for every value.
 * Default 1 (also settable with the environment variable CSV_CHUNKS): on B200 both kernels are limited by the warps
 * a register file holds, so running them side by side buys nothing today. */
int csv_ctx_set_pipeline_chunks(csv_ctx* ctx, int n_chunks);
/* How csv_depth_fetch / csv_depth_fetch_all / csv_depth bring the map back.  The reference's container is a uint32 per
 * base (sv_caller.cpp:788,801); depths are small, so the map crosses PCIe as bytes plus a short list of the values
 * >= 255 and `threads` host threads widen it into the caller's array -- bit-identical to a plain copy, and the
 * destination need not be pinned.  threads = 0 selects the plain 32-bit DMA.  Default: host cores - 2 (at most 16), or the
 * environment variable CSV_FETCH_THREADS.  A negative argument keeps the current value.  chunk_positions (multiple of
 * 512, default 2 Mi) is the pipeline granule, exception_slots (default 2048) the list length per chunk (a chunk that
 * overflows it is fetched again as plain words), min_positions (default 256 Ki) the shortest fetch that takes this path. */
int csv_ctx_set_fetch(csv_ctx* ctx, int threads, int64_t chunk_positions, int64_t exception_slots, int64_t min_positions);
/* Chunks fetched narrow / re-fetched plain since the context was created. */
int csv_ctx_fetch_stats(const csv_ctx* ctx, uint64_t* narrow_chunks_out, uint64_t* fallback_chunks_out);
/* Host helper of the narrow fetch: dst[i] = src[i] for i < n, streaming stores (no device needed). */
void csv_host_widen_u8(const uint8_t* src, uint32_t* dst, size_t n);
/* Kernels launched by this context since creation. */
uint64_t csv_ctx_launch_count(const csv_ctx* ctx);
/* Per-stage device time of the scan pipeline (event pairs around each stage while
 * enabled).  on = 1: every stage; on = 2: only the dominant kernel (k_depth_tiles16) -- every event pair is a stream
 * operation between two kernels, a dozen of them per pass cost the pass about 1 %; on = 0: off.
 * csv_profile_read synchronises, fills up to max_stages entries
 * (stage name, accumulated ms, number of timed calls) and returns the number of
 * stages, or a negative csv_status. */
int csv_profile_enable(csv_ctx* ctx, int on);
int csv_profile_read(csv_ctx* ctx, int max_stages, const char** names_out, double* ms_out,
                     uint32_t* calls_out, int reset);

/* --------------------------------------------------------------- input SoA */

/* Packed alignment records, BAM file order (coordinate-sorted).  What the host
 * packer extracts from each bam1_t: core.tid, core.pos, core.flag, core.qual
 * and the raw CIGAR words of bam_get_cigar() (cnv_caller.cpp:491-503,
 * sv_caller.cpp:526,542-546).  Caller-owned; pinned memory makes the upload
 * asynchronous but is not required. */
typedef struct {
    uint32_t        n_reads;
    uint64_t        n_ops;     /* == cig_off[n_reads]; n_ops + n_reads must be < 2^31 per batch (larger inputs: shards) */
    const int32_t*  tid;       /* [n_reads] contig id, or NULL (all reads on contig 0) */
    const int32_t*  pos0;      /* [n_reads] 0-based leftmost position                  */
    const uint16_t* flag;      /* [n_reads] BAM FLAG                                   */
    const uint8_t*  mapq;      /* [n_reads] MAPQ                                       */
    const uint64_t* cig_off;   /* [n_reads+1] prefix offsets into cigar[]              */
    const uint32_t* cigar;     /* [n_ops] len<<4 | op                                  */
    const uint32_t* n_gap;     /* [n_reads] optional (NULL = not counted): number of D / N ops (BAM_CDEL,   */
                               /* BAM_CREF_SKIP) in each record's CIGAR.  The packer touches every op anyway; */
                               /* with the counts the device finds where each record's depth events go from a  */
                               /* scan over RECORDS instead of a second pass over every CIGAR word.  Checked   */
                               /* against the CIGAR during the scan: a wrong count fails with CSV_ERR_ARG.    */
    const uint32_t* ref_len;   /* [n_reads] optional, used together with n_gap: reference bases the record's   */
                               /* CIGAR consumes (M, D, N, =, X) -- bam_endpos - pos, which a packer has at    */
                               /* hand.  With it the record ranges of the depth tiles are computed beside the  */
                               /* CIGAR walk instead of after it.  Checked like n_gap.                         */
} csv_reads;

/* A slice [beg,end) of one contig's depth map.  Indices are those of the
 * reference's per-chromosome vector<uint32_t>: index == 1-based coordinate,
 * slot 0 unused, map_size == chromosome length + 1 (sv_caller.cpp:801,
 * cnv_caller.cpp:482-487).  A whole contig is {tid, 0, map_size, map_size}.
 * Regions of one batch must not overlap.  A region also OWNS the reads whose
 * depth index pos0+1 falls in [beg,end) -- the region with end == map_size
 * additionally owns reads starting at or beyond map_size -- and only owned
 * reads emit signatures, so a read is reported exactly once however the
 * genome is sharded. */
typedef struct {
    int32_t  tid;
    uint32_t beg;
    uint32_t end;
    uint32_t map_size;
} csv_region;

/* ------------------------------------------------------- device-side batch */

/* Reads uploaded once and kept in HBM for every pass over them (the reference
 * re-reads the BAM three times). */
typedef struct csv_batch csv_batch;

/* Enqueues the copies on the context's stream and returns: pageable arrays are staged before the call returns, PINNED
 * arrays (csv_host_alloc) are read by the copy engine afterwards and must stay untouched until a call that waits for the
 * stream -- csv_ctx_sync, csv_depth_stats, csv_sigs_count or any fetch. */
int  csv_batch_upload(csv_ctx* ctx, const csv_reads* reads, uint32_t n_regions,
                      const csv_region* regions, csv_batch** out);
void csv_batch_free(csv_ctx* ctx, csv_batch* b);
/* Drops everything of a scanned batch except its RESULTS: the depth slabs (with the region tables csv_window_sums,
 * csv_depth_at, csv_depth_at_tid and the fetches need), the per-region stats, the signature columns and labels.  The
 * records, CIGAR words and event lists go back to the context's pool -- for a 30x genome 12.4 of 16 GB stay, for dense
 * ONT shards a few percent.  A caller that keeps the depth map on the device for the rest of the run (the reference reads
 * it in querySNPRegion, cnv_caller.cpp:76-113, and getReadDepth, sv_caller.cpp:1332-1344) holds one such batch per shard.
 * Afterwards csv_scan_run, csv_record_summary and csv_batch_reserve_sigs fail with CSV_ERR_STATE. */
int  csv_batch_release_inputs(csv_ctx* ctx, csv_batch* b);

/* Signature capacity of a batch.  Upload reserves max(2^20, n_ops / 16) entries (never more than n_ops); a pass that
 * emits more fails with CSV_ERR_CAPACITY when its results are fetched, and csv_sigs_count / csv_sigs_fetch then report
 * in *n_out how many entries to RESERVE (the number emitted plus the slack the 32 slot counters of a large batch need,
 * at least 1/8 more than the current capacity).  Reserve that many and run the pass again -- it deals dense slots and fits
 * as soon as the capacity covers the count.  The reference's vector has no limit, so a caller must not drop calls.
 * csv_cigar_scan does this by itself. */
int csv_batch_reserve_sigs(csv_ctx* ctx, csv_batch* b, uint64_t n_sigs);

typedef struct {
    uint32_t min_len;      /* signature length threshold; reference: 50 (sv_caller.cpp:566)   */
    uint8_t  min_mapq;     /* reference: 20 (sv_caller.h:72)                                   */
    uint8_t  want_depth;   /* run the depth part                                               */
    uint8_t  want_sigs;    /* run the signature part                                           */
    uint8_t  reserved;
} csv_scan_params;

/* One pass of the hot path over a batch, enqueued on the context's stream
 * (asynchronous, no host round trip inside):
 *   depth  == CNVCaller::calculateMeanChromosomeCoverage inner loops and
 *             reductions (cnv_caller.cpp:488-535) for every region;
 *   sigs   == SVCaller::findCIGARSVs -> processCIGARRecord -> addSVCall
 *             (sv_caller.cpp:506-661, sv_object.cpp:17-33) for every region.
 * Results stay on the device until fetched. */
int csv_scan_run(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p);

/* Per-region sum of depths and count of non-zero positions
 * (cnv_caller.cpp:534-535); the caller adds shards of one contig and divides
 * (cnv_caller.cpp:538).  Arrays have n_regions entries. */
int csv_depth_stats(csv_ctx* ctx, csv_batch* b, uint64_t* sum_out, uint32_t* nonzero_out);
/* Depth slice of one region: depth_out[i] == reference map[beg + i]. */
int csv_depth_fetch(csv_ctx* ctx, csv_batch* b, uint32_t region, uint32_t* depth_out);
/* All regions in one pipeline: depth_out[r] receives region r (NULL = skip it). */
int csv_depth_fetch_all(csv_ctx* ctx, csv_batch* b, uint32_t* const* depth_out);
/* Device address of a region's depth slice (for device-side consumers). */
int csv_depth_device_ptr(csv_ctx* ctx, csv_batch* b, uint32_t region, const uint32_t** dptr_out);

/* Signature records in the exact order of the reference's per-chromosome
 * vector<SVCall> after all addSVCall() insertions: ascending (start,end),
 * equal keys in reverse insertion order (sv_object.cpp:31-32). */
typedef struct {
    uint32_t* start;       /* SVCall.start                                                      */
    uint32_t* end;         /* SVCall.end                                                        */
    uint8_t*  kind;        /* 0 CIGARINS, 1 CIGARDEL, 2 CIGARCLIP (sv_types.h SVDataType)       */
    uint32_t* read_idx;    /* record index in csv_reads                                         */
    uint32_t* op_idx;      /* CIGAR op index inside the record                                  */
    uint32_t* query_pos;   /* reference's query_pos at the op (sv_caller.cpp:547,653-655): where */
                           /* the 50-base literal ALT starts in the read (sv_caller.cpp:572-591) */
} csv_sigs;

int csv_sigs_count(csv_ctx* ctx, csv_batch* b, uint64_t* n_out);
/* region_off_out: [n_regions+1] offsets of each region's run inside the arrays. */
int csv_sigs_fetch(csv_ctx* ctx, csv_batch* b, csv_sigs* out, uint64_t cap,
                   uint64_t* n_out, uint64_t* region_off_out);

/* DBSCAN1D::fit (dbscan1d.cpp:8-66) over the batch's sorted signature starts,
 * one independent fit per (region, SVType) group in vector order -- the
 * grouping mergeSVs uses (sv_object.cpp:61-83).  labels_out[i] belongs to
 * signature i of csv_sigs_fetch(); may be NULL to leave labels on the device. */
int csv_sigs_dbscan1d(csv_ctx* ctx, csv_batch* b, double eps, int min_pts,
                      int32_t* labels_out, uint64_t cap);

/* ------------------------------------------------ one-shot host-to-host API */

/* CNVCaller::calculateMeanChromosomeCoverage for one region (cnv_caller.h:104). */
int csv_depth(csv_ctx* ctx, const csv_reads* reads, const csv_region* region,
              uint32_t* depth_out, uint64_t* sum_out, uint32_t* nonzero_out);

/* SVCaller::findCIGARSVs for one region (sv_caller.h:86). */
int csv_cigar_scan(csv_ctx* ctx, const csv_reads* reads, const csv_region* region,
                   uint32_t min_len, uint8_t min_mapq, csv_sigs* out, uint64_t cap, uint64_t* n_out);

/* DBSCAN1D::fit + getClusters (dbscan1d.h:13-17).  labels: cluster id >= 0,
 * -2 noise, exactly as the reference assigns them.  n_clusters_out may be NULL.
 * Inputs of at most 1024 points (eps >= 0) are one kernel launch -- the reference fits one small set per cluster of
 * split alignments (sv_caller.cpp:270); the environment variable CSV_DB_SMALL=0 sends them through the general path. */
int csv_dbscan1d(csv_ctx* ctx, const int32_t* pts, uint64_t n, double eps, int min_pts,
                 int32_t* labels_out, int32_t* n_clusters_out);

/* Many independent fits in one launch sequence: seg_id[i] < n_seg names the
 * fit point i belongs to (NULL = one fit); input order inside a fit is the
 * order of appearance in pts[]. */
int csv_dbscan1d_seg(csv_ctx* ctx, const int32_t* pts, const uint32_t* seg_id, uint64_t n,
                     uint32_t n_seg, double eps, int min_pts, int32_t* labels_out,
                     int32_t* n_clusters_out /* [n_seg] or NULL */);

/* DBSCAN::fit + getClusters (include/dbscan.h:11-33, src/dbscan.cpp:9-81): the 2-D clustering mergeSVs
 * (src/sv_object.cpp:45-269) runs on the SV calls of one type -- points are intervals (start, end), the distance is
 * the minimum reciprocal overlap.  labels: cluster id >= 0, -2 noise (-1 only in the reference's own degenerate
 * min_pts <= 0 cases), exactly as the reference assigns them for this input order. */
int csv_dbscan2d(csv_ctx* ctx, const uint32_t* start, const uint32_t* end, uint64_t n, double eps, int min_pts,
                 int32_t* labels_out);

/* DBSCAN1D::getLargestCluster (dbscan1d.cpp:72-90) on labels from a fit: host-side
 * helper, no device work.  Returns the number of points written to out. */
uint64_t csv_largest_cluster(const int32_t* pts, const int32_t* labels, uint64_t n, int32_t* out);

/* Window depth sums for CNVCaller::querySNPRegion (cnv_caller.cpp:76-113):
 * integer sum and position count of each of the sample_size windows of
 * [start_pos,end_pos], read from a region's device-resident depth.  The
 * caller finishes with the same libm log2 as the reference.  A region that
 * is a slice of its contig returns its share (the positions inside [beg,end)):
 * shares of the shards of one contig add up. */
int csv_window_sums(csv_ctx* ctx, csv_batch* b, uint32_t region, uint32_t n_sv,
                    const uint32_t* start_pos, const uint32_t* end_pos, int sample_size,
                    uint64_t* sum_out /* [n_sv*sample_size] */, uint32_t* count_out);

/* SVCaller::getReadDepth (sv_caller.cpp:1332-1344) for many positions at once, read from a region's
 * device-resident depth: depth_out[i] == map[positions[i]], 0 beyond the map (the reference catches the out_of_range
 * and adds nothing) and 0 outside the region's slice [beg,end).  With csv_window_sums this serves every consumer of
 * the depth map -- its size, the log2 windows, the VCF's DP / SUPPORT -- without the map crossing PCIe. */
int csv_depth_at(csv_ctx* ctx, csv_batch* b, uint32_t region, uint64_t n, const uint32_t* positions, uint32_t* depth_out);
/* The same for (contig, position) pairs anywhere in the batch: the slice that holds each position is looked up on the
 * device.  0xffffffff marks a position inside its contig that none of this batch's regions covers (another shard's). */
int csv_depth_at_tid(csv_ctx* ctx, csv_batch* b, uint64_t n, const int32_t* tid, const uint32_t* positions, uint32_t* depth_out);
/* ... and for every signature of the batch at its start (what saveToVCF asks for each call, sv_caller.cpp:1306), in
 * csv_sigs_fetch order, without the positions leaving the device.  Same 0xffffffff convention. */
int csv_sigs_depth(csv_ctx* ctx, csv_batch* b, uint32_t* depth_out, uint64_t cap);
/* Position-weighted checksum of every region's depth slice: sum of depth[i] * mix64(tid << 32 | index) modulo 2^64
 * (splitmix64 finaliser).  Additive: the checksums of the shards of a contig add up to the whole contig's, so a
 * sharded run can be verified against a single-device run without the maps leaving the devices. */
int csv_depth_checksum(csv_ctx* ctx, csv_batch* b, uint64_t* checksum_out /* [n_regions] */);

/* Diagnostics: copies `bytes` bytes at `offset` of one of the batch's intermediate device arrays to the host (waits for
 * the pass).  name: "events", "ev_start", "ref_end", "span_desc", "pmax", "tile_q", "tile_r", "meta", "key".  *size_out
 * (optional) receives the bytes the array holds; bytes == 0 only queries it.  Not part of the drop-in path. */
int csv_debug_fetch(csv_ctx* ctx, csv_batch* b, const char* name, uint64_t offset, uint64_t bytes, void* out, uint64_t* size_out);

/* Per-record summaries the split-read pass starts from (SVCaller::detectSVsFromSplitReads, sv_caller.cpp:150-162):
 * bam_endpos(b) and SVCaller::getAlignmentReadPositions(b) (sv_caller.cpp:663-690) for every record of the batch, in
 * csv_reads order.  Any output may be NULL.  The batch needs no scan first. */
int csv_record_summary(csv_ctx* ctx, csv_batch* b, int32_t* endpos_out, int32_t* query_start_out, int32_t* query_end_out);

/* Host helper for packers that fill csv_reads::n_gap after the fact: n_gap_out[i] = number of D / N ops of record i
 * (threads = 0: all cores).  No device needed. */
void csv_host_count_gaps(const uint32_t* cigar, const uint64_t* cig_off, uint32_t n_reads, uint32_t* n_gap_out, int threads);
/* ... and csv_reads::ref_len with them (either output may be NULL). */
void csv_host_record_stats(const uint32_t* cigar, const uint64_t* cig_off, uint32_t n_reads, uint32_t* n_gap_out, uint32_t* ref_len_out, int threads);

/* ------------------------------------------------------- synthetic inputs */

/* Seeded generator of coordinate-sorted long-read alignments.
 * Host-only; lives in libcsvsynth.so, which has no CUDA dependency. */
typedef struct {
    uint64_t seed;
    int32_t  profile;           /* 0 HiFi (normal lengths), 1 ONT (lognormal, read_len_mean = N50) */
    double   coverage;
    double   read_len_mean, read_len_sd;
    double   indel_rate;        /* small indel events per reference base                           */
    uint32_t indel_len_max;
    uint64_t n_sv;              /* structural variants over all contigs                            */
    uint32_t sv_len_max;
    double   sv_jitter_sd;      /* per-read breakpoint jitter (config 5)                           */
    double   frac_len50;        /* SVs of length exactly 50                                        */
    double   frac_softclip, frac_supplementary, frac_secondary, frac_dup, frac_qcfail, frac_lowmapq;
    int32_t  use_eqx;           /* '=' ops instead of 'M'                                          */
    int32_t  threads;           /* 0 = all cores                                                   */
} csv_synth_params;

void     csv_synth_default_params(csv_synth_params* p);
uint64_t csv_synth_num_reads(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len);
int      csv_synth_reads(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len,
                         int32_t* tid, int32_t* pos0, uint16_t* flag, uint8_t* mapq,
                         uint64_t* cig_off /* [n_reads+1] */, uint64_t* n_ops_out);
int      csv_synth_cigar(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len,
                         const uint64_t* cig_off, uint32_t* cigar);

#ifdef __cplusplus
}
#endif
#endif

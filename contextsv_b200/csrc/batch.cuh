// batch.cuh -- device-resident batch of packed reads and the per-pass state.
#pragma once
#include "common.cuh"

namespace csv {
constexpr uint32_t kNone = 0xffffffffu;

// u32 slots of csv_batch::d_scalars
enum { SC_N_NONEMPTY = 0, SC_N_SIG = 1, SC_UNSORTED = 2, SC_N_SIG_EFF = 3 /* min(SC_N_SIG, sig_cap) */, SC_N_WIDE = 4 /* tiles on the wide list */, SC_ABSURD = 5 /* a record spans >= 2^31 reference bases */, SC_HAS_EMPTY = 6 /* a record without CIGAR: tables need compaction */, SC_BAD_GAPS = 7 /* csv_reads::n_gap does not match the CIGAR */, SC_SIG_DROPPED = 8 /* a signature found its slot beyond sig_cap */,
       SC_SIG_SUB0 = 16 /* kSigSub slot counters: a signature of CTA c draws from counter c % kSigSub and owns raw slot index * kSigSub + c % kSigSub (one counter for all was 1.4 ns per signature: same-address atomics); SC_N_SIG is their sum, formed on the host */,
       SC_COUNT = 48 };
constexpr uint32_t kSigSub = 32;

struct SigRaw {          // emission-order signature records (device)
    unsigned long long* key_hi;   // owner region << 32 | start
    unsigned long long* key_lo;   // end << 32 | ~global op index  (ties: reverse insertion order)
    uint32_t* k;                  // compact (non-empty) read index
    uint8_t* kind;                // bit 7: query_pos needs the exact sequential recomputation
    uint32_t* bucket;             // depth tile (clamped into the owner region's) the start falls into: the ordering's bucket
    uint32_t* arrival;            // arrival number inside the bucket (any order: ranked by key afterwards)
};
}  // namespace csv

namespace csv {
// One stage of the pipelined pass: the walk covers spans [span0, span1); once the walk of the NEXT chunk is done,
// every record of the contigs [first_tid, next chunk's first_tid) has its events, and the tiles of their regions run
// on the tile stream beside the walk of later chunks.
struct PipeChunk {
    uint32_t span0 = 0, span1 = 0;
    uint32_t first_tid = 0;
    uint32_t rec_upper = 0;                                   // records of the chunk's contigs (upper bound: empty ones included)
    std::vector<std::pair<uint32_t, uint32_t>> tiles;         // contiguous tile ranges of the chunk's regions
};
}  // namespace csv

struct csv_batch {
    uint32_t n_reads = 0;
    uint64_t n_ops = 0;
    uint32_t n_regions = 0, n_tids = 0, n_tiles = 0, n_spans = 0;
    bool has_tid = false, multi_region_tid = false;
    std::vector<csv_region> regions;          // caller order
    std::vector<uint32_t> tile_base;          // caller order, n_regions + 1
    uint64_t ev_cap = 0, sig_cap = 0;
    uint32_t last_min_len = 50;
    bool scanned = false, have_depth = false, have_sigs = false, have_labels = false;
    bool rec_prepass = false;                 // the caller supplied n_gap[] and records are short: record-level pre-pass (walk.cu)
    bool claimed_ref = false;                 // ... and csv_reads::ref_len: the tile ranges are ready before the walk ends (depth_tiles.cu)
    bool split_upload = false;                // the CIGAR words went up chunk by chunk on the upload stream (csv_ctx::ev_upload): the walk of chunk c only waits for its own
    bool inputs_released = false;             // csv_batch_release_inputs: only the results are left (depth slabs, signature columns, labels)

    // input SoA (device)
    csv::DevBuf d_tid, d_pos0, d_flag, d_mapq, d_cig_off, d_cigar, d_n_gap, d_ref_len, d_ref_chk, d_ev_check;
    // derived per-read tables
    csv::DevBuf d_meta;      // uint4 {pos0, map_size | 0 (contig not requested), flag | mapq << 16, owner region | kNone} per non-empty read
    csv::DevBuf d_key;       // u64 (tid << 32 | pos0 + 1) per non-empty read: the batch's sort key
    csv::DevBuf d_ne_idx;    // compact index -> record index
    csv::DevBuf d_headbits;  // one bit per op: op is the first of its record (+ sentinel bit at n_ops)
    csv::DevBuf d_scalars;   // SC_* counters
    csv::DevBuf d_regs, d_tids;
    // per-pass state, ONE allocation (d_scalars owns it) and one memset at the start of a pass:
    // [scalars (SC_COUNT) | tile tickets of the chained kernels (one per pipeline chunk, + 4) | signatures per region (n_regions)]
    csv::DevView d_tickets, d_reg_sig_cnt;
    size_t state_bytes = 0;
    uint32_t sig_sub_mask = csv::kSigSub - 1;  // see SC_SIG_SUB0
    csv::DevBuf d_reg_tab;   // u32 [tile_base (n_regions + 1) | len (n_regions) | beg (n_regions)], caller order
    // walk
    csv::DevBuf d_span_agg, d_span_pre, d_span_status, d_scan_carry, d_span_desc, d_span_rq;
    std::vector<csv::PipeChunk> chunks;
    csv::DevBuf d_chunk_tid, d_chunk_bounds;
    // depth
    csv::DevBuf d_events;    // uint32 depth-map indices, sign = slot parity
    csv::DevBuf d_ev_start, d_ref_end, d_pmax, d_pmax_part;   // per non-empty read
    csv::DevBuf d_depth, d_sum, d_nz, d_tile_desc, d_tile_ev, d_tile_sum, d_tile_nz, d_wide_list, d_tile_q, d_tile_r;
    // signatures
    csv::DevBuf d_sig_hi, d_sig_lo, d_sig_k, d_sig_kind, d_sig_payload, d_sig_bucket, d_sig_arrival;
    csv::DevBuf d_bucket_cnt, d_bucket_base;   // u32 per depth tile: signatures that start there (zero between passes), first slot of the bucket
    csv::DevBuf d_out_start, d_out_end, d_out_kind, d_out_read, d_out_op, d_out_qpos, d_out_seg, d_labels;

    void release(csv::DevPool* pool = nullptr) {
        csv::DevBuf* all[] = {&d_tid, &d_pos0, &d_flag, &d_mapq, &d_cig_off, &d_cigar, &d_n_gap, &d_ref_len, &d_ref_chk, &d_ev_check, &d_span_rq, &d_meta, &d_key, &d_ne_idx, &d_headbits, &d_scalars,
                              &d_regs, &d_tids, &d_reg_tab, &d_span_agg, &d_span_pre, &d_span_status, &d_scan_carry, &d_span_desc, &d_chunk_tid, &d_chunk_bounds,
                              &d_events, &d_ev_start, &d_ref_end, &d_pmax, &d_pmax_part, &d_depth, &d_sum, &d_nz, &d_tile_desc, &d_tile_ev, &d_tile_sum, &d_tile_nz, &d_wide_list, &d_tile_q, &d_tile_r, &d_sig_hi, &d_sig_lo, &d_sig_k,
                              &d_sig_kind, &d_sig_payload, &d_sig_bucket, &d_sig_arrival, &d_bucket_cnt, &d_bucket_base, &d_out_start, &d_out_end, &d_out_kind, &d_out_read, &d_out_op,
                              &d_out_qpos, &d_out_seg, &d_labels};
        for (auto* b : all) b->release(pool);
    }
    // everything a pass reads or writes on the way to the results; what stays serves the depth consumers and the fetches
    void release_inputs(csv::DevPool* pool) {
        csv::DevBuf* in[] = {&d_tid, &d_pos0, &d_flag, &d_mapq, &d_cig_off, &d_cigar, &d_n_gap, &d_ref_len, &d_ref_chk, &d_ev_check, &d_span_rq, &d_meta, &d_key, &d_ne_idx, &d_headbits,
                             &d_span_agg, &d_span_pre, &d_span_status, &d_scan_carry, &d_span_desc, &d_chunk_tid, &d_chunk_bounds,
                             &d_events, &d_ev_start, &d_ref_end, &d_pmax, &d_pmax_part, &d_tile_desc, &d_tile_ev, &d_tile_sum, &d_tile_nz, &d_wide_list, &d_tile_q, &d_tile_r,
                             &d_sig_hi, &d_sig_lo, &d_sig_k, &d_sig_kind, &d_sig_payload, &d_sig_bucket, &d_sig_arrival, &d_bucket_cnt, &d_bucket_base};
        for (auto* b : in) b->release(pool);
        inputs_released = true;
    }
};

namespace csv {
// kernels' host launchers (each enqueues on ctx->stream and bumps ctx->launches)
int launch_prep(csv_ctx* ctx, csv_batch* b, uint32_t min_mapq);
int launch_record_prepass(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p, int what = 3, uint32_t span_from = 0, uint32_t span_to = 0xffffffffu);
int wait_upload(csv_ctx* ctx, csv_batch* b, uint32_t chunk_or_all);   // main stream waits for the CIGAR words of a chunk (0xffffffff: of all)
int launch_claim_check(csv_ctx* ctx, csv_batch* b);
int launch_ev_check(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p);
int launch_walk(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p, uint32_t span0, uint32_t span1);
int launch_chunk_bounds(csv_ctx* ctx, csv_batch* b);
int launch_tile_hi(csv_ctx* ctx, csv_batch* b);
int launch_tile_ranges(csv_ctx* ctx, csv_batch* b, uint32_t chunk, int what = 3);
int launch_depth_begin(csv_ctx* ctx, csv_batch* b);
int launch_depth_tiles(csv_ctx* ctx, csv_batch* b, uint32_t chunk);
int launch_depth_finish(csv_ctx* ctx, csv_batch* b, int what = 3);
int launch_sig_finish(csv_ctx* ctx, csv_batch* b);
int launch_sig_dbscan(csv_ctx* ctx, csv_batch* b, double eps, int min_pts);
bool sig_order_radix();   // CSV_SIG_ORDER=radix: the sort-based order (one slot counter, dense raw slots)
int launch_window_sums(csv_ctx* ctx, csv_batch* b, uint32_t region, uint32_t n_sv, const uint32_t* d_start,
                       const uint32_t* d_end, int sample_size, unsigned long long* d_sum, uint32_t* d_cnt);

// radix sort of (hi, lo) 128-bit keys with a u32 payload; hi may be null (64-bit keys).
// n_dev (device u32) optionally overrides n_upper.
struct SortBufs {
    unsigned long long *hi, *lo; uint32_t* val;          // input / ping
    unsigned long long *hi2, *lo2; uint32_t* val2;       // pong
};
// digit_mask: bit d set = byte d of the 128-bit key (lo bytes 0-7, hi bytes 8-15) may vary.
// The sorted data always ends up in bufs.hi / lo / val.
// first (optional): the sort opens a pipeline whose element count is still raw -- its first kernel clamps *n_raw to
// clamp_cap, publishes the result in *n_clamped_out (which must then be the sort's n_dev) and, with iota, fills bufs.val
// with 0, 1, 2, ...: two tiny launches less on the signature side stream.
struct SortFirst { const uint32_t* n_raw; uint32_t clamp_cap; uint32_t* n_clamped_out; bool iota; };
int radix_sort_pairs(csv_ctx* ctx, SortBufs bufs, uint64_t n_upper, const uint32_t* n_dev, uint32_t digit_mask, const SortFirst* first = nullptr);

// DBSCAN1D on device-resident points.  d_seg may be null (single fit).
int dbscan1d_sorted_sigs(csv_ctx* ctx, const int32_t* d_pts, const uint32_t* d_seg, uint64_t n_upper, const uint32_t* n_dev,
                         uint32_t n_seg, double eps, int min_pts, int32_t* d_labels);
int dbscan1d_device(csv_ctx* ctx, const int32_t* d_pts, const uint32_t* d_seg, uint64_t n_upper, const uint32_t* n_dev,
                    uint32_t n_seg, double eps, int min_pts, int32_t* d_labels, int32_t* d_n_clusters /* [n_seg] or null */,
                    bool value_sorted = false /* points of one segment come in ascending order */);
// one fit of at most kDbSmallMax host-resident points, eps >= 0: one launch (dbscan_small.h)
int dbscan1d_small(csv_ctx* ctx, const int32_t* pts, uint32_t n, double eps, int min_pts, int32_t* labels_out, int32_t* n_clusters_out);
// DBSCAN::fit (2-D, reciprocal-overlap distance) on device-resident intervals
int dbscan2d_device(csv_ctx* ctx, const uint32_t* d_start, const uint32_t* d_end, uint64_t n, double eps, int min_pts, int32_t* d_labels);
}  // namespace csv

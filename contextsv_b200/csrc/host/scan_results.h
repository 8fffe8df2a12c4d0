// scan_results.h -- what ONE decode of the BAM leaves behind for the rest of SVCaller::run.
//
// The reference decodes the file three times (depth pass, CIGAR pass, split-read pass) and keeps a uint32 per base of
// every chromosome in host memory between them.  The drop-in decodes once, in the depth pass, and parks per contig:
//   * the depth map, DEVICE-resident: one csv_batch per shard with everything but its results released
//     (csv_batch_release_inputs).  The three places the reference reads the map -- its size (sv_caller.cpp:602), the
//     log2 windows (cnv_caller.cpp:76-113) and getReadDepth (sv_caller.cpp:1332-1344) -- are served from there
//     (csv_window_sums / csv_depth_at), found through the address of the caller's std::vector;
//   * the CIGAR signatures in the order of the reference's vector (what findCIGARSVs produces for the contig), with the
//     4-bit bases of the records that need a literal 50-base ALT;
//   * the per-record summaries the split-read pass starts from (sv_caller.cpp:140-162), in file order.
// Devices come from CONTEXTSV_GPUS; every device has ONE context for these batches, guarded by a mutex (a csv_ctx is
// not thread-safe, and the queries come from whichever thread runs the consumer).
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "contextsv_b200.h"

namespace csvhost {

struct Device {
    int id = 0;
    csv_ctx* ctx = nullptr;             // created on first use
    std::mutex m;                       // held around every use of ctx
};
size_t device_count();                  // entries of CONTEXTSV_GPUS (default "0")
Device& device_at(size_t i);            // context created lazily by ensure_context()
csv_ctx* ensure_context(Device& d);     // call with d.m held; throws std::runtime_error without a usable GPU (no CPU fallback)

struct DepthShard {
    Device* dev = nullptr;
    csv_batch* batch = nullptr;
    uint32_t region = 0, beg = 0, end = 0;
};

struct SigColumns {                     // one entry per signature, addSVCall order of the contig
    std::vector<uint32_t> start, end, op_idx, query_pos;
    std::vector<uint64_t> serial;       // record number in the contig's iterator order
    std::vector<uint8_t> kind;
    size_t size() const { return start.size(); }
};

struct SplitRecords {                   // records that pass the split-read filter's flag part, iterator order
    std::vector<int32_t> pos, endpos, query_start, query_end;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq;
    std::vector<char> names;
    std::vector<uint64_t> name_off{0};
    size_t size() const { return pos.size(); }
};

struct ContigResults {
    std::string bam_path;
    int tid = -1;
    uint32_t map_size = 0;
    std::vector<DepthShard> shards;                                   // ascending beg, disjoint, covering [0, map_size)
    bool have_sigs = false;
    uint8_t sig_min_mapq = 0;
    SigColumns sigs;
    std::unordered_map<uint64_t, std::vector<uint8_t>> seq4;          // by record serial
    bool have_split = false;
    SplitRecords split;
};

// ---- registry (thread-safe).  Results are owned by the registry until clear_results() / process exit.
void put_results(const void* depth_vector_key, std::shared_ptr<ContigResults> r);
std::shared_ptr<ContigResults> results_for_vector(const void* depth_vector_key);
std::shared_ptr<ContigResults> results_for_contig(const char* bam_path, int tid);
void clear_results();                                                 // frees the device batches

// ---- consumers of the device-resident map
// depth_out[i] = map[pos[i]], 0 beyond the map (what getReadDepth adds after catching std::out_of_range)
bool device_depth_at(const ContigResults& r, const uint32_t* pos, size_t n, uint32_t* depth_out);
// window sums / counts of querySNPRegion for n_sv regions of sample_size windows each (shares of the shards added)
bool device_window_sums(const ContigResults& r, uint32_t n_sv, const uint32_t* start, const uint32_t* end, int sample_size,
                        uint64_t* sum_out, uint32_t* count_out);

// ---- prefetch caches: a consumer that knows its queries in advance asks for all of them in one launch; the
// per-call overrides look here first.
void prefetch_depth_at(const void* depth_vector_key, const std::vector<uint32_t>& positions);
bool prefetched_depth(const void* depth_vector_key, uint32_t pos, uint32_t* out);
void prefetch_windows(const void* depth_vector_key, const std::vector<uint32_t>& start, const std::vector<uint32_t>& end, int sample_size);
// the sums of [start, end] with sample_size windows, if prefetched
bool prefetched_windows(const void* depth_vector_key, uint32_t start, uint32_t end, int sample_size, const uint64_t** sums, const uint32_t** counts);
void drop_prefetch(const void* depth_vector_key);

bool host_depth_requested();            // CONTEXTSV_HOST_DEPTH=1: the depth pass also fills the caller's vectors

}  // namespace csvhost

// cnv_caller_gpu.cpp -- drop-in definition of CNVCaller::calculateMeanChromosomeCoverage
// (include/cnv_caller.h:104, src/cnv_caller.cpp:415-556): same signature, same messages, same
// containers filled; the per-base loop and the two reductions run on the GPU through the C ABI.
// Chromosomes are scanned one after the other, each in as many shards as its CIGAR ops need.
#include "cnv_caller.h"

#include <htslib/sam.h>

#include <algorithm>

#include "contextsv_b200.h"
#include "gpu_context.h"
#include "packed_reads.h"

void CNVCaller::calculateMeanChromosomeCoverage(const std::vector<std::string>& chromosomes,
                                                std::unordered_map<std::string, std::vector<uint32_t>>& chr_pos_depth_map,
                                                std::unordered_map<std::string, double>& chr_mean_cov_map,
                                                const std::string& bam_filepath, int thread_count) const
{
    csvhost::StatTimer st_all(csvhost::STAT_DEPTH, chromosomes.size());
    csvhost::warm_up_async();
    printMessage("Opening BAM file: " + bam_filepath);
    samFile* bam_file = sam_open(bam_filepath.c_str(), "r");
    if (!bam_file) { printError("ERROR: Could not open BAM file: " + bam_filepath); return; }
    hts_set_threads(bam_file, thread_count);
    bam_hdr_t* bam_header = sam_hdr_read(bam_file);
    if (!bam_header) { sam_close(bam_file); printError("ERROR: Could not read header from BAM file: " + bam_filepath); return; }
    hts_idx_t* bam_index = sam_index_load(bam_file, bam_filepath.c_str());
    if (!bam_index) { bam_hdr_destroy(bam_header); sam_close(bam_file); printError("ERROR: Could not load index for BAM file: " + bam_filepath); return; }
    bam1_t* bam_record = bam_init1();
    if (!bam_record) { bam_hdr_destroy(bam_header); sam_close(bam_file); printError("ERROR: Could not initialize BAM record."); return; }

    int current_chr = 0;
    const int total_chr_count = (int)chromosomes.size();
    std::vector<std::pair<int, std::string>> by_tid;
    for (const std::string& chr : chromosomes) {
        hts_itr_t* bam_iter = sam_itr_querys(bam_index, bam_header, chr.c_str());
        if (!bam_iter) { printError("ERROR: Could not create iterator for chromosome: " + chr + ", check if the chromosome exists in the BAM file."); continue; }
        printMessage("(" + std::to_string(++current_chr) + "/" + std::to_string(total_chr_count) + ") Reading BAM file for chromosome: " + chr);
        std::vector<uint32_t>& pos_depth_map = chr_pos_depth_map[chr];
        const int tid = bam_name2id(bam_header, chr.c_str());
        if (tid < 0) { printError("ERROR: Could not find chromosome " + chr + " in BAM file."); hts_itr_destroy(bam_iter); continue; }
        const uint32_t chr_length = bam_header->target_len[tid] + 1;
        if (pos_depth_map.size() != static_cast<size_t>(chr_length)) {
            printError("ERROR: Chromosome length mismatch for " + chr + ": expected " + std::to_string(chr_length) + ", found " +
                       std::to_string(pos_depth_map.size()) + ", resizing to " + std::to_string(chr_length));
            pos_depth_map.resize(chr_length, 0);
        }
        by_tid.emplace_back(tid, chr);
        hts_itr_destroy(bam_iter);
    }
    // Contigs in header order.  Records stream into a shard; whenever the shard would exceed the ops one batch may hold
    // (a batch takes < 2^31: 60x ONT is ~3.7e10 over the genome), the depth slice up to the current position is
    // scanned, and only the records that reach past the cut stay as the halo of the next shard -- the region sharding
    // of SURVEY 8e, in time instead of across GPUs, with bounded host memory.
    std::sort(by_tid.begin(), by_tid.end());
    csv_ctx* ctx = nullptr;                                    // created at the first flush: CUDA start-up runs beside the decoding
    const uint64_t max_ops = csvhost::max_ops_per_batch();
    bool failed = false;
    for (const auto& tc : by_tid) {
        hts_itr_t* it = sam_itr_querys(bam_index, bam_header, tc.second.c_str());
        if (!it) continue;
        std::vector<uint32_t>& depth = chr_pos_depth_map[tc.second];
        const uint32_t size = bam_header->target_len[tc.first] + 1;
        uint64_t cum_depth = 0;
        uint32_t pos_count = 0, beg = 0;
        csvhost::PackedReads reads;
        auto flush = [&](uint32_t end) {                       // depth slice [beg, end) from the records packed so far
            if (end <= beg || failed) { beg = std::max(beg, end); return; }
            const csv_region reg = {tc.first, beg, end, size};
            const csv_reads view = reads.view();
            uint64_t sum = 0; uint32_t nz = 0;
            if (!ctx) ctx = csvhost::thread_context();
            csvhost::StatTimer st(csvhost::STAT_DEPTH_GPU, view.n_reads);
            if (csv_depth(ctx, &view, &reg, depth.data() + beg, &sum, &nz) != CSV_OK) {
                printError(std::string("ERROR: GPU depth pass failed: ") + csv_last_error());
                failed = true;
            }
            cum_depth += sum; pos_count += nz;
            beg = end;
        };
        int32_t last_pos = -2;
        bool whole = true;                                     // the contig is still one shard
        while (sam_itr_next(bam_file, it, bam_record) >= 0) {
            const int32_t pos = (int32_t)bam_record->core.pos;
            if (reads.ops() + reads.size() + bam_record->core.n_cigar + 1 > max_ops && pos != last_pos && reads.size() > 0) {   // ops + records: a batch counts both
                const uint32_t cut = std::min<uint32_t>((uint32_t)pos + 1u, size);       // records from here on start at or after the cut
                flush(cut);
                reads.keep_reaching(cut);
                whole = false;
            }
            reads.append(bam_record, true);                    // with the bases of the few records the CIGAR pass will want them for
            last_pos = pos;
        }
        hts_itr_destroy(it);
        flush(size);
        if (failed) break;
        if (whole) csvhost::cache_put(bam_filepath, tc.first, std::move(reads));   // the CIGAR pass scans the same records
        const double mean_chr_cov = (pos_count > 0) ? static_cast<double>(cum_depth) / static_cast<double>(pos_count) : 0.0;
        printMessage("Mean coverage for chromosome " + tc.second + ": " + std::to_string(mean_chr_cov));
        if (mean_chr_cov != 0.0) chr_mean_cov_map[tc.second] = mean_chr_cov;
    }

    printMessage("Closing BAM file " + bam_filepath);
    bam_destroy1(bam_record);
    hts_idx_destroy(bam_index);
    bam_hdr_destroy(bam_header);
    sam_close(bam_file);
    printMessage("BAM file closed.");
}

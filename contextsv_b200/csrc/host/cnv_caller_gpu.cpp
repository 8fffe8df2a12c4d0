// cnv_caller_gpu.cpp -- drop-in definition of CNVCaller::calculateMeanChromosomeCoverage
// (include/cnv_caller.h:104, src/cnv_caller.cpp:415-556): same signature, same messages, same
// containers filled; the per-base loop and the two reductions run on the GPU through the C ABI.
// One decode of the BAM feeds every chromosome in one batch (the reference iterates per chromosome).
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

    // pack every requested chromosome (file order == coordinate order inside a chromosome)
    csvhost::PackedReads reads;
    std::vector<csv_region> regions;
    std::vector<std::string> region_chr;
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
    // the batch must be coordinate-sorted across contigs: visit the contigs in header order
    std::sort(by_tid.begin(), by_tid.end());
    for (const auto& tc : by_tid) {
        hts_itr_t* it = sam_itr_querys(bam_index, bam_header, tc.second.c_str());
        if (!it) continue;
        csvhost::pack_iterator(bam_file, it, bam_record, reads, false);
        hts_itr_destroy(it);
        const uint32_t size = bam_header->target_len[tc.first] + 1;
        regions.push_back(csv_region{tc.first, 0u, size, size});
        region_chr.push_back(tc.second);
    }

    if (!regions.empty()) {
        csv_ctx* ctx = csvhost::thread_context();
        const csv_reads view = reads.view();
        csv_batch* batch = nullptr;
        const csv_scan_params params = {50, 20, 1, 0, 0};
        std::vector<uint64_t> sums(regions.size());
        std::vector<uint32_t> nonzero(regions.size());
        int rc = csv_batch_upload(ctx, &view, (uint32_t)regions.size(), regions.data(), &batch);
        if (rc == CSV_OK) rc = csv_scan_run(ctx, batch, &params);
        if (rc == CSV_OK) rc = csv_depth_stats(ctx, batch, sums.data(), nonzero.data());
        for (size_t i = 0; rc == CSV_OK && i < regions.size(); i++)
            rc = csv_depth_fetch(ctx, batch, (uint32_t)i, chr_pos_depth_map[region_chr[i]].data());
        csv_batch_free(ctx, batch);
        if (rc != CSV_OK) printError(std::string("ERROR: GPU depth pass failed: ") + csv_last_error());
        else for (size_t i = 0; i < regions.size(); i++) {
            const uint64_t cum_depth = sums[i];
            const uint32_t pos_count = nonzero[i];
            const double mean_chr_cov = (pos_count > 0) ? static_cast<double>(cum_depth) / static_cast<double>(pos_count) : 0.0;
            printMessage("Mean coverage for chromosome " + region_chr[i] + ": " + std::to_string(mean_chr_cov));
            if (mean_chr_cov != 0.0) chr_mean_cov_map[region_chr[i]] = mean_chr_cov;
        }
    }

    printMessage("Closing BAM file " + bam_filepath);
    bam_destroy1(bam_record);
    hts_idx_destroy(bam_index);
    bam_hdr_destroy(bam_header);
    sam_close(bam_file);
    printMessage("BAM file closed.");
}

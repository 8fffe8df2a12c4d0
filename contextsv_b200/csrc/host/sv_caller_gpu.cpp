// sv_caller_gpu.cpp -- drop-in definition of SVCaller::findCIGARSVs (include/sv_caller.h:86,
// src/sv_caller.cpp:506-537).  The records of the region are packed through the same iterator the
// reference uses, the CIGAR walk + sorted insertion order come from the GPU (csv_cigar_scan returns the
// signatures in the order of the reference's vector after all addSVCall() insertions), and the
// SVCall objects -- including the 50-base literal ALT allele -- are materialised here.
#include "sv_caller.h"

#include <htslib/sam.h>

#include <algorithm>

#include "contextsv_b200.h"
#include "gpu_context.h"
#include "packed_reads.h"

void SVCaller::findCIGARSVs(samFile* fp_in, hts_idx_t* idx, bam_hdr_t* bamHdr, const std::string& region, std::vector<SVCall>& sv_calls,
                            const std::vector<uint32_t>& pos_depth_map)
{
    bam1_t* bam1 = bam_init1();
    if (!bam1) { printError("ERROR: failed to initialize BAM record"); return; }
    hts_itr_t* itr = sam_itr_querys(idx, bamHdr, region.c_str());
    if (!itr) { bam_destroy1(bam1); printError("ERROR: failed to query region " + region); return; }
    csvhost::PackedReads reads;
    while (readNextAlignment(fp_in, itr, bam1) >= 0) reads.append(bam1, true);
    hts_itr_destroy(itr);
    bam_destroy1(bam1);
    if (reads.size() == 0) return;

    const int tid = reads.tid[0];
    const uint32_t map_size = (uint32_t)pos_depth_map.size();       // only the size of the depth map is consulted (sv_caller.cpp:602)
    if (map_size == 0) return;
    const csv_region reg = {tid, 0u, map_size, map_size};
    const csv_reads view = reads.view();
    csv_ctx* ctx = csvhost::thread_context();
    uint64_t n = 0, cap = 1u << 16;
    std::vector<uint32_t> start, end, read_idx, op_idx, query_pos;
    std::vector<uint8_t> kind;
    for (;;) {
        start.resize(cap); end.resize(cap); read_idx.resize(cap); op_idx.resize(cap); query_pos.resize(cap); kind.resize(cap);
        csv_sigs out = {start.data(), end.data(), kind.data(), read_idx.data(), op_idx.data(), query_pos.data()};
        const int rc = csv_cigar_scan(ctx, &view, &reg, 50, (uint8_t)this->min_mapq, &out, cap, &n);
        if (rc == CSV_OK) break;
        if (rc == CSV_ERR_CAPACITY && n > cap) { cap = n; continue; }
        printError(std::string("ERROR: GPU CIGAR scan failed: ") + csv_last_error());
        return;
    }
    std::vector<SVCall> found;
    found.reserve(n);
    const double default_lh = 0.0;
    for (uint64_t i = 0; i < n; i++) {
        SVEvidenceFlags aln_type;
        if (kind[i] == 1) {
            aln_type.set(static_cast<size_t>(SVDataType::CIGARDEL));
            found.emplace_back(start[i], end[i], SVType::DEL, getSVTypeSymbol(SVType::DEL), aln_type, Genotype::UNKNOWN, default_lh, 0, 0, 0);
            continue;
        }
        aln_type.set(static_cast<size_t>(kind[i] == 0 ? SVDataType::CIGARINS : SVDataType::CIGARCLIP));
        std::string alt_allele = "<INS>";
        const uint32_t op_len = end[i] - start[i] + 1;
        if (op_len <= 50) {                                         // literal sequence for a 50-base event (sv_caller.cpp:587-591)
            const auto it = reads.seq4.find(read_idx[i]);
            if (it != reads.seq4.end()) {
                alt_allele.assign(op_len, ' ');
                for (uint32_t j = 0; j < op_len; j++) alt_allele[j] = csvhost::base_at(it->second, query_pos[i] + j);
            }
        }
        found.emplace_back(start[i], end[i], SVType::INS, alt_allele, aln_type, Genotype::UNKNOWN, default_lh, 0, 0, 0);
    }
    if (sv_calls.empty()) { sv_calls.swap(found); return; }
    // a non-empty target vector: replay addSVCall in the reference's insertion order (record, op)
    std::vector<size_t> order(found.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        return read_idx[a] != read_idx[b] ? read_idx[a] < read_idx[b] : op_idx[a] < op_idx[b];
    });
    for (size_t i : order) addSVCall(sv_calls, found[i]);
}

// sv_caller_gpu.cpp -- drop-in definition of SVCaller::findCIGARSVs (include/sv_caller.h:86,
// src/sv_caller.cpp:506-537).  The records of the region are packed through the same iterator the
// reference uses, the CIGAR walk + sorted insertion order come from the GPU (csv_cigar_scan returns the
// signatures in the order of the reference's vector after all addSVCall() insertions), and the
// SVCall objects -- including the 50-base literal ALT allele -- are materialised here.
#include "sv_caller.h"

#include <htslib/sam.h>

#include <algorithm>
#include <memory>

#include "contextsv_b200.h"
#include "gpu_context.h"
#include "packed_reads.h"

void SVCaller::findCIGARSVs(samFile* fp_in, hts_idx_t* idx, bam_hdr_t* bamHdr, const std::string& region, std::vector<SVCall>& sv_calls,
                            const std::vector<uint32_t>& pos_depth_map)
{
    csvhost::StatTimer st_all(csvhost::STAT_CIGAR, 1);
    csvhost::warm_up_async();
    bam1_t* bam1 = bam_init1();
    if (!bam1) { printError("ERROR: failed to initialize BAM record"); return; }
    hts_itr_t* itr = sam_itr_querys(idx, bamHdr, region.c_str());
    if (!itr) { bam_destroy1(bam1); printError("ERROR: failed to query region " + region); return; }
    const uint32_t map_size = (uint32_t)pos_depth_map.size();       // only the size of the depth map is consulted (sv_caller.cpp:602)
    csv_ctx* ctx = nullptr;                                         // created at the first flush
    const uint64_t max_ops = csvhost::max_ops_per_batch();
    std::vector<SVCall> found;
    std::vector<std::pair<uint64_t, uint32_t>> seq;                 // insertion order of found[i]: (record, op)
    std::vector<uint32_t> start, end, read_idx, op_idx, query_pos;
    std::vector<uint8_t> kind;
    uint64_t read_base = 0;
    bool failed = false;
    const double default_lh = 0.0;
    csvhost::PackedReads reads;
    // one slice of consecutive records: signatures only depend on the record itself, so slices need no halo
    auto flush = [&]() {
        if (reads.size() == 0 || map_size == 0 || failed) { read_base += reads.size(); reads.clear(); return; }
        const csv_region reg = {reads.tid[0], 0u, map_size, map_size};
        const csv_reads view = reads.view();
        uint64_t n = 0, cap = std::max<uint64_t>(start.size(), 1u << 16);
        if (!ctx) ctx = csvhost::thread_context();
        csvhost::StatTimer st(csvhost::STAT_CIGAR_GPU, view.n_reads);
        for (;;) {
            start.resize(cap); end.resize(cap); read_idx.resize(cap); op_idx.resize(cap); query_pos.resize(cap); kind.resize(cap);
            csv_sigs out = {start.data(), end.data(), kind.data(), read_idx.data(), op_idx.data(), query_pos.data()};
            const int rc = csv_cigar_scan(ctx, &view, &reg, 50, (uint8_t)this->min_mapq, &out, cap, &n);
            if (rc == CSV_OK) break;
            if (rc == CSV_ERR_CAPACITY && n > cap) { cap = n; continue; }
            printError(std::string("ERROR: GPU CIGAR scan failed: ") + csv_last_error());
            failed = true; n = 0;
            break;
        }
        for (uint64_t i = 0; i < n; i++) {
            seq.emplace_back(read_base + read_idx[i], op_idx[i]);
            SVEvidenceFlags aln_type;
            if (kind[i] == 1) {
                aln_type.set(static_cast<size_t>(SVDataType::CIGARDEL));
                found.emplace_back(start[i], end[i], SVType::DEL, getSVTypeSymbol(SVType::DEL), aln_type, Genotype::UNKNOWN, default_lh, 0, 0, 0);
                continue;
            }
            aln_type.set(static_cast<size_t>(kind[i] == 0 ? SVDataType::CIGARINS : SVDataType::CIGARCLIP));
            std::string alt_allele = "<INS>";
            const uint32_t op_len = end[i] - start[i] + 1;
            if (op_len <= 50) {                                     // literal sequence for a 50-base event (sv_caller.cpp:587-591)
                const auto it = reads.seq4.find(read_idx[i]);
                if (it != reads.seq4.end()) {
                    alt_allele.assign(op_len, ' ');
                    for (uint32_t j = 0; j < op_len; j++) alt_allele[j] = csvhost::base_at(it->second, query_pos[i] + j);
                }
            }
            found.emplace_back(start[i], end[i], SVType::INS, alt_allele, aln_type, Genotype::UNKNOWN, default_lh, 0, 0, 0);
        }
        read_base += reads.size();
        reads.clear();
    };
    // the depth pass has usually packed this contig already (packed_reads.h): no second decode
    const int whole_tid = sam_hdr_name2tid(bamHdr, region.c_str());
    std::unique_ptr<csvhost::PackedReads> cached = whole_tid >= 0 ? csvhost::cache_take(csvhost::file_name(fp_in), whole_tid) : nullptr;
    if (cached && cached->ops() + cached->size() <= max_ops) {
        reads = std::move(*cached);
        csvhost::stat_add(csvhost::STAT_CACHE_HIT, 0.0, reads.size());
    } else {
        while (readNextAlignment(fp_in, itr, bam1) >= 0) {
            if (reads.ops() + reads.size() + bam1->core.n_cigar + 1 > max_ops && reads.size() > 0) flush();   // ops + records: a batch counts both
            reads.append(bam1, true);
        }
    }
    hts_itr_destroy(itr);
    bam_destroy1(bam1);
    const bool one_slice = read_base == 0;
    flush();
    if (failed) return;
    if (!one_slice) {
        // several slices: each came back in vector order; the order of the whole is (start, end) ascending with equal
        // keys in reverse insertion order (sv_object.cpp:17-33)
        std::vector<size_t> ord(found.size());
        for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
            if (found[a].start != found[b].start) return found[a].start < found[b].start;
            if (found[a].end != found[b].end) return found[a].end < found[b].end;
            return seq[a] > seq[b];
        });
        std::vector<SVCall> sorted; sorted.reserve(found.size());
        std::vector<std::pair<uint64_t, uint32_t>> sseq; sseq.reserve(found.size());
        for (size_t i : ord) { sorted.push_back(std::move(found[i])); sseq.push_back(seq[i]); }
        found.swap(sorted); seq.swap(sseq);
    }
    if (sv_calls.empty()) { sv_calls.swap(found); return; }
    // a non-empty target vector: replay addSVCall in the reference's insertion order (record, op)
    std::vector<size_t> order(found.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return seq[a] < seq[b]; });
    for (size_t i : order) addSVCall(sv_calls, found[i]);
}

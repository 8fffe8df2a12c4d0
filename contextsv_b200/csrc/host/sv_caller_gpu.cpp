// sv_caller_gpu.cpp -- drop-in definitions of the SVCaller members on the alignment-scan path:
//
//   SVCaller::findCIGARSVs      include/sv_caller.h:86, src/sv_caller.cpp:506-537 (and through it processCIGARRecord
//       539-661 and the addSVCall order).  The signatures of the contig were produced by the GPU scan of the depth pass
//       (one decode, one upload: scan_results.h) in the order of the reference's vector after all addSVCall()
//       insertions; here the SVCall objects -- including the 50-base literal ALT allele -- are materialised.  A
//       region the depth pass did not see (another file, a sub-region, another min_mapq) is decoded and scanned here.
//   SVCaller::getReadDepth      include/sv_caller.h:97, src/sv_caller.cpp:1332-1344: read from the device-resident map.
//   SVCaller::saveToVCF         include/sv_caller.h:93, src/sv_caller.cpp:1067-1330: asks for the depth at every call's
//       position in one launch per chromosome, then runs the reference's own body.
//   SVCaller::runSplitReadCopyNumberPredictions   include/sv_caller.h:91, src/sv_caller.cpp:983-1065: the same for the
//       log2 windows of the split-read candidates.
#include "sv_caller.h"

#include <htslib/sam.h>

#include <algorithm>
#include <memory>
#include <stdexcept>

#include "contextsv_b200.h"
#include "gpu_context.h"
#include "packed_reads.h"
#include "scan_results.h"

namespace {

// one signature -> the SVCall processCIGARRecord builds for it (sv_caller.cpp:566-646)
template <class BaseAt>
SVCall make_call(uint32_t start, uint32_t end, uint8_t kind, uint32_t query_pos, bool have_seq, BaseAt base_at)
{
    SVEvidenceFlags aln_type;
    if (kind == 1) {
        aln_type.set(static_cast<size_t>(SVDataType::CIGARDEL));
        return SVCall(start, end, SVType::DEL, getSVTypeSymbol(SVType::DEL), aln_type, Genotype::UNKNOWN, 0.0, 0, 0, 0);
    }
    aln_type.set(static_cast<size_t>(kind == 0 ? SVDataType::CIGARINS : SVDataType::CIGARCLIP));
    std::string alt_allele = "<INS>";
    const uint32_t op_len = end - start + 1;
    if (op_len <= 50 && have_seq) {                                  // literal sequence for a 50-base event (sv_caller.cpp:587-591)
        alt_allele.assign(op_len, ' ');
        for (uint32_t j = 0; j < op_len; j++) alt_allele[j] = base_at(query_pos + j);
    }
    return SVCall(start, end, SVType::INS, alt_allele, aln_type, Genotype::UNKNOWN, 0.0, 0, 0, 0);
}

}  // namespace

void SVCaller::findCIGARSVs(samFile* fp_in, hts_idx_t* idx, bam_hdr_t* bamHdr, const std::string& region, std::vector<SVCall>& sv_calls,
                            const std::vector<uint32_t>& pos_depth_map)
{
    csvhost::StatTimer st_all(csvhost::STAT_CIGAR, 1);
    const uint32_t map_size = (uint32_t)pos_depth_map.size();       // only the size of the depth map is consulted (sv_caller.cpp:602)
    std::vector<SVCall> found;
    std::vector<std::pair<uint64_t, uint32_t>> seq;                 // insertion order of found[i]: (record, op)

    // ---- the depth pass has scanned this contig already: no second decode, no second upload
    const int whole_tid = sam_hdr_name2tid(bamHdr, region.c_str());
    std::shared_ptr<csvhost::ContigResults> parked = whole_tid >= 0 ? csvhost::results_for_contig(csvhost::file_name(fp_in), whole_tid) : nullptr;
    if (parked && parked->have_sigs && parked->sig_min_mapq == (uint8_t)this->min_mapq && parked->map_size == map_size) {
        const csvhost::SigColumns& s = parked->sigs;
        csvhost::stat_add(csvhost::STAT_CACHE_HIT, 0.0, s.size());
        found.reserve(s.size()); seq.reserve(s.size());
        for (size_t i = 0; i < s.size(); i++) {
            const auto it = parked->seq4.find(s.serial[i]);
            const bool have = it != parked->seq4.end();
            found.push_back(make_call(s.start[i], s.end[i], s.kind[i], s.query_pos[i], have, [&](uint32_t q) { return csvhost::base_at(it->second, q); }));
            seq.emplace_back(s.serial[i], s.op_idx[i]);
        }
    } else {
        // ---- a region the depth pass did not see: decode it here, slice by slice
        csvhost::warm_up_async();
        bam1_t* bam1 = bam_init1();
        if (!bam1) { printError("ERROR: failed to initialize BAM record"); return; }
        hts_itr_t* itr = sam_itr_querys(idx, bamHdr, region.c_str());
        if (!itr) { bam_destroy1(bam1); printError("ERROR: failed to query region " + region); return; }
        csv_ctx* ctx = nullptr;                                     // created at the first flush
        const uint64_t max_ops = csvhost::max_ops_per_batch();
        std::vector<uint32_t> start, end, read_idx, op_idx, query_pos;
        std::vector<uint8_t> kind;
        uint64_t read_base = 0;
        csvhost::PackedReads reads;
        // one slice of consecutive records: signatures only depend on the record itself, so slices need no halo
        auto flush = [&]() {
            if (reads.size() == 0 || map_size == 0) { read_base += reads.size(); reads.clear(); return; }
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
                // fatal, like a failed DBSCAN fit: a chromosome without its CIGAR calls is not a result (caught per chromosome, sv_caller.cpp:838-842)
                throw std::runtime_error(std::string("contextsv_b200 CIGAR scan: ") + csv_last_error());
            }
            for (uint64_t i = 0; i < n; i++) {
                seq.emplace_back(read_base + read_idx[i], op_idx[i]);
                const auto it = reads.seq4.find(read_idx[i]);
                const bool have = it != reads.seq4.end();
                found.push_back(make_call(start[i], end[i], kind[i], query_pos[i], have, [&](uint32_t q) { return csvhost::base_at(it->second, q); }));
            }
            read_base += reads.size();
            reads.clear();
        };
        try {
            while (readNextAlignment(fp_in, itr, bam1) >= 0) {
                if (reads.ops() + reads.size() + bam1->core.n_cigar + 1 > max_ops && reads.size() > 0) flush();   // ops + records: a batch counts both
                reads.append(bam1, true);
            }
            const bool one_slice = read_base == 0;
            flush();
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
        } catch (...) {
            hts_itr_destroy(itr);
            bam_destroy1(bam1);
            throw;
        }
        hts_itr_destroy(itr);
        bam_destroy1(bam1);
    }
    if (sv_calls.empty()) { sv_calls.swap(found); return; }
    // a non-empty target vector: replay addSVCall in the reference's insertion order (record, op)
    std::vector<size_t> order(found.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return seq[a] < seq[b]; });
    for (size_t i : order) addSVCall(sv_calls, found[i]);
}

// SVCaller::getReadDepth (sv_caller.cpp:1332-1344): map.at(start), 0 (and the reference's warning) beyond the map.
int SVCaller::getReadDepth(const std::vector<uint32_t>& pos_depth_map, uint32_t start) const
{
    int read_depth = 0;
    if (start >= pos_depth_map.size()) {
        printError("Warning: Read depth for position " + std::to_string(start) + " is out of range of size " + std::to_string(pos_depth_map.size()));
        return read_depth;
    }
    uint32_t d = 0;
    if (csvhost::host_depth_requested()) d = pos_depth_map[start];
    else if (!csvhost::prefetched_depth(&pos_depth_map, start, &d)) {
        const std::shared_ptr<csvhost::ContigResults> res = csvhost::results_for_vector(&pos_depth_map);
        if (!res) d = pos_depth_map[start];                         // a map that did not come from the depth pass
        else if (!csvhost::device_depth_at(*res, &start, 1, &d)) throw std::runtime_error(std::string("contextsv_b200 getReadDepth: ") + csv_last_error());
    }
    read_depth += d;
    return read_depth;
}

// The reference's own bodies under a second name (oracle/Makefile: an alias symbol added to the unmodified object; in a
// source-level integration, the functions renamed).  Itanium C++ ABI: `this` first, references as pointers.
extern "C" void csv_ref_saveToVCF(const SVCaller* self, const std::unordered_map<std::string, std::vector<SVCall>>& sv_calls, const InputData& input_data,
                                  const ReferenceGenome& ref_genome, const std::unordered_map<std::string, std::vector<uint32_t>>& chr_pos_depth_map);
extern "C" void csv_ref_runSplitReadCopyNumberPredictions(SVCaller* self, const std::string& chr, std::vector<SVCall>& split_sv_calls, const CNVCaller& cnv_caller,
                                                          const CHMM& hmm, double mean_chr_cov, const std::vector<uint32_t>& pos_depth_map, const InputData& input_data);

void SVCaller::saveToVCF(const std::unordered_map<std::string, std::vector<SVCall>>& sv_calls, const InputData& input_data, const ReferenceGenome& ref_genome,
                         const std::unordered_map<std::string, std::vector<uint32_t>>& chr_pos_depth_map) const
{
    // every position the loop at sv_caller.cpp:1185-1306 can ask for: the call's start, or the base before it (the
    // record of a deletion / insertion is anchored there, :1253-1275)
    if (!csvhost::host_depth_requested()) {
        for (const auto& pair : sv_calls) {
            const auto m = chr_pos_depth_map.find(pair.first);
            if (m == chr_pos_depth_map.end() || !csvhost::results_for_vector(&m->second)) continue;
            std::vector<uint32_t> pos;
            pos.reserve(2 * pair.second.size());
            for (const SVCall& sv : pair.second) {
                if (sv.sv_type == SVType::UNKNOWN || sv.sv_type == SVType::NEUTRAL) continue;
                if (sv.start < m->second.size()) pos.push_back(sv.start);
                const uint32_t before = (uint32_t)std::max(1, static_cast<int>(sv.start) - 1);
                if (before < m->second.size()) pos.push_back(before);
            }
            std::sort(pos.begin(), pos.end());
            pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
            csvhost::prefetch_depth_at(&m->second, pos);
        }
    }
    csv_ref_saveToVCF(this, sv_calls, input_data, ref_genome, chr_pos_depth_map);
    for (const auto& pair : chr_pos_depth_map) csvhost::drop_prefetch(&pair.second);
}

void SVCaller::runSplitReadCopyNumberPredictions(const std::string& chr, std::vector<SVCall>& split_sv_calls, const CNVCaller& cnv_caller, const CHMM& hmm,
                                                 double mean_chr_cov, const std::vector<uint32_t>& pos_depth_map, const InputData& input_data)
{
    // runCopyNumberPrediction queries [start, end] of every candidate (cnv_caller.cpp:200), and the flanks as well when
    // the CNV data is saved (:178-196): all windows in one launch
    if (!csvhost::host_depth_requested() && csvhost::results_for_vector(&pos_depth_map)) {
        std::vector<uint32_t> start, end;
        const int last = static_cast<int>(pos_depth_map.size()) - 1;
        for (const SVCall& sv : split_sv_calls) {
            if (sv.start > sv.end) continue;
            start.push_back(sv.start); end.push_back(sv.end);
            if (input_data.getSaveCNVData()) {
                const int half = (static_cast<int>(sv.end) - static_cast<int>(sv.start)) / 2;
                const int b0 = std::max(1, static_cast<int>(sv.start) - half), b1 = std::max(1, static_cast<int>(sv.start) - 1);
                if (b0 < b1) { start.push_back((uint32_t)b0); end.push_back((uint32_t)b1); }
                const int a0 = std::min(last, static_cast<int>(sv.end) + 1), a1 = std::min(last, static_cast<int>(sv.end) + half);
                if (a0 < a1) { start.push_back((uint32_t)a0); end.push_back((uint32_t)a1); }
            }
        }
        csvhost::prefetch_windows(&pos_depth_map, start, end, input_data.getSampleSize());
    }
    csv_ref_runSplitReadCopyNumberPredictions(this, chr, split_sv_calls, cnv_caller, hmm, mean_chr_cov, pos_depth_map, input_data);
    csvhost::drop_prefetch(&pos_depth_map);
}

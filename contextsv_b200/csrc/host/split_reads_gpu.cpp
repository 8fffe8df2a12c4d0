// split_reads_gpu.cpp -- drop-in definition of SVCaller::findSplitSVSignatures (include/sv_caller.h:77,
// src/sv_caller.cpp:68-504) on top of what the single decode left behind (scan_results.h).
//
// The reference opens and inflates the BAM a third time here.  The drop-in takes the records from the depth pass: flags,
// positions and query names as packed on the host, bam_endpos and getAlignmentReadPositions (sv_caller.cpp:663-690) as
// computed on the device from the resident CIGAR words (csv_record_summary).  A file the depth pass did not see whole
// is decoded here instead, through the same htslib calls as the reference.
//
// What follows the record loop is re-stated, not re-ordered: the reference's results depend on the iteration order of its
// std::unordered_map / std::unordered_set containers (which primary alignment seeds an overlap group, in which order a
// group's members enter the DBSCAN1D input -- cluster ids, and with them the "largest" cluster on ties, are
// order-dependent), so the same containers are filled by the same sequence of operations and the same interval tree
// (SVCaller::insert / findOverlaps, the reference's own members) answers the overlap queries.  The part that is
// data-parallel -- six DBSCAN1D(100, 5) fits per overlap group (sv_caller.cpp:270-372) -- is collected for a whole
// chromosome and runs as ONE csv_dbscan1d_seg call (the fits are independent); the reference's DBSCAN1D::fit would be
// one launch per fit.
#include "sv_caller.h"

#include <htslib/sam.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <stdexcept>
#include <unordered_set>

#include "contextsv_b200.h"
#include "gpu_context.h"
#include "packed_reads.h"
#include "scan_results.h"

namespace {

// the six point sets of one overlap group, in the order the reference fits them
enum { FIT_PRIMARY_START = 0, FIT_PRIMARY_END, FIT_SUPP_START, FIT_SUPP_END, FIT_READ_DIST, FIT_REF_DIST, FIT_COUNT };

struct GroupFits {
    std::vector<int> pts[FIT_COUNT];
    bool inversion = false;
};

// DBSCAN1D::getLargestCluster (dbscan1d.cpp:72-90) on the labels of one fit
std::vector<int> largest_cluster(const int32_t* pts, const int32_t* labels, size_t n)
{
    std::vector<int> out(n);
    out.resize(csv_largest_cluster(pts, labels, n, out.data()));
    return out;
}

}  // namespace

void SVCaller::findSplitSVSignatures(std::unordered_map<std::string, std::vector<SVCall>>& sv_calls, const InputData& input_data)
{
    csvhost::StatTimer st_all(csvhost::STAT_SPLIT, 1);
    const std::string bam_filepath = input_data.getLongReadBam();
    samFile* fp_in = sam_open(bam_filepath.c_str(), "r");
    if (!fp_in) { printError("ERROR: failed to open " + bam_filepath); return; }
    const int thread_count = input_data.getThreadCount();
    hts_set_threads(fp_in, thread_count);
    printMessage("Using " + std::to_string(thread_count) + " threads for split read analysis");
    bam_hdr_t* bamHdr = sam_hdr_read(fp_in);
    if (!bamHdr) { sam_close(fp_in); printError("ERROR: failed to read header from " + bam_filepath); return; }

    // ---- alignment tables, filled exactly as sv_caller.cpp:101-170 fills them
    std::unordered_map<int, std::unordered_map<std::string, PrimaryAlignment>> primary_map;     // tid -> qname -> primary alignment
    std::unordered_map<std::string, std::vector<SuppAlignment>> supp_map;                      // qname -> supplementary alignments
    std::unordered_set<int> alignment_tids;
    std::unordered_set<std::string> supp_qnames;
    uint32_t num_alignments = 0;
    auto feed = [&](int tid, int32_t pos, int32_t endpos, int32_t query_start, int32_t query_end, uint16_t flag, uint8_t mapq, const std::string& qname) {
        if (flag & BAM_FSECONDARY || flag & BAM_FUNMAP || flag & BAM_FDUP || flag & BAM_FQCFAIL || mapq < this->min_mapq) return;
        if (!(flag & BAM_FSUPPLEMENTARY)) {
            primary_map[tid][qname] = PrimaryAlignment{pos + 1, endpos, query_start, query_end, !(flag & BAM_FREVERSE), 0};
            alignment_tids.insert(tid);
        } else {
            supp_map[qname].push_back(SuppAlignment{tid, pos + 1, endpos, query_start, query_end, !(flag & BAM_FREVERSE)});
            alignment_tids.insert(tid);
            supp_qnames.insert(qname);
        }
        num_alignments++;
        if (num_alignments % 1000000 == 0) printMessage("Processed " + std::to_string(num_alignments) + " alignments");
    };

    // ---- the records: parked by the depth pass for every contig the reference's iterator would walk, else decoded here
    std::vector<int> tids;
    if (input_data.isSingleChr()) {
        const int t = sam_hdr_name2tid(bamHdr, input_data.getChromosome().c_str());
        if (t < 0) { bam_hdr_destroy(bamHdr); sam_close(fp_in); printError("ERROR: failed to create iterator for " + input_data.getChromosome()); return; }
        tids.push_back(t);
    } else {
        for (int t = 0; t < bamHdr->n_targets; t++) tids.push_back(t);
    }
    std::vector<std::shared_ptr<csvhost::ContigResults>> parked;
    for (int t : tids) {
        std::shared_ptr<csvhost::ContigResults> r = csvhost::results_for_contig(bam_filepath.c_str(), t);
        if (!r || !r->have_split) { parked.clear(); break; }
        parked.push_back(std::move(r));
    }
    printMessage("Processing alignments from " + bam_filepath);
    if (!parked.empty()) {
        std::string qname;
        for (const auto& r : parked) {
            const csvhost::SplitRecords& s = r->split;
            for (size_t i = 0; i < s.size(); i++) {
                qname.assign(s.names.data() + s.name_off[i], s.names.data() + s.name_off[i + 1]);
                feed(r->tid, s.pos[i], s.endpos[i], s.query_start[i], s.query_end[i], s.flag[i], s.mapq[i], qname);
            }
        }
    } else {
        hts_idx_t* idx = sam_index_load(fp_in, bam_filepath.c_str());
        if (!idx) { bam_hdr_destroy(bamHdr); sam_close(fp_in); printError("ERROR: failed to load index for " + bam_filepath); return; }
        bam1_t* bam1 = bam_init1();
        hts_itr_t* itr = input_data.isSingleChr() ? sam_itr_querys(idx, bamHdr, input_data.getChromosome().c_str()) : sam_itr_queryi(idx, HTS_IDX_START, 0, 0);
        if (!bam1 || !itr) {
            if (bam1) bam_destroy1(bam1);
            hts_idx_destroy(idx); bam_hdr_destroy(bamHdr); sam_close(fp_in);
            printError("ERROR: failed to create iterator for " + bam_filepath);
            return;
        }
        while (readNextAlignment(fp_in, itr, bam1) >= 0) {
            if (bam1->core.flag & (BAM_FSECONDARY | BAM_FUNMAP | BAM_FDUP | BAM_FQCFAIL) || bam1->core.qual < this->min_mapq) continue;
            const std::pair<int, int> qpos = getAlignmentReadPositions(bam1);
            feed(bam1->core.tid, (int32_t)bam1->core.pos, (int32_t)bam_endpos(bam1), qpos.first, qpos.second, bam1->core.flag, bam1->core.qual, bam_get_qname(bam1));
        }
        hts_itr_destroy(itr);
        bam_destroy1(bam1);
        hts_idx_destroy(idx);
    }
    sam_close(fp_in);

    // ---- primary alignments without a supplementary one are dropped (sv_caller.cpp:180-200)
    std::unordered_map<int, std::unordered_set<std::string>> to_remove;
    for (auto& chr_primary : primary_map)
        for (const auto& entry : chr_primary.second)
            if (supp_qnames.find(entry.first) == supp_qnames.end()) to_remove[chr_primary.first].insert(entry.first);
    int total_removed = 0;
    for (auto& chr_primary : primary_map) {
        total_removed += to_remove[chr_primary.first].size();
        for (const auto& qname : to_remove[chr_primary.first]) chr_primary.second.erase(qname);
    }
    printMessage("Removed " + std::to_string(total_removed) + " primary alignments without supplementary alignments");

    const int min_length = 2000, max_length = 1000000;
    csv_ctx* ctx = nullptr;
    for (const auto& chr_primary : primary_map) {
        const int primary_tid = chr_primary.first;
        const std::string chr_name = bamHdr->target_name[primary_tid];
        printMessage("Processing chromosome " + chr_name + " with " + std::to_string(chr_primary.second.size()) + " primary alignments");
        std::vector<SVCall> chr_sv_calls;
        chr_sv_calls.reserve(1000);
        const std::unordered_map<std::string, PrimaryAlignment>& chr_primary_map = chr_primary.second;

        // overlap groups, seeded in the map's iteration order (sv_caller.cpp:212-235)
        std::unique_ptr<IntervalNode> root = nullptr;
        for (const auto& entry : chr_primary_map) insert(root, entry.second, entry.first);
        std::vector<std::vector<std::string>> primary_clusters;
        std::set<std::string> processed;
        for (const auto& entry : chr_primary_map) {
            if (processed.find(entry.first) != processed.end()) continue;
            std::vector<std::string> overlap_group;
            findOverlaps(root, entry.second, overlap_group);
            for (const std::string& q : overlap_group) processed.insert(q);
            if (overlap_group.size() > 1) primary_clusters.push_back(std::move(overlap_group));
        }

        // the point sets of every group (sv_caller.cpp:244-352) ...
        std::vector<GroupFits> groups(primary_clusters.size());
        size_t n_points = 0;
        for (size_t g = 0; g < primary_clusters.size(); g++) {
            const std::vector<std::string>& primary_cluster = primary_clusters[g];
            GroupFits& f = groups[g];
            int num_supp_opposite_strand = 0;
            for (const std::string& qname : primary_cluster) {
                const PrimaryAlignment& primary_aln = chr_primary_map.at(qname);
                const std::vector<SuppAlignment>& supp_alns = supp_map[qname];
                bool has_opposite_strand = false;
                f.pts[FIT_PRIMARY_START].push_back(primary_aln.start);
                f.pts[FIT_PRIMARY_END].push_back(primary_aln.end);
                for (const SuppAlignment& supp_aln : supp_alns) {
                    if (supp_aln.tid != primary_tid) continue;                   // other chromosome: translocations are not called
                    if (supp_aln.strand != primary_aln.strand) { has_opposite_strand = true; }
                    f.pts[FIT_SUPP_START].push_back(supp_aln.start);
                    f.pts[FIT_SUPP_END].push_back(supp_aln.end);
                    if (supp_aln.strand == primary_aln.strand) {
                        const bool primary_5p = primary_aln.start < supp_aln.start;
                        int read_distance = std::max(0, std::max(supp_aln.query_start, primary_aln.query_start) - std::min(supp_aln.query_end, primary_aln.query_end));
                        const int ref_distance = std::max(0, std::max(supp_aln.start, primary_aln.start) - std::min(supp_aln.end, primary_aln.end));
                        if (!primary_5p) read_distance = -read_distance;         // negative: the primary alignment is not 5'-most
                        f.pts[FIT_READ_DIST].push_back(read_distance);
                        f.pts[FIT_REF_DIST].push_back(ref_distance);
                    }
                }
                if (has_opposite_strand) num_supp_opposite_strand++;
            }
            f.inversion = static_cast<double>(num_supp_opposite_strand) / static_cast<double>((int)primary_cluster.size()) > 0.5;
            for (int k = 0; k < FIT_COUNT; k++) n_points += f.pts[k].size();
        }

        // ... fitted together: segment 6 g + k is fit k of group g; DBSCAN1D(100, 5) as at sv_caller.cpp:270
        std::vector<int32_t> pts, labels(n_points);
        std::vector<uint32_t> seg;
        std::vector<size_t> fit_off(groups.size() * FIT_COUNT + 1, 0);
        pts.reserve(n_points); seg.reserve(n_points);
        for (size_t g = 0; g < groups.size(); g++)
            for (int k = 0; k < FIT_COUNT; k++) {
                const std::vector<int>& p = groups[g].pts[k];
                fit_off[g * FIT_COUNT + k] = pts.size();
                pts.insert(pts.end(), p.begin(), p.end());
                seg.insert(seg.end(), p.size(), (uint32_t)(g * FIT_COUNT + k));
            }
        fit_off[groups.size() * FIT_COUNT] = pts.size();
        if (n_points) {
            if (!ctx) ctx = csvhost::thread_context();
            csvhost::StatTimer st(csvhost::STAT_DBSCAN1D, n_points);
            if (csv_dbscan1d_seg(ctx, pts.data(), seg.data(), n_points, (uint32_t)(groups.size() * FIT_COUNT), 100.0, 5, labels.data(), nullptr) != CSV_OK)
                throw std::runtime_error(std::string("contextsv_b200 split-read DBSCAN1D: ") + csv_last_error());
        }
        auto cluster_of = [&](size_t g, int k) {
            const size_t o = fit_off[g * FIT_COUNT + k], n = fit_off[g * FIT_COUNT + k + 1] - o;
            return n ? largest_cluster(pts.data() + o, labels.data() + o, n) : std::vector<int>();
        };

        // ... and turned into candidates group by group, in the reference's order (sv_caller.cpp:281-483)
        for (size_t g = 0; g < groups.size(); g++) {
            std::vector<int> primary_start_cluster = cluster_of(g, FIT_PRIMARY_START), primary_end_cluster = cluster_of(g, FIT_PRIMARY_END);
            if (primary_start_cluster.empty() && primary_end_cluster.empty()) continue;
            std::vector<int> supp_start_cluster = cluster_of(g, FIT_SUPP_START), supp_end_cluster = cluster_of(g, FIT_SUPP_END);
            std::vector<int> read_distance_cluster = cluster_of(g, FIT_READ_DIST), ref_distance_cluster = cluster_of(g, FIT_REF_DIST);
            if (supp_start_cluster.empty() && supp_end_cluster.empty() && read_distance_cluster.empty() && ref_distance_cluster.empty()) continue;

            // medians of the largest clusters are the coordinates
            std::vector<int> primary_positions, supp_positions;
            int primary_cluster_size = 0, supp_cluster_size = 0;
            bool primary_end = false, supp_end = false;
            if (!primary_start_cluster.empty()) {
                std::sort(primary_start_cluster.begin(), primary_start_cluster.end());
                primary_positions.push_back(primary_start_cluster[primary_start_cluster.size() / 2]);
                primary_cluster_size = primary_start_cluster.size();
            }
            if (!primary_end_cluster.empty()) {
                std::sort(primary_end_cluster.begin(), primary_end_cluster.end());
                primary_positions.push_back(primary_end_cluster[primary_end_cluster.size() / 2]);
                primary_cluster_size = std::max(primary_cluster_size, (int)primary_end_cluster.size());
                primary_end = true;
            }
            if (!supp_start_cluster.empty()) {
                std::sort(supp_start_cluster.begin(), supp_start_cluster.end());
                supp_positions.push_back(supp_start_cluster[supp_start_cluster.size() / 2]);
                supp_cluster_size = supp_start_cluster.size();
            }
            if (!supp_end_cluster.empty()) {
                std::sort(supp_end_cluster.begin(), supp_end_cluster.end());
                supp_positions.push_back(supp_end_cluster[supp_end_cluster.size() / 2]);
                supp_cluster_size = std::max(supp_cluster_size, (int)supp_end_cluster.size());
                supp_end = true;
            }

            // split insertion / unknown call from the distance between the two alignments on the read and on the reference
            if (!read_distance_cluster.empty() && !ref_distance_cluster.empty()) {
                std::sort(read_distance_cluster.begin(), read_distance_cluster.end());
                int read_distance = read_distance_cluster[read_distance_cluster.size() / 2];
                const bool primary_5p_most = read_distance > 0;
                read_distance = std::abs(read_distance);
                std::sort(ref_distance_cluster.begin(), ref_distance_cluster.end());
                const int ref_distance = ref_distance_cluster[ref_distance_cluster.size() / 2];
                int sv_start = 0;
                bool split_candidate_sv = false;
                if (primary_5p_most && primary_end) {
                    std::sort(primary_positions.begin(), primary_positions.end());
                    sv_start = primary_positions.back();
                    split_candidate_sv = true;
                } else if (!primary_5p_most && supp_end) {
                    std::sort(supp_positions.begin(), supp_positions.end());
                    sv_start = supp_positions.back();
                    split_candidate_sv = true;
                }
                if (split_candidate_sv) {
                    SVEvidenceFlags aln_type;
                    aln_type.set(static_cast<size_t>(SVDataType::SPLITDIST1));
                    const int aln_offset = ref_distance - read_distance;
                    if (read_distance > ref_distance && read_distance >= min_length && read_distance <= max_length) {
                        SVCall sv_candidate(sv_start, sv_start + (read_distance - 1), SVType::INS, getSVTypeSymbol(SVType::INS), aln_type, Genotype::UNKNOWN, 0.0, 0, aln_offset, primary_cluster_size);
                        addSVCall(chr_sv_calls, sv_candidate);
                    } else if (ref_distance > read_distance && ref_distance >= min_length && ref_distance <= max_length) {
                        SVCall sv_candidate(sv_start, sv_start + (ref_distance - 1), SVType::UNKNOWN, getSVTypeSymbol(SVType::UNKNOWN), aln_type, Genotype::UNKNOWN, 0.0, 0, aln_offset, primary_cluster_size);
                        addSVCall(chr_sv_calls, sv_candidate);
                    }
                }
            }

            // one candidate per (primary, supplementary) coordinate pair for the copy-number pass
            const int cluster_size = std::max(primary_cluster_size, supp_cluster_size);
            const SVType sv_type = groups[g].inversion ? SVType::INV : SVType::UNKNOWN;
            const std::string alt = (sv_type == SVType::INV) ? "<INV>" : ".";
            for (int primary_pos : primary_positions) {
                for (int supp_pos : supp_positions) {
                    const int sv_start = std::min(primary_pos, supp_pos), sv_end = std::max(primary_pos, supp_pos) - 1;
                    const int sv_length = sv_end - sv_start + 1;
                    if (sv_length < min_length || sv_length > max_length) continue;
                    SVEvidenceFlags aln_type;
                    aln_type.set(static_cast<size_t>(SVDataType::SPLIT));
                    SVCall sv_candidate(sv_start, sv_end, sv_type, alt, aln_type, Genotype::UNKNOWN, 0.0, 0, 0, cluster_size);
                    addSVCall(chr_sv_calls, sv_candidate);
                }
            }
        }

        std::sort(chr_sv_calls.begin(), chr_sv_calls.end(), [](const SVCall& a, const SVCall& b) { return a.start < b.start || (a.start == b.start && a.end < b.end); });
        mergeDuplicateSVs(chr_sv_calls);
        if (const char* dump = std::getenv("CONTEXTSV_B200_DUMP_SPLIT")) {          // test hook: the candidates, one line each
            if (FILE* f = std::fopen(dump, "a")) {
                for (const SVCall& c : chr_sv_calls)
                    std::fprintf(f, "%s\t%u\t%u\t%d\t%s\t%lu\t%d\t%d\n", chr_name.c_str(), c.start, c.end, (int)c.sv_type, c.alt_allele.c_str(), c.aln_type.to_ulong(), c.aln_offset, c.cluster_size);
                std::fclose(f);
            }
        }
        sv_calls[chr_name] = std::move(chr_sv_calls);
        printMessage(chr_name + ": Found " + std::to_string(sv_calls[chr_name].size()) + " SV candidates");
    }
    bam_hdr_destroy(bamHdr);
}

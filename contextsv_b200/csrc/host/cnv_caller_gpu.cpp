// cnv_caller_gpu.cpp -- drop-in definitions of the CNVCaller members on the alignment-scan path:
//
//   CNVCaller::calculateMeanChromosomeCoverage   include/cnv_caller.h:104, src/cnv_caller.cpp:415-556
//       same signature, same messages, same containers; the per-base loop and the two reductions run on the GPU.
//       This is also the ONE decode of the BAM (the reference decodes it three times): besides the depth it leaves
//       the CIGAR signatures of every contig and the per-record summaries of the split-read pass behind
//       (scan_results.h), and the depth map stays in HBM -- the caller's vectors keep their size (all the reference's
//       CIGAR pass looks at, sv_caller.cpp:602) and are only filled when CONTEXTSV_HOST_DEPTH=1.
//   CNVCaller::querySNPRegion                    include/cnv_caller.h:57, src/cnv_caller.cpp:53-160
//       the log2 windows come from csv_window_sums on the device-resident map; everything else as the reference.
//   CNVCaller::runCIGARCopyNumberPrediction      include/cnv_caller.h:102, src/cnv_caller.cpp:290-389
//       asks for the windows of ALL its candidates in one launch, then runs the reference's own body.
//
// Pipeline of the depth pass: the calling thread decodes and packs records into a slab (pinned once CUDA is up); a
// contig -- or, when it holds more CIGAR ops than a batch takes, a shard of it plus the halo records that reach
// into the next shard -- goes to one worker thread per device of CONTEXTSV_GPUS, which uploads, scans, fetches the
// small results and releases the batch's inputs.  Slabs cycle through a free list (devices + 2), so contig k+1 is
// decoded while contig k is uploaded and scanned, and several GPUs take contigs round-robin.
#include "cnv_caller.h"

#include <htslib/sam.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <stdexcept>
#include <thread>

#include "contextsv_b200.h"
#include "gpu_context.h"
#include "packed_reads.h"
#include "scan_results.h"

namespace {

using csvhost::ContigResults;
using csvhost::PackedReads;

constexpr uint8_t kScanMinMapq = 20;        // SVCaller::min_mapq (include/sv_caller.h:72); findCIGARSVs checks it before it trusts the parked signatures

struct ShardOut {                            // what one shard contributes to its contig
    csvhost::DepthShard depth;
    uint64_t sum = 0; uint32_t nz = 0;
    csvhost::SigColumns sigs;                // shard order = addSVCall order of the shard's own records
    std::unordered_map<uint64_t, std::vector<uint8_t>> seq4;
    csvhost::SplitRecords split;
};

struct ContigState {
    std::string name;
    int tid = -1;
    uint32_t size = 0;
    std::vector<uint32_t>* depth = nullptr;
    std::vector<std::unique_ptr<ShardOut>> shards;       // by shard number
    size_t done = 0;
    bool closed = false;                                  // the producer has cut its last shard
};

struct Job {
    std::unique_ptr<PackedReads> reads;
    csv_region reg{};
    size_t contig = 0, shard = 0;
    uint64_t first_own = 0;                               // serial of the first record this shard owns (records before it are halo)
};

class DepthPipeline {
public:
    DepthPipeline(std::vector<ContigState>& contigs, const std::string& bam_path)
        : contigs_(contigs), bam_path_(bam_path)
    {
        const size_t n_dev = csvhost::device_count();
        for (size_t i = 0; i < n_dev + 2; i++) free_.emplace_back(new PackedReads);
        for (size_t i = 0; i < n_dev; i++) workers_.emplace_back([this, i] { work(csvhost::device_at(i)); });
    }
    ~DepthPipeline() { finish(); }

    std::unique_ptr<PackedReads> take_slab()
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_free_.wait(lk, [&] { return !free_.empty() || failed_; });
        if (free_.empty()) return std::unique_ptr<PackedReads>(new PackedReads);
        std::unique_ptr<PackedReads> s = std::move(free_.back());
        free_.pop_back();
        return s;
    }
    void submit(Job&& j)
    {
        std::lock_guard<std::mutex> lk(m_);
        jobs_.push_back(std::move(j));
        cv_jobs_.notify_one();
    }
    void finish()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            if (stopping_) return;
            stopping_ = true;
            cv_jobs_.notify_all();
        }
        for (auto& t : workers_) t.join();
    }
    bool failed() { std::lock_guard<std::mutex> lk(m_); return failed_; }
    std::string error() { std::lock_guard<std::mutex> lk(m_); return error_; }
    // shards of contig c that are through (caller compares with what it submitted)
    size_t done(size_t c) { std::lock_guard<std::mutex> lk(m_); return contigs_[c].done; }

private:
    void work(csvhost::Device& dev)
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_jobs_.wait(lk, [&] { return !jobs_.empty() || stopping_; });
                if (jobs_.empty()) return;
                j = std::move(jobs_.front());
                jobs_.pop_front();
            }
            std::unique_ptr<ShardOut> out(new ShardOut);
            std::string err;
            try {
                if (!failed()) err = scan(dev, j, *out);
            } catch (const std::exception& e) { err = e.what(); }
            {
                std::lock_guard<std::mutex> lk(m_);
                if (!err.empty() && !failed_) { failed_ = true; error_ = err; }
                ContigState& c = contigs_[j.contig];
                if (c.shards.size() <= j.shard) c.shards.resize(j.shard + 1);
                c.shards[j.shard] = std::move(out);
                c.done++;
                j.reads->clear();
                free_.push_back(std::move(j.reads));
                cv_free_.notify_one();
            }
        }
    }

    // upload + scan + small results of one shard; returns an error message or ""
    std::string scan(csvhost::Device& dev, Job& j, ShardOut& out)
    {
        const PackedReads& reads = *j.reads;
        const csv_reads view = reads.view();
        ContigState& c = contigs_[j.contig];
        std::lock_guard<std::mutex> lk(dev.m);
        csv_ctx* ctx = csvhost::ensure_context(dev);
        csvhost::set_slab_allocator(csv_host_alloc, csv_host_free);      // CUDA is up: slabs that grow from now on are pinned
        csvhost::StatTimer st(csvhost::STAT_DEPTH_GPU, view.n_reads);
        csv_batch* b = nullptr;
        if (csv_batch_upload(ctx, &view, 1, &j.reg, &b) != CSV_OK) return csv_last_error();
        struct Guard { csv_ctx* c; csv_batch*& b; ~Guard() { if (b) csv_batch_free(c, b); } } guard{ctx, b};
        csv_scan_params p = {50, kScanMinMapq, 1, 1, 0};
        if (csv_scan_run(ctx, b, &p) != CSV_OK) return csv_last_error();
        if (csv_depth_stats(ctx, b, &out.sum, &out.nz) != CSV_OK) return csv_last_error();
        // signatures of the shard's own records (halo records belong to the shard before), in vector order
        uint64_t n = 0;
        int rc = csv_sigs_count(ctx, b, &n);
        if (rc == CSV_ERR_CAPACITY) {                                    // more than the batch reserved: the reference has no limit
            if (csv_batch_reserve_sigs(ctx, b, n) != CSV_OK || csv_scan_run(ctx, b, &p) != CSV_OK) return csv_last_error();
            rc = csv_sigs_count(ctx, b, &n);
        }
        if (rc != CSV_OK) return csv_last_error();
        std::vector<uint32_t> read_idx(n);
        out.sigs.start.resize(n); out.sigs.end.resize(n); out.sigs.op_idx.resize(n); out.sigs.query_pos.resize(n); out.sigs.kind.resize(n); out.sigs.serial.resize(n);
        csv_sigs cols = {out.sigs.start.data(), out.sigs.end.data(), out.sigs.kind.data(), read_idx.data(), out.sigs.op_idx.data(), out.sigs.query_pos.data()};
        if (csv_sigs_fetch(ctx, b, &cols, n, &n, nullptr) != CSV_OK) return csv_last_error();
        for (uint64_t i = 0; i < n; i++) {
            out.sigs.serial[i] = reads.serial[read_idx[i]];
            const uint32_t len = out.sigs.end[i] - out.sigs.start[i] + 1u;
            if (out.sigs.kind[i] != 1 && len <= 50) {                    // literal ALT allele (sv_caller.cpp:587-591): keep the record's bases
                const auto it = reads.seq4.find(read_idx[i]);
                if (it != reads.seq4.end()) out.seq4.emplace(out.sigs.serial[i], it->second);
            }
        }
        // what the split-read pass reads off every record (sv_caller.cpp:140-162): bam_endpos and the query interval
        if (!reads.name_off.empty() && view.n_reads) {
            std::vector<int32_t> endpos(view.n_reads), qs(view.n_reads), qe(view.n_reads);
            if (csv_record_summary(ctx, b, endpos.data(), qs.data(), qe.data()) != CSV_OK) return csv_last_error();
            for (uint32_t i = 0; i < view.n_reads; i++) {
                if (reads.serial[i] < j.first_own) continue;             // halo: reported by the shard that owns it
                if (reads.flag[i] & (BAM_FSECONDARY | BAM_FUNMAP | BAM_FDUP | BAM_FQCFAIL)) continue;     // sv_caller.cpp:136
                csvhost::SplitRecords& s = out.split;
                s.pos.push_back(reads.pos0[i]); s.endpos.push_back(endpos[i]); s.query_start.push_back(qs[i]); s.query_end.push_back(qe[i]);
                s.flag.push_back(reads.flag[i]); s.mapq.push_back(reads.mapq[i]);
                s.names.insert(s.names.end(), reads.names.data() + reads.name_off[i], reads.names.data() + reads.name_off[i + 1]);
                s.name_off.push_back(s.names.size());
            }
        }
        if (csvhost::host_depth_requested() && c.depth) {
            if (csv_depth_fetch(ctx, b, 0, c.depth->data() + j.reg.beg) != CSV_OK) return csv_last_error();
        }
        if (csv_batch_release_inputs(ctx, b) != CSV_OK) return csv_last_error();
        out.depth.dev = &dev; out.depth.batch = b; out.depth.region = 0; out.depth.beg = j.reg.beg; out.depth.end = j.reg.end;
        b = nullptr;                                                     // owned by the results from here on
        return "";
    }

    std::vector<ContigState>& contigs_;
    std::string bam_path_;
    std::mutex m_;
    std::condition_variable cv_jobs_, cv_free_;
    std::deque<Job> jobs_;
    std::vector<std::unique_ptr<PackedReads>> free_;
    std::vector<std::thread> workers_;
    bool stopping_ = false, failed_ = false;
    std::string error_;
};

// the shards of one contig -> its parked results
std::shared_ptr<ContigResults> assemble(ContigState& c, const std::string& bam_path, uint64_t* sum, uint32_t* nz)
{
    std::shared_ptr<ContigResults> r(new ContigResults);
    r->bam_path = bam_path; r->tid = c.tid; r->map_size = c.size;
    r->have_sigs = true; r->sig_min_mapq = kScanMinMapq; r->have_split = true;
    *sum = 0; *nz = 0;
    for (auto& sp : c.shards) {
        ShardOut& s = *sp;
        *sum += s.sum; *nz += s.nz;
        r->shards.push_back(s.depth);
        for (auto& e : s.seq4) r->seq4.emplace(e.first, std::move(e.second));
        csvhost::SplitRecords& d = r->split;
        d.pos.insert(d.pos.end(), s.split.pos.begin(), s.split.pos.end());
        d.endpos.insert(d.endpos.end(), s.split.endpos.begin(), s.split.endpos.end());
        d.query_start.insert(d.query_start.end(), s.split.query_start.begin(), s.split.query_start.end());
        d.query_end.insert(d.query_end.end(), s.split.query_end.begin(), s.split.query_end.end());
        d.flag.insert(d.flag.end(), s.split.flag.begin(), s.split.flag.end());
        d.mapq.insert(d.mapq.end(), s.split.mapq.begin(), s.split.mapq.end());
        const uint64_t base = d.names.size();
        d.names.insert(d.names.end(), s.split.names.begin(), s.split.names.end());
        for (size_t i = 1; i < s.split.name_off.size(); i++) d.name_off.push_back(base + s.split.name_off[i]);
    }
    if (c.shards.size() == 1) r->sigs = std::move(c.shards[0]->sigs);
    else {
        // every shard's run is in vector order; the order of the whole is (start, end) ascending with equal keys in
        // reverse insertion order = descending (record, op) (sv_object.cpp:17-33)
        csvhost::SigColumns all;
        for (auto& sp : c.shards) {
            const csvhost::SigColumns& s = sp->sigs;
            all.start.insert(all.start.end(), s.start.begin(), s.start.end()); all.end.insert(all.end.end(), s.end.begin(), s.end.end());
            all.op_idx.insert(all.op_idx.end(), s.op_idx.begin(), s.op_idx.end()); all.query_pos.insert(all.query_pos.end(), s.query_pos.begin(), s.query_pos.end());
            all.serial.insert(all.serial.end(), s.serial.begin(), s.serial.end()); all.kind.insert(all.kind.end(), s.kind.begin(), s.kind.end());
        }
        std::vector<size_t> ord(all.size());
        for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
            if (all.start[a] != all.start[b]) return all.start[a] < all.start[b];
            if (all.end[a] != all.end[b]) return all.end[a] < all.end[b];
            if (all.serial[a] != all.serial[b]) return all.serial[a] > all.serial[b];
            return all.op_idx[a] > all.op_idx[b];
        });
        csvhost::SigColumns& o = r->sigs;
        o.start.reserve(ord.size()); o.end.reserve(ord.size()); o.op_idx.reserve(ord.size()); o.query_pos.reserve(ord.size()); o.serial.reserve(ord.size()); o.kind.reserve(ord.size());
        for (size_t i : ord) {
            o.start.push_back(all.start[i]); o.end.push_back(all.end[i]); o.op_idx.push_back(all.op_idx[i]); o.query_pos.push_back(all.query_pos[i]);
            o.serial.push_back(all.serial[i]); o.kind.push_back(all.kind[i]);
        }
    }
    std::sort(r->shards.begin(), r->shards.end(), [](const csvhost::DepthShard& a, const csvhost::DepthShard& b) { return a.beg < b.beg; });
    c.shards.clear();
    return r;
}

}  // namespace

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

    // the reference's checks, chromosome by chromosome in the caller's order (cnv_caller.cpp:462-487)
    std::vector<ContigState> contigs;
    for (const std::string& chr : chromosomes) {
        hts_itr_t* bam_iter = sam_itr_querys(bam_index, bam_header, chr.c_str());
        if (!bam_iter) { printError("ERROR: Could not create iterator for chromosome: " + chr + ", check if the chromosome exists in the BAM file."); continue; }
        hts_itr_destroy(bam_iter);
        std::vector<uint32_t>& pos_depth_map = chr_pos_depth_map[chr];
        const int tid = bam_name2id(bam_header, chr.c_str());
        if (tid < 0) { printError("ERROR: Could not find chromosome " + chr + " in BAM file."); continue; }
        const uint32_t chr_length = bam_header->target_len[tid] + 1;
        if (pos_depth_map.size() != static_cast<size_t>(chr_length)) {
            printError("ERROR: Chromosome length mismatch for " + chr + ": expected " + std::to_string(chr_length) + ", found " +
                       std::to_string(pos_depth_map.size()) + ", resizing to " + std::to_string(chr_length));
            pos_depth_map.resize(chr_length, 0);
        }
        ContigState c;
        c.name = chr; c.tid = tid; c.size = chr_length; c.depth = &pos_depth_map;
        contigs.push_back(std::move(c));
    }
    // Records stream into a slab; whenever it would exceed the ops one batch may hold (60x ONT is ~3.7e10 over the
    // genome), the depth slice up to the current position goes to a GPU, and the records that reach past the cut are
    // copied into the next slab as its halo -- the region sharding of SURVEY 8e, in time and across CONTEXTSV_GPUS.
    const uint64_t max_ops = csvhost::max_ops_per_batch();
    const int total_chr_count = (int)chromosomes.size();
    std::string failure;
    {
        DepthPipeline pipe(contigs, bam_filepath);
        std::vector<size_t> submitted(contigs.size(), 0);
        size_t next_report = 0;
        auto report_done = [&](bool wait) {                      // mean coverage lines, in the caller's order, as contigs complete
            while (next_report < contigs.size() && contigs[next_report].closed) {
                ContigState& c = contigs[next_report];
                while (wait && pipe.done(next_report) < submitted[next_report] && !pipe.failed()) std::this_thread::yield();
                if (pipe.failed() || pipe.done(next_report) < submitted[next_report]) return;
                uint64_t cum_depth = 0; uint32_t pos_count = 0;
                std::shared_ptr<ContigResults> res = assemble(c, bam_filepath, &cum_depth, &pos_count);
                csvhost::put_results(c.depth, res);
                const double mean_chr_cov = (pos_count > 0) ? static_cast<double>(cum_depth) / static_cast<double>(pos_count) : 0.0;
                printMessage("Mean coverage for chromosome " + c.name + ": " + std::to_string(mean_chr_cov));
                if (mean_chr_cov != 0.0) chr_mean_cov_map[c.name] = mean_chr_cov;
                next_report++;
            }
        };
        for (size_t ci = 0; ci < contigs.size() && !pipe.failed(); ci++) {
            ContigState& c = contigs[ci];
            printMessage("(" + std::to_string(ci + 1) + "/" + std::to_string(total_chr_count) + ") Reading BAM file for chromosome: " + c.name);
            hts_itr_t* it = sam_itr_querys(bam_index, bam_header, c.name.c_str());
            if (!it) { c.closed = true; continue; }
            std::unique_ptr<PackedReads> reads = pipe.take_slab();
            uint32_t beg = 0;
            uint64_t serial = 0, first_own = 0;
            int32_t last_pos = -2;
            auto cut_here = [&](uint32_t end, bool last) {
                if (end <= beg && !last) return;
                std::unique_ptr<PackedReads> next;
                if (!last) {
                    next = pipe.take_slab();
                    reads->copy_reaching(end, *next);
                }
                Job j;
                j.reg = csv_region{c.tid, beg, std::max(end, beg + (last ? 0u : 1u)), c.size};
                if (last) j.reg.end = c.size;
                j.reads = std::move(reads);
                j.contig = ci; j.shard = submitted[ci]++; j.first_own = first_own;
                if (j.reg.beg < j.reg.end) pipe.submit(std::move(j)); else submitted[ci]--;
                reads = std::move(next);
                beg = end; first_own = serial;
            };
            {
                csvhost::StatTimer st_dec(csvhost::STAT_DECODE);
                while (sam_itr_next(bam_file, it, bam_record) >= 0) {
                    const int32_t pos = (int32_t)bam_record->core.pos;
                    if (reads->ops() + reads->size() + bam_record->core.n_cigar + 1 > max_ops && pos != last_pos && reads->size() > 0) {   // ops + records: a batch counts both
                        const uint32_t cut = std::min<uint32_t>((uint32_t)pos + 1u, c.size);      // records from here on start at or after the cut
                        if (cut > beg) cut_here(cut, false);
                    }
                    reads->append(bam_record, true, true, serial++);
                    last_pos = pos;
                    if (pipe.failed()) break;
                }
            }
            hts_itr_destroy(it);
            cut_here(c.size, true);
            c.closed = true;
            report_done(false);
        }
        pipe.finish();
        if (pipe.failed()) failure = pipe.error();
        else report_done(true);
    }

    printMessage("Closing BAM file " + bam_filepath);
    bam_destroy1(bam_record);
    hts_idx_destroy(bam_index);
    bam_hdr_destroy(bam_header);
    sam_close(bam_file);
    printMessage("BAM file closed.");
    // A GPU failure is fatal: partial maps would silently change every later result (the reference cannot get here).
    if (!failure.empty()) throw std::runtime_error("contextsv_b200 depth pass: " + failure);
}

// ---------------------------------------------------------------------------------------------------------------
// CNVCaller::querySNPRegion with the window sums taken from the device-resident depth map.  Same observable
// behaviour as src/cnv_caller.cpp:53-160: the SNP query, the sample-size rule, the window keys "<start>-<end>" held in
// an unordered_map (its iteration order decides the order of the HMM observations, so the same container is used),
// the dummy observation for a window without SNPs.
void CNVCaller::querySNPRegion(std::string chr, uint32_t start_pos, uint32_t end_pos, const std::vector<uint32_t>& pos_depth_map, double mean_chr_cov,
                               SNPData& snp_data, const InputData& input_data) const
{
    int sample_size = input_data.getSampleSize();
    std::vector<uint32_t> snp_pos;
    std::unordered_map<uint32_t, double> snp_baf_map, snp_pfb_map;
    this->readSNPAlleleFrequencies(chr, start_pos, end_pos, snp_pos, snp_baf_map, snp_pfb_map, input_data);
    sample_size = std::max((int)snp_pos.size(), sample_size);
    if (start_pos > end_pos) {
        printError("ERROR: Invalid SNP region for copy number prediction: " + chr + ":" + std::to_string((int)start_pos) + "-" + std::to_string((int)end_pos));
        return;
    }
    // integer window sums and position counts: prefetched by the caller's batch, else one launch for this region; a map
    // that never went through the depth pass (or CONTEXTSV_HOST_DEPTH=1) is read on the host like the reference does
    const double pos_step = static_cast<double>(end_pos - start_pos + 1) / static_cast<double>(sample_size);
    std::vector<uint64_t> sums_own;
    std::vector<uint32_t> counts_own;
    const uint64_t* sums = nullptr; const uint32_t* counts = nullptr;
    if (sample_size > 0 && !csvhost::prefetched_windows(&pos_depth_map, start_pos, end_pos, sample_size, &sums, &counts)) {
        sums_own.assign((size_t)sample_size, 0); counts_own.assign((size_t)sample_size, 0);
        const std::shared_ptr<ContigResults> res = csvhost::host_depth_requested() ? nullptr : csvhost::results_for_vector(&pos_depth_map);
        if (res) {
            if (!csvhost::device_window_sums(*res, 1, &start_pos, &end_pos, sample_size, sums_own.data(), counts_own.data()))
                throw std::runtime_error(std::string("contextsv_b200 querySNPRegion: ") + csv_last_error());
        } else {
            for (int i = 0; i < sample_size; i++) {
                for (int j = 0; j < pos_step; j++) {
                    const uint32_t pos = (uint32_t)(start_pos + i * pos_step + j);
                    if (pos > end_pos) break;
                    if (pos < pos_depth_map.size()) { sums_own[i] += pos_depth_map[pos]; counts_own[i]++; }
                }
            }
        }
        sums = sums_own.data(); counts = counts_own.data();
    }
    std::unordered_map<std::string, double> window_log2_map;
    for (int i = 0; i < sample_size; i++) {
        const uint32_t window_start = (uint32_t)(start_pos + i * pos_step);
        const uint32_t window_end = (uint32_t)(start_pos + (i + 1) * pos_step);
        double log2_cov = 0.0;
        if (counts[i] > 0) {
            double cov_sum = (double)sums[i];                     // a sum of uint32 depths: exact in a double either way
            if (cov_sum == 0) cov_sum = 1e-9;
            log2_cov = log2((cov_sum / (double)counts[i]) / mean_chr_cov);
        }
        window_log2_map[std::to_string(window_start) + "-" + std::to_string(window_end)] = log2_cov;
    }
    std::vector<uint32_t> pos_hmm;
    std::vector<double> baf_hmm, pfb_hmm, log2_hmm;
    std::vector<bool> is_snp_hmm;
    for (const auto& window : window_log2_map) {
        const size_t dash = window.first.find('-');
        const uint32_t window_start = std::stoi(window.first.substr(0, dash));
        const uint32_t window_end = std::stoi(window.first.substr(dash + 1));
        bool snp_found = false;
        for (uint32_t pos : snp_pos) {
            if (pos < window_start || pos > window_end) continue;
            pos_hmm.push_back(pos); baf_hmm.push_back(snp_baf_map[pos]); pfb_hmm.push_back(snp_pfb_map[pos]);
            log2_hmm.push_back(window.second); is_snp_hmm.push_back(true);
            snp_found = true;
        }
        if (!snp_found) {
            pos_hmm.push_back((window_start + window_end) / 2); baf_hmm.push_back(-1.0); pfb_hmm.push_back(0.5);
            log2_hmm.push_back(window.second); is_snp_hmm.push_back(false);
        }
    }
    snp_data.pos = std::move(pos_hmm);
    snp_data.baf = std::move(baf_hmm);
    snp_data.pfb = std::move(pfb_hmm);
    snp_data.log2_cov = std::move(log2_hmm);
    snp_data.is_snp = std::move(is_snp_hmm);
}

// The reference's own body of runCIGARCopyNumberPrediction, reachable under a second name (oracle/Makefile adds the
// alias to the unmodified object; in a source-level integration it is the function renamed): Itanium C++ ABI, `this`
// is the first argument, references are pointers.
extern "C" void csv_ref_runCIGARCopyNumberPrediction(const CNVCaller* self, std::string chr, std::vector<SVCall>& sv_candidates, const CHMM& hmm,
                                                     double mean_chr_cov, const std::vector<uint32_t>& pos_depth_map, const InputData& input_data);

void CNVCaller::runCIGARCopyNumberPrediction(std::string chr, std::vector<SVCall>& sv_candidates, const CHMM& hmm, double mean_chr_cov,
                                             const std::vector<uint32_t>& pos_depth_map, const InputData& input_data) const
{
    // the candidates the loop at cnv_caller.cpp:300-322 will query, all in one launch
    if (!csvhost::host_depth_requested() && csvhost::results_for_vector(&pos_depth_map)) {
        std::vector<uint32_t> start, end;
        for (const SVCall& sv : sv_candidates) {
            if (sv.start > sv.end || (sv.end - sv.start) < input_data.getMinCNVLength()) continue;
            start.push_back(sv.start); end.push_back(sv.end);
        }
        csvhost::prefetch_windows(&pos_depth_map, start, end, input_data.getSampleSize());
    }
    csv_ref_runCIGARCopyNumberPrediction(this, chr, sv_candidates, hmm, mean_chr_cov, pos_depth_map, input_data);
    csvhost::drop_prefetch(&pos_depth_map);
}

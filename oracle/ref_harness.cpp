/*
 * ref_harness.cpp -- C entry points into the UNMODIFIED ContextSV sources.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into
 * oracle/_ref/libcontextsv_ref.so together with the reference's own src/*.cpp
 * (compiled where they lie under /root/reference; nothing is copied) and the
 * htslib shim.  Compiled with -fno-access-control so the private members
 * SVCaller::findCIGARSVs / processCIGARRecord (sv_caller.h:80,86) and
 * CNVCaller::querySNPRegion (cnv_caller.h:56) can be called directly.
 */
#include "sv_caller.h"
#include "cnv_caller.h"
#include "dbscan1d.h"
#include "dbscan.h"
#include "sv_object.h"
#include "input_data.h"

#include "htslib_shim/shim_mem.h"

#include <atomic>
#include <mutex>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

namespace {
std::atomic<uint64_t> g_counter{0};
std::string unique_name() { return "h" + std::to_string(g_counter++); }

/* the reference chats on stdout/stderr for every call; keep test logs readable.
 * Process-wide fd redirection, reference-counted so concurrent harness calls
 * (bench.py runs one contig per thread) nest correctly. */
std::mutex g_quiet_mu;
int g_quiet_depth = 0, g_saved_out = -1, g_saved_err = -1;
struct Quiet {
    bool on;
    explicit Quiet(bool o) : on(o) {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_quiet_mu);
        if (g_quiet_depth++ == 0) {
            fflush(stdout); fflush(stderr);
            g_saved_out = dup(1); g_saved_err = dup(2);
            int nul = open("/dev/null", O_WRONLY);
            dup2(nul, 1); dup2(nul, 2); close(nul);
        }
    }
    ~Quiet() {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_quiet_mu);
        if (--g_quiet_depth == 0) {
            fflush(stdout); fflush(stderr);
            dup2(g_saved_out, 1); dup2(g_saved_err, 2); close(g_saved_out); close(g_saved_err);
        }
    }
};
int g_quiet = 1;
}  // namespace

extern "C" {

void ref_set_quiet(int q) { g_quiet = q; }

/* CNVCaller::calculateMeanChromosomeCoverage (cnv_caller.cpp:415-556) on one
 * contig of an in-memory table.  alloc_size is the size the caller
 * (sv_caller.cpp:801) gives the depth map (FASTA length + 1); the function
 * itself resizes to target_len+1 on mismatch.  depth_out must hold
 * target_len[tid]+1 entries. */
int ref_depth(const csvshim_mem* m, int32_t tid, uint32_t alloc_size,
              uint32_t* depth_out, uint64_t* sum_out, uint32_t* nonzero_out, double* mean_out)
{
    std::string name = unique_name();
    csvshim_register_mem(name.c_str(), m);
    std::string chr = m->target_name[tid];
    std::unordered_map<std::string, std::vector<uint32_t>> depth_map;
    std::unordered_map<std::string, double> mean_map;
    depth_map[chr] = std::vector<uint32_t>(alloc_size, 0);
    std::shared_mutex mu;
    CNVCaller cnv(mu);
    {
        Quiet q(g_quiet);
        cnv.calculateMeanChromosomeCoverage({chr}, depth_map, mean_map, "mem:" + name, 1);
    }
    csvshim_unregister_mem(name.c_str());
    const std::vector<uint32_t>& d = depth_map[chr];
    if (d.size() != (size_t)m->target_len[tid] + 1) return -1;
    uint64_t s = 0; uint32_t c = 0;
    for (size_t i = 0; i < d.size(); i++) { s += d[i]; c += d[i] > 0; }
    if (depth_out) memcpy(depth_out, d.data(), d.size() * sizeof(uint32_t));
    *sum_out = s; *nonzero_out = c;
    *mean_out = mean_map.count(chr) ? mean_map[chr] : 0.0;
    return 0;
}

/* SVCaller::findCIGARSVs (sv_caller.cpp:506-537) -> processCIGARRecord ->
 * addSVCall on one contig.  Returns the number of SVCalls; writes up to cap.
 * alt_out holds cap slots of 64 bytes (NUL-terminated ALT allele). */
int64_t ref_cigar_scan(const csvshim_mem* m, int32_t tid, uint32_t depth_map_size, int min_mapq,
                       uint32_t* start_out, uint32_t* end_out, int32_t* svtype_out,
                       uint32_t* evidence_out, char* alt_out, uint64_t cap)
{
    std::string name = unique_name();
    csvshim_register_mem(name.c_str(), m);
    std::string path = "mem:" + name;
    std::vector<SVCall> calls;
    {
        Quiet q(g_quiet);
        samFile* fp = sam_open(path.c_str(), "r");
        bam_hdr_t* hdr = sam_hdr_read(fp);
        hts_idx_t* idx = sam_index_load(fp, path.c_str());
        SVCaller caller;
        caller.min_mapq = min_mapq;
        std::vector<uint32_t> depth(depth_map_size, 0);
        caller.findCIGARSVs(fp, idx, hdr, m->target_name[tid], calls, depth);
        hts_idx_destroy(idx); bam_hdr_destroy(hdr); sam_close(fp);
    }
    csvshim_unregister_mem(name.c_str());
    for (size_t i = 0; i < calls.size() && i < cap; i++) {
        start_out[i] = calls[i].start; end_out[i] = calls[i].end;
        svtype_out[i] = (int32_t)calls[i].sv_type;
        evidence_out[i] = (uint32_t)calls[i].aln_type.to_ulong();
        if (alt_out) { strncpy(alt_out + 64 * i, calls[i].alt_allele.c_str(), 63); alt_out[64 * i + 63] = 0; }
    }
    return (int64_t)calls.size();
}

/* DBSCAN1D::fit + getClusters (dbscan1d.cpp:8-23) */
void ref_dbscan1d(const int* pts, uint64_t n, double eps, int min_pts, int* labels)
{
    std::vector<int> p(pts, pts + n);
    DBSCAN1D db(eps, min_pts);
    db.fit(p);
    const std::vector<int>& c = db.getClusters();
    if (n) memcpy(labels, c.data(), n * sizeof(int));
}

/* DBSCAN1D::getLargestCluster (dbscan1d.cpp:72-90) after fit */
uint64_t ref_largest_cluster(const int* pts, uint64_t n, double eps, int min_pts, int* out)
{
    std::vector<int> p(pts, pts + n);
    DBSCAN1D db(eps, min_pts);
    db.fit(p);
    std::vector<int> l = db.getLargestCluster(p);
    if (!l.empty()) memcpy(out, l.data(), l.size() * sizeof(int));
    return l.size();
}

/* CNVCaller::querySNPRegion (cnv_caller.cpp:53-164) with no SNP file: returns
 * the dummy-SNP positions (window centres) and log2 ratios in the order the
 * reference produces them (unordered_map iteration order).  Returns count. */
int ref_log2_windows(const uint32_t* depth, uint64_t map_size, uint32_t start_pos, uint32_t end_pos,
                     int sample_size, double mean_chr_cov, uint32_t* pos_out, double* log2_out, int cap)
{
    std::vector<uint32_t> d(depth, depth + map_size);
    InputData in;
    in.setSampleSize(sample_size);
    std::shared_mutex mu;
    CNVCaller cnv(mu);
    SNPData snp;
    {
        Quiet q(g_quiet);
        cnv.querySNPRegion("chrT", start_pos, end_pos, d, mean_chr_cov, snp, in);
    }
    int n = (int)snp.pos.size();
    for (int i = 0; i < n && i < cap; i++) { pos_out[i] = snp.pos[i]; log2_out[i] = snp.log2_cov[i]; }
    return n;
}

/* 2-D DBSCAN::fit (dbscan.cpp:9-81) on (start,end) pairs -- "next" row #1 */
void ref_dbscan2d(const uint32_t* start, const uint32_t* end, uint64_t n, double eps, int min_pts, int* labels)
{
    std::vector<SVCall> calls(n);
    for (uint64_t i = 0; i < n; i++) { calls[i].start = start[i]; calls[i].end = end[i]; }
    DBSCAN db(eps, min_pts);
    db.fit(calls);
    const std::vector<int>& c = db.getClusters();
    if (n) memcpy(labels, c.data(), n * sizeof(int));
}

/* What the split-read pass reads off every record (sv_caller.cpp:150-162): bam_endpos and
 * SVCaller::getAlignmentReadPositions (sv_caller.cpp:663-690), for the records of one contig in file order. */
int64_t ref_record_summary(const csvshim_mem* m, int32_t tid, int32_t* endpos_out, int32_t* qstart_out, int32_t* qend_out, uint64_t cap)
{
    std::string name = unique_name();
    csvshim_register_mem(name.c_str(), m);
    std::string path = "mem:" + name;
    uint64_t n = 0;
    {
        Quiet q(g_quiet);
        samFile* fp = sam_open(path.c_str(), "r");
        bam_hdr_t* hdr = sam_hdr_read(fp);
        hts_idx_t* idx = sam_index_load(fp, path.c_str());
        hts_itr_t* itr = sam_itr_querys(idx, hdr, m->target_name[tid]);
        bam1_t* b = bam_init1();
        SVCaller caller;
        while (itr && sam_itr_next(fp, itr, b) >= 0) {
            if (n < cap) {
                const std::pair<int, int> qp = caller.getAlignmentReadPositions(b);
                endpos_out[n] = (int32_t)bam_endpos(b); qstart_out[n] = qp.first; qend_out[n] = qp.second;
            }
            n++;
        }
        bam_destroy1(b);
        if (itr) hts_itr_destroy(itr);
        hts_idx_destroy(idx); bam_hdr_destroy(hdr); sam_close(fp);
    }
    csvshim_unregister_mem(name.c_str());
    return (int64_t)n;
}


/* SVCaller::findSplitSVSignatures (sv_caller.cpp:68-504) on a BAM file: the candidates of every chromosome, appended to
 * out_path one per line as chr, start, end, SVType, ALT, evidence bits, aln_offset, cluster_size -- the format the
 * drop-in's CONTEXTSV_B200_DUMP_SPLIT hook writes.  Chromosomes come in the reference's own (hash map) order.
 * Returns the number of candidates, or -1. */
int64_t ref_split_dump(const char* bam_path, const char* chr_or_empty, int threads, const char* out_path)
{
    std::unordered_map<std::string, std::vector<SVCall>> calls;
    try {
        Quiet q(g_quiet);
        InputData in;
        in.setLongReadBam(bam_path);
        in.setThreadCount(threads > 0 ? threads : 1);
        if (chr_or_empty && *chr_or_empty) in.setChromosome(chr_or_empty);
        SVCaller caller;
        caller.findSplitSVSignatures(calls, in);
    } catch (const std::exception&) { return -1; }
    FILE* f = fopen(out_path, "a");
    if (!f) return -1;
    int64_t n = 0;
    for (const auto& e : calls)
        for (const SVCall& c : e.second) {
            fprintf(f, "%s\t%u\t%u\t%d\t%s\t%lu\t%d\t%d\n", e.first.c_str(), c.start, c.end, (int)c.sv_type, c.alt_allele.c_str(), c.aln_type.to_ulong(), c.aln_offset, c.cluster_size);
            n++;
        }
    fclose(f);
    return n;
}
}  // extern "C"

// Test harness (CPU): the C++ host packer above the C ABI (contextsv_b200/csrc/host/packed_reads.*) driven through
// ctypes.  Links the packer and the htslib shim only -- no CUDA, no reference sources.
#include <cstring>
#include <memory>
#include <string>

#include "packed_reads.h"

using csvhost::PackedReads;

extern "C" {

// every record overlapping `chrom`, in file order, as the depth / CIGAR glue packs them
static int g_threads = 0;
// hts_set_threads for the files hp_pack opens from now on (the shim inflates BGZF blocks ahead on that many threads)
void hp_set_threads(int n) { g_threads = n; }

void* hp_pack(const char* bam, const char* chrom, int keep_seq)
{
    samFile* fp = sam_open(bam, "r");
    if (!fp) return nullptr;
    if (g_threads) hts_set_threads(fp, g_threads);
    bam_hdr_t* hdr = sam_hdr_read(fp);
    hts_idx_t* idx = hdr ? sam_index_load(fp, bam) : nullptr;
    hts_itr_t* it = idx ? sam_itr_querys(idx, hdr, chrom) : nullptr;
    PackedReads* p = nullptr;
    if (it) {
        bam1_t* b = bam_init1();
        p = new PackedReads;
        csvhost::pack_iterator(fp, it, b, *p, keep_seq != 0);
        bam_destroy1(b);
        hts_itr_destroy(it);
    }
    if (idx) hts_idx_destroy(idx);
    if (hdr) bam_hdr_destroy(hdr);
    sam_close(fp);
    return p;
}
void hp_free(void* p) { delete static_cast<PackedReads*>(p); }
uint64_t hp_size(void* p) { return static_cast<PackedReads*>(p)->size(); }
uint64_t hp_ops(void* p) { return static_cast<PackedReads*>(p)->ops(); }
void hp_copy(void* p, int32_t* tid, int32_t* pos0, uint16_t* flag, uint8_t* mapq, uint64_t* cig_off, uint32_t* cigar, uint32_t* ref_end)
{
    const PackedReads& r = *static_cast<PackedReads*>(p);
    const csv_reads v = r.view();
    memcpy(tid, v.tid, v.n_reads * 4); memcpy(pos0, v.pos0, v.n_reads * 4); memcpy(flag, v.flag, v.n_reads * 2);
    memcpy(mapq, v.mapq, v.n_reads); memcpy(cig_off, v.cig_off, (v.n_reads + 1) * 8); memcpy(cigar, v.cigar, v.n_ops * 4);
    memcpy(ref_end, r.ref_end.data(), v.n_reads * 4);
}
uint64_t hp_seq_count(void* p) { return static_cast<PackedReads*>(p)->seq4.size(); }
// bases [q, q + n) of record idx through base_at(); returns 0 if the record kept no sequence
int hp_bases(void* p, uint32_t idx, uint32_t q, uint32_t n, char* out)
{
    const PackedReads& r = *static_cast<PackedReads*>(p);
    const auto it = r.seq4.find(idx);
    if (it == r.seq4.end()) return 0;
    for (uint32_t j = 0; j < n; j++) out[j] = csvhost::base_at(it->second, q + j);
    return 1;
}
void hp_keep_reaching(void* p, uint32_t cut) { static_cast<PackedReads*>(p)->keep_reaching(cut); }
void hp_cache_put(const char* bam, int tid, void* p) { csvhost::cache_put(bam, tid, std::move(*static_cast<PackedReads*>(p))); }
void* hp_cache_take(const char* bam, int tid) { return csvhost::cache_take(bam, tid).release(); }
const char* hp_file_name(const char* bam)
{
    static std::string s;
    samFile* fp = sam_open(bam, "r");
    if (!fp) return nullptr;
    s = csvhost::file_name(fp);
    sam_close(fp);
    return s.c_str();
}
uint64_t hp_max_ops(void) { return csvhost::max_ops_per_batch(); }

}  // extern "C"

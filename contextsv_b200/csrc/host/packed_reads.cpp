#include "packed_reads.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>

namespace csvhost {

namespace {
std::mutex g_alloc_m;
void* (*g_alloc)(size_t) = nullptr;
void (*g_release)(void*) = nullptr;
void* heap_alloc(size_t n) { void* p = std::malloc(n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void heap_release(void* p) { std::free(p); }
}  // namespace

void set_slab_allocator(void* (*alloc)(size_t), void (*release)(void*))
{
    std::lock_guard<std::mutex> lk(g_alloc_m);
    g_alloc = alloc; g_release = release;
}

void* slab_alloc(size_t bytes, void (**release_out)(void*))
{
    void* (*a)(size_t); void (*r)(void*);
    { std::lock_guard<std::mutex> lk(g_alloc_m); a = g_alloc; r = g_release; }
    if (a && r) {
        if (void* p = a(bytes)) { *release_out = r; return p; }       // pinned memory can run out: fall back to the heap
    }
    *release_out = heap_release;
    return heap_alloc(bytes);
}

void PackedReads::append(const bam1_t* b, bool keep_seq, bool keep_name, uint64_t serial_no)
{
    const uint32_t idx = (uint32_t)pos0.size();
    tid.push_back(b->core.tid);
    pos0.push_back((int32_t)b->core.pos);
    flag.push_back(b->core.flag);
    mapq.push_back(b->core.qual);
    serial.push_back(serial_no);
    const uint32_t* c = bam_get_cigar(b);
    const uint32_t n = b->core.n_cigar;
    cigar.append(c, n);                                     // one copy of the record's CIGAR words
    bool want_seq = false;
    uint32_t rlen = 0, gaps = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t op = bam_cigar_op(c[i]), len = bam_cigar_oplen(c[i]);
        want_seq |= len == 50 && (op == BAM_CINS || op == BAM_CSOFT_CLIP);
        if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF) rlen += len;
        gaps += (op == BAM_CDEL) | (op == BAM_CREF_SKIP);
    }
    cig_off.push_back(cigar.size());
    ref_end.push_back((uint32_t)b->core.pos + 1u + rlen);
    n_gap.push_back(gaps);                                  // csv_reads::n_gap / ref_len: the device's record-level pre-pass starts from them
    ref_len.push_back(rlen);
    if (keep_seq && want_seq) {
        const uint8_t* s = bam_get_seq(b);
        seq4[idx].assign(s, s + ((size_t)b->core.l_qseq + 1) / 2);
    }
    if (keep_name) {
        if (name_off.empty()) name_off.push_back(0);
        const char* q = bam_get_qname(b);
        names.append(q, std::strlen(q));
        name_off.push_back(names.size());
    }
}

void PackedReads::clear()
{
    tid.clear(); pos0.clear(); flag.clear(); mapq.clear(); cigar.clear(); ref_end.clear(); n_gap.clear(); ref_len.clear(); serial.clear(); seq4.clear();
    names.clear(); name_off.clear();
    cig_off.clear(); cig_off.push_back(0);
}

void PackedReads::copy_reaching(uint32_t cut, PackedReads& k) const
{
    const bool with_names = !name_off.empty();
    for (size_t i = 0; i < pos0.size(); i++) {
        if (ref_end[i] <= cut) continue;
        const uint32_t j = (uint32_t)k.pos0.size();
        k.tid.push_back(tid[i]); k.pos0.push_back(pos0[i]); k.flag.push_back(flag[i]); k.mapq.push_back(mapq[i]); k.ref_end.push_back(ref_end[i]);
        k.n_gap.push_back(n_gap[i]); k.ref_len.push_back(ref_len[i]); k.serial.push_back(serial[i]);
        k.cigar.append(cigar.data() + cig_off[i], (size_t)(cig_off[i + 1] - cig_off[i]));
        k.cig_off.push_back(k.cigar.size());
        const auto s = seq4.find((uint32_t)i);
        if (s != seq4.end()) k.seq4[j] = s->second;
        if (with_names) {
            if (k.name_off.empty()) k.name_off.push_back(0);
            k.names.append(names.data() + name_off[i], (size_t)(name_off[i + 1] - name_off[i]));
            k.name_off.push_back(k.names.size());
        }
    }
}

void PackedReads::keep_reaching(uint32_t cut)
{
    PackedReads k;
    copy_reaching(cut, k);
    *this = std::move(k);
}

uint64_t max_ops_per_batch()
{
    static const uint64_t v = [] {
        const char* e = std::getenv("CONTEXTSV_MAX_OPS");
        const uint64_t hard = (1ull << 31) - (1ull << 20);
        if (!e) return hard;
        const uint64_t x = std::strtoull(e, nullptr, 10);
        return x == 0 || x > hard ? hard : x;
    }();
    return v;
}

namespace {
struct Cache {
    std::mutex m;
    std::map<std::pair<std::string, int>, std::unique_ptr<PackedReads>> entries;
    uint64_t ops = 0;
    uint64_t budget = [] {
        const char* e = std::getenv("CONTEXTSV_CACHE_OPS");
        return e ? std::strtoull(e, nullptr, 10) : (1ull << 30);
    }();
};
Cache g_cache;
}  // namespace

void cache_put(const std::string& bam_path, int tid, PackedReads&& reads)
{
    std::lock_guard<std::mutex> lk(g_cache.m);
    const uint64_t n = reads.ops() + reads.size();
    if (g_cache.ops + n > g_cache.budget) return;
    auto& slot = g_cache.entries[{bam_path, tid}];
    if (slot) g_cache.ops -= slot->ops() + slot->size();
    slot.reset(new PackedReads(std::move(reads)));
    g_cache.ops += n;
}

std::unique_ptr<PackedReads> cache_take(const char* bam_path, int tid)
{
    if (!bam_path) return nullptr;
    std::lock_guard<std::mutex> lk(g_cache.m);
    auto it = g_cache.entries.find({std::string(bam_path), tid});
    if (it == g_cache.entries.end()) return nullptr;
    std::unique_ptr<PackedReads> r = std::move(it->second);
    g_cache.entries.erase(it);
    g_cache.ops -= r->ops() + r->size();
    return r;
}

const char* file_name(samFile* fp)
{
#ifdef CSVSHIM_HTSLIB
    return hts_get_fn(fp);
#else
    return fp ? fp->fn : nullptr;
#endif
}

csv_reads PackedReads::view() const
{
    csv_reads r = {};
    r.n_reads = (uint32_t)pos0.size();
    r.n_ops = cigar.size();
    r.tid = tid.data(); r.pos0 = pos0.data(); r.flag = flag.data(); r.mapq = mapq.data();
    r.cig_off = cig_off.data(); r.cigar = cigar.data(); r.n_gap = n_gap.data(); r.ref_len = ref_len.data();
    return r;
}

void pack_iterator(samFile* fp, hts_itr_t* itr, bam1_t* scratch, PackedReads& out, bool keep_seq)
{
    while (sam_itr_next(fp, itr, scratch) >= 0) out.append(scratch, keep_seq);
}

char base_at(const std::vector<uint8_t>& seq4, uint32_t i)
{
    // SEQ '*' (l_qseq == 0) or a CIGAR longer than SEQ: the reference reads whatever follows inside the bam1_t data block;
    // here the base is reported as N instead of reading past the packed bases
    if ((size_t)(i >> 1) >= seq4.size()) return 'N';
    const char base = seq_nt16_str[bam_seqi(seq4.data(), i)];
    switch (base) {   // ambiguous bases -> N, either case
        case 'R': case 'Y': case 'K': case 'M': case 'S': case 'W': case 'B': case 'D': case 'H': case 'V':
        case 'r': case 'y': case 'k': case 'm': case 's': case 'w': case 'b': case 'd': case 'h': case 'v':
            return 'N';
        default: return base;
    }
}

}  // namespace csvhost

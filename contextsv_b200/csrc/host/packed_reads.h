// packed_reads.h -- host packer: htslib records -> the flat SoA of include/contextsv_b200.h.
//
// This is the C++ host side above the C ABI.  It only uses the htslib calls the reference itself
// uses (sam_itr_next, bam_get_cigar, ... -- src/cnv_caller.cpp:488-503, src/sv_caller.cpp:509-546),
// so it links against real htslib in the reference's own build and against oracle/htslib_shim here.
//
// The SoA lives in slabs that grow by doubling and are REUSED from shard to shard (clear() keeps the
// memory), allocated with a pluggable allocator: the GPU glue installs csv_host_alloc / csv_host_free
// (pinned memory: the upload is a plain DMA) as soon as the CUDA runtime is up; before that, and in
// builds without the CUDA library (tests/native), slabs are ordinary heap memory.  Each slab remembers
// which allocator it came from.
#pragma once
#include <htslib/sam.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "contextsv_b200.h"

namespace csvhost {

// Allocator of the slabs that are created from now on (nullptr, nullptr = heap).  Thread-safe.
void set_slab_allocator(void* (*alloc)(size_t), void (*release)(void*));
void* slab_alloc(size_t bytes, void (**release_out)(void*));

template <class T>
class Slab {
public:
    Slab() = default;
    Slab(const Slab&) = delete;
    Slab& operator=(const Slab&) = delete;
    Slab(Slab&& o) noexcept { steal(o); }
    Slab& operator=(Slab&& o) noexcept { if (this != &o) { drop(); steal(o); } return *this; }
    ~Slab() { drop(); }

    T* data() { return p_; }
    const T* data() const { return p_; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    bool empty() const { return n_ == 0; }
    T& operator[](size_t i) { return p_[i]; }
    const T& operator[](size_t i) const { return p_[i]; }
    T& back() { return p_[n_ - 1]; }
    void clear() { n_ = 0; }                                 // keeps the memory
    void reserve(size_t want)
    {
        if (want <= cap_) return;
        size_t cap = cap_ ? cap_ : 1024;
        while (cap < want) cap *= 2;
        void (*rel)(void*) = nullptr;
        T* q = static_cast<T*>(slab_alloc(cap * sizeof(T), &rel));
        if (n_) std::memcpy(q, p_, n_ * sizeof(T));
        drop_memory();
        p_ = q; cap_ = cap; release_ = rel;
    }
    void push_back(const T& v) { if (n_ == cap_) reserve(n_ + 1); p_[n_++] = v; }
    // n uninitialised elements at the end; returns where they start
    T* grow(size_t n) { reserve(n_ + n); T* at = p_ + n_; n_ += n; return at; }
    void append(const T* src, size_t n) { if (n) std::memcpy(grow(n), src, n * sizeof(T)); }
    void assign(size_t n, const T& v) { clear(); reserve(n); for (size_t i = 0; i < n; i++) p_[i] = v; n_ = n; }

private:
    T* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
    void (*release_)(void*) = nullptr;
    void drop_memory() { if (p_) release_(p_); p_ = nullptr; cap_ = 0; }
    void drop() { drop_memory(); n_ = 0; }
    void steal(Slab& o) { p_ = o.p_; n_ = o.n_; cap_ = o.cap_; release_ = o.release_; o.p_ = nullptr; o.n_ = o.cap_ = 0; }
};

struct PackedReads {
    Slab<int32_t> tid, pos0;
    Slab<uint16_t> flag;
    Slab<uint8_t> mapq;
    Slab<uint64_t> cig_off;          // [size() + 1]
    Slab<uint32_t> cigar;
    Slab<uint32_t> ref_end;          // pos0 + 1 + reference bases consumed: depth index one past the last covered base
    Slab<uint32_t> n_gap;            // D / N ops per record (csv_reads::n_gap)
    Slab<uint32_t> ref_len;          // reference bases consumed per record (csv_reads::ref_len)
    Slab<uint64_t> serial;           // number of the record in its iterator's order (stays with a record that moves on as a halo)
    // 4-bit bases of the few records that carry an I / S op of exactly 50 bases: the only place the
    // reference looks at the sequence on this path (literal ALT allele, src/sv_caller.cpp:572-591).  Keyed by index.
    std::unordered_map<uint32_t, std::vector<uint8_t>> seq4;
    // What the split-read pass reads off a record besides its CIGAR (src/sv_caller.cpp:140-162), kept when
    // append() is asked for names: the query name, as offsets into one character pool.
    Slab<char> names;
    Slab<uint64_t> name_off;         // [size() + 1] when names are kept, else empty

    PackedReads() { cig_off.push_back(0); }
    void append(const bam1_t* b, bool keep_seq, bool keep_name = false, uint64_t serial_no = 0);
    csv_reads view() const;
    size_t size() const { return pos0.size(); }
    uint64_t ops() const { return cigar.size(); }
    void clear();
    // keeps only the records that reach past depth index `cut` (the halo of the next shard), in order
    void keep_reaching(uint32_t cut);
    // ... or copies them to the end of another (empty) packing
    void copy_reaching(uint32_t cut, PackedReads& into) const;
};

// CIGAR ops + records one batch may hold: the C ABI takes < 2^31; CONTEXTSV_MAX_OPS lowers it (tests, small GPUs)
uint64_t max_ops_per_batch();

// Every record an iterator yields, in file order.
void pack_iterator(samFile* fp, hts_itr_t* itr, bam1_t* scratch, PackedReads& out, bool keep_seq);

// Packing cache between two passes over the same contig (kept for hosts that run the passes apart; the drop-in itself
// hands the depth pass's RESULTS on, see scan_results.h): at most CONTEXTSV_CACHE_OPS CIGAR ops + records in total
// (default 2^30; 0 disables).
void cache_put(const std::string& bam_path, int tid, PackedReads&& reads);
std::unique_ptr<PackedReads> cache_take(const char* bam_path, int tid);      // removes the entry; null if absent
// Path the file was opened with (htsFile::fn in htslib; an accessor in the shim, whose htsFile is opaque).
const char* file_name(samFile* fp);

// seq_nt16_str[bam_seqi(seq, i)] with the IUPAC -> N mapping of src/sv_caller.cpp:554-559,576-580
char base_at(const std::vector<uint8_t>& seq4, uint32_t i);

}  // namespace csvhost

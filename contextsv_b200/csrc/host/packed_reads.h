// packed_reads.h -- host packer: htslib records -> the flat SoA of include/contextsv_b200.h.
//
// This is the C++ host side above the C ABI.  It only uses the htslib calls the reference itself
// uses (sam_itr_next, bam_get_cigar, ... -- src/cnv_caller.cpp:488-503, src/sv_caller.cpp:509-546),
// so it links against real htslib in the reference's own build and against oracle/htslib_shim here.
#pragma once
#include <htslib/sam.h>

#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "contextsv_b200.h"

namespace csvhost {

struct PackedReads {
    std::vector<int32_t> tid, pos0;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq;
    std::vector<uint64_t> cig_off{0};
    std::vector<uint32_t> cigar;
    std::vector<uint32_t> ref_end;   // pos0 + 1 + reference bases consumed: depth index one past the last covered base
    std::vector<uint32_t> n_gap;     // D / N ops per record (csv_reads::n_gap)
    // 4-bit bases of the few records that carry an I / S op of exactly 50 bases: the only place the
    // reference looks at the sequence on this path (literal ALT allele, src/sv_caller.cpp:572-591)
    std::unordered_map<uint32_t, std::vector<uint8_t>> seq4;

    void append(const bam1_t* b, bool keep_seq);
    csv_reads view() const;
    size_t size() const { return pos0.size(); }
    uint64_t ops() const { return cigar.size(); }
    void clear();
    // keeps only the records that reach past depth index `cut` (the halo of the next shard), in order
    void keep_reaching(uint32_t cut);
};

// CIGAR ops one batch may hold: the C ABI takes < 2^31; CONTEXTSV_MAX_OPS lowers it (tests, small GPUs)
uint64_t max_ops_per_batch();

// Every record an iterator yields, in file order.
void pack_iterator(samFile* fp, hts_itr_t* itr, bam1_t* scratch, PackedReads& out, bool keep_seq);

// Single decode: the depth pass packs every record of a contig, and the CIGAR pass of the same contig
// (src/sv_caller.cpp:692-745, run later from a pool thread) needs exactly the same records -- the reference decodes
// the BAM again for it.  The depth pass parks the packing of every contig it scanned in one piece here and the CIGAR
// pass takes it instead of re-reading the file.  Bounded: at most CONTEXTSV_CACHE_OPS CIGAR ops in total (default 2^30,
// 4 GB; a 30x HiFi genome holds 4e8; 0 disables), beyond that the CIGAR pass decodes as before.
void cache_put(const std::string& bam_path, int tid, PackedReads&& reads);
std::unique_ptr<PackedReads> cache_take(const char* bam_path, int tid);      // removes the entry; null if absent
// Path the file was opened with (htsFile::fn in htslib; an accessor in the shim, whose htsFile is opaque).
const char* file_name(samFile* fp);

// seq_nt16_str[bam_seqi(seq, i)] with the IUPAC -> N mapping of src/sv_caller.cpp:554-559,576-580
char base_at(const std::vector<uint8_t>& seq4, uint32_t i);

}  // namespace csvhost

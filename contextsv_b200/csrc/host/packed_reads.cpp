#include "packed_reads.h"

#include <cstring>

namespace csvhost {

void PackedReads::append(const bam1_t* b, bool keep_seq)
{
    const uint32_t idx = (uint32_t)pos0.size();
    tid.push_back(b->core.tid);
    pos0.push_back((int32_t)b->core.pos);
    flag.push_back(b->core.flag);
    mapq.push_back(b->core.qual);
    const uint32_t* c = bam_get_cigar(b);
    const uint32_t n = b->core.n_cigar;
    bool want_seq = false;
    for (uint32_t i = 0; i < n; i++) {
        cigar.push_back(c[i]);
        const uint32_t op = bam_cigar_op(c[i]), len = bam_cigar_oplen(c[i]);
        if (len == 50 && (op == BAM_CINS || op == BAM_CSOFT_CLIP)) want_seq = true;
    }
    cig_off.push_back(cigar.size());
    if (keep_seq && want_seq) {
        const uint8_t* s = bam_get_seq(b);
        seq4[idx].assign(s, s + ((size_t)b->core.l_qseq + 1) / 2);
    }
}

csv_reads PackedReads::view() const
{
    csv_reads r;
    r.n_reads = (uint32_t)pos0.size();
    r.n_ops = cigar.size();
    r.tid = tid.data(); r.pos0 = pos0.data(); r.flag = flag.data(); r.mapq = mapq.data();
    r.cig_off = cig_off.data(); r.cigar = cigar.data();
    return r;
}

void pack_iterator(samFile* fp, hts_itr_t* itr, bam1_t* scratch, PackedReads& out, bool keep_seq)
{
    while (sam_itr_next(fp, itr, scratch) >= 0) out.append(scratch, keep_seq);
}

char base_at(const std::vector<uint8_t>& seq4, uint32_t i)
{
    const char base = seq_nt16_str[bam_seqi(seq4.data(), i)];
    switch (base) {   // ambiguous bases -> N, either case
        case 'R': case 'Y': case 'K': case 'M': case 'S': case 'W': case 'B': case 'D': case 'H': case 'V':
        case 'r': case 'y': case 'k': case 'm': case 's': case 'w': case 'b': case 'd': case 'h': case 'v':
            return 'N';
        default: return base;
    }
}

}  // namespace csvhost

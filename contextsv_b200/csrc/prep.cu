// prep.cu -- per-batch tables built once after upload.
//
// One chained-scan launch over the records: compacts the non-empty ones
// (records with n_cigar == 0 touch nothing in either reference loop), packs
// their metadata into one 16-byte word, resolves the region that owns each
// record and raises the "first op of a record" bit of every record in the
// op-indexed bitmap the walk kernel reads.
#include "batch.cuh"
#include "scan.cuh"

namespace csv {

__device__ __forceinline__ uint32_t find_owner(const TidDev td, const RegionDev* __restrict__ regs, uint32_t idx)
{
    for (uint32_t r = td.first; r < td.first + td.count; r++)
        if (idx >= regs[r].beg && idx < regs[r].end) return regs[r].orig;
    if (idx >= td.map_size)
        for (uint32_t r = td.first; r < td.first + td.count; r++)
            if (regs[r].end == td.map_size) return regs[r].orig;
    return kNone;
}

struct PrepParams {
    const unsigned long long* cig_off;
    const int32_t *tid, *pos0;
    const uint16_t* flag;
    const uint8_t* mapq;
    uint4* meta;
    uint32_t* ne_idx;
    unsigned long long* key;
    uint32_t* headbits;
    const TidDev* tids;
    const RegionDev* regs;
    uint32_t n_tids, n_reads, min_mapq;
    uint32_t* scalars;
};

// tables of record i, written at compact index k
__device__ __forceinline__ void prep_record(const PrepParams& P, uint32_t i, uint32_t k)
{
    const unsigned long long o = P.cig_off[i];
    atomicOr(&P.headbits[o >> 5], 1u << (o & 31));
    const int32_t t = P.tid ? P.tid[i] : 0;
    uint4 m;
    m.x = (uint32_t)P.pos0[i];
    m.z = (uint32_t)P.flag[i] | ((uint32_t)P.mapq[i] << 16);
    if (t < 0 || (uint32_t)t >= P.n_tids || P.tids[t].count == 0) { m.y = 0u; m.w = kNone; }   // contig not requested
    else {
        m.y = P.tids[t].map_size; m.w = find_owner(P.tids[t], P.regs, m.x + 1u);
        // the two record filters, decided once: depth (cnv_caller.cpp:491-495) and signatures (sv_caller.cpp:526)
        if (!(m.z & kDepthSkipFlags)) m.z |= 1u << 30;
        if (!(m.z & kSigSkipFlags) && (uint32_t)P.mapq[i] >= P.min_mapq && m.w != kNone) m.z |= 1u << 31;
    }
    P.meta[k] = m;
    P.key[k] = ((unsigned long long)(uint32_t)t << 32) | (uint32_t)(m.x + 1u);     // coordinate sort key of the batch
    P.ne_idx[k] = i;
}

// Optimistic pass: a batch without empty CIGARs (the rule: only unmapped records have none) needs no compaction,
// compact index == record index, and the tables are one fully parallel kernel.  A record with an empty CIGAR raises
// SC_HAS_EMPTY instead; the chained-scan kernel below then redoes the tables with compaction (it exits at once otherwise).
__global__ void __launch_bounds__(256) k_prep_dense(const PrepParams P)
{
    bool empty = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_reads; i += gridDim.x * blockDim.x) {
        if (P.cig_off[i + 1] > P.cig_off[i]) prep_record(P, i, i); else empty = true;
        if (i == P.n_reads - 1) {   // sentinel: the op after the last one starts "a new record"
            const unsigned long long e = P.cig_off[P.n_reads];
            atomicOr(&P.headbits[e >> 5], 1u << (e & 31));
            P.scalars[SC_N_NONEMPTY] = P.n_reads;
        }
    }
    if (__any_sync(0xffffffffu, empty) && (threadIdx.x & 31) == 0) P.scalars[SC_HAS_EMPTY] = 1u;
}

int launch_prep(csv_ctx* ctx, csv_batch* b, uint32_t min_mapq)
{
    // (the head-bit map was zeroed at upload and only ever receives the same bits; the scalars at the start of the pass)
    if (b->n_reads == 0) return CSV_OK;
    PrepParams P;
    P.cig_off = b->d_cig_off.as<unsigned long long>();
    P.tid = b->has_tid ? b->d_tid.as<int32_t>() : nullptr;
    P.pos0 = b->d_pos0.as<int32_t>();
    P.flag = b->d_flag.as<uint16_t>();
    P.mapq = b->d_mapq.as<uint8_t>();
    P.meta = b->d_meta.as<uint4>();
    P.ne_idx = b->d_ne_idx.as<uint32_t>();
    P.key = b->d_key.as<unsigned long long>();
    P.headbits = b->d_headbits.as<uint32_t>();
    P.tids = b->d_tids.as<TidDev>();
    P.regs = b->d_regs.as<RegionDev>();
    P.n_tids = b->n_tids; P.n_reads = b->n_reads; P.min_mapq = min_mapq;
    P.scalars = b->d_scalars.as<uint32_t>();
    const uint32_t grid = (b->n_reads + 255) / 256 < (uint32_t)ctx->sm_count * 16 ? (b->n_reads + 255) / 256 : (uint32_t)ctx->sm_count * 16;
    k_prep_dense<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    // compaction path, only if the dense pass met an empty CIGAR (head bits set twice are harmless: atomicOr)
    const uint32_t* has_empty = P.scalars + SC_HAS_EMPTY;
    const PrepParams Q = P;
    auto in = [=] __device__(uint64_t i) -> uint32_t { return Q.cig_off[i + 1] > Q.cig_off[i] ? 1u : 0u; };
    auto out = [=] __device__(uint64_t i, uint32_t k, uint32_t v) { if (v) prep_record(Q, (uint32_t)i, k); };
    return chained_scan(ctx, in, out, b->n_reads, nullptr, P.scalars + SC_N_NONEMPTY, has_empty);
}

}  // namespace csv

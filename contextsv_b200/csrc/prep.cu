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

int launch_prep(csv_ctx* ctx, csv_batch* b, uint32_t min_mapq)
{
    const uint64_t n_ops = b->n_ops;
    CSV_CUDA(cudaMemsetAsync(b->d_headbits.p, 0, n_ops / 8 + 16, ctx->stream));
    CSV_CUDA(cudaMemsetAsync(b->d_scalars.p, 0, SC_COUNT * sizeof(uint32_t), ctx->stream));
    if (b->n_reads == 0) return CSV_OK;
    const unsigned long long* cig_off = b->d_cig_off.as<unsigned long long>();
    const int32_t* tid = b->has_tid ? b->d_tid.as<int32_t>() : nullptr;
    const int32_t* pos0 = b->d_pos0.as<int32_t>();
    const uint16_t* flag = b->d_flag.as<uint16_t>();
    const uint8_t* mapq = b->d_mapq.as<uint8_t>();
    uint4* meta = b->d_meta.as<uint4>();
    uint32_t* ne_idx = b->d_ne_idx.as<uint32_t>();
    unsigned long long* key = b->d_key.as<unsigned long long>();
    uint32_t* headbits = b->d_headbits.as<uint32_t>();
    const TidDev* tids = b->d_tids.as<TidDev>();
    const RegionDev* regs = b->d_regs.as<RegionDev>();
    const uint32_t n_tids = b->n_tids, n_reads = b->n_reads;
    auto in = [=] __device__(uint64_t i) -> uint32_t { return cig_off[i + 1] > cig_off[i] ? 1u : 0u; };
    auto out = [=] __device__(uint64_t i, uint32_t k, uint32_t v) {
        if (i == n_reads - 1) {   // sentinel: the op after the last one starts "a new record"
            unsigned long long e = cig_off[n_reads];
            atomicOr(&headbits[e >> 5], 1u << (e & 31));
        }
        if (!v) return;
        unsigned long long o = cig_off[i];
        atomicOr(&headbits[o >> 5], 1u << (o & 31));
        int32_t t = tid ? tid[i] : 0;
        uint4 m;
        m.x = (uint32_t)pos0[i];
        m.z = (uint32_t)flag[i] | ((uint32_t)mapq[i] << 16);
        if (t < 0 || (uint32_t)t >= n_tids || tids[t].count == 0) { m.y = 0u; m.w = kNone; }   // contig not requested
        else {
            m.y = tids[t].map_size; m.w = find_owner(tids[t], regs, m.x + 1u);
            // the two record filters, decided once: depth (cnv_caller.cpp:491-495) and signatures (sv_caller.cpp:526)
            if (!(m.z & kDepthSkipFlags)) m.z |= 1u << 30;
            if (!(m.z & kSigSkipFlags) && (uint32_t)mapq[i] >= min_mapq && m.w != kNone) m.z |= 1u << 31;
        }
        meta[k] = m;
        key[k] = ((unsigned long long)(uint32_t)t << 32) | (uint32_t)(m.x + 1u);     // coordinate sort key of the batch
        ne_idx[k] = (uint32_t)i;
    };
    return chained_scan(ctx, in, out, n_reads, nullptr, b->d_scalars.as<uint32_t>() + SC_N_NONEMPTY);
}

}  // namespace csv

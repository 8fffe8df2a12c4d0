// sigs.cu -- order the emitted signatures exactly like the reference's
// vector<SVCall> (addSVCall, sv_object.cpp:22-33: ascending (start,end), equal
// keys in reverse insertion order) and materialise the output SoA.
//
// The walk appends signatures in arbitrary order with the 128-bit key
//   hi = owner region << 32 | start      lo = end << 32 | ~(global op index).
// Insertion order in the reference is (record order, op order) == ascending global
// op index, hence ~index puts equal (start,end) in reverse insertion order, and the
// key is unique.  Only hi is radix-sorted (5 byte passes instead of 13: the tiny
// passes of this side stream compete with the tile kernel for SM slots, so their
// NUMBER is what costs); the order inside a run of equal (region, start) -- a
// handful of entries, one per read over the same breakpoint -- is fixed by ranking
// lo inside the run.
//
// Default ordering (launch_sig_finish): no sort at all.  Every signature was counted into the depth tile its start
// falls into when the walk emitted it (SigRaw::bucket / arrival), tiles are in (region, position) order, so
//   one chained scan over the per-tile counts  -> first slot of every bucket (and the counts zeroed for the next pass),
//   one scatter                                -> the signatures of a bucket side by side, in arrival order,
//   one rank + gather kernel                   -> position inside the bucket = number of its entries with a smaller
//                                                 128-bit key, outputs written in place.
// A bucket holds the signatures of 8192 reference positions: one or two at 30x.  Three launches instead of the radix
// sort's ten on a stream that runs beside the HBM-bound tile kernel, where every launch costs the tiles SM slots.  The
// cost is quadratic in the bucket size like the tie ranking it replaces (2e5 signatures in one tile: a few ms); the
// radix path stays as CSV_SIG_ORDER=radix.
#include "batch.cuh"
#include "scan.cuh"

#include <cstdlib>

namespace csv {

// position of every entry inside its run of equal hi = number of entries of the run with a smaller lo.
// The run's bounds come from two binary searches over the sorted hi (O(log n) however long the run is); the ranking
// itself reads the run once per entry.  Runs are a handful of entries (one per read over a breakpoint); a pile-up of n
// entries at one (region, start) costs n^2 comparisons, but every lane of a warp inside the run reads the SAME entry in
// the same iteration (two broadcast loads for 32 comparisons): 2e5 entries at one coordinate take a few milliseconds
// (tests/test_gpu_parity.py::test_signature_pileup_at_one_start).
__global__ void k_sig_tiefix(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ raw_lo, const uint32_t* __restrict__ val,
                             const uint32_t* scalars, uint32_t* out)
{
    const uint32_t n = scalars[SC_N_SIG_EFF];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long h = hi[i];
        uint32_t a = i, b = i + 1;
        const bool tie_left = i > 0 && hi[i - 1] == h, tie_right = i + 1 < n && hi[i + 1] == h;
        if (tie_left) {                                      // first index with hi == h: gallop down, then bisect
            uint32_t lo_ = 0, hi_ = i - 1;                   // hi[hi_] == h
            for (uint32_t step = 1; hi_ > 0; step <<= 1) {
                const uint32_t probe = hi_ > step ? hi_ - step : 0u;
                if (hi[probe] != h) { lo_ = probe + 1; break; }
                hi_ = probe;
            }
            while (lo_ < hi_) { const uint32_t mid = lo_ + ((hi_ - lo_) >> 1); if (hi[mid] == h) hi_ = mid; else lo_ = mid + 1; }
            a = hi_;
        }
        if (tie_right) {                                     // one past the last index with hi == h
            uint32_t lo_ = i + 1, hi_ = n;                   // hi[lo_] == h
            for (uint32_t step = 1; lo_ + 1 < n; step <<= 1) {
                const uint32_t probe = n - lo_ > step ? lo_ + step : n;
                if (probe >= n || hi[probe] != h) { hi_ = probe < n ? probe : n; break; }
                lo_ = probe;
            }
            while (lo_ + 1 < hi_) { const uint32_t mid = lo_ + ((hi_ - lo_) >> 1); if (hi[mid] == h) lo_ = mid; else hi_ = mid; }
            b = lo_ + 1;
        }
        const uint32_t mine = val[i];
        uint32_t rank = 0;
        if (b - a > 1) {
            const unsigned long long my = raw_lo[mine];
            for (uint32_t j = a; j < b; j++) rank += raw_lo[val[j]] < my ? 1u : 0u;
        }
        out[a + rank] = mine;
    }
}

struct GatherParams {
    const unsigned long long *hi, *lo;      // hi: sorted; lo: emission order (indexed by slot)
    const uint32_t* val;
    const uint32_t* raw_k;
    const uint2* span_rq;                   // pre-pass: {reference, query} consumed since the last record head before the span
    const uint8_t* raw_kind;
    const uint32_t* ne_idx;
    const unsigned long long* cig_off;
    const uint32_t* cigar;
    const uint4* meta;
    const TidDev* tids;
    const uint32_t* scalars;
    uint32_t cap, min_len;
    uint32_t *o_start, *o_end, *o_read, *o_op, *o_qpos, *o_seg;
    uint8_t* o_kind;
};

// output row i <- emitted signature `slot` with keys (hi, lo)
__device__ __forceinline__ void gather_one(const GatherParams& P, uint32_t i, uint32_t slot, unsigned long long hi, unsigned long long lo)
{
    {
        const uint32_t g = 0xffffffffu - (uint32_t)lo;
        const uint32_t k = P.raw_k[slot];
        const uint32_t read = P.ne_idx[k];
        const unsigned long long c0 = P.cig_off[read];
        const uint8_t kr = P.raw_kind[slot];
        // query offset of the op (sv_caller.cpp:547,653-655): the pre-pass knows it (and the reference offset) at the
        // start of the op's span; the ops between the span start (or the record head, if later) and g are summed here.
        // Records that reach the end of their contig need the reference's own sequence of steps from there on: a soft
        // clip of >= min_len at pos + 1 >= map_size skips the query advance (sv_caller.cpp:602-604).  Positions never
        // decrease along a record, so the prefix is still valid as long as the record was inside the contig at the
        // span start; only a record that already left it before falls back to its head.
        const uint32_t sp = g / (uint32_t)kWalkSpan;
        const unsigned long long span_op0 = (unsigned long long)sp * kWalkSpan;
        const bool exact = kr & 0x80u;
        const uint4 m = P.meta[k];
        unsigned long long from = c0;
        uint32_t q = 0, pos = m.x;
        if (c0 < span_op0) {
            const uint2 rq = P.span_rq[sp];
            const uint32_t ref_pre = rq.x;
            if (!exact || m.x + ref_pre + 1u < m.y) {
                q = rq.y;
                pos = m.x + ref_pre;
                from = span_op0;
            }
        }
        for (unsigned long long o = from; o < (unsigned long long)g; o++) {
            const uint32_t w = P.cigar[o], op = w & 15u, len = w >> 4;
            if (exact && len >= P.min_len && op == 4u && (uint32_t)(pos + 1u) >= m.y) continue;
            if ((kRefMask >> op) & 1u) pos += len;
            if ((kQryMask >> op) & 1u) q += len;
        }
        const uint32_t qpos = q;
        const uint32_t kind = kr & 0x7fu;
        const uint32_t region = (uint32_t)(hi >> 32);
        P.o_start[i] = (uint32_t)hi;
        P.o_end[i] = (uint32_t)(lo >> 32);
        P.o_kind[i] = (uint8_t)kind;
        P.o_read[i] = read;
        P.o_op[i] = (uint32_t)((unsigned long long)g - c0);
        P.o_qpos[i] = qpos;
        P.o_seg[i] = region * 2u + (kind == 1u ? 0u : 1u);   // group = (region, SVType): DEL | INS
    }
}

__global__ void k_sig_gather(const GatherParams P)
{
    const uint32_t n = P.scalars[SC_N_SIG_EFF];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = P.val[i];
        gather_one(P, i, slot, P.hi[i], P.lo[slot]);
    }
}

// ---- bucket ordering
// raw slot i = index * kSigSub + counter (walk.cu): a lane always meets the same counter (the stride is a multiple of 32)
__global__ void __launch_bounds__(256) k_bucket_scatter(const uint32_t* __restrict__ bucket, const uint32_t* __restrict__ arrival,
                                                         const uint32_t* __restrict__ bucket_base, const uint32_t* scalars, uint32_t cap, uint32_t sub_mask,
                                                         uint32_t* perm, const uint32_t* __restrict__ reg_tab, uint32_t n_regions, uint32_t* reg_sig_cnt)
{
    static_assert(kSigSub == 32, "one slot counter per lane");
    // signatures per region (caller order): the buckets of a region are its depth tiles, so the counts are differences of
    // the bucket scan -- the walk does not count them (one same-address atomic less per signature)
    if (blockIdx.x == 0 && sub_mask != 0)
        for (uint32_t r = threadIdx.x; r < n_regions; r += blockDim.x) reg_sig_cnt[r] = bucket_base[reg_tab[r + 1u]] - bucket_base[reg_tab[r]];
    if (sub_mask == 0) {                                      // one counter, dense slots
        const uint32_t n = min(scalars[SC_SIG_SUB0], cap);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) perm[bucket_base[bucket[i]] + arrival[i]] = i;
        return;
    }
    const uint32_t mine = scalars[SC_SIG_SUB0 + (threadIdx.x & 31u)];
    const uint32_t most = __reduce_max_sync(0xffffffffu, mine);
    const uint64_t end = (uint64_t)most * kSigSub;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (uint64_t)gridDim.x * blockDim.x)
        if ((uint32_t)(i / kSigSub) < mine && i < cap) perm[bucket_base[bucket[i]] + arrival[i]] = (uint32_t)i;
}

// P.hi / P.lo: the keys in emission order; perm: emitted slots bucket by bucket
__global__ void __launch_bounds__(256) k_bucket_rank_gather(const GatherParams P, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ bucket,
                                                             const uint32_t* __restrict__ bucket_base)
{
    const uint32_t n = P.scalars[SC_N_SIG_EFF];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const uint32_t slot = perm[s];
        const uint32_t bk = bucket[slot];
        const uint32_t b0 = bucket_base[bk], b1 = bucket_base[bk + 1u];
        const unsigned long long hi = P.hi[slot], lo = P.lo[slot];
        uint32_t rank = 0;
        for (uint32_t j = b0; j < b1; j++) {
            const uint32_t o = perm[j];
            const unsigned long long h2 = P.hi[o], l2 = P.lo[o];
            rank += (h2 < hi || (h2 == hi && l2 < lo)) ? 1u : 0u;
        }
        gather_one(P, b0 + rank, slot, hi, lo);
    }
}

static GatherParams gather_params(csv_batch* b)
{
    GatherParams P;
    P.hi = b->d_sig_hi.as<unsigned long long>(); P.lo = b->d_sig_lo.as<unsigned long long>(); P.val = nullptr;
    P.raw_k = b->d_sig_k.as<uint32_t>(); P.span_rq = b->d_span_rq.as<uint2>(); P.raw_kind = b->d_sig_kind.as<uint8_t>();
    P.ne_idx = b->d_ne_idx.as<uint32_t>(); P.cig_off = b->d_cig_off.as<unsigned long long>(); P.cigar = b->d_cigar.as<uint32_t>();
    P.meta = b->d_meta.as<uint4>(); P.tids = b->d_tids.as<TidDev>(); P.scalars = b->d_scalars.as<uint32_t>(); P.cap = (uint32_t)b->sig_cap;
    P.o_start = b->d_out_start.as<uint32_t>(); P.o_end = b->d_out_end.as<uint32_t>(); P.o_read = b->d_out_read.as<uint32_t>();
    P.o_op = b->d_out_op.as<uint32_t>(); P.o_qpos = b->d_out_qpos.as<uint32_t>(); P.o_seg = b->d_out_seg.as<uint32_t>();
    P.o_kind = b->d_out_kind.as<uint8_t>();
    P.min_len = b->last_min_len;
    return P;
}

bool sig_order_radix()
{
    static const bool use_radix = getenv("CSV_SIG_ORDER") && !strcmp(getenv("CSV_SIG_ORDER"), "radix");
    return use_radix;
}

int launch_sig_finish(csv_ctx* ctx, csv_batch* b)
{
    const uint32_t cap = (uint32_t)b->sig_cap;
    uint32_t* scalars = b->d_scalars.as<uint32_t>();
    uint32_t grid = cap_grid(ctx, ctx->sm_count * grid_mult(ctx, 4));
    const bool use_radix = sig_order_radix();
    if (!use_radix) {
        // the scan always runs over every bucket: it is what leaves the counters at zero for the next pass
        uint32_t* cnt = b->d_bucket_cnt.as<uint32_t>();
        uint32_t* base = b->d_bucket_base.as<uint32_t>();
        const uint32_t n_tiles = b->n_tiles;
        CSV_TRY(chained_scan(ctx,
                             [=] __device__(uint64_t t) -> uint32_t { return cnt[t]; },
                             [=] __device__(uint64_t t, uint32_t ex, uint32_t v) { base[t] = ex; cnt[t] = 0u; if (t + 1 == n_tiles) base[n_tiles] = ex + v; },
                             n_tiles, nullptr, scalars + SC_N_SIG_EFF));
        uint32_t* perm = b->d_sig_payload.as<uint32_t>();
        k_bucket_scatter<<<grid, 256, 0, ctx->stream>>>(b->d_sig_bucket.as<uint32_t>(), b->d_sig_arrival.as<uint32_t>(), base, scalars, cap, b->sig_sub_mask, perm,
                                                        b->d_reg_tab.as<uint32_t>(), b->n_regions, b->d_reg_sig_cnt.as<uint32_t>());
        const GatherParams G = gather_params(b);
        k_bucket_rank_gather<<<grid, 256, 0, ctx->stream>>>(G, perm, b->d_sig_bucket.as<uint32_t>(), base);
        ctx->launches += 2;
        CSV_CUDA(cudaGetLastError());
        return CSV_OK;
    }
    // radix path: the bucket counters still have to return to zero
    CSV_CUDA(cudaMemsetAsync(b->d_bucket_cnt.p, 0, (size_t)b->n_tiles * 4 + 16, ctx->stream));
    CSV_TRY(ctx->sort_tmp[1].ensure((size_t)cap * 8));
    CSV_TRY(ctx->sort_tmp[3].ensure((size_t)cap * 4));
    SortBufs sb;
    sb.hi = nullptr; sb.hi2 = nullptr;
    sb.lo = b->d_sig_hi.as<unsigned long long>(); sb.lo2 = ctx->sort_tmp[1].as<unsigned long long>();
    sb.val = b->d_sig_payload.as<uint32_t>(); sb.val2 = ctx->sort_tmp[3].as<uint32_t>();
    uint32_t mask = 0x0fu;                                                       // start
    for (int d = 0; d < 4; d++) if (d == 0 ? b->n_regions > 1 : (b->n_regions >> (8 * d))) mask |= 1u << (4 + d);   // owner region
    // the sort's first kernel clamps the emitted count to the capacity (SC_N_SIG_EFF) and numbers the entries
    const SortFirst first = {scalars + SC_SIG_SUB0, cap, scalars + SC_N_SIG_EFF, true};   // one counter, dense slots (csv_batch::sig_sub_mask == 0)
    CSV_TRY(radix_sort_pairs(ctx, sb, cap, scalars + SC_N_SIG_EFF, mask, &first));
    k_sig_tiefix<<<grid, 256, 0, ctx->stream>>>(sb.lo, b->d_sig_lo.as<unsigned long long>(), sb.val, scalars, sb.val2);
    ctx->launches++;
    GatherParams P;
    P.hi = sb.lo; P.lo = b->d_sig_lo.as<unsigned long long>(); P.val = sb.val2;
    P.raw_k = b->d_sig_k.as<uint32_t>(); P.span_rq = b->d_span_rq.as<uint2>(); P.raw_kind = b->d_sig_kind.as<uint8_t>();
    P.ne_idx = b->d_ne_idx.as<uint32_t>(); P.cig_off = b->d_cig_off.as<unsigned long long>(); P.cigar = b->d_cigar.as<uint32_t>();
    P.meta = b->d_meta.as<uint4>(); P.tids = b->d_tids.as<TidDev>(); P.scalars = scalars; P.cap = cap;
    P.o_start = b->d_out_start.as<uint32_t>(); P.o_end = b->d_out_end.as<uint32_t>(); P.o_read = b->d_out_read.as<uint32_t>();
    P.o_op = b->d_out_op.as<uint32_t>(); P.o_qpos = b->d_out_qpos.as<uint32_t>(); P.o_seg = b->d_out_seg.as<uint32_t>();
    P.o_kind = b->d_out_kind.as<uint8_t>();
    P.min_len = b->last_min_len;
    k_sig_gather<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

int launch_sig_dbscan(csv_ctx* ctx, csv_batch* b, double eps, int min_pts)
{
    CSV_TRY(b->d_labels.ensure((size_t)b->sig_cap * 4 + 16));
    static const bool fused = !(getenv("CSV_DB_SIGS") && !strcmp(getenv("CSV_DB_SIGS"), "general"));
    if (fused && eps >= 0.0)
        return dbscan1d_sorted_sigs(ctx, b->d_out_start.as<int32_t>(), b->d_out_seg.as<uint32_t>(), b->sig_cap, b->d_scalars.as<uint32_t>() + SC_N_SIG_EFF,
                                    b->n_regions * 2, eps, min_pts, b->d_labels.as<int32_t>());
    return dbscan1d_device(ctx, b->d_out_start.as<int32_t>(), b->d_out_seg.as<uint32_t>(), b->sig_cap,
                           b->d_scalars.as<uint32_t>() + SC_N_SIG_EFF, b->n_regions * 2, eps, min_pts, b->d_labels.as<int32_t>(), nullptr,
                           true /* the signature list is sorted by (region, start) */);
}

}  // namespace csv

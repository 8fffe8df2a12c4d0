// walk.cu -- the CIGAR walk: one flat, op-parallel pass over every CIGAR word of
// the batch (sv_caller.cpp:539-661 and cnv_caller.cpp:488-531 at once).
//
// Work is cut into spans of kWalkSpan consecutive ops regardless of record
// boundaries, so HiFi (64 ops/record) and ONT (1e4+ ops/record) batches are
// equally balanced.  Per span:
//   1. each thread loads 8 ops with two 128-bit loads plus the 9 head bits that
//      say where records start;
//   2. a segmented scan (reset at record heads) of (reference, query)
//      consumption runs thread -> warp -> CTA; the carry-in of every span comes
//      from a cheap aggregate pre-pass (k_span_agg) and a two-level scan over the
//      span aggregates -- deterministic, no spinning on other CTAs.  Plain sums of
//      head bits (-> record index) and of depth-event counts (-> event slot) ride
//      along;
//   3. every op now knows its reference position, query offset and event slot:
//        - I/D/S ops >= min_len become signatures;
//        - depth events are written in op order: a record contributes +1 at its
//          first index, -1/+1 around every D/N gap, -1 one past its last base --
//          exactly the bases its M/=/X ops cover (cnv_caller.cpp:507-519).
//          Every record writes an EVEN number of events that alternate +,-,+,-...
//          so the sign of an event is the parity of its slot: the event list is
//          a plain array of uint32 depth-map indices (kNone = clipped/filtered),
//          grouped by record, hence sorted by contig and (nearly) by position.
// Per record the walk also leaves ev_start[k] (first event slot) and
// ref_end[k] (one past the last covered index, 0 if filtered) for the tile kernel.
#include "batch.cuh"
#include "scan.cuh"

namespace csv {

struct WalkParams {
    const uint32_t* cigar;
    uint32_t n_ops;
    const uint8_t* headbits;
    const uint4* meta;          // {pos0, map_size (0 = contig not requested), flag | mapq << 16 | depth-live << 30 | sig-ok << 31, owner region}
    WalkAgg* span_agg;          // per-span aggregate
    WalkAgg* span_pre;          // exclusive prefix of the span inside its chunk of kSpanChunk spans
    WalkAgg* chunk_agg;         // per-chunk aggregate, then exclusive prefix over chunks
    uint32_t n_spans;
    uint32_t* events;
    uint32_t ev_cap;
    uint32_t* ev_start;     // [n_nonempty + 1]
    uint32_t* ref_total;    // [n_nonempty] reference bases consumed by the record
    uint32_t n_meta;        // entries allocated in meta
    uint32_t min_len, min_mapq;
    uint32_t* scalars;
    SigRaw sig;
    uint32_t sig_cap;
    uint32_t* reg_sig_cnt;
    int want_depth, want_sigs;
};

constexpr int kSpanChunk = 2048;    // spans per scan chunk (256 threads x 8)

__device__ __forceinline__ WalkAgg combine(const WalkAgg a, const WalkAgg b)
{
    WalkAgg r;
    r.heads = a.heads + b.heads;
    r.ref = b.heads ? b.ref : a.ref + b.ref;
    r.qry = b.heads ? b.qry : a.qry + b.qry;
    r.ev = a.ev + b.ev;
    return r;
}

__device__ __forceinline__ WalkAgg warp_incl_scan_agg(WalkAgg inc, uint32_t lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        WalkAgg t;
        t.heads = __shfl_up_sync(0xffffffffu, inc.heads, d);
        t.ref = __shfl_up_sync(0xffffffffu, inc.ref, d);
        t.qry = __shfl_up_sync(0xffffffffu, inc.qry, d);
        t.ev = __shfl_up_sync(0xffffffffu, inc.ev, d);
        if (lane >= (uint32_t)d) inc = combine(t, inc);
    }
    return inc;
}
__device__ __forceinline__ WalkAgg shfl_up1_agg(const WalkAgg inc, uint32_t lane)
{
    WalkAgg e;
    e.heads = __shfl_up_sync(0xffffffffu, inc.heads, 1);
    e.ref = __shfl_up_sync(0xffffffffu, inc.ref, 1);
    e.qry = __shfl_up_sync(0xffffffffu, inc.qry, 1);
    e.ev = __shfl_up_sync(0xffffffffu, inc.ev, 1);
    if (lane == 0) e = WalkAgg{0, 0, 0, 0};
    return e;
}

// 8 ops of one thread + the 9 head bits around them
struct ThreadOps { uint32_t w[kWalkOpsPerThread]; uint32_t hb, n_valid; };

__device__ __forceinline__ ThreadOps load_ops(const uint32_t* __restrict__ cigar, const uint8_t* __restrict__ headbits, uint32_t n_ops, uint32_t g0)
{
    ThreadOps t;
    t.hb = 0;
    if (g0 + kWalkOpsPerThread <= n_ops) {
        uint4 a = ld_nc_v4(cigar + g0), c = ld_nc_v4(cigar + g0 + 4);
        t.w[0] = a.x; t.w[1] = a.y; t.w[2] = a.z; t.w[3] = a.w; t.w[4] = c.x; t.w[5] = c.y; t.w[6] = c.z; t.w[7] = c.w;
    } else {
#pragma unroll
        for (int j = 0; j < kWalkOpsPerThread; j++) t.w[j] = (g0 + j < n_ops) ? __ldg(cigar + g0 + j) : 0u;
    }
    if (g0 < n_ops) t.hb = (uint32_t)__ldg(headbits + (g0 >> 3)) | ((uint32_t)__ldg(headbits + (g0 >> 3) + 1) << 8);
    t.n_valid = g0 >= n_ops ? 0u : (n_ops - g0 < kWalkOpsPerThread ? n_ops - g0 : kWalkOpsPerThread);
    return t;
}

__device__ __forceinline__ uint32_t bit_of(uint32_t mask, uint32_t op) { return (mask >> op) & 1u; }

// Aggregate of the 8 ops of one thread.  Heads / event counts are popcounts of the head bits;
// (ref, qry) only count the ops from the LAST record head of the thread onwards.
__device__ __forceinline__ WalkAgg thread_aggregate(const ThreadOps& t)
{
    WalkAgg a;
    const uint32_t vmask = (1u << t.n_valid) - 1u;          // n_valid <= 8
    const uint32_t hbv = t.hb & vmask;
    a.heads = __popc(hbv);
    const int lh = 31 - __clz(hbv);                         // index of the last head, -1 if none
    uint32_t ref = 0, qry = 0, gaps = 0;
#pragma unroll
    for (int j = 0; j < kWalkOpsPerThread; j++) {
        const uint32_t op = t.w[j] & 15u;                   // ops beyond n_valid are zero words: M of length 0
        const uint32_t len = (j >= lh) ? (t.w[j] >> 4) : 0u;
        ref += bit_of(kRefMask, op) * len;
        qry += bit_of(kQryMask, op) * len;
        gaps += bit_of(kGapMask, op) & (uint32_t)((t.w[j] >> 4) != 0u);
    }
    a.ref = ref; a.qry = qry;
    a.ev = a.heads + __popc((t.hb >> 1) & vmask) + 2u * gaps;
    return a;
}

// pre-pass: aggregate of every span (one CTA per span; pure streaming read of the CIGAR words)
__global__ void __launch_bounds__(kWalkThreads) k_span_agg(const WalkParams P)
{
    __shared__ WalkAgg s_warp[kWalkThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t span = blockIdx.x;
    const ThreadOps t = load_ops(P.cigar, P.headbits, P.n_ops, span * (uint32_t)kWalkSpan + tid * kWalkOpsPerThread);
    const WalkAgg a = thread_aggregate(t);
    // warp aggregate with the redux unit: plain sums for heads / events; (ref, qry) count from the last lane
    // that saw a record head (that lane's own values already start at its last head)
    WalkAgg w;
    w.heads = __reduce_add_sync(0xffffffffu, a.heads);
    w.ev = __reduce_add_sync(0xffffffffu, a.ev);
    const uint32_t hm = __ballot_sync(0xffffffffu, a.heads != 0u);
    const bool counts = hm == 0u || lane >= (uint32_t)(31 - __clz(hm));
    w.ref = __reduce_add_sync(0xffffffffu, counts ? a.ref : 0u);
    w.qry = __reduce_add_sync(0xffffffffu, counts ? a.qry : 0u);
    if (lane == 0) s_warp[warp] = w;
    __syncthreads();
    if (tid == 0) {
        WalkAgg total = s_warp[0];
#pragma unroll
        for (int i = 1; i < kWalkThreads / 32; i++) total = combine(total, s_warp[i]);
        P.span_agg[span] = total;
    }
}

// level 1: exclusive segmented scan of the span aggregates inside chunks of kSpanChunk spans
__global__ void __launch_bounds__(256) k_span_scan_local(const WalkParams P)
{
    __shared__ WalkAgg s_warp[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t base = blockIdx.x * (uint32_t)kSpanChunk + tid * 8u;
    WalkAgg v[8], run = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; j++) { v[j] = (base + j < P.n_spans) ? P.span_agg[base + j] : WalkAgg{0, 0, 0, 0}; run = combine(run, v[j]); }
    const WalkAgg inc = warp_incl_scan_agg(run, lane);
    WalkAgg pre = shfl_up1_agg(inc, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    WalkAgg wpre = {0, 0, 0, 0};
    for (uint32_t i = 0; i < warp; i++) wpre = combine(wpre, s_warp[i]);
    pre = combine(wpre, pre);
#pragma unroll
    for (int j = 0; j < 8; j++) { if (base + j < P.n_spans) P.span_pre[base + j] = pre; pre = combine(pre, v[j]); }
    if (tid == 255) P.chunk_agg[blockIdx.x] = pre;
}

// level 2: one CTA turns the chunk aggregates into exclusive prefixes (in place)
__global__ void __launch_bounds__(1024) k_span_scan_chunks(const WalkParams P, uint32_t n_chunks)
{
    __shared__ WalkAgg s_warp[32];
    __shared__ WalkAgg s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = WalkAgg{0, 0, 0, 0};
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_chunks; b0 += 1024) {
        const uint32_t i = b0 + tid;
        const WalkAgg v = i < n_chunks ? P.chunk_agg[i] : WalkAgg{0, 0, 0, 0};
        const WalkAgg inc = warp_incl_scan_agg(v, lane);
        WalkAgg pre = shfl_up1_agg(inc, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        WalkAgg wpre = s_carry;
        for (uint32_t w = 0; w < warp; w++) wpre = combine(wpre, s_warp[w]);
        pre = combine(wpre, pre);
        if (i < n_chunks) P.chunk_agg[i] = pre;
        __syncthreads();
        if (tid == 1023) s_carry = combine(pre, v);
        __syncthreads();
    }
}

constexpr int kMetaStage = 128;     // record metadata of a span staged in shared memory (typical span: ~35 records)

// Replay of one thread's ops with their full prefixes.  FULL = all 8 ops valid (every span but the last).
// Only the per-op work lives here: D/N gap events and (rarely) signatures.  The two events every record
// owes at its first and last index are written by k_record_events from (ev_start, ref_total), so a record
// head / tail costs this loop a handful of instructions instead of a divergent block.
template <bool DEPTH, bool SIGS, bool FULL>
__device__ __forceinline__ void walk_replay(const WalkParams& P, const ThreadOps& t, const WalkAgg T, uint32_t g0,
                                            const uint4* s_meta, uint32_t k_first)
{
    uint32_t k = T.heads - 1u, Rc = T.ref, Qc = T.qry, slot = T.ev;
    auto fetch = [&](uint32_t kk) -> uint4 {
        const uint32_t l = kk - k_first;
        return l < (uint32_t)kMetaStage ? s_meta[l] : __ldg(P.meta + kk);
    };
    uint4 m = make_uint4(0, 0, 0, kNone);
    if (!(t.hb & 1u) && (FULL || t.n_valid)) m = fetch(k);                   // first op continues an earlier record
#pragma unroll
    for (int j = 0; j < kWalkOpsPerThread; j++) {
        if (!FULL && (uint32_t)j >= t.n_valid) break;
        const uint32_t op = t.w[j] & 15u, len = t.w[j] >> 4;
        if ((t.hb >> j) & 1u) {                                              // first op of a record
            k++; Rc = 0; Qc = 0;
            m = fetch(k);
            if (DEPTH) slot++;                                               // the record's +1 event (k_record_events)
        }
        const uint32_t pos1 = m.x + Rc + 1u;                                 // reference's `pos + 1` at this op (uint32)
        if (SIGS && len >= P.min_len && bit_of(kSigMask, op) && (m.z >> 31)) {
            const bool beyond = pos1 >= m.y;
            if (!(op == 4 && beyond)) {                                      // sv_caller.cpp:602-604
                const uint32_t start = pos1, end = start + len - 1u;
                if (start <= end) {                                          // sv_object.cpp:25-28
                    const uint32_t sl = atomicAdd(&P.scalars[SC_N_SIG], 1u);
                    if (sl < P.sig_cap) {
                        P.sig.key_hi[sl] = ((unsigned long long)m.w << 32) | start;
                        P.sig.key_lo[sl] = ((unsigned long long)end << 32) | (0xffffffffu - (g0 + j));
                        P.sig.k[sl] = k;
                        P.sig.qpos[sl] = Qc;
                        const uint32_t kind = op == 1 ? 0u : (op == 2 ? 1u : 2u);
                        P.sig.kind[sl] = (uint8_t)(kind | ((beyond || (int32_t)m.x < 0) ? 0x80u : 0u));
                        atomicAdd(&P.reg_sig_cnt[m.w], 1u);
                    }
                }
            }
        }
        if (DEPTH) {                                                         // D / N: -1 at its first index, +1 one past its last
            const bool gap = bit_of(kGapMask, op) && len;
            const uint32_t ia = pos1, ib = pos1 + len;                       // no overflow when ia < map_size <= 2^31
            const bool in = ((m.z >> 30) & 1u) && ia < m.y;
            const uint32_t va = in ? ia : kNone, vb = (in && ib < m.y) ? ib : kNone;
            if (gap) { P.events[slot] = va; P.events[slot + 1] = vb; }
            slot += gap ? 2u : 0u;
        }
        Rc += bit_of(kRefMask, op) * len;
        Qc += bit_of(kQryMask, op) * len;
        if (DEPTH && ((t.hb >> (j + 1)) & 1u)) {                             // last op of the record
            slot++;                                                          // the record's -1 event (k_record_events)
            P.ev_start[k + 1] = slot;
            P.ref_total[k] = Rc;
        }
    }
}

template <bool DEPTH, bool SIGS>
__global__ void __launch_bounds__(kWalkThreads, 4) k_walk(const WalkParams P)
{
    __shared__ WalkAgg s_warp[kWalkThreads / 32];
    __shared__ uint4 s_meta[kMetaStage];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t span = blockIdx.x;
    const uint32_t g0 = span * (uint32_t)kWalkSpan + tid * kWalkOpsPerThread;
    // independent loads first: ops, head bits, the span's carry-in and the metadata of the span's records
    const ThreadOps t = load_ops(P.cigar, P.headbits, P.n_ops, g0);
    const WalkAgg span_excl = combine(P.chunk_agg[span / kSpanChunk], P.span_pre[span]);
    const uint32_t k_first = span_excl.heads - 1u;                           // record running into this span (may be -1)
    if (tid < (uint32_t)kMetaStage) {
        const uint32_t kk = k_first + tid;
        s_meta[tid] = (kk < P.n_meta) ? __ldg(P.meta + kk) : make_uint4(0, 0, 0, kNone);
    }
    const WalkAgg inc = warp_incl_scan_agg(thread_aggregate(t), lane);
    const WalkAgg lane_excl = shfl_up1_agg(inc, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    WalkAgg wpre = {0, 0, 0, 0};
    for (uint32_t i = 0; i < warp; i++) wpre = combine(wpre, s_warp[i]);
    const WalkAgg T = combine(span_excl, combine(wpre, lane_excl));
    if (t.n_valid == kWalkOpsPerThread) walk_replay<DEPTH, SIGS, true>(P, t, T, g0, s_meta, k_first);
    else walk_replay<DEPTH, SIGS, false>(P, t, T, g0, s_meta, k_first);
}

// The two events every record owes: +1 at its first index, -1 one past its last covered base; plus ref_end,
// the input of the prefix-max that finds the records overlapping a tile.
__global__ void __launch_bounds__(256) k_record_events(const uint4* __restrict__ meta, const uint32_t* __restrict__ ev_start,
                                                       const uint32_t* __restrict__ ref_total, const uint32_t* scalars,
                                                       uint32_t* events, uint32_t* ref_end)
{
    const uint32_t n = scalars[SC_N_NONEMPTY];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint4 m = __ldg(meta + k);
        const uint32_t a0 = m.x + 1u;                                        // (uint32)pos + 1, cnv_caller.cpp:499
        const bool live = ((m.z >> 30) & 1u) && a0 < m.y;
        const uint32_t ie = a0 + ref_total[k];
        const bool in = live && ie >= a0 && ie < m.y;                        // ie < a0: 32-bit wrap (absurd record)
        events[ev_start[k]] = live ? a0 : kNone;
        events[ev_start[k + 1] - 1u] = in ? ie : kNone;
        ref_end[k] = live ? (in ? ie : m.y) : 0u;
    }
}

int launch_walk(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p)
{
    if (b->n_ops == 0) return CSV_OK;
    WalkParams P;
    P.cigar = b->d_cigar.as<uint32_t>();
    P.n_ops = (uint32_t)b->n_ops;
    P.headbits = b->d_headbits.as<uint8_t>();
    P.meta = b->d_meta.as<uint4>();
    P.span_agg = b->d_span_agg.as<WalkAgg>();
    P.span_pre = b->d_span_pre.as<WalkAgg>();
    P.chunk_agg = b->d_span_status.as<WalkAgg>();
    P.n_spans = b->n_spans;
    P.events = b->d_events.as<uint32_t>();
    P.ev_cap = (uint32_t)b->ev_cap;
    P.ev_start = b->d_ev_start.as<uint32_t>();
    P.ref_total = b->d_ref_total.as<uint32_t>();
    P.n_meta = b->n_reads;
    P.min_len = p->min_len; P.min_mapq = p->min_mapq;
    P.scalars = b->d_scalars.as<uint32_t>();
    P.sig.key_hi = b->d_sig_hi.as<unsigned long long>();
    P.sig.key_lo = b->d_sig_lo.as<unsigned long long>();
    P.sig.k = b->d_sig_k.as<uint32_t>();
    P.sig.qpos = b->d_sig_qpos.as<uint32_t>();
    P.sig.kind = b->d_sig_kind.as<uint8_t>();
    P.sig_cap = (uint32_t)b->sig_cap;
    P.reg_sig_cnt = b->d_reg_sig_cnt.as<uint32_t>();
    P.want_depth = p->want_depth; P.want_sigs = p->want_sigs;
    const uint32_t n_chunks = (b->n_spans + kSpanChunk - 1) / kSpanChunk;
    k_span_agg<<<b->n_spans, kWalkThreads, 0, ctx->stream>>>(P);
    k_span_scan_local<<<n_chunks, 256, 0, ctx->stream>>>(P);
    k_span_scan_chunks<<<1, 1024, 0, ctx->stream>>>(P, n_chunks);
    if (p->want_depth && p->want_sigs) k_walk<true, true><<<b->n_spans, kWalkThreads, 0, ctx->stream>>>(P);
    else if (p->want_depth) k_walk<true, false><<<b->n_spans, kWalkThreads, 0, ctx->stream>>>(P);
    else k_walk<false, true><<<b->n_spans, kWalkThreads, 0, ctx->stream>>>(P);
    ctx->launches += 4;
    if (p->want_depth) {
        const uint32_t grid = (b->n_reads + 255) / 256 < (uint32_t)ctx->sm_count * 16 ? (b->n_reads + 255) / 256 : (uint32_t)ctx->sm_count * 16;
        k_record_events<<<grid, 256, 0, ctx->stream>>>(P.meta, P.ev_start, P.ref_total, P.scalars, P.events, b->d_ref_end.as<uint32_t>());
        ctx->launches++;
    }
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

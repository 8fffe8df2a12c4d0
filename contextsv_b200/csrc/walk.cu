// walk.cu -- the CIGAR walk: one flat, op-parallel pass over every CIGAR word of
// the batch (sv_caller.cpp:539-661 and cnv_caller.cpp:488-531 at once).
//
// Work is cut into spans of kWalkSpan consecutive ops regardless of record
// boundaries, so HiFi (64 ops/record) and ONT (1e4+ ops/record) batches are
// equally balanced.  Per span:
//   1. each thread loads 8 ops with two 128-bit loads plus the 9 head bits that
//      say where records start;
//   2. a segmented scan (reset at record heads) of (reference, query)
//      consumption runs thread -> warp -> CTA -> chained look-back across spans;
//      plain sums of head bits (-> record index) and of depth-event counts
//      (-> event slot) ride along;
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
    const uint4* meta;
    const TidDev* tids;
    WalkAgg* span_agg;
    WalkAgg* span_pre;
    uint32_t* span_status;
    uint32_t* ticket;
    uint32_t epoch;
    uint32_t n_spans;
    uint32_t* events;
    uint32_t ev_cap;
    uint32_t* ev_start;     // [n_nonempty + 1]
    uint32_t* ref_end;      // [n_nonempty]
    uint32_t min_len, min_mapq;
    uint32_t* scalars;
    SigRaw sig;
    uint32_t sig_cap;
    uint32_t* reg_sig_cnt;
    int want_depth, want_sigs;
};

__device__ __forceinline__ WalkAgg combine(const WalkAgg a, const WalkAgg b)
{
    WalkAgg r;
    r.heads = a.heads + b.heads;
    r.ref = b.heads ? b.ref : a.ref + b.ref;
    r.qry = b.heads ? b.qry : a.qry + b.qry;
    r.ev = a.ev + b.ev;
    return r;
}

__device__ __forceinline__ WalkAgg shfl_agg(const WalkAgg v, int src)
{
    WalkAgg r;
    r.heads = __shfl_sync(0xffffffffu, v.heads, src);
    r.ref = __shfl_sync(0xffffffffu, v.ref, src);
    r.qry = __shfl_sync(0xffffffffu, v.qry, src);
    r.ev = __shfl_sync(0xffffffffu, v.ev, src);
    return r;
}

// Chained look-back over spans for the 4-word state.  Payloads live in span_agg /
// span_pre; span_status carries (epoch << 2 | flag) and is published after a
// __threadfence().  Called by warp 0; returns the exclusive state of span s.
__device__ __forceinline__ WalkAgg walk_lookback(const WalkParams& P, uint32_t s, const WalkAgg total)
{
    const uint32_t lane = lane_id();
    const uint32_t tag = P.epoch << 2;
    WalkAgg excl = {0, 0, 0, 0};
    if (lane == 0) {
        if (s == 0) { P.span_pre[0] = total; __threadfence(); st_volatile_u32(&P.span_status[0], tag | kFlagPrefix); }
        else { P.span_agg[s] = total; __threadfence(); st_volatile_u32(&P.span_status[s], tag | kFlagAgg); }
    }
    if (s == 0) return excl;
    int64_t look = (int64_t)s - 1;
    bool have = false;   // excl currently holds the combination of spans (look, s)
    for (;;) {
        int64_t idx = look - lane;
        uint32_t flag;
        do {
            if (idx >= 0) { uint32_t w = ld_volatile_u32(&P.span_status[idx]); flag = ((w >> 2) == P.epoch) ? (w & 3u) : 0u; }
            else flag = kFlagPrefix;
        } while (__any_sync(0xffffffffu, flag == 0));
        __threadfence();
        WalkAgg v = {0, 0, 0, 0};
        if (idx >= 0) v = (flag == kFlagPrefix) ? P.span_pre[idx] : P.span_agg[idx];
        uint32_t pm = __ballot_sync(0xffffffffu, flag == kFlagPrefix);
        int last = pm ? (__ffs(pm) - 1) : 31;     // farthest lane that contributes
        WalkAgg acc = shfl_agg(v, last);
        for (int i = last - 1; i >= 0; i--) acc = combine(acc, shfl_agg(v, i));
        excl = have ? combine(acc, excl) : acc;
        have = true;
        if (pm) break;
        look -= 32;
    }
    if (lane == 0) { P.span_pre[s] = combine(excl, total); __threadfence(); st_volatile_u32(&P.span_status[s], tag | kFlagPrefix); }
    return excl;
}

__global__ void __launch_bounds__(kWalkThreads, 4) k_walk(const WalkParams P)
{
    __shared__ WalkAgg s_warp[kWalkThreads / 32];
    __shared__ WalkAgg s_excl;
    __shared__ uint32_t s_span;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_span = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const uint32_t span = s_span;
        if (span >= P.n_spans) break;
        const uint32_t g0 = span * (uint32_t)kWalkSpan + tid * kWalkOpsPerThread;

        // ---- 1. load 8 ops + head bits
        uint32_t w[kWalkOpsPerThread];
        uint32_t hb = 0;
        if (g0 + kWalkOpsPerThread <= P.n_ops) {
            uint4 a = ld_nc_v4(P.cigar + g0), c = ld_nc_v4(P.cigar + g0 + 4);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = c.x; w[5] = c.y; w[6] = c.z; w[7] = c.w;
        } else {
#pragma unroll
            for (int j = 0; j < kWalkOpsPerThread; j++) w[j] = (g0 + j < P.n_ops) ? __ldg(P.cigar + g0 + j) : 0u;
        }
        if (g0 < P.n_ops) hb = (uint32_t)P.headbits[g0 >> 3] | ((uint32_t)P.headbits[(g0 >> 3) + 1] << 8);
        const uint32_t n_valid = g0 >= P.n_ops ? 0u : (P.n_ops - g0 < kWalkOpsPerThread ? P.n_ops - g0 : kWalkOpsPerThread);

        // ---- 2a. thread aggregate
        WalkAgg a = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < kWalkOpsPerThread; j++) {
            if ((uint32_t)j < n_valid) {
                const uint32_t op = w[j] & 15u, len = w[j] >> 4;
                if ((hb >> j) & 1u) { a.heads++; a.ref = 0; a.qry = 0; }
                a.ref += ((kRefMask >> op) & 1u) ? len : 0u;
                a.qry += ((kQryMask >> op) & 1u) ? len : 0u;
                a.ev += ((hb >> j) & 1u) + ((hb >> (j + 1)) & 1u) + ((((kGapMask >> op) & 1u) && len) ? 2u : 0u);
            }
        }
        // ---- 2b. warp segmented inclusive scan
        WalkAgg inc = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            WalkAgg t;
            t.heads = __shfl_up_sync(0xffffffffu, inc.heads, d);
            t.ref = __shfl_up_sync(0xffffffffu, inc.ref, d);
            t.qry = __shfl_up_sync(0xffffffffu, inc.qry, d);
            t.ev = __shfl_up_sync(0xffffffffu, inc.ev, d);
            if (lane >= (uint32_t)d) inc = combine(t, inc);
        }
        WalkAgg lane_excl;
        lane_excl.heads = __shfl_up_sync(0xffffffffu, inc.heads, 1);
        lane_excl.ref = __shfl_up_sync(0xffffffffu, inc.ref, 1);
        lane_excl.qry = __shfl_up_sync(0xffffffffu, inc.qry, 1);
        lane_excl.ev = __shfl_up_sync(0xffffffffu, inc.ev, 1);
        if (lane == 0) lane_excl = WalkAgg{0, 0, 0, 0};
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        // ---- 2c. CTA level + chained look-back
        WalkAgg wpre = {0, 0, 0, 0};
        for (uint32_t i = 0; i < warp; i++) wpre = combine(wpre, s_warp[i]);
        if (warp == 0) {
            WalkAgg total = s_warp[0];
#pragma unroll
            for (int i = 1; i < kWalkThreads / 32; i++) total = combine(total, s_warp[i]);
            const WalkAgg ex = walk_lookback(P, span, total);
            if (lane == 0) s_excl = ex;
        }
        __syncthreads();
        const WalkAgg T = combine(s_excl, combine(wpre, lane_excl));

        // ---- 3. replay with full prefixes
        uint32_t Hc = T.heads, Rc = T.ref, Qc = T.qry, slot = T.ev;
        uint32_t kcur = kNone;
        uint4 m = make_uint4(0, 0, 0x80000000u, kNone);
        uint32_t map_size = 0;
        bool dok = false, sok = false;
#pragma unroll
        for (int j = 0; j < kWalkOpsPerThread; j++) {
            if ((uint32_t)j >= n_valid) break;
            const uint32_t op = w[j] & 15u, len = w[j] >> 4;
            const bool head = (hb >> j) & 1u, tail = (hb >> (j + 1)) & 1u;
            if (head) { Hc++; Rc = 0; Qc = 0; }
            const uint32_t k = Hc - 1u;
            if (k != kcur) {
                kcur = k; m = __ldg(P.meta + k);
                const bool ignored = m.z >> 31;
                map_size = ignored ? 0u : P.tids[m.y].map_size;
                const uint32_t flags = m.z & 0xffffu, mq = (m.z >> 16) & 0xffu;
                dok = !ignored && !(flags & kDepthSkipFlags);
                sok = !ignored && !(flags & kSigSkipFlags) && mq >= P.min_mapq && m.w != kNone;
            }
            const uint32_t rl = ((kRefMask >> op) & 1u) ? len : 0u;
            const uint32_t pos = m.x + Rc;   // 0-based position at this op, uint32 like the reference's `pos`

            if (P.want_sigs && sok && len >= P.min_len && ((kSigMask >> op) & 1u)) {
                const bool beyond = (pos + 1u) >= map_size;
                if (!(op == 4 && beyond)) {                                     // sv_caller.cpp:602-604
                    const uint32_t start = pos + 1u, end = start + len - 1u;
                    if (start <= end) {                                         // sv_object.cpp:25-28
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
            if (P.want_depth) {
                const uint32_t a0 = m.x + 1u;                                   // (uint32)pos + 1, cnv_caller.cpp:499
                const bool live = dok && a0 < map_size;
                if (head) {
                    if (slot < P.ev_cap) P.events[slot] = live ? a0 : kNone;
                    slot++;
                }
                if (((kGapMask >> op) & 1u) && len) {
                    const unsigned long long ia = (unsigned long long)(uint32_t)(pos + 1u), ib = ia + len;
                    if (slot + 1 < P.ev_cap) {
                        P.events[slot] = (live && ia < map_size) ? (uint32_t)ia : kNone;
                        P.events[slot + 1] = (live && ib < map_size) ? (uint32_t)ib : kNone;
                    }
                    slot += 2;
                }
                if (tail) {
                    const unsigned long long ie = (unsigned long long)a0 + Rc + rl;
                    if (slot < P.ev_cap) P.events[slot] = (live && ie < map_size) ? (uint32_t)ie : kNone;
                    slot++;
                    P.ev_start[k + 1] = slot;
                    P.ref_end[k] = live ? (uint32_t)(ie < map_size ? ie : map_size) : 0u;
                }
            }
            Rc += rl;
            Qc += ((kQryMask >> op) & 1u) ? len : 0u;
        }
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
    P.tids = b->d_tids.as<TidDev>();
    P.span_agg = b->d_span_agg.as<WalkAgg>();
    P.span_pre = b->d_span_pre.as<WalkAgg>();
    P.span_status = b->d_span_status.as<uint32_t>();
    CSV_TRY(next_ticket(ctx, &P.ticket));
    P.epoch = next_epoch(ctx);
    P.n_spans = b->n_spans;
    P.events = b->d_events.as<uint32_t>();
    P.ev_cap = (uint32_t)b->ev_cap;
    P.ev_start = b->d_ev_start.as<uint32_t>();
    P.ref_end = b->d_ref_end.as<uint32_t>();
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
    uint32_t grid = b->n_spans < (uint32_t)ctx->sm_count * 4 ? b->n_spans : (uint32_t)ctx->sm_count * 4;
    k_walk<<<grid, kWalkThreads, 0, ctx->stream>>>(P);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

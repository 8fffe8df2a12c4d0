// walk.cu -- the CIGAR walk: one flat, op-parallel pass over every CIGAR word of
// the batch (sv_caller.cpp:539-661 and cnv_caller.cpp:488-531 at once).
//
// Work is cut into spans of kWalkSpan consecutive ops regardless of record
// boundaries, so HiFi (64 ops/record) and ONT (1e4+ ops/record) batches are
// equally balanced.  A cheap pre-pass (k_span_agg) reduces every span to one
// aggregate -- record heads, depth events, reference / query consumed since the
// last record head -- and a two-level scan turns the aggregates into the carry-in
// of every span: deterministic, no spinning on other CTAs.  k_walk then gives
// every op its reference position and event slot (see the comment above it):
//   - I/D/S ops >= min_len become signatures (query_pos is filled in by
//     k_sig_gather from the pre-pass's query prefix);
//   - depth events are written in op order: a record contributes +1 at its first
//     index, -1/+1 around every D/N gap, -1 one past its last base -- exactly the
//     bases its M/=/X ops cover (cnv_caller.cpp:507-519).  Every record writes an
//     EVEN number of events that alternate +,-,+,-... so the sign of an event is
//     the parity of its slot: the event list is a plain array of uint32 depth-map
//     indices, grouped by record, hence sorted by contig and (nearly) by position.
//     Events are not clipped to the map: a tile ignores what lies outside it.
// Per record the walk also leaves ev_start[k] (first event slot) and ref_end[k]
// (one past the last covered index, 0 if the record takes no part) for the tiles.
#include "batch.cuh"
#include "scan.cuh"

#include <cstdlib>

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
    uint32_t span_base;         // first span of this launch (the scan is pipelined in chunks of spans)
    uint32_t span_end;          // one past the last span of this launch
    uint4* span_desc;           // per span {record running in (may be -1), first event slot, ref consumed since the last head, records touched}
    uint2* span_rq;             // per span {reference, query} consumed since the last record head before the span (k_sig_gather)
    // record-level pre-pass (host supplied the D/N count of every record, csv_reads::n_gap)
    const uint32_t* n_gap;      // [n_reads] or null
    const unsigned long long* cig_off;
    const uint32_t* ne_idx;     // compact index -> record index
    uint32_t ev_given;          // ev_start[] comes from the record scan (csv_reads::n_gap): the walk leaves what it finds in ev_check[]
    uint32_t* ev_check;         // [n_nonempty + 1] event slot reached at the end of every record (k_ev_check compares after the walk)
    uint32_t* events;
    uint32_t ev_cap;
    uint32_t* ev_start;     // [n_nonempty + 1] first event slot of each record
    uint32_t* ref_end;      // [n_nonempty] one past the last covered index (may lie beyond the map); 0 = takes no part in the depth
    uint32_t n_meta;        // entries allocated in meta
    uint32_t min_len, thr;   // thr = min_len << 4 (0xffffffff when no length can reach min_len)
    uint32_t* scalars;
    SigRaw sig;
    uint32_t sig_cap;
    uint32_t* reg_sig_cnt;
    const uint32_t* reg_tab;    // [tile_base (n_regions + 1) | len (n_regions) | beg (n_regions)], caller order
    uint32_t n_regions;
    uint32_t* bucket_cnt;       // signatures per depth tile
    uint32_t sig_sub_mask;      // kSigSub - 1: slot counters in use (0: one counter, dense slots -- the radix order needs them dense)
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

// Class bit of a CIGAR word without isolating its op nibble: the per-op bit masks are replicated in both
// 16-bit halves, so a funnel shift by (w & 31) lands on the right bit whatever the length's low bit is.
constexpr uint32_t kRefLut = kRefMask | (kRefMask << 16);
constexpr uint32_t kGapLut = kGapMask | (kGapMask << 16);
constexpr uint32_t kSigLut = kSigMask | (kSigMask << 16);
__device__ __forceinline__ uint32_t class_bit(uint32_t lut, uint32_t w) { return __funnelshift_r(lut, lut, w) & 1u; }
// ... and the same bit delivered at position J of the result (0 elsewhere): the table is rotated left by J at compile time,
// so the funnel shift lands the op's bit on J -- one SHF + one LOP3 per op and mask instead of SHF, AND, shift, OR
template <int J> __device__ __forceinline__ uint32_t class_bit_at(uint32_t lut, uint32_t w)
{
    const uint32_t r = J ? ((lut << J) | (lut >> (32 - J))) : lut;
    return __funnelshift_r(r, r, w) & (1u << J);
}

template <int V> struct IntC { static constexpr int value = V; };

// one step of an inclusive warp scan: SHFL + predicated add
__device__ __forceinline__ uint32_t scan_step_u32(uint32_t x, int d)
{
    asm volatile("{\n\t.reg .u32 t;\n\t.reg .pred p;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.u32 %0, %0, t;\n\t}" : "+r"(x) : "r"(d));
    return x;
}
__device__ __forceinline__ uint32_t warp_incl_scan_fast(uint32_t x)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) x = scan_step_u32(x, d);
    return x;
}

// Per-op class table in shared memory: {consumes reference, consumes query} as 0/1 multipliers.  One LDS.64 per op
// replaces two funnel-shift + mask pairs: the pre-pass is bound by the ALU pipe, the multipliers feed IMADs on the
// FMA pipe instead.
__device__ __forceinline__ void init_class_table(uint2* s_cls)
{
    if (threadIdx.x < 16) s_cls[threadIdx.x] = make_uint2((kRefMask >> threadIdx.x) & 1u, (kQryMask >> threadIdx.x) & 1u);
}

// Aggregate of the 8 ops of one thread for the pre-pass.  Heads / event counts are popcounts of the head
// bits; (ref, qry) only count the ops from the LAST record head of the thread onwards.  Every D/N op owns two
// events, zero-length ones included (their -1/+1 land on the same index and cancel).
__device__ __forceinline__ WalkAgg thread_aggregate(const ThreadOps& t, const uint2* s_cls)
{
    WalkAgg a;
    const uint32_t vmask = (1u << t.n_valid) - 1u;          // n_valid <= 8
    const uint32_t hbv = t.hb & vmask;
    a.heads = __popc(hbv);
    const int lh = 31 - __clz(hbv);                         // index of the last head, -1 if none
    uint32_t ref = 0, qry = 0, gaps = 0;
#pragma unroll
    for (int j = 0; j < kWalkOpsPerThread; j++) {
        const uint32_t w = t.w[j];                          // ops beyond n_valid are zero words: M of length 0
        const uint2 cl = s_cls[w & 15u];
        const uint32_t len = (j >= lh) ? (w >> 4) : 0u;
        ref += cl.x * len;
        qry += cl.y * len;
        gaps += cl.x * (1u - cl.y);                         // D / N: reference only
    }
    a.ref = ref; a.qry = qry;
    a.ev = a.heads + __popc((t.hb >> 1) & vmask) + 2u * gaps;
    return a;
}

// pre-pass: aggregate of every span (one CTA per span; pure streaming read of the CIGAR words)
__global__ void __launch_bounds__(kWalkThreads) k_span_agg(const WalkParams P)
{
    __shared__ WalkAgg s_warp[kWalkThreads / 32];
    __shared__ uint2 s_cls[16];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t span = blockIdx.x + P.span_base;
    const ThreadOps t = load_ops(P.cigar, P.headbits, P.n_ops, span * (uint32_t)kWalkSpan + tid * kWalkOpsPerThread);
    init_class_table(s_cls);
    __syncthreads();
    const WalkAgg a = thread_aggregate(t, s_cls);
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
__global__ void __launch_bounds__(256) k_span_scan_local(const WalkParams P, uint32_t sc_base)
{
    __shared__ WalkAgg s_warp[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sc = blockIdx.x + sc_base;
    const uint32_t base = sc * (uint32_t)kSpanChunk + tid * 8u;
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
    if (tid == 255) P.chunk_agg[sc] = pre;
}

// level 2: one CTA turns the chunk aggregates [sc0, sc1) into exclusive prefixes (in place).  *carry is the
// aggregate of everything before sc0 on entry and of everything before sc1 on exit: the pipeline calls this
// once per chunk of spans, in order.
__global__ void __launch_bounds__(1024) k_span_scan_chunks(const WalkParams P, uint32_t sc0, uint32_t sc1, WalkAgg* carry)
{
    __shared__ WalkAgg s_warp[32];
    __shared__ WalkAgg s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = *carry;
    __syncthreads();
    for (uint32_t b0 = sc0; b0 < sc1; b0 += 1024) {
        const uint32_t i = b0 + tid;
        const WalkAgg v = i < sc1 ? P.chunk_agg[i] : WalkAgg{0, 0, 0, 0};
        const WalkAgg inc = warp_incl_scan_agg(v, lane);
        WalkAgg pre = shfl_up1_agg(inc, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        WalkAgg wpre = s_carry;
        for (uint32_t w = 0; w < warp; w++) wpre = combine(wpre, s_warp[w]);
        pre = combine(wpre, pre);
        if (i < sc1) P.chunk_agg[i] = pre;
        __syncthreads();
        if (tid == 1023) s_carry = combine(pre, v);
        __syncthreads();
    }
    if (tid == 0) *carry = s_carry;
}

// level 3: everything the walk needs to start a span, in one 16-byte word
__global__ void __launch_bounds__(256) k_span_finalize(const WalkParams P)
{
    const uint32_t s = P.span_base + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.span_end) return;
    const WalkAgg e = combine(P.chunk_agg[s / kSpanChunk], P.span_pre[s]);
    P.span_desc[s] = make_uint4(e.heads - 1u, e.ev, e.ref, 0u);
    P.span_rq[s] = make_uint2(e.ref, e.qry);
    // the walk reads a span's descriptor together with its successor's (.x): close the last span of this launch
    if (s + 1u == P.span_end) P.span_desc[s + 1u] = make_uint4(e.heads + P.span_agg[s].heads - 1u, 0u, 0u, 0u);
}

// ---- record-level pre-pass (launch_walk): a span's carry-in only concerns the ONE record that runs into it.  The events
// before the span are sum(2 + 2 gaps) over whole records plus the started part of that record; reference / query
// consumed since its head are a walk over that record's ops up to the span start.  With the D/N count of every record
// from the host packer (csv_reads::n_gap; the packer touches every op anyway) the pre-pass is ONE chained scan over
// RECORDS (16 bytes each) in which the record that crosses a span start also walks its own first ops (a few dozen for
// HiFi), instead of a second pass over every CIGAR word plus a three-launch span scan.

// Second half of the record-level pre-pass: the record scan left {record k, its first event slot, its op range [o0, o1)}
// in the descriptor of every span start B = s * kWalkSpan the record runs into (o0 < B <= o1).  A group of kCarryLanes
// lanes sums the class-weighted lengths of the ops [o0, B) -- a few dozen for HiFi -- and turns the entry into what the
// walk starts the span from: {k, event slot at B, reference consumed since the head, 0} and {reference, query} for the
// gather.  A record far longer than a span pays O(ops) per span start it crosses: the batch-level test in
// csv_batch_upload keeps such batches (ONT) on the op-level pre-pass.
constexpr uint32_t kCarryLanes = 8;
__global__ void __launch_bounds__(256) k_span_carry(const WalkParams P, uint32_t s_first, uint32_t s_last)   // span starts s_first .. s_last (inclusive)
{
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) / kCarryLanes, gl = threadIdx.x % kCarryLanes;
    const uint32_t s = g + s_first;
    const bool live = s <= s_last;
    uint4 d = make_uint4(0u, 0u, 0u, 0u);
    if (live) d = P.span_desc[s];
    const uint32_t B = s * (uint32_t)kWalkSpan;
    uint32_t ref = 0, qry = 0, gaps = 0;
    if (live) {
        uint32_t o = d.z + gl;
        for (; o + kCarryLanes < B; o += 2u * kCarryLanes) {                       // two independent loads in flight
            const uint32_t w0 = __ldg(P.cigar + o), w1 = __ldg(P.cigar + o + kCarryLanes);
            const uint32_t r0 = (kRefMask >> (w0 & 15u)) & 1u, r1 = (kRefMask >> (w1 & 15u)) & 1u;
            const uint32_t q0 = (kQryMask >> (w0 & 15u)) & 1u, q1 = (kQryMask >> (w1 & 15u)) & 1u;
            ref += r0 * (w0 >> 4) + r1 * (w1 >> 4);
            qry += q0 * (w0 >> 4) + q1 * (w1 >> 4);
            gaps += (r0 & ~q0) + (r1 & ~q1);
        }
        if (o < B) {
            const uint32_t w = __ldg(P.cigar + o), r = (kRefMask >> (w & 15u)) & 1u, q = (kQryMask >> (w & 15u)) & 1u;
            ref += r * (w >> 4); qry += q * (w >> 4); gaps += r & ~q;
        }
    }
#pragma unroll
    for (uint32_t m = kCarryLanes / 2; m > 0; m >>= 1) {
        ref += __shfl_xor_sync(0xffffffffu, ref, m);
        qry += __shfl_xor_sync(0xffffffffu, qry, m);
        gaps += __shfl_xor_sync(0xffffffffu, gaps, m);
    }
    if (live && gl == 0) {
        // head event of the record, two per D/N op so far, and its tail event if it ends right at the span start
        P.span_desc[s] = make_uint4(d.x, d.y + 1u + 2u * gaps + (B == d.w ? 1u : 0u), ref, 0u);
        P.span_rq[s] = make_uint2(ref, qry);
    }
}

// ---- TMA bulk copies (global -> shared, completion on an mbarrier): the walk is a persistent kernel whose next
// span is in flight while the current one is processed; no registers are spent on the prefetch.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// c[b] for a run-time b in 0..8 without local memory
__device__ __forceinline__ uint32_t sel9(const uint32_t (&c)[kWalkOpsPerThread + 1], uint32_t b)
{
    const bool b0 = b & 1u, b1 = b & 2u, b2 = b & 4u;
    const uint32_t a0 = b0 ? c[1] : c[0], a1 = b0 ? c[3] : c[2], a2 = b0 ? c[5] : c[4], a3 = b0 ? c[7] : c[6];
    const uint32_t d0 = b1 ? a1 : a0, d1 = b1 ? a3 : a2;
    const uint32_t r = b2 ? d1 : d0;
    return (b & 8u) ? c[8] : r;
}

constexpr uint32_t kDeadPos = 0x80000000u;   // "first index" of a record that takes no part in the depth: every event
                                             // of such a record lands at or beyond 2^31 >= any map_size and no tile sees it

// The walk proper.  Persistent CTAs, one span of kWalkSpan = 1024 ops at a time (next one in flight by TMA), 8 consecutive ops per thread.
//   A. thread-local exclusive prefix c[0..8] of reference consumption (NOT reset at record heads), bit masks of
//      the D/N ops and of the signature candidates.
//   B. warp scans with SHFL + predicated add: one packed scan for (record heads, event counts), one plain scan of
//      the reference totals; the segmented part -- reference consumed since the last record head -- comes from
//      one ballot and one extra shuffle.  Warp aggregates meet in shared memory; the span's carry-in comes from
//      the pre-pass.
//   C. replay: position of op j = c[j] + bias, where bias = (first index of the record) - (c at its head); a
//      record head only swaps the bias (one predicated shared-memory load).  D/N ops store their two events;
//      nothing is clipped here -- the tile kernel ignores what lies outside its tile, and records that do not
//      count get kDeadPos as their first index.  Candidate signatures branch to a rare path.
//   D. record boundaries (about one per warp and pass for long reads) are handled in a sparse loop: the last
//      event, ref_end and ev_start of the record that ends, the first event of the record that begins.
template <bool DEPTH, bool SIGS, int MINB>
__global__ void __launch_bounds__(kWalkThreads, MINB * 256 / kWalkThreads) k_walk(const WalkParams P)
{
    __shared__ __align__(128) uint32_t s_ops[2][kWalkSpan];                  // CIGAR words of the span in hand and of the next one
    __shared__ __align__(16) uint8_t s_hb[2][kWalkSpan / 8 + 16];            // their head bits (+ the byte that follows)
    __shared__ __align__(16) uint4 s_desc[2][2];                             // their span descriptors, each with its successor's
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ uint32_t s_pos1_[2][kWalkSpan + 4];                           // first depth index of the span's records       } double-buffered by
    __shared__ uint4 s_wagg_[2][kWalkThreads / 32];                          // {heads << 16 | events, ref total, ref since   } span parity: one CTA
                                                                             //  last head, has head} per warp                } barrier per span
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t stride = gridDim.x;
    auto issue = [&](uint32_t span, int buf) {                               // one thread: three bulk copies, one barrier phase
        const uint32_t o0 = span * (uint32_t)kWalkSpan;
        const uint32_t n = P.n_ops - o0 < (uint32_t)kWalkSpan ? P.n_ops - o0 : (uint32_t)kWalkSpan;
        const uint32_t bytes_ops = (n * 4u + 15u) & ~15u, bytes_hb = kWalkSpan / 8 + 16;
        mbar_expect_tx(&s_bar[buf], bytes_ops + bytes_hb + 32u);
        tma_bulk_g2s(s_ops[buf], P.cigar + o0, bytes_ops, &s_bar[buf]);
        tma_bulk_g2s(s_hb[buf], P.headbits + (o0 >> 3), bytes_hb, &s_bar[buf]);
        tma_bulk_g2s(&s_desc[buf][0], P.span_desc + span, 32u, &s_bar[buf]);
    };
    if (tid == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t first = P.span_base + blockIdx.x;
    if (tid == 0) {
        if (first < P.span_end) issue(first, 0);
        if (first + stride < P.span_end) issue(first + stride, 1);
    }
    uint32_t it = 0;
    for (uint32_t span = first; span < P.span_end; span += stride, it++) {
        const int buf = it & 1;
        uint32_t* const s_pos1 = s_pos1_[buf];
        uint4* const s_wagg = s_wagg_[buf];
        mbar_wait(&s_bar[buf], (it >> 1) & 1u);
        const uint32_t g0 = span * (uint32_t)kWalkSpan + tid * kWalkOpsPerThread;
        const uint4 desc = s_desc[buf][0];
        const uint32_t k_first = desc.x, n_rec = s_desc[buf][1].x - desc.x + 1u;   // record heads inside the span + the record running in
        // first index of record k_first + tid: the load is issued here and consumed after the scans
        uint32_t my_p1 = kDeadPos;
        const uint32_t my_k = k_first + tid;                                 // wraps for the span that starts the batch (k_first == -1)
        if (tid < n_rec && my_k < P.n_meta) {
            const uint4 m = __ldg(P.meta + my_k);
            if (((m.z >> 30) & 1u) && m.x + 1u < m.y) my_p1 = m.x + 1u;      // (uint32)pos + 1, cnv_caller.cpp:499
        }
        ThreadOps t;
        {
            const uint4 a = reinterpret_cast<const uint4*>(s_ops[buf])[2 * tid], c4 = reinterpret_cast<const uint4*>(s_ops[buf])[2 * tid + 1];
            t.w[0] = a.x; t.w[1] = a.y; t.w[2] = a.z; t.w[3] = a.w; t.w[4] = c4.x; t.w[5] = c4.y; t.w[6] = c4.z; t.w[7] = c4.w;
            t.hb = (uint32_t)s_hb[buf][tid] | ((uint32_t)s_hb[buf][tid + 1] << 8);
            t.n_valid = kWalkOpsPerThread;
            if (g0 + kWalkOpsPerThread > P.n_ops) {                          // last span of the batch: what lies beyond n_ops is not CIGAR
                t.n_valid = g0 >= P.n_ops ? 0u : P.n_ops - g0;
#pragma unroll
                for (int j = 0; j < kWalkOpsPerThread; j++) if ((uint32_t)j >= t.n_valid) t.w[j] = 0u;
                if (t.n_valid == 0) t.hb = 0;
            }
        }
        // ---- A
        uint32_t c[kWalkOpsPerThread + 1];
        uint32_t gm = 0, sm = 0;
        c[0] = 0;
        const uint32_t thr = P.thr;                                          // min_len << 4: a word at or above it is an op of at least min_len
        auto phase_a = [&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const uint32_t w = t.w[j];
            c[j + 1] = c[j] + class_bit(kRefLut, w) * (w >> 4);
            if (DEPTH) gm |= class_bit_at<j>(kGapLut, w);
            if (SIGS) sm |= w >= thr ? class_bit_at<j>(kSigLut, w) : 0u;
        };
        phase_a(IntC<0>()); phase_a(IntC<1>()); phase_a(IntC<2>()); phase_a(IntC<3>());
        phase_a(IntC<4>()); phase_a(IntC<5>()); phase_a(IntC<6>()); phase_a(IntC<7>());
        static_assert(kWalkOpsPerThread == 8, "phase A is unrolled by hand");
        const uint32_t vmask = (1u << t.n_valid) - 1u, hbv = t.hb & vmask, tails = (t.hb >> 1) & vmask;
        const uint32_t heads = __popc(hbv);
        const uint32_t evn = heads + __popc(tails) + 2u * __popc(gm);
        const int lh = 31 - __clz(hbv);
        const uint32_t reftail = c[kWalkOpsPerThread] - (lh < 0 ? 0u : sel9(c, (uint32_t)lh));
        // ---- B
        const uint32_t he = (heads << 16) | evn;                             // a span holds <= kWalkSpan heads and <= 4 kWalkSpan events: 16 bits each
        const uint32_t he_inc = warp_incl_scan_fast(he);
        const uint32_t S = warp_incl_scan_fast(c[kWalkOpsPerThread]);
        const uint32_t hm = __ballot_sync(0xffffffffu, heads != 0u);
        const uint32_t lower = hm & lanemask_lt();
        const uint32_t X = reftail - S;
        const uint32_t Xs = __shfl_sync(0xffffffffu, X, (31 - __clz(lower)) & 31);
        const uint32_t Xl = __shfl_sync(0xffffffffu, X, (31 - __clz(hm)) & 31);
        if (lane == 31) s_wagg[warp] = make_uint4(he_inc, S, hm ? S + Xl : S, hm != 0u);
        s_pos1[tid] = my_p1;
        for (uint32_t i = tid + kWalkThreads; i < n_rec; i += kWalkThreads) {    // spans of very short records
            const uint32_t kk = k_first + i;
            uint32_t v = kDeadPos;
            if (kk < P.n_meta) {
                const uint4 m = __ldg(P.meta + kk);
                if (((m.z >> 30) & 1u) && m.x + 1u < m.y) v = m.x + 1u;
            }
            s_pos1[i] = v;
        }
        __syncthreads();                                                     // the only CTA barrier of the span
        // everybody holds its ops in registers: the buffer is free for the span after the next
        if (tid == 0 && span + 2 * stride < P.span_end && span + 2 * stride >= span) issue(span + 2 * stride, buf);
        uint4 wa = make_uint4(0, 0, 0, 0);
        if (lane < kWalkThreads / 32) wa = s_wagg[lane];
        const uint32_t he_pre = __reduce_add_sync(0xffffffffu, lane < warp ? wa.x : 0u);
        const uint32_t wh = __ballot_sync(0xffffffffu, lane < warp && wa.w);
        uint32_t carry;                                                      // reference consumed since the last head before my warp
        if (wh) {
            const uint32_t lw = 31 - __clz(wh);
            carry = __shfl_sync(0xffffffffu, wa.z, lw) + __reduce_add_sync(0xffffffffu, (lane > lw && lane < warp) ? wa.y : 0u);
        } else carry = desc.z + __reduce_add_sync(0xffffffffu, lane < warp ? wa.y : 0u);
        if (t.n_valid != 0) {
            const uint32_t he_ex = he_pre + (he_inc - he);
            const uint32_t T_ev = desc.y + (he_ex & 0xffffu);
            uint32_t kl = he_ex >> 16;                                       // s_pos1 slot of the record running into my ops
            const uint32_t rc_entry = lower ? (S - c[kWalkOpsPerThread]) + Xs : carry + (S - c[kWalkOpsPerThread]);
            const uint32_t kl_entry = kl, bias_entry = s_pos1[kl] + rc_entry;
            // ---- C: the D / N events, in op order (32-bit slot arithmetic; the address is formed where the store is)
            if (DEPTH) {
                uint32_t bias = bias_entry, ev = T_ev;
                const uint32_t* sp = s_pos1 + kl;
#pragma unroll
                for (int j = 0; j < kWalkOpsPerThread; j++) {
                    if ((hbv >> j) & 1u) { sp++; bias = *sp - c[j]; ev += (j ? 2u : 1u); }                  // tail event of the record before + my head event
                    if ((gm >> j) & 1u) { uint32_t* e = P.events + ev; e[0] = c[j] + bias; e[1] = c[j + 1] + bias; ev += 2u; }   // -1 at its first index, +1 one past its last
                }
            }
            // ---- signatures: I / D / S of at least min_len -- about one op in 600, so the ops are not tested one by one in
            // the loop above; the record and the reference consumed since its head are rebuilt from the masks for the few
            if (SIGS && sm) {
                uint32_t todo = sm;
                while (todo) {
                    const uint32_t j = __ffs(todo) - 1u;
                    todo &= todo - 1u;
                    const uint32_t hb_low = hbv & ((2u << j) - 1u);          // record heads at my ops 0..j
                    kl = kl_entry + __popc(hb_low);
                    const uint32_t cj = sel9(c, j);
                    const uint32_t since = hb_low ? cj - sel9(c, 31u - __clz(hb_low)) : rc_entry + cj;   // reference consumed since the head
                    uint32_t w = t.w[0];
#pragma unroll
                    for (int u = 1; u < kWalkOpsPerThread; u++) w = j == (uint32_t)u ? t.w[u] : w;
                    const uint32_t op = w & 15u, len = w >> 4;
                    const uint32_t k = k_first + kl;
                    const uint4 m = __ldg(P.meta + k);
                    if (m.z >> 31) {
                        const uint32_t pos1 = m.x + 1u + since;             // reference's `pos + 1` at this op (uint32)
                        const bool beyond = pos1 >= m.y;
                        const uint32_t start = pos1, end = start + len - 1u;
                        if (!(op == 4 && beyond) && start <= end) {          // sv_caller.cpp:602-604, sv_object.cpp:25-28
                            const uint32_t sub = blockIdx.x & P.sig_sub_mask;
                            const unsigned long long sl64 = (unsigned long long)atomicAdd(&P.scalars[SC_SIG_SUB0 + sub], 1u) * (P.sig_sub_mask + 1u) + sub;   // (64 bits: a counter that ran far beyond the capacity must not wrap into it)
                            const uint32_t sl = (uint32_t)sl64;
                            if (sl64 >= P.sig_cap) P.scalars[SC_SIG_DROPPED] = 1u;
                            else {
                                P.sig.key_hi[sl] = ((unsigned long long)m.w << 32) | start;
                                P.sig.key_lo[sl] = ((unsigned long long)end << 32) | (0xffffffffu - (g0 + j));
                                P.sig.k[sl] = k;
                                const uint32_t kind = op == 1 ? 0u : (op == 2 ? 1u : 2u);
                                P.sig.kind[sl] = (uint8_t)(kind | ((beyond || (int32_t)m.x < 0) ? 0x80u : 0u));
                                if (P.sig_sub_mask == 0) atomicAdd(&P.reg_sig_cnt[m.w], 1u);   // (with the bucket order the counts per region fall out of the bucket scan: sigs.cu)
                                // the ordering's bucket: the depth tile of the owner region the start falls into (clamped into
                                // the region: monotone in start, which is all the bucket order needs)
                                const uint32_t tb = P.reg_tab[m.w], nt = P.reg_tab[m.w + 1u] - tb, beg = P.reg_tab[2u * P.n_regions + 1u + m.w];
                                const uint32_t rel = start >= beg ? (start - beg) >> kTileShift : 0u;
                                const uint32_t bk = tb + (rel < nt ? rel : nt - 1u);
                                P.sig.bucket[sl] = bk;
                                P.sig.arrival[sl] = atomicAdd(&P.bucket_cnt[bk], 1u);
                            }
                        }
                    }
                }
            }
            // ---- D
            if (DEPTH) {
                uint32_t bm = t.hb & ((2u << t.n_valid) - 1u);               // bit b: a record ends with op b-1 and (b < n_valid) one begins at op b
                uint32_t klb = kl_entry, biasb = bias_entry;
                while (bm) {
                    const uint32_t b = __ffs(bm) - 1u;
                    bm &= bm - 1u;
                    const uint32_t cb = sel9(c, b);
                    const uint32_t low = (1u << b) - 1u;
                    const uint32_t slot = T_ev + __popc((hbv & low) | ((tails & low) << 9)) + 2u * __popc(gm & low);   // after the tail event of op b-1
                    if (b) {
                        const uint32_t kt = k_first + klb;
                        const uint32_t p1 = s_pos1[klb];
                        const uint32_t ie = cb + biasb;                      // one past the last covered index
                        if (ie - p1 >= 0x80000000u) P.scalars[SC_ABSURD] = 1u;   // 2^31 reference bases in one record: not an alignment
                        P.events[slot - 1u] = ie;
                        const uint32_t re = p1 != kDeadPos ? ie : 0u;        // not clipped to the map: only compared with tile starts
                        // what the caller's per-record counts promised (csv_reads::n_gap / ref_len) is checked against what the
                        // CIGAR says: the event slots right here, the reference length by k_claim_check after the walk (then
                        // P.ref_end is the batch's check array); a wrong count voids the pass (CSV_ERR_ARG at the first fetch)
                        P.ref_end[kt] = re;
                        (P.ev_given ? P.ev_check : P.ev_start)[kt + 1u] = slot;   // a store, not a dependent load in the sparse loop: the claim is compared by k_ev_check
                    }
                    if (b < t.n_valid) {
                        klb++;
                        const uint32_t p1 = s_pos1[klb];
                        P.events[slot] = p1;
                        biasb = p1 - cb;
                    }
                }
            }
        }
    }
}

static WalkParams walk_params(csv_batch* b, const csv_scan_params* p)
{
    WalkParams P;
    P.cigar = b->d_cigar.as<uint32_t>();
    P.n_ops = (uint32_t)b->n_ops;
    P.headbits = b->d_headbits.as<uint8_t>();
    P.meta = b->d_meta.as<uint4>();
    P.span_agg = b->d_span_agg.as<WalkAgg>();
    P.span_pre = b->d_span_pre.as<WalkAgg>();
    P.chunk_agg = b->d_span_status.as<WalkAgg>();
    P.n_spans = b->n_spans;
    P.span_base = 0;
    P.span_end = b->n_spans;
    P.span_desc = b->d_span_desc.as<uint4>();
    P.events = b->d_events.as<uint32_t>();
    P.ev_cap = (uint32_t)b->ev_cap;
    P.ev_start = b->d_ev_start.as<uint32_t>();
    // with csv_reads::ref_len the tile ranges were derived from the claim before the walk (d_ref_end is theirs): the walk
    // leaves what it finds in the check array and k_claim_check compares the two
    P.ref_end = (b->claimed_ref && p->want_depth ? b->d_ref_chk : b->d_ref_end).as<uint32_t>();
    P.n_meta = b->n_reads;
    P.min_len = p->min_len < (1u << 28) ? p->min_len : 0u;
    P.thr = p->min_len < (1u << 28) ? p->min_len << 4 : 0xffffffffu;          // CIGAR lengths have 28 bits: nothing reaches a larger threshold (0xffffffff itself is op 15)
    P.scalars = b->d_scalars.as<uint32_t>();
    P.sig.key_hi = b->d_sig_hi.as<unsigned long long>();
    P.sig.key_lo = b->d_sig_lo.as<unsigned long long>();
    P.sig.k = b->d_sig_k.as<uint32_t>();
    P.sig.kind = b->d_sig_kind.as<uint8_t>();
    P.sig.bucket = b->d_sig_bucket.as<uint32_t>();
    P.sig.arrival = b->d_sig_arrival.as<uint32_t>();
    P.reg_tab = b->d_reg_tab.as<uint32_t>();
    P.n_regions = b->n_regions;
    P.bucket_cnt = b->d_bucket_cnt.as<uint32_t>();
    P.sig_sub_mask = b->sig_sub_mask;
    P.sig_cap = (uint32_t)b->sig_cap;
    P.reg_sig_cnt = b->d_reg_sig_cnt.as<uint32_t>();
    P.span_rq = b->d_span_rq.as<uint2>();
    P.n_gap = b->d_n_gap.as<uint32_t>();
    P.cig_off = b->d_cig_off.as<unsigned long long>();
    P.ne_idx = b->d_ne_idx.as<uint32_t>();
    P.ev_given = b->rec_prepass ? 1u : 0u;
    P.ev_check = b->d_ev_check.as<uint32_t>();
    return P;
}

// Record-level pre-pass of a batch whose caller counted the D / N ops per record: once per pass, before the walk.
// what: 1 = the record scan (event slots, span starts parked), 2 = the span carry, 3 = both
// span_from / span_to (what & 2): only the span starts in (span_from, span_to] -- the pipeline chunks of a batch whose CIGAR
// words arrive chunk by chunk; every span start must be done exactly once (the kernel rewrites the parked entry)
int launch_record_prepass(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p, int what, uint32_t span_from, uint32_t span_to)
{
    if (!b->rec_prepass || b->n_ops == 0) return CSV_OK;
    const WalkParams P = walk_params(b, p);
    if (what & 1) {
            const WalkParams Q = P;
            const uint32_t* n_rec = P.scalars + SC_N_NONEMPTY;
            auto in = [=] __device__(uint64_t k) -> uint32_t {
                const uint32_t i = Q.ne_idx[k];
                const uint32_t ops = (uint32_t)(Q.cig_off[i + 1] - Q.cig_off[i]), g = Q.n_gap[i];
                return 2u + 2u * (g < ops ? g : ops);                       // clamped: event slots stay inside the event array whatever the caller claims
            };
            auto out = [=] __device__(uint64_t k, uint32_t ex, uint32_t v) {
                const uint32_t i = Q.ne_idx[k];
                Q.ev_start[k] = ex;
                const unsigned long long o0 = Q.cig_off[i], o1 = Q.cig_off[i + 1];
                const uint32_t s_lo = (uint32_t)(o0 / kWalkSpan) + 1u, s_hi = (uint32_t)(o1 / kWalkSpan);
                if (k == 0) { Q.span_desc[0] = make_uint4(0xffffffffu, 0u, 0u, 0u); Q.span_rq[0] = make_uint2(0u, 0u); }
                // span starts B with o0 < B <= o1: this record runs into them (or ends right there).  Parked for k_span_carry,
                // which sums the record's ops up to B with a group of lanes per span (n_ops < 2^31: offsets fit 32 bits).
                for (uint32_t s = s_lo; s <= s_hi; s++) Q.span_desc[s] = make_uint4((uint32_t)k, ex, (uint32_t)o0, (uint32_t)o1);
                if (k + 1 == (uint64_t)*n_rec) {
                    Q.ev_start[k + 1] = ex + v;
                    if (s_hi != Q.n_spans) Q.span_desc[Q.n_spans] = make_uint4((uint32_t)k, 0u, 0u, 0u);      // sentinel: closes the last span
                }
            };
            CSV_TRY(chained_scan(ctx, in, out, b->n_reads, n_rec, nullptr));
    }
    if (what & 2) {
            const uint32_t n_carry = (uint32_t)(b->n_ops / kWalkSpan);               // span starts B = s * kWalkSpan, 1 <= s <= n_carry
            const uint32_t s_first = span_from + 1u, s_last = span_to < n_carry ? span_to : n_carry;
            if (s_first <= s_last) {
                const uint32_t per_cta = 256 / kCarryLanes, n = s_last - s_first + 1u;
                k_span_carry<<<(n + per_cta - 1) / per_cta, 256, 0, ctx->stream>>>(P, s_first, s_last);
                ctx->launches++;
            }
    }
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// csv_reads::n_gap against the CIGARs: the event slot the record scan derived for the end of every record and the one the
// walk reached there.  Beside the tiles; only the fetches wait for the verdict.
__global__ void __launch_bounds__(256) k_ev_check(const uint32_t* __restrict__ ev_start, const uint32_t* __restrict__ ev_check, uint32_t* scalars)
{
    const uint32_t n = scalars[SC_N_NONEMPTY];
    bool bad = false;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) bad |= ev_start[k + 1u] != ev_check[k + 1u];
    if (bad) scalars[SC_BAD_GAPS] = 1u;
}

int launch_ev_check(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p)
{
    if (!b->rec_prepass || !p->want_depth || b->n_ops == 0) return CSV_OK;
    const uint32_t grid = (b->n_reads + 255) / 256 < (uint32_t)ctx->sm_count * 2 ? (b->n_reads + 255) / 256 : (uint32_t)ctx->sm_count * 2;
    k_ev_check<<<grid, 256, 0, ctx->stream>>>(b->d_ev_start.as<uint32_t>(), b->d_ev_check.as<uint32_t>(), b->d_scalars.as<uint32_t>());
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// csv_reads::ref_len against the CIGARs: claimed[k] (k_pmax_chained derived it before the walk) and found[k] (the walk)
__global__ void __launch_bounds__(256) k_claim_check(const uint32_t* __restrict__ claimed, const uint32_t* __restrict__ found, uint32_t* scalars)
{
    const uint32_t n = scalars[SC_N_NONEMPTY];
    bool bad = false;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) bad |= claimed[k] != found[k];
    if (bad) scalars[SC_BAD_GAPS] = 1u;
}

// after the last walk chunk of a pass whose tile ranges came from the claim; beside the tiles, nothing waits for it but the fetches
int launch_claim_check(csv_ctx* ctx, csv_batch* b)
{
    if (!b->claimed_ref || b->n_reads == 0) return CSV_OK;
    const uint32_t grid = (b->n_reads + 255) / 256 < (uint32_t)ctx->sm_count * 4 ? (b->n_reads + 255) / 256 : (uint32_t)ctx->sm_count * 4;
    k_claim_check<<<grid, 256, 0, ctx->stream>>>(b->d_ref_end.as<uint32_t>(), b->d_ref_chk.as<uint32_t>(), b->d_scalars.as<uint32_t>());
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// Walks the spans [span0, span1); span0 must be a multiple of kSpanChunk and the chunks of one pass must come in
// order on one stream (the carry of the span scan lives in b->d_scan_carry).
int launch_walk(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p, uint32_t span0, uint32_t span1)
{
    if (b->n_ops == 0 || span0 >= span1) return CSV_OK;
    WalkParams P = walk_params(b, p);
    P.span_base = span0;
    P.span_end = span1;
    const uint32_t n = span1 - span0;
    if (!b->rec_prepass) {
        if (span0 == 0) CSV_CUDA(cudaMemsetAsync(b->d_scan_carry.p, 0, sizeof(WalkAgg), ctx->stream));
        const uint32_t sc0 = span0 / kSpanChunk, sc1 = (span1 + kSpanChunk - 1) / kSpanChunk;
        k_span_agg<<<n, kWalkThreads, 0, ctx->stream>>>(P);
        k_span_scan_local<<<sc1 - sc0, 256, 0, ctx->stream>>>(P, sc0);
        k_span_scan_chunks<<<1, 1024, 0, ctx->stream>>>(P, sc0, sc1, b->d_scan_carry.as<WalkAgg>());
        k_span_finalize<<<(n + 255) / 256, 256, 0, ctx->stream>>>(P);
        ctx->launches += 4;
    }
    static const int minb = getenv("CSV_WALK_MINB") ? atoi(getenv("CSV_WALK_MINB")) : 4;    // tuning knob: CTAs per SM the compiler targets
    static const int gmul = getenv("CSV_WALK_GRID") ? atoi(getenv("CSV_WALK_GRID")) : 128;   // CTAs per SM in the grid (8 resident): each walks ~10-20 spans, the next one prefetched
    const uint32_t per_sm = (uint32_t)(minb == 5 || minb == 6 ? minb : 4);
    const uint32_t cap = (uint32_t)ctx->sm_count * (gmul > 0 ? (uint32_t)gmul : per_sm);
    const uint32_t grid = n < cap ? n : cap;
    if (p->want_depth && p->want_sigs) {
        if (minb == 5) k_walk<true, true, 5><<<grid, kWalkThreads, 0, ctx->stream>>>(P);
        else if (minb == 6) k_walk<true, true, 6><<<grid, kWalkThreads, 0, ctx->stream>>>(P);
        else k_walk<true, true, 4><<<grid, kWalkThreads, 0, ctx->stream>>>(P);
    } else if (p->want_depth) k_walk<true, false, 4><<<grid, kWalkThreads, 0, ctx->stream>>>(P);
    else k_walk<false, true, 4><<<grid, kWalkThreads, 0, ctx->stream>>>(P);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

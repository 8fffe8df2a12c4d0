// depth_tiles.cu -- per-base read depth (cnv_caller.cpp:507-519) and the two
// reductions that follow it (:534-535), from the walk's op-ordered event list.
//
// The depth map of every region is cut into tiles of kTile positions.
//
// launch_tile_ranges: records are coordinate-sorted, so the records that can
//   touch a tile [T0,T1) of contig c are a contiguous run [r_lo, r_hi):
//     r_hi = first record with (tid, pos0+1) >= (c, T1)
//     r_lo = first record whose running maximum of (tid, ref_end) exceeds (c, T0)
//   (a u64 prefix-max over records, then two binary searches per tile).  Their
//   events are one contiguous slice of the event array.
//
// k_depth_tiles16: one CTA owns one tile.  Events come in PAIRS: every record starts on an even slot and
//   owns an even number of events, so slots (2i, 2i+1) are the two ends [p0, p1) of one covered stretch of a
//   record: +1 at p0, -1 at p1.  The CTA streams the pairs of the slice with 64-bit loads: an end left of the
//   tile moves the tile's carry-in (depth at T0), an end inside goes into a shared-memory difference array,
//   the rest is skipped.  No inter-tile dependency, no global atomics.  Slices of long reads (> 4096 pairs)
//   are not streamed whole: two binary searches per record find the pairs that can touch the tile.
//   The difference array holds 16-bit counters, two per word, biased by 0x8000 so that no borrow ever crosses
//   the halves: 16 KB per tile, a quarter of the shared-memory traffic of 32-bit counters.  A record moves any
//   counter (and any run of them) by -1, 0 or +1 in total, so only a tile that more than 32767 records overlap
//   (local coverage in the tens of thousands) could overflow them: those tiles go to k_depth_tiles_wide
//   (32-bit counters) through a list k_tile_ranges builds.
//   Thread ownership is chosen for the memory system, not for the scan: lane l of warp w owns the four
//   8-position chunks l, l+32, l+64, l+96 of the warp's 1024 positions, so every 128-bit shared-memory access
//   and every 256-bit global store of a warp covers one contiguous kilobyte.  The prefix sum is IDP.2A
//   (dot product of the two halves with {1,0} / {1,1}, accumulate in the running depth): one FMA-pipe
//   instruction per position, beside the ALU-pipe reductions (sum, min).  Two packed warp scans (two chunk rows
//   each) order the chunks.  sum(depth) / count(depth > 0) are reduced per tile and summed per
//   region by k_region_stats.
#include "batch.cuh"
#include "scan.cuh"

#include <cstdlib>

namespace csv {

// A record's events alternate +,-,+,- in position order, so its net contribution to any position (and to any run of
// consecutive positions) is -1, 0 or +1: a 16-bit counter cannot overflow while at most this many records overlap the tile.
constexpr uint32_t kNarrowMaxRecords = 32767;
// Slices longer than this many pairs (long reads: a 50 kb ONT record owns ~5000 events, 70 of them overlap a tile) are
// not streamed whole: the events of a record are sorted, so two binary searches per record find the pairs that can
// touch the tile.
constexpr uint32_t kSearchMinPairs = 4096;
constexpr uint32_t kSearchShortRecord = 64;     // events; shorter records are taken whole

// pair range [a, b) of record k that can touch [T0, T1): everything before a is an even number of events left of the
// tile (net zero), everything from b on lies at or beyond T1
// first slot in [lo, hi) whose event is >= x.  The events of a long record are spread almost evenly over its span, so the
// slot is guessed by interpolation and bracketed with a doubling step before the binary search: ~3 + 5 probes, most of
// them in one or two cache lines, instead of 13 scattered ones.
__device__ __forceinline__ uint32_t event_lower_bound(const uint32_t* __restrict__ events, uint32_t lo, uint32_t hi, uint32_t x)
{
    if (lo >= hi) return lo;
    const uint32_t first = __ldg(events + lo), last = __ldg(events + hi - 1u);
    if (x <= first) return lo;
    if (x > last) return hi;
    // first < x <= last: the answer is in (lo, hi - 1]
    uint32_t g = lo + (uint32_t)(((unsigned long long)(x - first) * (hi - 1u - lo)) / (last - first));
    uint32_t a, b;                                          // invariant: events[a] < x <= events[b]
    if (__ldg(events + g) < x) {
        a = g; uint32_t step = 16;
        for (;;) { b = a + step < hi - 1u ? a + step : hi - 1u; if (b == hi - 1u || __ldg(events + b) >= x) break; a = b; step <<= 1; }
    } else {
        b = g; uint32_t step = 16;
        for (;;) { a = b > lo + step ? b - step : lo; if (a == lo || __ldg(events + a) < x) break; b = a; step <<= 1; }
    }
    while (b - a > 1u) { const uint32_t mid = a + ((b - a) >> 1); if (__ldg(events + mid) < x) a = mid; else b = mid; }
    return b;
}

// pair range [a, b) of record k that can touch [T0, T1): everything before a is an even number of events left of the
// tile (net zero), everything from b on lies at or beyond T1
__device__ __forceinline__ uint2 record_pair_range(const uint32_t* __restrict__ events, const uint32_t* __restrict__ ev_start, uint32_t k,
                                                  uint32_t T0, uint32_t T1)
{
    const uint32_t es = ev_start[k], ee = ev_start[k + 1];
    if (ee - es <= kSearchShortRecord) return make_uint2(es >> 1, ee >> 1);
    const uint32_t lb0 = event_lower_bound(events, es, ee, T0);
    const uint32_t lb1 = event_lower_bound(events, lb0, ee, T1);
    return make_uint2(lb0 >> 1, (lb1 + 1u) >> 1);
}

// ------------------------------------------------------- prefix max over records
constexpr int kPmThreads = 256, kPmItems = 8, kPmTile = kPmThreads * kPmItems;

__device__ __forceinline__ unsigned long long pm_value(const unsigned long long* key, const uint32_t* ref_end, uint32_t k)
{
    return (key[k] & 0xffffffff00000000ull) | ref_end[k];
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// Running maximum of (tid, ref_end) over the records [bounds[0], bounds[1]) of one pipeline chunk, ONE launch: tiles of
// kPmTile records are claimed by ticket and chained by a look-back.  Records are sorted by contig, so the u64 maximum
// is a maximum of ref_end that restarts at every contig change: a tile publishes (epoch | flag, max ref_end among the
// records of its LAST contig) in one 64-bit word -- complete (kFlagPrefix) at once when the tile starts the chunk,
// starts a new contig or contains a contig change, otherwise its own share first (kFlagAgg) and the complete value after
// the look-back.  A successor only looks back if its first record continues the contig of the record right before it.
// The coordinate-order check of the batch rides along.
// complete: the published value needs nothing from earlier tiles; look: the tile's first records continue the contig of
// the tile before, so the carry-in has to be fetched (both can hold: a tile that continues a contig AND starts another).
__device__ __forceinline__ uint32_t lookback_max_u32(unsigned long long* status, uint32_t t, uint32_t aggregate, bool complete, bool look_back, uint32_t epoch)
{
    const uint32_t lane = lane_id();
    if (lane == 0) st_volatile_u64(&status[t], lb_pack(epoch, complete ? kFlagPrefix : kFlagAgg, aggregate));
    if (!look_back) return 0u;
    uint32_t excl = 0;
    int64_t look = (int64_t)t - 1;
    for (;;) {
        const int64_t idx = look - lane;
        uint32_t flag, val;
        do {
            if (idx >= 0) {
                const unsigned long long w = ld_volatile_u64(&status[idx]);
                const uint32_t hi = (uint32_t)(w >> 32);
                flag = ((hi >> 2) == epoch) ? (hi & 3u) : 0u;
                val = (uint32_t)w;
            } else { flag = kFlagPrefix; val = 0; }
        } while (__any_sync(0xffffffffu, flag == 0));
        const uint32_t pm = __ballot_sync(0xffffffffu, flag == kFlagPrefix);
        if (pm) {
            const uint32_t j = __ffs(pm) - 1;
            excl = max(excl, __reduce_max_sync(0xffffffffu, lane <= j ? val : 0u));
            break;
        }
        excl = max(excl, __reduce_max_sync(0xffffffffu, val));
        look -= 32;
    }
    if (lane == 0 && !complete) st_volatile_u64(&status[t], lb_pack(epoch, kFlagPrefix, max(excl, aggregate)));
    return excl;
}

// claim (optional): the pass runs BEFORE the walk on what the caller's csv_reads::ref_len promises -- ref_end[k] is derived
// here from the record's metadata and claimed reference length (and stored: the walk compares it with what it finds).
struct PmClaim { const uint4* meta4; const uint32_t* ref_len; const uint32_t* ne_idx; };
__global__ void __launch_bounds__(kPmThreads) k_pmax_chained(const unsigned long long* __restrict__ meta, uint32_t* __restrict__ ref_end,
                                                             uint32_t* scalars, const uint32_t* bounds, unsigned long long* pmax,
                                                             uint32_t* ticket, unsigned long long* status, uint32_t epoch, const PmClaim claim)
{
    // Global loads and stores are striped (lane-contiguous, one 256-byte row per warp and instruction); the scan wants
    // kPmItems consecutive records per thread.  The tile changes hands in shared memory, rows padded by one word per
    // kPmItems so that the blocked accesses spread over the banks.
    __shared__ unsigned long long s_v[kPmTile + kPmTile / kPmItems];
    __shared__ unsigned long long s_w[kPmThreads / 32];
    __shared__ uint32_t s_tile, s_excl;
    const uint32_t kb = bounds[0], n = bounds[1] - kb;
    meta += kb; ref_end += kb; pmax += kb;
    const uint4* meta4 = claim.meta4 ? claim.meta4 + kb : nullptr;
    const uint32_t* ne_idx = claim.meta4 ? claim.ne_idx + kb : nullptr;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = (uint32_t)(((uint64_t)n + kPmTile - 1) / kPmTile);
    for (;;) {
        __syncthreads();                                        // s_v, s_w, s_tile are rewritten by the next tile
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= n_tiles) break;
        const uint64_t base = (uint64_t)t * kPmTile;
        bool unsorted = false;
#pragma unroll
        for (int j = 0; j < kPmItems; j++) {
            const uint32_t i = j * kPmThreads + threadIdx.x;
            const uint64_t k = base + i;
            unsigned long long v = 0ull;
            if (k < n) {
                const unsigned long long key = meta[k];
                uint32_t re;
                if (meta4) {
                    // the walk's rule (k_walk, phase D): a record that takes part in the depth ends one past its last covered
                    // index = (uint32)pos0 + 1 + reference bases consumed; any other record counts as 0
                    const uint4 m = meta4[k];
                    const bool live = ((m.z >> 30) & 1u) && m.x + 1u < m.y;
                    re = live ? m.x + 1u + claim.ref_len[ne_idx[k]] : 0u;
                    ref_end[k] = re;
                } else re = ref_end[k];
                v = (key & 0xffffffff00000000ull) | re;
                // coordinate order check rides along: (tid, pos0 + 1) must not decrease (also across chunk borders)
                if (kb + k > 0 && meta[(long long)k - 1] > key) unsorted = true;
            }
            s_v[i + i / kPmItems] = v;
        }
        if (unsorted) scalars[SC_UNSORTED] = 1;
        __syncthreads();
        // blocked arrangement: thread owns kPmItems consecutive records
        const uint32_t o = threadIdx.x * (kPmItems + 1);
        unsigned long long v[kPmItems], run = 0;
#pragma unroll
        for (int j = 0; j < kPmItems; j++) { v[j] = s_v[o + j]; run = umax64(run, v[j]); }
        unsigned long long inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc = umax64(inc, x); }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            // the tile as a whole: its last contig and the maximum there; the carry-in only concerns its first contig
            unsigned long long tot = s_w[0];
#pragma unroll
            for (int i = 1; i < kPmThreads / 32; i++) tot = umax64(tot, s_w[i]);
            const uint32_t n_here = n - base < (uint64_t)kPmTile ? (uint32_t)(n - base) : (uint32_t)kPmTile;
            const uint32_t tid_first = (uint32_t)(meta[base] >> 32), tid_last = (uint32_t)(meta[base + n_here - 1] >> 32);
            const bool continues = t > 0 && (uint32_t)(meta[(long long)base - 1] >> 32) == tid_first;
            const uint32_t e = lookback_max_u32(status, t, (uint32_t)tot, !continues || tid_first != tid_last, continues, epoch);
            // with an unsorted batch (reported above) the high words may disagree: results are void anyway
            if (lane == 0) s_excl = e;
        }
        __syncthreads();
        unsigned long long pre = s_excl ? (((unsigned long long)(uint32_t)(meta[base] >> 32) << 32) | s_excl) : 0ull;
        for (uint32_t w = 0; w < warp; w++) pre = umax64(pre, s_w[w]);
        unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        pre = umax64(pre, excl);
#pragma unroll
        for (int j = 0; j < kPmItems; j++) { pre = umax64(pre, v[j]); s_v[o + j] = pre; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPmItems; j++) {
            const uint32_t i = j * kPmThreads + threadIdx.x;
            if (base + i < n) pmax[base + i] = s_v[i + i / kPmItems];
        }
    }
}

// ------------------------------------------------------------ tile -> event slice
// Static half: r_hi = first record with (tid, pos0 + 1) >= (tid, T1) only depends on the sort keys, so it runs on the
// tile stream while the walk is busy; it is parked in tile_ev[t].y.
__global__ void k_tile_hi(const uint4* __restrict__ tile_desc, uint32_t t_begin, uint32_t t_end, const unsigned long long* __restrict__ key,
                          const uint32_t* scalars, uint2* tile_ev)
{
    const uint32_t n = scalars[SC_N_NONEMPTY];
    for (uint32_t t = t_begin + blockIdx.x * blockDim.x + threadIdx.x; t < t_end; t += gridDim.x * blockDim.x) {
        const uint4 d = tile_desc[t];            // {region, positions, T0, tid}
        const unsigned long long key_hi = ((unsigned long long)d.w << 32) | (unsigned long long)(d.z + d.y);   // (tid, T1)
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (key[mid] < key_hi) lo = mid + 1; else hi = mid;
        }
        tile_ev[t] = make_uint2(0u, lo);
    }
}

// After the walk: r_lo = first record whose running max (tid, ref_end) exceeds (tid, T0); the records [r_lo, r_hi) are
// the ones that can touch the tile, their events one contiguous slice.
__global__ void k_tile_ranges(const uint4* __restrict__ tile_desc, uint32_t t_begin, uint32_t t_end,
                              const unsigned long long* __restrict__ pmax, const uint32_t* __restrict__ ev_start,
                              uint32_t* scalars, const uint32_t* bounds, uint2* tile_ev, uint4* tile_q, uint2* tile_r, uint32_t* wide_list)
{
    const uint32_t kb = bounds[0];
    for (uint32_t t = t_begin + blockIdx.x * blockDim.x + threadIdx.x; t < t_end; t += gridDim.x * blockDim.x) {
        const uint4 d = tile_desc[t];
        const unsigned long long key_lo = ((unsigned long long)d.w << 32) | (unsigned long long)d.z;           // (tid, T0)
        const uint32_t r_hi = tile_ev[t].y;
        // first record of [kb, r_hi) with pmax > key_lo.  It lies a few dozen records below r_hi (the records that overlap
        // one tile), so gallop down from r_hi before bisecting: ~12 dependent loads instead of 23 over 6 M records.
        uint32_t lo = kb < r_hi ? kb : r_hi, hi = r_hi;
        for (uint32_t step = 64; lo < hi; step <<= 2) {
            const uint32_t probe = hi - lo > step ? hi - step : lo;
            if (pmax[probe] <= key_lo) { lo = probe + 1; break; }
            hi = probe;                                          // pmax[probe] > key_lo: the answer is at or below probe
            if (probe == lo) break;
        }
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (pmax[mid] <= key_lo) lo = mid + 1; else hi = mid;
        }
        const uint32_t r_lo = lo;
        const uint2 er = r_lo < r_hi ? make_uint2(ev_start[r_lo], ev_start[r_hi]) : make_uint2(0u, 0u);
        tile_ev[t] = er;
        tile_q[t] = make_uint4(d.z, d.y, er.x, er.y);
        tile_r[t] = make_uint2(r_lo, r_hi);
        if (r_hi - r_lo > kNarrowMaxRecords && r_lo < r_hi) wide_list[atomicAdd(&scalars[SC_N_WIDE], 1u)] = t;   // rare: 32-bit counters
    }
}

// first compact record of every pipeline chunk: bounds[c] = first record with tid >= first_tid[c]
__global__ void k_chunk_bounds(const unsigned long long* __restrict__ key, const uint32_t* scalars, const uint32_t* first_tid, uint32_t n_chunks,
                               uint32_t* bounds)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    const uint32_t n = scalars[SC_N_NONEMPTY];
    if (c == n_chunks) { bounds[c] = n; return; }
    const unsigned long long want = (unsigned long long)first_tid[c] << 32;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (key[mid] < want) lo = mid + 1; else hi = mid;
    }
    bounds[c] = c == 0 ? 0u : lo;
}

int launch_chunk_bounds(csv_ctx* ctx, csv_batch* b)
{
    const uint32_t nc = (uint32_t)b->chunks.size();
    k_chunk_bounds<<<(nc + 1 + 63) / 64, 64, 0, ctx->stream>>>(b->d_key.as<unsigned long long>(), b->d_scalars.as<uint32_t>(),
                                                              b->d_chunk_tid.as<uint32_t>(), nc, b->d_chunk_bounds.as<uint32_t>());
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// static half of the tile ranges, all tiles: needs the sort keys only (prep)
int launch_tile_hi(csv_ctx* ctx, csv_batch* b)
{
    if (b->n_tiles == 0) return CSV_OK;
    const uint32_t grid_t = (b->n_tiles + 255) / 256;
    k_tile_hi<<<grid_t, 256, 0, ctx->stream>>>(b->d_tile_desc.as<uint4>(), 0u, b->n_tiles, b->d_key.as<unsigned long long>(), b->d_scalars.as<uint32_t>(),
                                              b->d_tile_ev.as<uint2>());
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// event slices of the tiles of pipeline chunk c (needs the walk of chunk c + 1: see csv_scan_run)
// what: 1 = the prefix max alone, 2 = the range searches alone (it is done), 3 = both
int launch_tile_ranges(csv_ctx* ctx, csv_batch* b, uint32_t c, int what)
{
    const PipeChunk& ch = b->chunks[c];
    if (ch.tiles.empty()) return CSV_OK;
    const unsigned long long* meta = b->d_key.as<unsigned long long>();
    uint32_t* ref_end = b->d_ref_end.as<uint32_t>();
    uint32_t* scalars = b->d_scalars.as<uint32_t>();
    const uint32_t* bounds = b->d_chunk_bounds.as<uint32_t>() + c;
    unsigned long long* part = b->d_pmax_part.as<unsigned long long>();
    unsigned long long* pmax = b->d_pmax.as<unsigned long long>();
    const uint32_t n_part = (uint32_t)(((uint64_t)ch.rec_upper + kPmTile - 1) / kPmTile);
    if (n_part && (what & 1)) {
        const uint32_t grid = n_part < (uint32_t)ctx->sm_count * 8 ? n_part : (uint32_t)ctx->sm_count * 8;
        // ticket and status words are the batch's own: this launch runs on the tile stream beside chained scans of the
        // signature side stream, which share the context's
        PmClaim claim = {nullptr, nullptr, nullptr};
        if (b->claimed_ref) claim = PmClaim{b->d_meta.as<uint4>(), b->d_ref_len.as<uint32_t>(), b->d_ne_idx.as<uint32_t>()};
        k_pmax_chained<<<grid, kPmThreads, 0, ctx->stream>>>(meta, ref_end, scalars, bounds, pmax, b->d_tickets.as<uint32_t>() + c, part, next_epoch(ctx), claim);
        ctx->launches++;
    }
    if (what & 2) for (const auto& tr : ch.tiles) {
        const uint32_t grid_t = (tr.second - tr.first + 255) / 256;
        k_tile_ranges<<<grid_t, 256, 0, ctx->stream>>>(b->d_tile_desc.as<uint4>(), tr.first, tr.second, pmax, b->d_ev_start.as<uint32_t>(),
                                                      scalars, bounds, b->d_tile_ev.as<uint2>(), b->d_tile_q.as<uint4>(), b->d_tile_r.as<uint2>(), b->d_wide_list.as<uint32_t>());
        ctx->launches++;
    }
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// --------------------------------------------------------------------- tile kernel
struct TileParams {
    const uint4* tile_desc;          // static per tile: {region, positions in tile, T0, tid}
    const uint2* tile_ev;            // event slice [x, y)
    const uint4* tile_q;             // {T0, positions, x, y}: all the 16-bit kernel needs, one 128-bit load
    const uint2* tile_r;             // records [r_lo, r_hi) that can touch the tile
    const uint32_t* ev_start;
    const uint32_t* events;
    uint32_t ev_cap;
    uint32_t* depth;                 // n_tiles * kTile words
    unsigned long long* tile_sum;    // per-tile partial reductions (no contended atomics)
    uint32_t* tile_nz;
    uint32_t n_tiles;                // all tiles of the batch (wide kernel)
    uint32_t t_begin, t_end;         // tiles of this launch (16-bit kernel)
    const uint32_t* wide_list;       // tiles that need 32-bit counters
    const uint32_t* scalars;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 256-bit streaming store (sm_100): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void st_na_v8(uint32_t* p, int a, int b, int c, int d, int e, int f, int g, int h)
{
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
                 "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}

// 32-bit counters, for the tiles on the wide list (and the reference implementation of the tile logic).
// One CTA per tile, PT consecutive positions per thread; shared-memory rows of PT words are padded by 4 words:
// 16-byte aligned and conflict-free for the 128-bit accesses of their owner.
template <int PT>
__global__ void __launch_bounds__(kTile / PT, (PT == 32 ? 4 : 2)) k_depth_tiles_wide(const TileParams P)
{
    constexpr int kThreads = kTile / PT, kWarps = kThreads / 32, kPadW = PT + 4;
    constexpr int kShift = (PT == 32 ? 5 : 4);
    constexpr int kPP = 1024 / kThreads;                           // pairs per thread and batch (2048 events per batch)
    static_assert(kWarps <= 32, "one lane per warp in the offset reduction");
    __shared__ __align__(16) int s_diff[kThreads * kPadW];
    // s_wcar_ is double-buffered by tile parity: a warp that runs ahead writes its share of the NEXT tile's carry-in
    // before the first barrier of that tile, while a slow warp may still be reading this tile's after the second one
    __shared__ int s_wtot[kWarps], s_wcar_[2][kWarps];
    __shared__ uint2 s_rng[kThreads];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint2* __restrict__ pairs = reinterpret_cast<const uint2*>(P.events);
    const uint32_t pair_cap = P.ev_cap >> 1;
    uint32_t parity = 0;

    for (uint32_t i = tid; i < kThreads * kPadW / 4; i += kThreads) reinterpret_cast<int4*>(s_diff)[i] = make_int4(0, 0, 0, 0);
    const uint32_t n_list = P.scalars[SC_N_WIDE];
    uint32_t li = blockIdx.x;
    uint32_t t = li < n_list ? P.wide_list[li] : P.n_tiles;
    uint4 desc = make_uint4(0, 0, 0, 0);
    uint2 er = make_uint2(0, 0);
    uint2 pf[kPP];
#pragma unroll
    for (int u = 0; u < kPP; u++) pf[u] = make_uint2(kNone, kNone);
    if (t < P.n_tiles) {
        desc = __ldg(P.tile_desc + t); er = __ldg(P.tile_ev + t);
        const uint32_t pb = er.x >> 1, pe = min(er.y >> 1, pair_cap);
#pragma unroll
        for (int u = 0; u < kPP; u++) { const uint32_t i = pb + tid + u * kThreads; if (i < pe) pf[u] = __ldg(pairs + i); }
    }
    __syncthreads();

    while (t < P.n_tiles) {
        // descriptor of the NEXT tile of this CTA: loaded now, used in the middle of this iteration
        li += gridDim.x;
        const uint32_t tn = li < n_list ? P.wide_list[li] : P.n_tiles;
        uint4 desc_n = make_uint4(0, 0, 0, 0);
        uint2 er_n = make_uint2(0, 0);
        if (tn < P.n_tiles) { desc_n = __ldg(P.tile_desc + tn); er_n = __ldg(P.tile_ev + tn); }

        const uint32_t n_here = desc.y, T0 = desc.z;
        const uint32_t pb = er.x >> 1, pe = min(er.y >> 1, pair_cap);
        int* const s_wcar = s_wcar_[parity];
        parity ^= 1u;
        // ---- stretches of the records that overlap the tile (s_diff is all zero here)
        int mycarry = 0;
        auto apply = [&](const uint2 pr) {
            const uint32_t q0 = pr.x - T0, q1 = pr.y - T0;         // wrap to huge values left of the tile
            mycarry += (pr.x < T0) ? 1 : 0;
            mycarry -= (pr.y < T0) ? 1 : 0;
            if (q0 < n_here) atomicAdd(&s_diff[q0 + (q0 >> kShift) * 4u], 1);
            if (q1 < n_here) atomicAdd(&s_diff[q1 + (q1 >> kShift) * 4u], -1);
        };
        if (pe - pb > kSearchMinPairs) {
            // long records (or a pile-up): per record, only the pairs that can touch the tile
            const uint2 rr = __ldg(P.tile_r + t);
            for (uint32_t rb = rr.x; rb < rr.y; rb += kThreads) {
                const uint32_t cnt = rr.y - rb < (uint32_t)kThreads ? rr.y - rb : (uint32_t)kThreads;
                __syncthreads();
                if (tid < cnt) s_rng[tid] = record_pair_range(P.events, P.ev_start, rb + tid, T0, T0 + n_here);
                __syncthreads();
                for (uint32_t r = warp; r < cnt; r += kWarps) {
                    const uint2 g2 = s_rng[r];
                    const uint32_t pend = g2.y < pair_cap ? g2.y : pair_cap;
                    for (uint32_t q = g2.x + lane; q < pend; q += 32) apply(__ldg(pairs + q));
                }
            }
        } else {
#pragma unroll
        for (int u = 0; u < kPP; u++) apply(pf[u]);
        for (uint32_t base = pb + kThreads * kPP; base < pe; base += kThreads * kPP) {      // rare: > 2048 events
            uint2 p[kPP];
#pragma unroll
            for (int u = 0; u < kPP; u++) { const uint32_t i = base + tid + u * kThreads; p[u] = i < pe ? __ldg(pairs + i) : make_uint2(kNone, kNone); }
#pragma unroll
            for (int u = 0; u < kPP; u++) apply(p[u]);
        }
        }
        // first batch of the next tile: in flight while this tile is scanned and written
#pragma unroll
        for (int u = 0; u < kPP; u++) pf[u] = make_uint2(kNone, kNone);
        if (tn < P.n_tiles) {
            const uint32_t nb = er_n.x >> 1, ne = min(er_n.y >> 1, pair_cap);
#pragma unroll
            for (int u = 0; u < kPP; u++) { const uint32_t i = nb + tid + u * kThreads; if (i < ne) pf[u] = __ldg(pairs + i); }
        }
        mycarry = (int)__reduce_add_sync(0xffffffffu, mycarry);
        if (lane == 0) s_wcar[warp] = mycarry;
        __syncthreads();
        // ---- my PT positions: read once, zero behind (the second barrier orders this zeroing before any thread
        // starts the next tile's atomics)
        int v[PT];
        int4* mine = reinterpret_cast<int4*>(s_diff + tid * kPadW);
        int tot = 0;
#pragma unroll
        for (int i = 0; i < PT / 4; i++) {
            const int4 x = mine[i];
            mine[i] = make_int4(0, 0, 0, 0);
            v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
            tot += (x.x + x.y) + (x.z + x.w);
        }
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += x; }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        // depth at the tile start + everything in the warps before mine, in one redux
        int part = 0;
        if (lane < (uint32_t)kWarps) part = s_wcar[lane] + (lane < warp ? s_wtot[lane] : 0);
        const int woff = (int)__reduce_add_sync(0xffffffffu, part);
        int carry_in = 0;
        if (lane < (uint32_t)kWarps) carry_in = s_wcar[lane];
        carry_in = (int)__reduce_add_sync(0xffffffffu, carry_in);
        int run = woff + incl - tot;
        // no depth in the tile exceeds carry_in + #events: PT of them fit a 32-bit sum unless that bound is absurd
        const bool wide = (unsigned long long)(uint32_t)carry_in + (er.y - er.x) >= (1ull << 26);
        const uint32_t q0 = tid * PT;
        const uint32_t cnt = q0 >= n_here ? 0u : (n_here - q0 < (uint32_t)PT ? n_here - q0 : (uint32_t)PT);
        uint32_t sum32 = 0, mn = 0xffffffffu, nz;
#pragma unroll
        for (int i = 0; i < PT; i++) { run += v[i]; v[i] = run; sum32 += (uint32_t)run; mn = min(mn, (uint32_t)run); }
        // ---- write my positions
        uint32_t* out = P.depth + (size_t)t * kTile + q0;
        if (cnt == (uint32_t)PT) {
#pragma unroll
            for (int i = 0; i < PT / 8; i++)
                st_na_v8(out + 8 * i, v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
        } else {
#pragma unroll
            for (int i = 0; i < PT; i++) if ((uint32_t)i < cnt) out[i] = (uint32_t)v[i];
        }
        unsigned long long sum = sum32;
        nz = cnt;
        if (cnt != (uint32_t)PT || mn == 0u || wide) {                    // rare: ragged tile end, zero depth, absurd depth
            sum = 0; nz = 0;
#pragma unroll
            for (int i = 0; i < PT; i++) if ((uint32_t)i < cnt) { sum += (uint32_t)v[i]; nz += v[i] != 0; }
        }
        // warp totals with the redux unit: the 32-bit thread sums are split in 16-bit halves so they cannot overflow
        if (__all_sync(0xffffffffu, sum <= 0xffffffffull)) {
            const uint32_t s32 = (uint32_t)sum;
            sum = (unsigned long long)__reduce_add_sync(0xffffffffu, s32 & 0xffffu) + ((unsigned long long)__reduce_add_sync(0xffffffffu, s32 >> 16) << 16);
        } else sum = warp_sum_u64(sum);
        nz = __reduce_add_sync(0xffffffffu, nz);
        if (lane == 0 && (sum | nz)) { atomicAdd(&P.tile_sum[t], sum); atomicAdd(&P.tile_nz[t], nz); }
        t = tn; desc = desc_n; er = er_n;
    }
}

// ---------------------------------------------------------------- 16-bit counters
constexpr uint32_t kBias2 = 0x80008000u;       // both halves of a counter word at zero

__device__ __forceinline__ int dp2(uint32_t w, int sel, int c) { return __dp2a_lo((int)w, sel, c); }   // c + lo*sel.b0 + hi*sel.b1

// predicated shared-memory reduction: no branch, no generic-address arithmetic
__device__ __forceinline__ void red_shared_if_lt(uint32_t q, uint32_t n, uint32_t saddr, uint32_t val)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], %3;\n\t}" ::"r"(q), "r"(n), "r"(saddr), "r"(val) : "memory");
}
// one step of an inclusive warp scan: SHFL + predicated add
__device__ __forceinline__ int scan_step(int x, int d)
{
    asm volatile("{\n\t.reg .s32 t;\n\t.reg .pred p;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.s32 %0, %0, t;\n\t}" : "+r"(x) : "r"(d));
    return x;
}

// Searched path of the 16-bit kernel (long records): per record only the pairs that can touch the tile.  Returns the
// thread's share of the tile's carry-in.  Out of line on purpose: the streaming path keeps its registers.
__device__ __noinline__ int tile16_searched(const uint32_t* __restrict__ events, const uint32_t* __restrict__ ev_start, const uint32_t pair_cap,
                                           const uint2 rr, const uint32_t T0, const uint32_t n_here, const uint32_t sbase, uint2* s_rng)
{
    constexpr int kThreads = 256, kWarps = 8;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint2* __restrict__ pairs = reinterpret_cast<const uint2*>(events);
    int mycarry = 0;
    for (uint32_t rb = rr.x; rb < rr.y; rb += kThreads) {
        const uint32_t cnt = rr.y - rb < (uint32_t)kThreads ? rr.y - rb : (uint32_t)kThreads;
        __syncthreads();
        if (tid < cnt) s_rng[tid] = record_pair_range(events, ev_start, rb + tid, T0, T0 + n_here);
        __syncthreads();
        for (uint32_t r = warp; r < cnt; r += kWarps) {
            const uint2 g2 = s_rng[r];
            const uint32_t pend = g2.y < pair_cap ? g2.y : pair_cap;
            for (uint32_t q = g2.x + lane; q < pend; q += 32 * 8) {        // 8 loads in flight per lane
                uint2 pr[8];
#pragma unroll
                for (int u = 0; u < 8; u++) pr[u] = q + 32 * u < pend ? __ldg(pairs + q + 32 * u) : make_uint2(kNone, kNone);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t q0 = pr[u].x - T0, q1 = pr[u].y - T0;
                    mycarry += (pr[u].x < T0) ? 1 : 0;
                    mycarry -= (pr[u].y < T0) ? 1 : 0;
                    red_shared_if_lt(q0, n_here, (sbase + 2u * q0) & ~3u, (q0 & 1u) * 0x0000ffffu + 1u);
                    red_shared_if_lt(q1, n_here, (sbase + 2u * q1) & ~3u, (q1 & 1u) * 0xffff0001u + 0xffffffffu);
                }
            }
        }
    }
    return mycarry;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_depth_tiles16(const TileParams P)
{
    constexpr int kThreads = 256, kWarps = kThreads / 32, kPP = 1024 / kThreads;
    static_assert(kTile == kWarps * 1024, "a warp owns 1024 positions: 4 rows of 32 chunks of 8");
    __shared__ __align__(16) uint32_t s_d[kTile / 2];
    // s_wcar_ is double-buffered by tile parity: a warp that runs ahead (the stores of the others are held up by HBM
    // back-pressure for thousands of cycles) writes its share of the NEXT tile's carry-in before the first barrier of that
    // tile, while a slow warp may not have read this tile's yet after the second barrier.  (Seen on B200 as +-k over one
    // warp's 1024 positions in about one tile in 10^5.)  s_wtot is written between the two barriers: no such window.
    __shared__ int s_wtot[kWarps], s_wcar_[2][kWarps];
    __shared__ uint2 s_rng[kThreads];                                    // searched path: pair range of 256 records at a time
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint2* __restrict__ pairs = reinterpret_cast<const uint2*>(P.events);
    const uint32_t pair_cap = P.ev_cap >> 1;
    const uint4 bias4 = make_uint4(kBias2, kBias2, kBias2, kBias2);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_d);

    for (uint32_t i = tid; i < kTile / 8; i += kThreads) reinterpret_cast<uint4*>(s_d)[i] = bias4;
    // tile descriptors {T0, positions, first event, end event} run two tiles ahead of the tile being scanned
    const uint32_t g = gridDim.x;
    uint32_t t = P.t_begin + blockIdx.x, parity = 0;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    uint4 cur = t < P.t_end ? __ldg(P.tile_q + t) : zero4;
    uint4 nxt = t + g < P.t_end ? __ldg(P.tile_q + t + g) : zero4;
    uint2 pf[kPP];
    {
        const uint32_t pb = cur.z >> 1, pe = min(cur.w >> 1, pair_cap);
#pragma unroll
        for (int u = 0; u < kPP; u++) { const uint32_t i = pb + tid + u * kThreads; pf[u] = i < pe ? __ldg(pairs + i) : make_uint2(kNone, kNone); }
    }
    __syncthreads();

    while (t < P.t_end) {
        const uint4 nn = t + 2 * g < P.t_end ? __ldg(P.tile_q + t + 2 * g) : zero4;
        const uint32_t T0 = cur.x, n_here = cur.y;
        const uint32_t pb = cur.z >> 1, pe = min(cur.w >> 1, pair_cap);
        const uint32_t np_slice = (cur.w - cur.z) >> 1;
        const bool searched = np_slice > kSearchMinPairs;                // CTA-uniform
        uint2 rr = make_uint2(0, 0);
        if (searched) rr = __ldg(P.tile_r + t);
        const bool narrow = !searched || rr.y - rr.x <= kNarrowMaxRecords;   // CTA-uniform; wide tiles belong to k_depth_tiles_wide
        // ---- stretches of the records that overlap the tile (every counter is at its bias here)
        int mycarry = 0;
        auto apply = [&](const uint2 pr) {
            const uint32_t q0 = pr.x - T0, q1 = pr.y - T0;         // wrap to huge values left of the tile
            mycarry += (pr.x < T0) ? 1 : 0;
            mycarry -= (pr.y < T0) ? 1 : 0;
            red_shared_if_lt(q0, n_here, (sbase + 2u * q0) & ~3u, (q0 & 1u) * 0x0000ffffu + 1u);             // +1 in its half
            red_shared_if_lt(q1, n_here, (sbase + 2u * q1) & ~3u, (q1 & 1u) * 0xffff0001u + 0xffffffffu);     // -1 in its half
        };
        if (searched && narrow) {
            mycarry = tile16_searched(P.events, P.ev_start, pair_cap, rr, T0, n_here, sbase, s_rng);  // long records: kept out of line, the streaming path stays lean
        } else if (narrow) {
#pragma unroll
            for (int u = 0; u < kPP; u++) apply(pf[u]);
            for (uint32_t base = pb + kThreads * kPP; base < pe; base += kThreads * kPP) {      // > 2048 events
                uint2 p[kPP];
#pragma unroll
                for (int u = 0; u < kPP; u++) { const uint32_t i = base + tid + u * kThreads; p[u] = i < pe ? __ldg(pairs + i) : make_uint2(kNone, kNone); }
#pragma unroll
                for (int u = 0; u < kPP; u++) apply(p[u]);
            }
        }
        // first batch of the next tile: in flight while this tile is scanned and written
        {
            const uint32_t nb = nxt.z >> 1, ne = min(nxt.w >> 1, pair_cap);
#pragma unroll
            for (int u = 0; u < kPP; u++) { const uint32_t i = nb + tid + u * kThreads; pf[u] = i < ne ? __ldg(pairs + i) : make_uint2(kNone, kNone); }
        }
        if (!narrow) { t += g; cur = nxt; nxt = nn; continue; }
        int* const s_wcar = s_wcar_[parity];                               // flips once per tile that passes the two barriers below
        parity ^= 1u;
        mycarry = (int)__reduce_add_sync(0xffffffffu, mycarry);
        if (lane == 0) s_wcar[warp] = mycarry;
        __syncthreads();
        // ---- my four chunks: read once, reset behind (the second barrier orders the reset before any thread
        // starts the next tile's atomics)
        uint4* rowp = reinterpret_cast<uint4*>(s_d) + warp * 128 + lane;
        uint32_t w[16];
        int tot[4], inc[4], R[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 x = rowp[32 * i];
            rowp[32 * i] = bias4;
            w[4 * i] = x.x ^ kBias2; w[4 * i + 1] = x.y ^ kBias2; w[4 * i + 2] = x.z ^ kBias2; w[4 * i + 3] = x.w ^ kBias2;
            tot[i] = dp2(w[4 * i + 3], 0x0101, dp2(w[4 * i + 2], 0x0101, dp2(w[4 * i + 1], 0x0101, dp2(w[4 * i], 0x0101, 0))));
        }
        // Two rows per scan: every partial sum of counters is bounded by the records overlapping the tile (<= 32767),
        // so the integer a + 65536 b carries both exactly
        int s01 = tot[0] + (tot[1] << 16), s23 = tot[2] + (tot[3] << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { s01 = scan_step(s01, d); s23 = scan_step(s23, d); }
        inc[0] = (int)(short)s01; inc[1] = (s01 - inc[0]) >> 16;
        inc[2] = (int)(short)s23; inc[3] = (s23 - inc[2]) >> 16;
        const int r01 = __shfl_sync(0xffffffffu, s01, 31), r23 = __shfl_sync(0xffffffffu, s23, 31);
        R[0] = (int)(short)r01; R[1] = (r01 - R[0]) >> 16;
        R[2] = (int)(short)r23; R[3] = (r23 - R[2]) >> 16;
        if (lane == 0) s_wtot[warp] = (R[0] + R[1]) + (R[2] + R[3]);
        __syncthreads();
        // depth at the tile start + everything in the warps before mine, in one redux
        int part = 0;
        if (lane < (uint32_t)kWarps) part = s_wcar[lane] + (lane < warp ? s_wtot[lane] : 0);
        const int woff = (int)__reduce_add_sync(0xffffffffu, part);
        int base[4];
        base[0] = woff + inc[0] - tot[0];
        base[1] = woff + R[0] + inc[1] - tot[1];
        base[2] = woff + R[0] + R[1] + inc[2] - tot[2];
        base[3] = woff + R[0] + R[1] + R[2] + inc[3] - tot[3];
        // depths never exceed the records overlapping the tile (<= 32767) here: 32-bit sums of 32 of them cannot overflow
        const uint32_t q_first = warp * 1024u + lane * 8u;                 // my chunk of row 0; row i is 256 positions further
        uint32_t* out = P.depth + (size_t)t * kTile + q_first;
        uint32_t sum32 = 0, mn = 0xffffffffu;
        const bool full = q_first + 3u * 256u + 8u <= n_here;              // all four chunks inside the tile
        if (full) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int v[8], run = base[i];
#pragma unroll
                for (int k = 0; k < 4; k++) { v[2 * k] = dp2(w[4 * i + k], 0x0001, run); run = v[2 * k + 1] = dp2(w[4 * i + k], 0x0101, run); }
                st_na_v8(out + 256 * i, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
#pragma unroll
                for (int k = 0; k < 8; k++) { sum32 += (uint32_t)v[k]; mn = min(mn, (uint32_t)v[k]); }
            }
        }
        uint32_t nz = 32;
        if (!full || mn == 0u) {                                            // ragged tile end or zero depth: count exactly
            sum32 = 0; nz = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int v[8], run = base[i];
#pragma unroll
                for (int k = 0; k < 4; k++) { v[2 * k] = dp2(w[4 * i + k], 0x0001, run); run = v[2 * k + 1] = dp2(w[4 * i + k], 0x0101, run); }
                const uint32_t qc = q_first + 256u * i;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (qc + k < n_here) { if (!full) out[256 * i + k] = (uint32_t)v[k]; sum32 += (uint32_t)v[k]; nz += v[k] != 0; }
            }
        }
        // warp totals with the redux unit: thread sums <= 32 * 32767 < 2^20, so the warp sum fits 32 bits
        sum32 = __reduce_add_sync(0xffffffffu, sum32);
        nz = __reduce_add_sync(0xffffffffu, nz);
        if (lane == 0 && (sum32 | nz)) { atomicAdd(&P.tile_sum[t], (unsigned long long)sum32); atomicAdd(&P.tile_nz[t], nz); }
        t += g; cur = nxt; nxt = nn;
    }
}

// sum the per-tile partials per region (cnv_caller.cpp:534-535): one thread per tile, one atomic per warp and region
__global__ void __launch_bounds__(256) k_region_stats(const uint4* __restrict__ tile_desc, uint32_t n_tiles, const unsigned long long* __restrict__ tile_sum,
                                                       const uint32_t* __restrict__ tile_nz, unsigned long long* reg_sum, uint32_t* reg_nz)
{
    for (uint32_t t0 = blockIdx.x * blockDim.x; t0 < n_tiles; t0 += gridDim.x * blockDim.x) {
        const uint32_t t = t0 + threadIdx.x;
        const bool ok = t < n_tiles;
        const uint32_t r = ok ? tile_desc[t].x : 0xffffffffu;
        unsigned long long s = ok ? tile_sum[t] : 0ull;
        uint32_t n = ok ? tile_nz[t] : 0u;
        const uint32_t r0 = __shfl_sync(0xffffffffu, r, 0);
        if (__all_sync(0xffffffffu, r == r0)) {                               // the common case: a warp inside one region
            s = warp_sum_u64(s); n = warp_sum_u32(n);
            if ((threadIdx.x & 31) == 0 && (s | n)) { atomicAdd(&reg_sum[r0], s); atomicAdd(&reg_nz[r0], n); }
        } else if (ok && (s | n)) { atomicAdd(&reg_sum[r], s); atomicAdd(&reg_nz[r], n); }
    }
}

static TileParams tile_params(csv_batch* b)
{
    TileParams P;
    P.tile_desc = b->d_tile_desc.as<uint4>();
    P.tile_ev = b->d_tile_ev.as<uint2>();
    P.tile_q = b->d_tile_q.as<uint4>();
    P.tile_r = b->d_tile_r.as<uint2>();
    P.ev_start = b->d_ev_start.as<uint32_t>();
    P.events = b->d_events.as<uint32_t>();
    P.ev_cap = (uint32_t)b->ev_cap;
    P.depth = b->d_depth.as<uint32_t>();
    P.tile_sum = b->d_tile_sum.as<unsigned long long>();
    P.tile_nz = b->d_tile_nz.as<uint32_t>();
    P.n_tiles = b->n_tiles;
    P.t_begin = 0; P.t_end = b->n_tiles;
    P.wide_list = b->d_wide_list.as<uint32_t>();
    P.scalars = b->d_scalars.as<uint32_t>();
    return P;
}

// before the first chunk of a pass
int launch_depth_begin(csv_ctx* ctx, csv_batch* b)
{
    if (b->n_tiles == 0) return CSV_OK;
    CSV_CUDA(cudaMemsetAsync(b->d_tile_sum.p, 0, (size_t)b->n_tiles * 8, ctx->stream));
    CSV_CUDA(cudaMemsetAsync(b->d_tile_nz.p, 0, (size_t)b->n_tiles * 4, ctx->stream));
    CSV_CUDA(cudaMemsetAsync(b->d_sum.p, 0, (size_t)b->n_regions * 8, ctx->stream));
    CSV_CUDA(cudaMemsetAsync(b->d_nz.p, 0, (size_t)b->n_regions * 4, ctx->stream));
    return CSV_OK;
}

// the tiles of pipeline chunk c
int launch_depth_tiles(csv_ctx* ctx, csv_batch* b, uint32_t c)
{
    TileParams P = tile_params(b);
    static const int mult = getenv("CSV_TILE_GRID") ? atoi(getenv("CSV_TILE_GRID")) : 96;   // tuning knob: CTAs per SM in the grid
    static const int minb = getenv("CSV_TILE_MINB") ? atoi(getenv("CSV_TILE_MINB")) : 4;
    for (const auto& tr : b->chunks[c].tiles) {
        P.t_begin = tr.first; P.t_end = tr.second;
        const uint32_t nt = tr.second - tr.first;
        const uint32_t grid = nt < (uint32_t)ctx->sm_count * mult ? nt : (uint32_t)ctx->sm_count * mult;
        StageTimer tk(ctx, ST_TILE_KERNEL);                                                 // the dominant kernel alone (bench.py's roofline)
        if (minb == 5) k_depth_tiles16<5><<<grid, 256, 0, ctx->stream>>>(P);
        else if (minb == 6) k_depth_tiles16<6><<<grid, 256, 0, ctx->stream>>>(P);
        else k_depth_tiles16<4><<<grid, 256, 0, ctx->stream>>>(P);
        ctx->launches++;
    }
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// after the last chunk: pile-up tiles with 32-bit counters, then the per-region reductions
// what: 1 = the pile-up tiles (32-bit counters; needs the tile ranges of every chunk), 2 = the per-region reductions
// (needs every tile), 3 = both, in this order on the current stream
int launch_depth_finish(csv_ctx* ctx, csv_batch* b, int what)
{
    if (b->n_tiles == 0) return CSV_OK;
    TileParams P = tile_params(b);
    if (what & 1) {
        const uint32_t grid_w = b->n_tiles < (uint32_t)ctx->sm_count * 4 ? b->n_tiles : (uint32_t)ctx->sm_count * 4;
        k_depth_tiles_wide<32><<<grid_w, kTile / 32, 0, ctx->stream>>>(P);    // exits at once when the wide list is empty
        ctx->launches++;
    }
    if (!(what & 2)) { CSV_CUDA(cudaGetLastError()); return CSV_OK; }
    const uint32_t grid_r = (b->n_tiles + 255) / 256 < (uint32_t)ctx->sm_count * 8 ? (b->n_tiles + 255) / 256 : (uint32_t)ctx->sm_count * 8;
    k_region_stats<<<grid_r, 256, 0, ctx->stream>>>(P.tile_desc, b->n_tiles, P.tile_sum, P.tile_nz, b->d_sum.as<unsigned long long>(), b->d_nz.as<uint32_t>());
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

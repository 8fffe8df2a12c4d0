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
// k_depth_tiles: one CTA owns one tile.  It zeroes a kTile-int difference array
//   in shared memory and streams the slice: an event left of the tile adds its
//   sign to the tile's carry-in (depth at T0), an event inside goes into the
//   difference array with a shared-memory atomic, the rest is skipped.  Sign =
//   parity of the event slot.  No inter-tile dependency, no global atomics.
//   Then prefix sum + carry, every depth word is written exactly once with
//   128-bit streaming stores, and sum(depth) / count(depth > 0) are reduced per
//   tile and summed per region by k_region_stats.
#include "batch.cuh"
#include "scan.cuh"

#include <cstdlib>

namespace csv {

// ------------------------------------------------------- prefix max over records
constexpr int kPmThreads = 256, kPmItems = 8, kPmTile = kPmThreads * kPmItems;

__device__ __forceinline__ unsigned long long pm_value(const unsigned long long* key, const uint32_t* ref_end, uint32_t k)
{
    return (key[k] & 0xffffffff00000000ull) | ref_end[k];
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

__global__ void __launch_bounds__(kPmThreads) k_pm_partials(const unsigned long long* __restrict__ meta, const uint32_t* __restrict__ ref_end,
                                                            const uint32_t* scalars, unsigned long long* part)
{
    __shared__ unsigned long long s[kPmThreads / 32];
    const uint32_t n = scalars[SC_N_NONEMPTY];
    for (uint32_t t = blockIdx.x; (uint64_t)t * kPmTile < n; t += gridDim.x) {
        unsigned long long v = 0;
        for (int j = 0; j < kPmItems; j++) {
            const uint64_t k = (uint64_t)t * kPmTile + j * kPmThreads + threadIdx.x;
            if (k < n) v = umax64(v, pm_value(meta, ref_end, (uint32_t)k));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = umax64(v, __shfl_xor_sync(0xffffffffu, v, d));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long r = s[0];
            for (int i = 1; i < kPmThreads / 32; i++) r = umax64(r, s[i]);
            part[t] = r;
        }
    }
}

// single CTA: part[t] <- max of part[0..t-1] (exclusive), in place
__global__ void __launch_bounds__(1024) k_pm_scan_partials(const uint32_t* scalars, unsigned long long* part)
{
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const uint32_t n = scalars[SC_N_NONEMPTY];
    const uint32_t n_part = (uint32_t)(((uint64_t)n + kPmTile - 1) / kPmTile);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_part; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_part ? part[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc = umax64(inc, t); }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (uint32_t w = 0; w < warp; w++) pre = umax64(pre, s_w[w]);
        unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        if (i < n_part) part[i] = umax64(pre, excl);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = umax64(pre, inc);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kPmThreads) k_pm_final(const unsigned long long* __restrict__ meta, const uint32_t* __restrict__ ref_end,
                                                         uint32_t* scalars, const unsigned long long* __restrict__ part,
                                                         unsigned long long* pmax)
{
    __shared__ unsigned long long s_w[kPmThreads / 32];
    const uint32_t n = scalars[SC_N_NONEMPTY];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t = blockIdx.x; (uint64_t)t * kPmTile < n; t += gridDim.x) {
        // blocked arrangement: thread owns kPmItems consecutive records
        const uint64_t k0 = (uint64_t)t * kPmTile + (uint64_t)threadIdx.x * kPmItems;
        unsigned long long v[kPmItems], run = 0;
#pragma unroll
        for (int j = 0; j < kPmItems; j++) {
            v[j] = (k0 + j < n) ? pm_value(meta, ref_end, (uint32_t)(k0 + j)) : 0ull;
            // coordinate order check rides along: (tid, pos0 + 1) must not decrease
            if (k0 + j < n && k0 + j > 0) {
                if (meta[k0 + j - 1] > meta[k0 + j]) scalars[SC_UNSORTED] = 1;
            }
            run = umax64(run, v[j]);
        }
        unsigned long long inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc = umax64(inc, x); }
        __syncthreads();
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        unsigned long long pre = part[t];
        for (uint32_t w = 0; w < warp; w++) pre = umax64(pre, s_w[w]);
        unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        pre = umax64(pre, excl);
#pragma unroll
        for (int j = 0; j < kPmItems; j++) { pre = umax64(pre, v[j]); if (k0 + j < n) pmax[k0 + j] = pre; }
    }
}

// ------------------------------------------------------------ tile -> event slice
__global__ void k_tile_ranges(const uint4* __restrict__ tile_desc, uint32_t n_tiles, const unsigned long long* __restrict__ key,
                              const unsigned long long* __restrict__ pmax, const uint32_t* __restrict__ ev_start,
                              const uint32_t* scalars, uint2* tile_ev)
{
    const uint32_t n = scalars[SC_N_NONEMPTY];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        const uint4 d = tile_desc[t];            // {region, positions, T0, tid}
        const unsigned long long key_hi = ((unsigned long long)d.w << 32) | (unsigned long long)(d.z + d.y);   // (tid, T1)
        const unsigned long long key_lo = ((unsigned long long)d.w << 32) | (unsigned long long)d.z;           // (tid, T0)
        uint32_t lo = 0, hi = n;
        while (lo < hi) {                        // r_hi: first record with (tid, pos0 + 1) >= (tid, T1)
            const uint32_t mid = (lo + hi) >> 1;
            if (key[mid] < key_hi) lo = mid + 1; else hi = mid;
        }
        const uint32_t r_hi = lo;
        lo = 0; hi = r_hi;
        while (lo < hi) {                        // r_lo: first record with running max (tid, ref_end) > (tid, T0)
            const uint32_t mid = (lo + hi) >> 1;
            if (pmax[mid] <= key_lo) lo = mid + 1; else hi = mid;
        }
        const uint32_t r_lo = lo;
        tile_ev[t] = r_lo < r_hi ? make_uint2(ev_start[r_lo], ev_start[r_hi]) : make_uint2(0u, 0u);
    }
}

int launch_tile_ranges(csv_ctx* ctx, csv_batch* b)
{
    if (b->n_tiles == 0) return CSV_OK;
    const unsigned long long* meta = b->d_key.as<unsigned long long>();
    const uint32_t* ref_end = b->d_ref_end.as<uint32_t>();
    uint32_t* scalars = b->d_scalars.as<uint32_t>();
    unsigned long long* part = b->d_pmax_part.as<unsigned long long>();
    unsigned long long* pmax = b->d_pmax.as<unsigned long long>();
    const uint32_t n_part = (uint32_t)(((uint64_t)b->n_reads + kPmTile - 1) / kPmTile);
    if (n_part) {
        const uint32_t grid = n_part < (uint32_t)ctx->sm_count * 8 ? n_part : (uint32_t)ctx->sm_count * 8;
        k_pm_partials<<<grid, kPmThreads, 0, ctx->stream>>>(meta, ref_end, scalars, part);
        k_pm_scan_partials<<<1, 1024, 0, ctx->stream>>>(scalars, part);
        k_pm_final<<<grid, kPmThreads, 0, ctx->stream>>>(meta, ref_end, scalars, part, pmax);
        ctx->launches += 3;
    }
    const uint32_t grid_t = (b->n_tiles + 255) / 256;
    k_tile_ranges<<<grid_t, 256, 0, ctx->stream>>>(b->d_tile_desc.as<uint4>(), b->n_tiles, meta, pmax, b->d_ev_start.as<uint32_t>(),
                                                  scalars, b->d_tile_ev.as<uint2>());
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

// --------------------------------------------------------------------- tile kernel
struct TileParams {
    const uint4* tile_desc;          // static per tile: {region, positions in tile, T0, tid}
    const uint2* tile_ev;            // event slice [x, y)
    const uint32_t* events;
    uint32_t ev_cap;
    uint32_t* depth;                 // n_tiles * kTile words
    unsigned long long* tile_sum;    // per-tile partial reductions (no contended atomics)
    uint32_t* tile_nz;
    uint32_t n_tiles;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 256-bit streaming store (sm_100): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void st_na_v8(uint32_t* p, int a, int b, int c, int d, int e, int f, int g, int h)
{
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
                 "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}

// One CTA per tile of kTile positions, PT consecutive positions per thread (kTile / PT threads).
// Shared-memory rows of PT words are padded by 4 words: 16-byte aligned and conflict-free for the 128-bit
// accesses of their owner (PT = 32: stride 36 words; PT = 16: stride 20 words).
// The kernel is limited by the L1/shared-memory pipe and by latency, not by HBM (profiles/r1_history.md), so
// shared memory is touched as little as possible -- events in (atomics), one 128-bit read + one zeroing write per
// 4 positions -- and the depths go from registers to HBM with 256-bit stores (full 32-byte sectors).
template <int PT>
__global__ void __launch_bounds__(kTile / PT, (PT == 32 ? 5 : 3)) k_depth_tiles(const TileParams P)
{
    constexpr int kThreads = kTile / PT, kWarps = kThreads / 32, kPadW = PT + 4;
    constexpr int kShift = (PT == 32 ? 5 : 4);
    __shared__ __align__(16) int s_diff[kThreads * kPadW];
    __shared__ uint32_t s_scan[40];
    __shared__ int s_wcarry[2][kWarps];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (uint32_t i = tid; i < kThreads * kPadW / 4; i += kThreads) reinterpret_cast<int4*>(s_diff)[i] = make_int4(0, 0, 0, 0);
    uint32_t t = blockIdx.x;
    uint4 desc = make_uint4(0, 0, 0, 0);
    uint2 er = make_uint2(0, 0);
    if (t < P.n_tiles) { desc = __ldg(P.tile_desc + t); er = __ldg(P.tile_ev + t); }
    __syncthreads();

    for (uint32_t it = 0; t < P.n_tiles; it++) {
        // descriptor of the NEXT tile of this CTA: loaded now, used one iteration later
        const uint32_t tn = t + gridDim.x;
        uint4 desc_n = make_uint4(0, 0, 0, 0);
        uint2 er_n = make_uint2(0, 0);
        if (tn < P.n_tiles) { desc_n = __ldg(P.tile_desc + tn); er_n = __ldg(P.tile_ev + tn); }

        const uint32_t n_here = desc.y, T0 = desc.z;
        const uint32_t e1 = er.y < P.ev_cap ? er.y : P.ev_cap;
        // ---- events of the records that overlap the tile (s_diff is all zero here).  All loads of a batch are
        // issued before the first shared-memory atomic: one memory round trip per batch.
        constexpr int kBatch = 2048 / kThreads;                    // 2048 events per batch
        int mycarry = 0;
        const int sgn = 1 - 2 * (int)((er.x + tid) & 1u);          // slot parity; the stride is even
        for (uint32_t base = er.x + tid; base < e1; base += kThreads * kBatch) {
            uint32_t p[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; u++) { const uint32_t e = base + u * kThreads; p[u] = e < e1 ? __ldg(P.events + e) : kNone; }
#pragma unroll
            for (int u = 0; u < kBatch; u++) {
                const uint32_t q = p[u] - T0;                      // wraps to a huge value for events left of the tile
                mycarry += (p[u] < T0) ? sgn : 0;
                if (q < n_here) atomicAdd(&s_diff[q + (q >> kShift) * 4u], sgn);
            }
        }
        mycarry = (int)__reduce_add_sync(0xffffffffu, mycarry);
        if (lane == 0) s_wcarry[it & 1][warp] = mycarry;
        // pull the next tile's event slice towards L2 while this tile is scanned and written
        if (tn < P.n_tiles) {
            const uint32_t pe = er_n.y < P.ev_cap ? er_n.y : P.ev_cap;
            for (uint32_t a = er_n.x + tid * 32u; a < pe; a += kThreads * 32u) prefetch_l2(P.events + a);
        }
        __syncthreads();
        // ---- my PT positions: read once, zero behind (the barriers of the block scan below order this zeroing
        // before any thread starts the next tile's atomics)
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
        uint32_t dummy;
        const uint32_t ex = block_excl_scan_u32((uint32_t)tot, s_scan, &dummy);
        int carry_in = 0;
#pragma unroll
        for (int i = 0; i < kWarps; i++) carry_in += s_wcarry[it & 1][i];
        int run = carry_in + (int)ex;
        // no depth in the tile exceeds carry_in + #events: PT of them fit a 32-bit sum unless that bound is absurd
        const bool wide = (unsigned long long)(uint32_t)carry_in + (e1 - er.x) >= (1ull << 26);
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

// one CTA per region: sum the per-tile partials (cnv_caller.cpp:534-535)
__global__ void __launch_bounds__(256) k_region_stats(const uint32_t* __restrict__ reg_tile_base, const unsigned long long* __restrict__ tile_sum,
                                                       const uint32_t* __restrict__ tile_nz, unsigned long long* reg_sum, uint32_t* reg_nz)
{
    __shared__ unsigned long long s_s[8];
    __shared__ uint32_t s_n[8];
    const uint32_t r = blockIdx.x, t0 = reg_tile_base[r], t1 = reg_tile_base[r + 1];
    unsigned long long s = 0; uint32_t n = 0;
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) { s += tile_sum[t]; n += tile_nz[t]; }
    s = warp_sum_u64(s); n = warp_sum_u32(n);
    if ((threadIdx.x & 31) == 0) { s_s[threadIdx.x >> 5] = s; s_n[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { s += s_s[i]; n += s_n[i]; }
        reg_sum[r] = s; reg_nz[r] = n;
    }
}

int launch_depth_tiles(csv_ctx* ctx, csv_batch* b)
{
    TileParams P;
    P.tile_desc = b->d_tile_desc.as<uint4>();
    P.tile_ev = b->d_tile_ev.as<uint2>();
    P.events = b->d_events.as<uint32_t>();
    P.ev_cap = (uint32_t)b->ev_cap;
    P.depth = b->d_depth.as<uint32_t>();
    P.tile_sum = b->d_tile_sum.as<unsigned long long>();
    P.tile_nz = b->d_tile_nz.as<uint32_t>();
    P.n_tiles = b->n_tiles;
    if (b->n_tiles == 0) return CSV_OK;
    CSV_CUDA(cudaMemsetAsync(P.tile_sum, 0, (size_t)b->n_tiles * 8, ctx->stream));
    CSV_CUDA(cudaMemsetAsync(P.tile_nz, 0, (size_t)b->n_tiles * 4, ctx->stream));
    static const int pt = getenv("CSV_TILE_PT") ? atoi(getenv("CSV_TILE_PT")) : 32;     // tuning knob: positions per thread
    if (pt == 32) {
        uint32_t grid = b->n_tiles < (uint32_t)ctx->sm_count * 20 ? b->n_tiles : (uint32_t)ctx->sm_count * 20;
        k_depth_tiles<32><<<grid, kTile / 32, 0, ctx->stream>>>(P);
    } else {
        uint32_t grid = b->n_tiles < (uint32_t)ctx->sm_count * 12 ? b->n_tiles : (uint32_t)ctx->sm_count * 12;
        k_depth_tiles<16><<<grid, kTile / 16, 0, ctx->stream>>>(P);
    }
    k_region_stats<<<b->n_regions, 256, 0, ctx->stream>>>(b->d_reg_tab.as<uint32_t>(), P.tile_sum, P.tile_nz,
                                                          b->d_sum.as<unsigned long long>(), b->d_nz.as<uint32_t>());
    ctx->launches += 2;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

// scan.cuh -- single-pass chained ("decoupled look-back") prefix sums.
//
// One launch per scan: tiles are claimed through a ticket counter so every
// predecessor of a tile is already running, each tile publishes
// (epoch|flag, value) as ONE 64-bit word -- no fences, no per-launch clearing
// (stale epochs read as "not ready").
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
namespace csv {

constexpr uint32_t kFlagAgg = 1, kFlagPrefix = 2;

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, uint32_t flag, uint32_t v)
{
    return ((unsigned long long)((epoch << 2) | flag) << 32) | v;
}

// Called by one full warp.  Publishes `aggregate` for tile t, walks back over the
// predecessors and returns the exclusive prefix of tile t; finally publishes the
// inclusive prefix.  Values wrap modulo 2^32.
__device__ __forceinline__ uint32_t lookback_u32(unsigned long long* status, uint32_t t, uint32_t aggregate, uint32_t epoch)
{
    const uint32_t lane = lane_id();
    if (lane == 0) st_volatile_u64(&status[t], lb_pack(epoch, t == 0 ? kFlagPrefix : kFlagAgg, aggregate));
    if (t == 0) return 0;
    uint32_t excl = 0;
    int64_t look = (int64_t)t - 1;
    for (;;) {
        int64_t idx = look - lane;
        uint32_t flag, val;
        do {
            if (idx >= 0) {
                unsigned long long w = ld_volatile_u64(&status[idx]);
                uint32_t hi = (uint32_t)(w >> 32);
                flag = ((hi >> 2) == epoch) ? (hi & 3u) : 0u;
                val = (uint32_t)w;
            } else { flag = kFlagPrefix; val = 0; }
        } while (__any_sync(0xffffffffu, flag == 0));
        uint32_t pm = __ballot_sync(0xffffffffu, flag == kFlagPrefix);
        if (pm) {
            uint32_t j = __ffs(pm) - 1;
            excl += warp_sum_u32(lane <= j ? val : 0u);
            break;
        }
        excl += warp_sum_u32(val);
        look -= 32;
    }
    if (lane == 0) st_volatile_u64(&status[t], lb_pack(epoch, kFlagPrefix, excl + aggregate));
    return excl;
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// out(i, exclusive_prefix, value) is called for every i < n; in(i) yields the u32 addend.
// total_out (optional) receives the sum of all n values.
// Global accesses are STRIPED (consecutive threads take consecutive elements: in() and out() usually read and write
// arrays indexed by the element, and a thread that took kScanItems consecutive ones turned every warp load into 32
// sectors -- the record scan of the walk's pre-pass was bound by L1 wavefronts, not by its look-back); the scan itself
// wants kScanItems consecutive elements per thread, so values and prefixes change hands in shared memory (rows padded by
// one word per kScanItems: the blocked accesses are conflict-free).
__device__ __forceinline__ uint32_t scan_pad(uint32_t i) { return i + i / kScanItems; }

template <class In, class Out>
__global__ void __launch_bounds__(kScanThreads) k_chained_scan(In in, Out out, const uint32_t* n_dev, uint64_t n_host,
                                                               uint32_t* ticket, unsigned long long* status, uint32_t epoch,
                                                               uint32_t* total_out, const uint32_t* run_if)
{
    __shared__ uint32_t s_x[kScanTile + kScanTile / kScanItems];
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_tile, s_excl;
    if (run_if && *run_if == 0u) return;            // optional device-side switch: the whole launch is a no-op
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_host;
    const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    if (n == 0) { if (total_out && blockIdx.x == 0 && threadIdx.x == 0) *total_out = 0; return; }
    for (;;) {
        __syncthreads();                             // s_x, s_tile of the tile before
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= n_tiles) break;
        const uint64_t tile0 = (uint64_t)t * kScanTile;
        uint32_t vs[kScanItems];                     // my striped elements: tile0 + j * kScanThreads + threadIdx.x
#pragma unroll
        for (int j = 0; j < kScanItems; j++) {
            const uint32_t e = j * kScanThreads + threadIdx.x;
            vs[j] = (tile0 + e < n) ? in(tile0 + e) : 0u;
            s_x[scan_pad(e)] = vs[j];
        }
        __syncthreads();
        const uint32_t b0 = threadIdx.x * kScanItems;   // my blocked elements: tile0 + b0 + j
        uint32_t v[kScanItems], sum = 0;
#pragma unroll
        for (int j = 0; j < kScanItems; j++) { v[j] = s_x[scan_pad(b0 + j)]; sum += v[j]; }
        uint32_t tile_total;
        uint32_t ex = block_excl_scan_u32(sum, s_scan, &tile_total);
        if (threadIdx.x < 32) {
            uint32_t e = lookback_u32(status, t, tile_total, epoch);
            if (threadIdx.x == 0) {
                s_excl = e;
                if (total_out && t == n_tiles - 1) *total_out = e + tile_total;
            }
        }
        __syncthreads();
        ex += s_excl;
#pragma unroll
        for (int j = 0; j < kScanItems; j++) { s_x[scan_pad(b0 + j)] = ex; ex += v[j]; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kScanItems; j++) {
            const uint32_t e = j * kScanThreads + threadIdx.x;
            if (tile0 + e < n) out(tile0 + e, s_x[scan_pad(e)], vs[j]);
        }
    }
}

template <class In, class Out>
int chained_scan(csv_ctx* ctx, In in, Out out, uint64_t n_upper, const uint32_t* n_dev, uint32_t* total_out, const uint32_t* run_if = nullptr)
{
    uint64_t tiles = (n_upper + kScanTile - 1) / kScanTile;
    if (tiles == 0) tiles = 1;
    CSV_TRY(ensure_status(ctx, tiles));
    uint32_t* ticket;
    CSV_TRY(next_ticket(ctx, &ticket));
    uint32_t epoch = next_epoch(ctx);
    const uint64_t cap = (uint64_t)ctx->sm_count * grid_mult(ctx, 8);
    uint64_t grid = cap_grid(ctx, tiles < cap ? tiles : cap);
    k_chained_scan<<<(unsigned)grid, kScanThreads, 0, ctx->stream>>>(in, out, n_dev, n_upper, ticket,
                                                                    ctx->scan_status.as<unsigned long long>(), epoch, total_out, run_if);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv
#endif

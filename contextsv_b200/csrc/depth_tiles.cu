// depth_tiles.cu -- per-base read depth from binned difference events
// (cnv_caller.cpp:507-519 and the two reductions at :534-535).
//
// The depth map of every region is cut into tiles of kTile positions.  One CTA
// owns one tile: it zeroes a kTile-int difference array in shared memory,
// applies the tile's +-1 events with shared-memory atomics, prefix-sums it
// (carry-in = net sign of all events in earlier tiles of the region) and
// writes every depth word exactly once with 128-bit streaming stores, while
// accumulating sum(depth) and count(depth > 0) for the region.
#include "batch.cuh"
#include "scan.cuh"

namespace csv {

int launch_tile_scan(csv_ctx* ctx, csv_batch* b)
{
    const unsigned long long* cn = b->d_tile_cn.as<unsigned long long>();
    uint32_t* off = b->d_tile_off.as<uint32_t>();
    uint32_t* net = b->d_tile_net.as<uint32_t>();
    // event slot bases: exclusive sum of the per-tile counts (low halves)
    CSV_TRY(chained_scan(ctx,
                         [=] __device__(uint64_t i) -> uint32_t { return (uint32_t)cn[i]; },
                         [=] __device__(uint64_t i, uint32_t ex, uint32_t) { off[i] = ex; },
                         b->n_tiles, nullptr, b->d_scalars.as<uint32_t>() + SC_EV_TOTAL));
    // carry-in: exclusive sum of the per-tile net signs (high halves), modulo 2^32
    CSV_TRY(chained_scan(ctx,
                         [=] __device__(uint64_t i) -> uint32_t { return (uint32_t)(cn[i] >> 32); },
                         [=] __device__(uint64_t i, uint32_t ex, uint32_t) { net[i] = ex; },
                         b->n_tiles, nullptr, nullptr));
    return CSV_OK;
}

constexpr int kTileThreads = 256;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kWarpChunk = kTile / kTileWarps;     // positions per warp
constexpr int kRows = kWarpChunk / 128;            // 128 positions (one int4 per lane) per row

struct TileParams {
    const uint32_t* reg_tile_base;   // caller order, n_regions + 1
    const uint32_t* reg_len;         // caller order: end - beg
    uint32_t n_regions;
    const uint32_t* tile_end;        // after the scatter walk: end of each tile's event run
    const uint32_t* tile_net;
    const uint16_t* events;
    uint32_t ev_cap;
    uint32_t* depth;                 // n_tiles * kTile words
    unsigned long long* reg_sum;
    uint32_t* reg_nz;
    uint32_t n_tiles;
};

__global__ void __launch_bounds__(kTileThreads) k_depth_tiles(const TileParams P)
{
    __shared__ __align__(16) int s_diff[kTile];
    __shared__ int s_wtot[kTileWarps];
    __shared__ uint32_t s_reg;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        // region of this tile: last r with reg_tile_base[r] <= t
        if (tid == 0) {
            uint32_t lo = 0, hi = P.n_regions;
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (P.reg_tile_base[mid] <= t) lo = mid; else hi = mid; }
            s_reg = lo;
        }
        int4* z = reinterpret_cast<int4*>(s_diff);
#pragma unroll
        for (int i = 0; i < kTile / 4 / kTileThreads; i++) z[tid + i * kTileThreads] = make_int4(0, 0, 0, 0);
        __syncthreads();
        const uint32_t r = s_reg;
        const uint32_t tb = P.reg_tile_base[r];
        const uint32_t p0 = (t - tb) << kTileShift;                  // first position of the tile inside the region
        const uint32_t reg_len = P.reg_len[r];
        const uint32_t n_here = reg_len - p0 < (uint32_t)kTile ? reg_len - p0 : (uint32_t)kTile;
        uint32_t e0 = t ? P.tile_end[t - 1] : 0u, e1 = P.tile_end[t];
        if (e1 > P.ev_cap) e1 = P.ev_cap;
        for (uint32_t e = e0 + tid; e < e1; e += kTileThreads) {
            const uint32_t ev = P.events[e];
            atomicAdd(&s_diff[ev & 0x7fffu], (ev & 0x8000u) ? -1 : 1);
        }
        __syncthreads();
        // pass A: per-warp totals
        const int4* row = reinterpret_cast<const int4*>(s_diff + warp * kWarpChunk);
        int tot = 0;
#pragma unroll
        for (int i = 0; i < kRows; i++) { int4 v = row[i * 32 + lane]; tot += v.x + v.y + v.z + v.w; }
        tot = (int)warp_sum_u32((uint32_t)tot);
        if (lane == 0) s_wtot[warp] = tot;
        __syncthreads();
        int carry = (int)(P.tile_net[t] - P.tile_net[tb]);
        for (uint32_t i = 0; i < warp; i++) carry += s_wtot[i];
        // pass B: scan rows, write, reduce
        unsigned long long sum = 0; uint32_t nz = 0;
        uint32_t* out = P.depth + (size_t)t * kTile + warp * kWarpChunk;
        const uint32_t wbase = warp * kWarpChunk;
#pragma unroll
        for (int i = 0; i < kRows; i++) {
            int4 v = row[i * 32 + lane];
            v.y += v.x; v.z += v.y; v.w += v.z;
            int incl = (int)warp_incl_scan_u32((uint32_t)v.w);
            int base = carry + incl - v.w;
            v.x += base; v.y += base; v.z += base; v.w += base;
            carry += __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t pos = wbase + i * 128 + lane * 4;
            if (pos + 4 <= n_here) {
                st_cs_v4(out + i * 128 + lane * 4, make_uint4((uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w));
                sum += (unsigned long long)(uint32_t)v.x + (uint32_t)v.y + (uint32_t)v.z + (uint32_t)v.w;
                nz += (v.x > 0) + (v.y > 0) + (v.z > 0) + (v.w > 0);
            } else if (pos < n_here) {
                const int vv[4] = {v.x, v.y, v.z, v.w};
                for (uint32_t q = 0; q < 4 && pos + q < n_here; q++) {
                    out[i * 128 + lane * 4 + q] = (uint32_t)vv[q];
                    sum += (uint32_t)vv[q]; nz += vv[q] > 0;
                }
            }
        }
        sum = warp_sum_u64(sum); nz = warp_sum_u32(nz);
        if (lane == 0 && (sum | nz)) { atomicAdd(&P.reg_sum[r], sum); atomicAdd(&P.reg_nz[r], nz); }
        __syncthreads();
    }
}

int launch_depth_tiles(csv_ctx* ctx, csv_batch* b)
{
    TileParams P;
    P.reg_tile_base = b->d_reg_tab.as<uint32_t>();
    P.reg_len = b->d_reg_tab.as<uint32_t>() + b->n_regions + 1;
    P.n_regions = b->n_regions;
    P.tile_end = b->d_tile_off.as<uint32_t>();
    P.tile_net = b->d_tile_net.as<uint32_t>();
    P.events = b->d_events.as<uint16_t>();
    P.ev_cap = (uint32_t)b->ev_cap;
    P.depth = b->d_depth.as<uint32_t>();
    P.reg_sum = b->d_sum.as<unsigned long long>();
    P.reg_nz = b->d_nz.as<uint32_t>();
    P.n_tiles = b->n_tiles;
    if (b->n_tiles == 0) return CSV_OK;
    uint32_t grid = b->n_tiles < (uint32_t)ctx->sm_count * 24 ? b->n_tiles : (uint32_t)ctx->sm_count * 24;
    k_depth_tiles<<<grid, kTileThreads, 0, ctx->stream>>>(P);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

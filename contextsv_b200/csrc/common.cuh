// common.cuh -- shared host/device plumbing of the sm_100a scan path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "contextsv_b200.h"

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "contextsv_b200 kernels are written for sm_100a only"
#endif

namespace csv {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define CSV_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            csv::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CSV_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)
#define CSV_TRY(call)                 \
    do {                              \
        int s__ = (call);             \
        if (s__ != CSV_OK) return s__; \
    } while (0)

// ------------------------------------------------------ geometry constants
constexpr int kSMs = 148;                       // B200
constexpr int kTile = 8192;                     // depth positions per tile (32 KB of int32 in smem)
constexpr int kTileShift = 13;
constexpr int kWalkThreads = 128;                           // 128 x 8 ops: measured best of 32 / 64 / 128 / 256 threads per span
constexpr int kWalkOpsPerThread = 8;
constexpr int kWalkSpan = kWalkThreads * kWalkOpsPerThread;   // CIGAR ops per span
constexpr int kSpanChunk = 2048;                              // spans per chunk of the two-level span scan (256 threads x 8)

// BAM constants (SAM spec): which ops consume reference / query
constexpr uint32_t kRefMask = (1u << 0) | (1u << 2) | (1u << 3) | (1u << 7) | (1u << 8);   // M D N = X
constexpr uint32_t kQryMask = (1u << 0) | (1u << 1) | (1u << 4) | (1u << 7) | (1u << 8);   // M I S = X
constexpr uint32_t kGapMask = (1u << 2) | (1u << 3);                                        // D N
constexpr uint32_t kSigMask = (1u << 1) | (1u << 2) | (1u << 4);                            // I D S
constexpr uint32_t kDepthSkipFlags = 0x4 | 0x100 | 0x200 | 0x400;          // cnv_caller.cpp:491-495
constexpr uint32_t kSigSkipFlags = 0x4 | 0x100 | 0x200 | 0x400 | 0x800;    // sv_caller.cpp:526

// ------------------------------------------------------------ device buffer
// Freed batch buffers are parked in a per-context pool so that the next upload of
// a similar batch does not pay cudaMalloc/cudaFree (both synchronise the device).
struct DevPool {
    std::vector<std::pair<void*, size_t>> free_list;
    void* take(size_t bytes) {
        int best = -1;
        for (int i = 0; i < (int)free_list.size(); i++)
            if (free_list[i].second >= bytes && free_list[i].second <= 2 * bytes + (1u << 20) &&
                (best < 0 || free_list[i].second < free_list[best].second)) best = i;
        if (best < 0) return nullptr;
        void* p = free_list[best].first;
        last_cap = free_list[best].second;
        free_list.erase(free_list.begin() + best);
        return p;
    }
    size_t last_cap = 0;
    void give(void* p, size_t cap) { free_list.emplace_back(p, cap); }
    void trim() { for (auto& e : free_list) cudaFree(e.first); free_list.clear(); }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, DevPool* pool = nullptr) {
        if (bytes <= cap) return CSV_OK;
        release(pool);
        if (pool) { p = pool->take(bytes); if (p) { cap = pool->last_cap; return CSV_OK; } }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && pool) { cudaGetLastError(); pool->trim(); e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) { p = nullptr; csv::set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e)); return CSV_ERR_CUDA; }
        cap = want;
        return CSV_OK;
    }
    void release(DevPool* pool = nullptr) {
        if (p) { if (pool) pool->give(p, cap); else cudaFree(p); }
        p = nullptr; cap = 0;
    }
    template <class T> T* as() const { return (T*)p; }
};

// a typed window into somebody else's DevBuf
struct DevView {
    void* p = nullptr;
    template <class T> T* as() const { return (T*)p; }
};

// Per-stage device timing (CUDA events on the context's stream), for bench.py's roofline.
enum Stage { ST_PREP = 0, ST_WALK, ST_TILE_RANGES, ST_DEPTH_TILES, ST_SIG_SORT, ST_DBSCAN, ST_TILE_KERNEL /* k_depth_tiles16 alone, inside ST_DEPTH_TILES */, ST_COUNT };

// Region tables (device copies live in the batch)
struct RegionDev {       // sorted by (tid, beg)
    uint32_t beg, end;   // end already clipped to map_size
    uint32_t tile_base;  // first global tile of the region
    uint32_t orig;       // caller's region index
};
struct TidDev { uint32_t first, count, map_size, pad; };

// Narrow depth fetch (fetch.cu): the map crosses PCIe as bytes (+ a short list of the values that do not fit) and
// host threads widen it into the caller's uint32 array.
struct FetchState {
    int threads = 0;                 // host threads that widen; 0 = plain 32-bit DMA
    uint32_t chunk = 2u << 20;       // positions per pipeline chunk (multiple of 512)
    uint32_t exc_cap = 2048;         // exception slots per chunk (values >= 255)
    uint32_t min_len = 1u << 18;     // shorter fetches use the plain DMA
    int slots = 0;                   // staging ring (allocated lazily)
    size_t slot_bytes = 0;
    uint8_t* h = nullptr;            // pinned
    uint8_t* d = nullptr;            // device
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> ev_narrow, ev_copy;
    uint64_t narrow_chunks = 0, fallback_chunks = 0;     // statistics since context creation
};
struct FetchSeg { const uint32_t* src; uint32_t* dst; size_t len; };

// Look-back scan state of the walk: record heads and depth events are plain sums,
// (ref, qry) consumption is segmented: it restarts at every record head.
struct WalkAgg { uint32_t heads, ref, qry, ev; };

}  // namespace csv

// ------------------------------------------------------------------ context
struct csv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // the stream kernels are launched on right now (main, or side inside a SideScope)
    cudaStream_t main_stream = nullptr; // depth pipeline, copies, timers
    cudaStream_t side_stream = nullptr; // signature sort + DBSCAN1D: only depend on the walk, run beside the depth tiles
    cudaStream_t tile_stream = nullptr; // tile ranges + depth tiles of chunk c, beside the walk of chunks c + 2, c + 3, ...
    cudaStream_t upload_stream = nullptr; // CIGAR words of a batch that is scanned in pipeline chunks: chunk by chunk, an event each
    std::vector<cudaEvent_t> ev_upload; // chunk c's CIGAR words have arrived (upload stream)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_tile_join = nullptr;
    std::vector<cudaEvent_t> ev_chunk;  // walk of chunk c finished (main stream)
    bool side_busy = false, tile_busy = false;
    int side_grid = 0;                  // CTAs per SM the side-stream kernels may take beside the tiles (0 = each kernel's own default)
    int side_ctas = 0;                  // ... and in total per launch (0 = no cap): every CTA of a side kernel waits for a slot a tile CTA frees
    int pipe_chunks = 1;                // batches uploaded from now on are scanned in up to this many pipelined chunks of contigs
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    uint32_t epoch = 1;                 // look-back epoch, bumped per chained launch
    csv::DevBuf tickets;                // zeroed u32 counters, one per chained launch
    uint32_t ticket_next = 0, ticket_cap = 0;
    csv::DevBuf scan_status;            // u64 status words for chained scans / radix passes
    csv::DevBuf sort_tmp[6];            // radix sort ping-pong buffers + histograms
    csv::DevBuf db[16];                 // DBSCAN scratch
    csv::DevBuf db2[14];                // 2-D DBSCAN scratch
    void* pinned_small = nullptr;       // 4 KB pinned staging for tiny D2H reads
    void* pinned_db = nullptr;          // mapped pinned buffer of the one-launch DBSCAN1D path (points | labels | cluster count)
    void* pinned_db_dev = nullptr;      // ... as the device sees it
    bool db_small = true;               // csv_dbscan1d takes the one-launch path for <= kDbSmallMax points (CSV_DB_SMALL=0 turns it off)
    int sm_count = csv::kSMs;
    csv::DevPool pool;                  // parked batch buffers
    int profile = 0;                    // 0 off, 1 = every stage (diagnostics), 2 = the dominant kernel only (bench.py's timed region)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> stage_events[csv::ST_COUNT];
    std::vector<cudaEvent_t> spare_events;
    double stage_ms[csv::ST_COUNT] = {0};
    uint32_t stage_calls[csv::ST_COUNT] = {0};
    csv::FetchState fetch;
};

namespace csv {
// Returns a device pointer to a zeroed u32 ticket counter (stream-ordered).
int next_ticket(csv_ctx* ctx, uint32_t** out);
inline uint32_t next_epoch(csv_ctx* ctx) { ctx->epoch++; if (ctx->epoch >= 0x3fffffffu) ctx->epoch = 1; return ctx->epoch; }
int ensure_status(csv_ctx* ctx, size_t words);   // u64 words; never needs clearing (epoch-tagged)
// Fork/join between the main and the side stream.
int side_fork(csv_ctx* ctx);            // side waits for everything enqueued on main so far
int side_join(csv_ctx* ctx);            // main waits for everything enqueued on side so far
struct SideScope {                      // kernels launched inside the scope go to the side stream
    csv_ctx* ctx;
    explicit SideScope(csv_ctx* c) : ctx(c) { ctx->stream = ctx->side_stream; ctx->side_busy = true; }
    ~SideScope() { ctx->stream = ctx->main_stream; }
};
struct TileScope {                      // ... to the tile stream
    csv_ctx* ctx;
    explicit TileScope(csv_ctx* c) : ctx(c) { ctx->stream = ctx->tile_stream; ctx->tile_busy = true; }
    ~TileScope() { ctx->stream = ctx->main_stream; }
};
// Depth slices device -> host (fetch.cu).  Blocks until every segment is in host memory.
int fetch_depth_segments(csv_ctx* ctx, const std::vector<FetchSeg>& segs);
void fetch_release(csv_ctx* ctx);
// Grid multiplier (CTAs per SM) of a grid-stride / ticket kernel.  On the side stream, beside the HBM-bound tile kernel,
// every CTA slot a tiny signature kernel holds is taken from the tiles: the context may cap the multiplier there.
inline uint32_t grid_mult(const csv_ctx* ctx, uint32_t dflt)
{
    if (ctx->stream == ctx->side_stream && ctx->side_grid > 0 && (uint32_t)ctx->side_grid < dflt) return (uint32_t)ctx->side_grid;
    return dflt;
}
// ... and the total a launch may take there (all side-stream kernels are grid-stride or ticket loops: any grid >= 1 is correct)
inline uint32_t cap_grid(const csv_ctx* ctx, uint64_t grid)
{
    if (ctx->stream == ctx->side_stream && ctx->side_ctas > 0 && grid > (uint64_t)ctx->side_ctas) return (uint32_t)ctx->side_ctas;
    return (uint32_t)grid;
}
// RAII stage timer: records an event pair around a pipeline stage when profiling is on.
struct StageTimer {
    csv_ctx* ctx; int stage; cudaEvent_t e1 = nullptr;
    StageTimer(csv_ctx* c, int s);
    ~StageTimer();
};
}

// ------------------------------------------------------------- device utils
#ifdef __CUDACC__
namespace csv {

__device__ __forceinline__ uint4 ld_nc_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_cs_v4(void* p, uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane_id() >= (uint32_t)d) v += t; }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// CTA-wide exclusive sum of one u32 per thread (blockDim multiple of 32, <= 1024).
// smem: at least 33 u32.  Returns exclusive prefix; *total receives the CTA sum.
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* smem, uint32_t* total)
{
    uint32_t incl = warp_incl_scan_u32(v);
    uint32_t w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane_id() == 31) smem[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane_id() < nw ? smem[lane_id()] : 0;
        uint32_t xi = warp_incl_scan_u32(x);
        smem[lane_id()] = xi - x;
        if (lane_id() == 31) smem[32] = xi;
    }
    __syncthreads();
    uint32_t base = smem[w];
    *total = smem[32];
    return base + incl - v;
}

}  // namespace csv
#endif

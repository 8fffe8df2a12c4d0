// radix_sort.cu -- CUB-free stable LSD radix sort of 64/128-bit keys + u32 payload.
//
// Onesweep layout: ONE histogram kernel counts all 8/16 byte digits up front; a
// digit on which every key agrees is skipped entirely (decided on the device, no
// host round trip).  Each remaining digit is ONE kernel: tiles are claimed by
// ticket, ranked stably with warp match_any, and the global offset of
// every (tile, digit value) comes from a chained look-back over the previous
// tiles' published counts (256 independent chains, one per thread).
// The element count may live on the device (n_dev), so sorts whose size is only
// known to an earlier kernel need no synchronisation either.
#include "batch.cuh"
#include "scan.cuh"

namespace csv {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;   // 2048 keys per tile
constexpr int kSortWarps = kSortThreads / 32;

struct SortState {              // device
    uint32_t hist[16][256];     // digit histograms, then exclusive bases
    uint32_t trivial[16];       // 1 = all keys share this digit
    uint32_t blocks_done;       // histogram CTAs that have added their share: the last one turns the counts into bases
};

__device__ __forceinline__ uint32_t key_digit(unsigned long long hi, unsigned long long lo, int d)
{
    return d < 8 ? (uint32_t)(lo >> (8 * d)) & 255u : (uint32_t)(hi >> (8 * (d - 8))) & 255u;
}

// Histogram of every digit + (last CTA to finish) exclusive bases and triviality per digit: one launch.  Optionally the
// first kernel of a pipeline whose element count is still raw: n = min(*n_raw, clamp_cap) is published to *n_clamped_out
// (block 0) for the kernels that follow, and val[i] = i is written along the way (payload = original position).
__global__ void __launch_bounds__(256) k_sort_hist(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ lo,
                                                   const uint32_t* n_dev, uint64_t n_host, uint32_t digit_mask, SortState* st,
                                                   const uint32_t* n_raw, uint32_t clamp_cap, uint32_t* n_clamped_out, uint32_t* iota_val)
{
    __shared__ uint32_t s_h[16][256];
    __shared__ uint32_t s_scan[40];
    __shared__ bool s_last;
    uint64_t n = n_dev ? (uint64_t)*n_dev : n_host;
    if (n_raw) {
        const uint32_t r = *n_raw;
        n = r < clamp_cap ? r : clamp_cap;
        if (blockIdx.x == 0 && threadIdx.x == 0) *n_clamped_out = (uint32_t)n;
    }
    for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long l = lo[i], h = hi ? hi[i] : 0ull;
        if (iota_val) iota_val[i] = (uint32_t)i;
#pragma unroll
        for (int d = 0; d < 16; d++)
            if ((digit_mask >> d) & 1u) atomicAdd(&s_h[d][key_digit(h, l, d)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) {
        uint32_t v = (&s_h[0][0])[i];
        if (v) atomicAdd(&(&st->hist[0][0])[i], v);
    }
    // last CTA done: per digit, exclusive scan of the 256 bins + triviality
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int d = 0; d < 16; d++) {
        if (!((digit_mask >> d) & 1u)) { if (threadIdx.x == 0) st->trivial[d] = 1; continue; }
        const uint32_t c = ld_volatile_u32(&st->hist[d][threadIdx.x]);
        uint32_t tot;
        const uint32_t ex = block_excl_scan_u32(c, s_scan, &tot);
        const int triv = __syncthreads_or(c == (uint32_t)n && n > 0);
        st->hist[d][threadIdx.x] = ex;
        if (threadIdx.x == 0) st->trivial[d] = (triv || n < 2) ? 1u : 0u;
    }
}

struct PassParams {
    unsigned long long *hi[2], *lo[2];
    uint32_t* val[2];
    const uint32_t* n_dev;
    uint64_t n_host;
    SortState* st;
    int digit;
    uint32_t* ticket;
    unsigned long long* status;   // [tile][256]
    uint32_t epoch;
};

template <bool HAS_HI>
__global__ void __launch_bounds__(kSortThreads) k_sort_pass(const PassParams P)
{
    __shared__ uint32_t s_cnt[kSortWarps][256];
    __shared__ uint32_t s_tile;
    if (P.st->trivial[P.digit]) return;
    int cur = 0;
    for (int d = 0; d < P.digit; d++) cur ^= (P.st->trivial[d] ? 0 : 1);
    const unsigned long long* __restrict__ src_hi = P.hi[cur];
    const unsigned long long* __restrict__ src_lo = P.lo[cur];
    const uint32_t* __restrict__ src_val = P.val[cur];
    unsigned long long* dst_hi = P.hi[cur ^ 1];
    unsigned long long* dst_lo = P.lo[cur ^ 1];
    uint32_t* dst_val = P.val[cur ^ 1];
    const uint64_t n = P.n_dev ? (uint64_t)*P.n_dev : P.n_host;
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
        for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= n_tiles) break;
        const uint64_t base = (uint64_t)t * kSortTile + (uint64_t)warp * (32 * kSortItems);
        unsigned long long klo[kSortItems], khi[kSortItems];
        uint32_t kval[kSortItems], rank[kSortItems];
#pragma unroll
        for (int i = 0; i < kSortItems; i++) {
            const uint64_t idx = base + i * 32 + lane;
            const bool valid = idx < n;
            klo[i] = valid ? src_lo[idx] : 0ull;
            khi[i] = (HAS_HI && valid) ? src_hi[idx] : 0ull;
            kval[i] = valid ? src_val[idx] : 0u;
            const uint32_t dg = key_digit(khi[i], klo[i], P.digit);
            const uint32_t active = __ballot_sync(0xffffffffu, valid);
            uint32_t peers = 0, pre = 0;
            if (valid) { peers = __match_any_sync(active, dg); pre = s_cnt[warp][dg]; }
            __syncwarp();
            if (valid && lane == (uint32_t)(__ffs(peers) - 1)) s_cnt[warp][dg] = pre + __popc(peers);
            __syncwarp();
            rank[i] = pre + __popc(peers & lanemask_lt());
        }
        __syncthreads();
        // thread `tid` owns digit value `tid`: exclusive over warps, chained look-back over tiles
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) { uint32_t c = s_cnt[w][tid]; s_cnt[w][tid] = run; run += c; }
        unsigned long long* my = P.status + (size_t)t * 256 + tid;
        uint32_t excl = 0;
        if (t == 0) st_volatile_u64(my, lb_pack(P.epoch, kFlagPrefix, run));
        else {
            st_volatile_u64(my, lb_pack(P.epoch, kFlagAgg, run));
            // The chain is walked kLook predecessors at a time: their words are loaded together (independent loads, one
            // round trip to L2) and consumed in order; a word that is not published yet is polled on its own.  Walking
            // one word per round trip made every pass cost (tiles in flight) x (L2 latency).
            constexpr int kLook = 8;
            bool done = false;
            for (int64_t p = (int64_t)t - 1; p >= 0 && !done; p -= kLook) {
                unsigned long long wv[kLook];
#pragma unroll
                for (int u = 0; u < kLook; u++) wv[u] = p - u >= 0 ? ld_volatile_u64(P.status + (size_t)(p - u) * 256 + tid) : 0ull;
#pragma unroll
                for (int u = 0; u < kLook; u++) {
                    if (done || p - u < 0) continue;
                    uint32_t h = (uint32_t)(wv[u] >> 32);
                    uint32_t flag = ((h >> 2) == P.epoch) ? (h & 3u) : 0u;
                    while (flag == 0) {
                        wv[u] = ld_volatile_u64(P.status + (size_t)(p - u) * 256 + tid);
                        h = (uint32_t)(wv[u] >> 32); flag = ((h >> 2) == P.epoch) ? (h & 3u) : 0u;
                    }
                    excl += (uint32_t)wv[u];
                    if (flag == kFlagPrefix) done = true;
                }
            }
            st_volatile_u64(my, lb_pack(P.epoch, kFlagPrefix, excl + run));
        }
        const uint32_t gbase = P.st->hist[P.digit][tid] + excl;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) s_cnt[w][tid] += gbase;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kSortItems; i++) {
            const uint64_t idx = base + i * 32 + lane;
            if (idx < n) {
                const uint32_t dg = key_digit(khi[i], klo[i], P.digit);
                const uint32_t pos = s_cnt[warp][dg] + rank[i];
                dst_lo[pos] = klo[i];
                if (HAS_HI) dst_hi[pos] = khi[i];
                dst_val[pos] = kval[i];
            }
        }
    }
}

// After the last pass: move the result back into the primary buffers if it sits in the alternates.
__global__ void k_sort_normalize(const PassParams P, int n_digits)
{
    int cur = 0;
    for (int d = 0; d < n_digits; d++) cur ^= (P.st->trivial[d] ? 0 : 1);
    if (!cur) return;
    const uint64_t n = P.n_dev ? (uint64_t)*P.n_dev : P.n_host;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        P.lo[0][i] = P.lo[1][i];
        if (P.hi[0]) P.hi[0][i] = P.hi[1][i];
        P.val[0][i] = P.val[1][i];
    }
}

int radix_sort_pairs(csv_ctx* ctx, SortBufs bufs, uint64_t n_upper, const uint32_t* n_dev, uint32_t digit_mask, const SortFirst* first)
{
    if (n_upper >= (1ull << 30)) { set_error("radix sort: %llu keys exceed the 2^30 limit", (unsigned long long)n_upper); return CSV_ERR_LIMIT; }
    const bool has_hi = bufs.hi != nullptr;
    const int n_digits = has_hi ? 16 : 8;
    if (!has_hi) digit_mask &= 0xffu;
    CSV_TRY(ctx->sort_tmp[0].ensure(sizeof(SortState)));
    SortState* st = ctx->sort_tmp[0].as<SortState>();
    CSV_CUDA(cudaMemsetAsync(st, 0, sizeof(SortState), ctx->stream));
    uint64_t tiles = (n_upper + kSortTile - 1) / kSortTile;
    if (tiles == 0) tiles = 1;
    CSV_TRY(ensure_status(ctx, tiles * 256));
    uint32_t grid_h = (uint32_t)((n_upper + 255) / 256);
    if (grid_h > (uint32_t)ctx->sm_count * grid_mult(ctx, 8)) grid_h = ctx->sm_count * grid_mult(ctx, 8);
    if (grid_h == 0) grid_h = 1;
    grid_h = cap_grid(ctx, grid_h);
    k_sort_hist<<<grid_h, 256, 0, ctx->stream>>>(bufs.hi, bufs.lo, n_dev, n_upper, digit_mask, st, first ? first->n_raw : nullptr, first ? first->clamp_cap : 0u,
                                                 first ? first->n_clamped_out : nullptr, first && first->iota ? bufs.val : nullptr);
    ctx->launches++;
    PassParams P;
    P.hi[0] = bufs.hi; P.hi[1] = bufs.hi2; P.lo[0] = bufs.lo; P.lo[1] = bufs.lo2; P.val[0] = bufs.val; P.val[1] = bufs.val2;
    P.n_dev = n_dev; P.n_host = n_upper; P.st = st;
    P.status = ctx->scan_status.as<unsigned long long>();
    const uint64_t pass_cap = (uint64_t)ctx->sm_count * grid_mult(ctx, 6);
    uint32_t grid = cap_grid(ctx, tiles < pass_cap ? tiles : pass_cap);
    for (int d = 0; d < n_digits; d++) {
        if (!((digit_mask >> d) & 1u)) continue;
        P.digit = d;
        CSV_TRY(next_ticket(ctx, &P.ticket));
        P.epoch = next_epoch(ctx);
        if (has_hi) k_sort_pass<true><<<grid, kSortThreads, 0, ctx->stream>>>(P);
        else k_sort_pass<false><<<grid, kSortThreads, 0, ctx->stream>>>(P);
        ctx->launches++;
    }
    k_sort_normalize<<<grid_h, 256, 0, ctx->stream>>>(P, n_digits);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

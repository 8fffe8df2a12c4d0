// dbscan1d.cu -- DBSCAN1D::fit (dbscan1d.cpp:8-66) with bit-identical labels,
// O(N log N), any number of independent fits ("segments") per launch sequence.
//
// The reference is a sequential O(N^2) expansion whose labels depend on input
// order.  They have a closed form (re-verified against the
// compiled reference by tests/test_oracle.py):
//   core(x)   <=> #{j : double(|p_j - x|) <= eps} >= minPts
//   clusters   =  maximal runs of CORE points, in value order, with gaps <= eps
//   id(C)      =  rank of C by the smallest input index among its core points
//                 (that point p_C is where the reference starts expanding C)
//   border b   :  candidates = clusters of the nearest core left / right within eps;
//                 none -> -2; else max(min candidate id,
//                                      max{id(C) : C candidate, |b - p_C| <= eps})
//                 (the first-discovered cluster claims b; a later one steals it only
//                  through the unconditional overwrite at dbscan1d.cpp:32-34)
// Integer distances make `double(d) <= eps` equal to `d <= floor(eps)` for eps >= 0.
//
// Pipeline: radix sort (segment, value) -> core flags by two binary searches ->
// compaction of core points (chained scan) -> run labelling (chained scan) with
// atomicMin of input indices -> radix sort of runs by (segment, min index) ->
// cluster ids -> labels.
#include "batch.cuh"
#include "dbscan_small.h"
#include "scan.cuh"

namespace csv {

// One small fit in one launch: a thread per point, the phases of dbscan_small.h, a block barrier between them.
__global__ void __launch_bounds__(kDbSmallMax) k_db_small(const int32_t* pts, uint32_t n, long long E, int min_pts, int32_t* labels, int32_t* n_clusters)
{
    __shared__ DbSmall S;
    const uint32_t t = threadIdx.x;
#pragma unroll 1
    for (int ph = 0; ph < kDbSmallPhases; ph++) {
        if (t < n) db_small_phase(S, ph, t, n, E, min_pts, pts, labels, n_clusters);
        __syncthreads();
    }
}

// Host side of the small path: points and labels travel through the context's mapped pinned buffer (no copy engine, no
// device allocation): one launch and one stream synchronisation per fit.
int dbscan1d_small(csv_ctx* ctx, const int32_t* pts, uint32_t n, double eps, int min_pts, int32_t* labels_out, int32_t* n_clusters_out)
{
    if (!ctx->pinned_db) {
        CSV_CUDA(cudaHostAlloc(&ctx->pinned_db, 2 * kDbSmallMax * sizeof(int32_t) + 64, cudaHostAllocMapped));
        CSV_CUDA(cudaHostGetDevicePointer(&ctx->pinned_db_dev, ctx->pinned_db, 0));
    }
    int32_t* h_in = (int32_t*)ctx->pinned_db;
    int32_t* h_lab = h_in + kDbSmallMax;
    int32_t* h_nc = h_lab + kDbSmallMax;
    int32_t* d_in = (int32_t*)ctx->pinned_db_dev;
    memcpy(h_in, pts, (size_t)n * sizeof(int32_t));
    const long long E = eps >= 4294967296.0 ? 4294967296ll : (long long)eps;      // floor for eps >= 0, as in the general path
    k_db_small<<<1, (n + 31u) & ~31u, 0, ctx->stream>>>(d_in, n, E, min_pts, d_in + kDbSmallMax, d_in + 2 * kDbSmallMax);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(labels_out, h_lab, (size_t)n * sizeof(int32_t));
    if (n_clusters_out) *n_clusters_out = *h_nc;
    return CSV_OK;
}

struct DbParams {
    const int32_t* pts;
    const uint32_t* seg;
    const uint32_t* n_dev;
    uint64_t n_host;
    long long E;            // floor(eps) clamped to [0, 2^32]
    int min_pts;
    unsigned long long* keys;      // sorted (seg << 32 | biased value)
    uint32_t* idx;                 // sorted payload: input index
    uint32_t* cc;                  // exclusive count of core points before sorted position i
    uint8_t* core;
    uint32_t* clist;               // compact list of core sorted positions
    uint32_t* rid;                 // run id of compact core j
    uint32_t* run_min;             // min input index per run
    unsigned long long* rkeys;     // sorted (seg << 32 | min idx)
    uint32_t* rval;                // sorted payload: run id
    uint32_t* cid;                 // cluster id per run
    uint32_t* counters;            // [0] n_core, [1] n_runs
    int32_t* labels;
    int32_t* n_clusters;
};

__device__ __forceinline__ uint64_t db_n(const DbParams& P) { return P.n_dev ? (uint64_t)*P.n_dev : P.n_host; }
__device__ __forceinline__ unsigned long long db_key(uint32_t seg, long long v)
{
    return ((unsigned long long)seg << 32) | (uint32_t)((uint32_t)(int32_t)v ^ 0x80000000u);
}
__device__ __forceinline__ long long key_value(unsigned long long k) { return (long long)(int32_t)((uint32_t)k ^ 0x80000000u); }

__global__ void k_db_keys(const DbParams P)
{
    const uint64_t n = db_n(P);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        P.keys[i] = db_key(P.seg ? P.seg[i] : 0u, P.pts[i]);
        P.idx[i] = (uint32_t)i;
        P.run_min[i] = 0xffffffffu;
    }
}

__device__ __forceinline__ uint64_t lower_bound_u64(const unsigned long long* a, uint64_t n, unsigned long long x)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint64_t upper_bound_u64(const unsigned long long* a, uint64_t n, unsigned long long x)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void k_db_core(const DbParams P)
{
    const uint64_t n = db_n(P);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = P.keys[i];
        const uint32_t seg = (uint32_t)(k >> 32);
        const long long v = key_value(k);
        long long lo_v = v - P.E, hi_v = v + P.E;
        if (lo_v < -2147483648ll) lo_v = -2147483648ll;
        if (hi_v > 2147483647ll) hi_v = 2147483647ll;
        const uint64_t lb = lower_bound_u64(P.keys, n, db_key(seg, lo_v));
        const uint64_t ub = upper_bound_u64(P.keys, n, db_key(seg, hi_v));
        P.core[i] = ((long long)(ub - lb) >= (long long)P.min_pts) ? 1 : 0;
    }
}

__global__ void k_db_runs_keys(const DbParams P)
{
    const uint32_t n_runs = P.counters[1];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
        const uint32_t mi = P.run_min[r];
        P.rkeys[r] = ((unsigned long long)(P.seg ? P.seg[mi] : 0u) << 32) | mi;
        P.rval[r] = r;
    }
}

__global__ void k_db_ids(const DbParams P)
{
    const uint32_t n_runs = P.counters[1];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += gridDim.x * blockDim.x) {
        const unsigned long long k = P.rkeys[j];
        const uint32_t seg = (uint32_t)(k >> 32);
        const uint32_t first = (uint32_t)lower_bound_u64(P.rkeys, n_runs, (unsigned long long)seg << 32);
        P.cid[P.rval[j]] = j - first;
        if (P.n_clusters && (j + 1 == n_runs || (uint32_t)(P.rkeys[j + 1] >> 32) != seg)) P.n_clusters[seg] = (int32_t)(j - first + 1);
    }
}

__global__ void k_db_labels(const DbParams P)
{
    const uint64_t n = db_n(P);
    const uint32_t n_core = P.counters[0];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = P.keys[i];
        const uint32_t j = P.cc[i];
        int32_t label;
        if (P.core[i]) label = (int32_t)P.cid[P.rid[j]];
        else {
            const uint32_t seg = (uint32_t)(k >> 32);
            const long long v = key_value(k);
            int32_t mn = 0x7fffffff, steal = -1;
            bool any = false;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                // side 0: nearest core to the left (compact j-1); side 1: to the right (compact j)
                if (side == 0 ? (j == 0) : (j >= n_core)) continue;
                const uint32_t cj = side == 0 ? j - 1 : j;
                const unsigned long long ck = P.keys[P.clist[cj]];
                if ((uint32_t)(ck >> 32) != seg) continue;
                long long d = key_value(ck) - v; if (d < 0) d = -d;
                if (d > P.E) continue;
                const uint32_t run = P.rid[cj];
                const int32_t id = (int32_t)P.cid[run];
                any = true;
                if (id < mn) mn = id;
                long long dp = (long long)P.pts[P.run_min[run]] - v; if (dp < 0) dp = -dp;
                if (dp <= P.E && id > steal) steal = id;
            }
            label = any ? (steal > mn ? steal : mn) : -2;
        }
        P.labels[P.idx[i]] = label;
    }
}

// eps < 0 (or NaN): no neighbourhood contains anything, not even the point itself.
// minPts > 0: every point is noise (-2).  minPts <= 0: expandCluster "succeeds" on an
// empty seed set, labels stay -1 and every point consumes one cluster id.
__global__ void k_db_degenerate(const DbParams P)
{
    const uint64_t n = db_n(P);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        P.labels[i] = P.min_pts > 0 ? -2 : -1;
        if (P.n_clusters && P.min_pts <= 0) atomicAdd(&P.n_clusters[P.seg ? P.seg[i] : 0u], 1);
    }
}

// ---- the batch's signature list: sorted by (region, start), segment = region * 2 + (DEL ? 0 : 1) ------------------------
// The general pipeline sorts (segment, value) keys and the runs; here both orders follow from the list's own: a DEL's place
// among the DELs of its region is the number of DELs before it (one scan), runs come out in (segment, smallest input
// index) order, and the first run of every segment is noted by the run scan itself.  Five launches instead of twelve on the
// stream that runs beside the depth tiles: scan, scatter, scan (core flags + compaction), scan (runs), labels.
struct DbSigParams {
    const int32_t* pts;            // out_start, final order
    const uint32_t* seg;           // out_seg, final order
    const uint32_t* n_dev;
    uint32_t* delb;                // [n + 1] DELs before index i
    uint32_t* seg_first;           // [n_seg] first run of the segment
};

__device__ __forceinline__ uint32_t first_of_region(const uint32_t* __restrict__ seg, uint32_t n, uint32_t region)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((seg[mid] >> 1) < region) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) k_dbsig_scatter(const DbParams P, const DbSigParams S)
{
    const uint32_t n = *S.n_dev;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t sg = S.seg[i], r = sg >> 1;
        const uint32_t i0 = first_of_region(S.seg, n, r), i1 = first_of_region(S.seg, n, r + 1u);
        const uint32_t d0 = S.delb[i0], n_del = S.delb[i1] - d0, mine = S.delb[i] - d0;
        const uint32_t pos = (sg & 1u) ? i0 + n_del + (i - i0 - mine) : i0 + mine;
        P.keys[pos] = db_key(sg, S.pts[i]);
        P.idx[pos] = i;
        P.run_min[i] = 0xffffffffu;
    }
}

__device__ __forceinline__ uint32_t dbsig_core(const DbParams& P, uint32_t n, uint32_t i)
{
    const unsigned long long k = P.keys[i];
    const uint32_t seg = (uint32_t)(k >> 32);
    const long long v = key_value(k);
    long long lo_v = v - P.E, hi_v = v + P.E;
    if (lo_v < -2147483648ll) lo_v = -2147483648ll;
    if (hi_v > 2147483647ll) hi_v = 2147483647ll;
    // the window lies a few entries either side of i: gallop out from i before bisecting
    const unsigned long long k_lo = db_key(seg, lo_v), k_hi = db_key(seg, hi_v);
    uint32_t a = i, step = 1;                                   // keys[a] >= k_lo so far; find an index below the window
    uint32_t lo = 0;
    for (;;) { if (a < step) { lo = 0; break; } const uint32_t pr = a - step; if (P.keys[pr] < k_lo) { lo = pr + 1; break; } a = pr; step <<= 1; }
    uint32_t hi = a;                                            // first index with key >= k_lo is in [lo, a]
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (P.keys[mid] < k_lo) lo = mid + 1; else hi = mid; }
    const uint32_t lb = lo;
    uint32_t b = i; step = 1;                                   // keys[b] <= k_hi so far; find an index above the window
    uint32_t top = n;
    for (;;) { const uint32_t pr = b + step; if (pr >= n) { top = n; break; } if (P.keys[pr] > k_hi) { top = pr; break; } b = pr; step <<= 1; }
    lo = b + 1; hi = top;                                       // first index with key > k_hi is in [b + 1, top]
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (P.keys[mid] <= k_hi) lo = mid + 1; else hi = mid; }
    return ((long long)(lo - lb) >= (long long)P.min_pts) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_dbsig_labels(const DbParams P, const DbSigParams S)
{
    const uint32_t n = *S.n_dev;
    const uint32_t n_core = P.counters[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = P.keys[i];
        const uint32_t seg = (uint32_t)(k >> 32);
        const uint32_t j = P.cc[i];
        int32_t label;
        if (P.core[i]) label = (int32_t)(P.rid[j] - S.seg_first[seg]);
        else {
            const long long v = key_value(k);
            int32_t mn = 0x7fffffff, steal = -1;
            bool any = false;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                if (side == 0 ? (j == 0) : (j >= n_core)) continue;
                const uint32_t cj = side == 0 ? j - 1 : j;
                const unsigned long long ck = P.keys[P.clist[cj]];
                if ((uint32_t)(ck >> 32) != seg) continue;
                long long d = key_value(ck) - v; if (d < 0) d = -d;
                if (d > P.E) continue;
                const uint32_t run = P.rid[cj];
                const int32_t id = (int32_t)(run - S.seg_first[seg]);
                any = true;
                if (id < mn) mn = id;
                long long dp = (long long)S.pts[P.run_min[run]] - v; if (dp < 0) dp = -dp;
                if (dp <= P.E && id > steal) steal = id;
            }
            label = any ? (steal > mn ? steal : mn) : -2;
        }
        P.labels[P.idx[i]] = label;
    }
}

int dbscan1d_sorted_sigs(csv_ctx* ctx, const int32_t* d_pts, const uint32_t* d_seg, uint64_t n_upper, const uint32_t* n_dev,
                         uint32_t n_seg, double eps, int min_pts, int32_t* d_labels)
{
    if (n_upper >= (1ull << 30)) { set_error("dbscan1d: %llu points exceed the 2^30 limit", (unsigned long long)n_upper); return CSV_ERR_LIMIT; }
    if (n_upper == 0) return CSV_OK;
    const size_t n = (size_t)n_upper;
    DevBuf* s = ctx->db;
    CSV_TRY(s[0].ensure(n * 8)); CSV_TRY(s[2].ensure(n * 4)); CSV_TRY(s[4].ensure(n * 4)); CSV_TRY(s[5].ensure(n)); CSV_TRY(s[6].ensure(n * 4));
    CSV_TRY(s[7].ensure(n * 4)); CSV_TRY(s[8].ensure(n * 4)); CSV_TRY(s[3].ensure((n + 1) * 4)); CSV_TRY(s[13].ensure((size_t)n_seg * 4 + 16)); CSV_TRY(s[14].ensure(64));
    DbParams P;
    P.pts = d_pts; P.seg = d_seg; P.n_dev = n_dev; P.n_host = n_upper; P.min_pts = min_pts;
    P.keys = s[0].as<unsigned long long>(); P.idx = s[2].as<uint32_t>();
    P.cc = s[4].as<uint32_t>(); P.core = s[5].as<uint8_t>(); P.clist = s[6].as<uint32_t>(); P.rid = s[7].as<uint32_t>();
    P.run_min = s[8].as<uint32_t>(); P.rkeys = nullptr; P.rval = nullptr; P.cid = nullptr; P.counters = s[14].as<uint32_t>();
    P.labels = d_labels; P.n_clusters = nullptr;
    P.E = eps >= 4294967296.0 ? 4294967296ll : (long long)eps;   // floor for eps >= 0
    DbSigParams S;
    S.pts = d_pts; S.seg = d_seg; S.n_dev = n_dev; S.delb = s[3].as<uint32_t>(); S.seg_first = s[13].as<uint32_t>();
    const uint32_t grid = cap_grid(ctx, ctx->sm_count * grid_mult(ctx, 8));
    CSV_CUDA(cudaMemsetAsync(P.counters, 0, 64, ctx->stream));
    {   // DELs before every index (and the total at index n)
        const uint32_t* seg = d_seg; uint32_t* delb = S.delb; const uint32_t* nd = n_dev;
        CSV_TRY(chained_scan(ctx,
                             [=] __device__(uint64_t i) -> uint32_t { return (seg[i] & 1u) ^ 1u; },
                             [=] __device__(uint64_t i, uint32_t ex, uint32_t v) { delb[i] = ex; if (i + 1 == (uint64_t)*nd) delb[i + 1] = ex + v; },
                             n_upper, n_dev, nullptr));
    }
    k_dbsig_scatter<<<grid, 256, 0, ctx->stream>>>(P, S);
    ctx->launches++;
    {   // core flags (window counts in the (segment, value) order) + compaction of the core points
        const DbParams Q = P; const uint32_t* nd = n_dev;
        CSV_TRY(chained_scan(ctx,
                             [=] __device__(uint64_t i) -> uint32_t { return dbsig_core(Q, *nd, (uint32_t)i); },
                             [=] __device__(uint64_t i, uint32_t ex, uint32_t v) { Q.cc[i] = ex; Q.core[i] = (uint8_t)v; if (v) Q.clist[ex] = (uint32_t)i; },
                             n_upper, n_dev, P.counters + 0));
    }
    {   // run ids over the compact core list, smallest input index per run, first run of every segment
        const unsigned long long* keys = P.keys; const uint32_t* clist = P.clist; uint32_t* rid = P.rid;
        uint32_t* run_min = P.run_min; const uint32_t* idx = P.idx; const long long E = P.E; uint32_t* seg_first = S.seg_first;
        CSV_TRY(chained_scan(ctx,
                             [=] __device__(uint64_t j) -> uint32_t {
                                 if (j == 0) return 1u;
                                 const unsigned long long a = keys[clist[j - 1]], c = keys[clist[j]];
                                 return ((uint32_t)(a >> 32) != (uint32_t)(c >> 32) || key_value(c) - key_value(a) > E) ? 1u : 0u;
                             },
                             [=] __device__(uint64_t j, uint32_t ex, uint32_t v) {
                                 const uint32_t r = ex + v - 1u;
                                 rid[j] = r;
                                 atomicMin(&run_min[r], idx[clist[j]]);
                                 const uint32_t sg = (uint32_t)(keys[clist[j]] >> 32);
                                 if (j == 0 || (uint32_t)(keys[clist[j - 1]] >> 32) != sg) seg_first[sg] = r;
                             },
                             n_upper, P.counters + 0, P.counters + 1));
    }
    k_dbsig_labels<<<grid, 256, 0, ctx->stream>>>(P, S);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

int dbscan1d_device(csv_ctx* ctx, const int32_t* d_pts, const uint32_t* d_seg, uint64_t n_upper, const uint32_t* n_dev,
                    uint32_t n_seg, double eps, int min_pts, int32_t* d_labels, int32_t* d_n_clusters, bool value_sorted)
{
    if (n_upper >= (1ull << 30)) { set_error("dbscan1d: %llu points exceed the 2^30 limit", (unsigned long long)n_upper); return CSV_ERR_LIMIT; }
    if (n_upper == 0) return CSV_OK;
    const size_t n = (size_t)n_upper;
    DevBuf* s = ctx->db;
    CSV_TRY(s[0].ensure(n * 8)); CSV_TRY(s[1].ensure(n * 8)); CSV_TRY(s[2].ensure(n * 4)); CSV_TRY(s[3].ensure(n * 4));
    CSV_TRY(s[4].ensure(n * 4)); CSV_TRY(s[5].ensure(n)); CSV_TRY(s[6].ensure(n * 4)); CSV_TRY(s[7].ensure(n * 4));
    CSV_TRY(s[8].ensure(n * 4)); CSV_TRY(s[9].ensure(n * 8)); CSV_TRY(s[10].ensure(n * 8)); CSV_TRY(s[11].ensure(n * 4));
    CSV_TRY(s[12].ensure(n * 4)); CSV_TRY(s[13].ensure(n * 4)); CSV_TRY(s[14].ensure(64));
    DbParams P;
    P.pts = d_pts; P.seg = d_seg; P.n_dev = n_dev; P.n_host = n_upper; P.min_pts = min_pts;
    P.keys = s[0].as<unsigned long long>(); P.idx = s[2].as<uint32_t>();
    P.cc = s[4].as<uint32_t>(); P.core = s[5].as<uint8_t>(); P.clist = s[6].as<uint32_t>(); P.rid = s[7].as<uint32_t>();
    P.run_min = s[8].as<uint32_t>(); P.rkeys = s[9].as<unsigned long long>(); P.rval = s[11].as<uint32_t>();
    P.cid = s[13].as<uint32_t>(); P.counters = s[14].as<uint32_t>();
    P.labels = d_labels; P.n_clusters = d_n_clusters;
    const uint32_t grid = cap_grid(ctx, ctx->sm_count * grid_mult(ctx, 8));
    if (d_n_clusters) CSV_CUDA(cudaMemsetAsync(d_n_clusters, 0, sizeof(int32_t) * n_seg, ctx->stream));
    if (!(eps >= 0.0)) {
        P.E = -1;
        k_db_degenerate<<<grid, 256, 0, ctx->stream>>>(P);
        ctx->launches++;
        CSV_CUDA(cudaGetLastError());
        return CSV_OK;
    }
    P.E = eps >= 4294967296.0 ? 4294967296ll : (long long)eps;   // floor for eps >= 0
    CSV_CUDA(cudaMemsetAsync(P.counters, 0, 64, ctx->stream));
    k_db_keys<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    SortBufs sb;
    sb.hi = nullptr; sb.hi2 = nullptr; sb.lo = P.keys; sb.lo2 = s[1].as<unsigned long long>(); sb.val = P.idx; sb.val2 = s[3].as<uint32_t>();
    // value_sorted: inside every segment the points already come in ascending order (the batch's signature list):
    // a stable sort on the segment bytes alone is then the sort by (segment, value)
    uint32_t mask = value_sorted ? 0u : 0x0fu;
    for (int d = 0; d < 4; d++) if (d == 0 ? n_seg > 1 : (n_seg >> (8 * d))) mask |= 1u << (4 + d);
    CSV_TRY(radix_sort_pairs(ctx, sb, n_upper, n_dev, mask));
    k_db_core<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    {   // compaction of core points
        const uint8_t* core = P.core; uint32_t* cc = P.cc; uint32_t* clist = P.clist;
        CSV_TRY(chained_scan(ctx,
                             [=] __device__(uint64_t i) -> uint32_t { return core[i]; },
                             [=] __device__(uint64_t i, uint32_t ex, uint32_t v) { cc[i] = ex; if (v) clist[ex] = (uint32_t)i; },
                             n_upper, n_dev, P.counters + 0));
    }
    {   // run ids over the compact core list + min input index per run
        const unsigned long long* keys = P.keys; const uint32_t* clist = P.clist; uint32_t* rid = P.rid;
        uint32_t* run_min = P.run_min; const uint32_t* idx = P.idx; const long long E = P.E;
        CSV_TRY(chained_scan(ctx,
                             [=] __device__(uint64_t j) -> uint32_t {
                                 if (j == 0) return 1u;
                                 const unsigned long long a = keys[clist[j - 1]], c = keys[clist[j]];
                                 return ((uint32_t)(a >> 32) != (uint32_t)(c >> 32) || key_value(c) - key_value(a) > E) ? 1u : 0u;
                             },
                             [=] __device__(uint64_t j, uint32_t ex, uint32_t v) {
                                 const uint32_t r = ex + v - 1u;
                                 rid[j] = r;
                                 atomicMin(&run_min[r], idx[clist[j]]);
                             },
                             n_upper, P.counters + 0, P.counters + 1));
    }
    k_db_runs_keys<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    SortBufs rb;
    rb.hi = nullptr; rb.hi2 = nullptr; rb.lo = P.rkeys; rb.lo2 = s[10].as<unsigned long long>(); rb.val = P.rval; rb.val2 = s[12].as<uint32_t>();
    uint32_t rmask = 0;
    for (int d = 0; d < 4; d++) if (d == 0 || (n_upper >> (8 * d))) rmask |= 1u << d;
    for (int d = 0; d < 4; d++) if (d == 0 ? n_seg > 1 : (n_seg >> (8 * d))) rmask |= 1u << (4 + d);
    // ... and the runs, found in (segment, value) order, are already in (segment, smallest input index) order
    if (!value_sorted) CSV_TRY(radix_sort_pairs(ctx, rb, n_upper, P.counters + 1, rmask));
    k_db_ids<<<grid, 256, 0, ctx->stream>>>(P);
    k_db_labels<<<grid, 256, 0, ctx->stream>>>(P);
    ctx->launches += 2;
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

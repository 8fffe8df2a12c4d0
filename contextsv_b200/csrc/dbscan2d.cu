// dbscan2d.cu -- DBSCAN::fit (include/dbscan.h:11-33, src/dbscan.cpp:9-81): the clustering mergeSVs
// (src/sv_object.cpp:45-269) runs on the CIGAR signatures of every (chromosome, SV type), with
// bit-identical labels.  Points are intervals (start, end); the distance is the minimum reciprocal
// overlap (dbscan.cpp:69-81)
//     ov = max(0, min(e1, e2) - max(s1, s2))          (int)
//     d  = 1.0 - std::min(double(ov) / double(e1 - s1), double(ov) / double(e2 - s2))
// evaluated here with the same IEEE double operations (__ddiv_rn / __dsub_rn are exactly rounded).
//
// Two paths, chosen on the host from (eps, min_pts) and on the device from the data:
//
// * Closed form (SURVEY.md 8a row A7 generalised, 8f-1), for 0 <= eps < 1 and intervals of positive
//   length -- the only case mergeSVs produces.  d <= eps then needs a real overlap and the relation is
//   symmetric, so with the intervals sorted by start every neighbour pair is found once by a forward
//   scan that stops at start_j >= end_i:
//     core(x)    <=> #{j : d(x, j) <= eps} >= minPts
//     clusters    =  connected components of the core-core neighbour graph (lock-free union-find)
//     id(C)       =  rank of C by the smallest input index among its cores (that point p_C is where the
//                    reference starts expanding C)
//     border b    :  candidates = clusters with a core neighbour of b; none -> -2; else
//                    max(min candidate id, max{id(C) : b is a neighbour of p_C})
//                    (the first-discovered cluster claims b; a later one steals it only through the
//                     unconditional overwrite of the initial point's seeds, dbscan.cpp:33-35)
//   O(N log N + pairs that overlap).
//
// * Literal (k_db2_literal): the reference's sequential expansion, statement by statement, run by ONE thread
//   block whose 1024 threads share each regionQuery (ordered compaction by block scans, LIFO seed stack).
//   O(N^2 / 1024); it takes every input the closed form does not -- eps >= 1 (everything is within reach),
//   eps < 0 or NaN, zero / negative lengths (0/0 = NaN makes the relation asymmetric through std::min).
#include "batch.cuh"
#include "scan.cuh"

namespace csv {

__device__ __forceinline__ double db2_distance(uint32_t s1, uint32_t e1, uint32_t s2, uint32_t e2)
{
    const int me = min((int)e1, (int)e2), ms = max((int)s1, (int)s2);
    const int ov = max(0, me - ms);
    const int l1 = (int)(e1 - s1), l2 = (int)(e2 - s2);
    const double a = __ddiv_rn((double)ov, (double)l1), b = __ddiv_rn((double)ov, (double)l2);
    const double m = (b < a) ? b : a;                       // std::min(a, b)
    return __dsub_rn(1.0, m);
}

// --------------------------------------------------------------------- literal path
struct Lit2Params {
    const uint32_t *start, *end;
    uint32_t n;
    double eps;
    int min_pts;
    int32_t* clusters;
    uint32_t* list;      // regionQuery result, ascending
    uint32_t* stack;     // seeds, capacity 2 n
};

// neighbours of point q in ascending index order -> P.list; returns their number (all threads)
__device__ uint32_t lit_region_query(const Lit2Params& P, uint32_t q, uint32_t* s_scan)
{
    const uint32_t sq = P.start[q], eq = P.end[q];
    uint32_t total = 0;
    for (uint32_t base = 0; base < P.n; base += blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        const uint32_t f = (j < P.n && db2_distance(sq, eq, P.start[j], P.end[j]) <= P.eps) ? 1u : 0u;
        uint32_t t;
        const uint32_t ex = block_excl_scan_u32(f, s_scan, &t);
        if (f) P.list[total + ex] = j;
        total += t;
        __syncthreads();
    }
    return total;
}

__global__ void __launch_bounds__(1024) k_db2_literal(const Lit2Params P)
{
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_top;
    for (uint32_t i = threadIdx.x; i < P.n; i += blockDim.x) P.clusters[i] = -1;
    if (threadIdx.x == 0) s_top = 0;
    __syncthreads();
    int cid = 0;
    for (uint32_t i = 0; i < P.n; i++) {
        if (P.clusters[i] != -1) continue;                               // uniform: global memory, read after a barrier
        // expandCluster(i, cid)
        const uint32_t m = lit_region_query(P, i, s_scan);
        if ((int)m < P.min_pts) {
            if (threadIdx.x == 0) P.clusters[i] = -2;
            __syncthreads();
            continue;
        }
        // every seed joins the cluster, whatever it was before; the seeds minus i go on the stack in order
        for (uint32_t base = 0; base < m; base += blockDim.x) {
            const uint32_t k = base + threadIdx.x;
            uint32_t x = 0, f = 0;
            if (k < m) { x = P.list[k]; P.clusters[x] = cid; f = x != i ? 1u : 0u; }
            uint32_t t;
            const uint32_t ex = block_excl_scan_u32(f, s_scan, &t);
            const uint32_t top = s_top;
            if (f) P.stack[top + ex] = x;
            __syncthreads();
            if (threadIdx.x == 0) s_top = top + t;
            __syncthreads();
        }
        while (s_top > 0) {
            __syncthreads();
            const uint32_t cur = P.stack[s_top - 1];
            __syncthreads();
            if (threadIdx.x == 0) s_top--;
            __syncthreads();
            const uint32_t r = lit_region_query(P, cur, s_scan);
            if ((int)r >= P.min_pts) {
                for (uint32_t base = 0; base < r; base += blockDim.x) {
                    const uint32_t k = base + threadIdx.x;
                    uint32_t y = 0, f = 0;
                    if (k < r) {
                        y = P.list[k];
                        const int c = P.clusters[y];
                        if (c == -1 || c == -2) { f = c == -1 ? 1u : 0u; P.clusters[y] = cid; }
                    }
                    uint32_t t;
                    const uint32_t ex = block_excl_scan_u32(f, s_scan, &t);
                    const uint32_t top = s_top;
                    if (f) P.stack[top + ex] = y;
                    __syncthreads();
                    if (threadIdx.x == 0) s_top = top + t;
                    __syncthreads();
                }
            }
        }
        cid++;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ closed-form path
struct Db2Params {
    const uint32_t *start, *end;   // input order
    uint32_t n;
    double eps;
    int min_pts;
    unsigned long long* keys;      // sorted start
    uint32_t* idx;                 // sorted position -> input index
    uint32_t *ss, *se;             // start / end in sorted order
    uint32_t* cnt;                 // neighbours (self included)
    uint32_t* parent;              // union-find over sorted positions (cores only)
    uint32_t* minidx;              // per root: smallest input index among the component's cores
    uint32_t* flag;                // per input index: 1 = initial point of a cluster; then its exclusive scan = cluster ids
    uint32_t* cid_at;              // exclusive scan of flag
    uint32_t* min_cand;            // per sorted position (non-core): smallest candidate id
    int32_t* max_init;             // ... largest id of a cluster whose initial point is a neighbour
    uint32_t* bad;                 // set when an interval has a non-positive length: the literal path takes over
    int32_t* labels;               // input order
};

__global__ void k_db2_keys(const Db2Params P)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
        P.keys[i] = P.start[i];
        P.idx[i] = i;
        if ((int)(P.end[i] - P.start[i]) <= 0 || (int)P.start[i] < 0 || (int)P.end[i] < 0) *P.bad = 1u;
    }
}

__global__ void k_db2_gather(const Db2Params P)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n; p += gridDim.x * blockDim.x) {
        const uint32_t i = P.idx[p];
        P.ss[p] = P.start[i]; P.se[p] = P.end[i];
        P.cnt[p] = 1u;                                   // d(x, x) = 0 <= eps
        P.parent[p] = p;
        P.minidx[p] = 0xffffffffu;
        P.flag[i] = 0u;
        P.min_cand[p] = 0xffffffffu;
        P.max_init[p] = -1;
    }
}

// calls f(p, q) for every neighbour pair p < q (sorted positions); one warp per p, lanes stride over the candidates
template <class F>
__device__ __forceinline__ void db2_for_pairs(const Db2Params& P, F f)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P.n; p += warps) {
        const uint32_t s = P.ss[p], e = P.se[p];
        for (uint32_t q0 = p + 1; q0 < P.n; q0 += 32) {
            const uint32_t q = q0 + lane;
            const bool in = q < P.n && P.ss[q] < e;      // sorted by start: beyond this nothing overlaps [s, e)
            if (in && db2_distance(s, e, P.ss[q], P.se[q]) <= P.eps) f(p, q);
            if (!__all_sync(0xffffffffu, in)) break;
        }
    }
}

__global__ void k_db2_count(const Db2Params P)
{
    if (*P.bad) return;
    db2_for_pairs(P, [&](uint32_t p, uint32_t q) { atomicAdd(&P.cnt[p], 1u); atomicAdd(&P.cnt[q], 1u); });
}

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x)
{
    uint32_t p = ld_volatile_u32(parent + x);
    while (p != x) {
        const uint32_t g = ld_volatile_u32(parent + p);
        if (g != p) atomicCAS(parent + x, p, g);         // path halving
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { const uint32_t t = a; a = b; b = t; }
        if (atomicCAS(parent + b, b, a) == b) return;    // hook the larger root under the smaller
    }
}

__global__ void k_db2_union(const Db2Params P)
{
    if (*P.bad) return;
    db2_for_pairs(P, [&](uint32_t p, uint32_t q) {
        if ((int)P.cnt[p] >= P.min_pts && (int)P.cnt[q] >= P.min_pts) uf_union(P.parent, p, q);
    });
}

__global__ void k_db2_minidx(const Db2Params P)
{
    if (*P.bad) return;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n; p += gridDim.x * blockDim.x)
        if ((int)P.cnt[p] >= P.min_pts) atomicMin(&P.minidx[uf_find(P.parent, p)], P.idx[p]);
}
__global__ void k_db2_flag(const Db2Params P)
{
    if (*P.bad) return;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n; p += gridDim.x * blockDim.x)
        if ((int)P.cnt[p] >= P.min_pts && P.parent[p] == p) P.flag[P.minidx[p]] = 1u;
}

__global__ void k_db2_borders(const Db2Params P)
{
    if (*P.bad) return;
    db2_for_pairs(P, [&](uint32_t p, uint32_t q) {
        const bool cp = (int)P.cnt[p] >= P.min_pts, cq = (int)P.cnt[q] >= P.min_pts;
        if (cp == cq) return;
        const uint32_t c = cp ? p : q, b = cp ? q : p;   // core, border
        const uint32_t mi = P.minidx[uf_find(P.parent, c)];
        const uint32_t id = P.cid_at[mi];
        atomicMin(&P.min_cand[b], id);
        if (P.idx[c] == mi) atomicMax(&P.max_init[b], (int32_t)id);
    });
}

__global__ void k_db2_labels(const Db2Params P)
{
    if (*P.bad) return;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n; p += gridDim.x * blockDim.x) {
        int32_t label;
        if ((int)P.cnt[p] >= P.min_pts) label = (int32_t)P.cid_at[P.minidx[uf_find(P.parent, p)]];
        else if (P.min_cand[p] == 0xffffffffu) label = -2;
        else label = max((int32_t)P.min_cand[p], P.max_init[p]);
        P.labels[P.idx[p]] = label;
    }
}

int dbscan2d_device(csv_ctx* ctx, const uint32_t* d_start, const uint32_t* d_end, uint64_t n64, double eps, int min_pts, int32_t* d_labels)
{
    if (n64 >= (1ull << 30)) { set_error("dbscan2d: %llu intervals exceed the 2^30 limit", (unsigned long long)n64); return CSV_ERR_LIMIT; }
    if (n64 == 0) return CSV_OK;
    const uint32_t n = (uint32_t)n64;
    DevBuf* s = ctx->db2;
    const size_t sz4 = (size_t)n * 4 + 16;
    CSV_TRY(s[0].ensure((size_t)n * 8)); CSV_TRY(s[1].ensure((size_t)n * 8));
    for (int i = 2; i <= 12; i++) CSV_TRY(s[i].ensure(i == 11 ? 2 * sz4 : sz4));
    CSV_TRY(s[13].ensure(64));
    uint32_t* bad = s[13].as<uint32_t>();
    CSV_CUDA(cudaMemsetAsync(bad, 0, 64, ctx->stream));
    const uint32_t grid = ctx->sm_count * 8;
    const bool closed_form = eps >= 0.0 && eps < 1.0;      // false for NaN too
    if (closed_form) {
        Db2Params P;
        P.start = d_start; P.end = d_end; P.n = n; P.eps = eps; P.min_pts = min_pts;
        P.keys = s[0].as<unsigned long long>(); P.idx = s[2].as<uint32_t>();
        P.ss = s[4].as<uint32_t>(); P.se = s[5].as<uint32_t>(); P.cnt = s[6].as<uint32_t>(); P.parent = s[7].as<uint32_t>();
        P.minidx = s[8].as<uint32_t>(); P.flag = s[9].as<uint32_t>(); P.cid_at = s[10].as<uint32_t>();
        P.min_cand = s[12].as<uint32_t>(); P.max_init = (int32_t*)s[11].p; P.bad = bad; P.labels = d_labels;
        k_db2_keys<<<grid, 256, 0, ctx->stream>>>(P);
        ctx->launches++;
        SortBufs sb;
        sb.hi = nullptr; sb.hi2 = nullptr; sb.lo = P.keys; sb.lo2 = s[1].as<unsigned long long>(); sb.val = P.idx; sb.val2 = s[3].as<uint32_t>();
        CSV_TRY(radix_sort_pairs(ctx, sb, n, nullptr, 0x0fu));
        k_db2_gather<<<grid, 256, 0, ctx->stream>>>(P);
        k_db2_count<<<grid, 256, 0, ctx->stream>>>(P);
        k_db2_union<<<grid, 256, 0, ctx->stream>>>(P);
        k_db2_minidx<<<grid, 256, 0, ctx->stream>>>(P);
        k_db2_flag<<<grid, 256, 0, ctx->stream>>>(P);
        ctx->launches += 5;
        {   // cluster ids: rank of every initial point among the initial points, in input-index order
            const uint32_t* flag = P.flag; uint32_t* cid_at = P.cid_at;
            CSV_TRY(chained_scan(ctx, [=] __device__(uint64_t i) -> uint32_t { return flag[i]; },
                                 [=] __device__(uint64_t i, uint32_t ex, uint32_t) { cid_at[i] = ex; }, n, nullptr, nullptr));
        }
        k_db2_borders<<<grid, 256, 0, ctx->stream>>>(P);
        k_db2_labels<<<grid, 256, 0, ctx->stream>>>(P);
        ctx->launches += 2;
    }
    // literal path: always launched, a no-op unless the closed form does not apply (decided on the device for the data)
    Lit2Params L;
    L.start = d_start; L.end = d_end; L.n = closed_form ? 0u : n; L.eps = eps; L.min_pts = min_pts;
    L.clusters = d_labels; L.list = s[4].as<uint32_t>(); L.stack = (uint32_t*)s[11].p;
    if (!closed_form) {
        k_db2_literal<<<1, 1024, 0, ctx->stream>>>(L);
        ctx->launches++;
    } else {
        // data-dependent switch: needs the flag on the host (one 4-byte read; mergeSVs-shaped input never sets it)
        CSV_CUDA(cudaMemcpyAsync(ctx->pinned_small, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CSV_CUDA(cudaStreamSynchronize(ctx->stream));
        if (*(uint32_t*)ctx->pinned_small) {
            L.n = n;
            k_db2_literal<<<1, 1024, 0, ctx->stream>>>(L);
            ctx->launches++;
        }
    }
    CSV_CUDA(cudaGetLastError());
    return CSV_OK;
}

}  // namespace csv

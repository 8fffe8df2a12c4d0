// dbscan_small.h -- DBSCAN1D::fit (src/dbscan1d.cpp:8-66) for one small input (<= kDbSmallMax points) as a sequence of
// phases in which every point is handled independently: on the device one thread per point and a block barrier between
// phases (k_db_small in dbscan1d.cu: ONE launch, points and labels through mapped pinned memory -- the split-read pass
// calls fit() once per cluster of alignments with 2..1000 points, src/sv_caller.cpp:270, and the general pipeline's
// ~20 launches cost the same 0.2 ms whatever the size); on the host the same code with a loop per phase
// (tests/native/dbscan_small_emul.cpp checks it against the oracle without a GPU).
//
// Same closed form as the general pipeline (dbscan1d.cu header): cores by eps-window counts, clusters = runs of cores
// with gaps <= eps in value order, id = rank by smallest input index, border points to the first-discovered candidate
// unless a later candidate's initial point is within eps.  eps >= 0 only (E = floor(eps)); eps < 0 / NaN stays with the
// general path's degenerate kernel.  No phase reads what another point writes in the same phase.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CSV_HD __host__ __device__ __forceinline__
#else
#define CSV_HD inline
#endif

namespace csv {

constexpr int kDbSmallMax = 1024;
constexpr int kDbSmallPhases = 8;

struct DbSmall {                       // working set of one fit (shared memory on the device): 24.5 KB
    int32_t in[kDbSmallMax];           // points in input order
    int32_t val[kDbSmallMax];          // sorted by (value, input index)
    uint32_t run_min[kDbSmallMax];     // per run: smallest input index among its points = where the reference starts the cluster
    uint16_t idx[kDbSmallMax];         // sorted position -> input index
    int16_t prev[kDbSmallMax];         // nearest core strictly left of a sorted position (-1: none)
    int16_t next[kDbSmallMax];         // nearest core strictly right (-1: none)
    uint16_t run[kDbSmallMax];         // run of a core sorted position
    uint16_t cid[kDbSmallMax];         // cluster id of a run
    uint8_t core[kDbSmallMax];
    uint8_t flag[kDbSmallMax];         // core point that starts a run
    uint32_t n_runs;
};

// phase `ph` for point / sorted position / run t (t < n).  pts and labels live in (mapped) global memory.
CSV_HD void db_small_phase(DbSmall& S, int ph, uint32_t t, uint32_t n, long long E, int min_pts, const int32_t* pts, int32_t* labels,
                           int32_t* n_clusters)
{
    switch (ph) {
    case 0:
        S.in[t] = pts[t];
        S.run_min[t] = 0xffffffffu;
        break;
    case 1: {                                               // stable rank sort
        const int32_t v = S.in[t];
        uint32_t r = 0;
        for (uint32_t j = 0; j < n; j++) { const int32_t w = S.in[j]; r += (w < v || (w == v && j < t)) ? 1u : 0u; }
        S.val[r] = v; S.idx[r] = (uint16_t)t;
        break;
    }
    case 2: {                                               // core <=> #{j : |p_j - v| <= E} >= minPts
        const long long v = S.val[t];
        long long lo_v = v - E, hi_v = v + E;
        if (lo_v < -2147483648ll) lo_v = -2147483648ll;
        if (hi_v > 2147483647ll) hi_v = 2147483647ll;
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((long long)S.val[mid] < lo_v) lo = mid + 1; else hi = mid; }
        const uint32_t lb = lo;
        lo = 0; hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((long long)S.val[mid] <= hi_v) lo = mid + 1; else hi = mid; }
        S.core[t] = ((long long)(lo - lb) >= (long long)min_pts) ? 1 : 0;
        break;
    }
    case 3: {                                               // nearest cores; a core starts a run if the core before it is further than E
        int p = -1, q = -1;
        for (int j = (int)t - 1; j >= 0; j--) if (S.core[j]) { p = j; break; }
        for (uint32_t j = t + 1; j < n; j++) if (S.core[j]) { q = (int)j; break; }
        S.prev[t] = (int16_t)p; S.next[t] = (int16_t)q;
        S.flag[t] = (S.core[t] && (p < 0 || (long long)S.val[t] - (long long)S.val[p] > E)) ? 1 : 0;
        break;
    }
    case 4: {                                               // run id = number of run starts up to here
        uint32_t r = 0;
        for (uint32_t j = 0; j <= t; j++) r += S.flag[j];
        S.run[t] = (uint16_t)(r ? r - 1u : 0u);
        if (t == n - 1) { S.n_runs = r; if (n_clusters) *n_clusters = (int32_t)r; }
        break;
    }
    case 5:                                                 // where the reference would start each cluster
        if (S.core[t]) {
#if defined(__CUDA_ARCH__)
            atomicMin(&S.run_min[S.run[t]], (uint32_t)S.idx[t]);
#else
            if ((uint32_t)S.idx[t] < S.run_min[S.run[t]]) S.run_min[S.run[t]] = S.idx[t];
#endif
        }
        break;
    case 6:                                                 // cluster id = rank of the run by that index
        if (t < S.n_runs) {
            const uint32_t m = S.run_min[t];
            uint32_t r = 0;
            for (uint32_t j = 0; j < S.n_runs; j++) r += S.run_min[j] < m ? 1u : 0u;
            S.cid[t] = (uint16_t)r;
        }
        break;
    case 7: {
        int32_t label;
        if (S.core[t]) label = (int32_t)S.cid[S.run[t]];
        else {
            const long long v = S.val[t];
            int32_t mn = 0x7fffffff, steal = -1;
            bool any = false;
            for (int side = 0; side < 2; side++) {
                const int c = side == 0 ? S.prev[t] : S.next[t];
                if (c < 0) continue;
                long long d = (long long)S.val[c] - v; if (d < 0) d = -d;
                if (d > E) continue;
                const uint32_t run = S.run[c];
                const int32_t id = (int32_t)S.cid[run];
                any = true;
                if (id < mn) mn = id;
                long long dp = (long long)S.in[S.run_min[run]] - v; if (dp < 0) dp = -dp;
                if (dp <= E && id > steal) steal = id;
            }
            label = any ? (steal > mn ? steal : mn) : -2;
        }
        labels[S.idx[t]] = label;
        break;
    }
    default: break;
    }
}

}  // namespace csv

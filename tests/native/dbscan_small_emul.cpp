// Test harness (CPU): the phases of the small-input DBSCAN1D kernel (contextsv_b200/csrc/dbscan_small.h) run as plain
// loops, one loop per phase -- the device runs the same code with one thread per point and a barrier per phase.
#include <memory>

#include "dbscan_small.h"

extern "C" int dbs_emul(const int32_t* pts, uint32_t n, double eps, int min_pts, int32_t* labels, int32_t* n_clusters)
{
    if (n == 0 || n > (uint32_t)csv::kDbSmallMax || !(eps >= 0.0)) return -1;
    const long long E = eps >= 4294967296.0 ? 4294967296ll : (long long)eps;
    std::unique_ptr<csv::DbSmall> S(new csv::DbSmall);
    for (int ph = 0; ph < csv::kDbSmallPhases; ph++)
        for (uint32_t t = 0; t < n; t++) csv::db_small_phase(*S, ph, t, n, E, min_pts, pts, labels, n_clusters);
    return 0;
}
// the same phases with the points of every phase visited in reverse order: a phase that depended on another point's
// result of the same phase would give a different answer
extern "C" int dbs_emul_reversed(const int32_t* pts, uint32_t n, double eps, int min_pts, int32_t* labels, int32_t* n_clusters)
{
    if (n == 0 || n > (uint32_t)csv::kDbSmallMax || !(eps >= 0.0)) return -1;
    const long long E = eps >= 4294967296.0 ? 4294967296ll : (long long)eps;
    std::unique_ptr<csv::DbSmall> S(new csv::DbSmall);
    for (int ph = 0; ph < csv::kDbSmallPhases; ph++)
        for (uint32_t t = n; t-- > 0;) csv::db_small_phase(*S, ph, t, n, E, min_pts, pts, labels, n_clusters);
    return 0;
}

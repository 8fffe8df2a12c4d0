// dbscan_gpu.cpp -- drop-in definitions of the reference's 2-D DBSCAN members (include/dbscan.h:11-33,
// src/dbscan.cpp:9-81) on top of the C ABI: the clustering mergeSVs (src/sv_object.cpp:45-269) runs on the
// calls of every SV type.  Compiled against the reference's own header; replaces src/dbscan.cpp in the link.
#include "dbscan.h"

#include <stdexcept>
#include <string>

#include "contextsv_b200.h"
#include "gpu_context.h"

void DBSCAN::fit(const std::vector<SVCall>& sv_calls)
{
    clusters.assign(sv_calls.size(), -1);
    if (sv_calls.empty()) return;
    std::vector<uint32_t> start(sv_calls.size()), end(sv_calls.size());
    for (size_t i = 0; i < sv_calls.size(); i++) { start[i] = sv_calls[i].start; end[i] = sv_calls[i].end; }
    csv_ctx* ctx = csvhost::thread_context();
    csvhost::StatTimer st(csvhost::STAT_DBSCAN2D, start.size());
    if (csv_dbscan2d(ctx, start.data(), end.data(), start.size(), epsilon, minPts, clusters.data()) != CSV_OK)
        throw std::runtime_error(std::string("DBSCAN::fit (GPU): ") + csv_last_error());   // caught by run() like any std::exception
}

const std::vector<int>& DBSCAN::getClusters() const { return clusters; }

// private helpers of the reference class: kept so that the class definition stays link-complete
bool DBSCAN::expandCluster(const std::vector<SVCall>&, size_t, int) { return false; }
std::vector<size_t> DBSCAN::regionQuery(const std::vector<SVCall>&, size_t) const { return {}; }
double DBSCAN::distance(const SVCall& a, const SVCall& b) const
{
    // 1 - minimum reciprocal overlap (same value as dbscan.cpp:69-81; unused once fit() runs on the GPU)
    const int lo = static_cast<int>(a.start) > static_cast<int>(b.start) ? static_cast<int>(a.start) : static_cast<int>(b.start);
    const int hi = static_cast<int>(a.end) < static_cast<int>(b.end) ? static_cast<int>(a.end) : static_cast<int>(b.end);
    const double shared = hi > lo ? static_cast<double>(hi - lo) : 0.0;
    const double fa = shared / static_cast<double>(static_cast<int>(a.end - a.start));
    const double fb = shared / static_cast<double>(static_cast<int>(b.end - b.start));
    return 1.0 - (fb < fa ? fb : fa);
}

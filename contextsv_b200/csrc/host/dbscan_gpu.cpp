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
    if (csv_dbscan2d(ctx, start.data(), end.data(), start.size(), epsilon, minPts, clusters.data()) != CSV_OK)
        throw std::runtime_error(std::string("DBSCAN::fit (GPU): ") + csv_last_error());   // caught by run() like any std::exception
}

const std::vector<int>& DBSCAN::getClusters() const { return clusters; }

// private helpers of the reference class: kept so that the class definition stays link-complete
bool DBSCAN::expandCluster(const std::vector<SVCall>&, size_t, int) { return false; }
std::vector<size_t> DBSCAN::regionQuery(const std::vector<SVCall>&, size_t) const { return {}; }
double DBSCAN::distance(const SVCall& point1, const SVCall& point2) const
{
    const int overlap = std::max(0, std::min(static_cast<int>(point1.end), static_cast<int>(point2.end)) - std::max(static_cast<int>(point1.start), static_cast<int>(point2.start)));
    const int length1 = static_cast<int>(point1.end - point1.start), length2 = static_cast<int>(point2.end - point2.start);
    return 1.0 - std::min(static_cast<double>(overlap) / static_cast<double>(length1), static_cast<double>(overlap) / static_cast<double>(length2));
}

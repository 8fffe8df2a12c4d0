// dbscan1d_gpu.cpp -- drop-in definitions of the reference's DBSCAN1D members
// (include/dbscan1d.h:11-32, src/dbscan1d.cpp:8-90) on top of the C ABI.
// Compiled against the reference's own header; replaces src/dbscan1d.cpp in the link.
#include "dbscan1d.h"

#include <stdexcept>

#include "contextsv_b200.h"
#include "gpu_context.h"

void DBSCAN1D::fit(const std::vector<int>& points)
{
    clusters.assign(points.size(), -1);
    if (points.empty()) return;
    csv_ctx* ctx = csvhost::thread_context();
    csvhost::StatTimer st(csvhost::STAT_DBSCAN1D, points.size());
    if (csv_dbscan1d(ctx, points.data(), points.size(), epsilon, minPts, clusters.data(), nullptr) != CSV_OK)
        throw std::runtime_error(std::string("DBSCAN1D::fit (GPU): ") + csv_last_error());   // caught by run() like any std::exception
}

const std::vector<int>& DBSCAN1D::getClusters() const { return clusters; }

std::vector<int> DBSCAN1D::getLargestCluster(const std::vector<int>& points)
{
    std::vector<int> out(points.size());
    out.resize(csv_largest_cluster(points.data(), clusters.data(), points.size(), out.data()));
    return out;
}

// private helpers of the reference class: kept so that the class definition stays link-complete
bool DBSCAN1D::expandCluster(const std::vector<int>&, size_t, int) { return false; }
std::vector<size_t> DBSCAN1D::regionQuery(const std::vector<int>&, size_t) const { return {}; }
double DBSCAN1D::distance(int point1, int point2) const { return std::abs(point1 - point2); }

#include "gpu_context.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace csvhost {

std::vector<int> device_list()
{
    std::vector<int> v;
    const char* e = std::getenv("CONTEXTSV_GPUS");
    std::string s = e ? e : "0";
    size_t p = 0;
    while (p < s.size()) {
        size_t q = s.find(',', p);
        if (q == std::string::npos) q = s.size();
        if (q > p) v.push_back(std::atoi(s.substr(p, q - p).c_str()));
        p = q + 1;
    }
    if (v.empty()) v.push_back(0);
    return v;
}
namespace {
std::atomic<unsigned> g_next{0};
struct Holder {
    csv_ctx* ctx = nullptr;
    ~Holder() { if (ctx) csv_ctx_destroy(ctx); }
};
}  // namespace

namespace {
struct Stats {
    std::mutex m;
    double seconds[STAT_COUNT] = {0};
    unsigned long long calls[STAT_COUNT] = {0}, items[STAT_COUNT] = {0};
    bool on = std::getenv("CONTEXTSV_B200_STATS") != nullptr;
    ~Stats()
    {
        if (!on) return;
        static const char* names[STAT_COUNT] = {"calculateMeanChromosomeCoverage", "  csv_depth", "findCIGARSVs", "  csv_cigar_scan",
                                                "DBSCAN1D::fit", "DBSCAN::fit", "csv_ctx_create", "findCIGARSVs served by the depth pass", "  BAM decode + packing",
                                                "log2 windows on the device", "getReadDepth on the device", "findSplitSVSignatures"};
        for (int i = 0; i < STAT_COUNT; i++)
            std::fprintf(stderr, "[contextsv_b200] %-34s %8llu calls %10.3f s %12llu items\n", names[i], calls[i], seconds[i], items[i]);
    }
};
Stats g_stats;
}  // namespace

void stat_add(int id, double seconds, unsigned long long items)
{
    if (!g_stats.on) return;
    std::lock_guard<std::mutex> lk(g_stats.m);
    g_stats.seconds[id] += seconds; g_stats.calls[id]++; g_stats.items[id] += items;
}

namespace {
struct WarmUp {
    std::once_flag once;
    std::thread t;
    ~WarmUp() { if (t.joinable()) t.join(); }
};
WarmUp g_warm;
}  // namespace

void warm_up_async()
{
    std::call_once(g_warm.once, [] {
        g_warm.t = std::thread([] {
            StatTimer st(STAT_CTX);
            csv_ctx* c = nullptr;
            if (csv_ctx_create(device_list()[0], &c) == CSV_OK) csv_ctx_destroy(c);    // errors surface at the first real call
        });
    });
}

csv_ctx* thread_context()
{
    thread_local Holder h;
    if (!h.ctx) {
        StatTimer st(STAT_CTX);
        static const std::vector<int> devs = device_list();
        const int dev = devs[g_next++ % devs.size()];
        if (csv_ctx_create(dev, &h.ctx) != CSV_OK)
            throw std::runtime_error(std::string("contextsv_b200: ") + csv_last_error());   // no CPU fallback
    }
    return h.ctx;
}

}  // namespace csvhost

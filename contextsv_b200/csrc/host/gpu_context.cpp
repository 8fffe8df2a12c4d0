#include "gpu_context.h"

#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace csvhost {

namespace {
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
std::atomic<unsigned> g_next{0};
struct Holder {
    csv_ctx* ctx = nullptr;
    ~Holder() { if (ctx) csv_ctx_destroy(ctx); }
};
}  // namespace

csv_ctx* thread_context()
{
    thread_local Holder h;
    if (!h.ctx) {
        static const std::vector<int> devs = device_list();
        const int dev = devs[g_next++ % devs.size()];
        if (csv_ctx_create(dev, &h.ctx) != CSV_OK)
            throw std::runtime_error(std::string("contextsv_b200: ") + csv_last_error());   // no CPU fallback
    }
    return h.ctx;
}

}  // namespace csvhost

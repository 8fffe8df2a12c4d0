#include "scan_results.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <stdexcept>

#include "gpu_context.h"

namespace csvhost {

namespace {

struct Prefetch {
    std::unordered_map<uint32_t, uint32_t> depth;                                       // position -> depth
    struct Win { int sample_size; std::vector<uint64_t> sums; std::vector<uint32_t> counts; };
    std::map<std::pair<uint32_t, uint32_t>, Win> windows;                               // (start, end) -> window sums
};

struct Registry {
    std::mutex m;
    std::unordered_map<const void*, std::shared_ptr<ContigResults>> by_vector;
    std::map<std::pair<std::string, int>, std::shared_ptr<ContigResults>> by_contig;
    std::unordered_map<const void*, Prefetch> prefetch;
};
Registry& registry() { static Registry r; return r; }

struct Devices {
    std::vector<std::unique_ptr<Device>> list;
    Devices()
    {
        registry();                                  // constructed first, so that it outlives this object (clear_results() in the destructor)
        for (int id : device_list()) { list.emplace_back(new Device); list.back()->id = id; }
    }
    ~Devices()
    {
        clear_results();
        for (auto& d : list) if (d->ctx) { csv_ctx_destroy(d->ctx); d->ctx = nullptr; }
    }
};
Devices& devices() { static Devices d; return d; }

void free_batches(ContigResults& r)
{
    for (DepthShard& s : r.shards) {
        if (!s.batch) continue;
        std::lock_guard<std::mutex> lk(s.dev->m);
        csv_batch_free(s.dev->ctx, s.batch);
        s.batch = nullptr;
    }
}

}  // namespace

size_t device_count() { return devices().list.size(); }
Device& device_at(size_t i) { return *devices().list[i % devices().list.size()]; }

csv_ctx* ensure_context(Device& d)
{
    if (!d.ctx) {
        StatTimer st(STAT_CTX);
        if (csv_ctx_create(d.id, &d.ctx) != CSV_OK) throw std::runtime_error(std::string("contextsv_b200: ") + csv_last_error());   // no CPU fallback
    }
    return d.ctx;
}

bool host_depth_requested()
{
    static const bool v = [] { const char* e = std::getenv("CONTEXTSV_HOST_DEPTH"); return e && std::atoi(e) != 0; }();
    return v;
}

void put_results(const void* key, std::shared_ptr<ContigResults> r)
{
    Registry& g = registry();
    std::shared_ptr<ContigResults> old_a, old_b;
    {
        std::lock_guard<std::mutex> lk(g.m);
        auto& a = g.by_vector[key]; old_a = a; a = r;
        auto& b = g.by_contig[{r->bam_path, r->tid}]; old_b = b; b = r;
        g.prefetch.erase(key);
    }
    if (old_a && old_a != r) free_batches(*old_a);
    if (old_b && old_b != r && old_b != old_a) free_batches(*old_b);
}

std::shared_ptr<ContigResults> results_for_vector(const void* key)
{
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    auto it = g.by_vector.find(key);
    return it == g.by_vector.end() ? nullptr : it->second;
}

std::shared_ptr<ContigResults> results_for_contig(const char* bam_path, int tid)
{
    if (!bam_path) return nullptr;
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    auto it = g.by_contig.find({std::string(bam_path), tid});
    return it == g.by_contig.end() ? nullptr : it->second;
}

void clear_results()
{
    Registry& g = registry();
    std::vector<std::shared_ptr<ContigResults>> all;
    {
        std::lock_guard<std::mutex> lk(g.m);
        for (auto& e : g.by_contig) all.push_back(e.second);
        for (auto& e : g.by_vector) all.push_back(e.second);
        g.by_contig.clear(); g.by_vector.clear(); g.prefetch.clear();
    }
    for (auto& r : all) free_batches(*r);
}

bool device_depth_at(const ContigResults& r, const uint32_t* pos, size_t n, uint32_t* out)
{
    StatTimer st(STAT_DEPTH_AT, n);
    std::fill(out, out + n, 0u);
    if (n == 0) return true;
    std::vector<uint32_t> sub_pos, sub_out;
    std::vector<size_t> sub_idx;
    for (const DepthShard& s : r.shards) {
        if (!s.batch) return false;
        const uint32_t* p = pos; uint32_t* o = out; size_t m = n;
        if (r.shards.size() > 1) {                           // only the positions inside this shard's slice
            sub_pos.clear(); sub_idx.clear();
            for (size_t i = 0; i < n; i++) if (pos[i] >= s.beg && pos[i] < s.end) { sub_pos.push_back(pos[i]); sub_idx.push_back(i); }
            if (sub_pos.empty()) continue;
            sub_out.resize(sub_pos.size());
            p = sub_pos.data(); o = sub_out.data(); m = sub_pos.size();
        }
        {
            std::lock_guard<std::mutex> lk(s.dev->m);
            if (csv_depth_at(s.dev->ctx, s.batch, s.region, m, p, o) != CSV_OK) return false;
        }
        if (r.shards.size() > 1) for (size_t j = 0; j < sub_idx.size(); j++) out[sub_idx[j]] = sub_out[j];
    }
    return true;
}

bool device_window_sums(const ContigResults& r, uint32_t n_sv, const uint32_t* start, const uint32_t* end, int sample_size,
                        uint64_t* sum_out, uint32_t* count_out)
{
    const size_t n_win = (size_t)n_sv * (size_t)sample_size;
    StatTimer st(STAT_WINDOWS, n_win);
    std::fill(sum_out, sum_out + n_win, 0ull);
    std::fill(count_out, count_out + n_win, 0u);
    if (n_win == 0) return true;
    // positions at or beyond the map never count (cnv_caller.cpp:93); inside it every position lies in exactly one
    // shard, and csv_window_sums returns each shard's share
    std::vector<uint64_t> part_sum;
    std::vector<uint32_t> part_cnt;
    for (const DepthShard& s : r.shards) {
        if (!s.batch) return false;
        if (r.shards.size() == 1) {
            std::lock_guard<std::mutex> lk(s.dev->m);
            return csv_window_sums(s.dev->ctx, s.batch, s.region, n_sv, start, end, sample_size, sum_out, count_out) == CSV_OK;
        }
        bool any = false;
        for (uint32_t i = 0; i < n_sv && !any; i++) any = start[i] <= end[i] && start[i] < s.end && end[i] >= s.beg;
        if (!any) continue;
        part_sum.resize(n_win); part_cnt.resize(n_win);
        {
            std::lock_guard<std::mutex> lk(s.dev->m);
            if (csv_window_sums(s.dev->ctx, s.batch, s.region, n_sv, start, end, sample_size, part_sum.data(), part_cnt.data()) != CSV_OK) return false;
        }
        for (size_t w = 0; w < n_win; w++) { sum_out[w] += part_sum[w]; count_out[w] += part_cnt[w]; }
    }
    return true;
}

void prefetch_depth_at(const void* key, const std::vector<uint32_t>& positions)
{
    auto r = results_for_vector(key);
    if (!r || positions.empty()) return;
    std::vector<uint32_t> out(positions.size());
    if (!device_depth_at(*r, positions.data(), positions.size(), out.data())) return;
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    auto& d = g.prefetch[key].depth;
    for (size_t i = 0; i < positions.size(); i++) d[positions[i]] = out[i];
}

bool prefetched_depth(const void* key, uint32_t pos, uint32_t* out)
{
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    auto it = g.prefetch.find(key);
    if (it == g.prefetch.end()) return false;
    auto d = it->second.depth.find(pos);
    if (d == it->second.depth.end()) return false;
    *out = d->second;
    return true;
}

void prefetch_windows(const void* key, const std::vector<uint32_t>& start, const std::vector<uint32_t>& end, int sample_size)
{
    auto r = results_for_vector(key);
    if (!r || start.empty() || sample_size <= 0) return;
    const size_t n = start.size();
    std::vector<uint64_t> sums(n * (size_t)sample_size);
    std::vector<uint32_t> counts(n * (size_t)sample_size);
    if (!device_window_sums(*r, (uint32_t)n, start.data(), end.data(), sample_size, sums.data(), counts.data())) return;
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    auto& w = g.prefetch[key].windows;
    for (size_t i = 0; i < n; i++) {
        Prefetch::Win& e = w[{start[i], end[i]}];
        e.sample_size = sample_size;
        e.sums.assign(sums.begin() + i * sample_size, sums.begin() + (i + 1) * sample_size);
        e.counts.assign(counts.begin() + i * sample_size, counts.begin() + (i + 1) * sample_size);
    }
}

bool prefetched_windows(const void* key, uint32_t start, uint32_t end, int sample_size, const uint64_t** sums, const uint32_t** counts)
{
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    auto it = g.prefetch.find(key);
    if (it == g.prefetch.end()) return false;
    auto w = it->second.windows.find({start, end});
    if (w == it->second.windows.end() || w->second.sample_size != sample_size) return false;
    *sums = w->second.sums.data(); *counts = w->second.counts.data();      // stable until drop_prefetch / put_results of this key
    return true;
}

void drop_prefetch(const void* key)
{
    Registry& g = registry();
    std::lock_guard<std::mutex> lk(g.m);
    g.prefetch.erase(key);
}

}  // namespace csvhost

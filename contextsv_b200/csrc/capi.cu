// capi.cu -- the extern "C" layer (include/contextsv_b200.h): contexts, uploads,
// the scan pipeline, result fetches and the one-shot host-to-host wrappers.
#include "batch.cuh"
#include "dbscan_small.h"

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <stdarg.h>
#include <thread>

namespace csv {

static thread_local std::string g_err;

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
}

int next_ticket(csv_ctx* ctx, uint32_t** out)
{
    if (ctx->ticket_next >= ctx->ticket_cap) {
        const uint32_t cap = 16384;
        // the counters are about to be zeroed for re-use: chained launches on the other streams may still be drawing from theirs
        if (ctx->ticket_cap) { CSV_CUDA(cudaStreamSynchronize(ctx->side_stream)); CSV_CUDA(cudaStreamSynchronize(ctx->tile_stream)); CSV_CUDA(cudaStreamSynchronize(ctx->main_stream)); }
        CSV_TRY(ctx->tickets.ensure(cap * sizeof(uint32_t)));
        CSV_CUDA(cudaMemsetAsync(ctx->tickets.p, 0, cap * sizeof(uint32_t), ctx->stream));   // stream-ordered after earlier users
        ctx->ticket_cap = cap; ctx->ticket_next = 0;
    }
    *out = ctx->tickets.as<uint32_t>() + ctx->ticket_next++;
    return CSV_OK;
}

int ensure_status(csv_ctx* ctx, size_t words)
{
    const size_t bytes = words * sizeof(unsigned long long);
    if (bytes <= ctx->scan_status.cap) return CSV_OK;
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    CSV_TRY(ctx->scan_status.ensure(bytes));
    CSV_CUDA(cudaMemsetAsync(ctx->scan_status.p, 0, ctx->scan_status.cap, ctx->stream));    // epoch 0 == never published
    return CSV_OK;
}

static cudaEvent_t get_event(csv_ctx* ctx)
{
    if (!ctx->spare_events.empty()) { cudaEvent_t e = ctx->spare_events.back(); ctx->spare_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
StageTimer::StageTimer(csv_ctx* c, int s) : ctx(c), stage(s)
{
    if (!ctx->profile || (ctx->profile == 2 && s != ST_TILE_KERNEL)) return;
    cudaEvent_t e0 = get_event(ctx); e1 = get_event(ctx);
    cudaEventRecord(e0, ctx->stream);
    ctx->stage_events[stage].emplace_back(e0, e1);
}
StageTimer::~StageTimer() { if (e1) cudaEventRecord(e1, ctx->stream); }

int side_fork(csv_ctx* ctx)
{
    CSV_CUDA(cudaEventRecord(ctx->ev_fork, ctx->main_stream));
    CSV_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    return CSV_OK;
}
int side_join(csv_ctx* ctx)
{
    if (ctx->side_busy) {
        CSV_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
        CSV_CUDA(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_join, 0));
        ctx->side_busy = false;
    }
    if (ctx->tile_busy) {
        CSV_CUDA(cudaEventRecord(ctx->ev_tile_join, ctx->tile_stream));
        CSV_CUDA(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_tile_join, 0));
        ctx->tile_busy = false;
    }
    return CSV_OK;
}

static int read_scalars(csv_ctx* ctx, csv_batch* b, uint32_t* out /* SC_COUNT */)
{
    CSV_TRY(side_join(ctx));
    CSV_CUDA(cudaMemcpyAsync(ctx->pinned_small, b->d_scalars.p, SC_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pinned_small, SC_COUNT * sizeof(uint32_t));
    uint64_t n_sig = 0;
    for (uint32_t i = 0; i < kSigSub; i++) n_sig += out[SC_SIG_SUB0 + i];
    out[SC_N_SIG] = (uint32_t)std::min<uint64_t>(n_sig, 0xffffffffu);      // emitted, whether stored or dropped
    return CSV_OK;
}

static int check_overflow(csv_ctx* ctx, csv_batch* b, uint32_t* sc)
{
    CSV_TRY(read_scalars(ctx, b, sc));
    if (b->have_depth && sc[SC_UNSORTED]) {
        set_error("records are not sorted by (contig, position): the depth path needs coordinate-sorted input, like the indexed BAM the reference requires");
        return CSV_ERR_ARG;
    }
    if (sc[SC_ABSURD]) {
        set_error("a record consumes 2^31 or more reference bases: not a valid alignment (BAM positions are int32)");
        return CSV_ERR_LIMIT;
    }
    if (sc[SC_BAD_GAPS]) {
        set_error("csv_reads::n_gap / ref_len do not match the CIGAR: they must hold the number of D / N ops and the reference bases consumed of every record");
        return CSV_ERR_ARG;
    }
    if (b->have_sigs && (sc[SC_SIG_DROPPED] || sc[SC_N_SIG] > b->sig_cap)) {
        // raw slots are dealt out by kSigSub counters: the one that ran ahead hits the capacity a little before the sum
        // does, so what to reserve is the emitted count plus that slack
        const uint64_t grow = std::max<uint64_t>(sc[SC_N_SIG], b->sig_cap);
        const uint64_t want = grow + grow / 8 + 64 * kSigSub;
        set_error("signatures: %u emitted, batch capacity is %llu: csv_batch_reserve_sigs(%llu) and scan again", sc[SC_N_SIG], (unsigned long long)b->sig_cap, (unsigned long long)want);
        sc[SC_N_SIG] = (uint32_t)std::min<uint64_t>(want, 0xffffffffu);
        b->sig_sub_mask = 0;      // the pass that follows deals dense slots: it fits as soon as the capacity covers the count
        return CSV_ERR_CAPACITY;
    }
    return CSV_OK;
}

// signature-side buffers of a batch, sized by b->sig_cap
static int alloc_sig_buffers(csv_ctx* ctx, csv_batch* b)
{
    const size_t sc = (size_t)b->sig_cap;
    CSV_TRY(b->d_sig_hi.ensure(sc * 8, &ctx->pool)); CSV_TRY(b->d_sig_lo.ensure(sc * 8, &ctx->pool)); CSV_TRY(b->d_sig_k.ensure(sc * 4, &ctx->pool));
    CSV_TRY(b->d_sig_kind.ensure(sc, &ctx->pool)); CSV_TRY(b->d_sig_payload.ensure(sc * 4, &ctx->pool));
    CSV_TRY(b->d_sig_bucket.ensure(sc * 4, &ctx->pool)); CSV_TRY(b->d_sig_arrival.ensure(sc * 4, &ctx->pool));
    CSV_TRY(b->d_out_start.ensure(sc * 4, &ctx->pool)); CSV_TRY(b->d_out_end.ensure(sc * 4, &ctx->pool)); CSV_TRY(b->d_out_kind.ensure(sc, &ctx->pool));
    CSV_TRY(b->d_out_read.ensure(sc * 4, &ctx->pool)); CSV_TRY(b->d_out_op.ensure(sc * 4, &ctx->pool)); CSV_TRY(b->d_out_qpos.ensure(sc * 4, &ctx->pool));
    CSV_TRY(b->d_out_seg.ensure(sc * 4, &ctx->pool));
    return CSV_OK;
}

int wait_upload(csv_ctx* ctx, csv_batch* b, uint32_t chunk_or_all)
{
    if (!b->split_upload) return CSV_OK;
    const uint32_t c = chunk_or_all == 0xffffffffu ? (uint32_t)b->chunks.size() - 1u : chunk_or_all;   // one stream, in order: the last event covers all
    CSV_CUDA(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_upload[c], 0));
    return CSV_OK;
}

}  // namespace csv

using namespace csv;

extern "C" {

const char* csv_last_error(void) { return g_err.c_str(); }
const char* csv_version(void) { return "contextsv_b200 0.1 (sm_100a)"; }

int csv_ctx_create(int device, csv_ctx** out)
{
    if (!out) { set_error("csv_ctx_create: out is NULL"); return CSV_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { set_error("no CUDA device available (%s): there is no CPU fallback", cudaGetErrorString(e)); return CSV_ERR_CUDA; }
    if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return CSV_ERR_ARG; }
    CSV_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CSV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_error("device %d is sm_%d%d; this library only carries sm_100a code", device, prop.major, prop.minor); return CSV_ERR_CUDA; }
    csv_ctx* ctx = new csv_ctx;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (getenv("CSV_CHUNKS")) ctx->pipe_chunks = std::max(1, atoi(getenv("CSV_CHUNKS")));
    if (getenv("CSV_SIDE_GRID")) ctx->side_grid = std::max(0, atoi(getenv("CSV_SIDE_GRID")));
    if (getenv("CSV_SIDE_CTAS")) ctx->side_ctas = std::max(0, atoi(getenv("CSV_SIDE_CTAS")));
    if (getenv("CSV_DB_SMALL")) ctx->db_small = atoi(getenv("CSV_DB_SMALL")) != 0;
    CSV_CUDA(cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking));
    // the signature kernels are many and tiny: with the highest priority their CTAs take the first slot a tile CTA frees
    int prio_lo = 0, prio_hi = 0;
    CSV_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (getenv("CSV_SIDE_PRIO") && atoi(getenv("CSV_SIDE_PRIO")) == 0) prio_hi = prio_lo;     // tuning knob: side stream at normal priority
    CSV_CUDA(cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, prio_hi));
    CSV_CUDA(cudaStreamCreateWithFlags(&ctx->tile_stream, cudaStreamNonBlocking));
    CSV_CUDA(cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->main_stream;
    CSV_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CSV_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CSV_CUDA(cudaEventCreateWithFlags(&ctx->ev_tile_join, cudaEventDisableTiming));
    CSV_CUDA(cudaEventCreate(&ctx->ev0));
    CSV_CUDA(cudaEventCreate(&ctx->ev1));
    CSV_CUDA(cudaHostAlloc(&ctx->pinned_small, 4096, cudaHostAllocDefault));
    // narrow depth fetch: on by default; widening threads = host cores - 2 (the thread that feeds the pipeline and the
    // CUDA driver want the rest), at most 16 (fetch.cu)
    const unsigned hw = std::thread::hardware_concurrency();
    ctx->fetch.threads = (int)std::min(16u, hw > 3 ? hw - 2 : 1u);
    if (getenv("CSV_FETCH_THREADS")) ctx->fetch.threads = std::max(0, atoi(getenv("CSV_FETCH_THREADS")));
    *out = ctx;
    return CSV_OK;
}

void csv_ctx_destroy(csv_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamSynchronize(ctx->tile_stream);
    cudaStreamSynchronize(ctx->main_stream);
    ctx->pool.trim();
    fetch_release(ctx);
    cudaStreamSynchronize(ctx->upload_stream);
    for (auto e : ctx->ev_chunk) cudaEventDestroy(e);
    for (auto e : ctx->ev_upload) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->upload_stream);
    cudaEventDestroy(ctx->ev_tile_join);
    cudaStreamDestroy(ctx->tile_stream);
    for (auto& v : ctx->stage_events) for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto e : ctx->spare_events) cudaEventDestroy(e);
    ctx->tickets.release(); ctx->scan_status.release();
    for (auto& b : ctx->sort_tmp) b.release();
    for (auto& b : ctx->db) b.release();
    for (auto& b : ctx->db2) b.release();
    if (ctx->pinned_small) cudaFreeHost(ctx->pinned_small);
    if (ctx->pinned_db) cudaFreeHost(ctx->pinned_db);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->side_stream);
    cudaStreamDestroy(ctx->main_stream);
    delete ctx;
}

int csv_ctx_sync(csv_ctx* ctx)
{
    if (!ctx) { set_error("null context"); return CSV_ERR_ARG; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    CSV_TRY(side_join(ctx));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    CSV_CUDA(cudaStreamSynchronize(ctx->upload_stream));     // uploads in flight from pinned host arrays
    return CSV_OK;
}

void* csv_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void csv_host_free(void* p) { if (p) cudaFreeHost(p); }

int csv_timer_begin(csv_ctx* ctx)
{
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    CSV_TRY(side_join(ctx));
    CSV_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return CSV_OK;
}
int csv_timer_end(csv_ctx* ctx, float* ms_out)
{
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    CSV_TRY(side_join(ctx));
    CSV_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    CSV_CUDA(cudaEventSynchronize(ctx->ev1));
    CSV_CUDA(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
    return CSV_OK;
}
uint64_t csv_ctx_launch_count(const csv_ctx* ctx) { return ctx ? ctx->launches : 0; }

int csv_ctx_set_pipeline_chunks(csv_ctx* ctx, int n_chunks)
{
    if (!ctx || n_chunks < 1) { set_error("csv_ctx_set_pipeline_chunks: need a context and n_chunks >= 1"); return CSV_ERR_ARG; }
    ctx->pipe_chunks = n_chunks;
    return CSV_OK;
}

int csv_ctx_set_fetch(csv_ctx* ctx, int threads, int64_t chunk_positions, int64_t exception_slots, int64_t min_positions)
{
    if (!ctx) { set_error("null context"); return CSV_ERR_ARG; }
    if (chunk_positions >= 0 && (chunk_positions < 512 || chunk_positions % 512 || chunk_positions > (1 << 28))) {
        set_error("csv_ctx_set_fetch: chunk_positions must be a multiple of 512 in [512, 2^28]"); return CSV_ERR_ARG;
    }
    if (exception_slots > (1 << 24) || min_positions > 0xffffffffll) { set_error("csv_ctx_set_fetch: value out of range"); return CSV_ERR_ARG; }
    if (threads >= 0) ctx->fetch.threads = std::min(threads, 256);
    if (chunk_positions >= 0) ctx->fetch.chunk = (uint32_t)chunk_positions;
    if (exception_slots >= 0) ctx->fetch.exc_cap = (uint32_t)exception_slots;
    if (min_positions >= 0) ctx->fetch.min_len = (uint32_t)min_positions;
    return CSV_OK;
}
int csv_ctx_fetch_stats(const csv_ctx* ctx, uint64_t* narrow_chunks_out, uint64_t* fallback_chunks_out)
{
    if (!ctx) { set_error("null context"); return CSV_ERR_ARG; }
    if (narrow_chunks_out) *narrow_chunks_out = ctx->fetch.narrow_chunks;
    if (fallback_chunks_out) *fallback_chunks_out = ctx->fetch.fallback_chunks;
    return CSV_OK;
}

int csv_profile_enable(csv_ctx* ctx, int on)
{
    if (!ctx) { set_error("null context"); return CSV_ERR_ARG; }
    ctx->profile = on < 0 ? 0 : on;
    return CSV_OK;
}

int csv_profile_read(csv_ctx* ctx, int max_stages, const char** names_out, double* ms_out, uint32_t* calls_out, int reset)
{
    static const char* kNames[ST_COUNT] = {"prep", "walk", "tile_ranges", "depth_tiles", "sig_sort", "dbscan1d", "k_depth_tiles16"};
    if (!ctx) { set_error("null context"); return -CSV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    if (side_join(ctx) != CSV_OK || cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("csv_profile_read: stream synchronisation failed"); return -CSV_ERR_CUDA; }
    for (int s = 0; s < ST_COUNT; s++) {
        for (auto& e : ctx->stage_events[s]) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) { ctx->stage_ms[s] += ms; ctx->stage_calls[s]++; }
            ctx->spare_events.push_back(e.first); ctx->spare_events.push_back(e.second);
        }
        ctx->stage_events[s].clear();
    }
    for (int s = 0; s < ST_COUNT && s < max_stages; s++) {
        if (names_out) names_out[s] = kNames[s];
        if (ms_out) ms_out[s] = ctx->stage_ms[s];
        if (calls_out) calls_out[s] = ctx->stage_calls[s];
    }
    if (reset) for (int s = 0; s < ST_COUNT; s++) { ctx->stage_ms[s] = 0; ctx->stage_calls[s] = 0; }
    return ST_COUNT;
}

/* ------------------------------------------------------------------ batch */

int csv_batch_upload(csv_ctx* ctx, const csv_reads* r, uint32_t n_regions, const csv_region* regions, csv_batch** out)
{
    if (!ctx || !r || !out || !regions || n_regions == 0) { set_error("csv_batch_upload: null argument or no regions"); return CSV_ERR_ARG; }
    *out = nullptr;
    if (r->n_reads && (!r->pos0 || !r->flag || !r->mapq || !r->cig_off)) { set_error("csv_batch_upload: missing SoA array"); return CSV_ERR_ARG; }
    if (r->n_ops && !r->cigar) { set_error("csv_batch_upload: cigar is NULL"); return CSV_ERR_ARG; }
    if (r->n_reads && r->cig_off[r->n_reads] != r->n_ops) { set_error("csv_batch_upload: cig_off[n_reads] != n_ops"); return CSV_ERR_ARG; }
    if (r->n_ops >= (1ull << 31)) { set_error("batch of %llu CIGAR ops exceeds the 2^31 per-batch limit: split it", (unsigned long long)r->n_ops); return CSV_ERR_LIMIT; }
    if (r->n_ops + (uint64_t)r->n_reads >= (1ull << 31)) { set_error("batch of %llu ops + %u records exceeds the 2^31 event-slot limit: split it", (unsigned long long)r->n_ops, r->n_reads); return CSV_ERR_LIMIT; }
    if (n_regions >= (1u << 30)) { set_error("too many regions"); return CSV_ERR_LIMIT; }
    CSV_CUDA(cudaSetDevice(ctx->device));

    // ---- region tables
    int32_t max_tid = -1;
    for (uint32_t i = 0; i < n_regions; i++) {
        const csv_region& g = regions[i];
        if (g.tid < 0 || g.beg >= g.end || g.end > g.map_size) { set_error("region %u: need tid >= 0 and beg < end <= map_size", i); return CSV_ERR_ARG; }
        if (g.map_size > 0x80000000u) { set_error("region %u: map_size beyond 2^31 (BAM positions are int32)", i); return CSV_ERR_LIMIT; }
        max_tid = std::max(max_tid, g.tid);
    }
    if (max_tid >= (1 << 24)) { set_error("contig id %d too large", max_tid); return CSV_ERR_LIMIT; }
    std::vector<uint32_t> order(n_regions);
    for (uint32_t i = 0; i < n_regions; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return regions[a].tid != regions[b].tid ? regions[a].tid < regions[b].tid : regions[a].beg < regions[b].beg;
    });
    // an error return below hands every buffer allocated so far back to the context's pool (out of memory is the likely one)
    struct Releaser { csv_ctx* ctx; void operator()(csv_batch* x) const { if (x) { x->release(&ctx->pool); delete x; } } };
    std::unique_ptr<csv_batch, Releaser> b(new csv_batch, Releaser{ctx});
    b->n_reads = r->n_reads; b->n_ops = r->n_ops; b->n_regions = n_regions; b->n_tids = (uint32_t)max_tid + 1;
    b->has_tid = r->tid != nullptr;
    b->regions.assign(regions, regions + n_regions);
    b->tile_base.resize(n_regions + 1);
    uint64_t tiles = 0;
    for (uint32_t i = 0; i < n_regions; i++) { b->tile_base[i] = (uint32_t)tiles; tiles += ((uint64_t)(regions[i].end - regions[i].beg) + kTile - 1) / kTile; }
    if (tiles >= (1ull << 31)) { set_error("too many depth tiles"); return CSV_ERR_LIMIT; }
    b->tile_base[n_regions] = (uint32_t)tiles; b->n_tiles = (uint32_t)tiles;
    std::vector<RegionDev> regs(n_regions);
    std::vector<TidDev> tids(b->n_tids, TidDev{0, 0, 0, 0});
    for (uint32_t s = 0; s < n_regions; s++) {
        const csv_region& g = regions[order[s]];
        regs[s] = RegionDev{g.beg, g.end, b->tile_base[order[s]], order[s]};
        TidDev& t = tids[g.tid];
        if (t.count == 0) { t.first = s; t.map_size = g.map_size; }
        else {
            if (t.map_size != g.map_size) { set_error("regions of contig %d disagree on map_size", g.tid); return CSV_ERR_ARG; }
            if (regs[s - 1].end > g.beg) { set_error("regions of contig %d overlap", g.tid); return CSV_ERR_ARG; }
            b->multi_region_tid = true;
        }
        t.count++;
    }
    std::vector<uint32_t> reg_tab(3 * n_regions + 1);
    for (uint32_t i = 0; i <= n_regions; i++) reg_tab[i] = b->tile_base[i];
    for (uint32_t i = 0; i < n_regions; i++) { reg_tab[n_regions + 1 + i] = regions[i].end - regions[i].beg; reg_tab[2 * n_regions + 1 + i] = regions[i].beg; }

    b->n_spans = (uint32_t)((r->n_ops + kWalkSpan - 1) / kWalkSpan);
    // ---- pipeline chunks: cuts only between contigs, walk ranges aligned to the chunks of the span scan
    {
        const int want = ctx->pipe_chunks;
        std::vector<uint32_t> utid;
        for (uint32_t s = 0; s < n_regions; s++) { const uint32_t t = (uint32_t)regions[order[s]].tid; if (utid.empty() || utid.back() != t) utid.push_back(t); }
        const uint32_t min_spans = std::max<uint32_t>((uint32_t)kSpanChunk, b->n_spans / (uint32_t)std::max(want, 1));
        PipeChunk cur;
        cur.span0 = 0; cur.first_tid = 0;
        uint64_t rec0 = 0;
        if (r->tid && want > 1) {
            for (size_t i = 1; i < utid.size(); i++) {
                const int32_t* it = std::lower_bound(r->tid, r->tid + r->n_reads, (int32_t)utid[i]);
                const uint64_t ri = (uint64_t)(it - r->tid);
                const uint64_t cut = (r->cig_off[ri] / kWalkSpan / kSpanChunk) * kSpanChunk;
                if (cut >= (uint64_t)cur.span0 + min_spans && cut + min_spans <= b->n_spans) {
                    cur.span1 = (uint32_t)cut; cur.rec_upper = (uint32_t)(ri - rec0);
                    b->chunks.push_back(cur);
                    cur = PipeChunk(); cur.span0 = (uint32_t)cut; cur.first_tid = utid[i]; rec0 = ri;
                }
            }
        }
        cur.span1 = b->n_spans; cur.rec_upper = (uint32_t)(r->n_reads - rec0);
        b->chunks.push_back(cur);
        for (size_t c = 0; c < b->chunks.size(); c++) {
            const uint32_t lo = b->chunks[c].first_tid, hi = c + 1 < b->chunks.size() ? b->chunks[c + 1].first_tid : 0xffffffffu;
            for (uint32_t i = 0; i < n_regions; i++) {
                const uint32_t t = (uint32_t)regions[i].tid;
                if (t < lo || t >= hi || b->tile_base[i] == b->tile_base[i + 1]) continue;
                auto& tl = b->chunks[c].tiles;
                if (!tl.empty() && tl.back().second == b->tile_base[i]) tl.back().second = b->tile_base[i + 1];
                else tl.emplace_back(b->tile_base[i], b->tile_base[i + 1]);
            }
        }
    }
    // record-level pre-pass: the caller counted the D/N ops per record and records are short (a span start lies a few
    // dozen ops into the record that crosses it; ONT-like batches keep the op-level pre-pass)
    static const bool rec_ok = !(getenv("CSV_REC_PREPASS") && atoi(getenv("CSV_REC_PREPASS")) == 0);
    b->rec_prepass = rec_ok && r->n_gap != nullptr && r->n_reads > 0 && r->n_ops <= (uint64_t)r->n_reads * 1024u;
    // ... and with the reference length of every record the tile ranges need nothing from the walk (single-chunk passes):
    // the prefix max runs beside the record scan, the range searches are enqueued ahead of the walk, the walk leaves what it
    // finds in d_ref_chk and k_claim_check compares after it.  Measured on B200 three ways (profiles/r2_history.md) and never a
    // gain: ranges on the high-priority stream beside the walk 3.88 ms per whole-genome step (the block scheduler drains SMs
    // for the larger CTAs), the same with the walk not touching the claim 3.75, ranges enqueued ahead of the walk at normal
    // priority 3.73 (they still end up behind the walk's grid) -- against 3.68 with the ranges after the walk.  Opt-in:
    // CSV_CLAIM_REFLEN=1.
    static const bool claim_ok = getenv("CSV_CLAIM_REFLEN") && atoi(getenv("CSV_CLAIM_REFLEN")) != 0;
    b->claimed_ref = claim_ok && b->rec_prepass && r->ref_len != nullptr && b->chunks.size() == 1;
    // small batches (and the radix order) keep ONE slot counter and dense raw slots: contention is no matter there, and a
    // handful of CTAs would not spread over the counters anyway
    b->sig_sub_mask = (sig_order_radix() || r->n_ops < (1u << 22)) ? 0u : kSigSub - 1u;
    b->ev_cap = 2 * (r->n_ops + (uint64_t)r->n_reads) + 2;      // exact bound: 2 per record + 2 per D/N op
    b->sig_cap = std::max<uint64_t>(16, std::min<uint64_t>(r->n_ops, std::max<uint64_t>(1u << 20, r->n_ops / 16)));

    // ---- allocations
    const size_t nr = r->n_reads, no = (size_t)r->n_ops, nt = b->n_tiles, sc = (size_t)b->sig_cap;
    if (b->has_tid) CSV_TRY(b->d_tid.ensure(nr * 4 + 16, &ctx->pool));
    CSV_TRY(b->d_pos0.ensure(nr * 4 + 16, &ctx->pool)); CSV_TRY(b->d_flag.ensure(nr * 2 + 16, &ctx->pool)); CSV_TRY(b->d_mapq.ensure(nr + 16, &ctx->pool));
    CSV_TRY(b->d_cig_off.ensure((nr + 1) * 8, &ctx->pool)); CSV_TRY(b->d_cigar.ensure(no * 4 + 64, &ctx->pool));
    if (b->rec_prepass) { CSV_TRY(b->d_n_gap.ensure(nr * 4 + 16, &ctx->pool)); CSV_TRY(b->d_ev_check.ensure((nr + 2) * 4, &ctx->pool)); }
    if (b->claimed_ref) { CSV_TRY(b->d_ref_len.ensure(nr * 4 + 16, &ctx->pool)); CSV_TRY(b->d_ref_chk.ensure(nr * 4 + 16, &ctx->pool)); }
    CSV_TRY(b->d_span_rq.ensure(((size_t)b->n_spans + 2) * 8, &ctx->pool));
    CSV_TRY(b->d_meta.ensure(nr * 16 + 16, &ctx->pool)); CSV_TRY(b->d_key.ensure(nr * 8 + 16, &ctx->pool)); CSV_TRY(b->d_ne_idx.ensure(nr * 4 + 16, &ctx->pool)); CSV_TRY(b->d_headbits.ensure(no / 8 + 512, &ctx->pool));   // the walk copies 272 bytes per span, also for the last one
    b->state_bytes = ((size_t)SC_COUNT + b->chunks.size() + 4 + n_regions) * 4;
    CSV_TRY(b->d_scalars.ensure(b->state_bytes, &ctx->pool));
    b->d_tickets.p = b->d_scalars.as<uint32_t>() + SC_COUNT;
    b->d_reg_sig_cnt.p = b->d_scalars.as<uint32_t>() + SC_COUNT + b->chunks.size() + 4;
    CSV_TRY(b->d_regs.ensure(regs.size() * sizeof(RegionDev), &ctx->pool)); CSV_TRY(b->d_tids.ensure(tids.size() * sizeof(TidDev), &ctx->pool));
    CSV_TRY(b->d_reg_tab.ensure(reg_tab.size() * 4, &ctx->pool));
    CSV_TRY(b->d_span_agg.ensure((size_t)b->n_spans * sizeof(WalkAgg) + 16, &ctx->pool)); CSV_TRY(b->d_span_pre.ensure((size_t)b->n_spans * sizeof(WalkAgg) + 16, &ctx->pool));
    CSV_TRY(b->d_span_status.ensure(((size_t)b->n_spans / kSpanChunk + 2) * sizeof(WalkAgg), &ctx->pool));   // per-chunk aggregates
    CSV_TRY(b->d_scan_carry.ensure(sizeof(WalkAgg), &ctx->pool)); CSV_TRY(b->d_span_desc.ensure((size_t)b->n_spans * 16 + 32, &ctx->pool));
    CSV_TRY(b->d_chunk_tid.ensure(b->chunks.size() * 4 + 16, &ctx->pool)); CSV_TRY(b->d_chunk_bounds.ensure(b->chunks.size() * 4 + 16, &ctx->pool));
    CSV_TRY(b->d_events.ensure((size_t)b->ev_cap * 4, &ctx->pool)); CSV_TRY(b->d_depth.ensure(nt * (size_t)kTile * 4, &ctx->pool));
    CSV_TRY(b->d_ev_start.ensure((nr + 2) * 4, &ctx->pool)); CSV_TRY(b->d_ref_end.ensure(nr * 4 + 16, &ctx->pool));
    CSV_TRY(b->d_pmax.ensure(nr * 8 + 16, &ctx->pool)); CSV_TRY(b->d_pmax_part.ensure((nr / 2048 + 2) * 8, &ctx->pool));
    CSV_TRY(b->d_tile_desc.ensure(nt * 16 + 16, &ctx->pool)); CSV_TRY(b->d_tile_ev.ensure(nt * 8 + 16, &ctx->pool));
    CSV_TRY(b->d_wide_list.ensure(nt * 4 + 16, &ctx->pool)); CSV_TRY(b->d_tile_q.ensure(nt * 16 + 16, &ctx->pool)); CSV_TRY(b->d_tile_r.ensure(nt * 8 + 16, &ctx->pool));
    CSV_TRY(b->d_tile_sum.ensure(nt * 8 + 16, &ctx->pool)); CSV_TRY(b->d_tile_nz.ensure(nt * 4 + 16, &ctx->pool));
    CSV_TRY(b->d_sum.ensure(n_regions * 8, &ctx->pool)); CSV_TRY(b->d_nz.ensure(n_regions * 4, &ctx->pool));
    CSV_TRY(b->d_bucket_cnt.ensure(nt * 4 + 16, &ctx->pool)); CSV_TRY(b->d_bucket_base.ensure(nt * 4 + 16, &ctx->pool));
    CSV_TRY(alloc_sig_buffers(ctx, b.get()));

    // ---- uploads (asynchronous when the host buffers are pinned)
    cudaStream_t st = ctx->stream;
    if (nr) {
        if (b->has_tid) CSV_CUDA(cudaMemcpyAsync(b->d_tid.p, r->tid, nr * 4, cudaMemcpyHostToDevice, st));
        CSV_CUDA(cudaMemcpyAsync(b->d_pos0.p, r->pos0, nr * 4, cudaMemcpyHostToDevice, st));
        CSV_CUDA(cudaMemcpyAsync(b->d_flag.p, r->flag, nr * 2, cudaMemcpyHostToDevice, st));
        CSV_CUDA(cudaMemcpyAsync(b->d_mapq.p, r->mapq, nr, cudaMemcpyHostToDevice, st));
        CSV_CUDA(cudaMemcpyAsync(b->d_cig_off.p, r->cig_off, (nr + 1) * 8, cudaMemcpyHostToDevice, st));
        if (b->rec_prepass) CSV_CUDA(cudaMemcpyAsync(b->d_n_gap.p, r->n_gap, nr * 4, cudaMemcpyHostToDevice, st));
        if (b->claimed_ref) CSV_CUDA(cudaMemcpyAsync(b->d_ref_len.p, r->ref_len, nr * 4, cudaMemcpyHostToDevice, st));
    } else {
        CSV_CUDA(cudaMemsetAsync(b->d_cig_off.p, 0, 8, st));
    }
    // The CIGAR words are nine tenths of the upload.  A batch that will be scanned in pipeline chunks gets them chunk by
    // chunk on the upload stream, an event behind each: the walk of chunk c waits for its own words only, and the scan
    // of the first chunks runs while the last ones are still crossing PCIe (csv_scan_run).
    b->split_upload = no != 0 && b->chunks.size() > 1;
    if (b->split_upload) {
        while (ctx->ev_upload.size() < b->chunks.size()) {
            cudaEvent_t e = nullptr;
            CSV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->ev_upload.push_back(e);
        }
        for (size_t c = 0; c < b->chunks.size(); c++) {
            const size_t o0 = (size_t)b->chunks[c].span0 * kWalkSpan, o1 = std::min<size_t>((size_t)b->chunks[c].span1 * kWalkSpan, no);
            if (o1 > o0) CSV_CUDA(cudaMemcpyAsync(b->d_cigar.as<uint32_t>() + o0, r->cigar + o0, (o1 - o0) * 4, cudaMemcpyHostToDevice, ctx->upload_stream));
            CSV_CUDA(cudaEventRecord(ctx->ev_upload[c], ctx->upload_stream));
        }
    } else if (no) CSV_CUDA(cudaMemcpyAsync(b->d_cigar.p, r->cigar, no * 4, cudaMemcpyHostToDevice, st));
    // table copies come from temporaries: make them synchronous with respect to the host
    CSV_CUDA(cudaMemcpyAsync(b->d_regs.p, regs.data(), regs.size() * sizeof(RegionDev), cudaMemcpyHostToDevice, st));
    CSV_CUDA(cudaMemcpyAsync(b->d_tids.p, tids.data(), tids.size() * sizeof(TidDev), cudaMemcpyHostToDevice, st));
    CSV_CUDA(cudaMemcpyAsync(b->d_reg_tab.p, reg_tab.data(), reg_tab.size() * 4, cudaMemcpyHostToDevice, st));
    std::vector<uint32_t> chunk_tid(b->chunks.size());
    for (size_t c = 0; c < b->chunks.size(); c++) chunk_tid[c] = b->chunks[c].first_tid;
    CSV_CUDA(cudaMemcpyAsync(b->d_chunk_tid.p, chunk_tid.data(), chunk_tid.size() * 4, cudaMemcpyHostToDevice, st));
    std::vector<uint4> tile_desc(nt);
    for (uint32_t i = 0; i < n_regions; i++) {
        const uint32_t len = regions[i].end - regions[i].beg;
        for (uint32_t t = b->tile_base[i]; t < b->tile_base[i + 1]; t++) {
            const uint32_t p0 = (t - b->tile_base[i]) * (uint32_t)kTile;
            tile_desc[t] = make_uint4(i, len - p0 < (uint32_t)kTile ? len - p0 : (uint32_t)kTile, regions[i].beg + p0, (uint32_t)regions[i].tid);
        }
    }
    if (nt) CSV_CUDA(cudaMemcpyAsync(b->d_tile_desc.p, tile_desc.data(), nt * sizeof(uint4), cudaMemcpyHostToDevice, st));
    CSV_CUDA(cudaMemsetAsync(b->d_pmax_part.p, 0, (nr / 2048 + 2) * 8, st));      // look-back status words of k_pmax_chained: epoch 0 == never published
    CSV_CUDA(cudaMemsetAsync(b->d_bucket_cnt.p, 0, nt * 4 + 16, st));              // signatures per depth tile: the ordering's scan leaves zeros behind for the next pass
    CSV_CUDA(cudaMemsetAsync(b->d_headbits.p, 0, no / 8 + 512, st));               // record-head bits: a function of cig_off alone, every pass ORs the same bits in (prep.cu)
    // No synchronisation here: the tables above come from pageable temporaries, which cudaMemcpyAsync stages before it
    // returns; the caller's SoA is either pageable (same) or pinned -- then the copies are in flight and the arrays must
    // stay untouched until a call that waits for the stream (csv_ctx_sync, csv_depth_stats, any fetch).
    *out = b.release();
    return CSV_OK;
}

void csv_batch_free(csv_ctx* ctx, csv_batch* b)
{
    if (!b) return;
    if (ctx) { cudaSetDevice(ctx->device); side_join(ctx); cudaStreamSynchronize(ctx->stream); if (b->split_upload) cudaStreamSynchronize(ctx->upload_stream); }
    b->release(ctx ? &ctx->pool : nullptr);
    delete b;
}

int csv_batch_release_inputs(csv_ctx* ctx, csv_batch* b)
{
    if (!ctx || !b) { set_error("csv_batch_release_inputs: null argument"); return CSV_ERR_ARG; }
    if (b->inputs_released) return CSV_OK;
    CSV_CUDA(cudaSetDevice(ctx->device));
    CSV_TRY(side_join(ctx));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));            // the pass may still be reading them
    if (b->split_upload) CSV_CUDA(cudaStreamSynchronize(ctx->upload_stream));
    b->release_inputs(&ctx->pool);
    return CSV_OK;
}

int csv_batch_reserve_sigs(csv_ctx* ctx, csv_batch* b, uint64_t n_sigs)
{
    if (!ctx || !b) { set_error("csv_batch_reserve_sigs: null argument"); return CSV_ERR_ARG; }
    if (b->inputs_released) { set_error("csv_batch_reserve_sigs: the batch's inputs were released"); return CSV_ERR_STATE; }
    if (n_sigs >= (1ull << 30)) { set_error("csv_batch_reserve_sigs: %llu signatures exceed the 2^30 limit", (unsigned long long)n_sigs); return CSV_ERR_LIMIT; }
    if (n_sigs <= b->sig_cap) return CSV_OK;
    CSV_CUDA(cudaSetDevice(ctx->device));
    CSV_TRY(side_join(ctx));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));            // the buffers are about to change hands
    b->sig_cap = std::min<uint64_t>(std::max<uint64_t>(n_sigs, 16), std::max<uint64_t>(b->n_ops, 16));
    b->d_labels.release(&ctx->pool);
    b->have_sigs = false; b->have_labels = false;            // results of the pass that overflowed are void
    return alloc_sig_buffers(ctx, b);
}

int csv_scan_run(csv_ctx* ctx, csv_batch* b, const csv_scan_params* p)
{
    if (!ctx || !b || !p) { set_error("csv_scan_run: null argument"); return CSV_ERR_ARG; }
    if (!p->want_depth && !p->want_sigs) { set_error("csv_scan_run: nothing requested"); return CSV_ERR_ARG; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    if (b->inputs_released) { set_error("csv_scan_run: the batch's inputs were released (csv_batch_release_inputs): only its results are left"); return CSV_ERR_STATE; }
    CSV_TRY(side_join(ctx));            // the previous pass's signature work may still be reading this batch
    cudaStream_t st = ctx->stream;
    b->last_min_len = p->min_len;
    CSV_CUDA(cudaMemsetAsync(b->d_scalars.p, 0, b->state_bytes, st));    // scalars, tile tickets, signatures per region: one memset
    if (p->want_depth && !b->rec_prepass) CSV_CUDA(cudaMemsetAsync(b->d_ev_start.p, 0, 4, st));   // (the record scan writes slot 0 itself)
    { StageTimer t(ctx, ST_PREP); CSV_TRY(launch_prep(ctx, b, p->min_mapq)); }
    // The pass is pipelined over chunks of whole contigs.  Main stream: the walk, chunk after chunk.  Tile stream:
    // tile ranges + depth tiles of chunk c as soon as the walk of chunk c + 1 is through (the records at the end of
    // chunk c's contigs share their last span-scan chunk with chunk c + 1).  The walk is bound by instruction issue,
    // the tiles by HBM writes: side by side they cost little more than the tiles alone.
    const uint32_t nc = (uint32_t)b->chunks.size();
    while (ctx->ev_chunk.size() < nc) {
        cudaEvent_t e = nullptr;
        CSV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_chunk.push_back(e);
    }
    const bool ranges_first = p->want_depth && b->claimed_ref;           // csv_reads::ref_len: the tile ranges are ready BEFORE the walk ends
    if (p->want_depth) {
        CSV_TRY(side_fork(ctx));
        CSV_CUDA(cudaStreamWaitEvent(ctx->tile_stream, ctx->ev_fork, 0));
        TileScope ts(ctx);
        CSV_TRY(launch_depth_begin(ctx, b));
        // (not timed as a stage: these run beside the walk, off the critical path; ST_TILE_RANGES is what follows the walk)
        CSV_TRY(launch_chunk_bounds(ctx, b));                            // only the tile ranges read them: off the walk's stream
        CSV_TRY(launch_tile_hi(ctx, b));                                 // beside the record scan / the walk
        if (ranges_first) CSV_TRY(launch_tile_ranges(ctx, b, 0, 1));     // prefix max of the claimed record ends: beside the record scan
    }
    { StageTimer t(ctx, ST_WALK); CSV_TRY(launch_record_prepass(ctx, b, p, 1)); }
    if (ranges_first) {
        // the range searches need the event slots of the record scan and nothing else: enqueued BEFORE the span carry and the
        // walk so that their few CTAs are dispatched ahead of the walk's grid (normal priority: a high-priority launch
        // beside the walk made the block scheduler drain SMs for its larger CTAs and cost the walk 0.15 ms)
        CSV_CUDA(cudaEventRecord(ctx->ev_join, ctx->main_stream));
        CSV_CUDA(cudaStreamWaitEvent(ctx->tile_stream, ctx->ev_join, 0));
        TileScope ts(ctx);
        StageTimer t(ctx, ST_TILE_RANGES);
        CSV_TRY(launch_tile_ranges(ctx, b, 0, 2));
    }
    if (nc == 1) { StageTimer t(ctx, ST_WALK); CSV_TRY(launch_record_prepass(ctx, b, p, 2)); }
    if (nc == 1) {
        // One chunk: the whole critical chain stays on the main stream (walk -> prefix max -> range searches -> tiles ->
        // reductions; no cross-stream hop in it).  The signature side stream forks right behind the walk; what the tile
        // stream did beside the walk (clears, chunk bounds, static halves of the ranges) is long finished when it is joined.
        { StageTimer t(ctx, ST_WALK); CSV_TRY(launch_walk(ctx, b, p, b->chunks[0].span0, b->chunks[0].span1)); }
        if (p->want_sigs && p->want_depth) {
            CSV_TRY(side_fork(ctx));
            SideScope side(ctx);
            StageTimer t(ctx, ST_SIG_SORT);
            CSV_TRY(launch_sig_finish(ctx, b));
        } else if (p->want_sigs) {
            StageTimer t(ctx, ST_SIG_SORT);
            CSV_TRY(launch_sig_finish(ctx, b));
        }
        if (p->want_depth) {
            CSV_CUDA(cudaEventRecord(ctx->ev_tile_join, ctx->tile_stream));
            CSV_CUDA(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_tile_join, 0));
            ctx->tile_busy = false;
            if (!ranges_first) { StageTimer t(ctx, ST_TILE_RANGES); CSV_TRY(launch_tile_ranges(ctx, b, 0)); }
            // the pile-up tiles (a launch that finds its list empty, as a rule) run on the tile stream BESIDE the 16-bit
            // kernel: different tiles, nothing shared; the reductions wait for both
            CSV_CUDA(cudaEventRecord(ctx->ev_join, ctx->main_stream));
            CSV_CUDA(cudaStreamWaitEvent(ctx->tile_stream, ctx->ev_join, 0));
            { TileScope ts(ctx); CSV_TRY(launch_depth_finish(ctx, b, 1)); CSV_TRY(launch_ev_check(ctx, b, p)); }   // ... and so does the check of the caller's D/N counts
            CSV_CUDA(cudaEventRecord(ctx->ev_tile_join, ctx->tile_stream));
            ctx->tile_busy = false;
            StageTimer t(ctx, ST_DEPTH_TILES);
            CSV_TRY(launch_depth_tiles(ctx, b, 0));
            CSV_CUDA(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_tile_join, 0));
            CSV_TRY(launch_depth_finish(ctx, b, 2));
        }
    } else {
    for (uint32_t c = 0; c < nc; c++) {
        CSV_TRY(wait_upload(ctx, b, c));                               // this chunk's CIGAR words (no-op unless they came up in pieces)
        {
            StageTimer t(ctx, ST_WALK);
            // the span starts of (span0, span1]: each exactly once over the chunks; the walk of chunk c reads the descriptors of span0 .. span1
            CSV_TRY(launch_record_prepass(ctx, b, p, 2, b->chunks[c].span0, b->chunks[c].span1));
            CSV_TRY(launch_walk(ctx, b, p, b->chunks[c].span0, b->chunks[c].span1));
        }
        CSV_CUDA(cudaEventRecord(ctx->ev_chunk[c], ctx->main_stream));
        if (p->want_depth && c >= 1) {
            const uint32_t tc = c - 1;                                   // its records are complete now
            CSV_CUDA(cudaStreamWaitEvent(ctx->tile_stream, ctx->ev_chunk[c], 0));
            TileScope ts(ctx);
            { StageTimer t(ctx, ST_TILE_RANGES); CSV_TRY(launch_tile_ranges(ctx, b, tc)); }
            { StageTimer t(ctx, ST_DEPTH_TILES); CSV_TRY(launch_depth_tiles(ctx, b, tc)); }
        }
    }
    if (p->want_depth) {
        TileScope ts(ctx);
        { StageTimer t(ctx, ST_TILE_RANGES); CSV_TRY(launch_tile_ranges(ctx, b, nc - 1)); }
        { StageTimer t(ctx, ST_DEPTH_TILES); CSV_TRY(launch_depth_tiles(ctx, b, nc - 1)); }
        StageTimer t(ctx, ST_DEPTH_TILES);
        CSV_TRY(launch_depth_finish(ctx, b));
    }
    if (p->want_sigs && p->want_depth) {   // sort + gather of the signatures run beside the depth kernels
        CSV_TRY(side_fork(ctx));
        SideScope side(ctx);
        StageTimer t(ctx, ST_SIG_SORT);
        CSV_TRY(launch_sig_finish(ctx, b));
    } else if (p->want_sigs) {
        StageTimer t(ctx, ST_SIG_SORT);
        CSV_TRY(launch_sig_finish(ctx, b));
    }
    }
    if (ranges_first) CSV_TRY(launch_claim_check(ctx, b));              // main stream, beside the tiles: only the fetches wait for it
    if (nc > 1 || !p->want_depth) CSV_TRY(launch_ev_check(ctx, b, p));   // (single-chunk depth passes ran it beside the tiles)
    b->scanned = true; b->have_depth = p->want_depth != 0; b->have_sigs = p->want_sigs != 0; b->have_labels = false;
    return CSV_OK;
}

int csv_depth_stats(csv_ctx* ctx, csv_batch* b, uint64_t* sum_out, uint32_t* nonzero_out)
{
    if (!ctx || !b) { set_error("null argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_stats: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    uint32_t sc[SC_COUNT];
    CSV_TRY(check_overflow(ctx, b, sc));
    if (sum_out) CSV_CUDA(cudaMemcpyAsync(sum_out, b->d_sum.p, b->n_regions * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nonzero_out) CSV_CUDA(cudaMemcpyAsync(nonzero_out, b->d_nz.p, b->n_regions * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    return CSV_OK;
}

int csv_depth_fetch(csv_ctx* ctx, csv_batch* b, uint32_t region, uint32_t* depth_out)
{
    if (!ctx || !b || !depth_out) { set_error("null argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_fetch: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    if (region >= b->n_regions) { set_error("region %u out of range", region); return CSV_ERR_ARG; }
    CSV_TRY(side_join(ctx));             // the tiles run on their own stream
    const size_t len = b->regions[region].end - b->regions[region].beg;
    const uint32_t* src = b->d_depth.as<uint32_t>() + (size_t)b->tile_base[region] * kTile;
    return fetch_depth_segments(ctx, {FetchSeg{src, depth_out, len}});
}

int csv_depth_fetch_all(csv_ctx* ctx, csv_batch* b, uint32_t* const* depth_out)
{
    if (!ctx || !b || !depth_out) { set_error("null argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_fetch_all: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    CSV_TRY(side_join(ctx));
    std::vector<FetchSeg> segs;
    for (uint32_t r = 0; r < b->n_regions; r++) {
        if (!depth_out[r]) continue;                           // regions the caller does not want
        segs.push_back(FetchSeg{b->d_depth.as<uint32_t>() + (size_t)b->tile_base[r] * kTile, depth_out[r], (size_t)(b->regions[r].end - b->regions[r].beg)});
    }
    return fetch_depth_segments(ctx, segs);
}

int csv_depth_device_ptr(csv_ctx* ctx, csv_batch* b, uint32_t region, const uint32_t** dptr_out)
{
    if (!ctx || !b || !dptr_out || region >= b->n_regions) { set_error("bad argument"); return CSV_ERR_ARG; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    CSV_TRY(side_join(ctx));             // work the caller enqueues on the context's stream after this call sees the finished map
    *dptr_out = b->d_depth.as<uint32_t>() + (size_t)b->tile_base[region] * kTile;
    return CSV_OK;
}

int csv_sigs_count(csv_ctx* ctx, csv_batch* b, uint64_t* n_out)
{
    if (!ctx || !b || !n_out) { set_error("null argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_sigs) { set_error("csv_sigs_count: run csv_scan_run with want_sigs first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    uint32_t sc[SC_COUNT];
    int s = check_overflow(ctx, b, sc);
    *n_out = sc[SC_N_SIG];
    return s;
}

int csv_sigs_fetch(csv_ctx* ctx, csv_batch* b, csv_sigs* out, uint64_t cap, uint64_t* n_out, uint64_t* region_off_out)
{
    if (!ctx || !b || !out || !n_out) { set_error("null argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_sigs) { set_error("csv_sigs_fetch: run csv_scan_run with want_sigs first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    uint32_t sc[SC_COUNT];
    sc[SC_N_SIG] = 0;
    const int ovf = check_overflow(ctx, b, sc);
    const uint64_t n = sc[SC_N_SIG];
    *n_out = n;                                              // also with CSV_ERR_CAPACITY: the size to reserve
    if (ovf != CSV_OK) return ovf;
    if (n > cap) { set_error("csv_sigs_fetch: %llu signatures, caller capacity %llu", (unsigned long long)n, (unsigned long long)cap); return CSV_ERR_CAPACITY; }
    cudaStream_t st = ctx->stream;
    if (n) {
        if (out->start) CSV_CUDA(cudaMemcpyAsync(out->start, b->d_out_start.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (out->end) CSV_CUDA(cudaMemcpyAsync(out->end, b->d_out_end.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (out->kind) CSV_CUDA(cudaMemcpyAsync(out->kind, b->d_out_kind.p, n, cudaMemcpyDeviceToHost, st));
        if (out->read_idx) CSV_CUDA(cudaMemcpyAsync(out->read_idx, b->d_out_read.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (out->op_idx) CSV_CUDA(cudaMemcpyAsync(out->op_idx, b->d_out_op.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (out->query_pos) CSV_CUDA(cudaMemcpyAsync(out->query_pos, b->d_out_qpos.p, n * 4, cudaMemcpyDeviceToHost, st));
    }
    if (region_off_out) {
        std::vector<uint32_t> cnt(b->n_regions);
        CSV_CUDA(cudaMemcpyAsync(cnt.data(), b->d_reg_sig_cnt.p, b->n_regions * 4, cudaMemcpyDeviceToHost, st));
        CSV_CUDA(cudaStreamSynchronize(st));
        uint64_t acc = 0;
        for (uint32_t i = 0; i < b->n_regions; i++) { region_off_out[i] = acc; acc += cnt[i]; }
        region_off_out[b->n_regions] = acc;
    }
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

int csv_sigs_dbscan1d(csv_ctx* ctx, csv_batch* b, double eps, int min_pts, int32_t* labels_out, uint64_t cap)
{
    if (!ctx || !b) { set_error("null argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_sigs) { set_error("csv_sigs_dbscan1d: run csv_scan_run with want_sigs first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    if (ctx->side_busy) {               // the signature sort is still on the side stream: follow it there
        SideScope side(ctx);
        StageTimer t(ctx, ST_DBSCAN);
        CSV_TRY(launch_sig_dbscan(ctx, b, eps, min_pts));
    } else {
        StageTimer t(ctx, ST_DBSCAN);
        CSV_TRY(launch_sig_dbscan(ctx, b, eps, min_pts));
    }
    b->have_labels = true;
    if (labels_out) {
        CSV_TRY(side_join(ctx));
        uint32_t sc[SC_COUNT];
        CSV_TRY(check_overflow(ctx, b, sc));
        const uint64_t n = sc[SC_N_SIG];
        if (n > cap) { set_error("csv_sigs_dbscan1d: %llu labels, caller capacity %llu", (unsigned long long)n, (unsigned long long)cap); return CSV_ERR_CAPACITY; }
        if (n) CSV_CUDA(cudaMemcpyAsync(labels_out, b->d_labels.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return CSV_OK;
}

/* ------------------------------------------------------- one-shot wrappers */

int csv_depth(csv_ctx* ctx, const csv_reads* reads, const csv_region* region, uint32_t* depth_out, uint64_t* sum_out, uint32_t* nonzero_out)
{
    csv_batch* b = nullptr;
    CSV_TRY(csv_batch_upload(ctx, reads, 1, region, &b));
    csv_scan_params p = {50, 20, 1, 0, 0};
    int s = csv_scan_run(ctx, b, &p);
    if (s == CSV_OK) s = csv_depth_stats(ctx, b, sum_out, nonzero_out);
    if (s == CSV_OK && depth_out) s = csv_depth_fetch(ctx, b, 0, depth_out);
    csv_batch_free(ctx, b);
    return s;
}

int csv_cigar_scan(csv_ctx* ctx, const csv_reads* reads, const csv_region* region, uint32_t min_len, uint8_t min_mapq,
                   csv_sigs* out, uint64_t cap, uint64_t* n_out)
{
    csv_batch* b = nullptr;
    CSV_TRY(csv_batch_upload(ctx, reads, 1, region, &b));
    csv_scan_params p = {min_len, min_mapq, 0, 1, 0};
    int s = csv_scan_run(ctx, b, &p);
    uint64_t n = 0;
    if (s == CSV_OK) s = csv_sigs_fetch(ctx, b, out, cap, &n, nullptr);
    if (s == CSV_ERR_CAPACITY && n > b->sig_cap) {           // more signatures than the batch reserved: the reference has no such limit
        s = csv_batch_reserve_sigs(ctx, b, n);
        if (s == CSV_OK) s = csv_scan_run(ctx, b, &p);
        if (s == CSV_OK) s = csv_sigs_fetch(ctx, b, out, cap, &n, nullptr);
    }
    if (n_out) *n_out = n;
    csv_batch_free(ctx, b);
    return s;
}

int csv_dbscan1d_seg(csv_ctx* ctx, const int32_t* pts, const uint32_t* seg_id, uint64_t n, uint32_t n_seg, double eps, int min_pts,
                     int32_t* labels_out, int32_t* n_clusters_out)
{
    if (!ctx || (n && (!pts || !labels_out))) { set_error("csv_dbscan1d: null argument"); return CSV_ERR_ARG; }
    if (n_seg == 0) n_seg = 1;
    if (n_clusters_out) for (uint32_t i = 0; i < n_seg; i++) n_clusters_out[i] = 0;
    if (n == 0) return CSV_OK;
    if (n >= (1ull << 30)) { set_error("csv_dbscan1d: %llu points exceed the 2^30 limit", (unsigned long long)n); return CSV_ERR_LIMIT; }
    CSV_CUDA(cudaSetDevice(ctx->device));
    CSV_TRY(side_join(ctx));            // shares scratch buffers with the batch pipeline
    DevBuf& d_pts = ctx->sort_tmp[4]; DevBuf& d_misc = ctx->sort_tmp[5];
    CSV_TRY(d_pts.ensure(n * 4));
    // d_misc: labels | seg | n_clusters
    const size_t off_seg = n * 4, off_nc = off_seg + (seg_id ? n * 4 : 0);
    CSV_TRY(d_misc.ensure(off_nc + (size_t)n_seg * 4 + 16));
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemcpyAsync(d_pts.p, pts, n * 4, cudaMemcpyHostToDevice, st));
    uint32_t* d_seg = nullptr;
    if (seg_id) { d_seg = (uint32_t*)((char*)d_misc.p + off_seg); CSV_CUDA(cudaMemcpyAsync(d_seg, seg_id, n * 4, cudaMemcpyHostToDevice, st)); }
    int32_t* d_nc = (int32_t*)((char*)d_misc.p + off_nc);
    CSV_TRY(dbscan1d_device(ctx, d_pts.as<int32_t>(), d_seg, n, nullptr, n_seg, eps, min_pts, d_misc.as<int32_t>(), d_nc));
    CSV_CUDA(cudaMemcpyAsync(labels_out, d_misc.p, n * 4, cudaMemcpyDeviceToHost, st));
    if (n_clusters_out) CSV_CUDA(cudaMemcpyAsync(n_clusters_out, d_nc, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

int csv_dbscan1d(csv_ctx* ctx, const int32_t* pts, uint64_t n, double eps, int min_pts, int32_t* labels_out, int32_t* n_clusters_out)
{
    if (ctx && ctx->db_small && n > 0 && n <= (uint64_t)kDbSmallMax && pts && labels_out && eps >= 0.0) {
        CSV_CUDA(cudaSetDevice(ctx->device));
        return dbscan1d_small(ctx, pts, (uint32_t)n, eps, min_pts, labels_out, n_clusters_out);
    }
    return csv_dbscan1d_seg(ctx, pts, nullptr, n, 1, eps, min_pts, labels_out, n_clusters_out);
}

int csv_dbscan2d(csv_ctx* ctx, const uint32_t* start, const uint32_t* end, uint64_t n, double eps, int min_pts, int32_t* labels_out)
{
    if (!ctx || (n && (!start || !end || !labels_out))) { set_error("csv_dbscan2d: null argument"); return CSV_ERR_ARG; }
    if (n == 0) return CSV_OK;
    if (n >= (1ull << 30)) { set_error("csv_dbscan2d: %llu intervals exceed the 2^30 limit", (unsigned long long)n); return CSV_ERR_LIMIT; }
    CSV_CUDA(cudaSetDevice(ctx->device));
    CSV_TRY(side_join(ctx));
    DevBuf& d_in = ctx->sort_tmp[4]; DevBuf& d_lab = ctx->sort_tmp[5];
    CSV_TRY(d_in.ensure(n * 8)); CSV_TRY(d_lab.ensure(n * 4));
    cudaStream_t st = ctx->stream;
    uint32_t* d_start = d_in.as<uint32_t>(); uint32_t* d_end = d_start + n;
    CSV_CUDA(cudaMemcpyAsync(d_start, start, n * 4, cudaMemcpyHostToDevice, st));
    CSV_CUDA(cudaMemcpyAsync(d_end, end, n * 4, cudaMemcpyHostToDevice, st));
    CSV_TRY(dbscan2d_device(ctx, d_start, d_end, n, eps, min_pts, d_lab.as<int32_t>()));
    CSV_CUDA(cudaMemcpyAsync(labels_out, d_lab.p, n * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

uint64_t csv_largest_cluster(const int32_t* pts, const int32_t* labels, uint64_t n, int32_t* out)
{
    // dbscan1d.cpp:72-90: std::map order = ascending id, strictly-greater keeps the first;
    // with no id >= 0 the reference returns cluster_map[-1]
    int32_t max_id = -1;
    for (uint64_t i = 0; i < n; i++) if (labels[i] > max_id) max_id = labels[i];
    int32_t best = -1;
    if (max_id >= 0) {
        std::vector<uint64_t> cnt((size_t)max_id + 1, 0);
        for (uint64_t i = 0; i < n; i++) if (labels[i] >= 0) cnt[labels[i]]++;
        uint64_t best_n = 0;
        for (int32_t c = 0; c <= max_id; c++) if (cnt[c] > best_n) { best_n = cnt[c]; best = c; }
    }
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++) if (labels[i] == best) out[m++] = pts[i];
    return m;
}

}  // extern "C"

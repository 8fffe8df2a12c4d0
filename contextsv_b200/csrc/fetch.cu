// fetch.cu -- the depth map's way back to the host (csv_depth_fetch / csv_depth_fetch_all / csv_depth).
//
// The reference's interface hands the caller a uint32 per base (unordered_map<string, vector<uint32_t>>,
// sv_caller.cpp:788,801): 12.4 GB for a human genome.  A plain D2H of that is PCIe-bound (57 GB/s into pinned memory,
// 21 GB/s into the pageable vector the drop-in really gets), and it was 75-85 % of the host-facing step.  Depths are
// small numbers, so the map crosses the link as BYTES and the host widens them:
//
//   device   k_depth_narrow: u32 -> u8, values >= 255 stored as 255 and appended to a short (index, value) list
//   PCIe     one cudaMemcpyAsync per chunk (header + list + bytes) into a pinned staging ring, own copy stream
//   host     worker threads widen u8 -> u32 straight into the caller's array with streaming stores (widen.cpp), 256 Ki
//            positions at a time, then patch the listed values; a chunk whose list overflowed is fetched again as
//            plain 32-bit words
//
// The result is bit-identical to the plain copy for every input; the plain copy stays for short fetches and for
// threads == 0.  B200 host (16 cores, scripts/fetch_sweep.py): the 12.4 GB of a genome in 101-108 ms (14 threads,
// 118 GB/s into host memory; widening alone peaks at 125-140 GB/s, scripts/fetch_probe.cu) instead of 216 ms into
// pinned / 576 ms into pageable memory.  The hosts of the pool differ and are shared: the same call took 200-240 ms on a
// busy box, where the plain copy took 269 / 724 ms.
#include "batch.cuh"

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace csv {

// layout of one staging slot: [u32 count | 3 x u32 pad | exc_cap x {u32 index in chunk, u32 value} | n bytes]
__global__ void __launch_bounds__(256) k_depth_narrow(const uint4* __restrict__ src, uint32_t n, uint32_t* __restrict__ hdr, uint32_t exc_cap,
                                                      uint32_t* __restrict__ out)
{
    const uint32_t nq = (n + 3) >> 2;                    // the slab behind a region is tile-padded: reading the last quad whole is safe
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const uint4 v = ld_nc_v4(src + q);
        out[q] = min(v.x, 255u) | (min(v.y, 255u) << 8) | (min(v.z, 255u) << 16) | (min(v.w, 255u) << 24);
        if (max(max(v.x, v.y), max(v.z, v.w)) >= 255u) {     // rare: coverage in the hundreds
            const uint32_t val[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (val[k] >= 255u && 4 * q + k < n) {
                    const uint32_t slot = atomicAdd(hdr, 1u);
                    if (slot < exc_cap) { hdr[4 + 2 * slot] = 4 * q + k; hdr[5 + 2 * slot] = val[k]; }
                }
        }
    }
}

static int fetch_plain(csv_ctx* ctx, const std::vector<FetchSeg>& segs)
{
    for (auto& s : segs)
        if (s.len) CSV_CUDA(cudaMemcpyAsync(s.dst, s.src, s.len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    return CSV_OK;
}

void fetch_release(csv_ctx* ctx)
{
    FetchState& f = ctx->fetch;
    for (auto e : f.ev_narrow) cudaEventDestroy(e);
    for (auto e : f.ev_copy) cudaEventDestroy(e);
    f.ev_narrow.clear(); f.ev_copy.clear();
    if (f.h) cudaFreeHost(f.h);
    if (f.d) cudaFree(f.d);
    f.h = nullptr; f.d = nullptr; f.slots = 0; f.slot_bytes = 0;
    if (f.copy_stream) { cudaStreamDestroy(f.copy_stream); f.copy_stream = nullptr; }
}

static int ensure_ring(csv_ctx* ctx, int slots)
{
    FetchState& f = ctx->fetch;
    const size_t slot_bytes = ((size_t)16 + (size_t)f.exc_cap * 8 + f.chunk + 255) & ~(size_t)255;
    if (f.slots >= slots && f.slot_bytes == slot_bytes) return CSV_OK;
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaStream_t keep = f.copy_stream; f.copy_stream = nullptr;
    fetch_release(ctx);
    f.copy_stream = keep;
    if (!f.copy_stream) CSV_CUDA(cudaStreamCreateWithFlags(&f.copy_stream, cudaStreamNonBlocking));
    CSV_CUDA(cudaHostAlloc((void**)&f.h, slot_bytes * slots, cudaHostAllocDefault));
    CSV_CUDA(cudaMalloc((void**)&f.d, slot_bytes * slots));
    for (int i = 0; i < slots; i++) {
        cudaEvent_t a = nullptr, b = nullptr;
        CSV_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); f.ev_narrow.push_back(a);
        CSV_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming)); f.ev_copy.push_back(b);
    }
    f.slots = slots; f.slot_bytes = slot_bytes;
    return CSV_OK;
}

namespace {
constexpr uint32_t kPart = 256u << 10;      // positions one worker widens at a time: a chunk is shared by chunk / kPart workers
struct Chunk { const uint32_t* src; uint32_t* dst; uint32_t n; uint32_t first_task, parts; };
struct Job {
    std::mutex m;
    std::condition_variable cv_issue, cv_done;
    size_t issued = 0, next = 0;             // tasks (parts of chunks) whose bytes are on their way / handed to a worker
    bool closed = false;
    std::vector<uint32_t> task_chunk;
    std::vector<uint32_t> remaining;         // parts of the chunk still being widened
    std::vector<uint8_t> done;               // chunk has left its staging slot
    std::vector<size_t> fallback;
    cudaError_t error = cudaSuccess;
};
}  // namespace

int fetch_depth_segments(csv_ctx* ctx, const std::vector<FetchSeg>& segs)
{
    FetchState& f = ctx->fetch;
    size_t total = 0;
    for (auto& s : segs) total += s.len;
    if (f.threads <= 0 || total < f.min_len) return fetch_plain(ctx, segs);

    // A chunk is widened by several workers (kPart positions each), so the staging ring stays at a few slots (12 MB by
    // default) whatever the thread count: what the copy engine writes is then still in the host's last-level cache when
    // a worker reads it, and only the 4 bytes per base of the result go to DRAM.  With one worker per chunk and a ring of
    // threads + 6 slots, 1 / 2 / 4 Mi-position chunks took 198 / 229 / 245 ms on the same box.
    Job job;
    std::vector<Chunk> chunks;
    for (auto& s : segs)
        for (size_t o = 0; o < s.len; o += f.chunk) {
            const uint32_t n = (uint32_t)std::min<size_t>(f.chunk, s.len - o), parts = (n + kPart - 1) / kPart;
            chunks.push_back({s.src + o, s.dst + o, n, (uint32_t)job.task_chunk.size(), parts});
            for (uint32_t p = 0; p < parts; p++) job.task_chunk.push_back((uint32_t)(chunks.size() - 1));
            job.remaining.push_back(parts);
        }
    const size_t n_tasks = job.task_chunk.size();
    const int n_threads = (int)std::min<size_t>(f.threads, n_tasks);
    const uint32_t parts_per_chunk = (f.chunk + kPart - 1) / kPart;
    const int slots = (n_threads + (int)parts_per_chunk - 1) / (int)parts_per_chunk + 4;   // chunks being widened + chunks in flight
    CSV_TRY(ensure_ring(ctx, slots));

    job.done.assign(chunks.size(), 0);
    const int device = ctx->device;
    auto worker = [&]() {
        cudaSetDevice(device);
        for (;;) {
            size_t t;
            {
                std::unique_lock<std::mutex> lk(job.m);
                job.cv_issue.wait(lk, [&] { return job.next < job.issued || job.closed; });
                if (job.next >= job.issued) return;
                t = job.next++;
            }
            const size_t c = job.task_chunk[t];
            const Chunk& ch = chunks[c];
            const int slot = (int)(c % (size_t)slots);
            const uint8_t* h = f.h + (size_t)slot * f.slot_bytes;
            const uint32_t* hdr = (const uint32_t*)h;
            const cudaError_t e = cudaEventSynchronize(f.ev_copy[slot]);
            const bool overflow = e == cudaSuccess && hdr[0] > f.exc_cap;
            if (e == cudaSuccess && !overflow) {
                const size_t o = (size_t)(t - ch.first_task) * kPart;
                csv_host_widen_u8(h + 16 + (size_t)f.exc_cap * 8 + o, ch.dst + o, std::min<size_t>(kPart, ch.n - o));
            }
            bool last;
            {
                std::lock_guard<std::mutex> lk(job.m);
                if (e != cudaSuccess && job.error == cudaSuccess) job.error = e;
                last = --job.remaining[c] == 0;
            }
            if (!last) continue;
            if (e == cudaSuccess && !overflow)                       // every part is in place: the values that did not fit a byte
                for (uint32_t i = 0; i < hdr[0]; i++) ch.dst[hdr[4 + 2 * i]] = hdr[5 + 2 * i];
            {
                std::lock_guard<std::mutex> lk(job.m);
                if (overflow) job.fallback.push_back(c);
                job.done[c] = 1;
            }
            job.cv_done.notify_all();
        }
    };
    std::vector<std::thread> pool;
    try {
        for (int t = 0; t < n_threads; t++) pool.emplace_back(worker);
    } catch (...) {                                                // out of threads: go on with the ones that started
    }
    if (pool.empty()) return fetch_plain(ctx, segs);               // nothing is in flight yet

    cudaError_t err = cudaSuccess;
    const uint32_t grid_cap = (uint32_t)ctx->sm_count * 8;
    for (size_t c = 0; c < chunks.size() && err == cudaSuccess; c++) {
        const int slot = (int)(c % (size_t)slots);
        if (c >= (size_t)slots) {                                  // the slot's previous chunk must have left the staging buffers
            std::unique_lock<std::mutex> lk(job.m);
            job.cv_done.wait(lk, [&] { return job.done[c - slots] != 0; });
            if (job.error != cudaSuccess) break;
        }
        uint8_t* d = f.d + (size_t)slot * f.slot_bytes;
        uint8_t* h = f.h + (size_t)slot * f.slot_bytes;
        const uint32_t n = chunks[c].n, nq = (n + 3) / 4;
        const size_t bytes = (size_t)16 + (size_t)f.exc_cap * 8 + (((size_t)n + 3) & ~(size_t)3);
        if ((err = cudaMemsetAsync(d, 0, 4, ctx->stream)) != cudaSuccess) break;
        k_depth_narrow<<<std::min((nq + 255) / 256, grid_cap), 256, 0, ctx->stream>>>((const uint4*)chunks[c].src, n, (uint32_t*)d, f.exc_cap,
                                                                                         (uint32_t*)(d + 16 + (size_t)f.exc_cap * 8));
        ctx->launches++;
        if ((err = cudaGetLastError()) != cudaSuccess) break;
        if ((err = cudaEventRecord(f.ev_narrow[slot], ctx->stream)) != cudaSuccess) break;
        if ((err = cudaStreamWaitEvent(f.copy_stream, f.ev_narrow[slot], 0)) != cudaSuccess) break;
        if ((err = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, f.copy_stream)) != cudaSuccess) break;
        if ((err = cudaEventRecord(f.ev_copy[slot], f.copy_stream)) != cudaSuccess) break;
        {
            std::lock_guard<std::mutex> lk(job.m);
            job.issued = chunks[c].first_task + chunks[c].parts;
        }
        job.cv_issue.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(job.m);
        job.closed = true;
    }
    job.cv_issue.notify_all();
    for (auto& t : pool) t.join();
    cudaStreamSynchronize(f.copy_stream);
    if (err == cudaSuccess) err = job.error;
    if (err != cudaSuccess) { set_error("depth fetch: %s", cudaGetErrorString(err)); cudaStreamSynchronize(ctx->stream); return CSV_ERR_CUDA; }
    f.narrow_chunks += chunks.size() - job.fallback.size();
    f.fallback_chunks += job.fallback.size();
    for (size_t c : job.fallback)
        CSV_CUDA(cudaMemcpyAsync(chunks[c].dst, chunks[c].src, (size_t)chunks[c].n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CSV_CUDA(cudaStreamSynchronize(ctx->stream));
    return CSV_OK;
}

}  // namespace csv

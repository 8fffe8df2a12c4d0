// windows.cu -- window depth sums for the log2 ratio
// (CNVCaller::querySNPRegion, cnv_caller.cpp:76-113), read from the
// device-resident depth of a whole-contig region so the depth map need not
// cross PCIe for this consumer.  Only the integer parts are computed here
// (sum of depths, number of positions); the caller divides and takes log2 with
// the host libm, exactly as the reference does.
//
// The window positions replicate the reference's double arithmetic
// `(uint32_t)(start_pos + i * pos_step + j)` operation by operation
// (__dmul_rn / __dadd_rn: no FMA contraction), one warp per window.
#include "batch.cuh"

namespace csv {

__global__ void k_window_sums(const uint32_t* __restrict__ depth, uint32_t beg, uint32_t lim /* min(end, map_size) */, uint32_t n_sv, const uint32_t* __restrict__ start,
                              const uint32_t* __restrict__ end, int sample_size, unsigned long long* sum_out, uint32_t* cnt_out)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_win = (uint64_t)n_sv * (uint64_t)sample_size;
    for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_win; w += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
        const uint32_t sv = (uint32_t)(w / (uint64_t)sample_size);
        const int i = (int)(w % (uint64_t)sample_size);
        const uint32_t s = start[sv], e = end[sv];
        unsigned long long sum = 0; uint32_t cnt = 0;
        if (s <= e) {                                                       // cnv_caller.cpp:69-73
            const double pos_step = __ddiv_rn((double)(uint32_t)(e - s + 1u), (double)sample_size);     // :76
            const double base = __dadd_rn((double)s, __dmul_rn((double)i, pos_step));
            for (int j = (int)lane; (double)j < pos_step; j += 32) {        // :86
                const uint32_t pos = (uint32_t)__dadd_rn(base, (double)j);
                if (pos > e) break;                                         // :89-92 (monotone in j)
                if (pos >= beg && pos < lim) { sum += depth[pos - beg]; cnt++; }   // :93-96, this slice's share
            }
        }
        sum = warp_sum_u64(sum); cnt = warp_sum_u32(cnt);
        if (lane == 0) { sum_out[w] = sum; cnt_out[w] = cnt; }
    }
}

// depth at arbitrary positions (SVCaller::getReadDepth, sv_caller.cpp:1332-1344): 0 beyond the map, like the caught
// std::out_of_range there
__global__ void k_depth_at(const uint32_t* __restrict__ depth, uint32_t beg, uint32_t lim, uint64_t n, const uint32_t* __restrict__ pos, uint32_t* out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = (pos[i] >= beg && pos[i] < lim) ? depth[pos[i] - beg] : 0u;
}

// getReadDepth for positions anywhere in the batch: (tid, pos) is looked up among the batch's regions
struct RegionLookup { const TidDev* tids; const RegionDev* regs; const uint32_t* reg_tab; uint32_t n_tids; const uint32_t* depth; };
__device__ __forceinline__ uint32_t depth_lookup(const RegionLookup& L, uint32_t tid, uint32_t pos)
{
    if (tid >= L.n_tids) return kNone;
    const TidDev td = L.tids[tid];
    if (td.count == 0) return kNone;
    if (pos >= td.map_size) return 0u;                                      // the reference catches the out_of_range and adds nothing
    for (uint32_t r = td.first; r < td.first + td.count; r++)
        if (pos >= L.regs[r].beg && pos < L.regs[r].end) return L.depth[(size_t)L.regs[r].tile_base * kTile + (pos - L.regs[r].beg)];
    return kNone;                                                           // inside the contig, outside this batch's slices: another shard's
}
__global__ void k_depth_at_tid(const RegionLookup L, uint64_t n, const int32_t* __restrict__ tid, const uint32_t* __restrict__ pos, uint32_t* out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = depth_lookup(L, (uint32_t)tid[i], pos[i]);
}
// ... and for every signature of the batch, at its start (sv_caller.cpp:1306), without the positions leaving the device
__global__ void k_sigs_depth(const RegionLookup L, const uint32_t* scalars, const uint32_t* __restrict__ o_start, const uint32_t* __restrict__ o_seg,
                             const int32_t* __restrict__ region_tid, uint32_t* out)
{
    const uint32_t n = scalars[SC_N_SIG_EFF];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = depth_lookup(L, (uint32_t)region_tid[o_seg[i] >> 1], o_start[i]);
}

// Position-weighted checksum of a depth slice: sum over i of depth[i] * mix(tid, beg + i) modulo 2^64.  Additive, so
// the checksums of the shards of a contig add up to the checksum of the whole contig however it was cut.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) k_depth_checksum(const uint32_t* __restrict__ depth, uint32_t tid, uint32_t beg, uint32_t len, unsigned long long* out)
{
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x)
        acc += (unsigned long long)depth[i] * mix64(((unsigned long long)tid << 32) | (beg + (uint32_t)i));
    acc = warp_sum_u64(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

}  // namespace csv

using namespace csv;

extern "C" int csv_depth_at(csv_ctx* ctx, csv_batch* b, uint32_t region, uint64_t n, const uint32_t* positions, uint32_t* depth_out)
{
    if (!ctx || !b || (n && (!positions || !depth_out))) { set_error("csv_depth_at: bad argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_at: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    if (region >= b->n_regions) { set_error("region %u out of range", region); return CSV_ERR_ARG; }
    const csv_region& g = b->regions[region];
    if (n == 0) return CSV_OK;
    CSV_TRY(side_join(ctx));
    DevBuf& io = ctx->sort_tmp[4];
    CSV_TRY(io.ensure((size_t)n * 8));
    uint32_t* d_pos = io.as<uint32_t>(); uint32_t* d_out = d_pos + n;
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemcpyAsync(d_pos, positions, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const uint32_t* depth = b->d_depth.as<uint32_t>() + (size_t)b->tile_base[region] * kTile;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    k_depth_at<<<grid, 256, 0, st>>>(depth, g.beg, g.end < g.map_size ? g.end : g.map_size, n, d_pos, d_out);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaMemcpyAsync(depth_out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

extern "C" int csv_window_sums(csv_ctx* ctx, csv_batch* b, uint32_t region, uint32_t n_sv, const uint32_t* start_pos,
                               const uint32_t* end_pos, int sample_size, uint64_t* sum_out, uint32_t* count_out)
{
    if (!ctx || !b || !start_pos || !end_pos || !sum_out || !count_out || sample_size <= 0) { set_error("csv_window_sums: bad argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_window_sums: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    if (region >= b->n_regions) { set_error("region %u out of range", region); return CSV_ERR_ARG; }
    const csv_region& g = b->regions[region];
    if (n_sv == 0) return CSV_OK;
    CSV_TRY(side_join(ctx));
    const size_t n_win = (size_t)n_sv * sample_size;
    DevBuf& in = ctx->sort_tmp[4]; DevBuf& out = ctx->sort_tmp[5];
    CSV_TRY(in.ensure((size_t)n_sv * 8));
    CSV_TRY(out.ensure(n_win * 12 + 16));
    uint32_t* d_s = in.as<uint32_t>(); uint32_t* d_e = d_s + n_sv;
    unsigned long long* d_sum = out.as<unsigned long long>(); uint32_t* d_cnt = (uint32_t*)(d_sum + n_win);
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemcpyAsync(d_s, start_pos, (size_t)n_sv * 4, cudaMemcpyHostToDevice, st));
    CSV_CUDA(cudaMemcpyAsync(d_e, end_pos, (size_t)n_sv * 4, cudaMemcpyHostToDevice, st));
    const uint32_t* depth = b->d_depth.as<uint32_t>() + (size_t)b->tile_base[region] * kTile;
    uint64_t warps = n_win; uint32_t grid = (uint32_t)std::min<uint64_t>((warps + 7) / 8, (uint64_t)ctx->sm_count * 16);
    k_window_sums<<<grid, 256, 0, st>>>(depth, g.beg, g.end < g.map_size ? g.end : g.map_size, n_sv, d_s, d_e, sample_size, d_sum, d_cnt);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaMemcpyAsync(sum_out, d_sum, n_win * 8, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaMemcpyAsync(count_out, d_cnt, n_win * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

static RegionLookup region_lookup(csv_batch* b)
{
    RegionLookup L;
    L.tids = b->d_tids.as<TidDev>(); L.regs = b->d_regs.as<RegionDev>(); L.reg_tab = b->d_reg_tab.as<uint32_t>(); L.n_tids = b->n_tids;
    L.depth = b->d_depth.as<uint32_t>();
    return L;
}

extern "C" int csv_depth_at_tid(csv_ctx* ctx, csv_batch* b, uint64_t n, const int32_t* tid, const uint32_t* positions, uint32_t* depth_out)
{
    if (!ctx || !b || (n && (!tid || !positions || !depth_out))) { set_error("csv_depth_at_tid: bad argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_at_tid: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    if (n == 0) return CSV_OK;
    CSV_TRY(side_join(ctx));
    DevBuf& io = ctx->sort_tmp[4];
    CSV_TRY(io.ensure((size_t)n * 12));
    uint32_t* d_pos = io.as<uint32_t>(); int32_t* d_tid = (int32_t*)(d_pos + n); uint32_t* d_out = d_pos + 2 * n;
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemcpyAsync(d_pos, positions, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CSV_CUDA(cudaMemcpyAsync(d_tid, tid, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    k_depth_at_tid<<<grid, 256, 0, st>>>(region_lookup(b), n, d_tid, d_pos, d_out);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaMemcpyAsync(depth_out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

extern "C" int csv_sigs_depth(csv_ctx* ctx, csv_batch* b, uint32_t* depth_out, uint64_t cap)
{
    if (!ctx || !b || !depth_out) { set_error("csv_sigs_depth: bad argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth || !b->have_sigs) { set_error("csv_sigs_depth: run csv_scan_run with want_depth and want_sigs first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    uint64_t n = 0;
    CSV_TRY(csv_sigs_count(ctx, b, &n));                                   // joins the side and tile streams, checks the scan
    if (n > cap) { set_error("csv_sigs_depth: %llu signatures, caller capacity %llu", (unsigned long long)n, (unsigned long long)cap); return CSV_ERR_CAPACITY; }
    if (n == 0) return CSV_OK;
    DevBuf& io = ctx->sort_tmp[4];
    CSV_TRY(io.ensure((size_t)n * 4 + (size_t)b->n_regions * 4));
    uint32_t* d_out = io.as<uint32_t>(); int32_t* d_rt = (int32_t*)(d_out + n);
    std::vector<int32_t> rt(b->n_regions);
    for (uint32_t i = 0; i < b->n_regions; i++) rt[i] = b->regions[i].tid;
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemcpyAsync(d_rt, rt.data(), rt.size() * 4, cudaMemcpyHostToDevice, st));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    k_sigs_depth<<<grid, 256, 0, st>>>(region_lookup(b), b->d_scalars.as<uint32_t>(), b->d_out_start.as<uint32_t>(), b->d_out_seg.as<uint32_t>(), d_rt, d_out);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaMemcpyAsync(depth_out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));                                  // also keeps rt alive until its copy is through
    return CSV_OK;
}

extern "C" int csv_depth_checksum(csv_ctx* ctx, csv_batch* b, uint64_t* checksum_out /* [n_regions] */)
{
    if (!ctx || !b || !checksum_out) { set_error("csv_depth_checksum: bad argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_checksum: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    CSV_TRY(side_join(ctx));
    DevBuf& out = ctx->sort_tmp[5];
    CSV_TRY(out.ensure((size_t)b->n_regions * 8));
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemsetAsync(out.p, 0, (size_t)b->n_regions * 8, st));
    for (uint32_t r = 0; r < b->n_regions; r++) {
        const csv_region& g = b->regions[r];
        const uint32_t len = g.end - g.beg;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(((uint64_t)len + 255) / 256, (uint64_t)ctx->sm_count * 16);
        k_depth_checksum<<<grid, 256, 0, st>>>(b->d_depth.as<uint32_t>() + (size_t)b->tile_base[r] * kTile, (uint32_t)g.tid, g.beg, len, out.as<unsigned long long>() + r);
        ctx->launches++;
    }
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaMemcpyAsync(checksum_out, out.p, (size_t)b->n_regions * 8, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

extern "C" int csv_debug_fetch(csv_ctx* ctx, csv_batch* b, const char* name, uint64_t offset, uint64_t bytes, void* out, uint64_t* size_out)
{
    if (!ctx || !b || !name) { set_error("csv_debug_fetch: bad argument"); return CSV_ERR_ARG; }
    CSV_CUDA(cudaSetDevice(ctx->device));       // the caller may be a thread that last used another device (CONTEXTSV_GPUS)
    const size_t nr = b->n_reads, nt = b->n_tiles;
    struct { const char* name; const DevBuf* buf; size_t bytes; } tab[] = {
        {"events", &b->d_events, (size_t)b->ev_cap * 4}, {"ev_start", &b->d_ev_start, (nr + 1) * 4}, {"ref_end", &b->d_ref_end, nr * 4},
        {"span_desc", &b->d_span_desc, ((size_t)b->n_spans + 1) * 16}, {"pmax", &b->d_pmax, nr * 8}, {"tile_q", &b->d_tile_q, nt * 16},
        {"tile_r", &b->d_tile_r, nt * 8}, {"meta", &b->d_meta, nr * 16}, {"key", &b->d_key, nr * 8},
    };
    for (const auto& e : tab) {
        if (strcmp(e.name, name)) continue;
        if (size_out) *size_out = e.bytes;
        if (bytes == 0) return CSV_OK;
        if (!out || !e.buf->p || offset + bytes > e.bytes) { set_error("csv_debug_fetch: %s holds %zu bytes", name, e.bytes); return CSV_ERR_ARG; }
        CSV_TRY(side_join(ctx));
        CSV_CUDA(cudaMemcpyAsync(out, (const char*)e.buf->p + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CSV_CUDA(cudaStreamSynchronize(ctx->stream));
        return CSV_OK;
    }
    set_error("csv_debug_fetch: unknown array %s", name);
    return CSV_ERR_ARG;
}

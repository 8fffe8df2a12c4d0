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

__global__ void k_window_sums(const uint32_t* __restrict__ depth, uint32_t map_size, uint32_t n_sv, const uint32_t* __restrict__ start,
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
                if (pos < map_size) { sum += depth[pos]; cnt++; }           // :93-96
            }
        }
        sum = warp_sum_u64(sum); cnt = warp_sum_u32(cnt);
        if (lane == 0) { sum_out[w] = sum; cnt_out[w] = cnt; }
    }
}

// depth at arbitrary positions (SVCaller::getReadDepth, sv_caller.cpp:1332-1344): 0 beyond the map, like the caught
// std::out_of_range there
__global__ void k_depth_at(const uint32_t* __restrict__ depth, uint32_t map_size, uint64_t n, const uint32_t* __restrict__ pos, uint32_t* out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = pos[i] < map_size ? depth[pos[i]] : 0u;
}

}  // namespace csv

using namespace csv;

extern "C" int csv_depth_at(csv_ctx* ctx, csv_batch* b, uint32_t region, uint64_t n, const uint32_t* positions, uint32_t* depth_out)
{
    if (!ctx || !b || (n && (!positions || !depth_out))) { set_error("csv_depth_at: bad argument"); return CSV_ERR_ARG; }
    if (!b->scanned || !b->have_depth) { set_error("csv_depth_at: run csv_scan_run with want_depth first"); return CSV_ERR_STATE; }
    if (region >= b->n_regions) { set_error("region %u out of range", region); return CSV_ERR_ARG; }
    const csv_region& g = b->regions[region];
    if (g.beg != 0 || g.end != g.map_size) { set_error("csv_depth_at needs a whole-contig region"); return CSV_ERR_ARG; }
    if (n == 0) return CSV_OK;
    CSV_TRY(side_join(ctx));
    DevBuf& io = ctx->sort_tmp[4];
    CSV_TRY(io.ensure((size_t)n * 8));
    uint32_t* d_pos = io.as<uint32_t>(); uint32_t* d_out = d_pos + n;
    cudaStream_t st = ctx->stream;
    CSV_CUDA(cudaMemcpyAsync(d_pos, positions, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const uint32_t* depth = b->d_depth.as<uint32_t>() + (size_t)b->tile_base[region] * kTile;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    k_depth_at<<<grid, 256, 0, st>>>(depth, g.map_size, n, d_pos, d_out);
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
    if (region >= b->n_regions) { set_error("region %u out of range", region); return CSV_ERR_ARG; }
    const csv_region& g = b->regions[region];
    if (g.beg != 0 || g.end != g.map_size) { set_error("csv_window_sums needs a whole-contig region"); return CSV_ERR_ARG; }
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
    k_window_sums<<<grid, 256, 0, st>>>(depth, g.map_size, n_sv, d_s, d_e, sample_size, d_sum, d_cnt);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    CSV_CUDA(cudaMemcpyAsync(sum_out, d_sum, n_win * 8, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaMemcpyAsync(count_out, d_cnt, n_win * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

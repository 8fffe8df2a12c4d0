// records.cu -- per-record alignment summaries for the split-read pass: what
// SVCaller::detectSVsFromSplitReads reads off every primary / supplementary record before it starts matching them
// (sv_caller.cpp:150-162): bam_endpos(b) and SVCaller::getAlignmentReadPositions(b) (sv_caller.cpp:663-690).
//   endpos      = pos + max(1, reference bases consumed)      (htslib bam_endpos; unmapped records consume none)
//   query_end   = sum of the lengths of M / I / S / = / X ops
//   query_start = that sum taken over the ops before the first M / I / = / X op (the leading soft clip); 0 if there is none
// One warp per record, lanes stride over its CIGAR words (coalesced for HiFi and ONT alike), records in csv_reads order
// (empty CIGARs included: endpos = pos + 1, 0, 0).
#include "batch.cuh"

namespace csv {

constexpr uint32_t kQStartMask = (1u << 0) | (1u << 1) | (1u << 7) | (1u << 8);     // M I = X

__global__ void __launch_bounds__(256) k_record_summary(const unsigned long long* __restrict__ cig_off, const uint32_t* __restrict__ cigar,
                                                         const int32_t* __restrict__ pos0, const uint16_t* __restrict__ flag, uint32_t n_reads,
                                                         int32_t* endpos, int32_t* qstart, int32_t* qend)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += warps) {
        const unsigned long long o0 = cig_off[r], o1 = cig_off[r + 1];
        uint32_t ref = 0, qry = 0;
        unsigned long long first = o1;                                      // first op that starts the aligned part of the query
        for (unsigned long long o = o0 + lane; o < o1; o += 32) {
            const uint32_t w = __ldg(cigar + o), op = w & 15u, len = w >> 4;
            if ((kRefMask >> op) & 1u) ref += len;
            if ((kQryMask >> op) & 1u) qry += len;
            if (((kQStartMask >> op) & 1u) && o < first) first = o;
        }
        ref = __reduce_add_sync(0xffffffffu, ref);
        qry = __reduce_add_sync(0xffffffffu, qry);
        const uint32_t f_lo = __reduce_min_sync(0xffffffffu, (uint32_t)(first - o0 < 0xffffffffull ? first - o0 : 0xffffffffull));
        uint32_t lead = 0;
        if ((unsigned long long)f_lo < o1 - o0) {                           // there is such an op: sum the query-consuming ops before it
            for (unsigned long long o = o0 + lane; o < o0 + f_lo; o += 32) {
                const uint32_t w = __ldg(cigar + o);
                if ((kQryMask >> (w & 15u)) & 1u) lead += w >> 4;
            }
            lead = __reduce_add_sync(0xffffffffu, lead);
        }
        if (lane == 0) {
            const uint32_t rlen = (flag[r] & 0x4u) ? 0u : ref;             // BAM_FUNMAP
            if (endpos) endpos[r] = (int32_t)((uint32_t)pos0[r] + (rlen ? rlen : 1u));
            if (qstart) qstart[r] = (int32_t)lead;
            if (qend) qend[r] = (int32_t)qry;
        }
    }
}

}  // namespace csv

using namespace csv;

extern "C" int csv_record_summary(csv_ctx* ctx, csv_batch* b, int32_t* endpos_out, int32_t* query_start_out, int32_t* query_end_out)
{
    if (!ctx || !b) { set_error("csv_record_summary: null argument"); return CSV_ERR_ARG; }
    if (b->inputs_released) { set_error("csv_record_summary: the batch's inputs were released"); return CSV_ERR_STATE; }
    const uint32_t n = b->n_reads;
    if (n == 0 || (!endpos_out && !query_start_out && !query_end_out)) return CSV_OK;
    CSV_CUDA(cudaSetDevice(ctx->device));
    CSV_TRY(side_join(ctx));
    CSV_TRY(wait_upload(ctx, b, 0xffffffffu));
    DevBuf& out = ctx->sort_tmp[5];
    CSV_TRY(out.ensure((size_t)n * 12));
    int32_t* d_e = out.as<int32_t>(); int32_t* d_s = d_e + n; int32_t* d_q = d_s + n;
    cudaStream_t st = ctx->stream;
    const uint64_t warps = n;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((warps + 7) / 8, (uint64_t)ctx->sm_count * 32);
    k_record_summary<<<grid, 256, 0, st>>>(b->d_cig_off.as<unsigned long long>(), b->d_cigar.as<uint32_t>(), b->d_pos0.as<int32_t>(),
                                           b->d_flag.as<uint16_t>(), n, d_e, d_s, d_q);
    ctx->launches++;
    CSV_CUDA(cudaGetLastError());
    if (endpos_out) CSV_CUDA(cudaMemcpyAsync(endpos_out, d_e, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (query_start_out) CSV_CUDA(cudaMemcpyAsync(query_start_out, d_s, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (query_end_out) CSV_CUDA(cudaMemcpyAsync(query_end_out, d_q, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CSV_CUDA(cudaStreamSynchronize(st));
    return CSV_OK;
}

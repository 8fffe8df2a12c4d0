// gpu_context.h -- one csv_ctx per host thread (the reference calls the hot path from ThreadPool
// workers, src/sv_caller.cpp:828-851; a csv_ctx is not thread-safe by design).
// Device selection comes from the environment, never from new CLI flags:
//   CONTEXTSV_GPUS   comma-separated device ids to round-robin over (default "0")
#pragma once
#include "contextsv_b200.h"

#include <chrono>

#include <vector>

namespace csvhost {
std::vector<int> device_list();      // CONTEXTSV_GPUS
csv_ctx* thread_context();
// Starts CUDA initialisation (driver context, module load: ~1.5 s on a B200 box) on a background thread, once per
// process, so that it runs beside the BAM decoding instead of in front of the first GPU call.
void warm_up_async();

// CONTEXTSV_B200_STATS=1: calls and wall time spent under each drop-in entry point, printed to stderr at exit.
enum StatId { STAT_DEPTH = 0, STAT_DEPTH_GPU, STAT_CIGAR, STAT_CIGAR_GPU, STAT_DBSCAN1D, STAT_DBSCAN2D, STAT_CTX, STAT_CACHE_HIT, STAT_DECODE, STAT_WINDOWS, STAT_DEPTH_AT, STAT_SPLIT, STAT_COUNT };
void stat_add(int id, double seconds, unsigned long long items);
struct StatTimer {
    int id; unsigned long long items; std::chrono::steady_clock::time_point t0;
    explicit StatTimer(int i, unsigned long long n = 0) : id(i), items(n), t0(std::chrono::steady_clock::now()) {}
    ~StatTimer() { stat_add(id, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), items); }
};
}

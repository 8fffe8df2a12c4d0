// gpu_context.h -- one csv_ctx per host thread (the reference calls the hot path from ThreadPool
// workers, src/sv_caller.cpp:828-851; a csv_ctx is not thread-safe by design).
// Device selection comes from the environment, never from new CLI flags:
//   CONTEXTSV_GPUS   comma-separated device ids to round-robin over (default "0")
#pragma once
#include "contextsv_b200.h"

namespace csvhost {
csv_ctx* thread_context();
}

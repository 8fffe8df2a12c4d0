// fetch_probe: what bounds the D2H of the depth map on this box?
//   (1) plain cudaMemcpy D2H into pinned / pageable host memory (what csv_depth_fetch does today),
//   (2) host threads widening u8 -> u32 with non-temporal stores (what a narrow transfer would need).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -mavx2 -o scripts/fetch_probe scripts/fetch_probe.cu -lpthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void widen(const uint8_t* src, uint32_t* dst, size_t n, bool nt)
{
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        __m128i b = _mm_loadu_si128((const __m128i*)(src + i));
        __m256i lo = _mm256_cvtepu8_epi32(b);
        __m256i hi = _mm256_cvtepu8_epi32(_mm_srli_si128(b, 8));
        if (nt) { _mm256_stream_si256((__m256i*)(dst + i), lo); _mm256_stream_si256((__m256i*)(dst + i + 8), hi); }
        else { _mm256_storeu_si256((__m256i*)(dst + i), lo); _mm256_storeu_si256((__m256i*)(dst + i + 8), hi); }
    }
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
}

int main(int argc, char** argv)
{
    size_t n = (size_t)1 << 31;                         // 2 Gi positions: 2 GiB narrow, 8 GiB wide
    if (argc > 1) n = (size_t)atoll(argv[1]);
    unsigned hw = std::thread::hardware_concurrency();
    printf("hardware_concurrency %u\n", hw);
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, n * 4) != cudaSuccess) { printf("no device\n"); return 1; }
    cudaMemset(d, 1, n * 4);
    uint32_t* pinned = nullptr; cudaHostAlloc(&pinned, n * 4, cudaHostAllocDefault);
    uint32_t* pageable = (uint32_t*)aligned_alloc(4096, n * 4); memset(pageable, 0, n * 4);
    uint8_t* narrow = nullptr; cudaHostAlloc(&narrow, n, cudaHostAllocDefault); memset(narrow, 7, n);
    for (int rep = 0; rep < 2; rep++) {
        double t0 = now(); cudaMemcpy(pinned, d, n * 4, cudaMemcpyDeviceToHost); double t1 = now();
        printf("D2H pinned    %.1f ms  %.1f GB/s\n", 1e3 * (t1 - t0), n * 4 / (t1 - t0) / 1e9);
    }
    for (int rep = 0; rep < 2; rep++) {
        double t0 = now(); cudaMemcpy(pageable, d, n * 4, cudaMemcpyDeviceToHost); double t1 = now();
        printf("D2H pageable  %.1f ms  %.1f GB/s\n", 1e3 * (t1 - t0), n * 4 / (t1 - t0) / 1e9);
    }
    { double t0 = now(); cudaMemcpy(narrow, d, n, cudaMemcpyDeviceToHost); double t1 = now();
      printf("D2H pinned (n bytes) %.1f ms  %.1f GB/s\n", 1e3 * (t1 - t0), n / (t1 - t0) / 1e9); }
    { double t0 = now(); cudaMemcpy(d, pinned, n * 4, cudaMemcpyHostToDevice); double t1 = now();
      printf("H2D pinned    %.1f ms  %.1f GB/s\n", 1e3 * (t1 - t0), n * 4 / (t1 - t0) / 1e9); }
    for (int dst_kind = 0; dst_kind < 2; dst_kind++)
        for (int nt = 1; nt >= 0; nt--)
            for (unsigned T : {1u, 2u, 4u, 8u, 16u, 32u, 64u}) {
                if (T > hw && T > 16) continue;
                uint32_t* dst = dst_kind ? pageable : pinned;
                double best = 1e9;
                for (int rep = 0; rep < 2; rep++) {
                    std::vector<std::thread> th;
                    double t0 = now();
                    for (unsigned t = 0; t < T; t++)
                        th.emplace_back([=] {
                            // 8 Mi-position chunks dealt round-robin, like a chunked fetch would
                            const size_t C = (size_t)8 << 20;
                            for (size_t c = t; c * C < n; c += T) { size_t b = c * C, e = b + C < n ? b + C : n; widen(narrow + b, dst + b, e - b, nt); }
                        });
                    for (auto& x : th) x.join();
                    double dt = now() - t0; if (dt < best) best = dt;
                }
                printf("widen %-8s %-3s T=%-2u  %.1f ms  %.1f GB/s written\n", dst_kind ? "pageable" : "pinned", nt ? "nt" : "st", T, 1e3 * best, n * 4 / best / 1e9);
            }
    // while a D2H of narrow bytes is in flight (shares the memory system)
    {
        unsigned T = hw < 16 ? hw : 16;
        cudaStream_t s; cudaStreamCreate(&s);
        double t0 = now();
        cudaMemcpyAsync(narrow, d, n, cudaMemcpyDeviceToHost, s);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([=] { const size_t C = (size_t)8 << 20; for (size_t c = t; c * C < n; c += T) { size_t b = c * C, e = b + C < n ? b + C : n; widen(narrow + b, pinned + b, e - b, true); } });
        for (auto& x : th) x.join();
        double t1 = now();
        cudaStreamSynchronize(s);
        double t2 = now();
        printf("widen nt T=%u beside a narrow D2H: widen %.1f ms, both %.1f ms\n", T, 1e3 * (t1 - t0), 1e3 * (t2 - t0));
    }
    uint64_t chk = 0; for (size_t i = 0; i < n; i += 4097) chk += pinned[i] + pageable[i];
    printf("check %llu\n", (unsigned long long)chk);
    return 0;
}

// bw_probe.cu -- calibration: what can a pure WRITE stream reach on this B200, versus the copy figure in
// MEASURED_PEAKS.json (read + write)?  The depth-tile kernel writes 12.35 GB per step and reads almost nothing.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void fill128(uint4* p, size_t n, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p + i), "r"(v) : "memory");
}
__global__ void fill256(uint4* p, size_t n, uint32_t v) {   // n in 32-byte units
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 2 * i), "r"(v) : "memory");
}
__global__ void fill256_line(uint4* p, size_t nlines, uint32_t v) {   // each thread writes one full 128-byte line (4 x 256-bit), like the tile kernel
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nlines; i += (size_t)gridDim.x * blockDim.x)
        for (int k = 0; k < 4; k++)
            asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 8 * i + 2 * k), "r"(v) : "memory");
}
__global__ void copy128(const uint4* a, uint4* b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void read128(const uint4* a, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = a[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678) *out = acc;
}
int main() {
    const size_t bytes = 12ull << 30;
    uint4 *a, *b; uint32_t* o;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, double gb, auto f) {
        float best = 1e9;
        for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        printf("%-28s %8.3f ms  %8.1f GB/s\n", name, best, gb / (best * 1e-3));
    };
    const int grid = 148 * 16;
    time("cudaMemsetAsync 12 GiB", bytes / 1e9, [&] { cudaMemsetAsync(a, 0, bytes); });
    time("fill st.cs.v4 (128-bit)", bytes / 1e9, [&] { fill128<<<grid, 256>>>(a, bytes / 16, 7); });
    time("fill st.v8 (256-bit)", bytes / 1e9, [&] { fill256<<<grid, 256>>>(a, bytes / 32, 7); });
    time("fill 4x st.v8 per line", bytes / 1e9, [&] { fill256_line<<<grid, 256>>>(a, bytes / 128, 7); });
    time("read 128-bit", bytes / 1e9, [&] { read128<<<grid, 256>>>(a, bytes / 16, o); });
    time("copy 128-bit (r+w bytes)", 2 * bytes / 1e9, [&] { copy128<<<grid, 256>>>(a, b, bytes / 16); });
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

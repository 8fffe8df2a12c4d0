// pipe_probe.cu -- issue rate of the integer instructions the tile/walk kernels lean on (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(256) k(int* out, int seed, int iters)
{
    int a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x;
    int b = seed * 3 + 1, c = seed ^ 0x55;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = a[i] + b + c;                                   // IADD3
                if (OP == 1) a[i] = a[i] * b + c;                                   // IMAD
                if (OP == 2) a[i] = __dp2a_lo(a[i], b, c);                          // IDP.2A
                if (OP == 3) a[i] = __dp4a(a[i], b, c);                             // IDP.4A
                if (OP == 4) a[i] = min(min(a[i], b), c) ^ r;                       // VIMNMX3 + LOP
                if (OP == 5) a[i] = (a[i] & b) ^ c;                                 // LOP3
                if (OP == 6) { a[i] = a[i] + b + c; a[(i + 1) & 7] = a[(i + 1) & 7] * b + i; }   // IADD3 + IMAD mix
                if (OP == 7) a[i] = __funnelshift_r(a[i], b, c);                    // SHF
                if (OP == 8) a[i] = __byte_perm(a[i], b, c);                        // PRMT
                if (OP == 9) a[i] = __popc(a[i]) + c;                               // POPC
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x12345678) out[0] = s;
}

template <int OP>
void run(const char* name, int per_iter)
{
    int* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = 148 * 8;
    k<OP><<<grid, 256>>>(d, 1, 16);
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(d, 1, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)grid * 8 * iters * per_iter;     // warp instructions
    printf("%-16s %8.3f ms  %7.1f G warp-inst/s  (%.2f /clk/SM at 1.965 GHz)\n", name, ms, winst / ms * 1e-6, winst / ms * 1e-6 / 148 / 1.965);
    cudaFree(d);
}

int main()
{
    run<0>("IADD3", 64); run<1>("IMAD", 64); run<2>("dp2a", 64); run<3>("dp4a", 64); run<4>("min3+xor", 128);
    run<5>("LOP3", 64); run<6>("IADD3+IMAD", 128); run<7>("SHF", 64); run<8>("PRMT", 64); run<9>("POPC+add", 128);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// probe_pipes.cu — how fast does a B200 issue DFMA, IMAD.WIDE.U32 and a mix of both?  (tools/; not part of the library)
// Answers whether the FP64 pipe could take over part of the field multiplication's partial products.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_pipes tools/probe_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: DFMA only, 1: IMAD.WIDE only, 2: both interleaved 1:1, 3: both 2 IMAD.WIDE : 1 DFMA
__global__ void __launch_bounds__(256) k_probe(int iters, double* outd, unsigned long long* outi, double seed, unsigned int iseed)
{
    double d0 = seed + threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3, d4 = d0 + 4, d5 = d0 + 5, d6 = d0 + 6, d7 = d0 + 7;
    const double m = 1.0000001, c = 1e-9;
    unsigned long long a0 = iseed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const unsigned int x = iseed | 1u, y = (iseed * 2654435761u) | 1u;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (MODE == 0 || MODE == 2 || MODE == 3) {
                d0 = fma(d0, m, c); d1 = fma(d1, m, c); d2 = fma(d2, m, c); d3 = fma(d3, m, c);
                if (MODE != 3) { d4 = fma(d4, m, c); d5 = fma(d5, m, c); d6 = fma(d6, m, c); d7 = fma(d7, m, c); }
            }
            if (MODE == 1 || MODE == 2 || MODE == 3) {
                // eight independent chains a <- lo32(a) * x + a: an IMAD.WIDE.U32 whose operand changes every iteration
                a0 += (unsigned long long)(unsigned int)a0 * x; a1 += (unsigned long long)(unsigned int)a1 * y;
                a2 += (unsigned long long)(unsigned int)a2 * x; a3 += (unsigned long long)(unsigned int)a3 * y;
                a4 += (unsigned long long)(unsigned int)a4 * x; a5 += (unsigned long long)(unsigned int)a5 * y;
                a6 += (unsigned long long)(unsigned int)a6 * x; a7 += (unsigned long long)(unsigned int)a7 * y;
            }
        }
    }
    outd[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    outi[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

template <int MODE>
static void run(const char* name, int sms, double dfma_per_iter, double imad_per_iter)
{
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double* od; unsigned long long* oi;
    cudaMalloc(&od, sizeof(double) * blocks * threads);
    cudaMalloc(&oi, sizeof(unsigned long long) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<MODE><<<blocks, threads>>>(100, od, oi, 1.5, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_probe<MODE><<<blocks, threads>>>(iters, od, oi, 1.5, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * threads * iters;
    printf("%-28s %8.3f ms   DFMA %7.2f T/s   IMAD.WIDE %7.2f T/s   (%s)\n", name, ms, n * dfma_per_iter / ms / 1e9, n * imad_per_iter / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(od); cudaFree(oi);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<0>("DFMA only", p.multiProcessorCount, 64, 0);
    run<1>("IMAD.WIDE only", p.multiProcessorCount, 0, 64);
    run<2>("DFMA + IMAD.WIDE 1:1", p.multiProcessorCount, 64, 64);
    run<3>("DFMA + IMAD.WIDE 1:2", p.multiProcessorCount, 32, 64);
    return 0;
}

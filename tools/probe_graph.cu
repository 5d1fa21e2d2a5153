// probe_graph.cu — gap between dependent kernel launches on a B200: plain stream launches against a CUDA graph of the same chain
// (tools/; not part of the library).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_graph tools/probe_graph.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_spin(long long cycles, int* sink)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && threadIdx.x == 1025) *sink = 1;
}
int main()
{
    const int N = 200;
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (long long us : {0LL, 5LL, 13LL}) {
        const long long cyc = us * 1965;
        for (int blocks : {1, 148, 592}) {
            // stream launches
            for (int w = 0; w < 2; w++) {
                cudaEventRecord(e0, st);
                for (int i = 0; i < N; i++) k_spin<<<blocks, 128, 0, st>>>(cyc, nullptr);
                cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            }
            float ms_s; cudaEventElapsedTime(&ms_s, e0, e1);
            // graph of the same chain
            cudaGraph_t g; cudaGraphExec_t ge;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            for (int i = 0; i < N; i++) k_spin<<<blocks, 128, 0, st>>>(cyc, nullptr);
            cudaStreamEndCapture(st, &g);
            cudaGraphInstantiate(&ge, g, 0);
            float ms_g = 0;
            for (int w = 0; w < 2; w++) {
                cudaEventRecord(e0, st);
                cudaGraphLaunch(ge, st);
                cudaEventRecord(e1, st); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms_g, e0, e1);
            }
            printf("kernel %2lld us x %d blocks: stream %.2f us/launch, graph %.2f us/launch\n", us, blocks, ms_s * 1e3 / N, ms_g * 1e3 / N);
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        }
    }
    return 0;
}

// experiment: event-to-event time of trivial kernels on this GPU (the floor under every per-kernel time in bench.py)
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>
struct Big { unsigned v[768]; };           // 3 KB of parameters
__global__ void k_empty() {}
__global__ void k_param(const __grid_constant__ Big b, unsigned* out) { if (b.v[threadIdx.x] == 0xdeadbeef) out[0] = 1; }
__global__ void k_smem(unsigned* out) { __shared__ unsigned s[1024]; s[threadIdx.x] = threadIdx.x; __syncthreads(); if (s[(threadIdx.x + 1) & 127] == 0xdeadbeef) out[0] = 1; }
__global__ void k_spin(long long cycles, unsigned* out) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} if (cycles < 0) out[0] = 1; }
template <class F> float timeit(cudaStream_t st, F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ts;
    for (int i = 0; i < 30; i++) { cudaStreamSynchronize(st); cudaEventRecord(e0, st); f(); cudaEventRecord(e1, st); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms * 1e3f); }
    std::sort(ts.begin(), ts.end()); return ts[ts.size() / 2];
}
int main() {
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    unsigned* out; cudaMalloc(&out, 4); Big b{}; 
    printf("no kernel (two events)      %.2f us\n", timeit(st, [&] {}));
    printf("empty <<<1,128>>>           %.2f us\n", timeit(st, [&] { k_empty<<<1, 128, 0, st>>>(); }));
    printf("empty <<<992,128>>>         %.2f us\n", timeit(st, [&] { k_empty<<<992, 128, 0, st>>>(); }));
    printf("3 KB params <<<1,128>>>     %.2f us\n", timeit(st, [&] { k_param<<<1, 128, 0, st>>>(b, out); }));
    printf("3 KB params <<<992,128>>>   %.2f us\n", timeit(st, [&] { k_param<<<992, 128, 0, st>>>(b, out); }));
    printf("smem <<<992,128>>>          %.2f us\n", timeit(st, [&] { k_smem<<<992, 128, 0, st>>>(out); }));
    for (long long c : {2000LL, 10000LL, 20000LL, 40000LL}) printf("spin %6lld cycles <<<148,128>>> %.2f us\n", c, timeit(st, [&] { k_spin<<<148, 128, 0, st>>>(c, out); }));
    printf("two empty kernels           %.2f us\n", timeit(st, [&] { k_empty<<<1, 128, 0, st>>>(); k_empty<<<1, 128, 0, st>>>(); }));
    // the same two events and kernel behind a kernel that is still running (the commands wait in the queue: no host latency between them)
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (long long c : {0LL, 20000LL}) {
            std::vector<float> ts;
            for (int i = 0; i < 30; i++) {
                cudaStreamSynchronize(st);
                k_spin<<<1, 32, 0, st>>>(40000, out);
                cudaEventRecord(e0, st); k_spin<<<148, 128, 0, st>>>(c, out); cudaEventRecord(e1, st); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms * 1e3f);
            }
            std::sort(ts.begin(), ts.end());
            printf("behind a running kernel: spin %6lld cycles <<<148,128>>> %.2f us\n", c, ts[ts.size() / 2]);
        }
        // one graph: event record node, kernel node, event record node
        for (long long c : {0LL, 20000LL}) {
            cudaGraph_t g; cudaGraphExec_t ge;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            cudaEventRecordWithFlags(e0, st, cudaEventRecordExternal);
            k_spin<<<148, 128, 0, st>>>(c, out);
            cudaEventRecordWithFlags(e1, st, cudaEventRecordExternal);
            cudaStreamEndCapture(st, &g);
            cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
            if (e != cudaSuccess) { printf("graph instantiate: %s\n", cudaGetErrorString(e)); break; }
            std::vector<float> ts;
            for (int i = 0; i < 30; i++) {
                cudaStreamSynchronize(st);
                cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
                float ms = -1; e = cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms * 1e3f);
            }
            std::sort(ts.begin(), ts.end());
            printf("one graph (event, kernel, event): spin %6lld cycles <<<148,128>>> %.2f us (%s)\n", c, ts[ts.size() / 2], cudaGetErrorString(e));
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        }
    }
    return 0;
}

// dependent-issue latencies on this GPU (cycles per op, one warp): DADD, DMUL, DFMA, SHFL(64-bit), LDS, ballot
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = lane; sm[lane + 32] = 0; __syncwarp();
    double x = a + lane;
    long long t0, t1;
    const int N = 2048;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dadd_rn(x, b);
    t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dmul_rn(x, b);
    t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __fma_rn(x, b, a);
    t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int)sm[idx & 31] ;
    t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
    unsigned m = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) m = __ballot_sync(0xffffffffu, (m >> (i & 7)) & 1) + lane;
    t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
    // two independent DADD chains (ILP 2)
    double y = b + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { x = __dadd_rn(x, b); y = __dmul_rn(y, a); }
    t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
    // DSETP + select chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = (x < y) ? __dadd_rn(x, b) : y;
    t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
    out[lane] = x + idx + m + y;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 256); cudaMalloc(&c, 64);
    for (int r = 0; r < 2; r++) k<<<1, 32>>>(o, c, 1.0000001, 0.9999999);
    long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    const char* n[8] = {"DADD", "DMUL", "DFMA", "SHFL64", "LDS+cvt", "BALLOT", "DADD||DMUL pair", "DSETP+sel+DADD"};
    for (int i = 0; i < 8; i++) printf("%-16s %.2f cycles/op\n", n[i], h[i] / 2048.0);
    return 0;
}

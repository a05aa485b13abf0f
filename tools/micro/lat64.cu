// dependent-issue latency of FP64 ops, shared-memory loads and shuffles for one warp (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fma(x, b, a);
    }
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = x + b;
    }
    long long t2 = clock64();
    int idx = threadIdx.x & 31;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) idx = (int)sm[idx & 63] + (idx & 1);
    }
    long long t3 = clock64();
    double y = x;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) y = __shfl_xor_sync(0xffffffffu, y, 1) + b;
    }
    long long t4 = clock64();
    float f = (float)a;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) f = fmaf(f, (float)b, (float)a);
    }
    long long t5 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    }
    out[threadIdx.x] = x + idx + y + f;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8 * 1024); cudaMalloc(&c, 8 * 8);
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(d, c, 0.0, 1.0000001);
        long long h[5]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %4d: DFMA %.1f  DADD %.1f  LDS+I2F chain %.1f  SHFL+DADD %.1f  FFMA %.1f cycles per dependent op\n", threads,
               h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0, h[4] / 4096.0);
    }
    return 0;
}

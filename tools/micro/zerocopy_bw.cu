// Micro-benchmark: PCIe bandwidth of SM-initiated zero-copy reads/writes of pinned host memory
// against the copy engines (decides whether a sparse host-array transfer can pay off).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zerocopy_bw zerocopy_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_read(const double* __restrict__ h, double* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) d[i] = h[i];
}
__global__ void k_write(double* __restrict__ h, const double* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) h[i] = d[i];
}

int main() {
    const size_t n = (size_t)88 << 20;   // 704 MB
    double *h, *h2, *d, *d2;
    cudaMallocHost(&h, n * 8); cudaMallocHost(&h2, n * 8);
    cudaMalloc(&d, n * 8); cudaMalloc(&d2, n * 8);
    for (size_t i = 0; i < n; i += 512) h[i] = 1.0;
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    auto report = [&](const char* name, double bytes) { cudaEventElapsedTime(&ms, e0, e1); printf("%-44s %7.2f ms  %6.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0, s1); cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice, s1); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        if (rep) report("copy engine H2D", n * 8.0);
        cudaEventRecord(e0, s1); cudaMemcpyAsync(h2, d2, n * 8, cudaMemcpyDeviceToHost, s1); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        if (rep) report("copy engine D2H", n * 8.0);
        for (int blocks : {16, 64, 296}) {
            cudaEventRecord(e0, s1); k_read<<<blocks, 512, 0, s1>>>(h, d, n); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
            if (rep) { char b[64]; snprintf(b, 64, "SM zero-copy read  (%d CTAs x 512)", blocks); report(b, n * 8.0); }
            cudaEventRecord(e0, s1); k_write<<<blocks, 512, 0, s1>>>(h2, d2, n); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
            if (rep) { char b[64]; snprintf(b, 64, "SM zero-copy write (%d CTAs x 512)", blocks); report(b, n * 8.0); }
        }
        // both directions at once
        cudaEventRecord(e0, s1);
        cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h2, d2, n * 8, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s2); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        if (rep) report("copy engines, both directions (per dir)", n * 8.0);
        cudaEventRecord(e0, s1);
        k_read<<<64, 512, 0, s1>>>(h, d, n); k_write<<<64, 512, 0, s2>>>(h2, d2, n);
        cudaStreamSynchronize(s2); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        if (rep) report("SM zero-copy, both directions (per dir)", n * 8.0);
    }
    return 0;
}

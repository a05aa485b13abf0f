// scan.cuh -- three-pass exclusive prefix scan (int32 counts -> int64 offsets) and
// flag compaction used by the grid build ("prefix-scan compaction" of the node lists
// and the CSR row offsets, replacing the serial loop at src/grid.cpp:230-232).
#pragma once
#include "common.cuh"

namespace pdscan {

constexpr int kBlock = 512;
constexpr int kItems = 8;
constexpr int kTile = kBlock * kItems;

__device__ __forceinline__ long long block_exclusive(long long v, long long* total, long long* sh) {
    // exclusive scan of one value per thread across a block of kBlock threads
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    if (wid == 0) {
        long long s = (lane < kBlock / 32) ? sh[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < kBlock / 32) sh[lane] = s;   // inclusive warp totals
    }
    __syncthreads();
    long long warp_off = wid ? sh[wid - 1] : 0;
    *total = sh[kBlock / 32 - 1];
    long long r = warp_off + x - v;
    __syncthreads();
    return r;
}

static __global__ void k_tile_sums(const int* __restrict__ in, long long n, long long* __restrict__ sums) {
    __shared__ long long sh[kBlock / 32];
    long long base = (long long)blockIdx.x * kTile;
    long long s = 0;
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        long long i = base + (long long)threadIdx.x * kItems + t;
        if (i < n) s += in[i];
    }
    long long total;
    block_exclusive(s, &total, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

static __global__ void k_scan_sums(long long* sums, long long nb, long long* grand_total) {
    __shared__ long long sh[kBlock / 32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nb; base += kBlock) {
        long long i = base + threadIdx.x;
        long long v = (i < nb) ? sums[i] : 0;
        long long total;
        long long ex = block_exclusive(v, &total, sh);
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

static __global__ void k_tile_scan(const int* __restrict__ in, long long n, const long long* __restrict__ sums,
                            long long* __restrict__ out) {
    __shared__ long long sh[kBlock / 32];
    long long base = (long long)blockIdx.x * kTile;
    int v[kItems];
    long long s = 0;
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        long long i = base + (long long)threadIdx.x * kItems + t;
        v[t] = (i < n) ? in[i] : 0;
        s += v[t];
    }
    long long total;
    long long ex = block_exclusive(s, &total, sh) + sums[blockIdx.x];
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        long long i = base + (long long)threadIdx.x * kItems + t;
        if (i < n) out[i] = ex;
        ex += v[t];
    }
}

// out[i] = sum_{t<i} in[t] for i in [0,n]; out has n+1 entries. Returns grand total.
inline int exclusive_scan(pdgpu_ctx* c, const int* d_in, long long n, long long* d_out, long long* total) {
    long long nb = (n + kTile - 1) / kTile;
    if (nb < 1) nb = 1;
    long long* d_sums = nullptr;
    CUDA_OK(cudaMalloc(&d_sums, sizeof(long long) * (nb + 1)));
    LAUNCH(c, k_tile_sums, (unsigned)nb, kBlock, 0, d_in, n, d_sums);
    LAUNCH(c, k_scan_sums, 1, kBlock, 0, d_sums, nb, d_sums + nb);
    LAUNCH(c, k_tile_scan, (unsigned)nb, kBlock, 0, d_in, n, d_sums, d_out);
    long long tot = 0;
    CUDA_OK(cudaMemcpyAsync(&tot, d_sums + nb, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(d_out + n, d_sums + nb, sizeof(long long), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_sums));
    if (total) *total = tot;
    return 0;
}

}  // namespace pdscan

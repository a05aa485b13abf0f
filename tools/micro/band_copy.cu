// PCIe copies of a 157 x 157 x Nz lattice of doubles: one linear copy per chunk of planes vs. B strided 3D
// copies per chunk that skip the corners outside the tube cross-section (tools for DESIGN 5.4, e2e path).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
    const int Nx = 157, Ny = 157, Nz = 707; const int chunks = argc > 1 ? atoi(argv[1]) : 8;
    const int comps_list[2] = {1, 3};
    const size_t N = (size_t)Nx * Ny * Nz;
    double *h_in, *h_out, *d_a, *d_b;
    CK(cudaMallocHost(&h_in, N * 5 * 8)); CK(cudaMallocHost(&h_out, N * 5 * 8));
    CK(cudaMalloc(&d_a, N * 5 * 8)); CK(cudaMalloc(&d_b, N * 5 * 8));
    for (size_t i = 0; i < N * 5; ++i) h_in[i] = (double)i;
    cudaStream_t up, down; CK(cudaStreamCreate(&up)); CK(cudaStreamCreate(&down));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int B : {0, 16}) {
        // bands of the disc of radius Nx/2 - 0.5
        struct Band { int j0, j1, x0, x1; };
        std::vector<Band> bands;
        double moved = 0;
        if (B == 0) { bands.push_back({0, Ny, 0, Nx}); }
        else {
            const double c = (Nx - 1) / 2.0, R = 78.0;
            for (int b = 0; b < B; ++b) {
                int j0 = b * Ny / B, j1 = (b + 1) * Ny / B, x0 = Nx, x1 = 0;
                for (int j = j0; j < j1; ++j) {
                    double dy = j - c; double h2 = R * R - dy * dy;
                    if (h2 < 0) continue;
                    int lo = (int)ceil(c - sqrt(h2)), hi = (int)floor(c + sqrt(h2));
                    x0 = lo < x0 ? lo : x0; x1 = hi + 1 > x1 ? hi + 1 : x1;
                }
                if (x1 > x0) bands.push_back({j0, j1, x0, x1});
            }
        }
        for (auto& b : bands) moved += (double)(b.j1 - b.j0) * (b.x1 - b.x0);
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, up));
            CK(cudaStreamWaitEvent(down, e0, 0));
            for (int k = 0; k < chunks; ++k) {
                int z0 = k * Nz / chunks, z1 = (k + 1) * Nz / chunks;
                for (int ci = 0; ci < 2; ++ci) {
                    const int comps = comps_list[ci];
                    for (int pass = 0; pass < (comps == 1 ? 2 : 1); ++pass) {   // rho, C: two scalar arrays; vel: one AoS array
                        const size_t base = (comps == 1 ? (size_t)pass * N : 2 * N);   // element offset of the array
                        for (auto& b : bands) {
                            if (B == 0) {
                                size_t off = base + (size_t)z0 * Nx * Ny * comps;
                                size_t bytes = (size_t)(z1 - z0) * Nx * Ny * comps * 8;
                                CK(cudaMemcpyAsync(d_a + off, h_in + off, bytes, cudaMemcpyHostToDevice, up));
                                CK(cudaMemcpyAsync(h_out + off, d_b + off, bytes, cudaMemcpyDeviceToHost, down));
                            } else {
                                size_t off = base + (((size_t)z0 * Ny + b.j0) * Nx + b.x0) * comps;
                                cudaMemcpy3DParms p = {};
                                p.srcPtr = make_cudaPitchedPtr(h_in + off, (size_t)Nx * comps * 8, Nx, Ny);
                                p.dstPtr = make_cudaPitchedPtr(d_a + off, (size_t)Nx * comps * 8, Nx, Ny);
                                p.extent = make_cudaExtent((size_t)(b.x1 - b.x0) * comps * 8, b.j1 - b.j0, z1 - z0);
                                p.kind = cudaMemcpyHostToDevice;
                                CK(cudaMemcpy3DAsync(&p, up));
                                cudaMemcpy3DParms q = {};
                                q.srcPtr = make_cudaPitchedPtr(d_b + off, (size_t)Nx * comps * 8, Nx, Ny);
                                q.dstPtr = make_cudaPitchedPtr(h_out + off, (size_t)Nx * comps * 8, Nx, Ny);
                                q.extent = p.extent;
                                q.kind = cudaMemcpyDeviceToHost;
                                CK(cudaMemcpy3DAsync(&q, down));
                            }
                        }
                    }
                }
            }
            CK(cudaEventRecord(e1, up));
            CK(cudaStreamSynchronize(down));
            CK(cudaStreamSynchronize(up));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 1)
                printf("bands %2d: %5.1f %% of the lattice, up+down of 5 doubles per node: %.2f ms (up stream), %.1f GB/s per direction\n", B,
                       100.0 * moved / (Nx * Ny), ms, moved * Nz * 5 * 8 / (ms * 1e6));
        }
    }
    return 0;
}

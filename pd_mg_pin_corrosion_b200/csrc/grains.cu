// grains.cu -- device side of GrainStructure::generate (src/grains.cpp:9-179).
//
// The reference assigns every SOLID_MG node to its nearest Voronoi seed by brute force
// (O(N_solid * n_grains), src/grains.cpp:55-70), marks grain-boundary nodes (a solid immediate
// neighbour of another grain, :75-89), dilates that set gb_width_cells times (:92-107) and, for
// clustered precipitates, grows every precipitate seed to a ball (:152-166).  Those lattice passes
// run here; the random draws (seed picks, the shuffle of the interior nodes) stay on the host
// because they go through libstdc++'s mt19937 / uniform_int_distribution / shuffle, whose call
// sequence the reference fixes (host/grains.cpp keeps that sequence and calls in here).
//
// Bit-exactness: node positions are fma(idx, dx, origin) and distances sqrt(fma-chain) exactly as
// the reference's Release build evaluates them (same pattern as host/grains.cpp and geom.cuh), and
// nearest-seed ties keep the first seed like the reference's strict `<`.
#include <algorithm>

#include "common.cuh"
#include "geom.cuh"

namespace {

struct GrainLat {
    Lat L;
    int dim, Nz;
    long long NL;
    double dx, ox, oy, oz;
};

__device__ __forceinline__ void node_pos(const GrainLat& g, long long l, double p[3], int* i, int* j, int* k) {
    int a;
    local_to_ijk(g.L, l, g.dim, i, j, k, &a);
    p[0] = geom_coord(g.ox, *i, g.dx);
    p[1] = geom_coord(g.oy, *j, g.dx);
    p[2] = g.dim == 3 ? geom_coord(g.oz, *k, g.dx) : 0.0;
}
__device__ __forceinline__ double dist3(const GrainLat& g, const double a[3], const double b[3]) {
    double s = 0.0;   // norm(a - b) of src/utils.h:16-24 as compiled: one fma per component, then sqrt
    for (int d = 0; d < g.dim; ++d) {
        const double t = a[d] - b[d];
        s = fma(t, t, s);
    }
    return sqrt(s);
}

// Voronoi assignment (src/grains.cpp:55-70) of every local node, ghost planes included (grain ids
// are a function of the position only; the neighbour passes below need them one plane out)
__global__ void __launch_bounds__(256)
k_voronoi(GrainLat g, const uint8_t* __restrict__ type, const double* __restrict__ seeds, int n_grains,
          int* __restrict__ gid) {
    extern __shared__ double s_seed[];   // tile of seeds
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool solid = l < g.NL && type[l] == PDGPU_SOLID_MG;
    double p[3] = {0.0, 0.0, 0.0};
    int i, j, k;
    if (solid) node_pos(g, l, p, &i, &j, &k);
    double best = 1.7976931348623157e308;
    int best_g = 0;
    for (int g0 = 0; g0 < n_grains; g0 += 256) {
        const int n = min(256, n_grains - g0);
        __syncthreads();
        for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) s_seed[t] = seeds[3 * (long long)g0 + t];
        __syncthreads();
        if (solid)
            for (int q = 0; q < n; ++q) {
                const double d = dist3(g, p, &s_seed[3 * q]);
                if (d < best) { best = d; best_g = g0 + q; }
            }
    }
    if (l < g.NL) gid[l] = solid ? best_g : -1;
}

// immediate neighbours that exist in the reference's CSR: in the box and not OUTSIDE
template <class F>
__device__ __forceinline__ void for_immediate(const GrainLat& g, long long l, const uint8_t* __restrict__ type, F&& f) {
    int i, j, k, a;
    local_to_ijk(g.L, l, g.dim, &i, &j, &k, &a);
    const int al = (int)(l / g.L.P);                 // local axial plane
    const int nlp = (int)(g.NL / g.L.P);
    const int klo = g.dim == 3 ? -1 : 0, khi = g.dim == 3 ? 1 : 0;
    for (int dk = klo; dk <= khi; ++dk)
        for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di) {
                if (!di && !dj && !dk) continue;
                const int ni = i + di;
                if (ni < 0 || ni >= g.L.Nx) continue;
                long long nn;
                if (g.dim == 3) {
                    const int nj = j + dj, nk = k + dk;
                    if (nj < 0 || nj >= g.L.Ny || nk < 0 || nk >= g.Nz) continue;
                    if (al + dk < 0 || al + dk >= nlp) continue;     // outside this slab's ghost planes
                    nn = l + (long long)dk * g.L.P + (long long)dj * g.L.Nx + di;
                } else {
                    const int nj = j + dj;                           // axial index in 2D
                    if (nj < 0 || nj >= g.L.Na) continue;
                    if (al + dj < 0 || al + dj >= nlp) continue;
                    nn = l + (long long)dj * g.L.P + di;
                }
                if (type[nn] == PDGPU_OUTSIDE) continue;
                if (f(nn)) return;
            }
}

__global__ void __launch_bounds__(256)
k_gb_detect(GrainLat g, const uint8_t* __restrict__ type, const int* __restrict__ gid, uint8_t* __restrict__ gb) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= g.NL) return;
    uint8_t r = 0;
    if (type[l] == PDGPU_SOLID_MG) {
        const int gi = gid[l];
        for_immediate(g, l, type, [&](long long nn) {
            if (type[nn] == PDGPU_SOLID_MG && gid[nn] != gi) { r = 1; return true; }
            return false;
        });
    }
    gb[l] = r;
}

__global__ void __launch_bounds__(256)
k_gb_dilate(GrainLat g, const uint8_t* __restrict__ type, const uint8_t* __restrict__ cur, uint8_t* __restrict__ next) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= g.NL) return;
    uint8_t r = cur[l];
    if (!r && type[l] == PDGPU_SOLID_MG)
        for_immediate(g, l, type, [&](long long nn) {
            if (cur[nn]) { r = 1; return true; }
            return false;
        });
    next[l] = r;
}

// clustered precipitates (src/grains.cpp:152-166): a solid, non-boundary, non-seed node joins when a
// seed lies within cluster_r.  The reference scans all seeds; the same predicate is evaluated here for
// the lattice nodes of the enclosing cube (a seed further away than cells+1 lattice steps along any
// axis cannot satisfy it).
__global__ void __launch_bounds__(256)
k_precip_grow(GrainLat g, const uint8_t* __restrict__ type, const uint8_t* __restrict__ gb,
              const uint8_t* __restrict__ seed, int cells, double cluster_r, uint8_t* __restrict__ out) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= g.NL) return;
    uint8_t r = seed[l];
    if (!r && type[l] == PDGPU_SOLID_MG && !gb[l]) {
        double p[3], q[3];
        int i, j, k;
        node_pos(g, l, p, &i, &j, &k);
        const int al = (int)(l / g.L.P), nlp = (int)(g.NL / g.L.P);
        const int c = cells + 1;
        const int klo = g.dim == 3 ? -c : 0, khi = g.dim == 3 ? c : 0;
        for (int dk = klo; dk <= khi && !r; ++dk)
            for (int dj = -c; dj <= c && !r; ++dj)
                for (int di = -c; di <= c && !r; ++di) {
                    const int ni = i + di;
                    if (ni < 0 || ni >= g.L.Nx) continue;
                    const int dax = g.dim == 3 ? dk : dj;
                    if (al + dax < 0 || al + dax >= nlp) continue;
                    long long nn;
                    if (g.dim == 3) {
                        const int nj = j + dj;
                        if (nj < 0 || nj >= g.L.Ny) continue;
                        nn = l + (long long)dk * g.L.P + (long long)dj * g.L.Nx + di;
                    } else {
                        nn = l + (long long)dj * g.L.P + di;
                    }
                    if (!seed[nn]) continue;
                    int qi, qj, qk;
                    node_pos(g, nn, q, &qi, &qj, &qk);
                    if (dist3(g, p, q) <= cluster_r) r = 1;
                }
    }
    out[l] = r;
}

__global__ void k_add_int(int* a, long long n, int s) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] += s;
}

GrainLat grain_lat(const pdgpu_ctx* c) {
    GrainLat g;
    g.L = make_lat(c);
    g.dim = c->dim;
    g.Nz = c->dim == 3 ? c->Nz : 1;
    g.NL = c->NL;
    g.dx = c->cfg.dx;
    g.ox = c->origin[0]; g.oy = c->origin[1]; g.oz = c->origin[2];
    return g;
}

// owned part of a local device array -> global host array; slab contexts merge over the ranks
int to_global(pdgpu_ctx* c, const void* d_local, int elem, void* host_global) {
    const long long n_own = c->own_hi - c->own_lo, goff = (long long)c->a0 * c->P;
    const size_t total = (size_t)c->N_total * elem;
    if (c->nranks == 1 || !c->comm) {
        CUDA_OK(cudaMemcpyAsync((char*)host_global + goff * elem, (const char*)d_local + c->own_lo * elem,
                                (size_t)n_own * elem, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    char* gbuf = nullptr;
    CUDA_OK(cudaMalloc(&gbuf, total));
    CUDA_OK(cudaMemsetAsync(gbuf, 0, total, c->stream));
    CUDA_OK(cudaMemcpyAsync(gbuf + goff * elem, (const char*)d_local + c->own_lo * elem, (size_t)n_own * elem,
                            cudaMemcpyDeviceToDevice, c->stream));
    int rc = pd_comm_allreduce_bytes(c, gbuf, total, elem);
    if (!rc && cudaMemcpyAsync(host_global, gbuf, total, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = 1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = 1;
    cudaFree(gbuf);
    return rc;
}

// global host flags -> local device array (owned planes + in-domain ghost planes)
int from_global(pdgpu_ctx* c, const uint8_t* host_global, uint8_t* d_local) {
    CUDA_OK(cudaMemsetAsync(d_local, 0, c->NL, c->stream));
    const int ga = std::max(c->a0 - c->R, 0), gb = std::min(c->a1 + c->R, c->Na);
    const long long loff = (long long)(ga - (c->a0 - c->R)) * c->P;
    CUDA_OK(cudaMemcpyAsync(d_local + loff, host_global + (long long)ga * c->P, (size_t)(gb - ga) * c->P,
                            cudaMemcpyHostToDevice, c->stream));
    return 0;
}

}  // namespace

// Voronoi assignment + grain-boundary detection + dilation.  seeds_xyz [n_grains][3] are the seed
// positions the host drew; grain_id_global [N_total] (int32, -1 off the wire) and is_gb_global
// [N_total] receive the WHOLE arrays on every rank.
extern "C" int pdgpu_grains_voronoi(pdgpu_ctx* c, const double* seeds_xyz, int n_grains, int gb_width_cells,
                                    int* grain_id_global, uint8_t* is_gb_global) {
    NEED_GRID(c);
    if (!seeds_xyz || n_grains < 1 || !grain_id_global || !is_gb_global || gb_width_cells < 0)
        PD_FAIL("pdgpu_grains_voronoi: bad arguments");
    if (c->nranks > 1 && 1 + gb_width_cells > c->R)
        PD_FAIL("pdgpu_grains_voronoi: gb_width_cells = %d needs more than the %d ghost planes of a slab context",
                gb_width_cells, c->R);
    GrainLat g = grain_lat(c);
    double* d_seeds = nullptr;
    int* d_gid = nullptr;
    uint8_t *d_a = nullptr, *d_b = nullptr;
    CUDA_OK(cudaMalloc(&d_seeds, sizeof(double) * 3 * (size_t)n_grains));
    CUDA_OK(cudaMalloc(&d_gid, sizeof(int) * c->NL));
    CUDA_OK(cudaMalloc(&d_a, c->NL));
    CUDA_OK(cudaMalloc(&d_b, c->NL));
    CUDA_OK(cudaMemcpyAsync(d_seeds, seeds_xyz, sizeof(double) * 3 * (size_t)n_grains, cudaMemcpyHostToDevice, c->stream));
    const unsigned nb = nblocks(c->NL, 256);
    LAUNCH(c, k_voronoi, nb, 256, sizeof(double) * 3 * 256, g, c->type, d_seeds, n_grains, d_gid);
    LAUNCH(c, k_gb_detect, nb, 256, 0, g, c->type, d_gid, d_a);
    for (int pass = 0; pass < gb_width_cells; ++pass) {
        LAUNCH(c, k_gb_dilate, nb, 256, 0, g, c->type, d_a, d_b);
        std::swap(d_a, d_b);
    }
    std::fill(grain_id_global, grain_id_global + c->N_total, -1);
    // merged over ranks as integers: -1 would not survive a sum, so ids travel as id + 1
    int rc = 0;
    if (c->nranks > 1 && c->comm) {
        LAUNCH(c, k_add_int, nb, 256, 0, d_gid, c->NL, 1);   // gid + 1 (0 = not on the wire)
        rc = to_global(c, d_gid, 4, grain_id_global);
        if (!rc) for (long long n = 0; n < c->N_total; ++n) grain_id_global[n] -= 1;
    } else {
        rc = to_global(c, d_gid, 4, grain_id_global);
    }
    std::fill(is_gb_global, is_gb_global + c->N_total, (uint8_t)0);
    if (!rc) rc = to_global(c, d_a, 1, is_gb_global);
    cudaFree(d_seeds); cudaFree(d_gid); cudaFree(d_a); cudaFree(d_b);
    if (rc) PD_FAIL("pdgpu_grains_voronoi failed: %s", pdgpu_last_error());
    return 0;
}

// Cluster growth of the precipitate seeds (src/grains.cpp:152-166).  seed_flags_global marks the seeds
// the host drew (already is_precipitate = 1); is_precip_global receives seeds + grown nodes (whole array).
extern "C" int pdgpu_grains_grow_precip(pdgpu_ctx* c, const uint8_t* is_gb_global, const uint8_t* seed_flags_global,
                                        int cluster_cells, uint8_t* is_precip_global) {
    NEED_GRID(c);
    if (!is_gb_global || !seed_flags_global || !is_precip_global || cluster_cells < 0)
        PD_FAIL("pdgpu_grains_grow_precip: bad arguments");
    if (c->nranks > 1 && cluster_cells + 1 > c->R)
        PD_FAIL("pdgpu_grains_grow_precip: cluster of %d cells needs more than the %d ghost planes of a slab context",
                cluster_cells, c->R);
    GrainLat g = grain_lat(c);
    uint8_t *d_gb = nullptr, *d_seed = nullptr, *d_out = nullptr;
    CUDA_OK(cudaMalloc(&d_gb, c->NL));
    CUDA_OK(cudaMalloc(&d_seed, c->NL));
    CUDA_OK(cudaMalloc(&d_out, c->NL));
    PD_TRY(from_global(c, is_gb_global, d_gb));
    PD_TRY(from_global(c, seed_flags_global, d_seed));
    LAUNCH(c, k_precip_grow, nblocks(c->NL, 256), 256, 0, g, c->type, d_gb, d_seed, cluster_cells,
           cluster_cells * c->cfg.dx, d_out);
    std::fill(is_precip_global, is_precip_global + c->N_total, (uint8_t)0);
    int rc = to_global(c, d_out, 1, is_precip_global);
    cudaFree(d_gb); cudaFree(d_seed); cudaFree(d_out);
    if (rc) PD_FAIL("pdgpu_grains_grow_precip failed: %s", pdgpu_last_error());
    return 0;
}

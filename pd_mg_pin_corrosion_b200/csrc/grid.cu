// grid.cu -- Grid::build / Grid::build_neighbors on device (reference src/grid.cpp:29-294)
// plus the static tables the boundary operators need (wall-mirror table of
// src/boundary.cpp:143-264, outlet Gauss-Seidel wavefront schedule).
#include <algorithm>
#include <numeric>

#include "common.cuh"
#include "geom.cuh"
#include "scan.cuh"
#include "stream.cuh"

// ---------------------------------------------------------------- host-only ----
static int build_stencil_host(const PdConfig& cfg, int dim, std::vector<OffEntry>& out) {
    // src/grid.cpp:161-187 (offset enumeration, dk -> dj -> di so that neighbours are
    // index-ascending) and :274-288 (partial-volume beta).
    int m = cfg.m_ratio, mext = m + 1;
    double dx = cfg.dx, delta = cfg.delta;
    double dx_dim = 1.0;
    for (int d = 0; d < dim; ++d) dx_dim *= dx;
    out.clear();
    int klo = dim == 3 ? -mext : 0, khi = dim == 3 ? mext : 0;
    for (int dk = klo; dk <= khi; ++dk)
        for (int dj = -mext; dj <= mext; ++dj)
            for (int di = -mext; di <= mext; ++di) {
                if (di == 0 && dj == 0 && dk == 0) continue;
                double r = std::sqrt((double)(di * di + dj * dj + dk * dk)) * dx;
                if (!(r <= delta + 0.5 * dx)) continue;
                OffEntry e;
                e.di = di; e.dj = dj; e.dk = dk; e.pad = 0; e.lin = 0;
                e.dist = r;
                e.ex = di * dx / r;
                e.ey = dj * dx / r;
                e.ez = dk * dx / r;
                double beta;
                if (r <= delta - 0.5 * dx) beta = 1.0;
                else if (r <= delta + 0.5 * dx) beta = (delta + 0.5 * dx - r) / dx;
                else beta = 0.0;
                e.vol = beta * dx_dim;
                double inv_xi = 1.0 / r;
                e.w1 = inv_xi * e.vol;
                e.w2 = inv_xi * inv_xi * e.vol;
                out.push_back(e);
            }
    return 0;
}

extern "C" int pdgpu_grid_extents(const PdConfig* cfg, int dim, int* Nx, int* Ny, int* Nz, double origin[3]) {
    if (!cfg || (dim != 2 && dim != 3)) PD_FAIL("pdgpu_grid_extents: bad arguments");
    geom_extents(*cfg, dim, Nx, Ny, Nz, origin);
    return 0;
}

extern "C" int pdgpu_partition(int n_axial, int nranks, int rank, int* a0, int* a1) {
    if (nranks < 1 || rank < 0 || rank >= nranks || n_axial < nranks) PD_FAIL("pdgpu_partition: bad arguments");
    long long q = n_axial / nranks, r = n_axial % nranks;
    *a0 = (int)(rank * q + std::min<long long>(rank, r));
    *a1 = (int)(*a0 + q + (rank < r ? 1 : 0));
    return 0;
}

// Cost-balanced partition of the axial planes (z-slabs of a slab context).  Planes within reach of the wire cost
// more than bulk-fluid planes: the ARD bond kernel takes its general (solid / wall aware) body there and the
// solid rows and the salt pre-pass live there -- measured on 8 x B200 (profiles/r2_bench_n8_per_rank.txt): NS + ARD
// 4.32 ms on a slab inside the wire region against 4.15 ms on a fluid-only slab of the same 701 planes, i.e. +4.1 %
// per plane.  With equal plane counts the wire slabs set the step time of every rank; equal COST gives them ~3 %
// fewer planes (4 x B200, profiles/r2b_bench_n4_*: 4.825 ms per step against 4.871 ms; the residual imbalance of
// that run put the surcharge at 4.8 %).  Deterministic in (cfg, dim, nranks): every rank computes the same boundaries.
// PDGPU_SLAB_PIN_COST overrides the surcharge (0 = equal plane counts).
static double pin_surcharge() {
    static const double w = [] {
        const char* e = getenv("PDGPU_SLAB_PIN_COST");
        return e ? atof(e) : 0.048;
    }();
    return w;
}
extern "C" int pdgpu_partition_balanced(const PdConfig* cfg, int dim, int nranks, int rank, int* a0, int* a1) {
    if (!cfg || !a0 || !a1 || (dim != 2 && dim != 3)) PD_FAIL("pdgpu_partition_balanced: bad arguments");
    int Nx, Ny, Nz;
    double org[3];
    geom_extents(*cfg, dim, &Nx, &Ny, &Nz, org);
    const int Na = dim == 2 ? Ny : Nz;
    const double oa = dim == 2 ? org[1] : org[2];
    if (nranks < 1 || rank < 0 || rank >= nranks || Na < nranks) PD_FAIL("pdgpu_partition_balanced: bad arguments");
    const double w = pin_surcharge();
    if (nranks == 1 || !(w > 0.0)) return pdgpu_partition(Na, nranks, rank, a0, a1);
    const double lo = -cfg->m_ratio * cfg->dx, hi = cfg->L_wire + cfg->m_ratio * cfg->dx;
    std::vector<double> cum(Na + 1, 0.0);
    for (int k = 0; k < Na; ++k) {
        const double z = geom_coord(oa, k, cfg->dx);
        cum[k + 1] = cum[k] + 1.0 + ((z >= lo && z <= hi) ? w : 0.0);
    }
    const int min_planes = 2 * cfg->m_ratio + 2;
    std::vector<int> b(nranks + 1, 0);
    b[nranks] = Na;
    for (int r = 1; r < nranks; ++r) {
        const double target = cum[Na] * r / nranks;
        int k = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        if (k > 0 && target - cum[k - 1] < cum[k] - target) --k;      // nearest boundary
        b[r] = std::max(k, b[r - 1] + min_planes);
    }
    for (int r = nranks - 1; r >= 1; --r) b[r] = std::min(b[r], b[r + 1] - min_planes);
    for (int r = 0; r < nranks; ++r)
        if (b[r + 1] - b[r] < 1) return pdgpu_partition(Na, nranks, rank, a0, a1);   // too short: equal counts
    *a0 = b[rank]; *a1 = b[rank + 1];
    return 0;
}

// Local layout of a z-slab: everything the halo exchange needs, as plain integers.
// out[0..9] = a0, a1, local planes, NL, own_lo, own_hi, send_lo, recv_lo, send_hi, recv_hi
// (node offsets into a local array; a halo block is reach*plane nodes).
extern "C" int pdgpu_slab_layout(int n_axial, long long plane, int reach, int nranks, int rank, long long* out) {
    if (!out || plane <= 0 || reach < 0) PD_FAIL("pdgpu_slab_layout: bad arguments");
    int a0 = 0, a1 = 0;
    PD_TRY(pdgpu_partition(n_axial, nranks, rank, &a0, &a1));
    return pdgpu_slab_layout_range(a0, a1, plane, reach, out);
}
// the same for an arbitrary owned range [a0, a1) (cost-balanced slabs)
extern "C" int pdgpu_slab_layout_range(int a0, int a1, long long plane, int reach, long long* out) {
    if (!out || plane <= 0 || reach < 0 || a1 <= a0) PD_FAIL("pdgpu_slab_layout_range: bad arguments");
    long long nlp = (long long)(a1 - a0) + 2 * reach;
    long long hp = (long long)reach * plane;
    out[0] = a0; out[1] = a1; out[2] = nlp; out[3] = nlp * plane;
    out[4] = hp;                                   // own_lo
    out[5] = hp + (long long)(a1 - a0) * plane;    // own_hi
    out[6] = out[4];                               // send to rank-1: first `reach` owned planes
    out[7] = 0;                                    // recv from rank-1: low ghost planes
    out[8] = out[5] - hp;                          // send to rank+1: last `reach` owned planes
    out[9] = out[5];                               // recv from rank+1: high ghost planes
    return 0;
}

extern "C" int pdgpu_stencil(const PdConfig* cfg, int dim, int* n_off, int* off_d, double* dist, double* evec,
                             double* vol) {
    if (!cfg || (dim != 2 && dim != 3) || !n_off) PD_FAIL("pdgpu_stencil: bad arguments");
    std::vector<OffEntry> st;
    build_stencil_host(*cfg, dim, st);
    *n_off = (int)st.size();
    for (size_t o = 0; o < st.size(); ++o) {
        if (off_d) { off_d[3 * o] = st[o].di; off_d[3 * o + 1] = st[o].dj; off_d[3 * o + 2] = st[o].dk; }
        if (dist) dist[o] = st[o].dist;
        if (evec) {
            evec[dim * o] = st[o].ex; evec[dim * o + 1] = st[o].ey;
            if (dim == 3) evec[dim * o + 2] = st[o].ez;
        }
        if (vol) vol[o] = st[o].vol;
    }
    return 0;
}

// --------------------------------------------------------------- kernels -------
__global__ void k_classify(GeomParams g, Lat L, long long NL, uint8_t* __restrict__ type) {
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= NL) return;
    int i, j, k, a;
    local_to_ijk(L, l, g.dim, &i, &j, &k, &a);
    uint8_t t = PDGPU_OUTSIDE;                 // ghost planes beyond the domain act as padding
    if (a >= 0 && a < L.Na) t = geom_classify(g, i, j, k);
    type[l] = t;
}

// Row length of every owned node (CSR count pass, src/grid.cpp:194-227).
__global__ void k_rowlen(Lat L, int dim, long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                         const OffEntry* __restrict__ off, int n_off, int* __restrict__ rowlen,
                         unsigned long long* __restrict__ sums /* [0]=ns [1]=ard [2]=nnz [3]=short rows */) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long s_ns = 0, s_ard = 0, s_all = 0, s_short = 0;
    if (t < own_n) {
        long long l = own_lo + t;
        uint8_t ty = type[l];
        int cnt = 0;
        if (ty != PDGPU_OUTSIDE) {
            int q = (int)(l % L.P);
            int jj = (dim == 3) ? q / L.Nx : 0;
            int ii = q - jj * L.Nx;
            for (int o = 0; o < n_off; ++o) cnt += nbr_local(L, off[o], dim, ii, jj, l, type) >= 0;
        }
        rowlen[t] = cnt;
        s_all = cnt;
        if (ty == PDGPU_FLUID) { s_ns = cnt; s_ard = cnt; s_short = (cnt != n_off); }
        if (ty == PDGPU_SOLID_MG) { s_ard = cnt; s_short = (cnt != n_off); }
    }
    // warp-aggregate, then one atomic per warp (integers: order-free)
    for (int o = 16; o > 0; o >>= 1) {
        s_ns += __shfl_xor_sync(0xffffffffu, s_ns, o);
        s_ard += __shfl_xor_sync(0xffffffffu, s_ard, o);
        s_all += __shfl_xor_sync(0xffffffffu, s_all, o);
        s_short += __shfl_xor_sync(0xffffffffu, s_short, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (s_ns) atomicAdd(&sums[0], s_ns);
        if (s_ard) atomicAdd(&sums[1], s_ard);
        if (s_all) atomicAdd(&sums[2], s_all);
        if (s_short) atomicAdd(&sums[3], s_short);
    }
}

__global__ void k_type_flags(const uint8_t* __restrict__ type, long long own_lo, long long own_n, int want,
                             int* __restrict__ flag) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    flag[t] = (type[own_lo + t] == want);
}

// SOLID_MG nodes with at least one fluid-like neighbour (FLUID / INLET / OUTLET): the only solids whose
// salt flag / interface diffusivity anybody reads (src/pd_ard.cpp:61-73,140-162)
template <int DIM>
__global__ void k_surface_solid_flags(Lat L, const uint8_t* __restrict__ type, long long own_lo, long long own_n,
                                      const OffEntry* __restrict__ off, int n_off, int* __restrict__ flag) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    const long long l = own_lo + t;
    int f = 0;
    if (type[l] == PDGPU_SOLID_MG) {
        const int q = (int)(l % L.P);
        const int jj = (DIM == 3) ? q / L.Nx : 0;
        const int ii = q - jj * L.Nx;
        for (int o = 0; o < n_off && !f; ++o) {
            const long long nn = nbr_local(L, off[o], DIM, ii, jj, l, type);
            if (nn < 0) continue;
            const uint8_t tj = type[nn];
            f = (tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET);
        }
    }
    flag[t] = f;
}

// histogram of owned node types (block-local shared counters, one global atomic per bin)
__global__ void k_type_hist(const uint8_t* __restrict__ type, long long own_lo, long long own_n,
                            unsigned long long* __restrict__ counts) {
    __shared__ unsigned int sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < own_n;
         t += (long long)gridDim.x * blockDim.x)
        atomicAdd(&sh[type[own_lo + t] & 7], 1u);
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void k_compact(const int* __restrict__ flag, const long long* __restrict__ pos, long long own_lo,
                          long long own_n, int* __restrict__ list) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    if (flag[t]) list[pos[t]] = (int)(own_lo + t);
}

// Static wall-mirror table (src/boundary.cpp:143-264).
__global__ void k_wall_mirror(GeomParams g, Lat L, const int* __restrict__ l_wall, long long n_wall,
                              const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
                              int* __restrict__ mirror) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_wall) return;
    long long l = l_wall[t];
    int i, j, k, a;
    local_to_ijk(L, l, g.dim, &i, &j, &k, &a);
    long long m = -1;
    int im, jm;
    if (geom_wall_mirror(g, i, j, &im, &jm)) {
        if (g.dim == 2) {
            // same axial row j (always inside: j_mirror = j_grid)
            if (im >= 0 && im < L.Nx) {
                long long cand = l - i + im;
                uint8_t mt = type[cand];
                if (mt == PDGPU_FLUID || mt == PDGPU_INLET || mt == PDGPU_OUTLET || mt == PDGPU_SOLID_MG) m = cand;
            }
        } else {
            if (im >= 0 && im < L.Nx && jm >= 0 && jm < L.Ny) {
                long long cand = l + (long long)(jm - j) * L.Nx + (im - i);
                uint8_t mt = type[cand];
                if (mt == PDGPU_FLUID || mt == PDGPU_INLET || mt == PDGPU_OUTLET || mt == PDGPU_SOLID_MG) m = cand;
            }
        }
    }
    if (m < 0) {   // fallback :254-263, nearest FLUID neighbour, strict '<', CSR order
        int q = (int)(l % L.P);
        int jj = (g.dim == 3) ? q / L.Nx : 0;
        int ii = q - jj * L.Nx;
        double best = 1e30;
        for (int o = 0; o < n_off; ++o) {
            long long nn = nbr_local(L, off[o], g.dim, ii, jj, l, type);
            if (nn >= 0 && type[nn] == PDGPU_FLUID && off[o].dist < best) {
                best = off[o].dist;
                m = nn;
            }
        }
    }
    mirror[t] = (int)m;
}

// ---- ghost-plane walls of a slab (multi-GPU) ----------------------------------------------
// The owner of a WALL node knows its exact mirror; the relative offset (mirror - node) is
// position independent, so it is exchanged with the halo and ghost-plane walls can be
// refreshed locally by the pre-step wall BC exactly as their owner does.
constexpr int kNoMirror = -2147483647 - 1;
__global__ void k_moff_write(const int* __restrict__ l_wall, const int* __restrict__ mirror, long long n,
                             int* __restrict__ moff) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int l = l_wall[t], m = mirror[t];
    moff[l] = m < 0 ? kNoMirror : m - l;
}
__global__ void k_ghost_wall_flags(const uint8_t* __restrict__ type, long long NL, long long own_lo,
                                   long long own_hi, int* __restrict__ flag) {
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= NL) return;
    flag[l] = (type[l] == PDGPU_WALL) && (l < own_lo || l >= own_hi);
}
__global__ void k_ghost_mirror(const int* __restrict__ list, long long n, const int* __restrict__ moff,
                               const uint8_t* __restrict__ type, long long NL, int* __restrict__ mirror,
                               int* __restrict__ bad) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int l = list[t], off = moff[l];
    long long m = (off == kNoMirror) ? -1 : (long long)l + off;
    if (m >= NL || (off != kNoMirror && m < 0)) { atomicAdd(bad, 1); m = -1; }
    mirror[t] = (int)m;
}
__global__ void k_count_ghost_inout(const uint8_t* __restrict__ type, long long NL, long long own_lo,
                                    long long own_hi, int* __restrict__ bad) {
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= NL || (l >= own_lo && l < own_hi)) return;
    if (type[l] == PDGPU_INLET || type[l] == PDGPU_OUTLET) atomicAdd(bad, 1);
}

__global__ void k_inlet_velocity(GeomParams g, Lat L, double U_in, const int* __restrict__ l_inlet,
                                 long long n_inlet, double* __restrict__ vax) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_inlet) return;
    int i, j, k, a;
    local_to_ijk(L, l_inlet[t], g.dim, &i, &j, &k, &a);
    vax[t] = geom_inlet_velocity(g, U_in, i, j);
}

// CSR fill (src/grid.cpp:246-291): one warp per owned row, ballot-compacted writes.
__global__ void k_csr_fill(Lat L, int dim, long long own_lo, long long own_n, long long halo_shift,
                           const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
                           const long long* __restrict__ row_off, int* __restrict__ idx,
                           double* __restrict__ dist, double* __restrict__ evec, double* __restrict__ vol) {
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= own_n) return;
    long long l = own_lo + w;
    if (type[l] == PDGPU_OUTSIDE) return;
    int q = (int)(l % L.P);
    int jj = (dim == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    long long wr = row_off[w];
    for (int base = 0; base < n_off; base += 32) {
        int o = base + lane;
        long long nn = -1;
        if (o < n_off) nn = nbr_local(L, off[o], dim, ii, jj, l, type);
        unsigned mask = __ballot_sync(0xffffffffu, nn >= 0);
        if (nn >= 0) {
            long long pos = wr + __popc(mask & ((1u << lane) - 1u));
            idx[pos] = (int)(nn + halo_shift);          // local -> global node index
            dist[pos] = off[o].dist;
            evec[pos * dim] = off[o].ex;
            evec[pos * dim + 1] = off[o].ey;
            if (dim == 3) evec[pos * dim + 2] = off[o].ez;
            vol[pos] = off[o].vol;
        }
        wr += __popc(mask);
    }
}

__global__ void k_mirror_to_global(const int* __restrict__ l_wall, const int* __restrict__ mirror,
                                   long long n_wall, long long own_lo, long long halo_shift,
                                   int* __restrict__ out_own) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_wall) return;
    int m = mirror[t];
    out_own[l_wall[t] - own_lo] = m < 0 ? -1 : (int)(m + halo_shift);
}

// ---------------------------------------------------------------- host side ----
static int build_list(pdgpu_ctx* c, int want, int* d_flag, long long* d_pos, int** list, long long* n) {
    long long own_n = c->own_hi - c->own_lo;
    LAUNCH(c, k_type_flags, nblocks(own_n, 256), 256, 0, c->type, c->own_lo, own_n, want, d_flag);
    long long total = 0;
    PD_TRY(pdscan::exclusive_scan(c, d_flag, own_n, d_pos, &total));
    if (*list) { CUDA_OK(cudaFree(*list)); *list = nullptr; }
    *n = total;
    CUDA_OK(cudaMalloc(list, sizeof(int) * std::max<long long>(total, 1)));
    if (total > 0) LAUNCH(c, k_compact, nblocks(own_n, 256), 256, 0, d_flag, d_pos, c->own_lo, own_n, *list);
    return 0;
}

static int build_outlet_schedule(pdgpu_ctx* c) {
    // Lexicographic Gauss-Seidel order of apply_outlet_bc (src/boundary.cpp:92-130) is
    // preserved by the hyperplane schedule tau = i + (R+1) j + (R+1)^2 k: two outlet nodes
    // of equal tau are never within each other's stencil, and every lexicographically
    // earlier stencil neighbour has a smaller tau.
    if (c->out_nodes) { cudaFree(c->out_nodes); c->out_nodes = nullptr; }
    if (c->out_level_off) { cudaFree(c->out_level_off); c->out_level_off = nullptr; }
    c->n_levels = 0;
    c->max_level_width = 0;
    if (c->n_outlet == 0) return 0;
    std::vector<int> nodes(c->n_outlet);
    CUDA_OK(cudaMemcpy(nodes.data(), c->l_outlet, sizeof(int) * c->n_outlet, cudaMemcpyDeviceToHost));
    std::vector<long long> tau(c->n_outlet);
    long long base = c->R + 1;
    long long al_min = nodes.front() / c->P;
    c->out_l0_any = al_min * c->P;
    for (long long t = 0; t < c->n_outlet; ++t) {
        long long l = nodes[t];
        long long al = l / c->P, q = l % c->P;
        if (c->dim == 2) tau[t] = q + base * (al - al_min);
        else tau[t] = (q % c->Nx) + base * (q / c->Nx) + base * base * (al - al_min);
    }
    std::vector<int> order(c->n_outlet);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tau[x] < tau[y]; });
    std::vector<int> sorted(c->n_outlet), level_off;
    long long prev = -1;
    for (long long t = 0; t < c->n_outlet; ++t) {
        sorted[t] = nodes[order[t]];
        if (tau[order[t]] != prev) { level_off.push_back((int)t); prev = tau[order[t]]; }
    }
    level_off.push_back((int)c->n_outlet);
    c->n_levels = (int)level_off.size() - 1;
    for (int lv = 0; lv < c->n_levels; ++lv)
        c->max_level_width = std::max(c->max_level_width, level_off[lv + 1] - level_off[lv]);
    CUDA_OK(cudaMalloc(&c->out_nodes, sizeof(int) * c->n_outlet));
    CUDA_OK(cudaMalloc(&c->out_level_off, sizeof(int) * level_off.size()));
    CUDA_OK(cudaMemcpy(c->out_nodes, sorted.data(), sizeof(int) * c->n_outlet, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->out_level_off, level_off.data(), sizeof(int) * level_off.size(), cudaMemcpyHostToDevice));
    return 0;
}

int pd_rebuild_tables(pdgpu_ctx* c) {
    pd_invalidate_graphs(c);
    c->tables_epoch++;
    c->types_epoch++;
    pd_touch_flow(c);   // node types may have changed: cached |v| is stale
    long long own_n = c->own_hi - c->own_lo;
    Lat L = make_lat(c);
    double org[3] = {c->origin[0], c->origin[1], c->origin[2]};
    GeomParams g = geom_params(c->cfg, c->dim, org);

    int* d_flag = nullptr;
    long long* d_pos = nullptr;
    unsigned long long* d_counts = nullptr;
    CUDA_OK(cudaMalloc(&d_flag, sizeof(int) * own_n));
    CUDA_OK(cudaMalloc(&d_pos, sizeof(long long) * (own_n + 1)));
    CUDA_OK(cudaMalloc(&d_counts, sizeof(unsigned long long) * 16));
    CUDA_OK(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * 16, c->stream));

    LAUNCH(c, k_type_hist, std::min<unsigned>(nblocks(own_n, 256), 148 * 8), 256, 0, c->type, c->own_lo, own_n, d_counts);
    PD_TRY(build_list(c, PDGPU_WALL, d_flag, d_pos, &c->l_wall, &c->n_wall));
    PD_TRY(build_list(c, PDGPU_INLET, d_flag, d_pos, &c->l_inlet, &c->n_inlet));
    PD_TRY(build_list(c, PDGPU_OUTLET, d_flag, d_pos, &c->l_outlet, &c->n_outlet));
    PD_TRY(build_list(c, PDGPU_SOLID_MG, d_flag, d_pos, &c->l_solid, &c->n_solid));
    {   // surface solids (the ARD salt pre-pass only needs these)
        if (c->dim == 2)
            LAUNCH(c, k_surface_solid_flags<2>, nblocks(own_n, 256), 256, 0, L, c->type, c->own_lo, own_n, c->d_off,
                   c->n_off, d_flag);
        else
            LAUNCH(c, k_surface_solid_flags<3>, nblocks(own_n, 256), 256, 0, L, c->type, c->own_lo, own_n, c->d_off,
                   c->n_off, d_flag);
        long long total = 0;
        PD_TRY(pdscan::exclusive_scan(c, d_flag, own_n, d_pos, &total));
        if (c->l_ssolid) { CUDA_OK(cudaFree(c->l_ssolid)); c->l_ssolid = nullptr; }
        c->n_ssolid = total;
        CUDA_OK(cudaMalloc(&c->l_ssolid, sizeof(int) * std::max<long long>(total, 1)));
        if (total > 0) LAUNCH(c, k_compact, nblocks(own_n, 256), 256, 0, d_flag, d_pos, c->own_lo, own_n, c->l_ssolid);
    }

    // row lengths + bond counts (d_flag reused as rowlen scratch)
    LAUNCH(c, k_rowlen, nblocks(own_n, 256), 256, 0, L, c->dim, c->own_lo, own_n, c->type, c->d_off, c->n_off,
           d_flag, d_counts + 8);
    unsigned long long h[16];
    CUDA_OK(cudaMemcpyAsync(h, d_counts, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int t = 0; t < 6; ++t) c->counts[t] = (long long)h[t];
    c->ns_bonds = (long long)h[8];
    c->ard_bonds = (long long)h[9];
    c->nnz_rows = (long long)h[10];
    c->full_rows = (h[11] == 0);

    // wall mirror table
    if (c->l_wall_mirror) { CUDA_OK(cudaFree(c->l_wall_mirror)); c->l_wall_mirror = nullptr; }
    CUDA_OK(cudaMalloc(&c->l_wall_mirror, sizeof(int) * std::max<long long>(c->n_wall, 1)));
    if (c->n_wall)
        LAUNCH(c, k_wall_mirror, nblocks(c->n_wall, 128), 128, 0, g, L, c->l_wall, c->n_wall, c->type, c->d_off,
               c->n_off, c->l_wall_mirror);
    // multi-GPU: mirrors of the ghost-plane walls from the owners' relative offsets
    if (c->l_gwall) { CUDA_OK(cudaFree(c->l_gwall)); c->l_gwall = nullptr; }
    if (c->l_gwall_mirror) { CUDA_OK(cudaFree(c->l_gwall_mirror)); c->l_gwall_mirror = nullptr; }
    c->n_gwall = 0;
    if (c->nranks > 1 && c->comm) {
        if (!c->moff) CUDA_OK(cudaMalloc(&c->moff, sizeof(int) * c->NL));
        CUDA_OK(cudaMemsetAsync(c->moff, 0, sizeof(int) * c->NL, c->stream));
        if (c->n_wall)
            LAUNCH(c, k_moff_write, nblocks(c->n_wall, 256), 256, 0, c->l_wall, c->l_wall_mirror, c->n_wall, c->moff);
        PD_TRY(pd_enqueue_halo(c, 4, 0, 0));
        int* gflag = nullptr;
        long long* gpos = nullptr;
        CUDA_OK(cudaMalloc(&gflag, sizeof(int) * c->NL));
        CUDA_OK(cudaMalloc(&gpos, sizeof(long long) * (c->NL + 1)));
        LAUNCH(c, k_ghost_wall_flags, nblocks(c->NL, 256), 256, 0, c->type, c->NL, c->own_lo, c->own_hi, gflag);
        long long ng = 0;
        PD_TRY(pdscan::exclusive_scan(c, gflag, c->NL, gpos, &ng));
        c->n_gwall = ng;
        CUDA_OK(cudaMalloc(&c->l_gwall, sizeof(int) * std::max<long long>(ng, 1)));
        CUDA_OK(cudaMalloc(&c->l_gwall_mirror, sizeof(int) * std::max<long long>(ng, 1)));
        CUDA_OK(cudaMemsetAsync(c->d_int, 0, sizeof(int) * 2, c->stream));
        if (ng) {
            LAUNCH(c, k_compact, nblocks(c->NL, 256), 256, 0, gflag, gpos, 0LL, c->NL, c->l_gwall);
            LAUNCH(c, k_ghost_mirror, nblocks(ng, 256), 256, 0, c->l_gwall, ng, c->moff, c->type, c->NL,
                   c->l_gwall_mirror, c->d_int);
        }
        LAUNCH(c, k_count_ghost_inout, nblocks(c->NL, 256), 256, 0, c->type, c->NL, c->own_lo, c->own_hi, c->d_int + 1);
        int bad[2] = {0, 0};
        CUDA_OK(cudaMemcpyAsync(bad, c->d_int, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        CUDA_OK(cudaFree(gflag));
        CUDA_OK(cudaFree(gpos));
        // every rank must take the same decision: a rank that returned alone would leave the others
        // hanging in their next collective
        double worst[2] = {(double)bad[0], (double)bad[1]};
        {
            c->h_red[0] = worst[0]; c->h_red[1] = worst[1];
            CUDA_OK(cudaMemcpyAsync(c->d_red, c->h_red, sizeof(double) * 2, cudaMemcpyHostToDevice, c->stream));
            PD_TRY(pd_comm_allreduce(c, c->d_red, 2, 1));
            CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(cudaStreamSynchronize(c->stream));
            worst[0] = c->h_red[0]; worst[1] = c->h_red[1];
        }
        if (worst[0] > 0.0 || worst[1] > 0.0) {
            cudaFree(d_flag); cudaFree(d_pos); cudaFree(d_counts);
            PD_FAIL("a slab boundary is within %d planes of the inlet/outlet planes (rank %d: %d wall mirrors, %d "
                    "inlet/outlet nodes fall into ghost planes; worst rank %d / %d): use fewer ranks or a longer tube",
                    c->R, c->rank, bad[0], bad[1], (int)worst[0], (int)worst[1]);
        }
    }
    // inlet velocity table
    if (c->inlet_vax) { CUDA_OK(cudaFree(c->inlet_vax)); c->inlet_vax = nullptr; }
    CUDA_OK(cudaMalloc(&c->inlet_vax, sizeof(double) * std::max<long long>(c->n_inlet, 1)));
    if (c->n_inlet)
        LAUNCH(c, k_inlet_velocity, nblocks(c->n_inlet, 128), 128, 0, g, L, c->cfg.U_in, c->l_inlet, c->n_inlet,
               c->inlet_vax);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_flag));
    CUDA_OK(cudaFree(d_pos));
    CUDA_OK(cudaFree(d_counts));
    PD_TRY(build_outlet_schedule(c));
    PD_TRY(pd_outlet_setup(c));
    // split point for overlapping the outlet sweep with the bulk bond kernels
    c->z_cut = -1;
    c->n_wall_lo = c->n_wall;
    c->solids_below_cut = true;
    if (c->n_outlet > 0 && c->out_fast) {
        const int RZ = 8;                                  // tile::TZ (planes per tile)
        int zo = (int)(c->out_l0 / c->P);                  // first outlet plane (local)
        int z_lo = c->R;
        int zc = z_lo + ((zo - c->R - z_lo) / RZ) * RZ;
        if (zo - c->R - z_lo > 0 && zc - z_lo >= 2 * c->R + RZ) {
            c->z_cut = zc;
            std::vector<int> w(c->n_wall);
            if (c->n_wall) CUDA_OK(cudaMemcpy(w.data(), c->l_wall, sizeof(int) * c->n_wall, cudaMemcpyDeviceToHost));
            c->n_wall_lo = std::lower_bound(w.begin(), w.end(), (int)c->out_l0) - w.begin();
            if (c->n_solid) {
                int last = 0;
                CUDA_OK(cudaMemcpy(&last, c->l_solid + (c->n_solid - 1), sizeof(int), cudaMemcpyDeviceToHost));
                c->solids_below_cut = (last / c->P) < (zc - c->R);
            }
        }
    }

    PD_TRY(pd_build_nbfast(c));
    // slab contexts: wall-list ranges of the boundary planes (exchanged while the interior is computed)
    c->n_wall_b0 = 0;
    c->n_wall_b1 = c->n_wall;
    if (c->nranks > 1 && c->n_wall) {
        std::vector<int> w(c->n_wall);
        CUDA_OK(cudaMemcpy(w.data(), c->l_wall, sizeof(int) * c->n_wall, cudaMemcpyDeviceToHost));
        const long long lo_end = (long long)(2 * c->R) * c->P;                     // local planes [R, 2R)
        const long long hi_beg = (long long)(c->R + (c->a1 - c->a0) - c->R) * c->P;   // last R owned planes
        if (c->rank > 0) c->n_wall_b0 = std::lower_bound(w.begin(), w.end(), (int)lo_end) - w.begin();
        if (c->rank < c->nranks - 1) c->n_wall_b1 = std::lower_bound(w.begin(), w.end(), (int)hi_beg) - w.begin();
    }
    // column tables + active tile list of the streaming / tiled bond kernels: built here, outside of
    // any stream capture (a negative result only means that those kernels do not apply)
    if (pd_stream_prepare(c) > 0) return 1;
    if (pd_ns2d_prepare(c)) return 1;     // persistent 2D flow loop (ns2d.cu)
    return 0;
}

extern "C" int pdgpu_grid_build(pdgpu_ctx* c) {
    CHECK_CTX(c);
    Lat L = make_lat(c);
    double org[3] = {c->origin[0], c->origin[1], c->origin[2]};
    GeomParams g = geom_params(c->cfg, c->dim, org);
    if (!c->type) PD_TRY(pd_alloc_fields(c));
    LAUNCH(c, k_classify, nblocks(c->NL, 256), 256, 0, g, L, c->NL, c->type);
    c->grid_built = true;
    PD_TRY(pd_rebuild_tables(c));
    return 0;
}

extern "C" int pdgpu_grid_set_types(pdgpu_ctx* c, const uint8_t* node_type_global) {
    CHECK_CTX(c);
    if (!node_type_global) PD_FAIL("pdgpu_grid_set_types: null array");
    if (!c->type) PD_TRY(pd_alloc_fields(c));
    CUDA_OK(cudaMemsetAsync(c->type, PDGPU_OUTSIDE, c->NL, c->stream));
    int ga = std::max(c->a0 - c->R, 0), gb = std::min(c->a1 + c->R, c->Na);
    long long loff = (long long)(ga - (c->a0 - c->R)) * c->P;
    CUDA_OK(cudaMemcpyAsync(c->type + loff, node_type_global + (long long)ga * c->P, (size_t)(gb - ga) * c->P,
                            cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->grid_built = true;
    PD_TRY(pd_rebuild_tables(c));
    return 0;
}

extern "C" int pdgpu_grid_info(pdgpu_ctx* c, PdGridInfo* out) {
    CHECK_CTX(c);
    if (!out) PD_FAIL("pdgpu_grid_info: null output");
    memset(out, 0, sizeof(*out));
    out->dim = c->dim; out->Nx = c->Nx; out->Ny = c->Ny; out->Nz = c->Nz;
    out->m = c->cfg.m_ratio; out->n_off = c->n_off; out->reach = c->R;
    out->a0 = c->a0; out->a1 = c->a1;
    out->N_total = c->N_total; out->plane = c->P;
    for (int t = 0; t < 6; ++t) out->counts[t] = c->counts[t];
    out->ns_bonds = c->ns_bonds; out->ard_bonds = c->ard_bonds; out->nnz = c->nnz_rows;
    for (int d = 0; d < 3; ++d) out->origin[d] = c->origin[d];
    return 0;
}

extern "C" int pdgpu_grid_free_neighbors(pdgpu_ctx* c) {
    CHECK_CTX(c);
    cudaFree(c->csr_off); cudaFree(c->csr_idx); cudaFree(c->csr_dist); cudaFree(c->csr_evec); cudaFree(c->csr_vol);
    c->csr_off = nullptr; c->csr_idx = nullptr; c->csr_dist = c->csr_evec = c->csr_vol = nullptr;
    c->nnz = -1;
    return 0;
}

extern "C" int pdgpu_grid_build_neighbors(pdgpu_ctx* c, long long* nnz_out) {
    NEED_GRID(c);
    PD_TRY(pdgpu_grid_free_neighbors(c));
    long long own_n = c->own_hi - c->own_lo;
    Lat L = make_lat(c);
    int* d_rowlen = nullptr;
    unsigned long long* d_sums = nullptr;
    CUDA_OK(cudaMalloc(&d_rowlen, sizeof(int) * own_n));
    CUDA_OK(cudaMalloc(&d_sums, sizeof(unsigned long long) * 4));
    CUDA_OK(cudaMemsetAsync(d_sums, 0, sizeof(unsigned long long) * 4, c->stream));
    LAUNCH(c, k_rowlen, nblocks(own_n, 256), 256, 0, L, c->dim, c->own_lo, own_n, c->type, c->d_off, c->n_off,
           d_rowlen, d_sums);
    CUDA_OK(cudaMalloc(&c->csr_off, sizeof(long long) * (own_n + 1)));
    long long total = 0;
    PD_TRY(pdscan::exclusive_scan(c, d_rowlen, own_n, c->csr_off, &total));
    c->nnz = total;
    size_t n = (size_t)std::max<long long>(total, 1);
    CUDA_OK(cudaMalloc(&c->csr_idx, sizeof(int) * n));
    CUDA_OK(cudaMalloc(&c->csr_dist, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&c->csr_evec, sizeof(double) * n * c->dim));
    CUDA_OK(cudaMalloc(&c->csr_vol, sizeof(double) * n));
    long long halo_shift = (long long)(c->a0 - c->R) * c->P;     // local index + shift = global index
    LAUNCH(c, k_csr_fill, nblocks(own_n * 32, 256), 256, 0, L, c->dim, c->own_lo, own_n, halo_shift, c->type,
           c->d_off, c->n_off, c->csr_off, c->csr_idx, c->csr_dist, c->csr_evec, c->csr_vol);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_rowlen));
    CUDA_OK(cudaFree(d_sums));
    if (nnz_out) *nnz_out = total;
    return 0;
}

extern "C" int pdgpu_grid_download_csr(pdgpu_ctx* c, long long* nbr_offset, int* nbr_index, double* nbr_dist,
                                       double* nbr_evec, double* nbr_vol) {
    NEED_GRID(c);
    if (c->nnz < 0) PD_FAIL("pdgpu_grid_download_csr: call pdgpu_grid_build_neighbors first");
    long long own_n = c->own_hi - c->own_lo;
    size_t n = (size_t)c->nnz;
    if (nbr_offset) CUDA_OK(cudaMemcpy(nbr_offset, c->csr_off, sizeof(long long) * (own_n + 1), cudaMemcpyDeviceToHost));
    if (nbr_index && n) CUDA_OK(cudaMemcpy(nbr_index, c->csr_idx, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (nbr_dist && n) CUDA_OK(cudaMemcpy(nbr_dist, c->csr_dist, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (nbr_evec && n) CUDA_OK(cudaMemcpy(nbr_evec, c->csr_evec, sizeof(double) * n * c->dim, cudaMemcpyDeviceToHost));
    if (nbr_vol && n) CUDA_OK(cudaMemcpy(nbr_vol, c->csr_vol, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pdgpu_grid_download_wall_mirror(pdgpu_ctx* c, int* mirror_global) {
    NEED_GRID(c);
    if (!mirror_global) PD_FAIL("pdgpu_grid_download_wall_mirror: null output");
    long long own_n = c->own_hi - c->own_lo;
    int* d_out = nullptr;
    CUDA_OK(cudaMalloc(&d_out, sizeof(int) * own_n));
    CUDA_OK(cudaMemsetAsync(d_out, 0xff, sizeof(int) * own_n, c->stream));
    long long halo_shift = (long long)(c->a0 - c->R) * c->P;
    if (c->n_wall)
        LAUNCH(c, k_mirror_to_global, nblocks(c->n_wall, 256), 256, 0, c->l_wall, c->l_wall_mirror, c->n_wall,
               c->own_lo, halo_shift, d_out);
    CUDA_OK(cudaMemcpyAsync(mirror_global + (long long)c->a0 * c->P, d_out, sizeof(int) * own_n,
                            cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_out));
    return 0;
}

// bc.cu -- boundary operators of the reference (src/boundary.cpp) on device.
//
// Every operator works on compact per-type node lists built by the grid build, so its
// cost is proportional to the number of boundary nodes, not to N.  Density writers also
// write p = EOS(rho) (the pressure field is kept as a shadow of rho, which removes the
// reference's separate compute_pressure pass, src/pd_ns.cpp:36-50).
#include "common.cuh"
#include "geom.cuh"

// ---------------------------------------------------------------- inlet --------
// apply_inlet_bc (src/boundary.cpp:31-75): prescribed Poiseuille velocity, rho = mean rho of
// FLUID neighbours (summed in CSR order: additions only, bit-identical to the reference),
// C = C_liquid_init.
template <int DIM>
__global__ void k_bc_inlet(Lat L, const int* __restrict__ list, long long n, const double* __restrict__ vax,
                           const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
                           double* __restrict__ rho, double* __restrict__ p, double* __restrict__ vx,
                           double* __restrict__ vy, double* __restrict__ vz, double* __restrict__ C,
                           double rho_f, double gamma, double B, double C_in) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long long l = list[t];
    int q = (int)(l % L.P);
    int jj = (DIM == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    // The sum runs in CSR order (bit-identical to the reference); only the loads are batched: the
    // 12 k INLET nodes give too few threads to hide one dependent type->rho round trip per neighbour.
    double s = 0.0;
    int cnt = 0;
    constexpr int UB = 8;
    for (int o0 = 0; o0 < n_off; o0 += UB) {
        double rv[UB];
        bool ok[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int o = o0 + u;
            long long nn = -1;
            if (o < n_off) {
                const OffEntry e = off[o];
                const int ni = ii + e.di;
                bool in = ni >= 0 && ni < L.Nx;
                if (DIM == 3) { const int nj = jj + e.dj; in = in && nj >= 0 && nj < L.Ny; }
                if (in) nn = l + e.lin;
            }
            const long long safe = nn >= 0 ? nn : l;
            ok[u] = nn >= 0 && type[safe] == PDGPU_FLUID;     // (OUTSIDE neighbours are not FLUID either)
            rv[u] = rho[safe];
        }
#pragma unroll
        for (int u = 0; u < UB; ++u)
            if (ok[u]) { s += rv[u]; ++cnt; }
    }
    double r = cnt > 0 ? s / cnt : rho_f;
    rho[l] = r;
    p[l] = eos_pressure(r, rho_f, gamma, B);
    vx[l] = 0.0;
    if (DIM == 2) vy[l] = vax[t];
    else { vy[l] = 0.0; vz[l] = vax[t]; }
    C[l] = C_in;
}

// ---------------------------------------------------------------- outlet -------
// apply_outlet_bc (src/boundary.cpp:88-131) is an in-place sweep in index order: an OUTLET
// node averages FLUID and OUTLET neighbours and therefore sees already-updated values of
// OUTLET neighbours with a smaller index (Gauss-Seidel).  The sweep is reproduced exactly
// by processing the wavefront levels of grid.cu::build_outlet_schedule in order; nodes of
// one level are independent.  Only the axial component of the averaged velocity is kept
// by the reference, so only that component is summed.
//
// One CTA: a level is a few dozen nodes, one warp per node, lanes stride over the offsets.
template <int DIM>
__global__ void __launch_bounds__(1024, 1)
k_bc_outlet(Lat L, const int* __restrict__ nodes, const int* __restrict__ level_off, int n_levels,
            const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
            double* rho, double* p, double* vx, double* vy, double* vz, double* C, double rho_f, double U_in) {
    double* vax = (DIM == 2) ? vy : vz;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int lv = 0; lv < n_levels; ++lv) {
        int lo = level_off[lv], hi = level_off[lv + 1];
        for (int t = lo + wid; t < hi; t += nw) {
            long long l = nodes[t];
            int q = (int)(l % L.P);
            int jj = (DIM == 3) ? q / L.Nx : 0;
            int ii = q - jj * L.Nx;
            double sv = 0.0, sc = 0.0;
            int cnt = 0;
            for (int o = lane; o < n_off; o += 32) {
                long long nn = nbr_local(L, off[o], DIM, ii, jj, l, type);
                if (nn >= 0) {
                    uint8_t tj = type[nn];
                    if (tj == PDGPU_FLUID || tj == PDGPU_OUTLET) {
                        sv += vax[nn];
                        sc += C[nn];
                        ++cnt;
                    }
                }
            }
            sv = warp_sum(sv);
            sc = warp_sum(sc);
            cnt = warp_sum_i(cnt);
            if (lane == 0) {
                rho[l] = rho_f;
                p[l] = 0.0;   // EOS(rho_f) = B*(1^gamma - 1) = 0 exactly
                vx[l] = 0.0;
                if (DIM == 3) vy[l] = 0.0;
                if (cnt > 0) {
                    double inv_c = 1.0 / cnt;
                    vax[l] = sv * inv_c;
                    C[l] = sc / cnt;
                } else {
                    vax[l] = U_in;
                    C[l] = 0.0;
                }
            }
        }
        __syncthreads();   // level lv is final (block-wide visibility) before level lv+1 reads it
    }
}

// ---------------------------------------------------------------- walls --------
// apply_wall_mirror_proper (src/boundary.cpp:266-283) with the precomputed mirror table.
template <int DIM>
__global__ void k_bc_wall(const int* __restrict__ list, const int* __restrict__ mirror, long long n,
                          double* __restrict__ rho, double* __restrict__ p, double* __restrict__ vx,
                          double* __restrict__ vy, double* __restrict__ vz, double rho_f) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int l = list[t], m = mirror[t];
    if (m >= 0) {
        vx[l] = -vx[m];
        vy[l] = -vy[m];
        if (DIM == 3) vz[l] = -vz[m];
        rho[l] = rho[m];
        p[l] = p[m];
    } else {
        vx[l] = 0.0;
        vy[l] = 0.0;
        if (DIM == 3) vz[l] = 0.0;
        rho[l] = rho_f;
        p[l] = 0.0;
    }
}

// apply_wall_concentration_bc (src/boundary.cpp:302-321)
template <int DIM>
__global__ void k_bc_wall_conc(Lat L, const int* __restrict__ list, long long n, const uint8_t* __restrict__ type,
                               const OffEntry* __restrict__ off, int n_off, const double* Csrc, double* C,
                               double* C_other) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long long l = list[t];
    int q = (int)(l % L.P);
    int jj = (DIM == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    double s = 0.0;
    int cnt = 0;
    for (int o = 0; o < n_off; ++o) {
        long long nn = nbr_local(L, off[o], DIM, ii, jj, l, type);
        if (nn >= 0 && type[nn] == PDGPU_FLUID) { s += Csrc[nn]; ++cnt; }
    }
    const double cw = cnt > 0 ? s / cnt : 0.0;
    C[l] = cw;
    if (C_other) C_other[l] = cw;   // the ARD step would copy it there (src/pd_ard.cpp:86-89)
}

// apply_solid_surface_bc (src/boundary.cpp:381-390)
template <int DIM>
__global__ void k_bc_solid(const int* __restrict__ list, long long n, double* __restrict__ vx,
                           double* __restrict__ vy, double* __restrict__ vz) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int l = list[t];
    vx[l] = 0.0;
    vy[l] = 0.0;
    if (DIM == 3) vz[l] = 0.0;
}

// ------------------------------------------------------------- enqueue ---------

int pd_enqueue_bc_inlet(pdgpu_ctx* c, int buf, int bufC) {
    if (!c->n_inlet) return 0;
    PdConsts k = pd_consts(c->cfg, c->dim);
    Lat L = make_lat(c);
    if (c->dim == 2)
        LAUNCH(c, k_bc_inlet<2>, nblocks(c->n_inlet, 128), 128, 0, L, c->l_inlet, c->n_inlet, c->inlet_vax, c->type,
               c->d_off, c->n_off, c->rho[buf], c->p[buf], VXYZ(c, buf), c->C[bufC], c->cfg.rho_f,
               c->cfg.gamma_eos, k.B_eos, c->cfg.C_liquid_init);
    else
        LAUNCH(c, k_bc_inlet<3>, nblocks(c->n_inlet, 128), 128, 0, L, c->l_inlet, c->n_inlet, c->inlet_vax, c->type,
               c->d_off, c->n_off, c->rho[buf], c->p[buf], VXYZ(c, buf), c->C[bufC], c->cfg.rho_f,
               c->cfg.gamma_eos, k.B_eos, c->cfg.C_liquid_init);
    return 0;
}

int pd_enqueue_bc_outlet(pdgpu_ctx* c, int buf, int bufC) {
    if (!c->n_outlet) return 0;
    int r = pd_enqueue_bc_outlet_fast(c, buf, bufC);
    if (r >= 0) return r;
    Lat L = make_lat(c);
    if (c->dim == 2)
        LAUNCH(c, k_bc_outlet<2>, 1, 1024, 0, L, c->out_nodes, c->out_level_off, c->n_levels, c->type, c->d_off,
               c->n_off, c->rho[buf], c->p[buf], VXYZ(c, buf), c->C[bufC], c->cfg.rho_f, c->cfg.U_in);
    else
        LAUNCH(c, k_bc_outlet<3>, 1, 1024, 0, L, c->out_nodes, c->out_level_off, c->n_levels, c->type, c->d_off,
               c->n_off, c->rho[buf], c->p[buf], VXYZ(c, buf), c->C[bufC], c->cfg.rho_f, c->cfg.U_in);
    return 0;
}

int pd_enqueue_bc_wall_range(pdgpu_ctx* c, int buf, long long first, long long n) {
    if (n <= 0) return 0;
    if (c->dim == 2)
        LAUNCH(c, k_bc_wall<2>, nblocks(n, 256), 256, 0, c->l_wall + first, c->l_wall_mirror + first, n, c->rho[buf],
               c->p[buf], VXYZ(c, buf), c->cfg.rho_f);
    else
        LAUNCH(c, k_bc_wall<3>, nblocks(n, 256), 256, 0, c->l_wall + first, c->l_wall_mirror + first, n, c->rho[buf],
               c->p[buf], VXYZ(c, buf), c->cfg.rho_f);
    return 0;
}

int pd_enqueue_bc_wall(pdgpu_ctx* c, int buf, int part) {
    if (part == 3) {   // ghost-plane walls of a slab (refreshed locally like their owner does)
        if (!c->n_gwall) return 0;
        if (c->dim == 2)
            LAUNCH(c, k_bc_wall<2>, nblocks(c->n_gwall, 256), 256, 0, c->l_gwall, c->l_gwall_mirror, c->n_gwall,
                   c->rho[buf], c->p[buf], VXYZ(c, buf), c->cfg.rho_f);
        else
            LAUNCH(c, k_bc_wall<3>, nblocks(c->n_gwall, 256), 256, 0, c->l_gwall, c->l_gwall_mirror, c->n_gwall,
                   c->rho[buf], c->p[buf], VXYZ(c, buf), c->cfg.rho_f);
        return 0;
    }
    long long first = (part == 2) ? c->n_wall_lo : 0;
    long long n = (part == 1) ? c->n_wall_lo : c->n_wall - first;
    if (n <= 0) return 0;
    if (c->dim == 2)
        LAUNCH(c, k_bc_wall<2>, nblocks(n, 256), 256, 0, c->l_wall + first, c->l_wall_mirror + first, n, c->rho[buf],
               c->p[buf], VXYZ(c, buf), c->cfg.rho_f);
    else
        LAUNCH(c, k_bc_wall<3>, nblocks(n, 256), 256, 0, c->l_wall + first, c->l_wall_mirror + first, n, c->rho[buf],
               c->p[buf], VXYZ(c, buf), c->cfg.rho_f);
    return 0;
}

int pd_enqueue_bc_wall_conc(pdgpu_ctx* c, int bufC, bool both_buffers, int srcC) {
    if (!c->n_wall) return 0;
    Lat L = make_lat(c);
    double* other = both_buffers ? c->C[1 - bufC] : nullptr;
    const double* src = c->C[srcC >= 0 ? srcC : bufC];   // FLUID values that are averaged
    if (c->dim == 2)
        LAUNCH(c, k_bc_wall_conc<2>, nblocks(c->n_wall, 128), 128, 0, L, c->l_wall, c->n_wall, c->type, c->d_off,
               c->n_off, src, c->C[bufC], other);
    else
        LAUNCH(c, k_bc_wall_conc<3>, nblocks(c->n_wall, 128), 128, 0, L, c->l_wall, c->n_wall, c->type, c->d_off,
               c->n_off, src, c->C[bufC], other);
    return 0;
}

int pd_enqueue_bc_solid(pdgpu_ctx* c, int buf) {
    if (!c->n_solid) return 0;
    if (c->dim == 2)
        LAUNCH(c, k_bc_solid<2>, nblocks(c->n_solid, 256), 256, 0, c->l_solid, c->n_solid, VXYZ(c, buf));
    else
        LAUNCH(c, k_bc_solid<3>, nblocks(c->n_solid, 256), 256, 0, c->l_solid, c->n_solid, VXYZ(c, buf));
    return 0;
}

// ------------------------------------------------------------- C ABI -----------

extern "C" int pdgpu_bc_inlet(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    pd_touch_flow(c);
    PD_TRY(pd_enqueue_bc_inlet(c, c->cur, c->curC));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int pdgpu_bc_outlet(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    pd_touch_flow(c);
    PD_TRY(pd_enqueue_bc_outlet(c, c->cur, c->curC));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int pdgpu_bc_wall(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    pd_touch_flow(c);
    PD_TRY(pd_enqueue_bc_wall(c, c->cur));
    PD_TRY(pd_enqueue_bc_wall(c, c->cur, 3));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int pdgpu_bc_wall_new(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    pd_touch_flow(c);
    PD_TRY(pd_enqueue_bc_wall(c, 1 - c->cur));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}
// The reference applies apply_wall_concentration_bc before every explicit step and the step copies
// the WALL values into C_new (src/coupling.cpp:235-238, src/pd_ard.cpp:86-89). Device-resident
// steps skip it (no bond reads WALL C) and leave it owed; it is evaluated here from the pre-step
// buffer as soon as anything could observe WALL concentrations.
int pd_flush_wall_c(pdgpu_ctx* c) {
    if (!c->wallC_pending) return 0;
    c->wallC_pending = false;
    return pd_enqueue_bc_wall_conc(c, 1 - c->wallC_src, true, c->wallC_src);
}

// smooth_boundary_concentration (src/boundary.cpp:332-376, implicit branch src/coupling.cpp:186): FLUID
// nodes within delta of the inlet / outlet end of the fluid column take the mean C of their FLUID
// neighbours on the interior side (lower planes at the outlet end, higher planes at the inlet end).
// The reference updates C IN PLACE in index order: an outlet-side node reads lower planes that were
// already smoothed, an inlet-side node reads higher planes that are still untouched.  Both are
// reproduced by sweeping the affected planes in ascending axial order, one launch per plane (nodes
// of one plane never read each other: same-plane neighbours are not "deeper in the interior").
template <int DIM>
__global__ void k_bc_smooth_plane(Lat L, long long plane_lo, int near_in, int near_out,
                                  const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
                                  double* __restrict__ C) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= L.P) return;
    const long long l = plane_lo + q;
    if (type[l] != PDGPU_FLUID) return;
    const int jj = (DIM == 3) ? (int)(q / L.Nx) : 0;
    const int ii = (int)(q - (long long)jj * L.Nx);
    double s = 0.0;
    int cnt = 0;
    for (int o = 0; o < n_off; ++o) {           // CSR order: the sum is bit-identical to the reference's
        const OffEntry e = off[o];
        const int dax = (DIM == 3) ? e.dk : e.dj;
        if (!((near_out && dax < 0) || (near_in && dax > 0))) continue;
        const int ni = ii + e.di;
        if (ni < 0 || ni >= L.Nx) continue;
        if (DIM == 3) { const int nj = jj + e.dj; if (nj < 0 || nj >= L.Ny) continue; }
        const long long nn = l + e.lin;
        if (type[nn] == PDGPU_FLUID) { s += C[nn]; ++cnt; }
    }
    if (cnt > 0) C[l] = s / cnt;
}

int pd_enqueue_bc_smooth(pdgpu_ctx* c, int bufC) {
    Lat L = make_lat(c);
    const double y_min = -c->cfg.L_upstream, y_max = c->cfg.L_wire + c->cfg.L_downstream, delta = c->cfg.delta;
    const double o_ax = (c->dim == 2) ? c->origin[1] : c->origin[2];
    for (int a = c->a0; a < c->a1; ++a) {       // ascending axial order = the reference's index order
        const double y = geom_coord(o_ax, a, c->cfg.dx);
        const int near_in = (y - y_min < delta), near_out = (y_max - y < delta);
        if (!near_in && !near_out) continue;
        const long long plane_lo = (long long)(a - c->a0 + c->R) * c->P;
        if (c->dim == 2)
            LAUNCH(c, k_bc_smooth_plane<2>, nblocks(c->P, 128), 128, 0, L, plane_lo, near_in, near_out, c->type,
                   c->d_off, c->n_off, c->C[bufC]);
        else
            LAUNCH(c, k_bc_smooth_plane<3>, nblocks(c->P, 128), 128, 0, L, plane_lo, near_in, near_out, c->type,
                   c->d_off, c->n_off, c->C[bufC]);
    }
    return 0;
}

extern "C" int pdgpu_bc_smooth_conc(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    PD_TRY(pd_flush_wall_c(c));
    PD_TRY(pd_enqueue_bc_smooth(c, c->curC));
    if (c->nranks > 1 && c->comm) PD_TRY(pd_enqueue_halo(c, 1, c->cur, c->curC));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int pdgpu_bc_wall_conc(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    PD_TRY(pd_flush_wall_c(c));
    PD_TRY(pd_enqueue_bc_wall_conc(c, c->curC));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int pdgpu_bc_solid(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    pd_touch_flow(c);
    PD_TRY(pd_enqueue_bc_solid(c, c->cur));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ns.cu -- PD_NS_Solver (reference src/pd_ns.cpp) on device: Tait EOS, the five bond sums
// (mass convection, density diffusion, momentum convection, pressure gradient, viscous
// Laplacian) and the forward-Euler update fused in one kernel; CFL dt, convergence
// reductions, steady-state loop.
#include <algorithm>

#include "common.cuh"
#include "geom.cuh"


struct NsParams {
    double rho_f, gamma, B;         // EOS
    double c_div;                   // alpha / V_H
    double dens_diff;               // beta_lap * eta * c0 * delta
    double visc;                    // mu * beta_lap
    double rho_lo, rho_hi;          // density clamp 0.5 rho_f, 2 rho_f
};

static NsParams ns_params(const pdgpu_ctx* c) {
    PdConsts k = pd_consts(c->cfg, c->dim);
    NsParams p;
    p.rho_f = c->cfg.rho_f; p.gamma = c->cfg.gamma_eos; p.B = k.B_eos;
    p.c_div = k.alpha * k.inv_VH;
    p.dens_diff = k.dens_diff_coeff;
    p.visc = c->cfg.mu_f * k.beta_lap;
    p.rho_lo = 0.5 * c->cfg.rho_f; p.rho_hi = 2.0 * c->cfg.rho_f;
    return p;
}

// ------------------------------------------------------------------------------
// Generic kernel: one thread per owned node, offsets from the table in global memory,
// bounds/OUTSIDE checks on every bond (any geometry, any m). Summation in CSR order.
// PD_NS_Solver::step, src/pd_ns.cpp:86-179.
// ------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128)
k_ns_step_generic(Lat L, long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                  const OffEntry* __restrict__ off, int n_off, NsParams P, const double* __restrict__ d_dt,
                  const double* __restrict__ rho, const double* __restrict__ pr, const double* __restrict__ vx,
                  const double* __restrict__ vy, const double* __restrict__ vz, double* __restrict__ rho_n,
                  double* __restrict__ pr_n, double* __restrict__ vx_n, double* __restrict__ vy_n,
                  double* __restrict__ vz_n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    long long l = own_lo + t;
    double rho_i = rho[l], p_i = pr[l];
    double vi0 = vx[l], vi1 = vy[l], vi2 = (DIM == 3) ? vz[l] : 0.0;
    if (type[l] != PDGPU_FLUID) {   // :93-97 copy-through
        rho_n[l] = rho_i; pr_n[l] = p_i; vx_n[l] = vi0; vy_n[l] = vi1;
        if (DIM == 3) vz_n[l] = vi2;
        return;
    }
    const double dt = *d_dt;
    int q = (int)(l % L.P);
    int jj = (DIM == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    double mass_conv = 0.0, mass_diff = 0.0;
    double mc0 = 0.0, mc1 = 0.0, mc2 = 0.0, mp0 = 0.0, mp1 = 0.0, mp2 = 0.0, mv0 = 0.0, mv1 = 0.0, mv2 = 0.0;
    for (int o = 0; o < n_off; ++o) {
        const OffEntry e = off[o];
        long long nn = nbr_local(L, e, DIM, ii, jj, l, type);
        if (nn < 0) continue;
        double rho_j = rho[nn], p_j = pr[nn];
        double vj0 = vx[nn], vj1 = vy[nn], vj2 = (DIM == 3) ? vz[nn] : 0.0;
        // mass convection :128-130
        double dd = (rho_j * vj0 - rho_i * vi0) * e.ex + (rho_j * vj1 - rho_i * vi1) * e.ey;
        if (DIM == 3) dd += (rho_j * vj2 - rho_i * vi2) * e.ez;
        mass_conv += dd * e.w1;
        // density diffusion :133
        mass_diff += (rho_j - rho_i) * e.w2;
        // momentum convection :136-143
        double c0 = (rho_j * vj0 * vj0 - rho_i * vi0 * vi0) * e.ex + (rho_j * vj0 * vj1 - rho_i * vi0 * vi1) * e.ey;
        double c1 = (rho_j * vj1 * vj0 - rho_i * vi1 * vi0) * e.ex + (rho_j * vj1 * vj1 - rho_i * vi1 * vi1) * e.ey;
        double c2 = 0.0;
        if (DIM == 3) {
            c0 += (rho_j * vj0 * vj2 - rho_i * vi0 * vi2) * e.ez;
            c1 += (rho_j * vj1 * vj2 - rho_i * vi1 * vi2) * e.ez;
            c2 = (rho_j * vj2 * vj0 - rho_i * vi2 * vi0) * e.ex + (rho_j * vj2 * vj1 - rho_i * vi2 * vi1) * e.ey +
                 (rho_j * vj2 * vj2 - rho_i * vi2 * vi2) * e.ez;
        }
        mc0 += c0 * e.w1; mc1 += c1 * e.w1; mc2 += c2 * e.w1;
        // pressure gradient :146-148
        double dp = (p_j - p_i) * e.w1;
        mp0 += dp * e.ex; mp1 += dp * e.ey; mp2 += dp * e.ez;
        // viscous Laplacian :151-153
        mv0 += (vj0 - vi0) * e.w2; mv1 += (vj1 - vi1) * e.w2; mv2 += (vj2 - vi2) * e.w2;
    }
    double rn = rho_i + dt * (-P.c_div * mass_conv + P.dens_diff * mass_diff);   // :160-168
    rn = fmin(fmax(rn, P.rho_lo), P.rho_hi);
    rho_n[l] = rn;
    pr_n[l] = eos_pressure(rn, P.rho_f, P.gamma, P.B);
    double s = dt / rho_i;                                                        // :171-178
    vx_n[l] = vi0 + s * (-P.c_div * mc0 - P.c_div * mp0 + P.visc * mv0);
    vy_n[l] = vi1 + s * (-P.c_div * mc1 - P.c_div * mp1 + P.visc * mv1);
    if (DIM == 3) vz_n[l] = vi2 + s * (-P.c_div * mc2 - P.c_div * mp2 + P.visc * mv2);
}

int pd_enqueue_ns_step_fast(pdgpu_ctx* c, int src, const double* d_dt, int zb, int ze);   // ns_tile.cu
int pd_enqueue_ns_stream(pdgpu_ctx* c, int src, const double* d_dt, int zb, int ze);      // ns_stream.cu

int pd_enqueue_ns_step(pdgpu_ctx* c, int src, const double* d_dt, int zb, int ze) {
    if (c->opt_ns_kernel == 3) return pd_enqueue_ns_step_csr(c, src, d_dt);
    if (c->opt_ns_kernel >= 1 && c->full_rows && c->cfg.m_ratio == 3) {
        int r = (c->opt_ns_kernel >= 2) ? pd_enqueue_ns_stream(c, src, d_dt, zb, ze)
                                        : pd_enqueue_ns_step_fast(c, src, d_dt, zb, ze);
        if (r >= 0) return r;   // <0: fast path not applicable -> generic
    }
    int dst = 1 - src;
    long long own_lo = c->own_lo, own_n = c->own_hi - c->own_lo;
    if (zb >= 0) { own_lo = (long long)zb * c->P; own_n = (long long)(ze - zb) * c->P; }
    if (own_n <= 0) return 0;
    Lat L = make_lat(c);
    NsParams P = ns_params(c);
    if (c->dim == 2)
        LAUNCH(c, k_ns_step_generic<2>, nblocks(own_n, 128), 128, 0, L, own_lo, own_n, c->type, c->d_off, c->n_off,
               P, d_dt, c->rho[src], c->p[src], VXYZ(c, src), c->rho[dst], c->p[dst], VXYZ(c, dst));
    else
        LAUNCH(c, k_ns_step_generic<3>, nblocks(own_n, 128), 128, 0, L, own_lo, own_n, c->type, c->d_off, c->n_off,
               P, d_dt, c->rho[src], c->p[src], VXYZ(c, src), c->rho[dst], c->p[dst], VXYZ(c, dst));
    return 0;
}

// ------------------------------------------------------------ reductions -------
// max |v| over owned FLUID nodes (src/pd_ns.cpp:56-62): max is order-free -> atomicMax on
// the bit pattern of the non-negative double.
template <int DIM>
__global__ void k_vmax_fluid(long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                             const double* __restrict__ vx, const double* __restrict__ vy,
                             const double* __restrict__ vz, unsigned long long* __restrict__ out) {
    double m = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < own_n;
         t += (long long)gridDim.x * blockDim.x) {
        long long l = own_lo + t;
        if (type[l] == PDGPU_FLUID) {
            double s = vx[l] * vx[l] + vy[l] * vy[l];
            if (DIM == 3) s += vz[l] * vz[l];
            double v = sqrt(s);
            if (v > m) m = v;
        }
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}


int pd_max_fluid_speed(pdgpu_ctx* c, double* vmax) {
    long long own_n = c->own_hi - c->own_lo;
    CUDA_OK(cudaMemsetAsync(c->d_u64, 0, sizeof(unsigned long long), c->stream));
    unsigned g = std::min<unsigned>(nblocks(own_n, 256), 148 * 8);
    if (c->dim == 2)
        LAUNCH(c, k_vmax_fluid<2>, g, 256, 0, c->own_lo, own_n, c->type, VXYZ(c, c->cur), c->d_u64);
    else
        LAUNCH(c, k_vmax_fluid<3>, g, 256, 0, c->own_lo, own_n, c->type, VXYZ(c, c->cur), c->d_u64);
    if (c->nranks > 1 && c->comm) PD_TRY(pd_comm_allreduce(c, (double*)c->d_u64, 1, 1));
    CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_u64, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *vmax = c->h_red[0];
    return 0;
}

extern "C" int pdgpu_ns_compute_dt(pdgpu_ctx* c, double* dt) {
    NEED_GRID(c);
    if (!dt) PD_FAIL("pdgpu_ns_compute_dt: null output");
    double v_max = 0.0;
    PD_TRY(pd_max_fluid_speed(c, &v_max));
    const PdConfig& k = c->cfg;                       // src/pd_ns.cpp:65-75
    double dt_cfl = k.dx / (k.c0 + v_max + 1e-30);
    double nu = k.mu_f / k.rho_f;
    double dt_visc = 0.25 * k.dx * k.dx / (nu + 1e-30);
    double D_v = k.eta_density * k.c0 * k.delta;
    double dt_dens = 0.25 * k.dx * k.dx / (D_v + 1e-30);
    *dt = k.cfl_factor * std::min(dt_cfl, std::min(dt_visc, dt_dens));
    return 0;
}

// Convergence block of solve_steady (src/pd_ns.cpp:273-301). Deterministic two-stage
// reduction: fixed grid, per-block partials combined by one block in index order.
constexpr int kResBlocks = 592;
template <int DIM>
__global__ void __launch_bounds__(256)
k_residual_partial(long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                   const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                   const double* __restrict__ vxn, const double* __restrict__ vyn, const double* __restrict__ vzn,
                   const double* __restrict__ rhon, double* __restrict__ part) {
    __shared__ double sh[6][8];
    double num = 0.0, den = 0.0, vmax = 0.0, rmin = 1e30, rmax = -1e30, nanf = 0.0;
    long long per = (own_n + gridDim.x - 1) / gridDim.x;
    long long lo = (long long)blockIdx.x * per, hi = lo + per < own_n ? lo + per : own_n;
    for (long long t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        long long l = own_lo + t;
        if (type[l] != PDGPU_FLUID) continue;
        double a0 = vx[l], a1 = vy[l], a2 = (DIM == 3) ? vz[l] : 0.0;
        double b0 = vxn[l], b1 = vyn[l], b2 = (DIM == 3) ? vzn[l] : 0.0;
        double r = rhon[l];
        if (isnan(b0) || isnan(r)) nanf = 1.0;
        double d0 = b0 - a0, d1 = b1 - a1, d2 = b2 - a2;
        num += d0 * d0 + d1 * d1 + d2 * d2;
        den += a0 * a0 + a1 * a1 + a2 * a2;
        double vn = sqrt(b0 * b0 + b1 * b1 + b2 * b2);
        vmax = fmax(vmax, vn);
        rmin = fmin(rmin, r);
        rmax = fmax(rmax, r);
    }
    num = warp_sum(num); den = warp_sum(den); vmax = warp_max(vmax);
    rmin = warp_min(rmin); rmax = warp_max(rmax); nanf = warp_max(nanf);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = num; sh[1][wid] = den; sh[2][wid] = vmax; sh[3][wid] = rmin; sh[4][wid] = rmax; sh[5][wid] = nanf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            num += sh[0][w]; den += sh[1][w]; vmax = fmax(vmax, sh[2][w]);
            rmin = fmin(rmin, sh[3][w]); rmax = fmax(rmax, sh[4][w]); nanf = fmax(nanf, sh[5][w]);
        }
        double* o = part + 6 * blockIdx.x;
        o[0] = num; o[1] = den; o[2] = vmax; o[3] = rmin; o[4] = rmax; o[5] = nanf;
    }
}
__global__ void k_residual_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double num = 0.0, den = 0.0, vmax = 0.0, rmin = 1e30, rmax = -1e30, nanf = 0.0;
    for (int b = 0; b < nb; ++b) {
        const double* o = part + 6 * b;
        num += o[0]; den += o[1]; vmax = fmax(vmax, o[2]); rmin = fmin(rmin, o[3]); rmax = fmax(rmax, o[4]);
        nanf = fmax(nanf, o[5]);
    }
    out[0] = num; out[1] = den; out[2] = vmax; out[3] = -rmin; out[4] = rmax; out[5] = nanf;   // -rmin: max-reducible
}

extern "C" int pdgpu_ns_residual(pdgpu_ctx* c, PdResidual* out) {
    NEED_GRID(c);
    if (!out) PD_FAIL("pdgpu_ns_residual: null output");
    long long own_n = c->own_hi - c->own_lo;
    int cur = c->cur, nw = 1 - c->cur;
    double* part = c->d_red;
    double* fin = c->d_red + 6 * kResBlocks;
    if (c->dim == 2)
        LAUNCH(c, k_residual_partial<2>, kResBlocks, 256, 0, c->own_lo, own_n, c->type, VXYZ(c, cur), VXYZ(c, nw),
               c->rho[nw], part);
    else
        LAUNCH(c, k_residual_partial<3>, kResBlocks, 256, 0, c->own_lo, own_n, c->type, VXYZ(c, cur), VXYZ(c, nw),
               c->rho[nw], part);
    LAUNCH(c, k_residual_final, 1, 32, 0, part, kResBlocks, fin);
    if (c->nranks > 1 && c->comm) {
        PD_TRY(pd_comm_allreduce(c, fin, 2, 0));        // num, den : sum
        PD_TRY(pd_comm_allreduce(c, fin + 2, 4, 1));    // vmax, -rmin, rmax, nan : max
    }
    CUDA_OK(cudaMemcpyAsync(c->h_red, fin, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    out->num = c->h_red[0]; out->den = c->h_red[1]; out->v_max = c->h_red[2];
    out->rho_min = -c->h_red[3]; out->rho_max = c->h_red[4];
    out->has_nan = c->h_red[5] > 0.0; out->pad = 0;
    return 0;
}

// ----------------------------------------------------------- loop body ----------

int pd_set_dt(pdgpu_ctx* c, int slot, double dt) {
    // pageable source: the runtime stages the 8 bytes before returning, so `dt` may die
    CUDA_OK(cudaMemcpyAsync(c->d_dt + slot, &dt, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

extern "C" int pdgpu_ns_step(pdgpu_ctx* c, double dt) {
    NEED_FIELDS(c);
    pd_touch_flow(c);
    PD_TRY(pd_set_dt(c, 0, dt));
    PD_TRY(pd_enqueue_ns_step(c, c->cur, c->d_dt));
    c->p_input = c->cur; pd_pressure_recomputed(c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Channel-flow corrections of solve_steady (src/pd_ns.cpp:209-270, Poiseuille validation only):
// v_transverse = 0 on FLUID nodes of the new buffer and rho_new := mean over the FLUID nodes of
// each axial plane. One CTA per owned plane, deterministic reduction; p follows rho.
template <int DIM>
__global__ void __launch_bounds__(256)
k_channel_corrections(Lat L, long long own_lo, NsParams P, const uint8_t* __restrict__ type,
                      double* __restrict__ rho, double* __restrict__ pr, double* __restrict__ vx,
                      double* __restrict__ vy) {
    __shared__ double sh_s[8];
    __shared__ int sh_c[8];
    __shared__ double sh_avg;
    __shared__ int sh_cnt;
    const long long base = own_lo + (long long)blockIdx.x * L.P;
    double s = 0.0;
    int cnt = 0;
    for (long long q = threadIdx.x; q < L.P; q += blockDim.x)
        if (type[base + q] == PDGPU_FLUID) { s += rho[base + q]; ++cnt; }
    s = warp_sum(s);
    cnt = warp_sum_i(cnt);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh_s[wid] = s; sh_c[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        int c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += sh_s[w]; c += sh_c[w]; }
        sh_cnt = c;
        sh_avg = c > 0 ? t / c : 0.0;
    }
    __syncthreads();
    const double avg = sh_avg;
    const bool have = sh_cnt > 0;
    const double pavg = have ? eos_pressure(avg, P.rho_f, P.gamma, P.B) : 0.0;
    for (long long q = threadIdx.x; q < L.P; q += blockDim.x) {
        const long long l = base + q;
        if (type[l] != PDGPU_FLUID) continue;
        vx[l] = 0.0;                       // d != ax: x (and y in 3D) are transverse
        if (DIM == 3) vy[l] = 0.0;
        if (have) { rho[l] = avg; pr[l] = pavg; }
    }
}

int pd_enqueue_channel_corrections(pdgpu_ctx* c, int buf) {
    if (!c->cfg.channel_flow_corrections) return 0;
    Lat L = make_lat(c);
    NsParams P = ns_params(c);
    unsigned planes = (unsigned)(c->a1 - c->a0);
    if (c->dim == 2)
        LAUNCH(c, k_channel_corrections<2>, planes, 256, 0, L, c->own_lo, P, c->type, c->rho[buf], c->p[buf],
               c->v[buf][0], c->v[buf][1]);
    else
        LAUNCH(c, k_channel_corrections<3>, planes, 256, 0, L, c->own_lo, P, c->type, c->rho[buf], c->p[buf],
               c->v[buf][0], c->v[buf][1]);
    return 0;
}

// one iteration of solve_steady without the convergence block (src/pd_ns.cpp:196-205):
// reads buffer `src`, leaves the new state (with wall mirror applied) in 1-src.
//
// Same arithmetic and the same order for every data dependence as the sequential loop body; the
// schedule forks where the dependences allow it:
//  * outlet rank (pd_can_overlap): the in-place outlet sweep is a long dependent chain on two SMs and
//    only the top few planes depend on it -> side stream = sweep, outlet-plane walls, top z-tiles;
//    main stream = inlet, lower walls, solid, bulk z-tiles.
//  * slab contexts (pd_comm_overlap): the 2*reach planes next to each neighbour and the walls of the
//    `reach` planes that are sent are computed first; their exchange then runs on a third stream next
//    to the interior tiles.  (A wall's mirror lies within `reach` planes and never in a ghost plane --
//    checked in pd_rebuild_tables -- so the boundary walls only read planes computed before them.)
static int enqueue_ns_body(pdgpu_ctx* c, int src) {
    cudaStream_t main_s = c->stream, side = c->stream2, comm = c->stream3;
    const bool fork = pd_can_overlap(c);
    const bool slab = c->nranks > 1 && c->comm;
    const bool ovl = pd_comm_overlap(c);
    const bool lower = slab && c->rank > 0, upper = slab && c->rank < c->nranks - 1;
    const int z_lo = c->R, z_hi = c->R + (c->a1 - c->a0), W = 2 * c->R;
    const int dst = 1 - src;

    // ---- boundary operators on the current buffers
    if (fork) {
        PD_TRY(pd_enqueue_bc_outlet_prepass(c, src, c->curC));   // before the fork: see outlet.cu
        CUDA_OK(cudaEventRecord(c->ev_a, main_s));
        CUDA_OK(cudaStreamWaitEvent(side, c->ev_a, 0));
        StreamSwap sw(c, side);
        PD_TRY(pd_enqueue_bc_outlet_sweep(c, src, c->curC));
        PD_TRY(pd_enqueue_bc_wall(c, src, 2));
    }
    PD_TRY(pd_enqueue_bc_inlet(c, src, c->curC));
    if (!fork) PD_TRY(pd_enqueue_bc_outlet(c, src, c->curC));
    PD_TRY(pd_enqueue_bc_wall(c, src, fork ? 1 : 0));
    PD_TRY(pd_enqueue_bc_wall(c, src, 3));
    PD_TRY(pd_enqueue_bc_solid(c, src));
    if (fork) CUDA_OK(cudaEventRecord(c->ev_b, main_s));

    // ---- bond kernel
    int i0 = z_lo, i1 = fork ? c->z_cut : z_hi;   // plane range of the main stream's bulk launch
    if (ovl) {
        if (lower) { PD_TRY(pd_enqueue_ns_step(c, src, c->d_dt, z_lo, z_lo + W)); i0 = z_lo + W; }
        if (upper) { PD_TRY(pd_enqueue_ns_step(c, src, c->d_dt, z_hi - W, z_hi)); i1 = z_hi - W; }
        PD_TRY(pd_enqueue_bc_wall_range(c, dst, 0, c->n_wall_b0));
        PD_TRY(pd_enqueue_bc_wall_range(c, dst, c->n_wall_b1, c->n_wall - c->n_wall_b1));
        CUDA_OK(cudaEventRecord(c->ev_d, main_s));
        CUDA_OK(cudaStreamWaitEvent(comm, c->ev_d, 0));
        {
            StreamSwap sw(c, comm);
            PD_TRY(pd_enqueue_halo(c, 0, dst, c->curC));
        }
        CUDA_OK(cudaEventRecord(c->ev_e, comm));
    }
    PD_TRY(pd_enqueue_ns_step(c, src, c->d_dt, i0, i1));
    if (fork) {
        CUDA_OK(cudaStreamWaitEvent(side, c->ev_b, 0));
        {
            StreamSwap sw(c, side);
            PD_TRY(pd_enqueue_ns_step(c, src, c->d_dt, c->z_cut, z_hi));
        }
        CUDA_OK(cudaEventRecord(c->ev_c, side));
        CUDA_OK(cudaStreamWaitEvent(main_s, c->ev_c, 0));
    }
    // ---- wall mirror of the new buffers, exchange
    if (ovl) {
        PD_TRY(pd_enqueue_bc_wall_range(c, dst, c->n_wall_b0, c->n_wall_b1 - c->n_wall_b0));
        CUDA_OK(cudaStreamWaitEvent(main_s, c->ev_e, 0));
    } else {
        PD_TRY(pd_enqueue_bc_wall(c, dst));
        PD_TRY(pd_enqueue_channel_corrections(c, dst));
        if (slab) PD_TRY(pd_enqueue_halo(c, 0, dst, c->curC));
    }
    return 0;
}

static int run_ns_body(pdgpu_ctx* c) {
    int src = c->cur;
    pd_touch_flow(c);
    // opt_graph >= 2 also captures the bodies of slab contexts (NCCL send/recv inside the graph)
    bool use_graph = c->opt_graph && (c->opt_graph >= 2 || !(c->nranks > 1 && c->comm));
    if (!use_graph) return enqueue_ns_body(c, src);
    // the inlet/outlet BCs of the body also write C of the current concentration buffer
    // (src/boundary.cpp:31-131): one graph per (flow buffer, C buffer)
    const int bC = c->curC;
    if (!c->g_ns[src][bC]) {
        cudaGraph_t g = nullptr;
        long long before = c->launches;
        CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int r = enqueue_ns_body(c, src);
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        c->launches = before;
        if (r) return r;
        if (e != cudaSuccess) PD_FAIL("graph capture failed: %s", cudaGetErrorString(e));
        size_t nn = 0;
        CUDA_OK(cudaGraphGetNodes(g, nullptr, &nn));
        c->g_ns_nodes[src][bC] = (long long)nn;
        CUDA_OK(cudaGraphInstantiate(&c->g_ns[src][bC], g, 0));
        CUDA_OK(cudaGraphDestroy(g));
    }
    CUDA_OK(cudaGraphLaunch(c->g_ns[src][bC], c->stream));
    c->launches += c->g_ns_nodes[src][bC];
    return 0;
}

// one loop body + swap, no host synchronisation (pdgpu_step_iterate, ard.cu)
int pd_ns_body_step(pdgpu_ctx* c) {
    if (pd_ns2d_ok(c)) {
        pd_touch_flow(c);
        PD_TRY(pd_enqueue_ns2d(c, c->cur, 1));
    } else {
        PD_TRY(run_ns_body(c));
    }
    c->p_input = c->cur; pd_pressure_recomputed(c);
    c->cur = 1 - c->cur;
    return 0;
}

extern "C" int pdgpu_ns_iterate(pdgpu_ctx* c, int iters, double dt) {
    NEED_FIELDS(c);
    PD_TRY(pd_set_dt(c, 0, dt));
    if (pd_ns2d_ok(c)) {   // 2D: batches of loop bodies as one persistent kernel (ns2d.cu)
        for (int done = 0; done < iters;) {
            const int n = std::min(iters - done, 500);
            pd_touch_flow(c);
            PD_TRY(pd_enqueue_ns2d(c, c->cur, n));
            c->cur ^= (n - 1) & 1;             // buffer the last body read
            c->p_input = c->cur; pd_pressure_recomputed(c);
            c->cur = 1 - c->cur;
            done += n;
        }
        iters = 0;
    }
    for (int it = 0; it < iters; ++it) {
        PD_TRY(run_ns_body(c));
        c->p_input = c->cur; pd_pressure_recomputed(c);
        c->cur = 1 - c->cur;
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Poiseuille L2 print of solve_steady (2D only, src/pd_ns.cpp:341-368)
__global__ void k_poiseuille_l2(GeomParams g, Lat L, long long own_lo, long long own_n,
                                const uint8_t* __restrict__ type, const double* __restrict__ vy, double y_check,
                                double U_in, double* __restrict__ out) {
    __shared__ double sh[3][32];
    double err = 0.0, nrm = 0.0, cnt = 0.0;
    for (long long t = threadIdx.x; t < own_n; t += blockDim.x) {
        long long l = own_lo + t;
        if (type[l] != PDGPU_FLUID) continue;
        int al = (int)(l / L.P), i = (int)(l % L.P);
        int j = al - L.R + L.a0;
        double py = geom_coord(g.oy, j, g.dx);
        if (fabs(py - y_check) > 0.6 * g.dx) continue;
        double px = geom_coord(g.ox, i, g.dx);
        double rn = px / g.R_tube;
        if (fabs(rn) > 1.0) continue;
        double va = 1.5 * U_in * (1.0 - rn * rn);
        double d = vy[l] - va;
        err += d * d; nrm += va * va; cnt += 1.0;
    }
    err = warp_sum(err); nrm = warp_sum(nrm); cnt = warp_sum(cnt);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = err; sh[1][wid] = nrm; sh[2][wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { err += sh[0][w]; nrm += sh[1][w]; cnt += sh[2][w]; }
        out[0] = err; out[1] = nrm; out[2] = cnt;
    }
}

// PD_NS_Solver::solve_steady (src/pd_ns.cpp:182-372)
extern "C" int pdgpu_ns_solve_steady(pdgpu_ctx* c, PdSteadyResult* out, int verbose) {
    NEED_FIELDS(c);
    if (!out) PD_FAIL("pdgpu_ns_solve_steady: null output");
    if (verbose) printf("\n--- Flow solver: solving to steady state ---\n");
    double dt = 0.0;
    PD_TRY(pdgpu_ns_compute_dt(c, &dt));
    if (verbose) printf("  Initial dt = %.4e s\n", dt);
    PD_TRY(pd_set_dt(c, 0, dt));
    double epsilon = 1.0;
    bool diverged = false;
    int status = 1, iter;
    PdResidual res;
    memset(&res, 0, sizeof(res));
    const int max_iters = c->cfg.flow_max_iters;
    const bool fused = pd_ns2d_ok(c);   // 2D: the bodies up to the next convergence poll are one kernel (ns2d.cu)
    for (iter = 1; iter <= max_iters; ++iter) {
        if (fused) {
            int n = 1;
            if (iter > 10) n = std::min(100 - (iter - 1) % 100, max_iters - iter + 1);
            pd_touch_flow(c);
            PD_TRY(pd_enqueue_ns2d(c, c->cur, n));
            c->cur ^= (n - 1) & 1;             // buffer the last body read; the swaps in between happened in the kernel
            iter += n - 1;
        } else {
            PD_TRY(run_ns_body(c));
        }
        c->p_input = c->cur; pd_pressure_recomputed(c);
        if (iter <= 10 || iter % 100 == 0) {   // :273-322
            PD_TRY(pdgpu_ns_residual(c, &res));
            if (res.has_nan) {
                if (verbose) printf("  Flow DIVERGED (NaN) at iter %d\n", iter);
                diverged = true; status = 2;
                break;
            }
            epsilon = (res.den > 1e-30) ? std::sqrt(res.num / res.den) : std::sqrt(res.num);
            bool print_it = (iter <= 10) || (iter % c->cfg.output_every_flow == 0);
            if (verbose && print_it)
                printf("  Flow iter %6d: eps=%.3e  v_max=%.4e  rho=[%.2f,%.2f]  dt=%.3e\n", iter, epsilon, res.v_max,
                       res.rho_min, res.rho_max, dt);
            if (res.v_max > 100.0 * c->cfg.U_in) {
                if (verbose) printf("  Flow DIVERGED (v_max=%.2e >> U_in=%.2e) at iter %d\n", res.v_max, c->cfg.U_in, iter);
                diverged = true; status = 3;
                break;
            }
            if (epsilon < c->cfg.flow_conv_tol && iter > 100) {
                if (verbose) printf("  Flow converged at iter %d, eps=%.3e\n", iter, epsilon);
                status = 0;
                break;
            }
        }
        c->cur = 1 - c->cur;   // fields.swap_buffers() :325
        if (iter % 200 == 0) { // :331-333
            PD_TRY(pdgpu_ns_compute_dt(c, &dt));
            PD_TRY(pd_set_dt(c, 0, dt));
        }
    }
    if (!diverged && iter > max_iters && verbose)
        printf("  Flow did NOT converge after %d iters, eps=%.3e\n", max_iters, epsilon);
    out->iters = iter; out->status = status; out->eps = epsilon; out->dt = dt;
    out->v_max = res.v_max; out->rho_min = res.rho_min; out->rho_max = res.rho_max;
    out->poiseuille_l2 = -1.0; out->poiseuille_nodes = 0; out->pad = 0;
    if (!diverged && c->dim == 2 && c->nranks == 1) {
        Lat L = make_lat(c);
        double org[3] = {c->origin[0], c->origin[1], c->origin[2]};
        GeomParams g = geom_params(c->cfg, c->dim, org);
        long long own_n = c->own_hi - c->own_lo;
        LAUNCH(c, k_poiseuille_l2, 1, 1024, 0, g, L, c->own_lo, own_n, c->type, c->v[c->cur][1],
               -c->cfg.L_upstream / 2.0, c->cfg.U_in, c->d_red);
        CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double) * 3, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (c->h_red[2] > 0 && c->h_red[1] > 1e-30) {
            out->poiseuille_l2 = std::sqrt(c->h_red[0] / c->h_red[1]);
            out->poiseuille_nodes = (int)c->h_red[2];
            if (verbose)
                printf("  Poiseuille validation (upstream, %d nodes): L2 rel error = %.3e\n", out->poiseuille_nodes,
                       out->poiseuille_l2);
        }
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

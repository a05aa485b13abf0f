// csr_path.cu -- bond kernels over the materialised reference-layout CSR (Grid::nbr_offset /
// nbr_index / nbr_dist / nbr_evec / nbr_vol, src/grid.h:35-40): the general-graph formulation of
// PD_NS_Solver::step (src/pd_ns.cpp:86-179) and PD_ARD_Solver::step (src/pd_ard.cpp:81-190).
//
// This is the HBM-bound regime of BASELINE.md: every bond-update streams one CSR entry
// (4 + 8 + 8*DIM + 8 = 44 B in 3D) and gathers the neighbour's fields through L2.  One warp per
// row, lanes stride over the row's entries (coalesced 44 B/entry streams), warp-shuffle
// reduction.  It is ~5x slower than the offset-table kernels on the uniform grid and exists
// (a) as the measured reference point for the "% of HBM roofline" reading of the metric and
// (b) as the path any non-lattice neighbour list (the reference's AMR grids) would use.
// Selected with pdgpu_set_option("ns_kernel"/"ard_kernel", 3) after pdgpu_grid_build_neighbors.
#include "common.cuh"

namespace {

struct CsrNsParams {
    double rho_f, gamma, B, c_div, dens_diff, visc, rho_lo, rho_hi;
};

template <int DIM>
__global__ void __launch_bounds__(256, 3)
k_ns_step_csr(long long own_lo, long long own_n, long long halo_shift, const uint8_t* __restrict__ type,
              const long long* __restrict__ row_off, const int* __restrict__ nbr_idx,
              const double* __restrict__ nbr_dist, const double* __restrict__ nbr_evec,
              const double* __restrict__ nbr_vol, CsrNsParams P, const double* __restrict__ d_dt,
              const double* __restrict__ rho, const double* __restrict__ pr, const double* __restrict__ vx,
              const double* __restrict__ vy, const double* __restrict__ vz, double* __restrict__ rho_n,
              double* __restrict__ pr_n, double* __restrict__ vx_n, double* __restrict__ vy_n,
              double* __restrict__ vz_n) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= own_n) return;
    const long long l = own_lo + row;
    const double rho_i = rho[l], p_i = pr[l];
    const double vi0 = vx[l], vi1 = vy[l], vi2 = (DIM == 3) ? vz[l] : 0.0;
    if (type[l] != PDGPU_FLUID) {
        if (lane == 0) {
            rho_n[l] = rho_i; pr_n[l] = p_i; vx_n[l] = vi0; vy_n[l] = vi1;
            if (DIM == 3) vz_n[l] = vi2;
        }
        return;
    }
    // (rho_j v_j v_j - rho_i v_i v_i) . e = v_j q_j - v_i q_i with q = rho (v . e): same sum as
    // src/pd_ns.cpp:136-143, fewer live values (occupancy matters: the kernel is latency bound)
    double mc = 0.0, md = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const long long beg = row_off[row], end = row_off[row + 1];
    for (long long jj = beg + lane; jj < end; jj += 32) {
        const long long j = (long long)nbr_idx[jj] - halo_shift;
        const double xi = nbr_dist[jj], V = nbr_vol[jj];
        const double e0 = nbr_evec[jj * DIM], e1 = nbr_evec[jj * DIM + 1], e2 = (DIM == 3) ? nbr_evec[jj * DIM + 2] : 0.0;
        if (V < 1e-30) continue;
        const double inv_xi = 1.0 / xi, w1 = inv_xi * V, w2 = inv_xi * inv_xi * V;
        const double rho_j = rho[j], dp = pr[j] - p_i;
        const double vj0 = vx[j], vj1 = vy[j], vj2 = (DIM == 3) ? vz[j] : 0.0;
        const double q_j = rho_j * (vj0 * e0 + vj1 * e1 + vj2 * e2);
        const double q_i = rho_i * (vi0 * e0 + vi1 * e1 + vi2 * e2);
        mc += (q_j - q_i) * w1;
        md += (rho_j - rho_i) * w2;
        a0 += (vj0 * q_j - vi0 * q_i + dp * e0) * w1;      // convection + pressure (both x -alpha/V_H)
        a1 += (vj1 * q_j - vi1 * q_i + dp * e1) * w1;
        a2 += (vj2 * q_j - vi2 * q_i + dp * e2) * w1;
        s0 += (vj0 - vi0) * w2; s1 += (vj1 - vi1) * w2; s2 += (vj2 - vi2) * w2;
    }
    mc = warp_sum(mc); md = warp_sum(md);
    a0 = warp_sum(a0); a1 = warp_sum(a1); s0 = warp_sum(s0); s1 = warp_sum(s1);
    if (DIM == 3) { a2 = warp_sum(a2); s2 = warp_sum(s2); }
    if (lane != 0) return;
    const double dt = *d_dt;
    double rn = rho_i + dt * (-P.c_div * mc + P.dens_diff * md);
    rn = fmin(fmax(rn, P.rho_lo), P.rho_hi);
    rho_n[l] = rn;
    pr_n[l] = eos_pressure(rn, P.rho_f, P.gamma, P.B);
    const double s = dt / rho_i;
    vx_n[l] = vi0 + s * (-P.c_div * a0 + P.visc * s0);
    vy_n[l] = vi1 + s * (-P.c_div * a1 + P.visc * s1);
    if (DIM == 3) vz_n[l] = vi2 + s * (-P.c_div * a2 + P.visc * s2);
}

struct CsrArdParams {
    double D_liquid, D_grain, D_gb, D_precip, decay, alpha_dx, beta, div_coeff;
};

template <int DIM>
__global__ void __launch_bounds__(256)
k_ard_step_csr(long long own_lo, long long own_n, long long halo_shift, const uint8_t* __restrict__ type,
               const long long* __restrict__ row_off, const int* __restrict__ nbr_idx,
               const double* __restrict__ nbr_dist, const double* __restrict__ nbr_evec,
               const double* __restrict__ nbr_vol, CsrArdParams P, const double* __restrict__ d_dt,
               const double* __restrict__ C, const double* __restrict__ vx, const double* __restrict__ vy,
               const double* __restrict__ vz, const double* __restrict__ vmag, const double* __restrict__ dsol,
               double* __restrict__ C_n) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= own_n) return;
    const long long l = own_lo + row;
    const uint8_t ti = type[l];
    const double C_i = C[l];
    if (ti != PDGPU_FLUID && ti != PDGPU_SOLID_MG) {
        if (lane == 0) C_n[l] = C_i;
        return;
    }
    const bool i_fluid = (ti == PDGPU_FLUID);
    double vi0 = 0.0, vi1 = 0.0, vi2 = 0.0, vi_mag = 0.0;
    if (i_fluid) { vi0 = vx[l]; vi1 = vy[l]; if (DIM == 3) vi2 = vz[l]; vi_mag = vmag[l]; }
    const double ds_i = dsol[l];
    double diff = 0.0, adv = 0.0;
    const long long beg = row_off[row], end = row_off[row + 1];
    for (long long jj = beg + lane; jj < end; jj += 32) {
        const long long j = (long long)nbr_idx[jj] - halo_shift;
        const double xi = nbr_dist[jj], V = nbr_vol[jj];
        if (V < 1e-30) continue;
        const uint8_t tj = type[j];
        if (tj == PDGPU_WALL || tj == PDGPU_OUTSIDE) continue;
        const bool j_fluid = (tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET);
        if (!i_fluid && !j_fluid) continue;
        const double inv_xi = 1.0 / xi, w1 = inv_xi * V, w2 = inv_xi * inv_xi * V;
        const double dC = C[j] - C_i;
        double D;
        if (i_fluid && j_fluid) {
            D = P.D_liquid + P.alpha_dx * fmax(vi_mag, vmag[j]);
            double vde = vi0 * nbr_evec[jj * DIM] + vi1 * nbr_evec[jj * DIM + 1];
            if (DIM == 3) vde += vi2 * nbr_evec[jj * DIM + 2];
            adv += dC * vde * w1;
        } else {
            D = i_fluid ? dsol[j] : ds_i;   // interface diffusivity of the solid end (0 when salt blocked)
        }
        diff += P.beta * D * dC * w2;
    }
    diff = warp_sum(diff);
    adv = warp_sum(adv);
    if (lane != 0) return;
    const double cn = C_i + (*d_dt) * (diff - P.div_coeff * adv);
    C_n[l] = cn < 0.0 ? 0.0 : cn;
}

}  // namespace

int pd_enqueue_ns_step_csr(pdgpu_ctx* c, int src, const double* d_dt) {
    if (c->nnz < 0) PD_FAIL("ns_kernel = 3 (CSR path) needs pdgpu_grid_build_neighbors first");
    PdConsts k = pd_consts(c->cfg, c->dim);
    CsrNsParams P;
    P.rho_f = c->cfg.rho_f; P.gamma = c->cfg.gamma_eos; P.B = k.B_eos; P.c_div = k.alpha * k.inv_VH;
    P.dens_diff = k.dens_diff_coeff; P.visc = c->cfg.mu_f * k.beta_lap;
    P.rho_lo = 0.5 * c->cfg.rho_f; P.rho_hi = 2.0 * c->cfg.rho_f;
    const long long own_n = c->own_hi - c->own_lo;
    const long long halo_shift = (long long)(c->a0 - c->R) * c->P;
    const int dst = 1 - src;
    if (c->dim == 2)
        LAUNCH(c, k_ns_step_csr<2>, nblocks(own_n * 32, 256), 256, 0, c->own_lo, own_n, halo_shift, c->type, c->csr_off,
               c->csr_idx, c->csr_dist, c->csr_evec, c->csr_vol, P, d_dt, c->rho[src], c->p[src], VXYZ(c, src),
               c->rho[dst], c->p[dst], VXYZ(c, dst));
    else
        LAUNCH(c, k_ns_step_csr<3>, nblocks(own_n * 32, 256), 256, 0, c->own_lo, own_n, halo_shift, c->type, c->csr_off,
               c->csr_idx, c->csr_dist, c->csr_evec, c->csr_vol, P, d_dt, c->rho[src], c->p[src], VXYZ(c, src),
               c->rho[dst], c->p[dst], VXYZ(c, dst));
    return 0;
}

// expects the ARD pre-pass (vmag, dsol) to have run
int pd_enqueue_ard_step_csr(pdgpu_ctx* c, int buf, int srcC, const double* d_dt) {
    if (c->nnz < 0) PD_FAIL("ard_kernel = 3 (CSR path) needs pdgpu_grid_build_neighbors first");
    PdConsts k = pd_consts(c->cfg, c->dim);
    CsrArdParams P;
    P.D_liquid = c->cfg.D_liquid; P.D_grain = c->cfg.D_grain; P.D_gb = c->cfg.D_gb; P.D_precip = c->cfg.D_precip;
    P.decay = 1.0; P.alpha_dx = c->cfg.alpha_art_diff * c->cfg.dx; P.beta = k.beta_lap; P.div_coeff = k.alpha / k.V_H;
    const long long own_n = c->own_hi - c->own_lo;
    const long long halo_shift = (long long)(c->a0 - c->R) * c->P;
    const int dstC = 1 - srcC;
    if (c->dim == 2)
        LAUNCH(c, k_ard_step_csr<2>, nblocks(own_n * 32, 256), 256, 0, c->own_lo, own_n, halo_shift, c->type,
               c->csr_off, c->csr_idx, c->csr_dist, c->csr_evec, c->csr_vol, P, d_dt, c->C[srcC], VXYZ(c, buf),
               c->vmag, c->dsol, c->C[dstC]);
    else
        LAUNCH(c, k_ard_step_csr<3>, nblocks(own_n * 32, 256), 256, 0, c->own_lo, own_n, halo_shift, c->type,
               c->csr_off, c->csr_idx, c->csr_dist, c->csr_evec, c->csr_vol, P, d_dt, c->C[srcC], VXYZ(c, buf),
               c->vmag, c->dsol, c->C[dstC]);
    return 0;
}

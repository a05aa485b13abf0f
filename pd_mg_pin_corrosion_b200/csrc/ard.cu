// ard.cu -- explicit PD_ARD_Solver (reference src/pd_ard.cpp) on device: salt-layer
// pre-pass, bond-classified diffusion + artificial diffusion + non-conservative advection
// with the forward-Euler update in one kernel, CFL dt, phase change, diagnostics.
#include <algorithm>

#include "common.cuh"

struct ArdParams {
    double D_liquid, D_grain, D_gb, D_precip, decay;
    double alpha_dx;        // alpha_art_diff * dx
    double beta;            // 4/(pi delta^2) or 12/(pi delta^2)
    double div_coeff;       // alpha_p / V_H
    double C_sat;
};

static ArdParams ard_params(const pdgpu_ctx* c) {
    PdConsts k = pd_consts(c->cfg, c->dim);
    ArdParams p;
    p.D_liquid = c->cfg.D_liquid; p.D_grain = c->cfg.D_grain; p.D_gb = c->cfg.D_gb; p.D_precip = c->cfg.D_precip;
    p.decay = 1.0;                                                      // src/pd_ard.cpp:75-79
    if (c->cfg.corrosion_decay_l > 0.0) p.decay = std::pow(10.0, -c->volume_loss / c->cfg.corrosion_decay_l);
    p.alpha_dx = c->cfg.alpha_art_diff * c->cfg.dx;
    p.beta = k.beta_lap;
    p.div_coeff = k.alpha / k.V_H;
    p.C_sat = c->cfg.C_sat;
    return p;
}

// vmag = |v| for fluid-like nodes (FLUID/INLET/OUTLET), -1 otherwise (feeds D_art,
// src/pd_ard.cpp:166-170).  wpack is what the tiled kernel stages instead of (vmag, dsol): the
// node's own fluid-fluid diffusivity  D_l + alpha dx |v|  >= +0 for fluid-like nodes -- the bond
// value D_l + alpha dx max(|v_i|, |v_j|) is the max of the two end values, bit for bit, because
// the rounded fma is monotone in |v| --, -dsol <= -0 for SOLID_MG nodes (written by the salt
// pre-pass below, never here) and -0.0 for WALL/OUTSIDE: the SIGN BIT says "not fluid-like".
// Velocities of FLUID nodes are frozen during the ARD phase, so the full pass runs once per flow
// state (pd_ensure_vmag); only the outlet planes are refreshed per step.
template <int DIM>
__global__ void k_ard_vmag(long long lo, long long hi, const uint8_t* __restrict__ type,
                           const double* __restrict__ vx, const double* __restrict__ vy,
                           const double* __restrict__ vz, double D_liquid, double alpha_dx,
                           double* __restrict__ vmag, double* __restrict__ wpack) {
    long long l = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= hi) return;
    uint8_t ty = type[l];
    double s = vx[l] * vx[l] + vy[l] * vy[l];
    if (DIM == 3) s += vz[l] * vz[l];
    bool fluid_like = (ty == PDGPU_FLUID || ty == PDGPU_INLET || ty == PDGPU_OUTLET);
    const double vm = sqrt(s);
    vmag[l] = fluid_like ? vm : -1.0;
    if (ty != PDGPU_SOLID_MG) wpack[l] = fluid_like ? fma(alpha_dx, vm, D_liquid) : -0.0;
}

// Per step, owned SOLID_MG nodes only: the salt-layer flag of src/pd_ard.cpp:61-73 (any FLUID
// neighbour with C >= C_sat) and the interface diffusivity dsol = 2 D_l D_s / (D_l + D_s + 1e-30)
// (0 when blocked), src/pd_ard.cpp:140-162.
template <int DIM>
__global__ void k_ard_prepass_solids(Lat L, const int* __restrict__ l_solid, long long n_solid,
                                     const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
                                     const double* __restrict__ C, const uint8_t* __restrict__ is_gb,
                                     const uint8_t* __restrict__ is_precip, ArdParams P,
                                     uint8_t* __restrict__ salt, double* __restrict__ dsol,
                                     double* __restrict__ wpack) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_solid) return;
    long long l = l_solid[t];
    int q = (int)(l % L.P);
    int jj = (DIM == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    // any FLUID neighbour at or above saturation blocks the node (order-free): loads in batches of 8
    uint8_t blocked = 0;
    constexpr int UB = 8;
    for (int o0 = 0; o0 < n_off && !blocked; o0 += UB) {
        bool hit[UB];
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
            hit[u] = nn >= 0 && type[safe] == PDGPU_FLUID && C[safe] >= P.C_sat;
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) blocked |= hit[u] ? 1 : 0;
    }
    double ds = 0.0;
    if (!blocked) {
        double D_s = is_gb[l] ? P.D_gb : (is_precip[l] ? P.D_precip : P.D_grain);
        D_s *= P.decay;
        ds = 2.0 * P.D_liquid * D_s / (P.D_liquid + D_s + 1e-30);
    }
    salt[l] = blocked;
    dsol[l] = ds;
    wpack[l] = -ds;   // sign bit set (-0.0 when blocked)
}

// Generic kernel: one thread per owned node (PD_ARD_Solver::step, src/pd_ard.cpp:81-190).
template <int DIM>
__global__ void __launch_bounds__(128)
k_ard_step_generic(Lat L, long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                   const OffEntry* __restrict__ off, int n_off, ArdParams P, const double* __restrict__ d_dt,
                   const double* __restrict__ C, const double* __restrict__ vx, const double* __restrict__ vy,
                   const double* __restrict__ vz, const double* __restrict__ vmag,
                   const uint8_t* __restrict__ is_gb, const uint8_t* __restrict__ is_precip,
                   const uint8_t* __restrict__ salt, double* __restrict__ C_n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    long long l = own_lo + t;
    uint8_t ti = type[l];
    double C_i = C[l];
    if (ti != PDGPU_FLUID && ti != PDGPU_SOLID_MG) { C_n[l] = C_i; return; }   // :86-89
    const double dt = *d_dt;
    bool i_fluid = (ti == PDGPU_FLUID);
    double vi0 = 0.0, vi1 = 0.0, vi2 = 0.0, vi_mag = 0.0;
    if (i_fluid) { vi0 = vx[l]; vi1 = vy[l]; if (DIM == 3) vi2 = vz[l]; vi_mag = vmag[l]; }
    int q = (int)(l % L.P);
    int jj = (DIM == 3) ? q / L.Nx : 0;
    int ii = q - jj * L.Nx;
    double diff_sum = 0.0, adv_sum = 0.0;
    for (int o = 0; o < n_off; ++o) {
        const OffEntry e = off[o];
        long long nn = nbr_local(L, e, DIM, ii, jj, l, type);
        if (nn < 0) continue;
        uint8_t tj = type[nn];
        if (tj == PDGPU_WALL) continue;                                          // :120
        bool j_fluid = (tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET);
        if (!i_fluid && !j_fluid) continue;                                      // solid-solid :134
        double dC = C[nn] - C_i;
        double D;
        if (i_fluid && j_fluid) {                                                // :137-139, :166-170
            D = P.D_liquid + P.alpha_dx * fmax(vi_mag, vmag[nn]);
            double vde = vi0 * e.ex + vi1 * e.ey;
            if (DIM == 3) vde += vi2 * e.ez;
            adv_sum += dC * vde * e.w1;                                          // :178-181
        } else {                                                                 // interface :140-162
            long long s = i_fluid ? nn : l;
            if (salt[s]) D = 0.0;
            else {
                double D_s = is_gb[s] ? P.D_gb : (is_precip[s] ? P.D_precip : P.D_grain);
                D_s *= P.decay;
                D = 2.0 * P.D_liquid * D_s / (P.D_liquid + D_s + 1e-30);
            }
        }
        diff_sum += P.beta * D * dC * e.w2;                                      // :173
    }
    adv_sum *= P.div_coeff;
    double cn = C_i + dt * (diff_sum - adv_sum);
    C_n[l] = cn < 0.0 ? 0.0 : cn;
}

int pd_enqueue_ard_tile(pdgpu_ctx* c, int buf, int srcC, const double* d_dt, int zb, int ze, bool do_solid,
                        bool skip_wall_copy);   // ard_tile.cu

int pd_enqueue_ard_vmag_range(pdgpu_ctx* c, int buf, long long lo, long long hi) {
    if (hi <= lo) return 0;
    long long n = hi - lo;
    if (c->dim == 2)
        LAUNCH(c, k_ard_vmag<2>, nblocks(n, 256), 256, 0, lo, hi, c->type, VXYZ(c, buf), c->cfg.D_liquid,
               c->cfg.alpha_art_diff * c->cfg.dx, c->vmag, c->wpack);
    else
        LAUNCH(c, k_ard_vmag<3>, nblocks(n, 256), 256, 0, lo, hi, c->type, VXYZ(c, buf), c->cfg.D_liquid,
               c->cfg.alpha_art_diff * c->cfg.dx, c->vmag, c->wpack);
    return 0;
}

// full |v| pass when the cached array does not belong to the current flow state (never inside a
// graph capture: called by the public entry points before the loop bodies are enqueued)
int pd_ensure_vmag(pdgpu_ctx* c, int buf) {
    if (c->vmag_epoch == c->flow_epoch && c->vmag_buf == buf) return 0;
    PD_TRY(pd_enqueue_ard_vmag_range(c, buf, 0, c->NL));
    c->vmag_epoch = c->flow_epoch;
    c->vmag_buf = buf;
    return 0;
}

// salt flag / interface diffusivity of the SURFACE solids (interior solids have no fluid-like
// neighbour: nobody reads their entries)
int pd_enqueue_ard_prepass_solids(pdgpu_ctx* c, int srcC) {
    if (!c->n_ssolid) return 0;
    Lat L = make_lat(c);
    ArdParams P = ard_params(c);
    if (c->dim == 2)
        LAUNCH(c, k_ard_prepass_solids<2>, nblocks(c->n_ssolid, 128), 128, 0, L, c->l_ssolid, c->n_ssolid, c->type,
               c->d_off, c->n_off, c->C[srcC], c->is_gb, c->is_precip, P, c->salt, c->dsol, c->wpack);
    else
        LAUNCH(c, k_ard_prepass_solids<3>, nblocks(c->n_ssolid, 128), 128, 0, L, c->l_ssolid, c->n_ssolid, c->type,
               c->d_off, c->n_off, c->C[srcC], c->is_gb, c->is_precip, P, c->salt, c->dsol, c->wpack);
    return 0;
}

// bond kernel over the local plane range [zb, ze) (negative = all owned planes)
int pd_enqueue_ard_main(pdgpu_ctx* c, int buf, int srcC, const double* d_dt, int zb, int ze, bool do_solid,
                        bool skip_wall_copy) {
    if (c->opt_ard_kernel == 3) return pd_enqueue_ard_step_csr(c, buf, srcC, d_dt);
    if (c->opt_ard_kernel >= 1) {
        int r = pd_enqueue_ard_tile(c, buf, srcC, d_dt, zb, ze, do_solid, skip_wall_copy);
        if (r >= 0) return r;
    }
    Lat L = make_lat(c);
    ArdParams P = ard_params(c);
    long long own_lo = c->own_lo, own_n = c->own_hi - c->own_lo;
    if (zb >= 0) { own_lo = (long long)zb * c->P; own_n = (long long)(ze - zb) * c->P; }
    if (own_n <= 0) return 0;
    int dstC = 1 - srcC;
    if (c->dim == 2)
        LAUNCH(c, k_ard_step_generic<2>, nblocks(own_n, 128), 128, 0, L, own_lo, own_n, c->type, c->d_off,
               c->n_off, P, d_dt, c->C[srcC], VXYZ(c, buf), c->vmag, c->is_gb, c->is_precip, c->salt, c->C[dstC]);
    else
        LAUNCH(c, k_ard_step_generic<3>, nblocks(own_n, 128), 128, 0, L, own_lo, own_n, c->type, c->d_off,
               c->n_off, P, d_dt, c->C[srcC], VXYZ(c, buf), c->vmag, c->is_gb, c->is_precip, c->salt, c->C[dstC]);
    return 0;
}

int pd_enqueue_ard_step(pdgpu_ctx* c, int buf, int srcC, const double* d_dt) {
    PD_TRY(pd_ensure_vmag(c, buf));
    PD_TRY(pd_enqueue_ard_prepass_solids(c, srcC));
    if (c->nranks > 1 && c->comm) PD_TRY(pd_enqueue_halo(c, 3, buf, srcC));   // salt flags + dsol of ghost solids
    return pd_enqueue_ard_main(c, buf, srcC, d_dt, -1, -1, true);
}

// ----------------------------------------------------------------- C ABI -------
extern "C" int pdgpu_ard_set_volume_loss(pdgpu_ctx* c, double vl) {
    if (!c) PD_FAIL("null context");
    if (vl != c->volume_loss) pd_invalidate_graphs(c);   // decay factor is baked into the kernel params
    c->volume_loss = vl;
    return 0;
}

extern "C" int pdgpu_ard_compute_dt(pdgpu_ctx* c, double* dt) {   // src/pd_ard.cpp:34-53
    NEED_GRID(c);
    if (!dt) PD_FAIL("pdgpu_ard_compute_dt: null output");
    const PdConfig& k = c->cfg;
    double D_max = std::max(k.D_liquid, std::max(k.D_grain, k.D_gb));
    double v_max = 0.0;
    PD_TRY(pd_max_fluid_speed(c, &v_max));
    double D_eff = D_max + k.alpha_art_diff * v_max * k.dx;
    double dt_diff = 0.25 * k.dx * k.dx / (D_eff + 1e-30);
    double dt_adv = k.dx / (v_max + 1e-30);
    *dt = k.cfl_factor_corr * std::min(dt_diff, dt_adv);
    return 0;
}

extern "C" int pdgpu_ard_step(pdgpu_ctx* c, double dt) {
    NEED_FIELDS(c);
    PD_TRY(pd_flush_wall_c(c));
    PD_TRY(pd_set_dt(c, 1, dt));
    PD_TRY(pd_enqueue_ard_step(c, c->cur, c->curC, c->d_dt + 1));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// explicit coupling-loop body (src/coupling.cpp:232-238).  Same arithmetic and dependence order as
// the sequential body; schedule as in enqueue_ns_body (ns.cu):
//  * outlet rank: sweep + |v| of the outlet planes + the z-tiles that see outlet planes on the side stream;
//  * slab contexts: the chain  {weights of ghost solids} -> {boundary tiles + SOLID_MG rows} -> {C of
//    the boundary planes}  runs on a third stream next to the interior tiles, which read owned planes only.
static int enqueue_ard_body(pdgpu_ctx* c, int buf, int srcC) {
    cudaStream_t main_s = c->stream, side = c->stream2, comm = c->stream3;
    const bool fork = pd_can_overlap(c);
    const bool slab = c->nranks > 1 && c->comm;
    const bool ovl = pd_comm_overlap(c);
    const bool lower = slab && c->rank > 0, upper = slab && c->rank < c->nranks - 1;
    const int z_lo = c->R, z_hi = c->R + (c->a1 - c->a0), W = c->R;
    const bool tiles = (c->opt_ard_kernel == 1 || c->opt_ard_kernel == 2);   // kernels that can skip the wall copy
    const double* d_dt = c->d_dt + 1;
    const bool skipw = tiles && fork;   // the side stream's wall-concentration BC writes WALL C of both buffers

    if (fork) {
        PD_TRY(pd_enqueue_bc_outlet_prepass(c, buf, srcC));   // before the fork: see outlet.cu
        CUDA_OK(cudaEventRecord(c->ev_a, main_s));
        CUDA_OK(cudaStreamWaitEvent(side, c->ev_a, 0));
        StreamSwap sw(c, side);
        PD_TRY(pd_enqueue_bc_outlet_sweep(c, buf, srcC));
        PD_TRY(pd_enqueue_ard_vmag_range(c, buf, c->out_l0, c->NL));
        // WALL concentrations are never read by a bond (src/pd_ard.cpp:120): off the critical path
        if (tiles && !c->opt_lazy_wallc) PD_TRY(pd_enqueue_bc_wall_conc(c, srcC, true));
    }
    PD_TRY(pd_enqueue_bc_inlet(c, buf, srcC));
    if (!fork) {
        PD_TRY(pd_enqueue_bc_outlet(c, buf, srcC));
        if (!c->opt_lazy_wallc) PD_TRY(pd_enqueue_bc_wall_conc(c, srcC));
        if (c->n_outlet) PD_TRY(pd_enqueue_ard_vmag_range(c, buf, c->out_l0_any, c->NL));   // outlet velocities just changed
    } else if (!tiles && !c->opt_lazy_wallc) {
        PD_TRY(pd_enqueue_bc_wall_conc(c, srcC));
    }
    PD_TRY(pd_enqueue_ard_prepass_solids(c, srcC));
    if (slab && !ovl) PD_TRY(pd_enqueue_halo(c, 3, buf, srcC));   // salt flags + weights of ghost solids
    if (fork) CUDA_OK(cudaEventRecord(c->ev_b, main_s));

    int i0 = z_lo, i1 = fork ? c->z_cut : z_hi;
    if (ovl) {
        CUDA_OK(cudaEventRecord(c->ev_d, main_s));
        CUDA_OK(cudaStreamWaitEvent(comm, c->ev_d, 0));
        {
            StreamSwap sw(c, comm);
            PD_TRY(pd_enqueue_halo(c, 3, buf, srcC));
            bool solids_done = false;
            if (lower) {
                PD_TRY(pd_enqueue_ard_main(c, buf, srcC, d_dt, z_lo, z_lo + W, true, skipw));
                solids_done = true;
            }
            if (upper) PD_TRY(pd_enqueue_ard_main(c, buf, srcC, d_dt, z_hi - W, z_hi, !solids_done, skipw));
            PD_TRY(pd_enqueue_halo(c, 1, buf, 1 - srcC));
        }
        CUDA_OK(cudaEventRecord(c->ev_e, comm));
        if (lower) i0 = z_lo + W;
        if (upper) i1 = z_hi - W;
    }
    // SOLID_MG rows: with the boundary chain (ovl); else with the top tiles of the outlet rank (they may read
    // OUTLET C of the sweep) or with the only launch there is
    PD_TRY(pd_enqueue_ard_main(c, buf, srcC, d_dt, i0, i1, !ovl && !fork, skipw));
    if (fork) {
        CUDA_OK(cudaStreamWaitEvent(side, c->ev_b, 0));
        {
            StreamSwap sw(c, side);
            PD_TRY(pd_enqueue_ard_main(c, buf, srcC, d_dt, c->z_cut, z_hi, !ovl, skipw));
        }
        CUDA_OK(cudaEventRecord(c->ev_c, side));
        CUDA_OK(cudaStreamWaitEvent(main_s, c->ev_c, 0));
    }
    if (ovl) CUDA_OK(cudaStreamWaitEvent(main_s, c->ev_e, 0));
    else if (slab) PD_TRY(pd_enqueue_halo(c, 1, buf, 1 - srcC));
    return 0;
}

// one corrosion loop body + swap of C, no host synchronisation (graph per (flow buffer, C buffer) where allowed)
static int ard_body_step(pdgpu_ctx* c) {
    // opt_graph >= 2 also captures the bodies of slab contexts (NCCL send/recv inside the graph)
    bool use_graph = c->opt_graph && (c->opt_graph >= 2 || !(c->nranks > 1 && c->comm));
    {
        int buf = c->cur, srcC = c->curC;
        if (!use_graph) {
            PD_TRY(enqueue_ard_body(c, buf, srcC));
        } else {
            if (!c->g_ard[buf][srcC]) {
                cudaGraph_t g = nullptr;
                long long before = c->launches;
                CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                int r = enqueue_ard_body(c, buf, srcC);
                cudaError_t e = cudaStreamEndCapture(c->stream, &g);
                c->launches = before;
                if (r) return r;
                if (e != cudaSuccess) PD_FAIL("graph capture failed: %s", cudaGetErrorString(e));
                size_t nn = 0;
                CUDA_OK(cudaGraphGetNodes(g, nullptr, &nn));
                c->g_ard_nodes[buf][srcC] = (long long)nn;
                CUDA_OK(cudaGraphInstantiate(&c->g_ard[buf][srcC], g, 0));
                CUDA_OK(cudaGraphDestroy(g));
            }
            CUDA_OK(cudaGraphLaunch(c->g_ard[buf][srcC], c->stream));
            c->launches += c->g_ard_nodes[buf][srcC];
        }
        c->curC = 1 - c->curC;   // std::swap(fields.C, fields.C_new)
        if (c->opt_lazy_wallc) { c->wallC_pending = true; c->wallC_src = 1 - c->curC; }
    }
    return 0;
}

extern "C" int pdgpu_ard_iterate(pdgpu_ctx* c, int steps, double dt) {
    NEED_FIELDS(c);
    if (steps <= 0) return 0;
    // an owed wall-C evaluation is superseded by the one of the first step of this call (both write
    // every WALL node from FLUID values only), unless nobody evaluates it again: keep it pending
    PD_TRY(pd_set_dt(c, 1, dt));
    PD_TRY(pd_ensure_vmag(c, c->cur));
    if (pd_ns2d_ok(c) && c->opt_lazy_wallc) {   // 2D: batches of loop bodies as one persistent kernel (ns2d.cu)
        for (int done = 0; done < steps;) {
            const int n = std::min(steps - done, 500);
            PD_TRY(pd_enqueue_ard2d(c, c->cur, c->curC, n));
            if (n & 1) c->curC = 1 - c->curC;
            c->wallC_pending = true; c->wallC_src = 1 - c->curC;   // the buffer the last step read
            done += n;
        }
        if (c->n_outlet) PD_TRY(pd_enqueue_ard_vmag_range(c, c->cur, c->out_l0_any, c->NL));   // |v| table: outlet velocities moved
        CUDA_OK(cudaStreamSynchronize(c->stream));
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    for (int it = 0; it < steps; ++it) PD_TRY(ard_body_step(c));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// `steps` x { NS loop body ; ARD loop body } without a host synchronisation in between: what a device-resident
// driver that alternates the two solvers per step issues (the unit of the throughput metric). Same state as
// pdgpu_ns_iterate(1) + pdgpu_ard_iterate(1) repeated.
int pd_ns_body_step(pdgpu_ctx* c);   // ns.cu
extern "C" int pdgpu_step_iterate(pdgpu_ctx* c, int steps, double dt_ns, double dt_ard) {
    NEED_FIELDS(c);
    if (steps <= 0) return 0;
    PD_TRY(pd_set_dt(c, 0, dt_ns));
    PD_TRY(pd_set_dt(c, 1, dt_ard));
    const bool fused2d = pd_ns2d_ok(c) && c->opt_lazy_wallc;
    for (int it = 0; it < steps; ++it) {
        PD_TRY(pd_ns_body_step(c));
        PD_TRY(pd_ensure_vmag(c, c->cur));               // |v| table of the new flow state
        if (fused2d) {
            PD_TRY(pd_enqueue_ard2d(c, c->cur, c->curC, 1));
            c->curC = 1 - c->curC;
            c->wallC_pending = true; c->wallC_src = 1 - c->curC;
            if (c->n_outlet) PD_TRY(pd_enqueue_ard_vmag_range(c, c->cur, c->out_l0_any, c->NL));
        } else {
            PD_TRY(ard_body_step(c));
        }
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// PD_ARD_Solver::apply_phase_change (src/pd_ard.cpp:193-212)
template <int DIM>
__global__ void k_phase_change(const int* __restrict__ l_solid, long long n_solid, uint8_t* __restrict__ type,
                               uint8_t* __restrict__ phase, double* __restrict__ C, double* __restrict__ rho,
                               double* __restrict__ p, double* __restrict__ vx, double* __restrict__ vy,
                               double* __restrict__ vz, double C_thresh, double rho_f, uint8_t* __restrict__ salt,
                               double* __restrict__ dsol, int* __restrict__ count, int* __restrict__ out,
                               double* __restrict__ out_rho_old, long long cap) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_solid) return;
    int l = l_solid[t];
    if (phase[l] == 0 && type[l] == PDGPU_SOLID_MG && C[l] < C_thresh) {
        const double rho_before = rho[l];   // the reference's `pressure` member keeps EOS(rho_before) until the next NS step
        phase[l] = 1;
        type[l] = PDGPU_FLUID;
        rho[l] = rho_f;
        p[l] = 0.0;
        vx[l] = 0.0; vy[l] = 0.0;
        if (DIM == 3) vz[l] = 0.0;
        C[l] = C_thresh;
        salt[l] = 0;
        dsol[l] = 0.0;
        const double rho_old = rho_before;
        int pos = atomicAdd(count, 1);
        if (pos < cap) { out[pos] = l; out_rho_old[pos] = rho_old; }
    }
}

extern "C" int pdgpu_phase_change(pdgpu_ctx* c, int* n_dissolved, int* dissolved_global, int cap) {
    NEED_FIELDS(c);
    PD_TRY(pd_flush_wall_c(c));   // node types are about to change
    if (!n_dissolved) PD_FAIL("pdgpu_phase_change: null output");
    *n_dissolved = 0;
    int n = 0;
    std::vector<int> h;
    if (c->n_solid > 0) {
        if (c->dissolved_cap < c->n_solid) {
            if (c->d_dissolved) CUDA_OK(cudaFree(c->d_dissolved));
            if (c->d_dissolved_rho) CUDA_OK(cudaFree(c->d_dissolved_rho));
            CUDA_OK(cudaMalloc(&c->d_dissolved, sizeof(int) * c->n_solid));
            CUDA_OK(cudaMalloc(&c->d_dissolved_rho, sizeof(double) * c->n_solid));
            c->dissolved_cap = c->n_solid;
        }
        CUDA_OK(cudaMemsetAsync(c->d_int, 0, sizeof(int), c->stream));
        int b = c->cur, bc = c->curC;
        if (c->dim == 2)
            LAUNCH(c, k_phase_change<2>, nblocks(c->n_solid, 256), 256, 0, c->l_solid, c->n_solid, c->type, c->phase,
                   c->C[bc], c->rho[b], c->p[b], VXYZ(c, b), c->cfg.C_thresh, c->cfg.rho_f, c->salt, c->dsol, c->d_int, c->d_dissolved,
                   c->d_dissolved_rho, c->dissolved_cap);
        else
            LAUNCH(c, k_phase_change<3>, nblocks(c->n_solid, 256), 256, 0, c->l_solid, c->n_solid, c->type, c->phase,
                   c->C[bc], c->rho[b], c->p[b], VXYZ(c, b), c->cfg.C_thresh, c->cfg.rho_f, c->salt, c->dsol, c->d_int, c->d_dissolved,
                   c->d_dissolved_rho, c->dissolved_cap);
        CUDA_OK(cudaMemcpyAsync(&n, c->d_int, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (n > 0) {
            h.resize(n);
            CUDA_OK(cudaMemcpy(h.data(), c->d_dissolved, sizeof(int) * n, cudaMemcpyDeviceToHost));
            // stale `pressure` of the dissolved nodes for snapshots written before the next NS step (vti.cu)
            std::vector<double> r(n);
            CUDA_OK(cudaMemcpy(r.data(), c->d_dissolved_rho, sizeof(double) * n, cudaMemcpyDeviceToHost));
            for (int t = 0; t < n; ++t) { c->stale_p_idx.push_back(h[t]); c->stale_p_rho.push_back(r[t]); }
            std::sort(h.begin(), h.end());
        }
    }
    // a rank must also learn about dissolved nodes inside its ghost planes
    int n_any = n;
    if (c->nranks > 1 && c->comm) {
        c->h_red[0] = (double)n;
        CUDA_OK(cudaMemcpyAsync(c->d_red, c->h_red, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        PD_TRY(pd_comm_allreduce(c, c->d_red, 1, 0));
        CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        n_any = (int)c->h_red[0];
        if (n_any > 0) PD_TRY(pd_enqueue_halo(c, 2, c->cur, c->curC));   // types, phase, rho, vel, p, C
    }
    if (n_any > 0) { pd_touch_flow(c); PD_TRY(pd_rebuild_tables(c)); }   // node lists, wall-mirror fallback, bond counts, graphs
    *n_dissolved = n;
    long long halo_shift = (long long)(c->a0 - c->R) * c->P;
    if (dissolved_global)
        for (int t = 0; t < n && t < cap; ++t) dissolved_global[t] = (int)(h[t] + halo_shift);
    return 0;
}

// number of SOLID_MG nodes whose concentration has fallen below C_thresh (the implicit coupling cycle ends
// at the first one, src/coupling.cpp:206-211)
__global__ void k_count_below(const int* __restrict__ l_solid, long long n, const uint8_t* __restrict__ type,
                              const double* __restrict__ C, double C_thresh, int* __restrict__ count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int l = l_solid[t];
    if (type[l] == PDGPU_SOLID_MG && C[l] < C_thresh) atomicAdd(count, 1);
}

extern "C" int pdgpu_solid_below_thresh(pdgpu_ctx* c, int* count) {
    NEED_FIELDS(c);
    if (!count) PD_FAIL("pdgpu_solid_below_thresh: null output");
    int n = 0;
    if (c->n_solid > 0) {
        CUDA_OK(cudaMemsetAsync(c->d_int, 0, sizeof(int), c->stream));
        LAUNCH(c, k_count_below, nblocks(c->n_solid, 256), 256, 0, c->l_solid, c->n_solid, c->type, c->C[c->curC],
               c->cfg.C_thresh, c->d_int);
        CUDA_OK(cudaMemcpyAsync(&n, c->d_int, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    if (c->nranks > 1 && c->comm) {
        double v = (double)n;
        PD_TRY(pdgpu_comm_allreduce(c, &v, 1, 0));
        n = (int)v;
    }
    *count = n;
    return 0;
}

// Reductions of write_diagnostics (src/coupling.cpp:20-49): solid count, max |v| and max C
// over FLUID nodes. All order-free (integer count, max of non-negative doubles).
template <int DIM>
__global__ void k_diag(long long own_lo, long long own_n, const uint8_t* __restrict__ type,
                       const double* __restrict__ C, const double* __restrict__ vx, const double* __restrict__ vy,
                       const double* __restrict__ vz, unsigned long long* __restrict__ out) {
    double vm = 0.0, cm = 0.0;
    unsigned long long ns = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < own_n;
         t += (long long)gridDim.x * blockDim.x) {
        long long l = own_lo + t;
        uint8_t ty = type[l];
        if (ty == PDGPU_SOLID_MG) ++ns;
        if (ty == PDGPU_FLUID) {
            double s = vx[l] * vx[l] + vy[l] * vy[l];
            if (DIM == 3) s += vz[l] * vz[l];
            vm = fmax(vm, sqrt(s));
            if (C[l] > cm) cm = C[l];
        }
    }
    vm = warp_max(vm);
    cm = warp_max(cm);
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if ((threadIdx.x & 31) == 0) {
        if (ns) atomicAdd(&out[0], ns);
        if (vm > 0.0) atomicMax(&out[1], (unsigned long long)__double_as_longlong(vm));
        if (cm > 0.0) atomicMax(&out[2], (unsigned long long)__double_as_longlong(cm));
    }
}

extern "C" int pdgpu_diag(pdgpu_ctx* c, PdDiag* out) {
    NEED_FIELDS(c);
    if (!out) PD_FAIL("pdgpu_diag: null output");
    long long own_n = c->own_hi - c->own_lo;
    CUDA_OK(cudaMemsetAsync(c->d_u64, 0, sizeof(unsigned long long) * 4, c->stream));
    unsigned g = std::min<unsigned>(nblocks(own_n, 256), 148 * 8);
    if (c->dim == 2)
        LAUNCH(c, k_diag<2>, g, 256, 0, c->own_lo, own_n, c->type, c->C[c->curC], VXYZ(c, c->cur), c->d_u64);
    else
        LAUNCH(c, k_diag<3>, g, 256, 0, c->own_lo, own_n, c->type, c->C[c->curC], VXYZ(c, c->cur), c->d_u64);
    unsigned long long h[4];
    CUDA_OK(cudaMemcpyAsync(h, c->d_u64, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    out->solid_count = (long long)h[0];
    memcpy(&out->v_max, &h[1], 8);
    memcpy(&out->C_max_fluid, &h[2], 8);
    if (c->nranks > 1 && c->comm) {
        c->h_red[0] = (double)out->solid_count; c->h_red[1] = out->v_max; c->h_red[2] = out->C_max_fluid;
        CUDA_OK(cudaMemcpyAsync(c->d_red, c->h_red, sizeof(double) * 3, cudaMemcpyHostToDevice, c->stream));
        PD_TRY(pd_comm_allreduce(c, c->d_red, 1, 0));
        PD_TRY(pd_comm_allreduce(c, c->d_red + 1, 2, 1));
        CUDA_OK(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double) * 3, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        out->solid_count = (long long)c->h_red[0]; out->v_max = c->h_red[1]; out->C_max_fluid = c->h_red[2];
    }
    return 0;
}

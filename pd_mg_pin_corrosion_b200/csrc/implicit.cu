// implicit.cu -- the implicit ARD branch: PD_ARD_ImplicitSolver of the reference
// (src/pd_ard_implicit.cpp) as a MATRIX-FREE operator on the device plus a restarted GMRES.
//
// The reference assembles the bond operator M once per coupling cycle into an Eigen sparse matrix
// (assemble, :104-346) and solves (I - dt M) C_new = C_old + dt bc_rhs per step with Eigen's
// GMRES + IncompleteLUT (step, :371-429).  Here nothing is assembled: every operator application
// walks the horizon-offset table and recomputes the bond weight
//     w_ij = beta D_avg V_j / xi^2                                   (diffusion; D_avg by bond class)
//     liquid-liquid:  w_adv = (alpha/V_H) (v_i . e) V_j / xi,  w_ij = (w_diff + max(0, w_adv - w_diff)) - w_adv
// from the node types, the velocity and the per-node interface diffusivity `dsol` (salt-layer
// blocking and volume-loss decay folded in by the pre-pass at "assemble" time, :70-89,127-132).
// Rows exist for FLUID and SOLID_MG nodes; INLET / OUTLET neighbours are known values that go to the
// right-hand side (:294-297), WALL / OUTSIDE and solid-solid bonds carry nothing (:196,212).
//
//     (M x)_i = sum_{j unknown} w_ij x_j - (sum_{j} w_ij) x_i          A = I - dt M
//     b_i     = C_i + dt sum_{j in INLET/OUTLET} w_ij C_j
//
// Linear solve: right-preconditioned GMRES(m), classical Gram-Schmidt applied twice with batched
// dot products (one host synchronisation per orthogonalisation pass).  Preconditioner: the operator
// is advection dominated along the tube axis (cell Peclet numbers >> 1), i.e. close to block lower
// triangular in the axial plane order, so P = (D + L_axial) -- diagonal plus the coupling to LOWER
// axial planes -- is applied by one forward sweep over the planes (one launch per plane).
//
// PARITY: the operator, the right-hand side and the adaptive step are checked against the line-by-line
// numpy restatement oracle/implicit_oracle.py; the solve is checked against the exact sparse solution.
// The reference's own solver (Eigen 3.4.0) is absent from its tree and from this image, so the
// solver itself is "unpinned" against the reference.  Single-GPU contexts only.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

struct ImplParams {
    double D_liquid, beta, div_coeff;
};

struct ImplState {
    bool assembled = false;
    double* diag = nullptr;      // 1 + dt sum_j w_ij of the last factor-free "setup" (per step)
    double* V = nullptr;         // Krylov basis, (m + 1) vectors of own_n
    double *w = nullptr, *z = nullptr, *x = nullptr, *b = nullptr, *r = nullptr;
    double* red = nullptr;       // reduction scratch
    double* h_red = nullptr;     // pinned
    int m_alloc = 0;
    long long n_alloc = 0;
};

constexpr int kRedBlocks = 296;
constexpr int kMaxM = 64;

// bond weight of (i -> neighbour through offset e); returns false when the bond carries nothing
template <int DIM>
__device__ __forceinline__ bool bond_weight(const ImplParams& P, const OffEntry& e, bool i_fluid, uint8_t tj,
                                            double dsol_i, double dsol_j, double vi0, double vi1, double vi2,
                                            double* w_out, bool* j_unknown) {
    if (tj == PDGPU_WALL) return false;                                                   // :196
    const bool j_fluid = (tj == PDGPU_FLUID || tj == PDGPU_INLET || tj == PDGPU_OUTLET);
    if (!i_fluid && !j_fluid) return false;                                               // solid-solid :212
    const double inv_xi = 1.0 / e.dist, inv_xi2 = inv_xi * inv_xi;
    double D_avg;
    if (i_fluid && j_fluid) D_avg = P.D_liquid;                                           // :215-217
    else D_avg = i_fluid ? dsol_j : dsol_i;                                               // interface :218-243
    const double w_diff = P.beta * D_avg * inv_xi2 * e.vol;                               // :267
    double w = w_diff;
    if (i_fluid && j_fluid) {                                                             // :272-282
        double vde = vi0 * e.ex + vi1 * e.ey;
        if (DIM == 3) vde += vi2 * e.ez;
        const double w_adv = P.div_coeff * vde * inv_xi * e.vol;
        const double w_stab = fmax(0.0, w_adv - w_diff);
        w = (w_diff + w_stab) - w_adv;
    }
    *w_out = w;
    *j_unknown = (tj == PDGPU_FLUID || tj == PDGPU_SOLID_MG);
    return true;
}

// One pass over the rows of the owned unknowns.  x, C are local arrays; outputs are owned-range arrays
// (index l - own_lo), any of them may be null:
//   y  = x - dt (M x)            (x read at neighbours; non-unknown rows: 0)
//   b  = C + dt * sum_bc w C_j
//   dg = 1 + dt * sum_j w
//   mc = (M C)_i + sum_bc w C_j  (dC/dt of the semi-discrete system; adaptive step)
template <int DIM>
__global__ void __launch_bounds__(128)
k_impl_rows(Lat L, long long own_lo, long long own_n, ImplParams P, double dt, const uint8_t* __restrict__ type,
            const OffEntry* __restrict__ off, int n_off, const double* __restrict__ vx,
            const double* __restrict__ vy, const double* __restrict__ vz, const double* __restrict__ dsol,
            const double* __restrict__ x, const double* __restrict__ C, double* __restrict__ y,
            double* __restrict__ b, double* __restrict__ dg, double* __restrict__ mc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= own_n) return;
    const long long l = own_lo + t;
    const uint8_t ti = type[l];
    if (ti != PDGPU_FLUID && ti != PDGPU_SOLID_MG) {
        if (y) y[t] = 0.0;
        if (b) b[t] = 0.0;
        if (dg) dg[t] = 1.0;
        if (mc) mc[t] = 0.0;
        return;
    }
    const bool i_fluid = ti == PDGPU_FLUID;
    double vi0 = 0.0, vi1 = 0.0, vi2 = 0.0;
    if (i_fluid) { vi0 = vx[l]; vi1 = vy[l]; if (DIM == 3) vi2 = vz[l]; }
    const double ds_i = dsol[l];
    const int q = (int)(l % L.P);
    const int jj = (DIM == 3) ? q / L.Nx : 0;
    const int ii = q - jj * L.Nx;
    double acc_x = 0.0, acc_c = 0.0, bc = 0.0, sum_w = 0.0;
    for (int o = 0; o < n_off; ++o) {
        const OffEntry e = off[o];
        const long long nn = nbr_local(L, e, DIM, ii, jj, l, type);
        if (nn < 0) continue;
        double w;
        bool ju;
        if (!bond_weight<DIM>(P, e, i_fluid, type[nn], ds_i, dsol[nn], vi0, vi1, vi2, &w, &ju)) continue;
        sum_w += w;
        if (ju) {
            if (x) acc_x += w * x[nn];
            if (mc) acc_c += w * C[nn];
        } else {
            bc += w * C[nn];
        }
    }
    if (y) y[t] = x[l] - dt * (acc_x - sum_w * x[l]);
    if (b) b[t] = C[l] + dt * bc;
    if (dg) dg[t] = 1.0 + dt * sum_w;
    if (mc) mc[t] = (acc_c - sum_w * C[l]) + bc;
}

// forward sweep of the preconditioner over one axial plane:
//   z_i = (r_i + dt * sum_{j unknown in LOWER planes} w_ij z_j) / diag_i
// z is a LOCAL array (ghost planes zero), r / diag owned-range arrays.
template <int DIM>
__global__ void __launch_bounds__(128)
k_impl_plane_fwd(Lat L, long long own_lo, long long plane_lo, ImplParams P, double dt,
                 const uint8_t* __restrict__ type, const OffEntry* __restrict__ off, int n_off,
                 const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                 const double* __restrict__ dsol, const double* __restrict__ r, const double* __restrict__ dg,
                 double* __restrict__ z) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= L.P) return;
    const long long l = plane_lo + q;
    const uint8_t ti = type[l];
    if (ti != PDGPU_FLUID && ti != PDGPU_SOLID_MG) { z[l] = 0.0; return; }
    const bool i_fluid = ti == PDGPU_FLUID;
    double vi0 = 0.0, vi1 = 0.0, vi2 = 0.0;
    if (i_fluid) { vi0 = vx[l]; vi1 = vy[l]; if (DIM == 3) vi2 = vz[l]; }
    const double ds_i = dsol[l];
    const int jj = (DIM == 3) ? (int)(q / L.Nx) : 0;
    const int ii = (int)(q - (long long)jj * L.Nx);
    double acc = 0.0;
    for (int o = 0; o < n_off; ++o) {
        const OffEntry e = off[o];
        const int dax = (DIM == 3) ? e.dk : e.dj;
        if (dax >= 0) continue;
        const long long nn = nbr_local(L, e, DIM, ii, jj, l, type);
        if (nn < 0) continue;
        double w;
        bool ju;
        if (!bond_weight<DIM>(P, e, i_fluid, type[nn], ds_i, dsol[nn], vi0, vi1, vi2, &w, &ju)) continue;
        if (ju) acc += w * z[nn];
    }
    const long long t = l - own_lo;
    z[l] = (r[t] + dt * acc) / dg[t];
}

// ---- small vector kernels over the owned range ------------------------------------------------
__global__ void k_scatter_local(const double* __restrict__ v, long long own_lo, long long n, double* __restrict__ loc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) loc[own_lo + t] = v[t];
}
__global__ void k_gather_local(const double* __restrict__ loc, long long own_lo, long long n, double* __restrict__ v) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) v[t] = loc[own_lo + t];
}
__global__ void k_div(const double* __restrict__ a, const double* __restrict__ d, long long n, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = a[t] / d[t];
}

__global__ void k_add_local(const double* __restrict__ z_loc, long long own_lo, long long n, double* __restrict__ x_loc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x_loc[own_lo + t] += z_loc[own_lo + t];
}
__global__ void k_scale_to(const double* __restrict__ a, double s, long long n, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = a[t] * s;
}
__global__ void k_sub(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = a[t] - b[t];
}
// w -= sum_i h[i] V_i
__global__ void k_orth_update(double* __restrict__ w, const double* __restrict__ V, long long n, int nvec,
                              const double* __restrict__ h) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double s = w[t];
    for (int i = 0; i < nvec; ++i) s -= h[i] * V[(long long)i * n + t];
    w[t] = s;
}
// out = sum_i y[i] V_i
__global__ void k_combine(const double* __restrict__ V, long long n, int nvec, const double* __restrict__ y,
                          double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double s = 0.0;
    for (int i = 0; i < nvec; ++i) s += y[i] * V[(long long)i * n + t];
    out[t] = s;
}
// partial[b][i] = <V_i, w> over block b's share (i < nvec), partial[b][nvec] = <w, w>; deterministic
__global__ void __launch_bounds__(256)
k_dots_partial(const double* __restrict__ V, const double* __restrict__ w, long long n, int nvec,
               double* __restrict__ partial) {
    __shared__ double sh[8];
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    for (int i = 0; i <= nvec; ++i) {
        const double* a = i < nvec ? V + (long long)i * n : w;
        double s = 0.0;
        for (long long t = lo + threadIdx.x; t < hi; t += blockDim.x) s += a[t] * w[t];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int k = 0; k < 8; ++k) tot += sh[k];
            partial[(long long)blockIdx.x * (kMaxM + 2) + i] = tot;
        }
        __syncthreads();
    }
}
__global__ void k_dots_final(const double* __restrict__ partial, int nb, int nvec, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i > nvec) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(long long)b * (kMaxM + 2) + i];
    out[i] = s;
}
// adaptive step (:452-480): min over SOLID_MG nodes with C > C_thresh and dC/dt < 0 of (C - C_thresh) / (-dC/dt)
__global__ void __launch_bounds__(256)
k_min_t_phase(long long own_lo, long long n, const uint8_t* __restrict__ type, const double* __restrict__ C,
              const double* __restrict__ mc, double C_thresh, unsigned long long* __restrict__ out) {
    double best = 1.7976931348623157e308;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long l = own_lo + t;
        if (type[l] != PDGPU_SOLID_MG) continue;
        const double Ci = C[l];
        if (Ci <= C_thresh) continue;
        const double d = mc[t];
        if (d >= 0.0) continue;
        const double rate = -d;
        if (rate < 1e-30) continue;
        const double tp = (Ci - C_thresh) / rate;
        if (tp > 0.0 && tp < best) best = tp;
    }
    best = warp_min(best);
    if ((threadIdx.x & 31) == 0) atomicMin(out, (unsigned long long)__double_as_longlong(best));   // positive doubles order like their bits
}
__global__ void k_clamp_store(const double* __restrict__ x, long long own_lo, long long n, const uint8_t* __restrict__ type,
                              double hi, double* __restrict__ C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long l = own_lo + t;
    const uint8_t ty = type[l];
    if (ty != PDGPU_FLUID && ty != PDGPU_SOLID_MG) return;
    double v = x[t];
    v = v < 0.0 ? 0.0 : v;
    C[l] = v > hi ? hi : v;                                                               // :417-425
}

ImplParams impl_params(const pdgpu_ctx* c) {
    PdConsts k = pd_consts(c->cfg, c->dim);
    ImplParams P;
    P.D_liquid = c->cfg.D_liquid;
    P.beta = k.beta_lap;
    P.div_coeff = k.alpha / k.V_H;
    return P;
}

ImplState* impl_state(pdgpu_ctx* c) {
    if (!c->impl_state) c->impl_state = new ImplState();
    return (ImplState*)c->impl_state;
}

int impl_reserve(pdgpu_ctx* c, int m) {
    ImplState* s = impl_state(c);
    const long long n = c->own_hi - c->own_lo;
    if (s->n_alloc == n && s->m_alloc >= m) return 0;
    double** bufs[] = {&s->diag, &s->V, &s->w, &s->z, &s->x, &s->b, &s->r, &s->red};
    for (double** p : bufs) { if (*p) cudaFree(*p); *p = nullptr; }
    CUDA_OK(cudaMalloc(&s->diag, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&s->V, sizeof(double) * n * (m + 1)));
    CUDA_OK(cudaMalloc(&s->w, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&s->z, sizeof(double) * c->NL));      // local array (neighbour reads)
    CUDA_OK(cudaMalloc(&s->x, sizeof(double) * c->NL));      // local array
    CUDA_OK(cudaMalloc(&s->b, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&s->r, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&s->red, sizeof(double) * ((size_t)kRedBlocks * (kMaxM + 2) + kMaxM + 2)));
    if (!s->h_red) CUDA_OK(cudaMallocHost(&s->h_red, sizeof(double) * (kMaxM + 2)));
    CUDA_OK(cudaMemsetAsync(s->z, 0, sizeof(double) * c->NL, c->stream));
    CUDA_OK(cudaMemsetAsync(s->x, 0, sizeof(double) * c->NL, c->stream));
    s->n_alloc = n;
    s->m_alloc = m;
    return 0;
}

// rows pass with any subset of outputs
int rows_pass(pdgpu_ctx* c, double dt, const double* x_local, double* y, double* b, double* dg, double* mc) {
    Lat L = make_lat(c);
    ImplParams P = impl_params(c);
    const long long n = c->own_hi - c->own_lo;
    const int fb = c->cur;
    if (c->dim == 2)
        LAUNCH(c, k_impl_rows<2>, nblocks(n, 128), 128, 0, L, c->own_lo, n, P, dt, c->type, c->d_off, c->n_off,
               VXYZ(c, fb), c->dsol, x_local, c->C[c->curC], y, b, dg, mc);
    else
        LAUNCH(c, k_impl_rows<3>, nblocks(n, 128), 128, 0, L, c->own_lo, n, P, dt, c->type, c->d_off, c->n_off,
               VXYZ(c, fb), c->dsol, x_local, c->C[c->curC], y, b, dg, mc);
    return 0;
}

// z_local = P^{-1} r   (precond 0: identity, 1: Jacobi, 2: forward axial sweep)
int apply_precond(pdgpu_ctx* c, ImplState* s, int precond, double dt, const double* r, double* z_local) {
    const long long n = c->own_hi - c->own_lo;
    if (precond == 2) {
        Lat L = make_lat(c);
        ImplParams P = impl_params(c);
        const int fb = c->cur;
        for (int a = c->a0; a < c->a1; ++a) {
            const long long plane_lo = (long long)(a - c->a0 + c->R) * c->P;
            if (c->dim == 2)
                LAUNCH(c, k_impl_plane_fwd<2>, nblocks(c->P, 128), 128, 0, L, c->own_lo, plane_lo, P, dt, c->type,
                       c->d_off, c->n_off, VXYZ(c, fb), c->dsol, r, s->diag, z_local);
            else
                LAUNCH(c, k_impl_plane_fwd<3>, nblocks(c->P, 128), 128, 0, L, c->own_lo, plane_lo, P, dt, c->type,
                       c->d_off, c->n_off, VXYZ(c, fb), c->dsol, r, s->diag, z_local);
        }
        return 0;
    }
    (void)n;
    return -1;   // identity / Jacobi are applied by the caller (k_div + k_scatter_local)
}

// batched dots: h_red[0..nvec] = { <V_i, w> }, <w, w>
int dots(pdgpu_ctx* c, ImplState* s, int nvec, const double* w) {
    const long long n = c->own_hi - c->own_lo;
    double* partial = s->red;
    double* fin = s->red + (size_t)kRedBlocks * (kMaxM + 2);
    LAUNCH(c, k_dots_partial, kRedBlocks, 256, 0, s->V, w, n, nvec, partial);
    LAUNCH(c, k_dots_final, 1, kMaxM + 2, 0, partial, kRedBlocks, nvec, fin);
    CUDA_OK(cudaMemcpyAsync(s->h_red, fin, sizeof(double) * (nvec + 1), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace

void pd_implicit_free(pdgpu_ctx* c) {
    ImplState* s = (ImplState*)c->impl_state;
    if (!s) return;
    double* bufs[] = {s->diag, s->V, s->w, s->z, s->x, s->b, s->r, s->red};
    for (double* p : bufs) if (p) cudaFree(p);
    if (s->h_red) cudaFreeHost(s->h_red);
    delete s;
    c->impl_state = nullptr;
}

// "assemble" (src/pd_ard_implicit.cpp:104-346): the salt-layer flags and interface diffusivities of the
// solid nodes from the CURRENT concentration (they stay frozen for the coupling cycle); nothing else
// is stored -- the operator is applied matrix-free.
extern "C" int pdgpu_implicit_assemble(pdgpu_ctx* c) {
    NEED_FIELDS(c);
    if (c->nranks > 1) PD_FAIL("pdgpu_implicit_*: single-GPU contexts only");
    PD_TRY(pd_flush_wall_c(c));
    CUDA_OK(cudaMemsetAsync(c->dsol, 0, sizeof(double) * c->NL, c->stream));
    PD_TRY(pd_enqueue_ard_prepass_solids(c, c->curC));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    impl_state(c)->assembled = true;
    return 0;
}

#define NEED_ASSEMBLED(c)                                                                              \
    do {                                                                                               \
        NEED_FIELDS(c);                                                                                \
        if (!c->impl_state || !((ImplState*)c->impl_state)->assembled)                                 \
            PD_FAIL("pdgpu_implicit_assemble has not been called for this coupling cycle");           \
    } while (0)

// y = (I - dt M) x for a GLOBAL host vector x (entries of non-unknown nodes are ignored, y there = 0)
extern "C" int pdgpu_implicit_matvec(pdgpu_ctx* c, double dt, const double* x_global, double* y_global) {
    NEED_ASSEMBLED(c);
    if (!x_global || !y_global) PD_FAIL("pdgpu_implicit_matvec: null array");
    PD_TRY(impl_reserve(c, 1));
    ImplState* s = impl_state(c);
    const long long n = c->own_hi - c->own_lo, goff = (long long)c->a0 * c->P;
    CUDA_OK(cudaMemcpyAsync(s->x + c->own_lo, x_global + goff, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    PD_TRY(rows_pass(c, dt, s->x, s->w, nullptr, nullptr, nullptr));
    CUDA_OK(cudaMemcpyAsync(y_global + goff, s->w, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// b = C_old + dt bc_rhs (global host array; 0 at non-unknown nodes)
extern "C" int pdgpu_implicit_rhs(pdgpu_ctx* c, double dt, double* b_global) {
    NEED_ASSEMBLED(c);
    if (!b_global) PD_FAIL("pdgpu_implicit_rhs: null array");
    PD_TRY(impl_reserve(c, 1));
    ImplState* s = impl_state(c);
    const long long n = c->own_hi - c->own_lo, goff = (long long)c->a0 * c->P;
    PD_TRY(rows_pass(c, dt, nullptr, nullptr, s->b, nullptr, nullptr));
    CUDA_OK(cudaMemcpyAsync(b_global + goff, s->b, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// compute_adaptive_dt (src/pd_ard_implicit.cpp:438-487)
extern "C" int pdgpu_implicit_compute_dt(pdgpu_ctx* c, double dt_fraction, double dt_max, double* dt_out) {
    NEED_ASSEMBLED(c);
    if (!dt_out) PD_FAIL("pdgpu_implicit_compute_dt: null output");
    PD_TRY(impl_reserve(c, 1));
    ImplState* s = impl_state(c);
    const long long n = c->own_hi - c->own_lo;
    PD_TRY(rows_pass(c, 0.0, nullptr, nullptr, nullptr, nullptr, s->r));
    const double big = 1.7976931348623157e308;
    unsigned long long init;
    memcpy(&init, &big, 8);
    CUDA_OK(cudaMemcpyAsync(c->d_u64, &init, 8, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_min_t_phase, std::min<unsigned>(nblocks(n, 256), 148 * 8), 256, 0, c->own_lo, n, c->type, c->C[c->curC],
           s->r, c->cfg.C_thresh, c->d_u64);
    double tmin = 0.0;
    CUDA_OK(cudaMemcpyAsync(&tmin, c->d_u64, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double min_t = std::min(dt_max, tmin);
    double dt = dt_fraction * min_t;
    dt = std::min(dt, dt_max);
    dt = std::max(dt, dt_max * 0.01);
    *dt_out = dt;
    return 0;
}

// PD_ARD_ImplicitSolver::step (:371-429): solve (I - dt M) C_new = C_old + dt bc_rhs, clamp to
// [0, C_solid_init], store into the current C buffer.  GMRES(restart) up to max_iters iterations or
// relative residual tol (the reference: 1e-10, restart 50, 200 iterations).
extern "C" int pdgpu_implicit_step(pdgpu_ctx* c, double dt, double tol, int restart, int max_iters, int precond,
                                   PdLinSolveInfo* info) {
    NEED_ASSEMBLED(c);
    if (restart < 1 || restart > kMaxM) PD_FAIL("pdgpu_implicit_step: restart must be in [1, %d]", kMaxM);
    if (precond < 0 || precond > 2) PD_FAIL("pdgpu_implicit_step: precond 0 (none), 1 (Jacobi) or 2 (axial sweep)");
    PD_TRY(pd_flush_wall_c(c));
    PD_TRY(impl_reserve(c, restart));
    ImplState* s = impl_state(c);
    const int m = restart;
    const long long n = c->own_hi - c->own_lo;
    const unsigned nb = nblocks(n, 256);
    double* Cbuf = c->C[c->curC];
    // b, diagonal; x0 = C_old
    PD_TRY(rows_pass(c, dt, nullptr, nullptr, s->b, s->diag, nullptr));
    CUDA_OK(cudaMemcpyAsync(s->x, Cbuf, sizeof(double) * c->NL, cudaMemcpyDeviceToDevice, c->stream));
    PD_TRY(dots(c, s, 0, s->b));
    const double bnorm = std::sqrt(s->h_red[0]);
    const double target = tol * (bnorm > 0.0 ? bnorm : 1.0);
    int iters = 0, converged = 0;
    double res = 0.0;
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), yv(m);
    double* d_y = s->red + (size_t)kRedBlocks * (kMaxM + 2);   // reuse the final-reduction slot for small vectors
    auto precond_apply = [&](const double* r_own, double* z_local) -> int {
        if (precond == 2) return apply_precond(c, s, 2, dt, r_own, z_local);
        if (precond == 1) LAUNCH(c, k_div, nb, 256, 0, r_own, s->diag, n, s->w);
        LAUNCH(c, k_scatter_local, nb, 256, 0, precond == 1 ? s->w : r_own, c->own_lo, n, z_local);
        return 0;
    };
    while (iters < max_iters) {
        // r = b - A x
        PD_TRY(rows_pass(c, dt, s->x, s->w, nullptr, nullptr, nullptr));
        LAUNCH(c, k_sub, nb, 256, 0, s->b, s->w, n, s->r);
        PD_TRY(dots(c, s, 0, s->r));
        double beta = std::sqrt(s->h_red[0]);
        res = beta;
        if (beta <= target) { converged = 1; break; }
        LAUNCH(c, k_scale_to, nb, 256, 0, s->r, 1.0 / beta, n, s->V);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int j = 0;
        for (; j < m && iters < max_iters; ++j, ++iters) {
            // w = A P^{-1} V_j
            PD_TRY(precond_apply(s->V + (long long)j * n, s->z));
            PD_TRY(rows_pass(c, dt, s->z, s->w, nullptr, nullptr, nullptr));
            // classical Gram-Schmidt, twice
            for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                PD_TRY(dots(c, s, j + 1, s->w));
                for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] += s->h_red[i];
                CUDA_OK(cudaMemcpyAsync(d_y, s->h_red, sizeof(double) * (j + 1), cudaMemcpyHostToDevice, c->stream));
                LAUNCH(c, k_orth_update, nb, 256, 0, s->w, s->V, n, j + 1, d_y);
                CUDA_OK(cudaStreamSynchronize(c->stream));   // h_red is reused by the next dots()
            }
            PD_TRY(dots(c, s, 0, s->w));
            const double hn = std::sqrt(s->h_red[0]);
            H[(size_t)(j + 1) * m + j] = hn;
            if (hn > 0.0) LAUNCH(c, k_scale_to, nb, 256, 0, s->w, 1.0 / hn, n, s->V + (long long)(j + 1) * n);
            // Givens rotations on column j
            for (int i = 0; i < j; ++i) {
                const double a = H[(size_t)i * m + j], b2 = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = cs[i] * a + sn[i] * b2;
                H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * b2;
            }
            const double a = H[(size_t)j * m + j], b2 = H[(size_t)(j + 1) * m + j];
            const double rr = std::hypot(a, b2);
            cs[j] = rr > 0.0 ? a / rr : 1.0;
            sn[j] = rr > 0.0 ? b2 / rr : 0.0;
            H[(size_t)j * m + j] = rr;
            H[(size_t)(j + 1) * m + j] = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            res = std::fabs(g[j + 1]);
            if (res <= target || hn == 0.0) { ++j; ++iters; break; }
        }
        // y = H^{-1} g ;  x += P^{-1} (V y)
        const int k = j;
        for (int i = k - 1; i >= 0; --i) {
            double sum = g[i];
            for (int l2 = i + 1; l2 < k; ++l2) sum -= H[(size_t)i * m + l2] * yv[l2];
            yv[i] = sum / H[(size_t)i * m + i];
        }
        CUDA_OK(cudaMemcpyAsync(d_y, yv.data(), sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, k_combine, nb, 256, 0, s->V, n, k, d_y, s->r);
        CUDA_OK(cudaStreamSynchronize(c->stream));   // yv may be reused
        PD_TRY(precond_apply(s->r, s->z));
        LAUNCH(c, k_add_local, nb, 256, 0, s->z, c->own_lo, n, s->x);   // x += P^{-1} (V y)
    }
    // true residual of the returned iterate
    PD_TRY(rows_pass(c, dt, s->x, s->w, nullptr, nullptr, nullptr));
    LAUNCH(c, k_sub, nb, 256, 0, s->b, s->w, n, s->r);
    PD_TRY(dots(c, s, 0, s->r));
    res = std::sqrt(s->h_red[0]);
    if (res <= target) converged = 1;
    LAUNCH(c, k_gather_local, nb, 256, 0, s->x, c->own_lo, n, s->w);
    LAUNCH(c, k_clamp_store, nb, 256, 0, s->w, c->own_lo, n, c->type, c->cfg.C_solid_init, Cbuf);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (info) { info->iters = iters; info->converged = converged; info->rel_res = bnorm > 0.0 ? res / bnorm : res; info->pad = 0; }
    return 0;
}
